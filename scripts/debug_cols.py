import os, sys
os.environ["OCEAN_B200_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gfx_ocean_b200 import Ocean
n, tiles = 1024, 8
with Ocean(n, 1000.0, n_tiles=tiles) as o:
    for i in range(tiles):
        o.generate_spectrum(i, 1234, stream_id=i)
    print("full frames (k_rows + k_cols):")
    full = []
    for rep in range(6):
        o.update(3.25)
        full.append(o.output_checksums())
    for r in full:
        print(" ".join("." if a == b else "X" for a, b in zip(r, full[0])))
    print("k_cols only, same intermediate:")
    os.environ["OCEAN_B200_DEBUG_SKIP_ROWS"] = "1"
    cols = []
    for rep in range(10):
        o.update(3.25)
        cols.append(o.output_checksums())
    for r in cols:
        print(" ".join("." if a == b else "X" for a, b in zip(r, cols[0])))
    del os.environ["OCEAN_B200_DEBUG_SKIP_ROWS"]
