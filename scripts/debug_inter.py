import os, sys
os.environ["OCEAN_B200_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gfx_ocean_b200 import Ocean
n, tiles = 1024, 8
def sums(o, what):
    if what: os.environ["OCEAN_B200_DEBUG_INTER"] = what
    r = o.output_checksums()
    os.environ.pop("OCEAN_B200_DEBUG_INTER", None)
    return r
with Ocean(n, 1000.0, n_tiles=tiles) as o:
    for i in range(tiles):
        o.generate_spectrum(i, 1234, stream_id=i)
    rows = []
    for rep in range(8):
        o.update(3.25)
        rows.append((sums(o, "h"), sums(o, "p"), sums(o, None)))
    for name, k in (("GH (first 512x512 float4 of each tile)", 0), ("GP", 1), ("output", 2)):
        print(name)
        for r in rows:
            print(" ".join("." if a == b else "X" for a, b in zip(r[k], rows[0][k])))
