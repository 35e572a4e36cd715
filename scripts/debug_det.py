import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gfx_ocean_b200 import Ocean
n, tiles = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "gen"
from gfx_ocean_b200.spectrum import synthetic_tile
with Ocean(n, 1000.0, n_tiles=tiles) as o:
    for i in range(tiles):
        if mode == "gen":
            o.generate_spectrum(i, 1234, stream_id=i)
        else:
            h0, w = synthetic_tile(n, i)
            o.set_spectrum(i, h0, w)
    rows = []
    for rep in range(8):
        if mode == "gen" and rep % 2 == 1:
            for i in range(tiles):
                o.generate_spectrum(i, 1234, stream_id=i)
        o.update(3.25)
        rows.append(o.output_checksums())
    rows = np.array(rows)
    ref = rows[0]
    for r in rows:
        print(" ".join("." if a == b else "X" for a, b in zip(r, ref)))
    # which texels differ for a mismatching tile
    o.update(3.25)
    base = [o.read_back(i) for i in range(tiles)]
    for rep in range(6):
        o.update(3.25)
        for i in range(tiles):
            cur = o.read_back(i)
            d = np.abs(cur - base[i])
            if d.max() > 0:
                ys, xs = np.nonzero(d[..., :3].max(axis=-1))
                ch = [float(d[..., c].max()) for c in range(3)]
                print(f"rep {rep} tile {i}: {len(ys)} texels differ, channels max {ch}, x range {xs.min()}..{xs.max()} unique x {len(np.unique(xs))}, y range {ys.min()}..{ys.max()} unique y {len(np.unique(ys))}")
