import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gfx_ocean_b200 import Ocean
from gfx_ocean_b200.spectrum import synthetic_tile
from oracle.ocean_oracle import COracle
n, tiles = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 3
o_ = COracle()
data = [synthetic_tile(n, g) for g in range(tiles)]
with Ocean(n, 1000.0, n_tiles=tiles) as o:
    for i, (h0, w) in enumerate(data):
        o.set_spectrum(i, h0, w)
    for rep in range(2):
        o.update(1.25)
        for i, (h0, w) in enumerate(data):
            out = o.read_back(i).astype(np.float64)
            ref = o_.frame(h0, w, 1.25, n, prec="f64")
            d = out[..., :3] - ref[..., :3]
            # spectral location of the error: which rows (ky) / columns (kx) of the 2-D spectrum are wrong
            s = np.fft.fft2(d[..., 1])
            e_row = np.abs(s).max(axis=1); e_col = np.abs(s).max(axis=0)
            print(f"rep {rep} tile {i}: max err {np.abs(d).max():.3e} / {np.abs(ref).max():.2f}; worst spectral rows {np.argsort(-e_row)[:6]} ({np.sort(e_row)[::-1][:3]}), cols {np.argsort(-e_col)[:6]}")
