"""Small driver for ncu captures: a few frames of the fused path on one context.
    ncu ... python scripts/prof_target.py [N] [tiles] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfx_ocean_b200 import Ocean, PIPELINE_FUSED, PIPELINE_LITERAL  # noqa: E402
from gfx_ocean_b200.spectrum import synthetic_tile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 6
pipeline = PIPELINE_LITERAL if len(sys.argv) > 4 and sys.argv[4] == "literal" else PIPELINE_FUSED
with Ocean(n, n_tiles=tiles, pipeline=pipeline) as o:
    h0, w = synthetic_tile(n, 0)
    for t in range(tiles):
        o.set_spectrum(t, h0, w)
    for i in range(frames):
        o.update(0.016 * i)
    o.sync()
    print("frames", frames, "launches", o.launch_count)
