"""Small driver for an ncu capture of the consumer step: update + normal map on a dx-plane context.
    ncu ... -k regex:k_normal_map_plane python scripts/prof_normals.py [N] [tiles] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfx_ocean_b200 import FLAG_DX_PLANE, Ocean  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
with Ocean(n, 1000.0, n_tiles=tiles, flags=FLAG_DX_PLANE) as o:
    for t in range(tiles):
        o.generate_spectrum(t, 7, stream_id=t)
    for i in range(frames):
        o.update(0.016 * i)
        o.compute_normals()
    o.sync()
    print("frames", frames, "launches", o.launch_count)
