"""Small multi-tile run for compute-sanitizer: python scripts/san_target.py N tiles frames"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gfx_ocean_b200 import Ocean
n, tiles, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lanes = len(sys.argv) > 4 and sys.argv[4] == "overlapped"      # frames through ocean_update_overlapped (two lanes)
with Ocean(n, 1000.0, n_tiles=tiles) as o:
    for i in range(tiles):
        o.generate_spectrum(i, 7, stream_id=i)
    for f in range(frames):
        if lanes:
            o.update_overlapped(0.1 * f)                       # all tiles: the column kernel is ordered behind the other lane
            o.update_overlapped(0.1 * f, f % tiles, 1)         # one tile: the lanes run side by side
        else:
            o.update(0.1 * f)
    o.sync()
    print("checksums", [hex(int(s)) for s in o.output_checksums()])
