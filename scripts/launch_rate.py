"""Host enqueue cost vs device time per frame (one tile per update): Python loop vs the library's C loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gfx_ocean_b200 import Ocean
from gfx_ocean_b200.spectrum import synthetic_tile

for n in (64, 512, 1024):
    h0, w = synthetic_tile(n, 0)
    with Ocean.new(n, 1000.0, w, h0) as o:
        o.update_sequence(0.0, 0.016, 200); o.sync()
        K = 4000
        t0 = time.perf_counter()
        for i in range(K):
            o.update(0.016 * i)
        t_enq = time.perf_counter() - t0
        o.sync()
        t_py = time.perf_counter() - t0
        t0 = time.perf_counter()
        o.update_sequence(0.0, 0.016, K)
        t_enq_c = time.perf_counter() - t0
        o.sync()
        t_c = time.perf_counter() - t0
        print(f"N={n}: python loop enqueue {1e6*t_enq/K:.2f} us/frame, total {1e6*t_py/K:.2f} us/frame | "
              f"C loop enqueue {1e6*t_enq_c/K:.2f} us/frame, total {1e6*t_c/K:.2f} us/frame")
