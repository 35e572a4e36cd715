#!/usr/bin/env python
"""bench.py -- ocean frames/sec at N=1024 (3 fields) + achieved HBM GB/s vs the B200 roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1024x1024 grids,
three fields (height, dx, dz), `--tiles` independent oceans per GPU (seeded synthetic spectra,
SURVEY.md 8d). One *step* = one ocean_update() of the rank's context = one frame of every tile;
one *frame* = one tile's propagate -> 2-D inverse FFT x3 -> correction -> RGBA32F displacement map.
With 8 tiles a step touches 96 MB of inputs + 128 MB of outputs (+ intermediates), more than the
126 MB L2, so every step's inputs come from HBM ("inputs larger than L2").

Prints ONE JSON line (rank 0). `value` = tile-frames/s over all GPUs with inputs resident in HBM;
`e2e` = same metric through Ocean.update() + read_back into pinned host memory every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ocean frames/sec at N=1024 (3 fields); achieved HBM GB/s vs B200 peak"
UNIT = "frames/s"
ALG_BYTES_PER_POINT = 76          # pass A 12 + 24, pass B 24 + 16
ALG_BYTES_ROWS, ALG_BYTES_COLS = 36, 40


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--resolution", type=int, default=1024)
    ap.add_argument("--tiles", type=int, default=8, help="independent oceans per GPU")
    ap.add_argument("--pipeline", default="fused", choices=["fused", "literal"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dt", type=float, default=0.016)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_fps(n: int, tiles_data, seconds: float, dt: float):
    """The reference's algorithm on the host cores: literal fp32 restatement (oracle/, OpenMP)."""
    from oracle.ocean_oracle import COracle
    o = COracle()
    o.set_num_threads(len(os.sched_getaffinity(0)))      # torchrun exports OMP_NUM_THREADS=1: use every host thread
    cores = o.num_threads()
    o.frame(tiles_data[0][0], tiles_data[0][1], 0.0, n, prec="f32")      # warm-up
    frames, t0 = 0, time.perf_counter()
    while True:
        h0, w = tiles_data[frames % len(tiles_data)]
        o.frame(h0, w, dt * frames, n, prec="f32")
        frames += 1
        el = time.perf_counter() - t0
        if el >= seconds or frames >= 4096:
            break
    return frames / el, cores, frames, el


def run_reference(args):
    """--impl reference: the reference's own CPU-runnable implementation of the path. The Rust +
    GLSL reference cannot be built in this image (no cargo, no Vulkan), so this is the literal fp32
    C restatement of its four shaders (oracle/ocean_oracle.c) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gfx_ocean_b200.spectrum import synthetic_tile
    from oracle.ocean_oracle import COracle
    n, tiles = args.resolution, args.tiles
    data = [synthetic_tile(n, t) for t in range(tiles)]
    o = COracle()
    o.set_num_threads(len(os.sched_getaffinity(0)))      # torchrun exports OMP_NUM_THREADS=1: use every host thread
    cores = o.num_threads()
    step = 0
    for _ in range(max(args.warmup, 1)):
        for h0, w in data[:1]:
            o.frame(h0, w, 0.0, n, prec="f32")
    # bounded sample: a step = one frame of every tile of ONE GPU's share
    t0 = time.perf_counter()
    for step in range(args.steps):
        for h0, w in data:
            o.frame(h0, w, args.dt * step, n, prec="f32")
        if time.perf_counter() - t0 > 150.0:
            break
    el = time.perf_counter() - t0
    steps = step + 1
    fps = steps * tiles / el
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}x{n} x3 fields, {tiles} tiles per step, CPU literal fp32 restatement of the "
                               "reference shaders (oracle/ocean_oracle.c, OpenMP)", "resolution": n, "tiles_per_step": tiles},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {tiles} tile-frames at {n}x{n} in {el:.1f} s"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gfx_ocean_b200 import Ocean, PIPELINE_FUSED, PIPELINE_LITERAL
    from gfx_ocean_b200.spectrum import synthetic_tile

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ocean path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, tiles, K, W = args.resolution, args.tiles, args.steps, max(args.warmup, 3)

    # tile-parallel sharding: rank r owns global tiles [r*tiles, (r+1)*tiles); no data-path collective
    from gfx_ocean_b200.shard import tiles_of_rank
    my_tiles = tiles_of_rank(rank, world, tiles * world)
    stream = torch.cuda.Stream()
    pipeline = PIPELINE_FUSED if args.pipeline == "fused" else PIPELINE_LITERAL
    ocean = Ocean(n, 1000.0, n_tiles=len(my_tiles), device=local_rank, pipeline=pipeline, stream=stream.cuda_stream)
    data = []
    for i, g in enumerate(my_tiles):
        h0, w = synthetic_tile(n, g)
        ocean.set_spectrum(i, h0, w)
        if rank == 0 and i < 2:
            data.append((h0, w))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- device-resident throughput: K steps, inputs already in HBM
    with torch.cuda.stream(stream):
        for i in range(W):
            ocean.update(args.dt * i)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            for i in range(max(1, int(0.4 / 1.2e-4))):      # keep the GPU loaded while nvidia-smi spins up
                ocean.update(args.dt * i)
        l0 = ocean.launch_count
        barrier()
        ev0.record(stream)
        for i in range(K):
            ocean.update(args.dt * (W + i))
        ev1.record(stream)
        barrier()
        launches = ocean.launch_count - l0
        clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    fps = world * len(my_tiles) * K / (ms * 1e-3)

    # ---- latency mode: ONE tile per ocean_update, rotating over the context's tiles so each frame's inputs
    #      are L2-cold (8 x 12 MB inputs + 8 x 16 MB outputs > 126 MB L2); this is BASELINE.json configs[2]
    #      taken literally (a single 1024^2 ocean per frame)
    single = None
    if len(my_tiles) >= 8:
        ks = max(200, K // 4)
        with torch.cuda.stream(stream):
            for i in range(W):
                ocean.update_tiles(args.dt * i, i % len(my_tiles), 1)
            barrier()
            ev0.record(stream)
            for i in range(ks):
                ocean.update_tiles(args.dt * i, i % len(my_tiles), 1)
            ev1.record(stream)
            barrier()
        ms1 = max_over_ranks(ev0.elapsed_time(ev1))
        single = {"value": world * ks / (ms1 * 1e-3), "unit": UNIT, "us_per_frame": 1e3 * ms1 / ks, "steps": ks,
                  "alg_GBps": ALG_BYTES_PER_POINT * n * n / (ms1 / ks * 1e-3) / 1e9,
                  "note": "one tile per ocean_update (2 launches), rotating over the tiles so inputs are L2-cold"}

    # ---- library sanity bar: cuFFT's batched 2-D C2C inverse transform ALONE on the same
    #      amount of data (3 complex fields per tile, via torch.fft.ifft2), without propagate or correction.
    #      A comparison line only: nothing of it is on the product path.
    cufft_ms = None
    if rank == 0:
        try:
            spec = torch.randn(len(my_tiles) * 3, n, n, dtype=torch.complex64, device="cuda")
            with torch.cuda.stream(stream):
                for _ in range(3):
                    torch.fft.ifft2(spec, norm="forward")
                torch.cuda.synchronize()
                ev0.record(stream)
                for _ in range(20):
                    torch.fft.ifft2(spec, norm="forward")
                ev1.record(stream)
                torch.cuda.synchronize()
            cufft_ms = ev0.elapsed_time(ev1) / 20
            del spec
        except Exception as exc:      # comparison only
            cufft_ms = f"unavailable: {exc}"

    # ---- per-kernel durations (CUDA events between the two launches, same stream)
    stage_ms = None
    if args.pipeline == "fused":
        acc = np.zeros(2)
        reps = 20
        for i in range(reps):
            acc += np.array(ocean.profile_update(args.dt * i))
        stage_ms = (acc / reps).tolist()

    # ---- end to end through the public API with host buffers: update + read back every step
    nbytes_out = n * n * 16 * len(my_tiles)
    host = torch.empty((len(my_tiles), n, n, 4), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(10, min(K, 100))
    with torch.cuda.stream(stream):
        for i in range(3):
            ocean.update(args.dt * i)
            for t in range(len(my_tiles)):
                ocean.read_back_async(t, host[t].data_ptr())
        barrier()
        ev0.record(stream)
        for i in range(e2e_steps):
            ocean.update(args.dt * (i + 7))          # PropagateLocals (12 B) travel as kernel parameters
            for t in range(len(my_tiles)):
                ocean.read_back_async(t, host[t].data_ptr())
            ocean.sync()                             # the caller owns the host frame before the next one
        ev1.record(stream)
        barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_fps = world * len(my_tiles) * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_step = ALG_BYTES_PER_POINT * n * n * len(my_tiles)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n}x{n} x3 fields (height, dx, dz), {len(my_tiles)} independent tiles per GPU per step; "
                                   "frame = one tile's propagate -> 2-D iFFT -> correction -> RGBA32F map",
                       "resolution": n, "tiles_per_gpu": len(my_tiles), "pipeline": args.pipeline,
                       "l2": "inputs larger than L2 (per-step working set %.0f MB > 126 MB)" % (
                           (12 + 12 + 16) * n * n * len(my_tiles) / 1e6),
                       "parallelism": f"tile-parallel x{world}, no data-path collective"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": 12, "d2h_bytes_per_step": nbytes_out,
                    "steps": e2e_steps, "note": "Ocean.update(t) + read_back of every tile into pinned host memory + sync, per step"},
            "step_alg_GBps": alg_step / (ms / K * 1e-3) / 1e9,
            "single_tile_per_update": single,
            "cufft_ifft2_only_ms_per_step": cufft_ms,
        }
        if stage_ms:
            alg = [ALG_BYTES_ROWS * n * n * len(my_tiles), ALG_BYTES_COLS * n * n * len(my_tiles)]
            names = ["k_rows", "k_cols"]
            d = int(np.argmax(stage_ms))
            ach = alg[d] / (stage_ms[d] * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": names[d], "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": recorded_traffic(names[d]), "peak_source": peak_src,
                                "alg_bytes_per_launch": alg[d], "launch_ms": stage_ms[d]}
            line["kernels"] = [{"name": names[i], "launch_ms": stage_ms[i], "alg_bytes": alg[i],
                                "alg_GBps": alg[i] / (stage_ms[i] * 1e-3) / 1e9} for i in range(2)]
        if not args.no_cpu_baseline and world == 1:
            cfps, cores, frames, el = cpu_reference_fps(n, data, args.cpu_seconds, args.dt)
            line["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{frames} frames at {n}x{n} in {el:.1f} s (literal fp32 restatement, OpenMP)"}
        print(json.dumps(line), flush=True)
    ocean.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
