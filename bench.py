#!/usr/bin/env python
"""bench.py -- ocean frames/sec at N=1024 (3 fields) + achieved HBM GB/s vs the B200 roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1024x1024 grids,
three fields (height, dx, dz), `--tiles` independent oceans per GPU (seeded synthetic spectra generated
ON the device by ocean_generate_spectrum). One *step* = one ocean_update() of the rank's
context = one frame of every tile; one *frame* = one tile's propagate -> 2-D inverse FFT x3 -> correction
-> RGBA32F displacement map. With 8 tiles a step touches 96 MB of inputs + 128 MB of outputs
(+ intermediates), more than the 126 MB L2, so every step's inputs come from HBM ("inputs larger than L2").

Headline call: ocean_update_overlapped (frames alternate between two internal lanes; same maps as ocean_update bit for
bit); `plain_updates` reports ocean_update beside it.

Timing: W warm-up steps, then R >= 5 repetitions of EXACTLY K steps, each bracketed by barrier +
synchronize and timed with CUDA events on the launching stream (max over ranks); R grows until the
repetitions add up to >= --min-time seconds, so a small K still gives a stable number. `value` is the
MEDIAN repetition; min/max are reported beside it.

Prints ONE JSON line (rank 0). `value` = tile-frames/s over all GPUs with inputs resident in HBM;
`e2e` = same metric through the public API with host buffers: Ocean.update() + read-back of every tile
into pinned host memory every step (double-buffered: frame n+1 is computed while frame n is copied).
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
ALG_BYTES_PER_POINT = 76          # pass A 12 + 24, pass B 24 + 16 (the contract denominator)
ALG_BYTES_ROWS, ALG_BYTES_COLS = 36, 40
REAL_BYTES_PER_POINT = 52         # what the Hermitian-folded kernels must move: 12 + 12 (rows), 12 + 16 (columns)
SEED = 1234                       # tile g of the benchmark = ocean_generate_spectrum(seed=SEED, stream_id=g)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--resolution", type=int, default=1024)
    ap.add_argument("--tiles", type=int, default=8, help="independent oceans per GPU (weak scaling)")
    ap.add_argument("--total-tiles", type=int, default=0,
                    help="strong scaling (BASELINE.json configs[4]): this many tiles split over the GPUs, e.g. 64")
    ap.add_argument("--pipeline", default="fused", choices=["fused", "literal"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other_configs / normals / graph / cuFFT side lines")
    ap.add_argument("--reps", type=int, default=5, help="minimum number of timed repetitions of --steps steps")
    ap.add_argument("--min-time", type=float, default=0.5, help="repeat until the timed repetitions add up to this many seconds")
    ap.add_argument("--dt", type=float, default=0.016)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json);
    ncu cannot run inside the bench, so this is the one number of the line that is not measured live."""
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


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank on the CPUs NVML reports as local to its GPU, so that pinned host buffers (first touch) and the
    copy-engine traffic stay on the GPU's NUMA node. Returns the affinity to restore for the CPU baseline leg."""
    before = os.sched_getaffinity(0)
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i for i in range(os.cpu_count()) if (mask[i // 64] >> (i % 64)) & 1} & before
        if cpus:
            os.sched_setaffinity(0, cpus)
        return before, sorted(cpus)
    except Exception:
        return before, None


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
    n = args.resolution
    tiles = args.total_tiles // max(args.gpus, 1) if args.total_tiles else args.tiles
    data = [synthetic_tile(n, t) for t in range(min(tiles, 8))]
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
        for i in range(tiles):
            h0, w = data[i % len(data)]
            o.frame(h0, w, args.dt * step, n, prec="f32")
        if time.perf_counter() - t0 > 150.0:
            break
    el = time.perf_counter() - t0
    steps = step + 1
    fps = steps * tiles / el
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
        "scaling": "strong" if args.total_tiles else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}x{n} x3 fields (height, dx, dz), {tiles} independent tiles per GPU per step; "
                               "frame = one tile's propagate -> 2-D iFFT -> correction -> RGBA32F map",
                   "resolution": n, "tiles_per_gpu": tiles, "total_tiles": tiles * max(args.gpus, 1),
                   "implementation": "CPU literal fp32 restatement of the reference shaders (oracle/ocean_oracle.c, OpenMP); "
                                     "rank 0 times one GPU's share of the tiles"},
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
    from gfx_ocean_b200 import FLAG_DOUBLE_BUFFER_OUTPUT, FLAG_DX_PLANE, Ocean, PIPELINE_FUSED, PIPELINE_LITERAL
    from gfx_ocean_b200.shard import fan_out, tiles_of_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ocean path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity_before, numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, K, W = args.resolution, args.steps, max(args.warmup, 3)
    strong = args.total_tiles > 0
    total_tiles = args.total_tiles if strong else args.tiles * world
    if total_tiles % world:
        raise SystemExit("--total-tiles must be a multiple of the GPU count")

    # tile-parallel sharding: rank r owns a contiguous block of the global tiles; no data-path collective
    my_tiles = tiles_of_rank(rank, world, total_tiles)
    tiles = len(my_tiles)
    stream = torch.cuda.Stream()
    pipeline = PIPELINE_FUSED if args.pipeline == "fused" else PIPELINE_LITERAL
    ocean = Ocean(n, 1000.0, n_tiles=tiles, device=local_rank, pipeline=pipeline, stream=stream.cuda_stream)
    for i, g in enumerate(my_tiles):
        ocean.generate_spectrum(i, SEED, stream_id=g)          # on the device: no host upload

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

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps):
        """ms for `steps` calls of fn(i), bracketed by barrier + synchronize, max over ranks."""
        barrier()
        ev0.record(stream)
        # launch fan-out, inside the timed region: rank 0 broadcasts the frame block (one NCCL broadcast per `steps`
        # frames); the frames themselves need no communication
        first, count, _ = fan_out(0 if rank == 0 else -1, steps if rank == 0 else -1, args.dt, device="cuda")
        for i in range(first, first + count):
            fn(i)
        ev1.record(stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    def repeated(fn, steps, min_reps, min_time):
        reps = [timed(fn, steps)]
        while len(reps) < min_reps or (sum(reps) < min_time * 1e3 and len(reps) < 400):
            reps.append(timed(fn, steps))
        return reps

    # ---- device-resident throughput: R repetitions of K steps, inputs already in HBM.
    #      Headline: ocean_update_overlapped -- consecutive frames alternate between two lanes (stream + own row-pass
    #      intermediate), so the row kernel of step n+1 runs beside the column kernel of step n; bit-identical maps
    #      (tests/test_gpu_features.py). The closing event is recorded behind ocean_join, i.e. after both lanes drained.
    #      plain_updates: the same steps through ocean_update (one stream, kernels chained by programmatic dependent launch).
    lanes_ok = args.pipeline == "fused"

    def headline_step(i):
        if lanes_ok:
            ocean.update_overlapped(args.dt * (W + i))
            if i == K - 1:
                ocean.join()
        else:
            ocean.update(args.dt * (W + i))

    with torch.cuda.stream(stream):
        for i in range(W):
            ocean.update(args.dt * i)
            if lanes_ok:
                ocean.update_overlapped(args.dt * i)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            for i in range(max(1, int(0.4 / 1.2e-4))):      # keep the GPU loaded while nvidia-smi spins up
                ocean.update(args.dt * i)
        l0 = ocean.launch_count
        reps = repeated(headline_step, K, args.reps, args.min_time)
        launches = (ocean.launch_count - l0) // len(reps)
        clocks = sampler.stop() if rank == 0 else None
        reps_plain = repeated(lambda i: ocean.update(args.dt * (W + i)), K, args.reps, args.min_time) if lanes_ok else reps
    ms = statistics.median(reps)
    fps = total_tiles * K / (ms * 1e-3)
    ms_plain = statistics.median(reps_plain)
    plain = {"value": total_tiles * K / (ms_plain * 1e-3), "unit": UNIT, "ms_per_step": ms_plain / K, "repetitions": len(reps_plain),
             "alg_GBps": ALG_BYTES_PER_POINT * n * n * tiles / (ms_plain / K * 1e-3) / 1e9,
             "note": "the same steps through ocean_update: one stream, k_rows_t / k_cols chained by programmatic dependent launch"}

    # ---- tile determinism: every tile's checksum at a fixed time, gathered; rank 0 recomputes ALL tiles on its own
    #      GPU (tile g = generate_spectrum(SEED, stream_id=g) anywhere) and compares bit for bit
    t_check = 3.25
    determinism = None
    with torch.cuda.stream(stream):
        ocean.update(t_check)
        mine = torch.from_numpy(ocean.output_checksums().astype(np.int64)).cuda()
        if world > 1:
            allsums = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allsums, mine)
            allsums = torch.cat(allsums).cpu().numpy()
        else:
            allsums = mine.cpu().numpy()
        if rank == 0:
            redo = []
            for g0 in range(0, total_tiles, tiles):
                for i in range(tiles):
                    ocean.generate_spectrum(i, SEED, stream_id=g0 + i)
                ocean.update(t_check)
                redo.append(ocean.output_checksums().astype(np.int64))
            redo = np.concatenate(redo)
            for i, g in enumerate(my_tiles):                   # restore rank 0's own tiles
                ocean.generate_spectrum(i, SEED, stream_id=g)
            determinism = {"tiles_checked": int(total_tiles), "bit_identical_to_single_gpu": bool(np.array_equal(redo, allsums)),
                           "checksum_xor": "%016x" % (int(np.bitwise_xor.reduce(allsums.astype(np.uint64))) & (2 ** 64 - 1)),
                           "time": t_check}
        barrier()

    extras = rank == 0 and world == 1 and not args.no_extras and args.pipeline == "fused"

    # ---- latency mode: ONE tile per ocean_update, rotating over the context's tiles so each frame's inputs
    #      are L2-cold (8 x 12 MB inputs + 8 x 16 MB outputs > 126 MB L2); this is BASELINE.json configs[2]
    #      taken literally (a single 1024^2 ocean per frame); (i) plain launches, (ii) CUDA-graph replay
    single = None
    if tiles >= 8 and args.pipeline == "fused":
        ks = max(200, K // 4)
        with torch.cuda.stream(stream):
            for i in range(W):
                ocean.update_tiles(args.dt * i, i % tiles, 1)
                ocean.update_graph(args.dt * i, i % tiles, 1)
            r_plain = repeated(lambda i: ocean.update_tiles(args.dt * i, i % tiles, 1), ks, 5, 0.2)
            r_graph = repeated(lambda i: ocean.update_graph(args.dt * i, i % tiles, 1), ks, 5, 0.2)
            for i in range(W):
                ocean.update_overlapped(args.dt * i, i % tiles, 1)
            ocean.sync()

            def lanes(i):
                ocean.update_overlapped(args.dt * i, i % tiles, 1)
                if i == ks - 1:
                    ocean.join()                 # the closing event is recorded behind both lanes
            r_lanes = repeated(lanes, ks, 5, 0.2)

        def lat(r):
            m = statistics.median(r)
            return {"value": world * ks / (m * 1e-3), "unit": UNIT, "us_per_frame": 1e3 * m / ks,
                    "alg_GBps": ALG_BYTES_PER_POINT * n * n / (m / ks * 1e-3) / 1e9}
        single = {"plain": lat(r_plain), "graph": lat(r_graph), "overlapped": lat(r_lanes), "steps": ks, "repetitions": len(r_plain),
                  "note": "one tile per ocean_update (2 launches), rotating over the tiles so inputs are L2-cold; "
                          "plain = cudaLaunchKernelEx per kernel, graph = ocean_update_graph replay, overlapped = "
                          "ocean_update_overlapped (consecutive frames, which are different tiles, alternate between two "
                          "lanes: the row kernel of frame n+1 runs beside the column kernel of frame n)"}
        single.update(single["plain"])      # keep round 1's flat keys

    # ---- per-kernel durations (CUDA events between the two launches, same stream)
    stage_ms = None
    if args.pipeline == "fused":
        acc = []
        for i in range(30):
            acc.append(ocean.profile_update(args.dt * i))
        stage_ms = np.median(np.array(acc), axis=0).tolist()

    # ---- consumer step: frames/s with the normal map of ocean.frag:50-66 computed after every frame; (i) gathered from
    #      the RGBA texels, (ii) on a context whose column kernel also writes a dense copy of channel .x
    with_normals = None
    if extras:
        kn = max(50, K // 10)

        def normals_rate(o):
            with torch.cuda.stream(stream):
                for i in range(3):
                    o.update(args.dt * i)
                    o.compute_normals()
                rn = repeated(lambda i: (o.update(args.dt * i), o.compute_normals()), kn, 5, 0.2)
            return statistics.median(rn)
        mn = normals_rate(ocean)
        with Ocean(n, 1000.0, n_tiles=tiles, device=local_rank, stream=stream.cuda_stream, flags=FLAG_DX_PLANE) as on:
            for i, g in enumerate(my_tiles):
                on.generate_spectrum(i, SEED, stream_id=g)
            mp = normals_rate(on)
        with_normals = {"value": tiles * kn / (mp * 1e-3), "unit": UNIT, "ms_per_step": mp / kn,
                        "normal_map_ms_per_step": mp / kn - ms / K,
                        "from_rgba_texels": {"value": tiles * kn / (mn * 1e-3), "ms_per_step": mn / kn},
                        "note": "ocean_update + ocean_compute_normals every frame. value: OCEAN_FLAG_DX_PLANE context (k_cols "
                                "also writes channel .x densely, +4 B/pt; the normal kernel reads 4 + writes 16 B/pt); "
                                "from_rgba_texels: the normal kernel gathers .x out of the RGBA map (16 + 16 B/pt)"}

    # ---- library sanity bar: cuFFT's batched 2-D C2C inverse transform ALONE on the same
    #      amount of data (3 complex fields per tile, via torch.fft.ifft2), without propagate or correction.
    #      A comparison line only: nothing of it is on the product path.
    cufft_ms = None
    if extras:
        try:
            spec = torch.randn(tiles * 3, n, n, dtype=torch.complex64, device="cuda")
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

    # ---- other BASELINE.json configurations on this GPU (device-resident, same timing method)
    other = None
    if extras:
        other = {}
        for name, on, ot in (("512x512 (configs[1], the reference's own size)", 512, 32), ("2048x2048 (configs[3])", 2048, 2),
                             ("1024x1024 x 64 tiles on one GPU (configs[4] at 1 GPU)", 1024, 64)):
            try:
                with Ocean(on, 1000.0, n_tiles=ot, device=local_rank, stream=stream.cuda_stream) as o2:
                    for i in range(ot):
                        o2.generate_spectrum(i, SEED, stream_id=i)
                    ko = max(20, int(K * (8 * 1024 * 1024) / (ot * on * on)) // 4)
                    with torch.cuda.stream(stream):
                        for i in range(5):
                            o2.update(args.dt * i)
                        ro = repeated(lambda i: o2.update(args.dt * i), ko, 5, 0.2)
                    mo = statistics.median(ro)
                    peak_o, _ = measured_peak()
                    gbps = ALG_BYTES_PER_POINT * on * on * ot / (mo / ko * 1e-3) / 1e9
                    other[name] = {"value": ot * ko / (mo * 1e-3), "unit": UNIT, "tiles_per_step": ot, "ms_per_step": mo / ko,
                                   "alg_GBps": gbps, "frac_of_measured_peak": gbps / peak_o}
            except Exception as exc:
                other[name] = f"failed: {exc}"

    # ---- end to end through the public API with host buffers: update + read back every step, double-buffered
    #      (frame n+1 is computed while frame n is copied; the host owns frame n-1 after download_fence(1))
    nbytes_out = n * n * 16 * tiles
    e2e_steps = max(10, min(K, 100))
    host = [torch.empty((tiles, n, n, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    with Ocean(n, 1000.0, n_tiles=tiles, device=local_rank, pipeline=pipeline, stream=stream.cuda_stream,
               flags=FLAG_DOUBLE_BUFFER_OUTPUT) as oe:
        for i, g in enumerate(my_tiles):
            oe.generate_spectrum(i, SEED, stream_id=g)

        def e2e_step(i):
            oe.update(args.dt * (i + 7))                 # PropagateLocals (12 B) travel as kernel parameters
            oe.read_back_all_async(host[i % 2].data_ptr())
            oe.download_fence(1)                         # frame i-1 is on the host; its buffer is free again

        with torch.cuda.stream(stream):
            for i in range(3):
                e2e_step(i)
            oe.sync()
            barrier()
            ev0.record(stream)
            for i in range(e2e_steps):
                e2e_step(i)
            oe.sync()                                    # the last frame has landed too
            ev1.record(stream)
            barrier()
        e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
        e2e_check = int(oe.output_checksums()[0])
    e2e_fps = total_tiles * e2e_steps / (e2e_ms * 1e-3)

    # ---- the ceiling of that path: the same D2H copies alone (same pinned buffers, all ranks at once)
    dev = torch.empty((tiles, n, n, 4), dtype=torch.float32, device="cuda")
    with torch.cuda.stream(stream):
        for i in range(2):
            host[i % 2].copy_(dev, non_blocking=True)
        barrier()
        ev0.record(stream)
        for i in range(e2e_steps):
            host[i % 2].copy_(dev, non_blocking=True)
        ev1.record(stream)
        barrier()
    d2h_ms_own = ev0.elapsed_time(ev1)
    d2h_gbps = sum_over_ranks(nbytes_out * e2e_steps / (d2h_ms_own * 1e-3) / 1e9)
    d2h_ms = max_over_ranks(d2h_ms_own)
    del dev

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_step = ALG_BYTES_PER_POINT * n * n * tiles
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n}x{n} x3 fields (height, dx, dz), {tiles} independent tiles per GPU per step; "
                                   "frame = one tile's propagate -> 2-D iFFT -> correction -> RGBA32F map",
                       "resolution": n, "tiles_per_gpu": tiles, "total_tiles": total_tiles, "pipeline": args.pipeline,
                       "update_call": "ocean_update_overlapped (two lanes)" if lanes_ok else "ocean_update",
                       "inputs": f"ocean_generate_spectrum(seed={SEED}, stream_id=global tile) on the device",
                       "l2": "inputs larger than L2 (per-step working set %.0f MB > 126 MB)" % (
                           (12 + 12 + 16) * n * n * tiles / 1e6),
                       "parallelism": f"tile-parallel x{world}, no data-path collective",
                       "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
            "repetitions": len(reps),
            "value_stats": {"min": total_tiles * K / (max(reps) * 1e-3), "median": fps, "max": total_tiles * K / (min(reps) * 1e-3),
                            "timed_seconds": sum(reps) * 1e-3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": 12, "d2h_bytes_per_step": nbytes_out,
                    "steps": e2e_steps, "d2h_only_GBps": d2h_gbps, "achieved_GBps": total_tiles * n * n * 16 * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "frac_of_d2h_ceiling": d2h_ms / e2e_ms, "last_frame_checksum": "%016x" % (e2e_check & (2 ** 64 - 1)),
                    "note": "Ocean.update(t) + ocean_download_all_async of every tile into pinned host memory per step on a "
                            "double-buffered context (copy stream overlaps the next frame's kernels); d2h_only_GBps = the same "
                            "copies with no compute, all ranks at once (the PCIe / host-memory ceiling of this box)"},
            "step_alg_GBps": alg_step / (ms / K * 1e-3) / 1e9,
            "step_frac_of_measured_peak": alg_step / (ms / K * 1e-3) / 1e9 / peak,
            "step_real_bytes_frac": REAL_BYTES_PER_POINT * n * n * tiles / (ms / K * 1e-3) / 1e9 / peak,
            "plain_updates": plain,
            "tile_determinism": determinism,
            "single_tile_per_update": single,
            "with_normals": with_normals,
            "cufft_ifft2_only_ms_per_step": cufft_ms,
            "other_configs": other,
        }
        if stage_ms:
            alg = [ALG_BYTES_ROWS * n * n * tiles, ALG_BYTES_COLS * n * n * tiles]
            real = [24 * n * n * tiles, 28 * n * n * tiles]
            names = ["k_rows", "k_cols"]
            d = int(np.argmax(stage_ms))
            ach = alg[d] / (stage_ms[d] * 1e-3) / 1e9
            traffic = recorded_traffic(names[d])
            line["roofline"] = {"bound": "hbm", "kernel": names[d], "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                                "alg_bytes_per_launch": alg[d], "launch_ms": stage_ms[d],
                                "launch_ms_clock": "CUDA events recorded between the launches on the launching stream "
                                                   "(ocean_profile_update, median of 30 frames); includes the inter-kernel gap. The event "
                                                   "between the kernels serialises them, so the two durations add up to MORE than ms_per_step: in "
                                                   "the timed loop programmatic dependent launch overlaps the head of k_rows with the tail of k_cols",
                                "real_bytes_per_launch": real[d], "real_bytes_frac": real[d] / (stage_ms[d] * 1e-3) / 1e9 / peak,
                                "traffic_frac": (traffic / (stage_ms[d] * 1e-3) / 1e9 / peak) if traffic else None,
                                "traffic_source": "profiles/traffic.json (ncu --set full capture of the same kernel, per launch)"}
            line["kernels"] = [{"name": names[i], "launch_ms": stage_ms[i], "alg_bytes": alg[i],
                                "alg_GBps": alg[i] / (stage_ms[i] * 1e-3) / 1e9,
                                "real_bytes_GBps": real[i] / (stage_ms[i] * 1e-3) / 1e9} for i in range(2)]
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, affinity_before)         # the CPU leg uses every host thread again
            data = [ocean.get_spectrum(i) for i in range(min(2, tiles))]
            cfps, cores, frames, el = cpu_reference_fps(n, data, args.cpu_seconds, args.dt)
            line["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{frames} frames at {n}x{n} in {el:.1f} s (literal fp32 restatement, OpenMP)"}
        print(json.dumps(line), flush=True)
    ocean.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
