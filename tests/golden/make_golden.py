"""Generate the committed golden fixtures under tests/golden/.

Run from the repo root in the DEV container (it reads /root/reference, which does not
exist on the GPU box):   python tests/golden/make_golden.py

* ref_data/omega.bin, ref_data/spectrum.bin -- byte copies of the reference-owned INPUT
  fixtures /root/reference/data/{omega,spectrum}.bin (bincode Vec<f32> / Vec<[f32;2]>,
  loaded at src/render.rs:769-771,808-810). They are data, not source.
* golden_512.npz -- outputs of the f64 oracle (oracle/ocean_oracle.c, literal pi32
  Stockham) on that data at t in {0, 1, 37.5, 600}: 4096 seeded probe texels + the named
  probes of SURVEY.md 8c, per-channel sum|.| and max|.|, and post-propagate spectra
  probes at t=0 (the uint-wrap quirk cases).
* golden_synth.npz -- same for the seeded synthetic 1024^2 tile 0 (t in {0, 1, 37.5}) and
  a 2048^2 tile at t=1, plus checksums of the synthetic inputs themselves.

The reference has no golden OUTPUTS of its own ("parity unpinned"): these vectors pin the
oracle against regressions and across machines, nothing more.
"""
import hashlib
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ocean_oracle import COracle  # noqa: E402
from gfx_ocean_b200.bincode import read_vec_f32, read_vec_f32x2  # noqa: E402
from gfx_ocean_b200.spectrum import synthetic_tile  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/data"


def probes(n, count=4096, seed=7):
    rng = np.random.default_rng(seed)
    ys = rng.integers(0, n, count)
    xs = rng.integers(0, n, count)
    named = np.array([[0, 0], [17, 300 % n], [n - 1, n - 1], [n // 2, n // 2], [n // 2 + 1, n // 2 + 4]])
    return np.concatenate([named[:, 0], ys]), np.concatenate([named[:, 1], xs])


def summarise(out):
    return (np.abs(out[..., :3]).sum(axis=(0, 1)), np.abs(out[..., :3]).max(axis=(0, 1)))


def main():
    os.makedirs(os.path.join(G, "ref_data"), exist_ok=True)
    for f in ("omega.bin", "spectrum.bin"):
        if os.path.exists(os.path.join(REF, f)):
            shutil.copyfile(os.path.join(REF, f), os.path.join(G, "ref_data", f))
    om = read_vec_f32(os.path.join(G, "ref_data", "omega.bin"))
    sp = read_vec_f32x2(os.path.join(G, "ref_data", "spectrum.bin"))
    o = COracle()

    n = 512
    py, px = probes(n)
    d = {"probe_y": py, "probe_x": px, "times": np.array([0.0, 1.0, 37.5, 600.0])}
    for i, t in enumerate(d["times"]):
        out = o.frame(sp, om, float(t), n, prec="f64")
        d[f"probe_out_{i}"] = out[py, px, :]
        d[f"sum_abs_{i}"], d[f"max_abs_{i}"] = summarise(out)
    h, dx, dz = o.propagate(sp, om, 0.0, n, prec="f64")
    qy = np.array([300, 100, 10, 400]); qx = np.array([200, 400, 20, 500])
    d["spec_y"], d["spec_x"] = qy, qx
    d["spec_h"], d["spec_dx"], d["spec_dz"] = h[qy, qx], dx[qy, qx], dz[qy, qx]
    d["input_sha256"] = np.array([hashlib.sha256(open(os.path.join(G, "ref_data", f), "rb").read()).hexdigest()
                                  for f in ("omega.bin", "spectrum.bin")])
    np.savez_compressed(os.path.join(G, "golden_512.npz"), **d)

    s = {}
    for n, times in ((1024, (0.0, 1.0, 37.5)), (2048, (1.0,))):
        h0, w = synthetic_tile(n, tile=0)
        py, px = probes(n, 2048, seed=11 + n)
        s[f"n{n}_probe_y"], s[f"n{n}_probe_x"] = py, px
        s[f"n{n}_times"] = np.array(times)
        s[f"n{n}_input_sums"] = np.array([np.abs(h0.astype(np.float64)).sum(), w.astype(np.float64).sum()])
        for i, t in enumerate(times):
            out = o.frame(h0, w, float(t), n, prec="f64")
            s[f"n{n}_probe_out_{i}"] = out[py, px, :]
            s[f"n{n}_sum_abs_{i}"], s[f"n{n}_max_abs_{i}"] = summarise(out)
    np.savez_compressed(os.path.join(G, "golden_synth.npz"), **s)
    print("wrote", os.listdir(G))


if __name__ == "__main__":
    main()
