"""Generate tests/golden/spv_512.npz by EXECUTING the reference's shipped SPIR-V on the CPU.

Run in the DEV container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_spv_golden.py

oracle/spv_exec.py interprets /root/reference/shader/spv/*.spv (the binaries the reference embeds with
`include_bytes!`, src/fft.rs:20-25, src/ocean.rs:26-28,195-197) and replays the per-frame dispatch
sequence of src/render.rs:1122-1287 on the reference-owned inputs data/omega.bin + data/spectrum.bin.
Unlike golden_512.npz (outputs of the self-authored oracle), these vectors come from reference-authored
code, so they PIN the oracle and the CUDA path:

* frames[t]        displacement image channels (dx, height, dz) at t in {0, 1, 37.5, 600}, full 512x512
* spectra_t1_*     post-propagate spectra probes at t = 1 (the uint-wrap quirk lives here)
* normals_t1       local `N` of ocean.frag.spv (ocean.frag:50-66) at every texel centre of frames[t=1]
* vertex_t1        p_PosWorld of ocean.vert.spv (ocean.vert:21-25) for the reference's 128x128 vertex
                   grid (src/render.rs:498-506) and patch offset (127, 0) (src/render.rs:544)
* spv_sha256       digests of the six modules that were executed
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import spv_exec  # noqa: E402
from gfx_ocean_b200.bincode import read_vec_f32, read_vec_f32x2  # noqa: E402

REF = "/root/reference"
SPV = os.path.join(REF, "shader", "spv")
G = os.path.join(ROOT, "tests", "golden")
TIMES = (0.0, 1.0, 37.5, 600.0)
HALF_RESOLUTION = 128          # src/render.rs:45
MODULES = ("propagate.comp", "fft_row.comp", "fft_col.comp", "correction.comp", "ocean.vert", "ocean.frag")


def vertex_grid(h=HALF_RESOLUTION):
    """a_Pos / a_Uv exactly as src/render.rs:498-506 builds them (f32 division)."""
    g = np.arange(h, dtype=np.float32)
    x, z = np.meshgrid(g, g, indexing="xy")
    pos = np.stack([x, np.zeros_like(x), z], -1).reshape(-1, 3)
    uv = np.stack([x / np.float32(h - 1), z / np.float32(h - 1)], -1).reshape(-1, 2)
    return pos, uv


def main():
    om = read_vec_f32(os.path.join(REF, "data", "omega.bin")).reshape(512, 512)
    sp = read_vec_f32x2(os.path.join(REF, "data", "spectrum.bin")).reshape(512, 512, 2)
    d = {"times": np.array(TIMES, np.float32)}
    d["spv_names"] = np.array(MODULES)
    d["spv_sha256"] = np.array([hashlib.sha256(open(os.path.join(SPV, m + ".spv"), "rb").read()).hexdigest() for m in MODULES])
    frames = np.zeros((len(TIMES), 512, 512, 3), np.float32)
    for i, t in enumerate(TIMES):
        if t == 1.0:
            img, (dy, dx, dz) = spv_exec.run_reference_frame(SPV, sp, om, t, keep_spectra=True)
            rng = np.random.default_rng(11)
            idx = np.concatenate([[0, 256 * 512 + 256, 300 * 512 + 200, 100 * 512 + 400, 512 * 512 - 1], rng.integers(0, 512 * 512, 4091)])
            d["spectra_t1_index"] = idx
            d["spectra_t1_h"], d["spectra_t1_dx"], d["spectra_t1_dz"] = dy[idx], dx[idx], dz[idx]
            xs = (np.arange(512, dtype=np.float32) + np.float32(0.5)) / np.float32(512)
            uv = np.stack(np.meshgrid(xs, xs, indexing="xy"), -1).reshape(-1, 2)
            nrm, _ = spv_exec.run_fragment_normals(SPV, img, uv)
            d["normals_t1"] = nrm.reshape(512, 512, 3).astype(np.float32)
            pos, vuv = vertex_grid()
            d["vertex_offset"] = np.array([HALF_RESOLUTION - 1, 0.0], np.float32)
            d["vertex_t1"] = spv_exec.run_vertex_displacement(SPV, img, pos, vuv, d["vertex_offset"]).astype(np.float32)
        else:
            img = spv_exec.run_reference_frame(SPV, sp, om, t)
        assert np.all(img[..., 3] == 0.0) and not np.signbit(img[..., 3]).any()
        frames[i] = img[..., :3]
        print(f"t={t}: max|dx,h,dz| = {np.abs(img[..., :3]).max(axis=(0, 1))}")
    d["frames"] = frames
    np.savez_compressed(os.path.join(G, "spv_512.npz"), **d)
    print("wrote", os.path.join(G, "spv_512.npz"), os.path.getsize(os.path.join(G, "spv_512.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
