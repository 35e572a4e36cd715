"""In-tree build of libocean_b200.so (hand-written sm_100a kernels + the C ABI).

    python -m gfx_ocean_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file, is
git-ignored, and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libocean_b200.so")
SOURCES = ["ocean_api.cu", "kernels_literal.cu", "kernels_consumer.cu", "kernels_fused.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    # no --use_fast_math: the propagate phase needs the full-range sincosf (phases reach 1e3..1e4 rad)
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libocean_b200.so cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ocean_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str | None = None, extra: list[str] | None = None) -> str:
    """Build the default library, or -- with `variant` -- an A/B build with extra nvcc flags into
    gfx_ocean_b200/variants/libocean_b200.<variant>.so (separate object directory; the default library is
    never touched, so later tests and benches cannot silently run an experimental configuration)."""
    extra = list(extra or [])
    if variant is None:
        if extra:
            raise ValueError("extra flags need a variant name: the default library is always the default configuration")
        if not force and not _stale():
            return LIB
        out, bdir = LIB, os.path.join(HERE, "build")
    else:
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        out, bdir = os.path.join(HERE, "variants", f"libocean_b200.{variant}.so"), os.path.join(HERE, "build", variant)
    objs = []
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            print(r.stdout)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stdout)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        print(r.stdout)
        raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    # python -m gfx_ocean_b200.build [--force] [-v] [--variant NAME -DFLAG=.. ...]
    argv = sys.argv[1:]
    variant = argv[argv.index("--variant") + 1] if "--variant" in argv else None
    extra = [a for a in argv if a.startswith("-D") or a.startswith("-maxrregcount")]
    print(build(force="--force" in argv, verbose="-v" in argv, variant=variant, extra=extra))
