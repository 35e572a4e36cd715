"""Host build of the in-register transform templates (csrc/fft_reg.cuh) checked against a
naive f64 DFT -- nvcc compiles the same __host__ __device__ code for the CPU, no GPU needed."""
import os
import subprocess

from conftest import ROOT


def test_register_fft_templates_on_host(tmp_path):
    src = os.path.join(ROOT, "gfx_ocean_b200", "csrc", "host_check_fft_reg.cu")
    exe = str(tmp_path / "host_check")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True,
                   capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "R=32" in r.stdout and "twiddle" in r.stdout


def test_sincos_reduced_accuracy(tmp_path):
    """The Cody-Waite + minimax sincos of the propagate phase (ocean_common.cuh): max abs error vs f64 < 1e-7
    over 2.7 M samples of [-1e5, 1e5] incl. the neighbourhood of every quadrant boundary."""
    src = os.path.join(ROOT, "gfx_ocean_b200", "csrc", "host_check_sincos.cu")
    exe = str(tmp_path / "host_check_sincos")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
