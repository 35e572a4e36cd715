"""Cross-lane ordering of ocean_update_overlapped (gfx_ocean_b200/csrc/host/lane_order.hpp), on the host: the
bookkeeping the library uses is run against a happens-before model of the streams and events it enqueues -- two
frames on different lanes that write a common tile are always ordered, for 4000 random call sequences and for the
returning-frame case; the latest-frame-only rule the entry point first shipped with is shown to leave gaps."""
import os
import subprocess

from conftest import ROOT


def test_lane_ordering_against_a_happens_before_model(tmp_path):
    csrc = os.path.join(ROOT, "gfx_ocean_b200", "csrc")
    exe = str(tmp_path / "host_check_lane_order")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", csrc, "-o", exe, os.path.join(csrc, "host_check_lane_order.cpp")],
                   check=True, capture_output=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "0 unordered pairs" in r.stdout and "latest-frame-only rule 1" in r.stdout
