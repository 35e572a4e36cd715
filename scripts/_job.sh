mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2q_pytest.log
for tool in racecheck memcheck; do echo "== $tool overlapped"; timeout 500 compute-sanitizer --tool $tool --print-limit 4 python - <<'PY' 2>&1 | grep -E "SUMMARY|Error|sums" | head
import sys; sys.path.insert(0, '.')
from gfx_ocean_b200 import Ocean
with Ocean(512, 1000.0, n_tiles=3) as o:
    for i in range(3): o.generate_spectrum(i, 7, stream_id=i)
    for f in range(6): o.update_overlapped(0.1 * f)
    for f in range(6): o.update_overlapped(0.1 * f, f % 3, 1)
    print("sums", [hex(int(s)) for s in o.output_checksums()])
PY
done > gpurun_out/r2q_sanitizer.log 2>&1
