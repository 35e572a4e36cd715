#!/bin/bash
# Produces the evidence kept under profiles/ (run on one B200 through gpurun):
#   scripts/gpu_profile.sh r02     -> gpurun_out/r02_*
# pytest log, the default bench line, the ncu launch list of the bench command, full ncu captures of the two
# kernels of the product path (1024^2 x 8 tiles and 2048^2 x 2 tiles), the row-kernel alternatives, the bulk-copy
# latency micro-benchmark and compute-sanitizer logs. FULL=1 also re-captures the alternative row kernels under ncu.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_default_1xB200.json 2> gpurun_out/${tag}_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/${tag}_launches_1024x8.csv \
    python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --reps 1 --min-time 0 > gpurun_out/${tag}_launches.log 2>&1
for k in k_rows_t k_cols; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k} -s 4 -c 1 -f -o gpurun_out/${tag}_${k} \
      python scripts/prof_target.py 1024 8 6 > gpurun_out/${tag}_ncu_${k}.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k} -s 4 -c 1 -f -o gpurun_out/${tag}_2048_${k} \
      python scripts/prof_target.py 2048 2 6 > gpurun_out/${tag}_ncu_2048_${k}.log 2>&1
done
if [ -n "$FULL" ]; then
OCEAN_B200_ROWS=staged timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_rows<" -s 4 -c 1 -f \
    -o gpurun_out/${tag}_k_rows_staged python scripts/prof_target.py 1024 8 6 > gpurun_out/${tag}_ncu_k_rows_staged.log 2>&1
OCEAN_B200_ROWS=persistent timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rows_p -s 4 -c 1 -f \
    -o gpurun_out/${tag}_k_rows_p python scripts/prof_target.py 1024 8 6 > gpurun_out/${tag}_ncu_k_rows_p.log 2>&1
fi
for m in tma staged persistent fold; do
  OCEAN_B200_ROWS=$m timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/${tag}_rows_${m}.json 2>/dev/null
done
timeout 600 python bench.py --total-tiles 64 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_64tiles_1xB200.json 2>/dev/null
# (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bulk_latency scripts/ubench/bulk_latency.cu, built before the call)
if [ -x build/bulk_latency ]; then for m in 0 1 2 3; do ./build/bulk_latency $m; done > gpurun_out/${tag}_bulk_latency.log 2>&1; fi
for tool in memcheck synccheck racecheck; do
  for cfg in "1024 4 2" "512 3 3" "2048 1 2"; do
    echo "== compute-sanitizer --tool $tool python scripts/san_target.py $cfg"
    timeout 600 compute-sanitizer --tool $tool --print-limit 4 python scripts/san_target.py $cfg 2>&1 | grep -E "SUMMARY|Error:|checksums|hazards" | cut -c1-260 | head -12
  done
  echo "== compute-sanitizer --tool $tool python scripts/san_target.py 512 3 6 overlapped"
  timeout 600 compute-sanitizer --tool $tool --print-limit 4 python scripts/san_target.py 512 3 6 overlapped 2>&1 | grep -E "SUMMARY|Error:|checksums|hazards" | cut -c1-260 | head -12
done > gpurun_out/${tag}_sanitizer.log 2>&1
echo done
