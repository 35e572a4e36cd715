#!/bin/bash
# round-2 GPU job C: full GPU test suite, TMA-fed persistent k_rows A/B, new bench.py, ncu captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2c_pytest.log
B="python bench.py --steps 1500 --no-cpu-baseline --no-extras"
OCEAN_B200_ROWS=legacy timeout 300 $B > gpurun_out/r2c_bench_legacy.json 2> gpurun_out/r2c_bench_legacy.err
timeout 300 $B > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
for v in s0 g4s3 g4s2; do
  OCEAN_B200_LIB=$PWD/gfx_ocean_b200/variants/libocean_b200.$v.so timeout 300 $B > gpurun_out/r2c_bench_$v.json 2> gpurun_out/r2c_bench_$v.err
done
timeout 600 python bench.py > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rows_p -s 4 -c 1 -f -o gpurun_out/r2c_rows python scripts/prof_target.py 1024 8 6 > gpurun_out/r2c_ncu.log 2>&1
for tool in racecheck memcheck; do echo "== $tool"; timeout 500 compute-sanitizer --tool $tool --print-limit 6 python scripts/san_target.py 1024 4 2 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         " | tail -12; done > gpurun_out/r2c_sanitizer.log 2>&1
echo done
