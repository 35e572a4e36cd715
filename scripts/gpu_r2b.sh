#!/bin/bash
# round-2 GPU job B: full GPU test suite, A/B timing of the persistent k_rows variants, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2b_pytest.log
B="python bench.py --steps 1500 --no-cpu-baseline"
OCEAN_B200_ROWS=legacy timeout 300 $B > gpurun_out/r2b_bench_legacy.json 2> gpurun_out/r2b_bench_legacy.err
timeout 300 $B > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
for v in g4 g5u2 g5u3; do
  OCEAN_B200_LIB=$PWD/gfx_ocean_b200/variants/libocean_b200.$v.so timeout 300 $B > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
done
OCEAN_B200_PDL=0 timeout 300 $B > gpurun_out/r2b_bench_default_pdl0.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rows_p -s 4 -c 1 -f -o gpurun_out/r2b_rows python scripts/prof_target.py 1024 8 6 > gpurun_out/r2b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cols -s 4 -c 1 -f -o gpurun_out/r2b_cols python scripts/prof_target.py 1024 8 6 >> gpurun_out/r2b_ncu.log 2>&1
echo done
