mkdir -p gpurun_out
V=$PWD/gfx_ocean_b200/variants
for lib in default emit; do
  if [ $lib = default ]; then unset OCEAN_B200_LIB; else export OCEAN_B200_LIB=$V/libocean_b200.$lib.so; fi
  python scripts/san_target.py 1024 3 3 > gpurun_out/r3a_sums_$lib.log 2>&1
  for rep in 1 2; do
  timeout 300 python bench.py --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/r3a_bench_${lib}_$rep.json 2>> gpurun_out/r3a_bench_$lib.err
  done
  timeout 300 python bench.py --resolution 512 --tiles 32 --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/r3a_bench512_${lib}.json 2>> gpurun_out/r3a_bench_$lib.err
  timeout 300 python bench.py --resolution 2048 --tiles 2 --steps 300 --no-extras --no-cpu-baseline > gpurun_out/r3a_bench2048_${lib}.json 2>> gpurun_out/r3a_bench_$lib.err
done
export OCEAN_B200_LIB=$V/libocean_b200.emit.so
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r3a_pytest_emit.log
echo done
