mkdir -p gpurun_out
V=$PWD/gfx_ocean_b200/variants
for lib in default w2048; do
  if [ $lib = default ]; then unset OCEAN_B200_LIB; else export OCEAN_B200_LIB=$V/libocean_b200.$lib.so; fi
  python scripts/san_target.py 2048 2 3 > gpurun_out/r2w_sums_$lib.log 2>&1
  for rep in 1 2; do
  timeout 300 python bench.py --resolution 2048 --tiles 2 --steps 300 --no-extras --no-cpu-baseline > gpurun_out/r2w_bench2048_${lib}_$rep.json 2>> gpurun_out/r2w_bench_$lib.err
  done
done
unset OCEAN_B200_LIB
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2w_pytest.log
echo done
