mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r3d_pytest.log
V=$PWD/gfx_ocean_b200/variants
for lib in notrim default; do
  if [ $lib = default ]; then unset OCEAN_B200_LIB; else export OCEAN_B200_LIB=$V/libocean_b200.$lib.so; fi
  for rep in 1 2; do
  timeout 300 python bench.py --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/r3d_bench_${lib}_$rep.json 2>> gpurun_out/r3d_bench_$lib.err
  done
  timeout 300 python bench.py --resolution 2048 --tiles 2 --steps 300 --no-extras --no-cpu-baseline > gpurun_out/r3d_bench2048_${lib}.json 2>> gpurun_out/r3d_bench_$lib.err
done
echo done
