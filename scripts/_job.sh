mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/r3c_bench_2gpu.json 2> gpurun_out/r3c_bench_2gpu.err
timeout 600 python bench.py --steps 500 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/r3c_bench_1gpu.json 2> gpurun_out/r3c_bench_1gpu.err
echo done
