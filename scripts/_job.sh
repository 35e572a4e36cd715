mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2z_pytest.log
python scripts/san_target.py 2048 2 3 > gpurun_out/r2z_sums_2048.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 10 > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2z_bench_2gpu_ref.json 2> gpurun_out/r2z_bench_2gpu_ref.err
echo done
