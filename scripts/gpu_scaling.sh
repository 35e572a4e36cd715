#!/bin/bash
# 1 -> 8 GPU scaling record (gpurun --gpus 8): weak (8 tiles per GPU) and strong (64 tiles in total, BASELINE.json configs[4]).
#   scripts/gpu_scaling.sh [tag] ["1 2 4 8"]   -> gpurun_out/<tag>_scale_{weak,strong}_<n>.json
tag=${1:-r02}
counts=${2:-"1 2 4 8"}
mkdir -p gpurun_out
for n in $counts; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n"; fi
  timeout 600 $L bench.py --gpus $n --steps 1000 --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/${tag}_scale_weak_$n.json
  timeout 600 $L bench.py --gpus $n --total-tiles 64 --steps 200 --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/${tag}_scale_strong_$n.json
done
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
echo done
