#!/bin/bash
# Mesh configurations on all GPUs of the box: BASELINE configs 3 and 4 at full size, and the config-5 scene with a mesh.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for sc in teapot gopher; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $NG --steps 3 --warmup 3 --scene $sc --aperture 0 --focal-length 0 --no-cpu-baseline 2>/dev/null | grep '^{' | tee gpurun_out/${sc}_$NG.json | cut -c1-220
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $NG --steps 2 --warmup 3 --scene cubemap --width 3840 --height 2160 --samples 4096 --aperture 0 --focal-length 0 --no-cpu-baseline 2>/dev/null | grep '^{' | tee gpurun_out/cfg5_cubemap_$NG.json | cut -c1-220
