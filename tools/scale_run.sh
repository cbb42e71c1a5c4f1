#!/bin/bash
# Strong-scaling run of bench.py on 1/2/4/8 GPUs of one box (whatever the box has).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python bench.py --gpus 1 --steps 5 --warmup 3 | tee gpurun_out/scale_1.json | cut -c1-300
for n in 2 4 8; do
  if [ "$n" -le "$NG" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 2>/dev/null | grep '^{' | tee gpurun_out/scale_$n.json | cut -c1-300
  fi
done
[ -n "$SCALE_ONLY" ] && exit 0
# BASELINE config 5 family (3840x2160, 4096 spp) on all GPUs: textured planes/spheres, env sphere, env cube + mesh
for sc in textures envmap cubemap; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $NG --steps 2 --warmup 3 --scene $sc --width 3840 --height 2160 --samples 4096 --aperture 0 --focal-length 0 --no-cpu-baseline 2>/dev/null | grep '^{' | tee gpurun_out/cfg5_${sc}_$NG.json | cut -c1-300
done
# BASELINE configs 3 and 4 (mesh scenes, full 2048 spp) on all GPUs
for sc in teapot gopher; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $NG --steps 3 --warmup 3 --scene $sc --aperture 0 --focal-length 0 --no-cpu-baseline 2>/dev/null | grep '^{' | tee gpurun_out/${sc}_$NG.json | cut -c1-300
done
