#!/bin/bash
# Strong-scaling curves from ONE multi-GPU lease: bench.py for every BASELINE config at the GPU counts the box has.
#   tools/scale_run.sh [configs...]      default: 2 3 4 5 5e 5c     env: NS="2 4 8" (N = 1 comes from the single-GPU runs)
# Writes gpurun_out/scale_cfg<config>_n<N>.json (one bench line each).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
CONFIGS=${@:-2 3 4 5 5e 5c}
NS=${NS:-2 4 8}
port=29500
for cfg in $CONFIGS; do
  for n in $NS; do
    [ "$n" -le "$NG" ] || continue
    port=$((port+1))
    if [ "$n" -eq 1 ]; then
      python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > gpurun_out/scale_cfg${cfg}_n$n.json
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_err.log | grep '^{' > gpurun_out/scale_cfg${cfg}_n$n.json
    fi
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_cfg${cfg}_n$n.json"))
    sp = d.get("single_process_multi_gpu", {})
    print("cfg$cfg N=$n value %.0f e2e %.0f ms %.2f kernel_ms %.2f frac %s one-process identical=%s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], sp.get("bit_identical_to_gathered")))
except Exception as e:
    print("cfg$cfg N=$n FAILED", e); print(open("gpurun_out/scale_err.log").read()[-1500:])
PY
  done
done
