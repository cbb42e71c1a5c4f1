#!/bin/bash
# Round-end single-GPU evidence: GPU tests, the default bench line (fp64 + other configs inside), configs 3 / 4 with their
# own CPU baselines, the reference arm, then the ncu launch list of the bench command and full captures of the trace
# kernel per workload (each only after the same command exited 0 without ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for c in 3 4; do python bench.py --config $c --no-extras > gpurun_out/bench_cfg$c.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --samples 64"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches.log 2>&1
tools/r2_ncu.sh $TAG ref teapot gopher tex ref64 cube
