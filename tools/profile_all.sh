#!/bin/bash
# Runs on the GPU box (one GPU): plain runs first, then the ncu launch list and full captures of the
# trace kernel on three workloads.  Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
REF="$B --samples 64"
TEA="$B --scene teapot --samples 4 --aperture 0 --focal-length 0"
GOP="$B --scene gopher --samples 4 --aperture 0 --focal-length 0"
F64="$B --samples 32 --precision fp64"
$REF > gpurun_out/plain_ref.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ref.csv $REF > gpurun_out/ncu_l.log 2>&1
$REF > gpurun_out/plain_ref.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_ref_fp32 $REF > gpurun_out/ncu_ref.log 2>&1
$TEA > gpurun_out/plain_tea.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_teapot_fp32 $TEA > gpurun_out/ncu_tea.log 2>&1
$GOP > gpurun_out/plain_gop.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_gopher_fp32 $GOP > gpurun_out/ncu_gop.log 2>&1
$F64 > gpurun_out/plain_f64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_ref_fp64 $F64 > gpurun_out/ncu_f64.log 2>&1
tail -1 gpurun_out/plain_ref.log gpurun_out/plain_tea.log gpurun_out/plain_gop.log gpurun_out/plain_f64.log | cut -c1-400
ls -la gpurun_out/*.ncu-rep
