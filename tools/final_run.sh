#!/bin/bash
# Round-end single-GPU measurements: GPU tests, bench lines for every config at its full size, the reference arm,
# then the ncu launch list and full captures (each only after the same command exited 0 without ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --rng fast --no-cpu-baseline > gpurun_out/final_ref_fast.json 2>/dev/null
python bench.py --precision fp64 --samples 256 --no-cpu-baseline > gpurun_out/final_ref_fp64.json 2>/dev/null
python bench.py --scene teapot --aperture 0 --focal-length 0 > gpurun_out/final_teapot.json 2>/dev/null
python bench.py --scene gopher --aperture 0 --focal-length 0 > gpurun_out/final_gopher.json 2>/dev/null
python bench.py --scene teapot --aperture 0 --focal-length 0 --precision fp64 --samples 256 --no-cpu-baseline > gpurun_out/final_teapot_fp64.json 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_reference_arm.json 2>/dev/null
cat gpurun_out/final_*.json | cut -c1-260
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
REF="$B --samples 64"
GOP="$B --scene gopher --samples 8 --aperture 0 --focal-length 0"
TEA="$B --scene teapot --samples 8 --aperture 0 --focal-length 0"
$REF > gpurun_out/plain_ref.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ref_v9.csv $REF > gpurun_out/ncu_l.log 2>&1
$REF > gpurun_out/plain_ref.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_ref_fp32_v9 $REF > gpurun_out/ncu_ref.log 2>&1
$GOP > gpurun_out/plain_gop.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_gopher_v9 $GOP > gpurun_out/ncu_gop.log 2>&1
$TEA > gpurun_out/plain_tea.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_teapot_v9b $TEA > gpurun_out/ncu_tea.log 2>&1
ls -la gpurun_out/*v9*.ncu-rep gpurun_out/launches_ref_v9.csv
