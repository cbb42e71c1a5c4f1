#!/bin/bash
# ncu --set full captures of the trace kernel on the reference scene and the teapot (reduced spp), each only after
# the same command exited 0 without ncu.  usage: tools/r2_ncu.sh TAG [ref] [teapot] [gopher] [ref64]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
for W in "$@"; do
  case $W in
    ref) CMD="python tools/r2_time.py ref"; export SPP_SCALE=0.03125;;
    ref64) CMD="python tools/r2_time.py ref64"; export SPP_SCALE=0.0625;;
    teapot) CMD="python tools/r2_time.py teapot"; export SPP_SCALE=0.125;;
    gopher) CMD="python tools/r2_time.py gopher"; export SPP_SCALE=0.125;;
    cube) CMD="python tools/r2_time.py cube"; export SPP_SCALE=0.0625;;
    tex) CMD="python tools/r2_time.py tex"; export SPP_SCALE=0.0625;;
  esac
  $CMD > gpurun_out/plain_$W.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$W $CMD > gpurun_out/ncu_$W.log 2>&1
  tail -1 gpurun_out/plain_$W.log
done
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
