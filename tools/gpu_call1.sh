#!/usr/bin/env bash
# round-2 GPU call 1: measured peaks, baseline of the round-1 build on this box, ME-segment / stream overlap sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
tools/microbench_peaks > gpurun_out/r02_peaks.json 2> gpurun_out/c1_peaks.err
cat gpurun_out/r02_peaks.json
echo "== baseline value"
python tools/value_only.py 2>&1 | tail -1
echo "== ME seg sweep under overlap (whole step, 4 streams)"
for seg in 22 11; do for st in 4 6 8; do
  ICSP_ME_SEG=$seg ICSP_STREAMS=$st python tools/value_only.py 2>&1 | tail -1 | sed "s/^/seg $seg /"
done; done
echo "== ME kernel alone per seg"
python tools/me_sweep.py 22 11 2>&1 | tail -3
