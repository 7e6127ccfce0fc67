#!/usr/bin/env bash
mkdir -p gpurun_out
echo "fused: $(python tools/value_only.py 2>&1 | tail -1)"
echo "unfused: $(ICSP_ME_FUSED=0 python tools/value_only.py 2>&1 | tail -1)"
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kernel_timeline.py 2> gpurun_out/kt2.txt >/dev/null
