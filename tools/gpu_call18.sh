#!/usr/bin/env bash
python tools/kernel_timeline.py 2> gpurun_out/kt.txt >/dev/null; grep -c "icsp kt" gpurun_out/kt.txt
