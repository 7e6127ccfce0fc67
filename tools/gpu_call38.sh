#!/usr/bin/env bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-120
