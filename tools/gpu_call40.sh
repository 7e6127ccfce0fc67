#!/usr/bin/env bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encode_matches or dc_chain or launch_shapes or copy_modes" 2>&1 | tail -2
