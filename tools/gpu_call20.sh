#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "batch or two_gpus" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -n 3 gpurun_out/r02_bench_2gpu.err
python tools/pcie_probe_multi.py 1 2>&1 | tail -3; python tools/pcie_probe_multi.py 2 2>&1 | tail -3
