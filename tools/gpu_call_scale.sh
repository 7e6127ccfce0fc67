#!/usr/bin/env bash
# usage: gpu_call_scale.sh N   -> bench.py at N GPUs (torchrun), N-rank PCIe probe, product CLI batch at N GPUs
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -n 2 gpurun_out/r02_bench_${N}gpu.err
python tools/pcie_probe_multi.py $N > gpurun_out/r02_pcie_${N}.txt 2>&1; cat gpurun_out/r02_pcie_${N}.txt
python tools/cli_bench.py 64 $N > gpurun_out/r02_cli_bench_${N}gpu.json 2> gpurun_out/r02_cli_bench_${N}gpu.err; cat gpurun_out/r02_cli_bench_${N}gpu.json
