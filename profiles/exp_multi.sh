#!/bin/bash
# round 2 multi-GPU evidence (N = $1): the driver's bench command under torchrun, and the multi-rank parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
  > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python profiles/show_bench.py gpurun_out/r2_bench_${N}gpu.json
if [ "$N" = "4" ] || [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2_multirank_${N}gpu.log 2>&1
  tail -3 gpurun_out/r2_multirank_${N}gpu.log
fi
