#!/bin/bash
# r72 (2 GPUs): bench at N=2 with the Python fault handler on (r71: rank 0 died with SIGSEGV)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r72; mkdir -p $O
export PYTHONFAULTHANDLER=1
timeout 900 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $O/bench_2gpu.json 2>$O/bench_2gpu.err
grep -n "Fatal\|File \"/root\|File \"/tmp\|bench.py\|lib.py\|comm.py" $O/bench_2gpu.err | head -30
tail -c 600 $O/bench_2gpu.json
