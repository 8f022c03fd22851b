#!/bin/bash
# r84 (2 GPUs): round-end sanity of the N > 1 path (NCCL inside the library) on the final tree
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r84; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2 | tee $O/gpu_multi_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs 2>$O/bench_2gpu.err | tail -1 > $O/bench_2gpu.json
python -c "
import json; d=json.load(open('$O/bench_2gpu.json')); print('N=2 value %.4e e2e %.4e ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']))"
