#!/bin/bash
# r71 (2 GPUs): the NCCL path after the round’s ABI changes: 2-GPU == 1-GPU test, bench at N=2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r71; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2>$O/bench_2gpu.err | tail -1 > $O/bench_2gpu.json
python -c "
import json; d=json.load(open('$O/bench_2gpu.json')); print('N=2 value %.4e e2e %.4e ms %.1f dev %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['device_ms_per_step']))"
tail -5 $O/bench_2gpu.err
