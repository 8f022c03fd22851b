#!/bin/bash
# r50 (8 GPUs): weak scaling through the library's NCCL path, value and e2e legs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r50; mkdir -p $O
nvidia-smi -L | wc -l | tee $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 2>$O/bench_8gpu.err | tail -1 > $O/bench_8gpu.json
python -c "
import json; d=json.load(open('$O/bench_8gpu.json')); print('N=8 value %.4e e2e %.4e ms %.1f dev %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['device_ms_per_step']))"
tail -3 $O/bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 2>$O/bench_4gpu.err | tail -1 > $O/bench_4gpu.json
python -c "
import json; d=json.load(open('$O/bench_4gpu.json')); print('N=4 value %.4e e2e %.4e ms %.1f dev %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['device_ms_per_step']))"
