#!/bin/bash
# Quick N-GPU check of the three multi-GPU workloads, most informative first (row strips + NCCL halos, pairs, batch):
#   bash tools/multi_gpu_quick.sh 2      -> gpurun_out/mg_c{5,3,4}_n2.json
N=$1
mkdir -p gpurun_out
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" --no-cpu-baseline 2> gpurun_out/$2.err | grep '^{' > gpurun_out/$2.json; python -c "
import json; d=json.load(open('gpurun_out/$2.json')); print('$2', round(d['value']/1e9,1), 'Gpxd/s', round(d['ms_per_step'],3), 'ms/step', round(d['fps'],1), 'fps', d['scaling'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
run 29611 mg_c5_n$N --steps 3 --warmup 3 --workload c5
run 29612 mg_c3_n$N --steps 10 --warmup 3
run 29613 mg_c4_n$N --steps 3 --warmup 3 --workload c4
