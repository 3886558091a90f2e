N=$1; shift
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" 2> gpurun_out/$2.err | grep '^{' > gpurun_out/$2.json; python -c "
import json; d=json.load(open('gpurun_out/$2.json')); print('$2', round(d['value']/1e9,1), 'Gpxd/s', round(d['ms_per_step'],3), 'ms/step', round(d['fps'],1), 'fps', d['scaling'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
run 29601 bench_r1_c3_n$N --steps 20 --warmup 3
if [ "$1" == "all" ]; then
run 29602 bench_r1_c4_n$N --steps 3 --warmup 3 --workload c4
run 29603 bench_r1_c5_n$N --steps 3 --warmup 3 --workload c5
fi
