TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { $TR --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --configs none --no-cpu-baseline "${@:2}" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('${*:2}', '| us/subcycle', round(1e3*d['ms_per_step']/d['config'].get('ndte',120),2) if 'ndte' in d['config'] else round(1e3*d['ms_per_step']/120,2), 'parity', d.get('parity_vs_1gpu'))
except Exception as e: print('${*:2}', 'FAILED', e)"; }
run 29571 --workload om025@1440x270
run 29572 --workload om025@1440x270 --variant 32768
run 29573 --workload om025@1440x270 --variant 32768 --tile-threads 128
run 29574 --workload om025@1440x540
run 29575 --workload om025@1440x540 --variant 32768
