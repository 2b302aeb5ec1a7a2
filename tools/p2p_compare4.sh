TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { $TR --master-port $1 bench.py --gpus 4 --steps 5 --warmup 3 --configs none --no-cpu-baseline "${@:2}" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('WBND=$EVP_B200_WBND ${*:2}', '| us/subcycle', round(1e3*d['ms_per_step']/120,2), 'parity', d.get('parity_vs_1gpu','')[:9], 'W', d['clocks']['power_w_max'])
except Exception as e: print('${*:2}', 'FAILED', e)"; }
run 29571 --workload om025@1440x540
run 29572 --workload om025@1440x540 --variant 32768
EVP_B200_WBND=0.3 run 29573 --workload om025@1440x540
EVP_B200_WBND=1.0 run 29574 --workload om025@1440x540
EVP_B200_WBND=0.3 run 29575 --workload om025@1440x540 --variant 32768
