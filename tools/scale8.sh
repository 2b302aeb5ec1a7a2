#!/bin/bash
# 8-GPU checks: bit-exact parity against the oracle, then bench.py variants (one JSON line each)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29551 tests/multigpu_parity.py --case om025 --nx 1440 --ny 1080 --steps 1 --ndte 24 2>&1 | grep multigpu_parity
run() { $TR --master-port $1 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline "${@:2}" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('${*:2}', '| value', '%.3e' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'us/subcycle', round(1e3*d['ms_per_step']/120,2), 'frac/GPU', round(d['roofline']['frac'],3), 'e2e ms', round(d['e2e'].get('ms_per_call',0),2))
except Exception as e: print('${*:2}', 'FAILED', e)"; }
run 29552
run 29553 --math-mode 1
run 29554 --tile-threads 128
run 29555 --workload p01
