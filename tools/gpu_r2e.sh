#!/bin/bash
# 2 GPUs: the warp-strip plane kernel with the peer-to-peer halo (bit-exact vs the oracle), then bench.py --gpus 2
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29561 tests/multigpu_parity.py --realistic --kernel-variant 32768 --tile-threads 128 2>&1 | grep -E "multigpu_parity|MISMATCH|Error" | cut -c1-400
timeout 100 $TR --master-port 29562 tests/multigpu_parity.py --case om1deg --nx 300 --ny 200 --steps 1 --ndte 60 --kernel-variant 32768 --tile-threads 128 2>&1 | grep -E "multigpu_parity|MISMATCH|Error" | cut -c1-400
timeout 150 $TR --master-port 29563 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --configs none 2> $O/r2e_err.log | tail -1 > $O/r2e_bench2.json
python -c "
import json
d=json.load(open('$O/r2e_bench2.json')); print('2 GPUs | us', round(d['roofline']['kernel_us'],2), 'frac/GPU', round(d['roofline']['frac'],3), 'parity', d.get('parity_vs_1gpu'), 'e2e ms', round(d['e2e']['ms_per_call'],2), 'fma us', round(d.get('fma_mode',{}).get('kernel_us',0),2), d['config']['parallelism'][:60])" || tail -5 $O/r2e_err.log
