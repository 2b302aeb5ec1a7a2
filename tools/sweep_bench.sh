#!/bin/bash
# bench.py on the headline workload with a list of kernel options (one line per run)
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$*', '| us', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'])
except Exception as e: print('$*', 'FAILED', e)"; }
while read -r line; do run $line; done
