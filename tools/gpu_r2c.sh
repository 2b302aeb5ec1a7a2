#!/bin/bash
# Round-2 sweep: late-load variant of the warp-strip plane kernel, warp-strip plane kernel vs tiled kernel on short slabs
mkdir -p gpurun_out
O=gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "tiling_invariance" > $O/r2c_tests.log 2>&1
el "tiling tests rc=$? : $(tail -1 $O/r2c_tests.log)"
run() { timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --configs none "$@" 2> $O/r2c_bench_err.log | tail -1 > $O/r2c_line.json; python -c "
import json,sys
try:
    d=json.load(open('$O/r2c_line.json')); print('$*', '| us', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_max'), 'e2e ms', round(d['e2e']['ms_per_call'],2), 'fma us', round(d.get('fma_mode',{}).get('kernel_us',0),2), 'tile', d['config']['tile'])
except Exception as e: print('$*', 'FAILED', e, open('$O/r2c_bench_err.log').read()[-400:])"; }
{
run --variant 0
run --variant 1024
run --variant 525312
run --workload om025@1440x270
run --workload om025@1440x270 --tile-threads 128 --variant 32768
run --workload om025@1440x270 --tile-threads 128 --variant 33792
run --workload om025@1440x135
run --workload om025@1440x135 --tile-threads 128 --variant 32768
run --workload om025@1440x135 --tile-threads 128 --variant 33792
run --workload gx1
run --workload gx1 --tile-threads 128 --variant 32768
run --workload om1deg
run --workload om1deg --tile-threads 128 --variant 32768
run --realistic
run --realistic --variant 2048
} > $O/r2c_sweep.txt 2>&1
el "sweep done"; cat $O/r2c_sweep.txt
