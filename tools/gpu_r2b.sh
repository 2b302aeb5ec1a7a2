#!/bin/bash
# Round-2 validation call: new-variant parity tests first, the A/B of the plane kernel's strip variants inside
# bench.py's timed region, then the whole GPU suite, then one ncu capture.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 120 python -c "import torch; print(torch.cuda.get_device_name(0))" 2>&1 | tail -1
el "torch imported"
timeout 200 python -c "from cice4_b200 import build as B; import time; t=time.time(); print(B.build(), 'build', round(time.time()-t,1), 's')" 2>&1 | tail -2
el "library ready"
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu --durations=8 \
    -k "strip_and_finish or finish_epilogue or tiling_invariance or golden or full_size_vs_oracle" > $O/r2b_newtests.log 2>&1
el "new-variant tests rc=$? : $(tail -1 $O/r2b_newtests.log)"
run() { timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --configs none "$@" 2> $O/r2b_bench_err.log | tail -1 > $O/r2b_line.json; python -c "
import json,sys
try:
    d=json.load(open('$O/r2b_line.json')); print('$*', '| us', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_max'), 'e2e ms', round(d['e2e']['ms_per_call'],2), 'finish ms', d['e2e']['device_breakdown_ms_rank0'].get('finish_ms'))
except Exception as e: print('$*', 'FAILED', e, open('$O/r2b_bench_err.log').read()[-400:])"; }
{
run --variant 0
run --variant 2097152
run --variant 0
run --variant 2097152
run --variant 64
run --variant 4194304
run --workload om025@1440x540 --variant 32768
run --workload om025@1440x540 --variant 2129920
run --math-mode 1
} > $O/r2b_sweep.txt 2>&1
el "sweep done"; cat $O/r2b_sweep.txt
timeout 420 python -m pytest tests -x -q -m gpu --durations=12 > $O/r2b_gputests.log 2>&1
el "gpu suite rc=$? : $(tail -1 $O/r2b_gputests.log)"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
el "smoke"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --configs none"
$CMD > $O/r2b_plain.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_subcycle -s 300 -c 1 -f -o $O/r2b_warpx $CMD > $O/r2b_ncu.log 2>&1
el "ncu rc=$?"
