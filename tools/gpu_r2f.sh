#!/bin/bash
# final check of the round: the driver's bench command, then as much of the GPU suite as the remaining budget allows
mkdir -p gpurun_out
O=gpurun_out
t0=$(date +%s)
timeout 215 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench.json 2> $O/r2f_bench_err.log
echo "[$(( $(date +%s) - t0 )) s] bench rc=$?"; tail -c 600 $O/r2f_bench_err.log
python -c "
import json
d=json.loads(open('$O/r2f_bench.json').read().strip().splitlines()[-1])
print('value %.4e us %.2f frac %.3f e2e %.4e (%.2f ms) cpu %s fma %s' % (d['value'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_call'], d['cpu_baseline'].get('value'), d.get('fma_mode',{}).get('kernel_us')))
for c in d.get('configs', []): print(c['workload'][:60], '| us', round(c['us_per_subcycle'],2), 'frac', round(c['roofline_frac_per_gpu'],3), 'e2e ms', round(c['e2e_ms_per_call'],2))
"
timeout 60 python -m pytest tests -x -q -m gpu -k "not full_size and not ieee and not strip_and_finish" > $O/r2f_tests.log 2>&1
echo "[$(( $(date +%s) - t0 )) s] tests rc=$? : $(tail -1 $O/r2f_tests.log)"
