#!/bin/bash
# quick first-person-view check on the GPU box: parity tests, then the bench workload (prints a short summary)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${FPV_TESTS:-fpv_golden or fpv_full or fpv_auto}" 2>&1 | tail -3
for c in ${FPV_CTAS:-0}; do
  XW_FPV_CTAS_PER_SM=$c python bench.py --workload fpv --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/fpv_$c.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ctas/sm $c fpv value %.1fM step %.4f render %.4f frac %.3f step_reset %.4f kernel %s e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], r['step_reset_ms'], r['kernel'], d['e2e']['value']/1e6))"
done
