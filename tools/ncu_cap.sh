#!/bin/bash
# One `ncu --set full` capture of one launch of a kernel of a bench workload, exported as CSV into gpurun_out/ (raw page + SASS
# source page); usage on the GPU box:  tools/ncu_cap.sh <tag> <kernel regex> <workload> [skip launches] [launches to capture]
# (the render kernels run twice per step -- all envs, then the auto-reset queue's envs again: capture two launches and read the long one)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-6} -c ${5:-1} -o /tmp/rep/$1 \
   python bench.py --workload $3 --steps 4 --warmup 4 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/$1_ncu.log 2>&1
ncu -i /tmp/rep/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i /tmp/rep/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1_source.csv.gz
