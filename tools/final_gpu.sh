#!/bin/bash
# End-of-round GPU pass (one gpurun call): full GPU test suite, the default bench line, every workload on its own, the CPU
# reference arm, the ncu launch list of the default command and one `ncu --set full` capture per hot kernel.  Everything lands
# in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
#   usage on the GPU box:  tools/final_gpu.sh <tag>
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r02z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2>> gpurun_out/${T}_bench.err
for w in c2 c4 fpv c5; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 20 > gpurun_out/${T}_bench_$w.json 2>> gpurun_out/${T}_bench.err
done
# launch list of the default command (per-launch durations: cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
tools/final_ncu.sh $T
python - <<P
import json
T="$T"
d=json.loads(open("gpurun_out/%s_bench.json" % T).read().strip().splitlines()[-1])
r=d["roofline"]; e=d["e2e"]
print("c3 value %.1fM step %.4f render %.4f frac %.3f e2e %.1fM pipe %.1fM launches %d" % (d["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"], e["value"]/1e6, e["pipelined"]["value"]/1e6, d["gpu_launches"]))
for k,v in d.get("configs",{}).items():
    print(k, "value %.4gM step %.4f frac %.3f e2e %.4gM" % (v["value"]/1e6, v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["value"]/1e6))
P
