set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 > gpurun_out/s3_bench0.json 2> gpurun_out/s3_bench0.err
for dbg in 0 1 3 7; do XW_RENDER_DEBUG=$dbg python tools/sweep_render.py c3 65536 '[[8,64,0]]' ; done > gpurun_out/s3_debug.txt 2>&1
cat gpurun_out/s3_debug.txt
tail -c 1500 gpurun_out/s3_bench0.json
