for dbg in 0 8 32 64; do XW_RENDER_DEBUG=$dbg timeout 100 python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' ; done 2>&1 | tee gpurun_out/s3_prof_sp2.txt
