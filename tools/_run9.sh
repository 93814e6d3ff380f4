for dbg in 0 8 16 32 64; do XW_RENDER_DEBUG=$dbg python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' ; done 2>&1 | tee gpurun_out/s3_debug_sp9.txt
