XW_RENDER_MODE=sp timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1],[9,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp7.txt
for dbg in 8 16 24; do XW_RENDER_DEBUG=$dbg timeout 60 python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' ; done 2>&1 | tee gpurun_out/s3_debug_sp10.txt
timeout 60 python tools/sweep_render.py c2 65536 '[[8,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp7_c2.txt
timeout 60 python tools/sweep_render.py c4 32768 '[[2,256,0,0,"sp",0,0],[2,128,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp7_c4.txt
