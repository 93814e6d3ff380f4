timeout 200 python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sb"],[8,64,0,0,"sp",0,0],[9,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp9.txt
timeout 60 python tools/sweep_render.py c2 65536 '[[8,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp9_c2.txt
timeout 60 python tools/sweep_render.py c4 32768 '[[2,256,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp9_c4.txt
