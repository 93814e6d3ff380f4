timeout 200 python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sb"],[8,64,0,0,"sp",0,0],[8,96,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp11.txt
timeout 60 python tools/sweep_render.py c2 65536 '[[8,64,0,0,"sb"],[8,96,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp11_c2.txt
