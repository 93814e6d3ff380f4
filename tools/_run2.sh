set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sb",0,0],[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1],[8,32,0,0,"sp",0,0],[8,32,0,0,"sp",0,1],[8,96,0,0,"sp",0,0],[8,96,0,0,"sp",0,1],[9,64,0,0,"sp",0,0],[9,64,0,0,"sp",0,1],[9,32,0,0,"sp",0,1],[8,128,0,0,"sp",0,0],[8,128,0,0,"sp",0,1]]' > gpurun_out/s3_sweep_sp.txt 2>&1
cat gpurun_out/s3_sweep_sp.txt
for dbg in 1 3; do XW_RENDER_DEBUG=$dbg python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1]]' ; done 2>&1 | tee gpurun_out/s3_debug_sp.txt
python tools/sweep_render.py c2 65536 '[[8,64,0,0,"sb",0,0],[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1]]' 2>&1 | tee gpurun_out/s3_sweep_sp_c2.txt
python tools/sweep_render.py c4 32768 '[[3,256,0,0,"sb",0,0],[3,256,0,0,"sp",0,0],[3,256,0,0,"sp",0,1],[3,128,0,0,"sp",0,1],[3,64,0,0,"sp",0,1]]' 2>&1 | tee gpurun_out/s3_sweep_sp_c4.txt
