XW_RENDER_MODE=sp python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sb"],[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1],[9,64,0,0,"sp",0,0],[7,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp6.txt
for dbg in 1 8 16 24 32 64; do XW_RENDER_DEBUG=$dbg python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' ; done 2>&1 | tee gpurun_out/s3_debug_sp8.txt
python tools/sweep_render.py c2 65536 '[[8,64,0,0,"sb"],[8,64,0,0,"sp",0,0],[8,64,0,0,"sp",0,1]]' 2>&1 | tee gpurun_out/s3_sweep_sp6_c2.txt
python tools/sweep_render.py c4 32768 '[[3,256,0,0,"sb"],[2,256,0,0,"sp",0,0],[2,64,0,0,"sp",0,0]]' 2>&1 | tee gpurun_out/s3_sweep_sp6_c4.txt
XW_RENDER_GROUPS=8 XW_RENDER_GROUP_THREADS=64 ncu --set full --import-source on --clock-control none -k regex:k_render_sp -c 1 -s 3 -o gpurun_out/s3_sp_f python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' > gpurun_out/s3_ncu_sp_f.log 2>&1
tail -2 gpurun_out/s3_ncu_sp_f.log
