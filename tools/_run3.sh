for dbg in 0 8 16 24; do XW_RENDER_DEBUG=$dbg python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,1],[8,128,0,0,"sp",0,1]]' ; done 2>&1 | tee gpurun_out/s3_debug_sp2.txt
XW_RENDER_GROUPS=8 XW_RENDER_GROUP_THREADS=64 XW_RENDER_SP_FILL=1 ncu --set full --import-source on --clock-control none -k regex:k_render_sp -c 1 -s 3 -o gpurun_out/s3_sp_a python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,1]]' > gpurun_out/s3_ncu_sp_a.log 2>&1
tail -3 gpurun_out/s3_ncu_sp_a.log
