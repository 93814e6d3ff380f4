timeout 250 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 250 python bench.py --steps 100 --warmup 10 > gpurun_out/r01m_bench.json 2> gpurun_out/r01m_bench.err; tail -c 600 gpurun_out/r01m_bench.err
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01m_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01m_ncu_bench.log 2>&1
timeout 250 ncu --set full --import-source on --clock-control none -k regex:k_render_sp -c 1 -s 3 -o gpurun_out/r01m_k_render_sp python tools/sweep_render.py c3 65536 '[[8,64,0,0,"sp",0,0]]' > gpurun_out/r01m_ncu.log 2>&1
tail -2 gpurun_out/r01m_ncu.log
cat gpurun_out/r01m_bench.json
