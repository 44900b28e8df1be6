python -m pytest tests/test_gpu_checkpoint.py -x -q 2>&1 | tail -3
python tools/prof_mc.py 262144 n=1200 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_07_sr python tools/prof_mc.py 262144 n=1200 > gpurun_out/ncu_sr.log 2>&1; tail -1 gpurun_out/ncu_sr.log
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_07_mrdyn python tools/prof_mc.py 262144 mr dyn n=600 > gpurun_out/ncu_mr.log 2>&1; tail -1 gpurun_out/ncu_mr.log
