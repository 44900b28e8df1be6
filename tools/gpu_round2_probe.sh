set -x
P=quadrotor_landing_b200/presets/rotors_sim.yaml
quadrotor_landing_b200/bin/qekf_replay --preset $P --seconds 60 --out gpurun_out/trace_mr.csv 2>&1 | tail -2
quadrotor_landing_b200/bin/qekf_replay --preset $P --seconds 60 --single-rate --out gpurun_out/trace_sr.csv 2>&1 | tail -2
rm -f gpurun_out/trace_*.csv
tools/bin/dfma_probe > gpurun_out/r2_06_dfma_probe.log 2>&1; cat gpurun_out/r2_06_dfma_probe.log
python tools/prof_mc.py 262144 > gpurun_out/r2_06_prof_mc.log 2>&1; cat gpurun_out/r2_06_prof_mc.log
python tools/prof_mc.py 262144 nostats >> gpurun_out/r2_06_prof_mc.log 2>&1
python tools/prof_mc.py 262144 nodrop >> gpurun_out/r2_06_prof_mc.log 2>&1
python tools/prof_mc.py 262144 mr >> gpurun_out/r2_06_prof_mc.log 2>&1
python tools/prof_mc.py 262144 mr dyn >> gpurun_out/r2_06_prof_mc.log 2>&1
tail -4 gpurun_out/r2_06_prof_mc.log
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_06_sr python tools/prof_mc.py 262144 > gpurun_out/ncu_sr.log 2>&1; tail -2 gpurun_out/ncu_sr.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:run_kernel -c 2 --csv --log-file gpurun_out/r2_06_traffic_nostats.csv python bench.py --no-stats --no-legs --no-parity --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_traffic.log 2>&1; tail -3 gpurun_out/r2_06_traffic_nostats.csv
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:run_kernel -c 2 --csv --log-file gpurun_out/r2_06_traffic_stats.csv python bench.py --no-legs --no-parity --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_traffic2.log 2>&1; tail -3 gpurun_out/r2_06_traffic_stats.csv
