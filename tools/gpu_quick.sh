# quick GPU check used while tuning: GPU tests, the 300-tick probes, the bench line without its side legs
python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
P=quadrotor_landing_b200/presets/rotors_sim.yaml
quadrotor_landing_b200/bin/qekf_replay --preset $P --seconds 60 --out gpurun_out/trace_mr.csv 2>&1 | grep latency
quadrotor_landing_b200/bin/qekf_replay --preset $P --seconds 60 --single-rate --out gpurun_out/trace_sr.csv 2>&1 | grep latency
rm -f gpurun_out/trace_*.csv
for a in "" "nostats" "mr" "mr dyn" "fp32"; do python tools/prof_mc.py 262144 $a 2>&1 | tail -1; done
python bench.py --no-legs --no-parity --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['frac'], d['clocks'])"
