# the probe set used while tuning (CUDA-event times of un-profiled launches, 262,144 filters)
for a in "n=1200" "n=1200 nostats" "n=1200 nodrop" "mr n=600" "mr dyn n=600" "fp32 n=1200"; do python tools/prof_mc.py 262144 $a 2>&1 | tail -1; done
