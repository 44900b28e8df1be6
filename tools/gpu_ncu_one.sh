# one full ncu capture of the fused replay: gpu_ncu_one.sh <tag> <prof_mc args...>
tag=$1; shift
python tools/prof_mc.py 262144 "$@" 2>&1 | tail -1
QEKF_DIAG=1 python tools/prof_mc.py 262144 "$@" 2>&1 | grep diag
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -f -o gpurun_out/prof_$tag python tools/prof_mc.py 262144 "$@" > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
