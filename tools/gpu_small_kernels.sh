# usage: bash tools/gpu_small_kernels.sh <tag>   launch list + one full ncu capture per product kernel (2nd invocation each)
cd /root/repo
TAG=${1:-r2x}
timeout 300 python tools/run_small_kernels.py > gpurun_out/small_plain_$TAG.log 2>&1 || { tail -5 gpurun_out/small_plain_$TAG.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' --csv --log-file gpurun_out/small_launches_$TAG.csv \
  python tools/run_small_kernels.py > gpurun_out/small_ncu1_$TAG.log 2>&1; echo "launch list rc=$?"
ITERS=2 timeout 900 ncu --set full --clock-control none --kernel-id ::regex:'k_':2 -f -o /tmp/prof_small_$TAG \
  python tools/run_small_kernels.py > gpurun_out/small_ncu2_$TAG.log 2>&1; echo "full rc=$?"
ncu -i /tmp/prof_small_$TAG.ncu-rep --page raw --csv > gpurun_out/small_raw_$TAG.csv 2>/dev/null
ls -la /tmp/prof_small_$TAG.ncu-rep gpurun_out/small_raw_$TAG.csv
