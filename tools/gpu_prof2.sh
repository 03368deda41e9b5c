# usage: bash tools/gpu_prof2.sh <tag> [kernel-regex]   full ncu capture of one kernel from a short eager bench run
cd /root/repo
TAG=${1:-r2x}; K=${2:-k_photometric}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-strong ${BENCH_ARGS} > gpurun_out/ncu2_$TAG.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -2
