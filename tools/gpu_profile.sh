# usage: bash tools/gpu_profile.sh <tag>   (on the GPU box, via gpurun)
set -x
cd /root/repo
TAG=${1:-r1x}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu1_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_photometric -s 3 -c 1 -f -o gpurun_out/prof_photometric_$TAG \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -5
