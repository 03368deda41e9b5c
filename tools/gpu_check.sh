set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -3 gpurun_out/bench_r1f.err; cat gpurun_out/bench_r1f.json
