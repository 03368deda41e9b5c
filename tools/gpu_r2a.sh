# round 2, first GPU pass: smoke, GPU tests (incl. the new full-batch cases), default bench, gradient-error table
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2a.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2a.err; cut -c1-300 gpurun_out/bench_r2a.json
timeout 900 python tools/grad_error_by_leaf.py > gpurun_out/grad_by_leaf_r2a.log 2>&1; echo "leaf rc=$?"; tail -3 gpurun_out/grad_by_leaf_r2a.log
