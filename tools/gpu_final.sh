# on the GPU box: the driver's round-end sequence for N=1 (reference arm, then ours), outputs under gpurun_out/
cd /root/repo
TAG=${1:-r1x}
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
