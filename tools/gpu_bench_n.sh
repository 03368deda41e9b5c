# usage: bash tools/gpu_bench_n.sh N tag   (run under gpurun --gpus N)
cd /root/repo
N=$1; TAG=${2:-r2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo rc=$?
grep -v "Warning\|warn\|run_backward\|^\*\*\*" gpurun_out/bench_${TAG}_n$N.err | tail -5
cut -c1-220 gpurun_out/bench_${TAG}_n$N.json
