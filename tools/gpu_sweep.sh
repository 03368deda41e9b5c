# on the GPU box: every BASELINE config on one GPU (100 steps), one JSON line each under gpurun_out/sweep_<tag>.jsonl
cd /root/repo
TAG=${1:-r1x}
OUT=gpurun_out/sweep_$TAG.jsonl
: > $OUT
run() { echo "== $*"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline "$@" >> $OUT 2>> gpurun_out/sweep_$TAG.err || echo "FAILED $*"; }
run --config C1
run --config C1 --family iid
run --config C2
run --config C3
run --config C3 --bf16-images
run --config C4
run --config C5
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(d["config"]["workload"][:70], "| value %.3g ms %.3f | e2e %.3g (%.1f MB) | e2e_f32 %.3g | k %.3f ms" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"] / 1e6,
        d["e2e_f32_host_tensors"]["value"], d["roofline"]["kernel_ms"]))
PY
