# on the GPU box (via gpurun): GPU test suite, then the default benchmark
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_last.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_last.json").read().strip().splitlines()[-1])
keep = {k: d[k] for k in ("value", "ms_per_step") if k in d}
keep["e2e"] = {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}
if "e2e_f32_host_tensors" in d: keep["e2e_f32"] = d["e2e_f32_host_tensors"]["ms_per_step"]
keep["kernel_ms"] = d["roofline"]["kernel_ms"]; keep["frac"] = d["roofline"]["frac"]
keep["issue_frac"] = d.get("roofline_issue", {}).get("frac")
keep["clocks"] = d.get("clocks")
print(json.dumps(keep))
PY
