cd /root/repo
for sk in pipeline h2d pipeline,h2d pipeline,h2d,readback; do
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-skip $sk > gpurun_out/diag.json 2> gpurun_out/diag.err || { tail -5 gpurun_out/diag.err; continue; }
  python -c "
import json; d=json.loads(open('gpurun_out/diag.json').read().strip().splitlines()[-1]); print('$sk', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
