"""N >= 2: loss step (graph) + 114.6 MB gradient all-reduce issued concurrently, for several NCCL CTA limits.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/probe_comm_overlap.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unsupervised_pose_estimation_b200 import synthetic  # noqa: E402
from unsupervised_pose_estimation_b200.graph import GraphedLossStep  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False
cfg = dict(synthetic.CONFIGS["C1"])
wl = bench.Workload(cfg, "smooth", dev, 4)
for i in range(3):
    wl.step(wl.sets[i % 4])
graphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"]) for st in wl.sets]


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


payload = torch.zeros(28641888, device=dev)
bucket = (32 << 20) // 4
buckets = [payload[i:i + bucket] for i in range(0, payload.numel(), bucket)]
loss_ms = bench.timed_loop(lambda i: graphs[i % 4].replay(), 100, barrier, dev, dist)
out = {"loss_ms": loss_ms}
for max_ctas in (None, 16, 8):
    for nb in (4, 1):
        prio = 0
        bucket = (payload.numel() + nb - 1) // nb
        buckets = [payload[i:i + bucket] for i in range(0, payload.numel(), bucket)]
        if max_ctas is None:
            group = None
        else:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = min(max_ctas, 4)
            group = dist.new_group(backend="nccl", pg_options=opts)
        comm = torch.cuda.Stream(device=dev, priority=prio)

        def allreduce_all():
            return [dist.all_reduce(b, async_op=True, group=group) for b in buckets]

        def alone(i):
            with torch.cuda.stream(comm):
                works = allreduce_all()
            for w in works:
                w.wait()

        def overlapped(i):
            ev = torch.cuda.Event()
            ev.record()
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                works = allreduce_all()
            graphs[i % 4].replay()
            for w in works:
                w.wait()
        for i in range(3):
            alone(i)
        a = bench.timed_loop(alone, 20, barrier, dev, dist)
        for i in range(3):
            overlapped(i)
        o = bench.timed_loop(overlapped, 100, barrier, dev, dist)
        prio = nb
        out["max_ctas=%s prio=%d" % (max_ctas, prio)] = {"allreduce_alone_ms": round(a, 4), "overlapped_ms": round(o, 4),
                                                        "hidden": round((loss_ms + a - o) / a, 3)}
        if rank == 0:
            print(max_ctas, prio, out["max_ctas=%s prio=%d" % (max_ctas, prio)], flush=True)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
