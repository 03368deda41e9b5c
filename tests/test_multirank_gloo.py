"""world_size-2 gloo test of the N>1 host logic (unsupervised_pose_estimation_b200/parallel.py):
batch sharding, loss all-reduce and the bucketed gradient all-reduce.  The per-rank loss is computed
with the oracle on CPU (test infrastructure) — what is under test is the sharding / reduction logic,
which is device independent."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden, golden_cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import vsl_oracle as O
    from unsupervised_pose_estimation_b200 import parallel

    g = load_golden(case)
    opt = g["opt"]
    B = opt.batch_size
    inputs = parallel.shard_batch(g["inputs"], rank, world)
    leaves = {k: v.clone().requires_grad_(True) for k, v in parallel.shard_batch(g["leaves"], rank, world).items()}
    noise = [parallel.shard_batch({"z": z}, rank, world)["z"] for z in g["noise"]]
    lo, hi = parallel.shard_range(B, rank, world)
    opt.batch_size = hi - lo  # modules are built with the LOCAL batch (layers.py:225-232)
    outputs = dict(leaves)
    for f in opt.frame_ids[1:]:
        if f != "s":
            outputs[("cam_T_cam", 0, f)] = O.transformation_from_parameters(
                leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
    losses = O.loss_step(opt, inputs, outputs, noise)
    glob = parallel.all_reduce_losses(losses, hi - lo)
    # a stand-in "network": one shared parameter feeding every image, so its gradient needs the all-reduce
    w = torch.nn.Parameter(torch.ones(3))
    proxy = sum((leaves[("disp", s)] * w[0]).sum() for s in opt.scales) * 0 + losses["loss"] * w.sum()
    proxy.backward()
    n_buckets = parallel.all_reduce_grads([w], bucket_bytes=8, local_batch=hi - lo, global_batch=B)
    if rank == 0:
        ret["losses"] = {k: v.item() for k, v in glob.items()}
        ret["w_grad"] = w.grad.tolist()
        ret["buckets"] = n_buckets
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_losses_and_grad_allreduce():
    case = [c for c in golden_cases() if c.startswith("mono_iid")][0]
    g = load_golden(case)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, ret), nprocs=2, join=True)
    # per-image means averaged over equal shards == the global-batch losses of the golden run
    # (min_loss is a batch mean; the smoothness term is a mean over images as well)
    for k, ref in g["losses"].items():
        assert ret["losses"][k] == pytest.approx(ref.item(), rel=2e-6), k
    mean_loss = ret["losses"]["loss"]
    assert ret["w_grad"] == pytest.approx([mean_loss] * 3, rel=1e-5)
    assert ret["buckets"] == 1


def _bucket_worker(rank, world, port, ret):
    """GradBuckets (hooks + flat views + overlapped bucket all-reduce) against the definition: the gradient of
    the global-batch mean, with UNEVEN shards (5 images over 2 ranks) and a parameter that never gets a grad."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from unsupervised_pose_estimation_b200 import parallel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                              torch.nn.Linear(16, 1))
    unused = torch.nn.Parameter(torch.ones(7))           # like torchvision's fc: never reached by backward
    params = list(net.parameters()) + [unused]
    gen = torch.Generator().manual_seed(1)
    X, Y = torch.randn(5, 6, generator=gen), torch.randn(5, 1, generator=gen)
    lo, hi = parallel.shard_range(5, rank, world)
    buckets = parallel.GradBuckets(params, local_batch=hi - lo, global_batch=5, bucket_bytes=256)
    per_step = []
    for step in range(2):   # second step: the views must have survived and been re-zeroed
        buckets.begin_step()
        loss = ((net(X[lo:hi]) - Y[lo:hi]) ** 2).mean() * (1.0 + step)
        loss.backward()
        n = buckets.finish()
        per_step.append([p.grad.clone() for p in params])
        assert all(p.grad.data_ptr() == buckets.grad_view(p).data_ptr() for p in params)
    if rank == 0:
        full = ((net(X) - Y) ** 2).mean()
        ref = torch.autograd.grad(full, list(net.parameters()))
        ret["err"] = max(float((g - r).abs().max() / r.abs().max()) for g, r in zip(per_step[0], ref))
        ret["err2"] = max(float((g - 2.0 * r).abs().max() / r.abs().max()) for g, r in zip(per_step[1], ref))
        ret["unused"] = float(per_step[0][-1].abs().max())
        ret["buckets"] = n
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_grad_buckets_overlapped_allreduce_uneven_shards():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["err"] <= 1e-6 and ret["err2"] <= 1e-6
    assert ret["unused"] == 0.0
    assert ret["buckets"] >= 3   # 256-byte buckets over ~1.6 KB of parameters: several exchanges, all the same on both ranks


def test_grad_buckets_single_process_is_plain_backward():
    from unsupervised_pose_estimation_b200 import parallel
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 3)
    x = torch.randn(5, 4)
    ref = torch.autograd.grad(net(x).pow(2).sum(), list(net.parameters()))
    b = parallel.GradBuckets(net.parameters(), bucket_bytes=16)
    for _ in range(2):
        b.begin_step()
        net(x).pow(2).sum().backward()
        b.finish()
        for p, r in zip(net.parameters(), ref):
            assert torch.equal(p.grad, r)
    # zero_grad(set_to_none=True) between begin_step and backward must not lose the gradient
    b.begin_step()
    for p in net.parameters():
        p.grad = None
    net(x).pow(2).sum().backward()
    b.finish()
    for p, r in zip(net.parameters(), ref):
        assert torch.equal(p.grad, r) and p.grad.data_ptr() == b.grad_view(p).data_ptr()


def test_shard_range_covers_batch():
    from unsupervised_pose_estimation_b200 import parallel
    for B in (1, 7, 12, 96):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
