#!/usr/bin/env python
"""Benchmark of the view-synthesis loss path (BASELINE.json metric: target-pixels/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C1] [--family smooth]

One "step" = one pass of the hot path over one batch: generate_images_pred + compute_losses +
backward to the leaves (disp_0..3 and cam_T_cam per temporal frame) through the package's public API
(unsupervised_pose_estimation_b200.trainer.LossPath).  `value` is measured with the inputs resident
in HBM (a ring of input sets larger than L2 is cycled, so no step re-reads L2-warm inputs); `e2e`
copies every step's inputs from pinned host memory and reads the loss dict back.  The `roofline`
object times the dominant kernel (k_photometric) with CUDA events recorded by the library around its
launch, inside the timed region.  `cpu_baseline` / `--impl reference` time the oracle port of the
reference (oracle/vsl_oracle.py, test infrastructure) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from unsupervised_pose_estimation_b200 import synthetic  # noqa: E402

METRIC = "view_synthesis_loss_target_pixels_per_s_fwd_bwd"
UNIT = "px/s"


def algorithmic_bytes_per_pixel(num_src, e_img=4):
    """SURVEY.md §8d bytes_min per target pixel: target + F sources + target pyramid levels 1-3 +
    disp pyramid read + disp-grad pyramid write."""
    p = 1 + 0.25 + 1 / 16 + 1 / 64
    return e_img * (3 + 3 * num_src + 3 * (p - 1)) + 4 * p + 4 * p


def profiled_traffic_bytes():
    """dram read + write of k_photometric per launch from the newest committed `ncu --set full` summary
    (profiles/*_k_photometric.md, written by tools/summarize_profile.py); None if there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_k_photometric.md")))
    if not files:
        return None, None
    text = open(files[-1]).read()
    m = re.search(r"traffic = dram read \+ write\*\* \| ([0-9.]+) \| MB", text)
    mi = re.search(r"smsp__inst_executed.sum \| ([0-9.]+) \| inst", text)
    profiled_traffic_bytes.warp_instructions = float(mi.group(1)) if mi else None
    pct = {}
    for key, name in (("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
                      ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                      ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
                      ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
                      ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
                      ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")):
        mm = re.search(re.escape(key) + r" \| ([0-9.]+) \|", text)
        if mm:
            pct[name] = round(float(mm.group(1)), 2)
    profiled_traffic_bytes.ncu_pct = pct
    return (float(m.group(1)) * 1e6, os.path.basename(files[-1])) if m else (None, None)


def small_kernel_rooflines():
    """HBM fractions of the kernels around k_photometric (k_epilogue, k_combine, the input-pipeline and side-output
    kernels ...) from the newest committed small-kernel capture (profiles/*_small_kernels.json, written by
    tools/summarize_small_kernels.py from one `ncu --set full` launch each at C1): algorithmic bytes / captured time
    against the measured HBM peak.  Not timed in this run: they are microsecond kernels inside a CUDA graph."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_small_kernels.json")))
    if not files:
        return None
    data = json.load(open(files[-1]))
    return {"source": os.path.basename(files[-1]), "bound": "hbm", "unit": "fraction of the measured HBM peak",
            "kernels": {k: {"ncu_us": v["ncu_us"], "algorithmic_bytes": v["algorithmic_bytes"], "frac": v["frac_of_hbm_peak"]}
                        for k, v in data.items() if v.get("frac_of_hbm_peak") is not None}}


def peak_hbm_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(device_index):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE it
    allocates pinned host buffers, so first-touch places them on the GPU's NUMA node and the per-step H2D
    copies of 2/4/8 ranks do not all cross one socket's memory controllers.  Best effort: returns the NUMA
    node or None."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return int(open(base + "/numa_node").read())
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def count(self):
        """Samples written so far."""
        try:
            with open(self.tmp.name) as f:
                return sum(1 for _ in f)
        except OSError:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, reasons = [], set()
        for line in self.tmp.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                out["sm_max_mhz"] = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        os.unlink(self.tmp.name)
        return out


# the option values both arms run with (monodepth2's) -- passed explicitly to the GPU arm's
# make_opt AND to the oracle's, so the two can never drift apart through their defaults
BENCH_OPT = dict(max_depth=100.0, disparity_smoothness=1e-3, min_depth=0.1)


def oracle_step_cpu(cfg, family, batch, threads, device="cpu"):
    """One forward+backward of the reference path via the oracle port: on the host cores (cpu_baseline,
    --impl reference) or, with device="cuda", as the eager PyTorch-CUDA op stream (the `eager_cuda` record)."""
    from oracle import vsl_oracle as O
    if device == "cpu":
        torch.set_num_threads(threads)
    opt = O.make_opt(height=cfg["height"], width=cfg["width"], batch_size=batch, frame_ids=list(cfg["frame_ids"]),
                     **BENCH_OPT)
    inputs, outputs, leaves = synthetic.make_batch(batch, cfg["height"], cfg["width"], cfg["frame_ids"], cfg["K"],
                                                   seed=0, family=family, device=device)

    def run():
        out = dict(outputs)
        for f in cfg["frame_ids"][1:]:
            if f != "s":
                out[("cam_T_cam", 0, f)] = O.transformation_from_parameters(
                    leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        losses = O.loss_step(opt, inputs, out)
        torch.autograd.grad(losses["loss"], list(leaves.values()))
        if device != "cpu":
            torch.cuda.synchronize()
        return time.perf_counter() - t0
    return run


def run_reference(args, cfg, rank):
    """`--impl reference`: the reference's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = cfg["batch"]
    t_full = oracle_step_cpu(cfg, args.family, B, threads)()
    budget = 150.0
    b_s = max(1, min(B, int(B * budget / max(1e-6, (args.steps + args.warmup) * t_full))))
    run = oracle_step_cpu(cfg, args.family, b_s, threads)
    for _ in range(args.warmup):
        run()
    t = sum(run() for _ in range(args.steps))
    px = b_s * cfg["height"] * cfg["width"] * args.steps
    val = px / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d images per step (%dx%d, %d source frames, 4 scales), fwd+bwd, %d steps"
                                   % (b_s, B, cfg["width"], cfg["height"], len(cfg["frame_ids"]) - 1, args.steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, cfg):
    return {"workload": "%s: %dx%d, batch %d per GPU, frames %s, 4 scales, %s, %s synthetic frames"
                        % (args.config, cfg["width"], cfg["height"], cfg["batch"], cfg["frame_ids"],
                           "bf16 image storage + fp32 arithmetic" if getattr(args, "bf16_images", False) else "fp32",
                           args.family),
            "l2": "ring of input sets larger than L2 (126 MB) cycled between timed steps",
            "side_outputs": "none (fused path; reference side outputs are materialised on logging steps only)",
            "tie_break_noise": ("drawn inside every step, in front of its loss kernels" if getattr(args, "no_noise_prefetch", False)
                                or getattr(args, "no_graph", False) else
                                "software-pipelined: every step draws the NEXT step's four randn tensors (same draws, same "
                                "order as the reference) while its own loss kernels run; `noise_inline` is the step without")}


def shard_parity(cfg, family, device, rank, world, dist):
    """N > 1, before anything is timed: the path's claim that it shards by image with no exchange, checked on
    the hardware path.  Every rank runs its shard (rows [rank*B, (rank+1)*B) of a global batch of B*world images,
    same tie-break noise rows), rank 0 also runs the global batch on one GPU; required: the auto-masks of the shard
    rows are the same bits as the global run's, the all-reduced losses (parallel.all_reduce_losses) equal the
    global losses to 1e-6, and shard gradients / world equal the global gradient rows to 1e-5 rel-L2
    (reference semantics: batch means, trainer.py:672-685)."""
    from unsupervised_pose_estimation_b200 import functional as VF
    from unsupervised_pose_estimation_b200 import layers as L
    from unsupervised_pose_estimation_b200 import parallel
    B, H, W, frames = cfg["batch"], cfg["height"], cfg["width"], cfg["frame_ids"]
    Bg, F, S = B * world, len(cfg["frame_ids"]) - 1, 4
    inputs, _, leaves = synthetic.make_batch(Bg, H, W, frames, cfg["K"], seed=4242, family=family, device="cpu",
                                             requires_grad=False)
    gen = torch.Generator().manual_seed(99)
    noise_g = [torch.randn(Bg, F, H, W, generator=gen) for _ in range(S)]

    def run(lo, hi):
        b = hi - lo
        dv = lambda t: t[lo:hi].to(device)
        plan = VF.FusedLossPlan(b, H, W, list(range(S)), F, BENCH_OPT["min_depth"], BENCH_OPT["max_depth"],
                                BENCH_OPT["disparity_smoothness"], arith=VF.calibrate_arith(b, H, W, device))
        disps = [dv(leaves[("disp", s)]).requires_grad_(True) for s in range(S)]
        Ts = []
        for f in frames[1:]:
            if f == "s":
                Ts.append(dv(inputs["stereo_T"]))
            else:
                T = L.transformation_from_parameters(dv(leaves[("axisangle", 0, f)])[:, 0],
                                                     dv(leaves[("translation", 0, f)])[:, 0], f < 0)
                Ts.append(T.detach().requires_grad_(True))
        vec, masks = VF.fused_loss(plan, [dv(inputs[("color", 0, s)]) for s in range(S)],
                                   [dv(inputs[("color", f, 0)]) for f in frames[1:]], disps, dv(inputs[("inv_K", 0)]),
                                   None, [dv(z) for z in noise_g], K=dv(inputs[("K", 0)]), Ts=Ts)
        wrt = disps + [T for T in Ts if T.requires_grad]
        grads = torch.autograd.grad(vec[2 * S], wrt)
        return vec.detach(), torch.stack(masks), [g.flatten(1) for g in grads]

    lo, hi = rank * B, (rank + 1) * B
    vec_l, masks_l, grads_l = run(lo, hi)
    names = ["min_loss/%d" % s for s in range(S)] + ["loss/%d" % s for s in range(S)] + ["loss"]
    reduced = parallel.all_reduce_losses({k: vec_l[i] for i, k in enumerate(names)}, B)
    gshape = [(Bg, g.shape[1]) for g in grads_l]
    if rank == 0:
        vec_g, masks_g, grads_g = run(0, Bg)
    else:
        vec_g = torch.empty(2 * S + 1, device=device)
        masks_g = torch.empty(S, Bg, H, W, device=device)
        grads_g = [torch.empty(sh, device=device) for sh in gshape]
    for t in [vec_g, masks_g] + grads_g:
        dist.broadcast(t, 0)
    mask_ok = torch.equal(masks_l, masks_g[:, lo:hi])
    loss_err = max(abs(float(reduced[k]) - float(vec_g[i])) / abs(float(vec_g[i])) for i, k in enumerate(names))
    grad_err = max(float((gl / world - gg[lo:hi]).norm() / gg[lo:hi].norm().clamp_min(1e-30)) for gl, gg in zip(grads_l, grads_g))
    ok = torch.tensor([1.0 if (mask_ok and loss_err <= 1e-6 and grad_err <= 1e-5) else 0.0, loss_err, grad_err],
                      device=device, dtype=torch.float64)
    worst = ok.clone()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if float(ok[0]) != 1.0:
        raise SystemExit("shard parity FAILED on some rank: masks_equal=%s loss_err=%.3g grad_err=%.3g (rank %d)"
                         % (mask_ok, loss_err, grad_err, rank))
    del masks_g, grads_g
    torch.cuda.empty_cache()
    return {"status": "ok", "global_batch": Bg, "ranks": world, "automask_rows_bit_equal": True,
            "max_rel_loss_err_allreduced_vs_global": float(worst[1]), "max_rel_l2_grad_err_shard_vs_global_rows": float(worst[2]),
            "checked": "every rank's shard of a B*N-image batch against rank 0's single-GPU run of the whole batch"}


def timed_loop(fn, n, barrier, device, dist):
    """ms per call of fn(i) over n calls: CUDA events on the current stream, barrier + synchronize on both sides,
    max over ranks."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n



class Workload:
    """A ring of device-resident input sets + the public-API step."""

    def __init__(self, cfg, family, device, ring, pinned=False, bf16_images=False):
        from unsupervised_pose_estimation_b200 import layers as L
        from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt
        self.cfg, self.device, self.L = cfg, device, L
        self.opt = make_opt(height=cfg["height"], width=cfg["width"], batch_size=cfg["batch"],
                            frame_ids=list(cfg["frame_ids"]), **BENCH_OPT)
        self.path = LossPath(self.opt, device=device, side_outputs="none")
        self.sets = []
        self.host = []
        for r in range(ring):
            inputs, outputs, leaves = synthetic.make_batch(cfg["batch"], cfg["height"], cfg["width"], cfg["frame_ids"],
                                                           cfg["K"], seed=r, family=family, device="cpu",
                                                           requires_grad=False)
            # only what the path reads
            keep = {k: v for k, v in inputs.items()
                    if k == "stereo_T" or k[0] in ("K", "inv_K") and k[1] == 0
                    or (k[0] == "color" and (k[1] == 0 or k[2] == 0))}
            if bf16_images:
                keep = {k: (v.bfloat16() if isinstance(k, tuple) and k[0] == "color" else v) for k, v in keep.items()}
            # the path starts at outputs[("cam_T_cam",0,f)] (trainer.py:513): the pose network's
            # axis-angle -> 4x4 conversion (predict_poses, trainer.py:437-438) is upstream of it
            poses = {}
            for f in cfg["frame_ids"][1:]:
                if f != "s":
                    poses[("cam_T_cam", 0, f)] = L.transformation_from_parameters(
                        leaves.pop(("axisangle", 0, f))[:, 0], leaves.pop(("translation", 0, f))[:, 0], f < 0)
            leaves.update(poses)
            host = {"inputs": keep, "leaves": leaves}
            if pinned:
                host = {"inputs": {k: v.pin_memory() for k, v in keep.items()},
                        "leaves": {k: v.pin_memory() for k, v in leaves.items()}}
                self.host.append(host)
            if not pinned or r == 0:
                self.sets.append({"inputs": {k: v.to(device) for k, v in keep.items()},
                                  "leaves": {k: v.to(device).requires_grad_(True) for k, v in leaves.items()}})
        self.h2d_bytes = sum(v.numel() * v.element_size() for h in [host] for d in h.values() for v in d.values())

    def step(self, s):
        inputs, leaves = s["inputs"], s["leaves"]
        outputs = dict(leaves)  # ("disp", s) and ("cam_T_cam", 0, f)
        self.path.generate_images_pred(inputs, outputs)
        losses = self.path.compute_losses(inputs, outputs)
        grads = torch.autograd.grad(losses["loss"], list(leaves.values()))
        return losses, grads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C1", choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--family", default="smooth", choices=["smooth", "iid"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--e2e-skip", default="", help="diagnostic: comma list of pipeline,h2d,readback to leave out of the "
                    "e2e loop (the line is then marked invalid)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling C5 sub-record")
    ap.add_argument("--no-noise-prefetch", action="store_true",
                    help="draw each step's tie-break noise in front of its loss kernels (the faithful default of "
                         "compute_losses) instead of one step ahead, behind the previous step's loss kernels")
    ap.add_argument("--bf16-images", action="store_true",
                    help="store the colour images as bf16 (BASELINE config 3); arithmetic stays fp32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(synthetic.CONFIGS[args.config])

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = False

    from unsupervised_pose_estimation_b200 import functional as VF
    B, H, W = cfg["batch"], cfg["height"], cfg["width"]
    F = len(cfg["frame_ids"]) - 1
    n0 = B * H * W
    parity = shard_parity(cfg, args.family, device, rank, world, dist) if dist is not None else None
    ring = 4  # 4 x ~79 MB of inputs (+ 47 MB of fresh tie-break noise per step) > 126 MB L2
    wl = Workload(cfg, args.family, device, ring, bf16_images=args.bf16_images)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM --------------------------------------------------------
    from unsupervised_pose_estimation_b200.graph import GraphedLossStep
    for i in range(args.warmup):
        wl.step(wl.sets[i % ring])
    barrier()
    # (a) eager public-API loop, instrumented: k_photometric is timed by events the library records
    events = VF.KernelEvents()
    wl.path._vsl_plan().kernel_events = events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_eager = min(args.steps, 50)
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(n_eager):
        wl.step(wl.sets[i % ring])
    e1.record()
    t_enqueue = time.perf_counter() - t_wall0   # host time to enqueue the steps (no sync inside)
    barrier()
    eager_ms = e0.elapsed_time(e1) / n_eager
    kernel_ms = events.drain_ms()
    wl.path._vsl_plan().kernel_events = None
    # (b) the timed region: the same step captured once per input set and replayed (one host call per step)
    prefetch = not args.no_noise_prefetch and not args.no_graph
    if args.no_graph:
        run_step = lambda i: wl.step(wl.sets[i % ring])
    else:
        graphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"], noise_prefetch=prefetch) for st in wl.sets]
        run_step = lambda i: graphs[i % ring].replay()
    sampler = ClockSampler(local_rank) if rank == 0 else None   # nvidia-smi needs a moment to start: begin before
    for i in range(args.warmup):                                # the warm-up replays, same load as the timed region
        run_step(i)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        run_step(i)
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler is not None:
        # a short timed region (100 steps are 83 ms) can end before the 50 ms sampler has three readings: keep the
        # same load running, untimed, until it has (at most one more second)
        t_keep = time.perf_counter()
        while sampler.proc is not None and sampler.count() < 3 and time.perf_counter() - t_keep < 1.0:
            for i in range(20):
                run_step(i)
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n0 * args.steps / (ms_total * 1e-3)

    # ---- the un-pipelined step: every replay draws its own noise in front of its loss kernels ----------------
    noise_inline = None
    if not args.no_graph and prefetch:
        igraphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"]) for st in wl.sets]
        for i in range(8):
            igraphs[i % ring].replay()
        n_i = min(args.steps, 100)
        i_ms = timed_loop(lambda i: igraphs[i % ring].replay(), n_i, barrier, device, dist)
        noise_inline = {"value": world * n0 / (i_ms * 1e-3), "unit": UNIT, "ms_per_step": i_ms,
                        "note": "GraphedLossStep(noise_prefetch=False): the four randn launches of a step run alone "
                                "in front of its k_photometric; the headline draws them one step ahead (same draws, "
                                "same order), where they fill the SMs the loss kernel's last wave leaves idle"}
        del igraphs

    # ---- L2-warm variant (SURVEY.md 8d asks for both): one input set re-used every step ----------------
    l2_warm = None
    if not args.no_graph:
        for i in range(5):
            graphs[0].replay()
        barrier()
        n_w = min(args.steps, 100)
        e0.record()
        for i in range(n_w):
            graphs[0].replay()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ms = float(t.item()) / n_w
        l2_warm = {"value": world * n0 / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms,
                   "note": "the same input set every step (inputs, 79 MB, stay in the 126 MB L2; the 47 MB of noise "
                           "is fresh every step); the headline cycles four sets"}

    # ---- contract mode: the same step with every reference-visible side output materialised ---------
    # (outputs[("depth",0,s)], ("sample",f,s), ("color",f,s): +192 B per target pixel), written by
    # the fused kernel itself (vsl_side_outputs = "fused")
    contract = None
    if not args.no_graph:
        wl.path.vsl_side_outputs = "fused"
        cgraphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"], noise_prefetch=prefetch) for st in wl.sets[:2]]
        wl.path.vsl_side_outputs = "none"
        for i in range(4):
            cgraphs[i % 2].replay()
        barrier()
        n_c = min(args.steps, 100)
        e0.record()
        for i in range(n_c):
            cgraphs[i % 2].replay()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c_ms = float(t.item()) / n_c
        contract = {"side_outputs": "fused (depth, sampling grids, warped colours of every scale and frame written "
                                    "by k_photometric)", "value": world * n0 / (c_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": c_ms, "extra_bytes_per_step": n0 * (4 * 4 + 8 * 4 * F + 12 * 4 * F)}
        # SURVEY.md 8d "contract-mode" bytes: bytes_min + side outputs + the four auto-masks, against the step time
        c_bytes = algorithmic_bytes_per_pixel(F, 2 if args.bf16_images else 4) * n0 + contract["extra_bytes_per_step"] + 4 * 4 * n0
        contract["contract_bytes_per_step"] = c_bytes
        contract["hbm_frac_of_step"] = c_bytes / (c_ms * 1e-3) / 1e9 / peak_hbm_gbs()[0]
        del cgraphs

    # ---- e2e: host buffers in, loss dict out ---------------------------------------------------
    # Per step: H2D of the step's inputs from pinned host memory (copy stream, into the staging slot the
    # step's graph reads), the step, D2H read of the loss dict.  The copies of step i+1 overlap step i.
    # Two forms of the host buffers:
    #   "u8"  (headline): the level-0 frames as 8-bit HWC arrays, as the reference's dataset holds them before
    #         transforms.ToTensor(); the pyramid and the /255 conversion run on the GPU inside the step
    #         (input_pipeline.LossInputPipeline, bit-exact with Pillow / torchvision);
    #   "f32": the fp32 ("color", f, s) tensors the reference's DataLoader hands to process_batch.
    from unsupervised_pose_estimation_b200.staging import HostBatchStager
    from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline
    wl_h = Workload(cfg, args.family, device, ring, pinned=True, bf16_images=args.bf16_images)
    is_leaf = lambda k: isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")
    is_color = lambda k: isinstance(k, tuple) and k[0] == "color"
    img_dtype = torch.bfloat16 if args.bf16_images else torch.float32

    def host_batches_for(mode):
        out = []
        for h in wl_h.host:
            hb = dict(list(h["inputs"].items()) + list(h["leaves"].items()))
            if mode == "u8":
                frames = {k[1]: v for k, v in hb.items() if is_color(k) and k[2] == 0}
                hb = {k: v for k, v in hb.items() if not is_color(k)}
                for f, v in frames.items():   # 8-bit HWC, what np.asarray(pil_image) gives
                    hb[("color_u8", f)] = (v.float().permute(0, 2, 3, 1) * 255).round().clamp(0, 255) \
                        .to(torch.uint8).contiguous().pin_memory()
            out.append(hb)
        return out

    def run_e2e(mode):
        host_batches = host_batches_for(mode)
        stager = HostBatchStager(device, depth=2)
        slot_steps = {}

        slot_inputs, slot_pipes = {}, {}

        def slot_parts(dev):
            key = id(dev)
            if key not in slot_inputs:
                slot_inputs[key] = ({k: v for k, v in dev.items() if not is_leaf(k) and k[0] != "color_u8"},
                                    {k: v.requires_grad_(True) for k, v in dev.items() if is_leaf(k)})
                if mode == "u8":
                    slot_pipes[key] = LossInputPipeline(wl_h.opt, device, img_dtype)
            return slot_inputs[key]

        def pipeline(dev):
            # runs on the stager's copy stream right behind the H2D copies of this slot
            inputs, _ = slot_parts(dev)
            slot_pipes[id(dev)]({k[1]: v for k, v in dev.items() if k[0] == "color_u8"}, inputs)

        post = pipeline if mode == "u8" else None
        skip = set(x for x in args.e2e_skip.split(",") if x)
        if "pipeline" in skip and post is not None:
            stager.submit(host_batches[0], post); stager.take(); stager.release()
            stager.submit(host_batches[1], post); stager.take(); stager.release()   # both slots hold valid pyramids
            post = None
        if "h2d" in skip:
            real_submit = stager.submit
            primed = []

            def submit_once(hb, post=None):
                if len(primed) < 2:
                    primed.append(1)
                    return real_submit(hb, post)
                k = stager.next_slot
                stager.next_slot = (k + 1) % stager.depth
                stager.queue.append(k)   # re-use what the slot already holds: no copies
            stager.submit = submit_once

        def slot_step(dev):
            key = id(dev)
            if key not in slot_steps:
                inputs, leaves = slot_parts(dev)
                if args.no_graph:
                    def eager():
                        losses, grads = wl_h.step({"inputs": inputs, "leaves": leaves})
                        return losses, grads, wl_h.path.vsl_last_loss_vector
                    slot_steps[key] = eager
                else:
                    g = GraphedLossStep(wl_h.path, inputs, leaves, noise_prefetch=not args.no_noise_prefetch)
                    slot_steps[key] = lambda g=g: g.replay() + (g.loss_vector,)
            return slot_steps[key]

        # D2H read of every step's loss dict: copied into pinned memory right behind the step and read on the
        # host one step later, after step i+1 has been enqueued (how a training loop logs without stalling)
        # The copy runs on its own stream behind an event, so the compute stream goes straight on to the next step.
        result = [torch.empty(9, dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        stepped = [torch.cuda.Event() for _ in range(2)]
        rb_stream = torch.cuda.Stream(device=device)

        def e2e_loop(n):
            checksum, pending = 0.0, None
            stager.submit(host_batches[0], post)
            for i in range(n):
                dev = stager.take()
                losses, _, vec = slot_step(dev)()                  # enqueue step i
                stager.release()
                if vec is None:   # no contiguous loss vector (predictive-mask weighting): gather the dict
                    vec = torch.stack([losses[k] for k in sorted(losses)])
                stepped[i % 2].record()
                with torch.cuda.stream(rb_stream):
                    rb_stream.wait_event(stepped[i % 2])
                    if "readback" not in skip:
                        result[i % 2][:vec.numel()].copy_(vec, non_blocking=True)
                    done[i % 2].record(rb_stream)
                if i + 1 < n:
                    stager.submit(host_batches[(i + 1) % ring], post)   # step i+1's H2D + pyramid overlap step i
                if pending is not None:
                    done[pending].synchronize()
                    checksum += float(result[pending][:vec.numel()].sum())   # the host really reads every result
                pending = i % 2
            done[pending].synchronize()
            checksum += float(result[pending][:vec.numel()].sum())
            assert checksum == checksum or "readback" in skip, "NaN loss"
            return vec

        e2e_loop(4)
        barrier()
        e0.record()
        vec = e2e_loop(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return {"value": world * n0 * args.steps / (ms * 1e-3), "unit": UNIT,
                **({"INVALID_diagnostic_skip": sorted(skip)} if skip else {}),
                "h2d_bytes_per_step": stager.bytes_per_batch, "d2h_bytes_per_step": int(vec.numel() * 4),
                "ms_per_step": ms / args.steps}

    e2e_u8 = run_e2e("u8")
    e2e_u8.update({
        "host_buffers": "8-bit HWC level-0 frames + fp32 disp pyramid, poses, K/inv_K (pinned)",
        "gpu_launches_per_step": 3 + 3 + F,   # loss step (3) + target pyramid (3 LANCZOS levels; level 0 fused) + F source conversions
        "note": "pyramid (Pillow-exact LANCZOS) + ToTensor on the GPU, on the copy stream behind the H2D of the same "
                "batch: both overlap the loss kernels of the previous step; loss dict read back every step",
        "numa_node_rank0": numa_node})
    e2e_f32 = run_e2e("f32")
    e2e_f32["host_buffers"] = "fp32 (\"color\", f, s) tensors as the reference's DataLoader yields them + disp, poses, K/inv_K"

    # ---- N > 1: the one exchange of data-parallel training with this loss -- the all-reduce of the depth/pose-net
    # gradients (28,641,888 fp32 parameters) -- alone, and issued CONCURRENTLY with the loss step
    # (comm stream; buckets as parallel.GradBuckets cuts them), so the line shows what the two contend for
    grad_allreduce = train_step = None
    if dist is not None and not args.no_graph:
        payload = torch.zeros(28641888, device=device)
        nbytes = payload.numel() * 4
        comm = torch.cuda.Stream(device=device)
        loss_ms = ms_total / args.steps
        n_t = min(args.steps, 100)

        def measure(n_buckets, max_ctas):
            group = None
            if max_ctas is not None:   # a communicator of its own with few channels: fewer SMs taken from the loss kernel
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = max_ctas
                opts.config.min_ctas = min(max_ctas, 4)
                group = dist.new_group(backend="nccl", pg_options=opts)
            per = (payload.numel() + n_buckets - 1) // n_buckets
            buckets = [payload[k:k + per] for k in range(0, payload.numel(), per)]

            def alone(i):
                for w in [dist.all_reduce(b, async_op=True, group=group) for b in buckets]:
                    w.wait()

            def overlapped(i):
                # the exchange of the previous step's network gradients rides next to this step's loss kernels
                ev = torch.cuda.Event()
                ev.record()
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    works = [dist.all_reduce(b, async_op=True, group=group) for b in buckets]
                run_step(i)
                for w in works:
                    w.wait()   # the compute stream waits for the collectives before the next step (optimizer.step would)
            for i in range(3):
                alone(i)
            ar = timed_loop(alone, 20, barrier, device, dist)
            for i in range(5):
                overlapped(i)
            ov = timed_loop(overlapped, n_t, barrier, device, dist)
            return {"buckets": n_buckets, "nccl_max_ctas": max_ctas, "ms_allreduce_alone": ar,
                    "bus_gbs_alone": 2 * (world - 1) / world * nbytes / (ar * 1e-3) / 1e9,
                    "ms_per_step_overlapped": ov, "ms_if_serial": loss_ms + ar,
                    "hidden_fraction_of_allreduce": max(0.0, min(1.0, (loss_ms + ar - ov) / ar)),
                    "value": world * n0 / (ov * 1e-3), "unit": UNIT}

        variants = [measure(4, None), measure(1, None), measure(1, 8)]
        grad_allreduce = {"bytes": nbytes, "ms": variants[1]["ms_allreduce_alone"], "bus_gbs": variants[1]["bus_gbs_alone"],
                          "note": "NCCL all-reduce of the 28,641,888 fp32 net gradients as one flat bucket, nothing else running"}
        best = min(variants, key=lambda v: v["ms_per_step_overlapped"])
        train_step = {"ms_per_step_loss_alone": loss_ms, "best": best, "variants": variants,
                      "note": "graph-replayed loss step on the compute stream, the gradient all-reduce on a side stream issued "
                              "at the same time.  k_photometric fills every SM's register file, so an NCCL kernel only gets its "
                              "SMs if it starts BEFORE the loss kernel: one flat bucket (parallel.GradBuckets with one bucket) "
                              "hides, four 32 MB buckets leave three of them waiting for the loss kernel to drain"}
        del payload

    # ---- strong scaling, BASELINE config 5: a GLOBAL batch of 96 images at 640x192 sharded 96/N per GPU ----------
    strong = None
    if args.config == "C1" and not args.no_graph and 96 % world == 0 and not args.no_strong:
        cfg5 = dict(synthetic.CONFIGS["C5"])
        cfg5["batch"] = 96 // world
        wl5 = Workload(cfg5, args.family, device, 2, bf16_images=args.bf16_images)
        g5 = [GraphedLossStep(wl5.path, st["inputs"], st["leaves"], noise_prefetch=not args.no_noise_prefetch) for st in wl5.sets]
        for i in range(4):
            g5[i % 2].replay()
        n5 = min(args.steps, 40)
        ms5 = timed_loop(lambda i: g5[i % 2].replay(), n5, barrier, device, dist)
        strong = {"workload": "C5: global batch 96 at %dx%d, %d images per GPU" % (cfg5["width"], cfg5["height"], cfg5["batch"]),
                  "scaling": "strong", "global_batch": 96, "per_gpu_batch": cfg5["batch"], "ms_per_step": ms5,
                  "value": 96 * cfg5["height"] * cfg5["width"] / (ms5 * 1e-3), "unit": UNIT, "steps": n5}
        del g5, wl5

    if rank == 0:
        peak, peak_src = peak_hbm_gbs()
        kms = sum(kernel_ms) / max(1, len(kernel_ms))
        alg_bytes = algorithmic_bytes_per_pixel(F, 2 if args.bf16_images else 4) * n0
        traffic, traffic_src = profiled_traffic_bytes() if args.config == "C1" else (None, None)
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg),
            "l2_warm": l2_warm,
            "noise_inline": noise_inline,
            "contract_mode": contract,
            "e2e": e2e_u8,
            "e2e_f32_host_tensors": e2e_f32,
            "gpu_launches": 3 * args.steps,
            "kernels_per_step": ["k_photometric", "k_epilogue", "k_combine"],
            "host_wall_ms_per_step": 1e3 * t_wall / args.steps,
            "step_mode": "eager launches" if args.no_graph else "CUDA graph replay of the public-API step",
            "eager": {"ms_per_step": eager_ms, "host_enqueue_ms_per_step": 1e3 * t_enqueue / n_eager},
            "roofline": {"bound": "hbm", "kernel": "k_photometric", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "note": "bytes_min roofline as SURVEY.md 8d defines it; the fused kernel is instruction-issue bound "
                                 "(~400 warp-instructions per target pixel, see roofline_issue), not HBM bound: "
                                 "DESIGN.md section 4 and profiles/"},
            "clocks": clocks,
        }
        winst = getattr(profiled_traffic_bytes, "warp_instructions", None)
        if traffic is not None and winst:
            # what actually bounds the kernel: warp-instruction issue (4 schedulers/SM, 1 instruction/clock each)
            sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
            peak_issue = 148 * 4 * sm_hz * 1e6
            line["roofline_issue"] = {"bound": "instruction issue", "kernel": "k_photometric",
                                      "warp_instructions_per_launch": winst, "achieved": winst / (kms * 1e-3) / 1e9,
                                      "peak": peak_issue / 1e9, "unit": "G warp-inst/s",
                                      "frac": winst / (kms * 1e-3) / peak_issue,
                                      "source": "smsp__inst_executed.sum of " + traffic_src + ", live kernel time"}
        if args.config == "C1" and not args.bf16_images:
            # SURVEY.md 8d asks for these next to the bytes_min roofline: the reference's un-fused op stream moves
            # 55.88 GB per C1 step (2,109 ATen ops, device-independent count); the fused step
            # does the same arithmetic in ms_per_step, i.e. at this multiple of what HBM could stream
            unfused = 55.88e9
            line["roofline"]["supplementary"] = {
                "ncu": getattr(profiled_traffic_bytes, "ncu_pct", None),
                "reference_unfused_bytes_per_step": unfused,
                "unfused_equivalent_gbs": unfused / (ms_total / args.steps * 1e-3) / 1e9,
                "unfused_equivalent_over_hbm_peak": unfused / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                }
        aux = small_kernel_rooflines()
        if aux is not None and args.config == "C1":
            line["roofline_aux"] = aux
        if grad_allreduce is not None:
            line["grad_allreduce"] = grad_allreduce
            line["train_step"] = train_step
        if parity is not None:
            line["shard_parity"] = parity["status"]
            line["shard_parity_detail"] = parity
        if strong is not None:
            line["strong_scaling_c5"] = strong
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            run = oracle_step_cpu(cfg, args.family, B, threads)
            run()
            ts = [run() for _ in range(3)]
            line["cpu_baseline"] = {"value": n0 / min(ts), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "3 full %s steps (batch %d) after 1 warm-up, best of 3, oracle port "
                                              "of the reference on the host cores" % (args.config, B)}
            # the same op stream as eager PyTorch-CUDA on this GPU, measured in this run: what the
            # reference does when it is NOT given --no_cuda.  Checker code used as a reported baseline, like cpu_baseline.
            run = oracle_step_cpu(cfg, args.family, B, threads, device=device)
            run()
            tg = [run() for _ in range(3)]
            line["eager_cuda"] = {"ms_per_step": 1e3 * min(tg), "value": n0 / min(tg), "unit": UNIT,
                                  "speedup_of_fused_step": 1e3 * min(tg) / (ms_total / args.steps),
                                  "sample": "3 full %s steps after 1 warm-up, best of 3: the reference's op stream "
                                            "(oracle port) as eager PyTorch on the same B200, wall clock around "
                                            "synchronize()" % args.config}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
