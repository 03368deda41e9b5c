"""Generate golden vectors by running the UNMODIFIED reference implementation.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so these files are the
pin for ``oracle/vsl_oracle.py``: reference ``layers.py`` is imported as is, and
``Trainer.generate_images_pred / compute_reprojection_loss / compute_losses``
(reference trainer.py:491-686) are bound as methods on a namespace carrying exactly the
attributes they read (``Trainer.__init__`` itself cannot run offline: wandb login,
dataset constructors).  Five third-party modules that ``trainer.py`` imports but the
path never touches are stubbed.  Outputs: ``tests/golden/<case>.npz``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VSL_REFERENCE", "/root/reference")

CASES = {
    # name: (batch, height, width, frame_ids, K, family, seed, opt overrides)
    "mono_iid_b2_32x64": dict(batch=2, height=32, width=64, frame_ids=[0, -1, 1], K="kitti",
                              family="iid", seed=0, opt={}),
    "mono_smooth_b2_32x96": dict(batch=2, height=32, width=96, frame_ids=[0, -1, 1], K="scared",
                                 family="smooth", seed=1, opt={"max_depth": 150.0, "disparity_smoothness": 1e-4}),
    "stereo_iid_b2_32x64": dict(batch=2, height=32, width=64, frame_ids=[0, -1, 1, "s"], K="kitti",
                                family="iid", seed=2, opt={}),
}


def load_reference():
    """Import reference layers.py / trainer.py with the absent, unused deps stubbed."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("tensorboardX", SummaryWriter=object)
    stub("torchview", draw_graph=lambda *a, **k: None)
    stub("IPython", embed=lambda *a, **k: None)
    sk = stub("skimage")
    sk.transform = stub("skimage.transform")
    mpl = stub("matplotlib")
    mpl.cm = stub("matplotlib.cm")
    mpl.pyplot = stub("matplotlib.pyplot")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import layers as ref_layers  # noqa
    import trainer as ref_trainer  # noqa
    return ref_layers, ref_trainer


def bind_reference(ref_layers, ref_trainer, opt, device="cpu"):
    """Namespace that can run the three Trainer methods (mirrors trainer.py:245-259)."""
    ns = types.SimpleNamespace()
    ns.opt = opt
    ns.device = torch.device(device)
    ns.num_scales = len(opt.scales)
    ns.ssim = ref_layers.SSIM().to(device)
    ns.backproject_depth, ns.project_3d = {}, {}
    for s in opt.scales:
        h, w = opt.height // 2 ** s, opt.width // 2 ** s
        ns.backproject_depth[s] = ref_layers.BackprojectDepth(opt.batch_size, h, w).to(device)
        ns.project_3d[s] = ref_layers.Project3D(opt.batch_size, h, w).to(device)
    for name in ("generate_images_pred", "compute_reprojection_loss", "compute_losses"):
        setattr(ns, name, types.MethodType(getattr(ref_trainer.Trainer, name), ns))
    return ns


def run_case(name, spec, ref_layers, ref_trainer):
    sys.path.insert(0, ROOT)
    from unsupervised_pose_estimation_b200 import synthetic
    from oracle import vsl_oracle

    K = {"kitti": synthetic.K_KITTI, "scared": synthetic.K_SCARED}[spec["K"]]
    opt = vsl_oracle.make_opt(height=spec["height"], width=spec["width"], batch_size=spec["batch"],
                              frame_ids=list(spec["frame_ids"]), **spec["opt"])
    inputs, outputs, leaves = synthetic.make_batch(
        spec["batch"], spec["height"], spec["width"], spec["frame_ids"], K, seed=spec["seed"],
        family=spec["family"], pose_fn=ref_layers.transformation_from_parameters)
    ns = bind_reference(ref_layers, ref_trainer, opt)

    ns.generate_images_pred(inputs, outputs)
    torch.manual_seed(123)
    losses = ns.compute_losses(inputs, outputs)
    losses["loss"].backward()

    # the tie-break noise the reference drew (trainer.py:656-657), re-drawn from the same seed
    torch.manual_seed(123)
    n_src = len(spec["frame_ids"]) - 1
    noise = [torch.randn(spec["batch"], n_src, spec["height"], spec["width"]) for _ in opt.scales]

    blob = {}

    def put(key, t):
        blob[key] = t.detach().cpu().numpy()

    for k, v in inputs.items():
        put("in|" + repr(k), v)
    for k, v in leaves.items():
        put("leaf|" + repr(k), v)
        put("grad|" + repr(k), v.grad)
    for k, v in outputs.items():
        if k in leaves or (isinstance(k, tuple) and k[0] == "color_identity"):
            continue  # leaves are stored above; color_identity aliases an input
        put("out|" + repr(k), v)
    for k, v in losses.items():
        put("loss|" + k, v)
    for i, z in enumerate(noise):
        put("noise|%d" % i, z)
    blob["meta|opt"] = np.array(repr({k: v for k, v in vars(opt).items()}))
    blob["meta|torch"] = np.array(torch.__version__)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), {k: float(v) for k, v in losses.items()})


def main():
    torch.set_num_threads(1)
    ref_layers, ref_trainer = load_reference()
    for name, spec in CASES.items():
        run_case(name, spec, ref_layers, ref_trainer)


if __name__ == "__main__":
    main()
