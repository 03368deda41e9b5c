"""Shared test utilities: golden loading and oracle drivers (test infrastructure)."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np
import torch

from oracle import vsl_oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name, device="cpu"):
    """-> dict(opt, inputs, leaves, grads, outputs, losses, noise) of torch tensors."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = dict(inputs={}, leaves={}, grads={}, outputs={}, losses={}, noise={})
    for key in z.files:
        kind, _, rest = key.partition("|")
        if kind == "meta":
            continue
        t = torch.from_numpy(z[key]).to(device)
        if kind == "in":
            g["inputs"][ast.literal_eval(rest)] = t
        elif kind == "leaf":
            g["leaves"][ast.literal_eval(rest)] = t
        elif kind == "grad":
            g["grads"][ast.literal_eval(rest)] = t
        elif kind == "out":
            g["outputs"][ast.literal_eval(rest)] = t
        elif kind == "loss":
            g["losses"][rest] = t
        elif kind == "noise":
            g["noise"][int(rest)] = t
    g["noise"] = [g["noise"][i] for i in sorted(g["noise"])]
    g["opt"] = vsl_oracle.make_opt(**ast.literal_eval(str(z["meta|opt"])))
    return g


def fresh_leaves(g):
    """Detached, grad-requiring copies of the golden leaves + an outputs dict built from them."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in g["leaves"].items()}
    outputs = {}
    for k, v in leaves.items():
        outputs[k] = v
    return leaves, outputs


def run_oracle(opt, inputs, leaves, noise=None, pose_fn=None, backward=True):
    """Oracle forward (+ backward of losses['loss']). Returns (outputs, losses, grads)."""
    pose_fn = pose_fn or vsl_oracle.transformation_from_parameters
    outputs = dict(leaves)
    for f in opt.frame_ids[1:]:
        if f == "s":
            continue
        outputs[("cam_T_cam", 0, f)] = pose_fn(
            leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
    losses = vsl_oracle.loss_step(opt, inputs, outputs, noise)
    grads = {}
    if backward:
        losses["loss"].backward()
        grads = {k: v.grad for k, v in leaves.items()}
    return outputs, losses, grads
