"""oracle/metrics_oracle.py (SLlog, compute_depth_errors, compute_depth_losses) against the golden values made
from the unmodified reference (tests/golden/metrics/make_golden_metrics.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics", "metrics.npz"))


@pytest.mark.parametrize("seed", [0, 1])
def test_metrics_oracle_reproduces_reference(seed):
    torch.set_num_threads(1)
    fake, real = M.metric_inputs(seed, "sllog")
    fake, real = fake.requires_grad_(True), real.requires_grad_(True)
    loss = M.sllog(fake, real)
    gf, gr = torch.autograd.grad(loss, [fake, real])
    assert np.allclose(loss.item(), GOLD["sllog|%d|loss" % seed], rtol=1e-6)
    assert np.allclose(gf.numpy(), GOLD["sllog|%d|grad_fake" % seed], rtol=1e-5, atol=1e-9)
    assert np.allclose(gr.numpy(), GOLD["sllog|%d|grad_real" % seed], rtol=1e-5, atol=1e-9)
    gt, pred = M.metric_inputs(seed, "errors")
    got = np.array([float(v) for v in M.compute_depth_errors(gt, pred)])
    assert np.allclose(got, GOLD["errors|%d" % seed], rtol=1e-6)
    dpred, dgt = M.metric_inputs(seed, "depth_losses")
    got = np.array([float(v) for v in M.compute_depth_losses(dpred, dgt)])
    assert np.allclose(got, GOLD["depth_losses|%d" % seed], rtol=1e-6)
