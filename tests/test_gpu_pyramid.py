"""GPU parity of the on-device input pipeline (vsl_pyramid_forward) against the Pillow-pinned oracle:
byte-exact 8-bit levels, bit-exact float tensors."""
import os

import numpy as np
import pytest
import torch

from oracle import pil_pyramid_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pyramid", "pyramid_pil.npz")


def _run(batch_u8, num_levels=4, dtype=torch.float32, levels=None):
    from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid
    B, H, W, _ = batch_u8.shape
    pyr = FramePyramid(B, H, W, num_levels, "cuda", dtype, levels=levels)
    out, u8 = pyr(torch.from_numpy(batch_u8).cuda(), want_u8=True)
    torch.cuda.synchronize()
    return {s: t.cpu() for s, t in out.items()}, {s: t.cpu().numpy() for s, t in u8.items()}


@pytest.mark.parametrize("name", ["iid_64x96", "smooth_96x160", "edges_32x64"])
def test_pyramid_equals_pillow_goldens(name):
    g = np.load(GOLDEN)
    img = g[name + "/u8_0"]
    out, u8 = _run(np.stack([img, img[::-1].copy()]))
    for s in range(1, 4):
        assert np.array_equal(u8[s][0], g["%s/u8_%d" % (name, s)]), (name, s)
    assert np.array_equal(out[3][0].numpy(), g[name + "/f32_3"])
    assert np.array_equal(out[0][0].numpy(), O.to_tensor(img))


@pytest.mark.parametrize("shape", [(12, 192, 640), (2, 256, 320), (1, 320, 1024), (3, 8, 8), (2, 24, 40), (1, 72, 200)])
def test_pyramid_equals_oracle(shape):
    rng = np.random.RandomState(sum(shape))
    B, H, W = shape
    batch = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    # smooth half, so that interior coefficients and the saturating clip both matter
    batch[: max(1, B // 2)] = (127 + 120 * np.sin(np.arange(W)[None, None, :, None] / 5.0 + np.arange(H)[None, :, None, None] / 3.0)).astype(np.uint8)
    out, u8 = _run(batch)
    levels, tensors = O.pyramid(batch, 4)
    for s in range(4):
        if s:
            assert np.array_equal(u8[s], levels[s]), s
        assert np.array_equal(out[s].numpy(), tensors[s]), s


def test_pyramid_level_subset_and_bf16():
    rng = np.random.RandomState(5)
    batch = rng.randint(0, 256, (2, 32, 64, 3)).astype(np.uint8)
    out, _ = _run(batch, levels=[0])
    assert sorted(out) == [0] and np.array_equal(out[0].numpy(), O.to_tensor(batch))
    out16, _ = _run(batch, dtype=torch.bfloat16)
    _, tensors = O.pyramid(batch, 4)
    for s in range(4):
        assert torch.equal(out16[s], torch.from_numpy(tensors[s]).bfloat16()), s


def test_loss_from_u8_frames_equals_loss_from_float_tensors():
    """LossInputPipeline output fed to compute_losses == the reference-style float inputs fed to it."""
    from unsupervised_pose_estimation_b200 import layers as L
    from unsupervised_pose_estimation_b200 import synthetic
    from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline
    from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt
    B, H, W, frames = 2, 64, 96, [0, -1, 1]
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, seed=3, family="smooth", device="cuda")
    opt = make_opt(height=H, width=W, batch_size=B, frame_ids=frames)
    u8 = {f: (inputs[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8).contiguous()
          for f in frames}
    ref_inputs = dict(inputs)
    for f in frames:   # what the reference's dataset would produce from these 8-bit frames
        _, tensors = O.pyramid(u8[f].cpu().numpy(), 4)
        for s in range(4):
            ref_inputs[("color", f, s)] = torch.from_numpy(tensors[s]).cuda()
    new_inputs = {k: v for k, v in inputs.items() if not (isinstance(k, tuple) and k[0] == "color")}
    LossInputPipeline(opt, "cuda")(u8, new_inputs)
    assert torch.equal(new_inputs[("color", 0, 2)], ref_inputs[("color", 0, 2)])
    assert ("color", -1, 1) not in new_inputs   # source frames: level 0 only

    def run(inp):
        out = dict(outputs)
        for f in frames[1:]:
            out[("cam_T_cam", 0, f)] = L.transformation_from_parameters(
                leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
        path = LossPath(opt, device="cuda", side_outputs="none")
        path.generate_images_pred(inp, out)
        torch.manual_seed(11)
        losses = path.compute_losses(inp, out)
        grads = torch.autograd.grad(losses["loss"], list(leaves.values()))
        return losses, grads, out
    la, ga, oa = run(ref_inputs)
    lb, gb, ob = run(new_inputs)
    for k in la:
        assert torch.equal(la[k], lb[k]), k
    for a, b in zip(ga, gb):
        assert torch.equal(a, b)
    assert torch.equal(oa["identity_selection/0"], ob["identity_selection/0"])


def test_host_batch_stager_with_pipeline_equals_direct_upload():
    """The e2e loop of bench.py in miniature: pinned host batch -> HostBatchStager (copy stream) with the input
    pipeline as its post hook -> loss step; twice through both slots, against a plain synchronous upload."""
    from unsupervised_pose_estimation_b200 import layers as L
    from unsupervised_pose_estimation_b200 import synthetic
    from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline
    from unsupervised_pose_estimation_b200.staging import HostBatchStager
    from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt
    B, H, W, frames = 2, 64, 96, [0, -1, 1]
    opt = make_opt(height=H, width=W, batch_size=B, frame_ids=frames)
    path = LossPath(opt, device="cuda", side_outputs="none")
    dev = torch.device("cuda", torch.cuda.current_device())

    def host_batch(seed):
        inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, seed=seed, family="smooth", device="cpu")
        hb = {k: v for k, v in inputs.items() if isinstance(k, tuple) and k[0] in ("K", "inv_K") and k[1] == 0}
        for f in frames:
            hb[("color_u8", f)] = (inputs[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
        for s in range(4):
            hb[("disp", s)] = outputs[("disp", s)].detach()
        for f in frames[1:]:
            hb[("cam_T_cam", 0, f)] = L.transformation_from_parameters(
                leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0).detach()
        return {k: v.pin_memory() for k, v in hb.items()}

    def step(devb, inputs):
        outputs = {k: v for k, v in devb.items() if k[0] in ("disp", "cam_T_cam")}
        path.generate_images_pred(inputs, outputs)
        torch.manual_seed(5)
        losses = path.compute_losses(inputs, outputs)
        return torch.stack([losses[k] for k in sorted(losses)]).cpu(), outputs["identity_selection/0"].cpu()

    stager = HostBatchStager(dev, depth=2)
    pipes, slot_inputs = {}, {}

    def post(devb):
        key = id(devb)
        if key not in pipes:
            pipes[key] = LossInputPipeline(opt, dev)
            slot_inputs[key] = {k: v for k, v in devb.items() if k[0] in ("K", "inv_K")}
        pipes[key]({k[1]: v for k, v in devb.items() if k[0] == "color_u8"}, slot_inputs[key])

    batches = [host_batch(s) for s in range(3)]
    got = []
    stager.submit(batches[0], post)
    for i in range(3):
        if i + 1 < 3:
            stager.submit(batches[i + 1], post)
        devb = stager.take()
        got.append(step(devb, slot_inputs[id(devb)]))
        stager.release()
    assert len(pipes) == 2   # two slots, re-used
    direct_pipe = LossInputPipeline(opt, dev)
    for i in range(3):
        devb = {k: v.to(dev) for k, v in batches[i].items()}
        inputs = {k: v for k, v in devb.items() if k[0] in ("K", "inv_K")}
        direct_pipe({k[1]: v for k, v in devb.items() if k[0] == "color_u8"}, inputs)
        want = step(devb, inputs)
        assert torch.equal(got[i][0], want[0]) and torch.equal(got[i][1], want[1]), i
