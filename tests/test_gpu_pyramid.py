"""GPU parity of the on-device input pipeline (vsl_pyramid_forward) against the Pillow-pinned oracle:
byte-exact 8-bit levels, bit-exact float tensors."""
import os
import sys

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


AUGMENT = os.path.join(os.path.dirname(__file__), "golden", "pyramid", "augment_pil.npz")


@pytest.mark.parametrize("name", ["iid_64x96", "smooth_96x160", "edges_32x64"])
def test_flip_equals_pillow_goldens(name):
    """MonoDataset's do_flip (`color.transpose(Image.FLIP_LEFT_RIGHT)` before the pyramid) applied per image while the
    raw frame is read: flagged images equal the goldens made with PIL's transpose + torchvision's Resize/ToTensor,
    unflagged images in the same batch are untouched."""
    from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid
    g, a = np.load(GOLDEN), np.load(AUGMENT)
    img = g[name + "/u8_0"]
    batch = torch.from_numpy(np.stack([img, img, img[::-1].copy()])).cuda()
    flip = torch.tensor([1, 0, 1], dtype=torch.uint8, device="cuda")
    H, W = img.shape[:2]
    for levels in (None, [0]):   # all levels (target frame: fused level-0 write) and level 0 only (source frames)
        pyr = FramePyramid(3, H, W, 4, "cuda", levels=levels)
        out, u8 = pyr(batch, want_u8=True, flip=flip)
        torch.cuda.synchronize()
        assert np.array_equal(out[0][0].cpu().numpy(), O.to_tensor(a["flip/%s/u8_0" % name]))
        assert np.array_equal(out[0][1].cpu().numpy(), O.to_tensor(img))
        if levels is None:
            for s in range(1, 4):
                assert np.array_equal(u8[s][0].cpu().numpy(), a["flip/%s/u8_%d" % (name, s)]), (name, s)
                assert np.array_equal(u8[s][1].cpu().numpy(), g["%s/u8_%d" % (name, s)]), (name, s)
            assert np.array_equal(out[3][0].cpu().numpy(), a["flip/%s/f32_3" % name])
            # image 2 is the vertically mirrored frame, flipped horizontally: the oracle on the mirrored input
            lv, tn = O.pyramid(img[::-1, ::-1].copy()[None], 4)
            for s in range(1, 4):
                assert np.array_equal(u8[s][2].cpu().numpy(), lv[s][0]), s


@pytest.mark.parametrize("shape", [(2, 24, 40), (3, 10, 18), (2, 192, 640)])
def test_flip_ragged_and_full_shapes(shape):
    """Widths that are not multiples of 4 take the scalar conversion and the byte-column staging."""
    from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid
    B, H, W = shape
    rng = np.random.RandomState(H * W)
    batch = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    n = 2 if (H % 8 or W % 8) else 4
    flip = np.arange(B) % 2 == 0
    pyr = FramePyramid(B, H, W, n, "cuda")
    out, u8 = pyr(torch.from_numpy(batch).cuda(), want_u8=True, flip=torch.from_numpy(flip.astype(np.uint8)).cuda())
    want = np.where(flip[:, None, None, None], batch[:, :, ::-1], batch)
    levels, tensors = O.pyramid(np.ascontiguousarray(want), n)
    for s in range(n):
        if s:
            assert np.array_equal(u8[s].cpu().numpy(), levels[s]), s
        assert np.array_equal(out[s].cpu().numpy(), tensors[s]), s


def test_stereo_T_and_cached_intrinsics():
    """inputs["stereo_T"] from the per-item flags (datasets/mono_dataset2.py:197-203) and the per-scale K / inv_K as
    plan constants (:168-177), equal to what the reference's dataset code produces."""
    from unsupervised_pose_estimation_b200 import synthetic
    from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline
    from unsupervised_pose_estimation_b200.trainer import make_opt
    opt = make_opt(height=32, width=64, batch_size=4, frame_ids=[0, -1, 1, "s"])
    pipe = LossInputPipeline(opt, "cuda")
    flip = torch.tensor([0, 1, 0, 1], dtype=torch.uint8, device="cuda")
    left = torch.tensor([0, 0, 1, 1], dtype=torch.uint8, device="cuda")
    T = pipe.stereo_T(flip, left).cpu().numpy()
    for b, (do_flip, side) in enumerate([(False, "r"), (True, "r"), (False, "l"), (True, "l")]):
        ref = np.eye(4, dtype=np.float32)   # the reference's own lines
        baseline_sign = -1 if do_flip else 1
        side_sign = -1 if side == "l" else 1
        ref[0, 3] = side_sign * baseline_sign * 0.1
        assert np.array_equal(T[b], ref), b
    intr = pipe.intrinsics(synthetic.K_KITTI)
    ref = synthetic.scaled_intrinsics(synthetic.K_KITTI, 32, 64, 4, 4)   # restates mono_dataset2.py:168-177
    for k, v in ref.items():
        assert torch.equal(intr[k].cpu(), v), k
    assert pipe.intrinsics(synthetic.K_KITTI)[("K", 0)] is intr[("K", 0)]   # cached: no per-step work or copy
    rng = np.random.RandomState(0)
    frames = {f: torch.from_numpy(rng.randint(0, 256, (4, 32, 64, 3)).astype(np.uint8)).cuda() for f in opt.frame_ids}
    inputs = pipe(frames, flip=flip, side_left=left)
    assert torch.equal(inputs["stereo_T"].cpu(), torch.from_numpy(T))
    assert torch.equal(inputs[("color", "s", 0)][1].cpu(), torch.from_numpy(O.to_tensor(frames["s"][1].cpu().numpy()[:, ::-1])))


def _resize_cases():
    here = os.path.join(os.path.dirname(__file__), "golden", "pyramid")
    if here not in sys.path:
        sys.path.insert(0, here)
    import resize_cases
    return resize_cases.CASES, resize_cases.make_input, np.load(os.path.join(here, "resize_pil.npz"))


def test_level0_resize_equals_pillow_goldens_and_oracle():
    """FrameResize (vsl_resize_forward): the decoded file image -> level 0 at arbitrary ratios, byte for byte."""
    from unsupervised_pose_estimation_b200.input_pipeline import FrameResize
    cases, make_input, g = _resize_cases()
    for name, h, w, oh, ow, family in cases:
        img = make_input(name, h, w, family)
        batch = np.stack([img, img[::-1].copy(), img[:, ::-1].copy()])
        out = FrameResize(3, h, w, oh, ow)(torch.from_numpy(batch).cuda()).cpu().numpy()
        assert np.array_equal(out[0], g[name]), name
        assert np.array_equal(out[1], O.resize_lanczos(batch[1], oh, ow)), name
        assert np.array_equal(out[2], O.resize_lanczos(batch[2], oh, ow)), name


def test_native_frames_through_resize_and_pyramid():
    """native frame -> FrameResize -> FramePyramid equals the reference's whole preprocess chain (oracle)."""
    from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid, FrameResize
    rng = np.random.RandomState(77)
    native = rng.randint(0, 256, (2, 375, 1242, 3)).astype(np.uint8)
    lvl0 = FrameResize(2, 375, 1242, 192, 640)(torch.from_numpy(native).cuda())
    out = FramePyramid(2, 192, 640, 4)(lvl0)
    torch.cuda.synchronize()
    ref0 = O.resize_lanczos(native, 192, 640)
    _, tensors = O.pyramid(ref0, 4)
    for s in range(4):
        assert np.array_equal(out[s].cpu().numpy(), tensors[s]), s
    with pytest.raises(Exception):
        FrameResize(2, 375, 1242, 192, 640, "cpu")
