"""Developer check (GPU box): very small images (smaller than one tile, down to 4x4) against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_gpu_parity as T
from oracle import vsl_oracle as O
from unsupervised_pose_estimation_b200 import synthetic

bad = 0
for (H, W, scales) in [(4, 4, [0]), (4, 8, [0, 1]), (8, 8, [0, 1, 2]), (6, 10, [0]), (12, 4, [0, 1]), (16, 16, [0, 1, 2, 3]),
                       (8, 40, [0, 1, 2]), (24, 8, [0, 1, 2]), (2, 2, [0]), (2, 34, [0])]:
    for frames in ([0, -1, 1], [0, -1, 1, "s"], [0, 1]):
        opt = O.make_opt(height=H, width=W, batch_size=2, frame_ids=list(frames), scales=scales)
        inputs, outputs, leaves = synthetic.make_batch(2, H, W, frames, synthetic.K_KITTI, scales=tuple(scales), seed=H * W,
                                                       family="iid", device=T.DEV)
        ref_out, ref_losses, ref_g = T.run_oracle(opt, inputs, outputs, leaves, seed=1)
        out, losses, g = T.run_ours(opt, inputs, outputs, leaves, seed=1, side="eager")
        ok = all(abs(losses[k].item() - ref_losses[k].item()) <= 2e-6 * abs(ref_losses[k].item()) for k in ref_losses)
        ok &= all(torch.equal(out["identity_selection/%d" % s], ref_out["identity_selection/%d" % s]) for s in scales)
        ok &= all(torch.equal(out[("sample", f, s)], ref_out[("sample", f, s)]) for s in scales for f in frames[1:])
        gerr = max(((g[k] - ref_g[k]).norm() / ref_g[k].norm().clamp_min(1e-30)).item() for k in ref_g)
        ok &= gerr <= 2e-4
        bad += not ok
        print(H, W, scales, frames, "ok" if ok else "MISMATCH", "grad err %.1e" % gerr)
print("mismatching cases:", bad)
