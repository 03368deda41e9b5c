"""Per-leaf gradient error over the named parity cases of tests/test_gpu_parity.py (to set the test gates)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
worst = {}
for name in sorted(T.CASES):
    opt, inputs, outputs, leaves = T.build_case(name)
    _, _, rg = T.run_oracle(opt, inputs, outputs, leaves)
    _, _, g = T.run_ours(opt, inputs, outputs, leaves)
    row = []
    for k in rg:
        e = ((g[k] - rg[k]).norm() / rg[k].norm()).item()
        kind = "disp_0" if k == ("disp", 0) else k[0]
        worst[kind] = max(worst.get(kind, 0), e)
        row.append("%s=%.1e" % ("/".join(str(x) for x in k), e))
    print(name, CASES := T.CASES[name][5], " ".join(row), flush=True)
print("WORST", worst)
