"""Build kernel variants next to the product library for side-by-side timing on the GPU box.

    python tools/build_variants.py name1:"-DFOO=1 -DBAR" name2:"" ...   ->  variants/libvsl_<name>.so (+ ptxas -v log)

`VSL_LIB_PATH=variants/libvsl_<name>.so python bench.py ...` then loads that build (same ABI, same compute path).
"""
import os
import re
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unsupervised_pose_estimation_b200 import build as B  # noqa: E402


def one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(ROOT, "variants", "libvsl_%s.so" % name)
    import subprocess
    nvcc = B.find_nvcc()
    cmd = [nvcc] + B.NVCC_FLAGS + flags.split() + ["-Xptxas", "-v", "-o", out] + [os.path.join(B.CSRC, f) for f in B.SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        return name, "FAILED\n" + res.stderr[-3000:]
    open(out + ".ptxas.log", "w").write(res.stderr)
    # registers / spills of the default C1 kernel
    m = re.search(r"Compiling entry function '(_ZN3vsl13k_photometricINS_7TileCfgILi32ELi16ELi2ELi256EfLb0ELb0EEELb1E[^']*)'.*?\n(.*?Used \d+ registers[^\n]*)",
                  res.stderr, re.S)
    info = m.group(2).strip().replace("\n", " | ") if m else "?"
    return name, info


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, info in ex.map(one, sys.argv[1:]):
            print("%-14s %s" % (name, info[-400:]))
