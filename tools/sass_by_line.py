"""Attribute an `ncu --set full --import-source on` capture of k_photometric to SOURCE LINES.

    python tools/sass_by_line.py gpurun_out/prof_photometric_<tag>.ncu-rep [--lib path/to/libvsl_b200.so] [--top 40]

ncu's CLI source page only lists SASS; the line table comes from `nvdisasm -g` of the same cubin (the library
must be the build that was profiled: the join is by instruction offset and is checked opcode by opcode).
Prints, per source line and per phase function: executed warp instructions, stall samples, excessive
shared-memory wavefronts (bank conflicts).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    kernel = lines[start - 1]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    return kernel, rows


def disasm_lines(lib, mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith("vsl_fused.")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    insts, cur, active = [], None, False
    for l in txt.splitlines():
        if l.startswith(".text."):
            active = mangled_substr in l
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            insts.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return insts


def mangle_hint(kernel_name):
    # "void vsl::k_photometric<vsl::TileCfg<(int)32, (int)16, (int)2, (int)256, float, (bool)0, (bool)0>, (bool)1>"
    nums = re.findall(r"\((?:int|bool)\)(\d+)", kernel_name)
    img = "NS_6bf16_tE" if "bf16" in kernel_name else "f"
    tw, th, f, nt, avg, pm, fast = nums[:7]
    return "k_photometricINS_7TileCfgILi%sELi%sELi%sELi%sE%sLb%sELb%sEEELb%sE" % (tw, th, f, nt, img, avg, pm, fast)


def main():
    rep = sys.argv[1]
    lib = os.path.join(ROOT, "unsupervised_pose_estimation_b200", "libvsl_b200.so")
    top = 40
    if "--lib" in sys.argv:
        lib = sys.argv[sys.argv.index("--lib") + 1]
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    kernel, rows = ncu_sass_rows(rep)
    insts = disasm_lines(lib, mangle_hint(kernel))
    if len(insts) != len(rows):
        print("WARNING: %d SASS rows in the capture, %d in the library's cubin" % (len(rows), len(insts)))
    n = min(len(insts), len(rows))
    mism = sum(1 for i in range(n) if rows[i]["Source"].split()[0].rstrip(";") != insts[i][1].split()[0])
    print("kernel:", kernel[:160])
    print("joined %d instructions, %d opcode mismatches" % (n, mism))
    src = {}
    for fn in ("vsl_tile.cuh", "vsl_math.cuh", "vsl_fused.cu"):
        src[fn] = open(os.path.join(ROOT, "unsupervised_pose_estimation_b200", "csrc", fn)).read().splitlines()
    by_line = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0, 0]
    for i in range(n):
        r = rows[i]
        vals = [int(float(r["Instructions Executed"] or 0)), int(float(r["# Samples"] or 0)),
                int(float(r["L1 Wavefronts Shared Excessive"] or 0)), int(float(r["L1 Wavefronts Shared"] or 0))]
        key = insts[i][2]
        for k in range(4):
            by_line[key][k] += vals[k]
            tot[k] += vals[k]
    print("totals: inst %d, samples %d, excessive shared wavefronts %d of %d" % tuple(tot))
    for title, idx in (("warp instructions", 0), ("stall samples", 1), ("excessive shared wavefronts", 2)):
        print("\n== top source lines by %s ==" % title)
        for key, v in sorted(by_line.items(), key=lambda kv: -kv[1][idx])[:top]:
            if v[idx] == 0:
                break
            text = src.get(key[0], [""] * 100000)[key[1] - 1].strip()[:110] if key else ""
            print("%5.2f%%  inst %5.2f%% samp %5.2f%% xwf %5.2f%%  %s:%d  %s" % (
                100.0 * v[idx] / max(1, tot[idx]), 100.0 * v[0] / tot[0], 100.0 * v[1] / max(1, tot[1]),
                100.0 * v[2] / max(1, tot[2]), key[0] if key else "?", key[1] if key else 0, text))


if __name__ == "__main__":
    main()
