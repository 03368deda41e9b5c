"""profiles/<tag>_small_kernels.md + profiles/<tag>_sass/*.sass from tools/gpu_small_kernels.sh's outputs.

    python tools/summarize_small_kernels.py <tag>      (reads gpurun_out/small_raw_<tag>.csv, small_launches_<tag>.csv)

Per kernel: the `ncu --set full` capture's duration, DRAM bytes, achieved DRAM GB/s against the measured HBM peak
(MEASURED_PEAKS.json), occupancy and issue utilisation; and the SASS of every product kernel, cut from the library.
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
peak = 6535.7
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

# what each kernel has to move at the profiled shape (C1: B=12, 640x192, F=2), bytes
n0 = 12 * 192 * 640
ALG = {
    "k_epilogue": ("tile partials + up-sample-adjoint partials in, d/d disp_s (s>=1) out", 1.4e6 + 2.9e6 + 4 * n0 * (1 / 4 + 1 / 16 + 1 / 64)),
    "k_combine": ("2 unit-gradient pyramids in, 1 out", 3 * 4 * n0 * (1 + 1 / 4 + 1 / 16 + 1 / 64)),
    "k_warp_forward": ("disp_s + 2 gathered frames in; depth, 2 grids, 2 warped frames out", n0 * (4 + 2 * 12 + 4 + 2 * 8 + 2 * 12)),
    "k_u8_to_tensor_x4": ("3 B in, 12 B out per pixel", n0 * 15),
    "k_lanczos_half": ("8-bit level in, 8-bit + fp32 level out (captured launch: 320x96 -> 160x48)", 12 * (320 * 96 * 3 + 160 * 48 * 15)),
    "k_source_grad_upstream": ("4 winner maps in, 2x4 + 2 weight maps out", n0 * (4 + 4 * 10)),
    "k_grid_sample_bwd_source": ("grid + d/d pred in, source gradient accumulated", n0 * (8 + 12 + 12)),
    "k_ssim_coef": ("pred, target, upstream in; 4 coefficient planes out", n0 * (12 + 12 + 4 + 48)),
    "k_reproj_bwd": ("pred, target, upstream, coefficient planes (3x3 gather) in; d/d pred out", n0 * (12 + 12 + 4 + 48 + 12)),
    "k_median_hist": ("depth_gt + depth_pred in (375x1242 ground truth)", 12 * 375 * 1242 * 4 + n0 * 4),
    "k_depth_errors": ("depth_gt + depth_pred in", 12 * 375 * 1242 * 4 + n0 * 4),
    "k_sllog_fwd": ("fake + real in", n0 * 8),
    "k_aug_stats": ("8-bit frames in (grey sum of the partially jittered image)", n0 * 3),
    "k_aug_apply": ("8-bit frames in, jittered 8-bit frames out", n0 * 6),
    "k_aug_finish": ("jittered 8-bit frames in, fp32 tensor out", n0 * 15),
}

rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "small_raw_%s.csv" % tag))))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = r[col[name]].replace(",", "") if name in col else ""
    try:
        return float(v)
    except ValueError:
        return None


def to_us(r):
    v, u = val(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]]
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)


def to_bytes(r, name):
    v, u = val(r, name), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


out = ["# %s — every product kernel except `k_photometric`: one `ncu --set full --clock-control none` capture each" % tag, "",
       "Driver: `tools/run_small_kernels.py` at config C1 (B=12, 640x192, two source frames), second invocation of each kernel",
       "(`tools/gpu_small_kernels.sh`).  Times under ncu are cold-cache and serialised.  `achieved` = measured DRAM bytes / time;",
       "`algorithmic` = the bytes the kernel has to move at this shape / time; both against the measured HBM peak of %.0f GB/s." % peak, "",
       "| kernel | grid x block | regs | time µs | DRAM MB (r+w) | achieved GB/s (frac) | algorithmic MB | algorithmic GB/s (frac) | warps active % | issue active % |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
aux = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("vsl::", "")
    base = re.sub(r"<.*", "", name)
    if base == "k_photometric" or base == "k_probe_bmm":
        continue
    us = to_us(r)
    dram = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    ach = dram / (us * 1e-6) / 1e9
    alg = ALG.get(base)
    alg_s = "%.1f | %.0f (%.3f)" % (alg[1] / 1e6, alg[1] / (us * 1e-6) / 1e9, alg[1] / (us * 1e-6) / 1e9 / peak) if alg else "- | -"
    out.append("| `%s` | %d x %d | %d | %.1f | %.1f | %.0f (%.3f) | %s | %.1f | %.1f |" % (
        name[:60], val(r, "launch__grid_size"), val(r, "launch__block_size"), val(r, "launch__registers_per_thread"), us, dram / 1e6,
        ach, ach / peak, alg_s, val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")))
    if alg:
        aux[base] = {"ncu_us": round(us, 2), "algorithmic_bytes": alg[1], "frac_of_hbm_peak": round(alg[1] / (us * 1e-6) / 1e9 / peak, 4),
                     "what": alg[0]}
out += ["", "What each kernel has to move:", ""]
for k, (what, b) in ALG.items():
    out.append("* `%s`: %s — %.1f MB" % (k, what, b / 1e6))
out += ["", "Latency-bound by construction (a few hundred bytes of work): `k_pose_fwd`, `k_pose_bwd`, `k_median_select`, `k_stereo_T`."]
open(os.path.join(ROOT, "profiles", "%s_small_kernels.md" % tag), "w").write("\n".join(out) + "\n")
json.dump(aux, open(os.path.join(ROOT, "profiles", "%s_small_kernels.json" % tag), "w"), indent=1)
print("\n".join(out[6:30]))

# SASS of every kernel in the library
sass_dir = os.path.join(ROOT, "profiles", "%s_sass" % tag)
os.makedirs(sass_dir, exist_ok=True)
lib = os.path.join(ROOT, "unsupervised_pose_estimation_b200", "libvsl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, buf, seen = None, [], collections.Counter()
KEEP_PHOTO = "k_photometricINS_7TileCfgILi32ELi16ELi2ELi256EfLb0ELb0EEELb1E"   # the C1 kernel; the other instantiations are variants of it


def flush():
    if cur is None:
        return
    m = re.search(r"(k_[a-z0-9_]+)", cur)
    base = m.group(1) if m else "other"
    if base == "k_photometric" and KEEP_PHOTO not in cur:
        return
    seen[base] += 1
    path = os.path.join(sass_dir, "%s%s.sass" % (base, "" if seen[base] == 1 else "_%d" % seen[base]))
    open(path, "w").write("\n".join(buf) + "\n")


for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, buf = m.group(1), [line]
    elif cur is not None:
        if re.match(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", line):
            continue   # second encoding word of an instruction: keep one line per instruction
        buf.append(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", line))
flush()
print("SASS:", dict(seen))
