"""Per-phase breakdown of a k_photometric ncu capture: the SASS page is split at the BAR.SYNC instructions
(phases are separated by __syncthreads) and executed instructions / stall samples are summed per segment.

    ncu -i gpurun_out/prof_photometric_<tag>.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
    python tools/phase_breakdown.py /tmp/sass.csv

Developer tool.
"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); it=hdr.index("Thread Instructions Executed")
stall_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
segs=[]; cur={"n":0,"inst":0,"samp":0,"thr":0,"ops":collections.Counter(),"st":collections.Counter(),"start":0}
tot=0; ts=0
for k,r in enumerate(rows[2:]):
    if len(r)<=it: continue
    op=r[ia].split()[0] if not r[ia].strip().startswith('@') else r[ia].split()[1]
    e=int(r[ie]); s=int(r[isamp]); t=int(r[it])
    cur["n"]+=1; cur["inst"]+=e; cur["samp"]+=s; cur["thr"]+=t; cur["ops"][op.split('.')[0]]+=e
    for c in stall_cols:
        cur["st"][hdr[c]]+=int(r[c] or 0)
    tot+=e; ts+=s
    if op.startswith("BAR"):
        segs.append(cur); cur={"n":0,"inst":0,"samp":0,"thr":0,"ops":collections.Counter(),"st":collections.Counter(),"start":k+1}
segs.append(cur)
print("total",tot,ts)
for i,s in enumerate(segs):
    top=", ".join(f"{o}:{100*c/max(1,s['inst']):.0f}" for o,c in s["ops"].most_common(8))
    st=", ".join(f"{o[6:]}:{100*c/max(1,s['samp']):.0f}" for o,c in s["st"].most_common(5))
    print(f"seg{i} sass[{s['start']}..+{s['n']}] inst {100*s['inst']/tot:5.1f}% samp {100*s['samp']/ts:5.1f}% lane {s['thr']/max(1,s['inst'])/32:.2f} | {top} | {st}")
