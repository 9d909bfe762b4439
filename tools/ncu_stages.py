"""Stall samples and warp instructions of an ncu report grouped by source-line RANGES of one file.
usage: python tools/ncu_stages.py report.ncu-rep file-substring name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys, collections
rep, want = sys.argv[1], sys.argv[2]
ranges = []
for spec in sys.argv[3:]:
    n, r = spec.split(':'); lo, hi = r.split('-'); ranges.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; f = None
agg = collections.defaultdict(lambda: collections.Counter())
for rec in csv.reader(out.splitlines()):
    if not rec: continue
    if rec[0] == "File Path": f = rec[1]; continue
    if rec[0] == "Line No": hdr = rec; continue
    if hdr is None or len(rec) != len(hdr) or not rec[0].isdigit(): continue
    d = dict(zip(hdr[4:], rec[4:]))
    try: inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError: continue
    line = int(rec[0]); name = "other:" + f.split('/')[-1]
    if want in f:
        name = "unassigned"
        for n, lo, hi in ranges:
            if lo <= line <= hi: name = n; break
    c = agg[name]; c["inst"] += inst; c["samp"] += samp
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: c[k[6:]] += int(v)
            except ValueError: pass
ti = sum(c["inst"] for c in agg.values()); ts = sum(c["samp"] for c in agg.values())
print(f"warp instructions {ti}, samples {ts}")
for n, c in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
    st = " ".join(f"{k}={v}" for k, v in c.most_common() if k not in ("inst", "samp") and v > 0.03 * c["samp"])
    print(f"{n:22s} inst {c['inst']/ti*100:5.1f}%  samples {c['samp']/ts*100:5.1f}%  | {st}")
