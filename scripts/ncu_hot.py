"""Hot spots of one kernel from an .ncu-rep captured with --import-source on:
    python scripts/ncu_hot.py rep.ncu-rep <kernel-name> [top]
prints the top stall-sampled SASS instructions (in program order), 100-instruction bucket totals and
the per-stall-reason totals."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern,
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
h = rows[hi]
si, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


seg, reasons = [], {}
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"):
        if seg and len(r) > 1 and r[0] == "Address":
            break                       # next launch of the same kernel
        continue
    seg.append((I(r[si]), r[1].strip()[:80], I(r[ie])))
    for c in stall_cols:
        reasons[h[c]] = reasons.get(h[c], 0) + I(r[c])
print("instructions", len(seg), "samples", sum(s[0] for s in seg), "warp-inst executed", sum(s[2] for s in seg))
top = sorted(range(len(seg)), key=lambda i: -seg[i][0])[:top_n]
for i in sorted(top):
    print(f"{i:5d} {seg[i][0]:6d} {seg[i][2]:8d}  {seg[i][1]}")
print("buckets of 100 instructions: (samples, executed)")
for b in range(0, len(seg), 100):
    print(f"  {b:5d} {sum(s[0] for s in seg[b:b+100]):6d} {sum(s[2] for s in seg[b:b+100]):9d}")
print({k: v for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:8]})
