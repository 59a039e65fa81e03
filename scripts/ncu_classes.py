"""Per-RHS instruction classes from an ncu source page: python scripts/ncu_classes.py rep [warp_rhs_count]"""
import csv, subprocess, collections, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:])); hdr = rows[0]
iex = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
data = []
for r in rows[1:]:
    try: data.append((int(r[iex]), int(r[ismp]), r[isrc]))
    except ValueError: pass
cnt = collections.Counter(ex for ex, _, s in data if ex > 0 and "SHFL" in s)
rhs = int(sys.argv[2]) if len(sys.argv) > 2 else max(cnt.items(), key=lambda kv: kv[1])[0]
print("warp-RHS count:", rhs)
cls = collections.Counter(); n = collections.Counter(); smp = collections.Counter(); fp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for ex, sm, src in data:
    key = round(ex / rhs, 2); cls[key] += ex; n[key] += 1; smp[key] += sm
    toks = src.split()
    op = (toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')).split('.')[0]
    ops[key][op] += ex
    if op in ("DFMA", "DMUL", "DADD"): fp[key] += ex
tot = sum(cls.values()); ts = sum(smp.values())
for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:8]:
    top = ", ".join(f"{o} {c/rhs:.0f}" for o, c in ops[k].most_common(8))
    print(f"  x{k:7.2f}: {n[k]:5d} static, {v/rhs:8.1f} dyn/RHS ({100*v/tot:5.1f}%), fp64 {fp[k]/rhs:7.1f}, samples {100*smp[k]/ts:5.1f}%  [{top}]")
print("total per RHS %.0f, fp64 per RHS %.0f" % (tot / rhs, sum(fp.values()) / rhs))
