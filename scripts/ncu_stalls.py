"""Stall-reason breakdown per execution-count class from an ncu source page: python scripts/ncu_stalls.py rep [warp_rhs_count]"""
import csv, subprocess, collections, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:])); hdr = rows[0]
iex = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
data = []
for r in rows[1:]:
    try: data.append((int(r[iex]), {h: int(r[idx[h]] or 0) for h in reasons}, r[isrc]))
    except ValueError: pass
cnt = collections.Counter(ex for ex, _, s in data if ex > 0 and "SHFL" in s)
rhs = int(sys.argv[2]) if len(sys.argv) > 2 else max(cnt.items(), key=lambda kv: kv[1])[0]
tot = collections.Counter(); per = collections.defaultdict(collections.Counter)
for ex, st, _ in data:
    key = round(ex / rhs, 2)
    for h, v in st.items(): per[key][h] += v; tot[h] += v
all_s = sum(tot.values())
print("overall:", ", ".join(f"{h[6:]} {100*v/all_s:.1f}%" for h, v in tot.most_common(9)))
for key, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:6]:
    s = sum(c.values())
    print(f"  x{key:7.2f} ({100*s/all_s:5.1f}% of samples):", ", ".join(f"{h[6:]} {100*v/s:.0f}%" for h, v in c.most_common(7)))
