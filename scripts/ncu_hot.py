"""Aggregate the ncu source page (SASS view): instruction mix and the hottest address ranges.
usage: python scripts/ncu_hot.py file.ncu-rep"""
import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0; mix = collections.Counter(); smp = collections.Counter(); data = []
for r in rows[1:]:
    try: ex = int(r[iex]); sm = int(r[ismp])
    except ValueError: continue
    op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
    op = op.split(".")[0]
    mix[op] += ex; smp[op] += sm; tot += ex; data.append((r[ia], r[isrc], ex, sm))
print("total warp-instructions:", tot)
for op, c in mix.most_common(25):
    print(f"  {op:10s} {c:14d} {100*c/tot:6.2f}%   samples {smp[op]}")
# hottest 16-instruction windows by samples
print("hottest instructions by stall samples:")
for a, s, ex, sm in sorted(data, key=lambda t: -t[3])[:25]:
    print(f"  {a} {sm:7d} {ex:12d}  {s[:90]}")
