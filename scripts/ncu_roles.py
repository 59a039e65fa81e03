"""Warp-state samples of a warp-specialised kernel split by role, from an ncu source page (SASS view):
python scripts/ncu_roles.py rep.ncu-rep [split_instruction_index]
The split is the SASS index of the first helper instruction (default: found from the largest backward gap: the walker code ends with
EXIT before the helper code starts)."""
import csv, subprocess, collections, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rows = list(csv.reader(out[start:])); hdr = rows[0]
iex = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
isamp = hdr.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in hdr else None
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
data = []
for r in rows[1:]:
    try: data.append((int(r[iex] or 0), {h: int(r[idx[h]] or 0) for h in reasons}, r[isrc]))
    except ValueError: pass
exits = [i for i, d in enumerate(data) if d[2].strip().startswith("EXIT")]
split = int(sys.argv[2]) if len(sys.argv) > 2 else (exits[0] + 1 if exits else len(data) // 2)
print(f"{len(data)} SASS instructions, split at {split} (EXITs at {exits[:6]})")
tot_all = sum(sum(d[1].values()) for d in data)
for name, seg in (("walkers", data[:split]), ("helpers", data[split:])):
    tot = collections.Counter()
    for ex, st, _ in seg:
        for h, v in st.items(): tot[h] += v
    s = sum(tot.values())
    print(f"{name}: instrs {len(seg)} executed {sum(d[0] for d in seg)} samples {s} ({100*s/max(tot_all,1):.1f}%)  " +
          ", ".join(f"{h[6:]} {100*v/max(s,1):.0f}%" for h, v in tot.most_common(8)))
top = sorted(range(len(data)), key=lambda i: -sum(data[i][1].values()))[:14]
for i in top:
    st = data[i][1]; s = sum(st.values())
    print(f"  #{i:5d} {s:7d} x{data[i][0]:>10d}  {data[i][2][:60]:60s} " + ", ".join(f"{h[6:]} {v}" for h, v in collections.Counter(st).most_common(3)))
