#!/usr/bin/env python3
"""Static instruction mix of a SASS address range (no GPU needed): cuobjdump -sass <obj> | this script <mangled-name-fragment> <start> <end>.
Used for DESIGN.md §11: the pair-iteration of K2's column walk (nyxb_k_coop<8,1,true>), with and without the column-switch block."""
import collections
import re
import subprocess
import sys


def main():
    obj, frag, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, rows = None, []
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur and frag in cur:
            addr = int(m.group(1), 16)
            if lo <= addr <= hi:
                rows.append((addr, m.group(2).strip()))
    mix = collections.Counter()
    for _, ins in rows:
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        base = op.split(".")[0]
        cls = ("FP64" if base in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU") else
               "LDS" if base == "LDS" else
               "LDC" if base in ("LDC", "LDCU") else
               "branch/sync" if base in ("BRA", "BSSY", "BSYNC", "WARPSYNC", "EXIT") else
               "int/move")
        mix[cls] += 1
    print(f"{frag}: {len(rows)} instructions in [{lo:#x}, {hi:#x}]")
    for k, v in mix.most_common():
        print(f"  {k:12s} {v}")


if __name__ == "__main__":
    main()
