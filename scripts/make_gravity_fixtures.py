"""Pack the public gravity-coefficient files shipped with the reference (data/01_planetary) into
small .npz fixtures (n, m, C̄, S̄ in file order) so tests/bench need no /root/reference at run time.

  python scripts/make_gravity_fixtures.py [/root/reference/data/01_planetary]

JGM-3 (70x70, all 2 556 records) and the GRAIL JGGRX lunar field truncated to degree 80.
Only parsed numbers are stored; no reference source is copied.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nyx_b200.frames import IAU_EARTH_FRAME, IAU_MOON_FRAME  # noqa: E402
from nyx_b200.gravity import GravityFieldData  # noqa: E402

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/01_planetary")
out = Path(__file__).resolve().parent.parent / "data"
out.mkdir(exist_ok=True)


def dump(gd: GravityFieldData, name: str):
    ns, ms, cs, ss = [], [], [], []
    for n in range(gd.degree + 1):
        for m in range(min(n, gd.order) + 1):
            if n < 2 and gd.c_nm[n, m] == 0.0 and gd.s_nm[n, m] == 0.0:
                continue
            ns.append(n); ms.append(m); cs.append(gd.c_nm[n, m]); ss.append(gd.s_nm[n, m])
    np.savez_compressed(out / f"{name}.npz", n=np.array(ns, dtype=np.int32), m=np.array(ms, dtype=np.int32),
                        c=np.array(cs), s=np.array(ss))
    print(name, "records:", len(ns), "degree", gd.degree, "order", gd.order)


dump(GravityFieldData.from_cof(src / "JGM3.cof.gz", 70, 70, True, IAU_EARTH_FRAME), "jgm3_70x70")
dump(GravityFieldData.from_shadr(src / "Luna_jggrx_1500e_sha.tab.gz", 80, 80, True, IAU_MOON_FRAME), "luna_jggrx_80x80")
