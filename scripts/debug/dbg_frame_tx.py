"""Debug (GPU box): integration_frame + transposed kernel, which trajectories come back wrong and under which variation."""
import sys
import numpy as np
sys.path.insert(0, ".")
import nyx_b200 as nb
from nyx_b200.frames import MOON, SUN
from tests.util import S
from tests.test_gpu_frames_fields import _cislunar, _fields
from oracle import pyoracle


def run(n, stagger, cap, kernel, deg=12):
    alm, (st_e, cs, ep) = _cislunar(n, seed=9)
    if stagger:
        ep = ep + (np.arange(n, dtype=np.int64) % 4) * 1200 * S
    prop = nb.Propagator.default(_fields(deg), mode=nb.MODE_FAST)
    prop.opts.integration_frame = nb.EARTH_J2000
    moon = alm.bodies[alm.body_index(MOON)]
    st_m = st_e.copy()
    st_m[:3] -= moon.position(0)[:, None]
    st_m[3:6] -= ((moon.position(S) - moon.position(-S)) / 2.0)[:, None]
    eng = prop.engine(nb.MOON_J2000, alm)
    eng.set_kernel(kernel)
    end = 6 * 3600 * S
    res = eng.propagate_batch(st_m, cs, ep, end, traj_capacity=cap)
    out, oep, det, status = res[:4]
    packed, opts_c = prop.lower(nb.MOON_J2000, alm)
    ref, rep, rdet, rstatus = pyoracle.propagate_batch(packed.c, opts_c, st_m, cs, ep, end)[:4]
    d = np.abs(out[:3] - ref[:3]).max(axis=0)
    bad = np.nonzero(d > 1e-6)[0]
    print(f"n={n} stagger={stagger} cap={cap} kernel={kernel} deg={deg}: max dr {d.max():.3e}, bad idx {bad.tolist()[:20]}, status {np.unique(status)}, steps {det['n_steps'][:4]} vs {rdet['n_steps'][:4]}")
    for i in bad[:3]:
        print("   i", i, "out", out[:6, i], "ref", ref[:6, i], "ep", oep[i], rep[i])


for args in [(48, True, 600, nb.KERNEL_TRANSPOSED), (48, True, 0, nb.KERNEL_TRANSPOSED), (48, False, 600, nb.KERNEL_TRANSPOSED),
             (32, True, 600, nb.KERNEL_TRANSPOSED), (64, True, 600, nb.KERNEL_TRANSPOSED), (48, True, 600, nb.KERNEL_COOP),
             (48, True, 600, nb.KERNEL_TRANSPOSED, 20)]:
    try:
        run(*args)
    except Exception as e:
        print(args, "EXC", repr(e)[:300])
