"""Independent arbiters for the CPU oracle (TEST INFRASTRUCTURE): textbook formulas evaluated with mpmath at 40 digits, scipy's
root finder / interpolator — nothing here shares code, recursions or operation order with oracle/ or nyx_b200/.

Spherical-harmonic acceleration (Vallado, *Fundamentals of Astrodynamics*, eq. 8-19/8-27; Montenbruck & Gill eq. 3.27-3.33):
    U = (mu / r) sum_{n>=1} sum_{m<=n} (R / r)^n  Pbar_nm(sin phi) (Cbar_nm cos m lam + Sbar_nm sin m lam)
    a_r   = -(mu / r^2) sum (n + 1) (R/r)^n Pbar_nm (C cos + S sin)
    a_phi =  (mu / r^2) sum (R/r)^n dPbar_nm/dphi (C cos + S sin),   dP_nm/dphi = P_{n,m+1} - m tan(phi) P_nm
    a_lam =  (mu / (r^2 cos phi)) sum (R/r)^n m Pbar_nm (S cos - C sin)
with the associated Legendre FUNCTIONS from their closed (hypergeometric) form, `mpmath.legenp`, no recursion; geodesy's
normalisation Pbar = sqrt((2 - delta_0m)(2n + 1)(n - m)! / (n + m)!) P and no Condon-Shortley phase."""
import mpmath as mp


def mp_harmonic_accel(c_nm, s_nm, degree, order, mu, r_eq, rb, dps=40):
    """Non-central acceleration [km/s^2] at the body-fixed position rb [km] (three floats) for normalised coefficients
    c_nm[n][m], s_nm[n][m]; returns three Python floats rounded from `dps`-digit arithmetic."""
    mp.mp.dps = dps
    x, y, z = (mp.mpf(float(v)) for v in rb)
    r = mp.sqrt(x * x + y * y + z * z)
    sphi = z / r
    cphi = mp.sqrt(x * x + y * y) / r
    tphi = sphi / cphi
    lam = mp.atan2(y, x)
    mu, r_eq = mp.mpf(float(mu)), mp.mpf(float(r_eq))
    ar = aphi = alam = mp.mpf(0)
    for n in range(1, degree + 1):
        rn = (r_eq / r) ** n
        # P_n^m(sin phi), m = 0..n+1, closed form; mpmath's type-2 function carries the Condon-Shortley phase (-1)^m: remove it
        P = [((-1) ** m) * mp.legenp(n, m, sphi) for m in range(0, n + 1)] + [mp.mpf(0)]
        for m in range(0, min(n, order) + 1):
            c, s = mp.mpf(float(c_nm[n][m])), mp.mpf(float(s_nm[n][m]))
            if c == 0 and s == 0:
                continue
            norm = mp.sqrt((2 if m else 1) * (2 * n + 1) * mp.factorial(n - m) / mp.factorial(n + m))
            cs = c * mp.cos(m * lam) + s * mp.sin(m * lam)
            sc = s * mp.cos(m * lam) - c * mp.sin(m * lam)
            dP = P[m + 1] - m * tphi * P[m]
            ar -= (n + 1) * rn * norm * P[m] * cs
            aphi += rn * norm * dP * cs
            alam += rn * m * norm * P[m] * sc
    k = mu / (r * r)
    ar, aphi, alam = k * ar, k * aphi, k * alam / cphi
    # spherical -> Cartesian (unit vectors e_r, e_phi, e_lam)
    cl, sl = mp.cos(lam), mp.sin(lam)
    ax = ar * cphi * cl - aphi * sphi * cl - alam * sl
    ay = ar * cphi * sl - aphi * sphi * sl + alam * cl
    az = ar * sphi + aphi * cphi
    return float(ax), float(ay), float(az)


def sun_visible_fraction(r_ls, r_body, d, n=1500):
    """Fraction of a disk of angular radius r_ls (light source) NOT covered by a disk of angular radius r_body whose centre is the
    angle d away, by brute-force area quadrature on a polar grid (small-angle, planar geometry: what `occultation` models)."""
    import numpy as np

    rr = (np.arange(n) + 0.5) / n * r_ls
    th = (np.arange(2 * n) + 0.5) / (2 * n) * 2 * np.pi
    R, T = np.meshgrid(rr, th, indexing="ij")
    px, py = R * np.cos(T), R * np.sin(T)
    covered = (px - d) ** 2 + py ** 2 < r_body ** 2
    w = R   # area element r dr dtheta (constant factors cancel in the ratio)
    return 1.0 - float((w * covered).sum() / w.sum())
