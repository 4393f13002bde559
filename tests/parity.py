"""Error metrics shared by the parity tests.

North-star tolerance: per-iteration H and b within 1e-6 relative of the double CPU result;
converged pose within 1e-6 m / 1e-6 rad at the same iteration count.  Entries of b (and
off-diagonal entries of H) cancel towards zero, so "relative" is taken against the scale of the
quantity: max|H| for H, max(|b|_max, 1e-12 * max|H|) ... concretely:
    err_H = max|H - H_ref| / max|H_ref|
    err_g = max|g - g_ref| / max(max|g_ref|, 1e-9 * max|H_ref|)
    err_c = |c - c_ref| / max(|c_ref|, 1e-300)
"""
import numpy as np

TOL = 1e-6  # BASELINE.json north_star


def rel_errors(H, g, cost, H_ref, g_ref, cost_ref):
    hs = max(np.max(np.abs(H_ref)), 1e-300)
    gs = max(np.max(np.abs(g_ref)), 1e-9 * hs, 1e-300)
    return (np.max(np.abs(H - H_ref)) / hs, np.max(np.abs(g - g_ref)) / gs,
            abs(cost - cost_ref) / max(abs(cost_ref), 1e-300))


def assert_sums_close(H, g, cost, H_ref, g_ref, cost_ref, tol=TOL):
    eh, eg, ec = rel_errors(H, g, cost, H_ref, g_ref, cost_ref)
    assert eh < tol and eg < tol and ec < tol, (eh, eg, ec)


def rotation_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    # small-angle safe: use the skew part
    d = Ra.T @ Rb
    v = np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]]) / 2.0
    return float(np.arctan2(np.linalg.norm(v), c))
