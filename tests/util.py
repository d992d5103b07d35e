import numpy as np


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|a|,|b|) over elements whose magnitude exceeds floor*max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(a), np.abs(b))
    mask = scale > floor * (np.abs(b).max() if b.size else 0.0)
    if not mask.any():
        return 0.0
    return float((np.abs(a - b)[mask] / scale[mask]).max())


def gamma_err(G, Gref):
    """Norm-wise per-depth error of a Gamma array [..., N, N, K]:
    max_k ||dG_k||_max / ||Gref_k||_max."""
    d = np.abs(G - Gref).max(axis=(-3, -2))
    s = np.abs(Gref).max(axis=(-3, -2))
    return float((d / s).max())


def compare_problems(p, q):
    """Errors of problem p's outputs against reference problem q."""
    out = {'I': rel_err(p.I, q.I), 'J': rel_err(p.J, q.J)}
    g, ge, r, n = 0.0, 0.0, 0.0, 0.0
    for a, b in zip(p.atoms, q.atoms):
        if not a.detailedStatic:
            g = max(g, gamma_err(a.Gamma, b.Gamma))
            ge = max(ge, rel_err(a.Gamma, b.Gamma, floor=1e-12))
        n = max(n, rel_err(a.n, b.n))
        for t, u in zip(a.trans, b.trans):
            r = max(r, rel_err(t.Rij, u.Rij, floor=1e-30), rel_err(t.Rji, u.Rji, floor=1e-30))
    out.update({'Gamma': g, 'GammaElem': ge, 'R': r, 'n': n})
    return out
