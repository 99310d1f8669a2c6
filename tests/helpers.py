"""Parity metrics: per-particle vector-norm relative errors."""
import numpy as np

TOL = 1e-6  # north-star: relative acc/jerk/pot error <= 1e-6 vs ph4's FP64 CPU loop


def rel_vec_err(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def check_forces(got, ref, tol=TOL, what=""):
    ea = rel_vec_err(got["acc"], ref["acc"]).max()
    ej = rel_vec_err(got["jerk"], ref["jerk"]).max()
    ep = rel_err(got["pot"], ref["pot"]).max()
    assert ea <= tol, "%s acc rel err %.3e" % (what, ea)
    assert ej <= tol, "%s jerk rel err %.3e" % (what, ej)
    assert ep <= tol, "%s pot rel err %.3e" % (what, ep)
    return ea, ej, ep


def check_nn(got_id, ref_j, jid, ipos, pred_pos, tie_tol=1e-6):
    """Nearest-neighbour ids must be exact, except at distance ties: a mismatch is accepted
    only if the FP64 distance of the returned neighbour is within tie_tol (relative, in r^2)
    of the true minimum.  Returns the number of tie-mismatches."""
    ref_id = np.where(ref_j >= 0, jid[np.maximum(ref_j, 0)], -1)
    bad = np.nonzero(got_id != ref_id)[0]
    if len(bad) == 0:
        return 0
    id2j = {int(v): k for k, v in enumerate(jid)}
    for i in bad:
        assert int(got_id[i]) in id2j, "i=%d returned unknown id %d (ref %d)" % (i, got_id[i], ref_id[i])
        jg = id2j[int(got_id[i])]
        r2g = ((pred_pos[jg] - ipos[i]) ** 2).sum()
        r2r = ((pred_pos[ref_j[i]] - ipos[i]) ** 2).sum()
        assert abs(r2g - r2r) <= tie_tol * r2r, "i=%d nn %d (r2 %.17g) vs ref %d (r2 %.17g): not a tie" % (
            i, got_id[i], r2g, ref_id[i], r2r)
    return len(bad)
