"""Parity metrics and tolerances (north-star).

Per-particle vector-norm relative errors against ph4's FP64 CPU loop:
    e_acc = |a - a_ref| / |a_ref|,  e_jerk likewise,  e_pot = |p - p_ref| / |p_ref|.

Tolerance written here and asserted by every parity test: the north-star's 1e-6 on the per-particle MAXIMUM of
all three, for every particle -- field probes inside the cluster and members next to the cluster centre, whose
acc and jerk are cancelling sums, included.  (Round 1 held jerk to 1e-5 and gave cancelling acc sums a conditioned
bound; since round 2 the library evaluates every pair closer than sqrt(K) x the i-particle's nearest-neighbour
distance in FP64 with the reference's expression tree, K = 16 by default, so the pairs that dominate a cancelling
sum no longer carry the 2^-24 rounding of dx: tools/emulate_v2.py, DESIGN.md "Accuracy".)
When the oracle provides the condition scales S_i = sum_j |a_ij| and sum_j |jerk_ij|, the backward-stable bound
|d_i| <= 1e-6 * S_i is asserted as well (it is the weaker statement for every particle).
"""
import numpy as np

TOL = 1e-6          # acc / jerk / pot: per-particle maximum
TOL_JERK_MAX = 1e-6
TOL_JERK_SCALED = 1e-6


def rel_vec_err(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def check_forces(got, ref, tol=TOL, what=""):
    """Asserts the tolerances above; returns (max e_acc, max e_jerk, max e_pot)."""
    ea = rel_vec_err(got["acc"], ref["acc"])
    ej = rel_vec_err(got["jerk"], ref["jerk"])
    ep = rel_err(got["pot"], ref["pot"])
    assert ea.max() <= tol, "%s acc rel err %.3e" % (what, ea.max())
    assert ep.max() <= tol, "%s pot rel err %.3e" % (what, ep.max())
    assert ej.max() <= min(tol, TOL_JERK_MAX) if tol < TOL else ej.max() <= max(tol, TOL_JERK_MAX), "%s jerk rel err max %.3e" % (what, ej.max())
    if "sjerk" in ref:
        sc = np.linalg.norm(got["jerk"] - ref["jerk"], axis=1) / ref["sjerk"]
        assert sc.max() <= TOL_JERK_SCALED, "%s jerk err / sum|terms| %.3e" % (what, sc.max())
        sa = np.linalg.norm(got["acc"] - ref["acc"], axis=1) / ref["sacc"]
        assert sa.max() <= TOL_JERK_SCALED, "%s acc err / sum|terms| %.3e" % (what, sa.max())
    return ea.max(), ej.max(), ep.max()


def error_report(got, ref):
    ea = rel_vec_err(got["acc"], ref["acc"])
    ej = rel_vec_err(got["jerk"], ref["jerk"])
    ep = rel_err(got["pot"], ref["pot"])
    f = lambda e: "max %.2e p99.9 %.2e p99 %.2e median %.2e" % (e.max(), np.percentile(e, 99.9), np.percentile(e, 99), np.median(e))
    return "acc[%s] jerk[%s] pot[%s]" % (f(ea), f(ej), f(ep))


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
