"""Device-resident Hermite block step (the steps either side of the force call,
idata::advance, src/amuse_ph4/src/idata.cc:832-870): g6x_hermite_init / _step / _evolve against the oracle
restatements of the i-predictor, the corrector + Aarseth step (pinned bit-exact against the reference in
tests/test_oracle.py) and against the unmodified ph4 integrator in CPU mode."""
import ctypes as C
import os

import numpy as np
import pytest

from amuse_b200 import plummer as P
from helpers import check_forces

pytestmark = pytest.mark.gpu


def _O():
    from oracle import oracle as O
    return O


def _state(g6, n):
    t = np.zeros(n); x = np.zeros((n, 3)); v = np.zeros((n, 3)); a = np.zeros((n, 3)); j = np.zeros((n, 3))
    g6.L.g6x_hermite_get_state(n, t.ctypes.data, x.ctypes.data, v.ctypes.data, a.ctypes.data, j.ctypes.data)
    return t, x, v, a, j


def _energy(O, m, t_sys, state, eps2):
    t, x, v, a, j = state
    pp, pv = O.predict(t_sys, t, x, v, a, j)
    f = O.force(pp, pv, m, pp, pv, eps2)
    return 0.5 * (m * (pv ** 2).sum(axis=1)).sum() + 0.5 * (m * f["pot"]).sum()


def _load(g6, n, seed, do_scale=False):
    m, x, v = P.new_plummer_model(n, seed=seed, do_scale=do_scale)
    ids = np.arange(1, n + 1, dtype=np.int32)
    g6.nj = 0
    g6.set_variant(0)
    g6.set_j_particles(ids, m, x, v)
    return m, x, v, ids


def test_init_and_block_steps_match_oracle(g6):
    O = _O()
    n, eta, eps2 = 2000, 0.14, 1e-4
    m, x, v, ids = _load(g6, n, 5)
    dt = np.zeros(n)
    g6.L.g6x_hermite_init(n, 0.0, eta, eps2, dt.ctypes.data)
    t, sx, sv, sa, sj = _state(g6, n)
    ref = O.force(x, v, m, x, v, eps2, scales=True)
    check_forces(dict(acc=sa, jerk=sj, pot=ref["pot"]), ref, what="init forces")
    assert np.array_equal(sx, x) and np.array_equal(sv, v) and np.all(t == 0.0)
    # first steps: the reference's rule applied to the forces the device stored
    dt_ref = O.initial_timestep(0.0, eta, sa, sj)
    assert np.mean(dt == dt_ref) > 0.999 and np.all((dt == dt_ref) | (dt == 2 * dt_ref) | (2 * dt == dt_ref))
    time = np.zeros(n)
    for step in range(12):
        tnext = (time + dt).min()
        ilist = np.nonzero(time + dt == tnext)[0].astype(np.int32)
        ni = len(ilist)
        before = _state(g6, n)
        new_dt = np.zeros(ni); pot = np.zeros(ni); nn = np.zeros(ni, dtype=np.int32)
        g6.L.g6x_hermite_step(n, ni, ilist, float(tnext), eta, eps2, np.ascontiguousarray(dt[ilist]), new_dt,
                              pot.ctypes.data, nn.ctypes.data)
        after = _state(g6, n)
        bt, bx, bv, ba, bj = before
        pp, pv = O.predict(tnext, bt, bx, bv, ba, bj)                    # j- and i-predictor (same expression)
        f = O.force(pp[ilist], pv[ilist], m, pp, pv, eps2, iid=ids[ilist], jid=ids, scales=True)
        check_forces(dict(acc=after[3][ilist], jerk=after[4][ilist], pot=pot), f, what="step %d forces" % step)
        assert np.array_equal(nn, ids[f["nn"]])
        # corrector on the forces the device used: positions/velocities to FP64 rounding, steps exact
        cp, cv, ct, cdt = O.correct(tnext, eta, bt[ilist], dt[ilist], ba[ilist], bj[ilist], after[3][ilist],
                                    after[4][ilist], pp[ilist], pv[ilist])
        assert np.abs(after[1][ilist] - cp).max() <= 1e-14 * np.abs(cp).max()
        assert np.abs(after[2][ilist] - cv).max() <= 1e-14 * np.abs(cv).max()
        assert np.array_equal(new_dt, cdt) and np.all(after[0][ilist] == tnext)
        others = np.setdiff1d(np.arange(n), ilist)
        for k in range(5):
            assert np.array_equal(after[k][others], before[k][others])   # nobody else was touched
        time[ilist] = tnext
        dt[ilist] = new_dt


def test_blocks_larger_than_npipes_go_in_one_pass(g6):
    """A synchronised (re)start makes every particle active: the block exceeds g6_npipes() and must still see
    only PREDICTED j (no particle corrected before all forces are in), like idata::advance."""
    O = _O()
    n, eta, eps2 = 20000, 0.14, 1e-4
    assert n > g6.npipes
    m, x, v, ids = _load(g6, n, 8)
    dt = np.zeros(n)
    g6.L.g6x_hermite_init(n, 0.0, eta, eps2, dt.ctypes.data)
    t, sx, sv, sa, sj = _state(g6, n)
    samp = np.arange(0, n, 97)
    ref = O.force(x[samp], v[samp], m, x, v, eps2, iid=ids[samp], jid=ids, scales=True)
    check_forces(dict(acc=sa[samp], jerk=sj[samp], pot=ref["pot"]), ref, what="init forces, n > npipes")
    assert np.mean(dt == O.initial_timestep(0.0, eta, sa, sj)) > 0.999
    # force every particle to step together to t = 2^-12 (all steps are powers of two >= that? use the smallest)
    tnext = float(dt.min())
    ilist = np.arange(n, dtype=np.int32)
    new_dt = np.zeros(n)
    g6.L.g6x_hermite_step(n, n, ilist, tnext, eta, eps2, dt, new_dt, None, None)
    after = _state(g6, n)
    pp, pv = O.predict(tnext, t, sx, sv, sa, sj)
    f = O.force(pp[samp], pv[samp], m, pp, pv, eps2, iid=ids[samp], jid=ids, scales=True)
    check_forces(dict(acc=after[3][samp], jerk=after[4][samp], pot=f["pot"]), f, what="all-particle block")
    cp, cv, ct, cdt = O.correct(tnext, eta, t, dt, sa, sj, after[3], after[4], pp, pv)
    assert np.abs(after[1] - cp).max() <= 1e-14 * np.abs(cp).max()
    assert np.array_equal(new_dt, cdt)


def test_evolve_tracks_ph4_cpu_mode(g6):
    O = _O()
    if not O.ref_available():
        pytest.skip("oracle/_ref not built")
    n, eta, eps2, t_end = 1024, 0.14, 1e-4, 0.25
    m, x, v, ids = _load(g6, n, 1, do_scale=True)
    g6.L.g6x_hermite_init(n, 0.0, eta, eps2, None)
    e0 = _energy(O, m, 0.0, _state(g6, n), eps2)
    stats = np.zeros(4)
    g6.L.g6x_hermite_evolve(n, t_end, eta, eps2, 0, stats)
    e1 = _energy(O, m, stats[0], _state(g6, n), eps2)
    cpu = O.ref_evolve(m, x, v, eps2, eta, t_end)
    print("device-resident Hermite N=%d: %d block steps, %d particle steps, %.3f s (%.1f us per block step), dE/E %.2e | "
          "ph4 CPU mode: %d block steps, %d particle steps, %.3f s, dE/E %.2e" % (
              n, stats[1], stats[2], stats[3], 1e6 * stats[3] / stats[1], abs((e1 - e0) / e0),
              cpu["block_steps"], cpu["particle_steps"], cpu["seconds"], abs((cpu["E1"] - cpu["E0"]) / cpu["E0"])))
    assert stats[0] >= t_end
    assert abs(e0 - cpu["E0"]) < 2e-7 * abs(cpu["E0"])
    assert abs(e1 - e0) < 2e-5 * abs(e0)
    assert abs(stats[2] - cpu["particle_steps"]) < 0.05 * cpu["particle_steps"]
    assert abs(stats[1] - cpu["block_steps"]) < 0.05 * cpu["block_steps"]
