"""GPU parity tests: the CUDA path, called through the g6 C ABI, against the CPU oracle
(tests/golden fixtures made from the real reference + the oracle restatement on seeded inputs).

Tolerance (north-star): per-particle vector-norm relative error of acc and jerk and relative
error of pot <= 1e-6 versus ph4's double-precision CPU loop; nearest-neighbour ids exact except
at distance ties (relative r^2 gap < 1e-6)."""
import ctypes as C
import os

import numpy as np
import pytest

from amuse_b200 import plummer as P
from helpers import TOL, check_forces, check_nn, error_report, rel_err, rel_vec_err

pytestmark = pytest.mark.gpu

VARIANTS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]   # see DESIGN.md; 0 = auto; 9, 10 = speculative kernels


def _O():
    from oracle import oracle as O
    return O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _fresh(g6, ids, m, x, v, acc=None, jerk=None, tj=None, ti=0.0):
    """(Re)load the j-memory from address 0 like jdata::initialize_gpu (gpu.cc:36-92)."""
    g6.nj = 0
    g6.set_variant(0)
    g6.set_j_particles(ids, m, x, v, acc=acc, jerk=jerk, tj=tj)
    g6.set_ti(ti)


@pytest.mark.parametrize("variant", [0] + VARIANTS)
@pytest.mark.parametrize("name", ["ph4_plummer1k_eps1e-4.npz", "ph4_plummer1k_eps0.npz"])
def test_golden_full_sweep_all_variants(g6, golden_dir, name, variant):
    g = _load(golden_dir, name)
    _fresh(g6, g["ids"], g["mass"], g["pos"], g["vel"])
    g6.set_variant(variant)
    out = g6.calc(g["ids"], g["pos"], g["vel"], float(g["eps2"]))
    g6.set_variant(0)
    print("%s variant %d: %s" % (name, variant, error_report(out, g)))
    check_forces(out, g, what="%s v%d" % (name, variant))
    check_nn(out["nn"], g["nn"], g["ids"], g["pos"], g["pos"])
    out2 = g6.calc(g["ids"], g["pos"], g["vel"], float(g["eps2"]), want_nn=False)   # lasthalf path
    check_forces(out2, g, what="%s v%d lasthalf" % (name, variant))
    assert rel_vec_err(out2["acc"], out["acc"]).max() < 1e-6


def test_refine_switch(g6, golden_dir):
    """g6x_set_refine(0) uses the raw MUFU.RSQ: same answers to ~1e-6, larger error than the default."""
    g = _load(golden_dir, "ph4_plummer1k_eps1e-4.npz")
    _fresh(g6, g["ids"], g["mass"], g["pos"], g["vel"])
    a = g6.calc(g["ids"], g["pos"], g["vel"], 1e-4)
    g6.L.g6x_set_refine(0)
    try:
        b = g6.calc(g["ids"], g["pos"], g["vel"], 1e-4)
    finally:
        g6.L.g6x_set_refine(1)
    print("refined : %s" % error_report(a, g))
    print("raw rsq : %s" % error_report(b, g))
    assert rel_vec_err(a["acc"], g["acc"]).max() < rel_vec_err(b["acc"], g["acc"]).max()
    assert rel_vec_err(b["acc"], g["acc"]).max() < 3e-6


def test_golden_predictor_and_block_step_forces(g6, golden_dir):
    g = _load(golden_dir, "ph4_predict_force_512.npz")
    n = len(g["mass"])
    ids = np.arange(n, dtype=np.int32)
    _fresh(g6, ids, g["mass"], g["pos"], g["vel"], acc=g["acc0"], jerk=g["jerk0"], tj=g["tj"], ti=float(g["t"]))
    il = g["ilist"]
    out = g6.calc(ids[il], g["ipos"], g["ivel"], float(g["eps2"]))
    pp, pv = g6.read_predicted(n)
    # predicted positions carry double-single precision (2^-48 relative), velocities FP32
    assert np.abs(pp - g["pred_pos"]).max() <= 2e-14 * max(1.0, np.abs(g["pred_pos"]).max())
    assert rel_vec_err(pv, g["pred_vel"]).max() <= 1.2e-7
    check_forces(out, g, what="block-step")
    check_nn(out["nn"], g["nn"], ids, g["ipos"], g["pred_pos"])


@pytest.mark.parametrize("ni,nj", [(1, 1), (1, 2), (3, 255), (5, 256), (7, 257), (33, 1000), (64, 513),
                                    (257, 300), (1025, 2049), (2100, 700)])
def test_ragged_sizes(g6, ni, nj):
    O = _O()
    m, x, v = P.new_plummer_model(max(nj, 2), seed=20 + nj % 7)
    m, x, v = m[:nj], x[:nj], v[:nj]
    ids = np.arange(100, 100 + nj, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    rnd = np.random.RandomState(ni)
    ipos = rnd.normal(size=(ni, 3)) * 0.7
    ivel = rnd.normal(size=(ni, 3)) * 0.5
    iid = -np.ones(ni, dtype=np.int32)           # field points (interface.cc:840-919)
    k = min(ni, nj) // 2                          # ... and some genuine members
    ipos[:k], ivel[:k], iid[:k] = x[:k], v[:k], ids[:k]
    out = g6.calc(iid, ipos, ivel, 1e-4, nj=nj)
    ref = O.force(ipos, ivel, m, x, v, 1e-4, iid=iid, jid=ids, scales=(ni > 1 and nj > 1))
    if nj == 1 and k == 1:
        assert np.all(out["acc"][0] == 0) and out["pot"][0] == 0 and out["nn"][0] == -1
        out = {kk: vv[1:] for kk, vv in out.items()}
        ref = {kk: vv[1:] for kk, vv in ref.items()}
        ipos = ipos[1:]
    if len(ipos):
        check_forces(out, ref, what="ragged %dx%d" % (ni, nj))
        check_nn(out["nn"], ref["nn"], ids, ipos, x)


def test_nj_prefix_argument(g6):
    """nj selects the address prefix [0, nj) (gpu.cc:394 passes localnj)."""
    O = _O()
    m, x, v = P.new_plummer_model(1500, seed=5)
    ids = np.arange(1500, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    for nj in (1500, 777, 256, 10):
        out = g6.calc(ids[:50], x[:50], v[:50], 0.0, nj=nj)
        ref = O.force(x[:50], v[:50], m[:nj], x[:nj], v[:nj], 0.0, iid=ids[:50], jid=ids[:nj])
        check_forces(out, ref, what="prefix %d" % nj)
        check_nn(out["nn"], ref["nn"], ids, x[:50], x[:nj])


def test_massless_coincident_and_unset_slots(g6):
    O = _O()
    m, x, v = P.new_plummer_model(600, seed=9)
    m = m.copy(); x = x.copy()
    m[10:20] = 0.0                      # massless j are skipped, also as neighbours (idata.cc:208)
    x[30] = x[31]                       # two different particles at the same place (r2 = 0)
    ids = np.arange(1, 601, dtype=np.int32)
    for eps2 in (0.0, 1e-4):
        _fresh(g6, ids, m, x, v)
        out = g6.calc(ids, x, v, eps2)
        ref = O.force(x, v, m, x, v, eps2, iid=ids, jid=ids)
        assert np.all(np.isfinite(out["acc"])) and np.all(np.isfinite(out["jerk"])) and np.all(np.isfinite(out["pot"]))
        check_forces(out, ref, what="degenerate eps2=%g" % eps2)
        check_nn(out["nn"], ref["nn"], ids, x, x)
    # unset addresses inside the prefix behave like massless particles (fresh device memory)
    g6.close()
    assert g6.L.g6_open_(C.byref(g6.cid)) == 0
    g6.nj = 0
    g6.set_j_particles(ids[:300], m[:300], x[:300], v[:300], address0=0)
    g6.set_j_particles(ids[300:], m[300:], x[300:], v[300:], address0=20000)
    out = g6.calc(ids[:40], x[:40], v[:40], 1e-4, nj=20300)
    ref = O.force(x[:40], v[:40], m, x, v, 1e-4, iid=ids[:40], jid=ids)
    check_forces(out, ref, what="holes")


@pytest.mark.parametrize("variant", [9, 10])
def test_speculative_kernel_equals_masked_kernel(g6, variant):
    """The speculative kernels (mask-free groups + verification) must return what the
    masked pair rule returns on inputs that exercise every fallback: ids shuffled against addresses
    (every tile's id range contains every i), coincident particles and a self pair that is NOT at
    r = 0 (host-predicted i), massless particles, a ragged last tile, and nearest neighbours that sit
    in the first, a middle and the last group of the j range."""
    O = _O()
    rng = np.random.RandomState(77)
    n = 3000                                   # 11 full tiles + a ragged one
    m, x, v = P.new_plummer_model(n, seed=21)
    m = m.copy(); x = x.copy(); v = v.copy()
    m[100:110] = 0.0
    x[500] = x[1500]                           # coincident pair across tiles
    x[1234] = x[1233]                          # coincident pair inside one group
    xn = x.copy()                              # near-coincident variants (r2 < 2^-52 but not 0): only with
    xn[2999] = xn[0] + 1e-9                    # softening, where such a pair carries no weight (without it
    xn[1234] = xn[1233] + np.array([3e-9, 0, 0])  # its force is ~1e11 and not representable to 1e-6)
    for ids in (np.arange(1, n + 1, dtype=np.int32), rng.permutation(n).astype(np.int32) + 5):
        for eps2 in (0.0, 1e-4):
            xj = x if eps2 == 0.0 else xn
            _fresh(g6, ids, m, xj, v)
            # with softening i is also "predicted on the host": the self pair is not exactly at r = 0
            xi = xj if eps2 == 0.0 else xj + 1e-12
            ref = O.force(xi, v, m, xj, v, eps2, iid=ids, jid=ids)
            g6.set_variant(7)
            masked = g6.calc(ids, xi, v, eps2)
            g6.set_variant(variant)
            out = g6.calc(ids, xi, v, eps2)
            out_nonn = g6.calc(ids, xi, v, eps2, want_nn=False)
            g6.set_variant(0)
            check_forces(out, ref, what="speculative v%d eps2=%g" % (variant, eps2))
            check_nn(out["nn"], ref["nn"], ids, xi, xj)
            assert np.array_equal(out["nn"], masked["nn"])
            # same arithmetic per pair, different summation grouping only (FP32 partial sums over 32 pairs here,
            # 16 in the masked kernel): both are within 1e-6 of the oracle (checked above for `out`), and a wrong
            # mask or a lost group would show as an O(1) difference
            assert rel_vec_err(out["acc"], masked["acc"]).max() < 1e-6
            assert rel_err(out["pot"], masked["pot"]).max() < 1e-6
            # lasthalf without the neighbour output: the same launch (the pairs evaluated in FP64 are added with
            # atomics, so two runs agree to FP64 rounding of a sum, not bit for bit)
            assert rel_vec_err(out_nonn["acc"], out["acc"]).max() < 1e-12
            assert rel_err(out_nonn["pot"], out["pot"]).max() < 1e-12


def test_j_update_visible_and_last_write_wins(g6):
    O = _O()
    m, x, v = P.new_plummer_model(400, seed=3)
    ids = np.arange(400, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    g6.calc(ids[:8], x[:8], v[:8], 1e-4)
    x2 = x.copy(); v2 = v.copy()
    x2[100:110] += 0.01
    v2[100:110] *= 0.5
    z = np.zeros(3)
    for j in range(100, 110):
        g6.set_j_particle(j, int(ids[j]), 0.0, 0.125, m[j], z, z, z, v[j], x[j] + 5.0)   # overwritten below
        g6.set_j_particle(j, int(ids[j]), 0.0, 0.125, m[j], z, z, z, v2[j], x2[j])
    out = g6.calc(ids[:64], x[:64], v[:64], 1e-4)
    ref = O.force(x[:64], v[:64], m, x2, v2, 1e-4, iid=ids[:64], jid=ids)
    check_forces(out, ref, what="update")


def test_block_step_sequence_across_latency_path_thresholds(g6):
    """The call pattern of a block time step (gpu.cc:163-230,365-407; gravity.F:54-112, update_grape.F:19-40):
    ni j-updates, set_ti, firsthalf, lasthalf2 -- with i-block sizes on both sides of every switch of the
    latency path (i-block in the kernel parameters <= 64 / <= 384, results written straight to host memory
    <= 2048, scatter fused with the predictor for <= 256 updates) and a moving prediction time."""
    O = _O()
    n = 3000
    m, x, v = P.new_plummer_model(n, seed=11)
    ids = np.arange(1, n + 1, dtype=np.int32)
    rnd = np.random.RandomState(5)
    acc = 0.1 * rnd.standard_normal((n, 3)); jerk = 0.1 * rnd.standard_normal((n, 3))
    tj = np.zeros(n)
    _fresh(g6, ids, m, x, v, acc=acc, jerk=jerk, tj=tj)
    x = x.copy(); v = v.copy()
    t = 0.0
    for step, ni in enumerate([1, 4, 5, 32, 33, 64, 65, 255, 256, 257, 384, 385, 1000, 2048, 2049, 3, 3000, 40]):
        t += 2.0 ** -10
        sel = np.sort(rnd.choice(n, ni, replace=False))
        pp, pv = O.predict(t, tj, x, v, acc, jerk)
        g6.set_ti(t)
        out = g6.calc(ids[sel], pp[sel], pv[sel], 1e-4)
        ref = O.force(pp[sel], pv[sel], m, pp, pv, 1e-4, iid=ids[sel], jid=ids, scales=True)
        check_forces(out, ref, what="block step %d ni %d" % (step, ni))
        check_nn(out["nn"], ref["nn"], ids, pp[sel], pp)
        # the caller's corrector stand-in: advance the block to t and send it back (idata::update_gpu)
        x[sel] = pp[sel]; v[sel] = pv[sel]; tj[sel] = t
        acc[sel] = out["acc"]; jerk[sel] = out["jerk"]
        g6.set_j_particles(ids[sel], m[sel], x[sel], v[sel], acc=acc[sel], jerk=jerk[sel], tj=tj[sel],
                           address=sel.astype(np.int32))


@pytest.mark.parametrize("batched", [False, True], ids=["per-particle", "batched-replace"])
def test_bhtree_interaction_list_call_pattern(g6, batched):
    """The third g6 caller, src/amuse_bhtree/src/BHtree.C:763-885: every tree walk
    re-sends its whole interaction list (address = index = position in the list, velocities zero,
    t = 0), then asks for the forces on the leaf's particles -- which are members of the list, so the
    self pair is excluded by id -- with nj = list length (shrinking and growing from call to call),
    h2 = 0 and plain lasthalf.  batched: the list replaced by ONE g6x_set_j_particles call (address0 = 0) instead of
    `length` g6_set_j_particle_ calls -- the stub INTEGRATION.md shows for BHtree.C:800-812."""
    O = _O()
    rnd = np.random.RandomState(21)
    g6.nj = 0
    g6.set_variant(0)
    z = np.zeros(3)
    for length, ni in ((3000, 64), (700, 300), (5000, 17), (260, 260)):
        pos = rnd.standard_normal((length, 3))
        mass = rnd.uniform(0.5, 1.5, length) / length
        first_leaf = int(rnd.randint(0, length - ni + 1))
        g6.set_ti(0.0)
        if batched:
            g6.set_j_particles(np.arange(length, dtype=np.int32), mass, pos, np.zeros_like(pos))
        else:
            for i in range(length):
                g6.set_j_particle(i, i, 0.0, 0.0, mass[i], z, z, z, z, pos[i])
        idx = np.arange(first_leaf, first_leaf + ni, dtype=np.int32)
        vel0 = np.zeros((ni, 3))
        out = g6.calc(idx, pos[idx], vel0, 1e-4, nj=length, want_nn=False)
        ref = O.force(pos[idx], vel0, mass, pos, np.zeros_like(pos), 1e-4, iid=idx, jid=np.arange(length, dtype=np.int32))
        assert rel_vec_err(out["acc"], ref["acc"]).max() <= TOL, (length, ni)
        assert rel_err(out["pot"], ref["pot"]).max() <= TOL, (length, ni)
        assert np.abs(out["jerk"]).max() == 0.0          # all velocities are zero


def test_big_update_batches_flush_early_and_last_write_still_wins(g6):
    """Batches of >= 8192 staged updates go out while the caller is still staging (stage_j); a particle written
    before and after such a flush must end up with its last value (sapporo.cpp:83-110 semantics across batches)."""
    O = _O()
    n = 20000
    m, x, v = P.new_plummer_model(n, seed=14)
    ids = np.arange(1, n + 1, dtype=np.int32)
    _fresh(g6, ids, m, x + 3.0, v)                       # wrong positions everywhere ...
    z = np.zeros(3)
    for j in (0, 5000, 8191, 8192, 19999):               # ... and again, one by one, around the flush boundary
        g6.set_j_particle(j, int(ids[j]), 0.0, 0.125, m[j], z, z, z, v[j], x[j] - 1.0)
    g6.set_j_particles(ids, m, x, v)                     # the right values: 20000 records = two early flushes + a rest
    out = g6.calc(ids[:300], x[:300], v[:300], 1e-4)
    ref = O.force(x[:300], v[:300], m, x, v, 1e-4, iid=ids[:300], jid=ids, scales=True)
    check_forces(out, ref, what="early flush")
    check_nn(out["nn"], ref["nn"], ids, x[:300], x)


def test_empty_calls_and_extreme_ids(g6):
    """Degenerate calls the callers can make: a force call before any j-particle exists (BHTree opens the
    device long before its first list), an empty i-block, ids far beyond 2^24 (sapporo_light returns the
    neighbour id through a float, sapporo.cpp:226-227; ph4's binaries get ids of 1e7 and more,
    test_multiples2.py:254) and negative ids, a softening far larger than the system."""
    O = _O()
    g6.close()
    assert g6.L.g6_open_(C.byref(g6.cid)) == 0                 # fresh device state: no j-memory at all
    g6.nj = 0
    g6.set_ti(0.0)
    x3 = np.array([[0.1, 0.2, 0.3], [1.0, -1.0, 0.5], [0.0, 0.0, 0.0]]); v3 = np.zeros((3, 3))
    out = g6.calc(np.array([5, 6, 7], dtype=np.int32), x3, v3, 1e-4, nj=0)
    assert np.all(out["acc"] == 0) and np.all(out["jerk"] == 0) and np.all(out["pot"] == 0) and np.all(out["nn"] == -1)
    n = 1500
    m, x, v = P.new_plummer_model(n, seed=17)
    ids = ((1 << 30) + 7 * np.arange(n)).astype(np.int32)
    ids[::3] = -2 - np.arange(len(ids[::3]), dtype=np.int32)    # a third of them negative (not -1: field points)
    _fresh(g6, ids, m, x, v)
    out = g6.calc(ids[:0], x[:0], v[:0], 1e-4)                  # empty i-block: nothing to do, nothing to wait for
    assert out["acc"].shape == (0, 3)
    for eps2 in (0.0, 1e4):
        out = g6.calc(ids[:200], x[:200], v[:200], eps2)
        ref = O.force(x[:200], v[:200], m, x, v, eps2, iid=ids[:200], jid=ids, scales=True)
        check_forces(out, ref, what="extreme ids eps2=%g" % eps2)
        assert np.array_equal(out["nn"], ids[ref["nn"]])


def test_debug_readback_of_a_j_particle(g6):
    """get_j_part_data (grape.h:118-119): stored state and last prediction of one j-particle."""
    O = _O()
    n = 300
    m, x, v = P.new_plummer_model(n, seed=19)
    ids = np.arange(1, n + 1, dtype=np.int32)
    rnd = np.random.RandomState(2)
    acc = rnd.standard_normal((n, 3)); jerk = rnd.standard_normal((n, 3)); tj = -2.0 ** -rnd.randint(3, 8, n)
    _fresh(g6, ids, m, x, v, acc=acc, jerk=jerk, tj=tj, ti=0.0)
    g6.predict(0.0)
    pp, pv = O.predict(0.0, tj, x, v, acc, jerk)
    out = [np.zeros(3) for _ in range(6)]
    for j in (0, 17, n - 1):
        assert g6.L.get_j_part_data(j, n, *out) == 0
        assert np.array_equal(out[0], x[j]) and np.array_equal(out[1], v[j])
        assert np.allclose(out[2], acc[j], rtol=1e-15) and np.allclose(out[3], jerk[j], rtol=1e-15)
        assert np.allclose(out[4], pp[j], rtol=2e-14, atol=1e-15) and np.allclose(out[5], pv[j], rtol=2e-7, atol=1e-9)
    assert g6.L.get_j_part_data(n, n, *out) == -1


def test_neighbour_lists(g6):
    O = _O()
    m, x, v = P.new_plummer_model(3000, seed=8)
    ids = np.arange(1, 3001, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    f = O.force(x[:300], v[:300], m, x, v, 0.0, iid=ids[:300], jid=ids)
    h2 = np.minimum(8 * f["dnn"] ** 2, 1.0)          # gpu.cc:629-630
    out = g6.calc(ids[:300], x[:300], v[:300], 0.0, h2=h2)
    assert g6.read_neighbour_list() == 0
    bad = []
    for i in range(300):
        n_ref, lst_ref = O.neighbours(int(ids[i]), x[i], h2[i], ids, m, x)
        rc, n, lst = g6.get_neighbour_list(i)
        r2 = ((x - x[i]) ** 2).sum(axis=1)
        edge = set(ids[np.abs(r2 - h2[i]) <= 1e-6 * h2[i]].tolist())     # members at the FP32 edge of the sphere
        good = (set(lst.tolist()) ^ set(lst_ref.tolist()) <= edge) and rc == 0 and n == len(lst) \
            and np.all(np.diff(lst) > 0) and ((out["nn"][i] in lst) or f["dnn"][i] ** 2 > h2[i] * (1 - 1e-6))
        if not good:
            bad.append((i, rc, n, n_ref, int(out["nn"][i]), int(ids[f["nn"][i]]), lst[:6].tolist(), lst_ref[:6].tolist()))
    assert not bad, bad[:10]
    # overflow flag: nblen >= maxlength (lib/sapporo_light/sapporo.cpp:262-265; ph4 shrinks h2 on it, gpu.cc:668-751)
    k = int(np.argmax([g6.get_neighbour_list(i)[1] for i in range(300)]))
    rc, n, lst = g6.get_neighbour_list(k)
    assert n >= 2
    for maxlength, want in ((n + 1, 0), (n, 1), (n - 1, 1), (1, 1)):
        rc, n2, lst2 = g6.get_neighbour_list(k, maxlength=maxlength)
        assert n2 == n and (rc != 0) == bool(want), (maxlength, rc, n2)
        assert np.array_equal(lst2, lst[:min(n, maxlength)])
    # h2 = 0 everywhere -> no lists
    g6.calc(ids[:10], x[:10], v[:10], 0.0)
    assert g6.read_neighbour_list() == 0
    assert g6.get_neighbour_list(0)[1] == 0


def test_config1_n16k_through_abi_vs_cpu_forces(g6):
    """BASELINE configs[1]: Plummer N=16k through the g6 ABI on one B200, checked against CPU forces."""
    O = _O()
    n = 16384
    m, x, v = P.new_plummer_model(n, seed=1)
    ids = np.arange(1, n + 1, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    for eps2 in (0.0, 1e-4):
        out = g6.calc(ids, x, v, eps2)
        ref = O.force(x, v, m, x, v, eps2, scales=True)
        print("N=16k eps2=%g: %s" % (eps2, error_report(out, ref)))
        check_forces(out, ref, what="16k eps2=%g" % eps2)
        nties = check_nn(out["nn"], ref["nn"], ids, x, x)
        print("N=16k eps2=%g: nn tie mismatches %d of %d" % (eps2, nties, n))


def test_binaries_n8k(g6):
    """Hard primordial binaries (config 4 family): one neighbour dominates acc and jerk."""
    O = _O()
    m, x, v = P.new_plummer_model(8000, seed=2)
    ids, m, x, v = P.add_binaries(m, x, v, fraction=0.1)
    _fresh(g6, ids, m, x, v)
    out = g6.calc(ids, x, v, 0.0)
    ref = O.force(x, v, m, x, v, 0.0, scales=True)
    print("binaries N=8.8k: %s" % error_report(out, ref))
    check_forces(out, ref, what="binaries")
    check_nn(out["nn"], ref["nn"], ids, x, x)


def test_close_open_cycle(g6):
    """phiGRAPE closes and reopens the device around every evolve (interface.F:552-594)."""
    O = _O()
    m, x, v = P.new_plummer_model(500, seed=6)
    ids = np.arange(500, dtype=np.int32)
    ref = O.force(x[:32], v[:32], m, x, v, 1e-4, iid=ids[:32], jid=ids)
    for _ in range(3):
        g6.close()
        assert g6.L.g6_open_(C.byref(g6.cid)) == 0
        _fresh(g6, ids, m, x, v)
        out = g6.calc(ids[:32], x[:32], v[:32], 1e-4)
        check_forces(out, ref, what="reopen")
    # a device id that does not exist: folded onto the devices present by default (standalone ph4 passes an
    # uninitialised gpu_id, jdata.h:137-170), the reference's -1 (sapporo.cpp:31-34) with G6_B200_STRICT_DEVICE=1
    os.environ["G6_B200_STRICT_DEVICE"] = "1"
    try:
        assert g6.L.g6_open_(C.byref(C.c_int(4096))) == -1
    finally:
        del os.environ["G6_B200_STRICT_DEVICE"]
    assert g6.L.g6_open_(C.byref(C.c_int(4096))) == 0 and g6.L.g6x_device_count_open() == 1
    g6.close()
    assert g6.L.g6_open_(C.byref(g6.cid)) == 0


def test_full_size_n1m_sampled_oracle_and_properties(g6):
    """BASELINE's headline size.  The CPU oracle evaluates a random i-sample against all 1M j;
    the whole sweep is checked through a size-independent property: Newton's third law,
    sum_i m_i a_i = 0 and sum_i m_i jerk_i = 0 (relative to sum_i m_i |a_i|)."""
    O = _O()
    n = 1 << 20
    m, x, v = P.new_plummer_model(n, seed=1)
    ids = np.arange(1, n + 1, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    out = g6.calc(ids, x, v, 0.0)
    rnd = np.random.RandomState(0)
    samp = np.sort(rnd.choice(n, 256, replace=False))
    ref = O.force(x[samp], v[samp], m, x, v, 0.0, iid=ids[samp], jid=ids, scales=True)
    got = {k: out[k][samp] for k in out}
    print("N=1M sample of 256: %s" % error_report(got, ref))
    ea, ej, ep = check_forces(got, ref, what="N=1M sample")
    check_nn(got["nn"], ref["nn"], ids, x[samp], x)
    fa = np.abs((m[:, None] * out["acc"]).sum(axis=0)).max() / (m * np.linalg.norm(out["acc"], axis=1)).sum()
    fj = np.abs((m[:, None] * out["jerk"]).sum(axis=0)).max() / (m * np.linalg.norm(out["jerk"], axis=1)).sum()
    u = 0.5 * (m * out["pot"]).sum()
    print("N=1M: sample max rel err acc %.2e jerk %.2e pot %.2e; momentum residual acc %.2e jerk %.2e; U=%.6f"
          % (ea, ej, ep, fa, fj, u))
    assert fa < 1e-7 and fj < 1e-6
    assert -0.6 < u < -0.4                                    # Plummer in N-body units: U ~ -0.5
    # nearest neighbours are mutual for the closest pair of the whole system
    assert np.all(out["nn"] >= 1)


def test_device_resident_entry_point_matches_abi(g6):
    import torch
    O = _O()
    n = 5000
    m, x, v = P.new_plummer_model(n, seed=12)
    ids = np.arange(1, n + 1, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    out = g6.calc(ids, x, v, 1e-4)
    dev = torch.device("cuda:0")
    L = g6.L
    L.g6x_set_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), 1)
    try:
        d_id = torch.from_numpy(ids).to(dev)
        d_x = torch.from_numpy(x).to(dev)
        d_v = torch.from_numpy(v).to(dev)
        d_sum = torch.empty((n, 7), dtype=torch.float64, device=dev)
        d_key = torch.empty(n, dtype=torch.int64, device=dev)
        d_nn = torch.empty(n, dtype=torch.int32, device=dev)
        L.g6x_calc_device(n, n, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, 1e-4, 1,
                          d_sum.data_ptr(), d_key.data_ptr(), d_nn.data_ptr())
        torch.cuda.synchronize()
        s = d_sum.cpu().numpy()
        # same kernels, same j; the second call finds every particle's nearest-neighbour distance refreshed by the
        # first, so a few more or fewer pairs take the FP64 path: equal to FP32 pair rounding, not bit for bit
        check_forces(dict(acc=s[:, 0:3], jerk=s[:, 3:6], pot=-s[:, 6]), out, tol=3e-7, what="device entry vs ABI")
        assert np.array_equal(d_nn.cpu().numpy(), out["nn"])
        # key = (float bits of r2min) << 32 | address
        key = d_key.cpu().numpy().astype(np.uint64)
        addr = (key & np.uint64(0xffffffff)).astype(np.int64)
        assert np.array_equal(ids[addr], out["nn"])
    finally:
        L.g6x_set_stream(None, 0)


def test_j_shards_combine_like_ph4_domains(g6):
    """j-domain decomposition (jdata.cc:56-67 + idata.cc:284-313) on one device: evaluate each
    shard separately with its global address offset, combine sums and min-keys on the host."""
    import torch
    O = _O()
    n, ni, nsh = 6000, 700, 4
    m, x, v = P.new_plummer_model(n, seed=13)
    ids = np.arange(1, n + 1, dtype=np.int32)
    ref = O.force(x[:ni], v[:ni], m, x, v, 0.0, iid=ids[:ni], jid=ids)
    dev = torch.device("cuda:0")
    L = g6.L
    d_id = torch.from_numpy(ids[:ni].copy()).to(dev)
    d_x = torch.from_numpy(x[:ni].copy()).to(dev)
    d_v = torch.from_numpy(v[:ni].copy()).to(dev)
    tot = np.zeros((ni, 7))
    keys = np.full(ni, np.iinfo(np.int64).max, dtype=np.int64)
    nnid = np.zeros(ni, dtype=np.int64)
    shards = []
    for r in range(nsh):
        a, b = O.define_domain(n, nsh, r)
        shards.append((a, b))
    for r, (a, b) in enumerate(shards):
        g6.close(); L.g6_open_(C.byref(g6.cid)); g6.nj = 0
        L.g6x_set_j_offset(a)
        g6.set_j_particles(ids[a:b], m[a:b], x[a:b], v[a:b])
        g6.set_ti(0.0)
        d_sum = torch.empty((ni, 7), dtype=torch.float64, device=dev)
        d_key = torch.empty(ni, dtype=torch.int64, device=dev)
        d_nn = torch.empty(ni, dtype=torch.int32, device=dev)
        L.g6x_calc_device(b - a, ni, d_id.data_ptr(), d_x.data_ptr(), d_v.data_ptr(), None, 0.0, 1,
                          d_sum.data_ptr(), d_key.data_ptr(), d_nn.data_ptr())
        L.g6x_synchronize()
        tot += d_sum.cpu().numpy()
        keys = np.minimum(keys, d_key.cpu().numpy())
    # second pass: resolve the ids of the winners (idata.cc:308-313: only the owner contributes)
    d_keys = torch.from_numpy(keys).to(dev)
    for r, (a, b) in enumerate(shards):
        g6.close(); L.g6_open_(C.byref(g6.cid)); g6.nj = 0
        L.g6x_set_j_offset(a)
        g6.set_j_particles(ids[a:b], m[a:b], x[a:b], v[a:b])
        g6.predict(0.0)
        d_nn = torch.empty(ni, dtype=torch.int32, device=dev)
        L.g6x_resolve_nn(ni, d_keys.data_ptr(), r, d_nn.data_ptr())
        L.g6x_synchronize()
        nnid += d_nn.cpu().numpy()
    L.g6x_set_j_offset(0)
    got = dict(acc=tot[:, 0:3], jerk=tot[:, 3:6], pot=-tot[:, 6], nn=nnid.astype(np.int32))
    check_forces(got, ref, what="shards")
    check_nn(got["nn"], ref["nn"], ids, x[:ni], x)


def test_launch_counter_and_no_cpu_path(g6):
    m, x, v = P.new_plummer_model(300, seed=1)
    ids = np.arange(300, dtype=np.int32)
    _fresh(g6, ids, m, x, v)
    c0 = g6.launch_count()
    g6.calc(ids, x, v, 1e-4)
    assert g6.launch_count() - c0 >= 3      # scatter + predictor + force
