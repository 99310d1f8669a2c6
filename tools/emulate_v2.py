"""numpy emulation of the round-2 force-kernel arithmetic, used to choose its two accuracy parameters
without GPU time (the FP64 reference is the same pair formula the oracle uses, idata.cc:206-233):

  * j-memory and i-blocks in Morton order; a (warp of 64 i) x (group of 32 j) block whose bounding
    boxes are further apart than FARC * (largest |coordinate| of either box) AND further than every
    i's near radius is "far": position differences use the hi parts only (31 instead of 37 FP32
    operations per pair);
  * pairs closer than the i-particle's near radius r_hp,i = sqrt(K) * d_est,i (d_est = an upper bound
    of the nearest-neighbour distance, here from the +-W Morton neighbours) are evaluated in FP64
    from the full double-single positions and velocities and added to the FP64 totals directly;
  * everything else: the FP32 pair function of the kernel (packed arithmetic is IEEE per lane), MUFU.RSQ
    modelled as the correctly rounded value times (1 + U(-1,1) 2^-22.9) + the sign-folded Newton step,
    FP32 sums over one group, flushed to FP64.

Usage: python tools/emulate_v2.py case [K list] [FARC]      case: plummer1k | plummer1k0 | binaries | ragged | pN
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import plummer as P  # noqa: E402

f = np.float32
F = np.float64
TINY = 2.220446049250313e-16


def fma(a, b, c):
    return (a.astype(F) * b.astype(F) + c.astype(F)).astype(f)


def split(x):
    h = x.astype(f)
    return h, (x - h.astype(F)).astype(f)


def morton_order(x, lo, hi, bits=10):
    q = np.clip(((x - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1).astype(np.uint64)
    key = np.zeros(len(x), dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return np.argsort(key, kind="stable"), key


def emulate(ipos, ivel, iid, m, x, v, jid, eps2, K=16.0, farc=0.125, W=32, grp=32, warp=64, mufu_err=2 ** -22.9,
            seed=0, stats=None, group_mode=True):
    rnd = np.random.RandomState(seed)
    ni, nj = len(ipos), len(m)
    lo, hi = x.min(axis=0), x.max(axis=0)
    x0 = 0.5 * (lo + hi)                                   # library-internal origin
    jo, jkey = morton_order(x, lo, hi)
    io, ikey = morton_order(ipos, lo, hi)
    xs, vs, ms, ids = x[jo] - x0, v[jo], m[jo], jid[jo]
    npad = (-nj) % grp
    if npad:
        xs = np.vstack((xs, np.full((npad, 3), 1e18)))
        vs = np.vstack((vs, np.zeros((npad, 3))))
        ms = np.concatenate((ms, np.zeros(npad)))
        ids = np.concatenate((ids, np.full(npad, -(2 ** 31), dtype=ids.dtype)))
    NJ = nj + npad
    G = NJ // grp
    xjh, xjl = split(xs)
    vjh, vjl = split(vs)
    mj = ms.astype(f)
    # per-group bounding boxes of the hi parts (massive members only)
    gx = np.where((ms > 0)[:, None], xjh, np.nan).reshape(G, grp, 3)
    with np.errstate(all="ignore"):
        gmin = np.nanmin(gx, axis=1)
        gmax = np.nanmax(gx, axis=1)
    gmin = np.where(np.isnan(gmin), 1e18, gmin)
    gmax = np.where(np.isnan(gmax), 1e18, gmax)
    e2 = f(f(eps2) + f(TINY))
    skey = jkey[jo]
    out = dict(acc=np.zeros((ni, 3)), jerk=np.zeros((ni, 3)), pot=np.zeros(ni))
    nhp = nfar = nblk = nredo = 0
    for w0 in range(0, ni, warp):
        sel = io[w0:w0 + warp]
        n = len(sel)
        xi, vi, idi = ipos[sel] - x0, ivel[sel], iid[sel]
        xih, xil = split(xi)
        vih, vil = split(vi)
        # near radius: +-W Morton neighbours of the slot where i's key falls
        pos = np.searchsorted(skey, ikey[sel])
        d2 = np.full(n, np.inf)
        for k in range(n):
            a, b = max(0, pos[k] - W), min(nj, pos[k] + W)
            dd = ((xs[a:b] - xi[k]) ** 2).sum(axis=1)
            dd = dd[(ids[a:b] != idi[k]) & (dd > TINY) & (ms[a:b] > 0)]
            if len(dd):
                d2[k] = dd.min()
        thr = (K * d2).astype(f)
        # warp box and far groups
        wmin, wmax = xih.min(axis=0), xih.max(axis=0)
        gap = np.maximum(0, np.maximum(gmin - wmax, wmin - gmax))
        gap2 = (gap * gap).sum(axis=1)
        scale = np.maximum(np.abs(gmin).max(axis=1), np.abs(gmax).max(axis=1))
        scale = np.maximum(scale, max(np.abs(wmin).max(), np.abs(wmax).max()))
        far = (gap2 > (farc * scale) ** 2) & (gap2 > np.nanmax(np.where(np.isfinite(thr), thr, 0))) & (scale < 1e17)
        if not np.isfinite(thr).all():
            far[:] = False
        farj = np.repeat(far, grp)
        nfar += far.sum()
        nblk += G
        # geometry [n, NJ]
        d = []
        for k in range(3):
            dh = (xjh[None, :, k] - xih[:, None, k])
            dl = (xjl[None, :, k] - xil[:, None, k])
            d.append(np.where(farj[None, :], dh, (dh + dl).astype(f)))
        dv = [(vjh[None, :, k] - vih[:, None, k]) for k in range(3)]
        r2 = fma(d[2], d[2], fma(d[1], d[1], (d[0] * d[0]).astype(f)))
        xv = fma(d[2], dv[2], fma(d[1], dv[1], (d[0] * dv[0]).astype(f)))
        r2e = (r2 + e2).astype(f)
        idok = ids[None, :] != idi[:, None]
        ok = idok & (r2 > f(TINY))
        if group_mode:   # the kernel's rule: the whole (warp x group) block goes to FP64 when the boxes are close
            # point-to-box distance of every i of the warp to every group box against its own close radius
            pg = np.maximum(0, np.maximum(gmin[None, :, :] - xih[:, None, :], xih[:, None, :] - gmax[None, :, :]))
            pg2 = (pg * pg).sum(axis=2)
            close = (pg2 <= np.where(np.isfinite(thr), thr, 1e30)[:, None]).any(axis=0)
            far &= ~close
            farj = np.repeat(far, grp)
            closej = np.repeat(close, grp)
            hp = idok & closej[None, :] & (mj[None, :] > 0)
        else:
            hp = ok & (r2 < thr[:, None]) & ~farj[None, :]
        nhp += hp.sum()
        nredo += (hp.reshape(n, G, grp).any(axis=(0, 2))).sum()
        with np.errstate(all="ignore"):
            y0 = ((1 / np.sqrt(r2e.astype(F))) * (1 + mufu_err * rnd.uniform(-1, 1, r2e.shape))).astype(f)
            yy = (y0 * y0).astype(f)
            ee = fma(r2e, yy, np.full(1, -2, f))
            rinv2 = (-(yy * ee)).astype(f)
            c = fma(ee, np.full(1, -0.5, f), np.full(1, 0.5, f))
            mrinv = ((mj[None, :] * y0).astype(f) * c).astype(f)
            use = idok & ~hp
            mr3 = np.where(use, (mrinv * rinv2).astype(f), f(0))
            a3 = np.where(use, ((xv * rinv2).astype(f) * f(-3)).astype(f), f(0))
            mpot = np.where(ok & ~hp, mrinv, f(0))
        S = np.zeros((n, G, 7), dtype=f)
        R = lambda t: t.reshape(n, G, grp)
        mr3g, a3g, mpg = R(mr3), R(a3), R(mpot)
        dg = [R(t) for t in d]
        dvg = [R(t.astype(f)) for t in dv]
        for u in range(grp):
            for k in range(3):
                S[:, :, k] = fma(mr3g[:, :, u], dg[k][:, :, u], S[:, :, k])
                S[:, :, 3 + k] = fma(mr3g[:, :, u], fma(a3g[:, :, u], dg[k][:, :, u], dvg[k][:, :, u]), S[:, :, 3 + k])
            S[:, :, 6] = (S[:, :, 6] + mpg[:, :, u]).astype(f)
        D = S.astype(F).sum(axis=1)
        # near pairs in FP64 from the double-single operands
        ii, jj = np.nonzero(hp)
        if len(ii):
            X = (xjh[jj].astype(F) - xih[ii].astype(F)) + (xjl[jj].astype(F) - xil[ii].astype(F))
            V = (vjh[jj].astype(F) + vjl[jj].astype(F)) - (vih[ii].astype(F) + vil[ii].astype(F))
            r2d = (X * X).sum(axis=1)
            xvd = (X * V).sum(axis=1)
            r2i = 1 / (r2d + eps2 + TINY)
            ri = np.sqrt(r2i)
            mri = mj[jj].astype(F) * ri
            m3 = mri * r2i
            A3 = -3 * xvd * r2i
            np.add.at(D, (ii, slice(0, 3)), m3[:, None] * X)
            np.add.at(D, (ii, slice(3, 6)), m3[:, None] * (V + A3[:, None] * X))
            np.add.at(D, (ii, 6), np.where(r2d > TINY, mri, 0.0))
        out["acc"][sel] = D[:, 0:3]
        out["jerk"][sel] = D[:, 3:6]
        out["pot"][sel] = -D[:, 6]
    if stats is not None:
        stats.update(hp_per_i=nhp / ni, far_frac=nfar / max(nblk, 1), redo_frac=nredo / max(nblk, 1))
    return out


def reference(ipos, ivel, iid, m, x, v, jid, eps2, chunk=512):
    ni = len(ipos)
    acc = np.zeros((ni, 3)); jerk = np.zeros((ni, 3)); pot = np.zeros(ni); sacc = np.zeros(ni)
    mf = m.astype(f).astype(F)   # the library carries masses in FP32; parity tests use FP32-exact masses
    for i0 in range(0, ni, chunk):
        s = slice(i0, i0 + chunk)
        X = x[None, :, :] - ipos[s, None, :]
        V = v[None, :, :] - ivel[s, None, :]
        R2 = (X * X).sum(2)
        XV = (X * V).sum(2)
        idok = (jid[None, :] != iid[s, None]) & (m[None, :] > TINY)
        ok = idok & (R2 > TINY)
        r2i = 1 / (R2 + eps2 + TINY)
        ri = np.sqrt(r2i)
        mri = mf[None, :] * ri
        mr3 = np.where(idok, mri * r2i, 0)
        a3 = -3 * XV * r2i
        acc[s] = (mr3[:, :, None] * X).sum(1)
        jerk[s] = (mr3[:, :, None] * (V + a3[:, :, None] * X)).sum(1)
        pot[s] = -np.where(ok, mri, 0).sum(1)
        sacc[s] = np.linalg.norm(mr3[:, :, None] * X, axis=2).sum(1)
    return dict(acc=acc, jerk=jerk, pot=pot, sacc=sacc)


def errs(out, ref):
    ea = np.linalg.norm(out["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ej = np.linalg.norm(out["jerk"] - ref["jerk"], axis=1) / np.linalg.norm(ref["jerk"], axis=1)
    ep = np.abs(out["pot"] - ref["pot"]) / np.abs(ref["pot"])
    return ea, ej, ep


def case(name):
    if name == "ragged":      # tests/test_gpu_parity.py::test_ragged_sizes[2100-700]
        ni, nj = 2100, 700
        m, x, v = P.new_plummer_model(max(nj, 2), seed=20 + nj % 7)
        ids = np.arange(100, 100 + nj, dtype=np.int32)
        rnd = np.random.RandomState(ni)
        ipos = rnd.normal(size=(ni, 3)) * 0.7
        ivel = rnd.normal(size=(ni, 3)) * 0.5
        iid = -np.ones(ni, dtype=np.int32)
        k = min(ni, nj) // 2
        ipos[:k], ivel[:k], iid[:k] = x[:k], v[:k], ids[:k]
        return ipos, ivel, iid, m, x, v, ids, 1e-4
    if name == "binaries":
        m, x, v = P.new_plummer_model(8000, seed=2)
        ids, m, x, v = P.add_binaries(m, x, v, 0.1, seed=7)
        return x, v, ids, m, x, v, ids, 0.0
    if name[0] == "p" and name[1:].isdigit():
        n = int(name[1:])
        m, x, v = P.new_plummer_model(n, seed=1)
        ids = np.arange(1, n + 1, dtype=np.int32)
        return x, v, ids, m, x, v, ids, 0.0
    g = dict(np.load(os.path.join(ROOT, "tests/golden/ph4_plummer1k_%s.npz" % ("eps0" if name.endswith("0") else "eps1e-4"))))
    return g["pos"], g["vel"], g["ids"], g["mass"], g["pos"], g["vel"], g["ids"], float(g["eps2"])


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "plummer1k"
    Ks = [float(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 16, 64]
    farc = float(sys.argv[3]) if len(sys.argv) > 3 else 0.125
    gm = (sys.argv[4] != "pair") if len(sys.argv) > 4 else True
    args = case(name)
    ref = reference(*args)
    print("case %s: ni %d nj %d eps2 %g" % (name, len(args[0]), len(args[3]), args[7]))
    for K in Ks:
        st = {}
        ea, ej, ep = errs(emulate(*args, K=K, farc=farc, stats=st, group_mode=gm), ref)
        print("K=%-4g farc=%g: hp pairs/i %.1f, far blocks %.1f%%, redone blocks %.2f%% | acc max %.2e p99 %.2e | "
              "jerk max %.2e p99 %.2e | pot max %.2e" % (K, farc, st["hp_per_i"], 100 * st["far_frac"],
                                                         100 * st["redo_frac"], ea.max(), np.percentile(ea, 99),
                                                         ej.max(), np.percentile(ej, 99), ep.max()), flush=True)
