"""numpy float32 emulation of the force kernels' arithmetic (double-single geometry, FFMA chains,
MUFU.RSQ + Newton variants, FP32 group sums flushed to FP64) against an FP64 evaluation of the same
pairs, to attribute the per-particle error to its sources without GPU time.
MUFU.RSQ is modelled as the correctly rounded value times (1 + U(-1,1) * 2^-22.9).
Usage: python tools/emulate_kernel.py [case]   (case: ragged | binaries | plummer1k)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amuse_b200 import plummer as P  # noqa: E402

f = np.float32
F = np.float64


def fma(a, b, c):
    return (a.astype(F) * b.astype(F) + c.astype(F)).astype(f)


def split(x):
    h = x.astype(f)
    return h, (x - h.astype(F)).astype(f)


def emulate(ipos, ivel, iid, m, x, v, jid, eps2, newton="B", grp=32, dual=False, mufu_err=2 ** -22.9, seed=0,
            exact=()):
    """exact: subset of {'geom','rsqrt','mr3','acc_sum'} evaluated in FP64 instead (attribution)."""
    rnd = np.random.RandomState(seed)
    ni, nj = len(ipos), len(m)
    xih, xil = split(ipos)
    xjh, xjl = split(x)
    vi = ivel.astype(f)
    vj = v.astype(f)
    mj = m.astype(f)
    e2 = f(f(eps2) + f(2.220446049250313e-16))
    D = np.zeros((ni, 7))
    S = np.zeros((2, ni, 7), dtype=f)
    for j in range(nj):
        if 'geom' in exact:
            d = [(x[j, k] - ipos[:, k]) for k in range(3)]
            d = [t.astype(f) for t in d]
        else:
            d = [((xjh[j, k] - xih[:, k]).astype(f) + (xjl[j, k] - xil[:, k]).astype(f)).astype(f) for k in range(3)]
        dv = [(vj[j, k] - vi[:, k]).astype(f) for k in range(3)]
        r2 = fma(d[2], d[2], fma(d[1], d[1], (d[0] * d[0]).astype(f)))
        xv = fma(d[2], dv[2], fma(d[1], dv[1], (d[0] * dv[0]).astype(f)))
        r2e = (r2 + e2).astype(f)
        ok = (jid[j] != iid) & (r2 > f(2.220446049250313e-16))
        idok = (jid[j] != iid)
        if 'rsqrt' in exact:
            rinv = (1 / np.sqrt(r2e.astype(F)))
            rinv2 = (rinv * rinv).astype(f)
            mrinv = (mj[j] * rinv).astype(f)
        else:
            y0 = ((1 / np.sqrt(r2e.astype(F))) * (1 + mufu_err * rnd.uniform(-1, 1, ni))).astype(f)
            if newton == "A":     # e = x*y0^2 - 1 ; rinv = y0 - (y0/2)*e ; rinv2 = rinv^2
                yy = (y0 * y0).astype(f)
                e = fma(r2e, yy, np.full(ni, -1, f))
                rinv = fma((y0 * f(-0.5)).astype(f), e, y0)
                rinv2 = (rinv * rinv).astype(f)
                mrinv = (mj[j] * rinv).astype(f)
            elif newton == "B":   # e2 = x*yy - 2 ; -rinv2 = yy*e2 ; rinv = y0*(0.5 - e2/2)
                yy = (y0 * y0).astype(f)
                ee = fma(r2e, yy, np.full(ni, -2, f))
                rinv2 = (-(yy * ee)).astype(f)
                c = fma(ee, np.full(ni, -0.5, f), np.full(ni, 0.5, f))
                mrinv = ((mj[j] * y0).astype(f) * c).astype(f)
            else:                 # raw MUFU
                rinv2 = (y0 * y0).astype(f)
                mrinv = (mj[j] * y0).astype(f)
        if 'mr3' in exact:
            mr3 = (mrinv.astype(F) * rinv2.astype(F))
            a3 = (-3.0 * xv.astype(F) * rinv2.astype(F))
            terms = [mr3 * d[k] for k in range(3)] + [mr3 * (a3 * d[k] + dv[k]) for k in range(3)] + [mrinv.astype(F)]
            terms = [np.where(idok, t, 0.0) for t in terms[:6]] + [np.where(ok, terms[6], 0.0)]
            D += np.stack(terms, axis=1)
            continue
        mr3 = np.where(idok, (mrinv * rinv2).astype(f), f(0))
        a3 = ((xv * rinv2).astype(f) * f(-3)).astype(f)
        s = S[j & 1] if dual else S[0]
        if 'acc_sum' in exact:
            terms = [mr3.astype(F) * d[k] for k in range(3)] + \
                    [mr3.astype(F) * fma(a3, d[k], dv[k]) for k in range(3)] + [np.where(ok, mrinv, 0).astype(F)]
            D += np.stack(terms, axis=1)
        else:
            for k in range(3):
                s[:, k] = fma(mr3, d[k], s[:, k])
                s[:, 3 + k] = fma(mr3, fma(a3, d[k], dv[k]), s[:, 3 + k])
            s[:, 6] = (s[:, 6] + np.where(ok, mrinv, f(0))).astype(f)
            if (j + 1) % grp == 0 or j == nj - 1:
                tot = (S[0] + S[1]).astype(f) if dual else S[0]
                D += tot.astype(F)
                S[:] = 0
    return dict(acc=D[:, 0:3], jerk=D[:, 3:6], pot=-D[:, 6])


def reference(ipos, ivel, iid, m, x, v, jid, eps2):
    X = x[None, :, :] - ipos[:, None, :]
    V = v[None, :, :] - ivel[:, None, :]
    R2 = (X * X).sum(2)
    XV = (X * V).sum(2)
    idok = jid[None, :] != iid[:, None]
    ok = idok & (R2 > 2.220446049250313e-16)
    r2i = 1 / (R2 + eps2 + 2.220446049250313e-16)
    ri = np.sqrt(r2i)
    mri = m[None, :] * ri
    mr3 = np.where(idok, mri * r2i, 0)
    a3 = -3 * XV * r2i
    acc = (mr3[:, :, None] * X).sum(1)
    jerk = (mr3[:, :, None] * (V + a3[:, :, None] * X)).sum(1)
    pot = -np.where(ok, mri, 0).sum(1)
    sacc = np.linalg.norm(mr3[:, :, None] * X, axis=2).sum(1)
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
        m, x, v, ids = P.new_plummer_with_binaries(2048, 0.1, seed=3) if hasattr(P, "new_plummer_with_binaries") else (None,) * 4
        return x, v, ids, m, x, v, ids, 0.0
    g = dict(np.load(os.path.join(ROOT, "tests/golden/ph4_plummer1k_eps1e-4.npz")))
    return g["pos"], g["vel"], g["ids"], g["mass"], g["pos"], g["vel"], g["ids"], float(g["eps2"])


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "ragged"
    args = case(name)
    ref = reference(*args)
    kappa = ref["sacc"] / np.linalg.norm(ref["acc"], axis=1)
    print("case %s: ni %d nj %d; cancellation sum|a_ij|/|a_i|: median %.1f max %.1f" % (
        name, len(args[0]), len(args[3]), np.median(kappa), kappa.max()))
    for label, kw in [("newton A grp32", dict(newton="A")), ("newton B grp32 (the kernel)", dict(newton="B")),
                      ("newton B grp8", dict(newton="B", grp=8)), ("newton B grp16", dict(newton="B", grp=16)),
                      ("newton B grp32 dual", dict(newton="B", grp=32, dual=True)),
                      ("newton B grp64 dual", dict(newton="B", grp=64, dual=True)),
                      ("raw mufu grp32", dict(newton="none")),
                      ("B, exact geom", dict(exact=("geom",))), ("B, exact rsqrt", dict(exact=("rsqrt",))),
                      ("B, exact mr3..", dict(exact=("mr3",))), ("B, exact sums", dict(exact=("acc_sum",))),
                      ("B, exact rsqrt+sums", dict(exact=("rsqrt", "acc_sum")))]:
        worst = [0, 0, 0]
        for seed in range(3):
            ea, ej, ep = errs(emulate(*args, seed=seed, **kw), ref)
            worst = [max(worst[0], ea.max()), max(worst[1], ej.max()), max(worst[2], ep.max())]
            i = int(np.argmax(ea))
        print("%-22s acc max %.2e (i=%d, kappa %.0f) p99 %.2e | jerk max %.2e p99 %.2e | pot max %.2e" % (
            label, worst[0], i, kappa[i], np.percentile(ea, 99), worst[1], np.percentile(ej, 99), worst[2]))
