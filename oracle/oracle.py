"""ctypes bindings of the parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  It loads

* ``oracle/liboracle_g6.so``   -- the plain-C restatement (oracle_g6.c), and
* ``oracle/_ref/libph4ref.so`` -- the unmodified reference ph4 CPU core built by
  ``make -C oracle ref`` (when present; it is built in the dev container from
  /root/reference and travels to the GPU box as a binary),
* ``oracle/_ref/libg6ref.so``  -- the reference lib/g6lib CPU emulation.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the C restatement and, if /root/reference exists, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(os.environ.get("REF", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle_g6.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_predict.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.oracle_force.argtypes = [C.c_int, C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_void_p,
                                   _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _ip, _dp]
        L.oracle_force_scales.argtypes = [C.c_int, C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_void_p,
                                          _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp]
        L.oracle_combine.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, _dp]
        L.oracle_define_domain.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_neighbours.argtypes = [C.c_int, _dp, C.c_double, C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, _ip]
        L.oracle_neighbours.restype = C.c_int
        _lib = L
    return _lib


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def predict(t, time, pos, vel, acc, jerk):
    """jdata::predict_all restatement; returns (pred_pos, pred_vel)."""
    nj = len(time)
    pp = np.empty((nj, 3))
    pv = np.empty((nj, 3))
    lib().oracle_predict(nj, float(t), _c(time), _c(pos), _c(vel), _c(acc), _c(jerk), pp, pv)
    return pp, pv


def force(ipos, ivel, mass, pred_pos, pred_vel, eps2, iid=None, jid=None, j_start=0, j_end=None, scales=False):
    """idata::get_partial_acc_and_jerk restatement.

    Returns dict(acc, jerk, pot, nn (j index), dnn)."""
    ipos = _c(ipos)
    ivel = _c(ivel)
    ni = len(ipos)
    nj = len(mass)
    if j_end is None:
        j_end = nj
    acc = np.empty((ni, 3))
    jerk = np.empty((ni, 3))
    pot = np.empty(ni)
    nn = np.empty(ni, dtype=np.int32)
    dnn = np.empty(ni)
    use_ids = int(iid is not None and jid is not None)
    iid_ = _c(iid, np.int32) if use_ids else None
    jid_ = _c(jid, np.int32) if use_ids else None
    lib().oracle_force(ni, iid_.ctypes.data if use_ids else None, ipos, ivel, j_start, j_end,
                       jid_.ctypes.data if use_ids else None, _c(mass), _c(pred_pos), _c(pred_vel),
                       float(eps2), use_ids, acc, jerk, pot, nn, dnn)
    sacc = np.empty(ni)
    sjerk = np.empty(ni)
    if scales:
        lib().oracle_force_scales(ni, iid_.ctypes.data if use_ids else None, ipos, ivel, j_start, j_end,
                                  jid_.ctypes.data if use_ids else None, _c(mass), _c(pred_pos), _c(pred_vel),
                                  float(eps2), use_ids, sacc, sjerk)
        return dict(acc=acc, jerk=jerk, pot=pot, nn=nn, dnn=dnn, sacc=sacc, sjerk=sjerk)
    return dict(acc=acc, jerk=jerk, pot=pot, nn=nn, dnn=dnn)


def combine(parts):
    """idata::get_acc_and_jerk reduction tail over a list of force() results."""
    nd = len(parts)
    ni = len(parts[0]["pot"])

    def st(k, dt=np.float64):
        return _c(np.stack([p[k] for p in parts]), dt)

    acc = np.empty((ni, 3))
    jerk = np.empty((ni, 3))
    pot = np.empty(ni)
    nn = np.empty(ni, dtype=np.int32)
    dnn = np.empty(ni)
    lib().oracle_combine(nd, ni, st("acc"), st("jerk"), st("pot"), st("nn", np.int32), st("dnn"),
                         acc, jerk, pot, nn, dnn)
    return dict(acc=acc, jerk=jerk, pot=pot, nn=nn, dnn=dnn)


def define_domain(nj, size, rank):
    a = C.c_int()
    b = C.c_int()
    lib().oracle_define_domain(nj, size, rank, C.byref(a), C.byref(b))
    return a.value, b.value


def neighbours(iid, ipos, h2, jid, mass, pred_pos, maxlen=1 << 20):
    lst = np.empty(maxlen, dtype=np.int32)
    n = lib().oracle_neighbours(int(iid), _c(ipos), float(h2), 0, len(mass), _c(jid, np.int32),
                                _c(mass), _c(pred_pos), maxlen, lst)
    return n, lst[:min(n, maxlen)].copy()


# --------------------------------------------------------------------------
# The real reference (oracle/_ref), when its binaries are present.
# --------------------------------------------------------------------------
_ref = {}

_EVOLVE_ARGS = [C.c_int, C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                C.c_int, _dp, C.c_void_p, C.c_void_p]


def ref_available(libname="libph4ref.so"):
    return os.path.exists(os.path.join(HERE, "_ref", libname))


def ref(libname="libph4ref.so"):
    if libname not in _ref:
        L = C.CDLL(os.path.join(HERE, "_ref", libname))
        L.ph4ref_full_sweep.argtypes = [C.c_int, C.c_void_p, _dp, _dp, _dp, C.c_double,
                                        _dp, _dp, _dp, _ip, _dp, C.POINTER(C.c_double)]
        L.ph4ref_predict_force.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double,
                                           C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp,
                                           C.POINTER(C.c_double)]
        L.ph4ref_evolve.argtypes = _EVOLVE_ARGS
        _ref[libname] = L
    return _ref[libname]


def ref_evolve_enc(mass, pos, vel, eps2, eta, t_end, manage_encounters=1, ids=None, use_gpu=True,
                   libname="libph4ref_gpu.so", max_block_steps=0):
    """The reference integrator with its own close-encounter management (ph4ref_evolve_enc in ref_driver.cc)."""
    n = len(mass)
    L = ref(libname)
    out = np.zeros(8)
    ids_ = _c(ids, np.int32) if ids is not None else None
    L.ph4ref_evolve_enc.argtypes = [C.c_int, C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.c_int, C.c_long, _dp]
    L.ph4ref_evolve_enc(n, ids_.ctypes.data if ids is not None else None, _c(mass), _c(pos), _c(vel), float(eps2),
                        float(eta), float(t_end), int(use_gpu), int(manage_encounters), int(max_block_steps), out)
    return dict(E0=out[0], E1=out[1], block_steps=int(out[2]), particle_steps=int(out[3]), seconds=out[4], t=out[5],
                nj_left=int(out[6]))


def ref_full_sweep(mass, pos, vel, eps2, ids=None):
    """Reference ph4 idata::setup() sweep (i = j, t = 0).  Returns dict + seconds."""
    n = len(mass)
    acc = np.empty((n, 3))
    jerk = np.empty((n, 3))
    pot = np.empty(n)
    nn = np.empty(n, dtype=np.int32)
    dnn = np.empty(n)
    sec = C.c_double()
    ids_ = _c(ids, np.int32) if ids is not None else None
    ref().ph4ref_full_sweep(n, ids_.ctypes.data if ids is not None else None, _c(mass), _c(pos), _c(vel),
                            float(eps2), acc, jerk, pot, nn, dnn, C.byref(sec))
    return dict(acc=acc, jerk=jerk, pot=pot, nn=nn, dnn=dnn, seconds=sec.value)


def ref_predict_force(mass, tj, pos, vel, acc, jerk, t, eps2, ipos, ivel):
    nj = len(mass)
    ni = len(ipos)
    pp = np.empty((nj, 3))
    pv = np.empty((nj, 3))
    ia = np.empty((ni, 3))
    ij = np.empty((ni, 3))
    ip = np.empty(ni)
    inn = np.empty(ni, dtype=np.int32)
    idn = np.empty(ni)
    sec = C.c_double()
    ref().ph4ref_predict_force(nj, _c(mass), _c(tj), _c(pos), _c(vel), _c(acc), _c(jerk), float(t),
                               float(eps2), ni, _c(ipos), _c(ivel), pp, pv, ia, ij, ip, inn, idn,
                               C.byref(sec))
    return dict(pred_pos=pp, pred_vel=pv, acc=ia, jerk=ij, pot=ip, nn=inn, dnn=idn, seconds=sec.value)


def correct(tnext, eta, itime, itimestep, old_acc, old_jerk, iacc, ijerk, ipos, ivel, use_ref=False):
    """idata::correct restatement (or the reference itself with use_ref): returns
    (pos, vel, time, timestep) after the corrector."""
    ni = len(itime)
    t = _c(itime).copy(); dt = _c(itimestep).copy(); p = _c(ipos).copy(); v = _c(ivel).copy()
    L = ref() if use_ref else lib()
    f = L.ph4ref_correct if use_ref else L.oracle_correct
    f.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    f(ni, float(tnext), float(eta), t, dt, _c(old_acc), _c(old_jerk), _c(iacc), _c(ijerk), p, v)
    return p, v, t, dt


def check_encounters(iid, inn, idnn, imass, iradius, jid, jmass, jradius, rmin, jvel=None, use_ref=False):
    """idata::check_encounters restatement (or the reference with use_ref): returns (close1, close2, coll1, coll2)."""
    out = np.zeros(4, dtype=np.int32)
    ni, nj = len(iid), len(jid)
    if use_ref:
        L = ref()
        L.ph4ref_check_encounters.argtypes = [C.c_int, _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, _ip]
        jv = _c(jvel) if jvel is not None else np.zeros((nj, 3))
        L.ph4ref_check_encounters(nj, _c(jid, np.int32), _c(jmass), _c(jradius), jv, ni, _c(iid, np.int32),
                                  _c(inn, np.int32), _c(idnn), _c(imass), _c(iradius), float(rmin), out)
    else:
        L = lib()
        L.oracle_check_encounters.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _ip, _dp, _dp, C.c_double, _ip]
        L.oracle_check_encounters(ni, _c(iid, np.int32), _c(inn, np.int32), _c(idnn), _c(imass), _c(iradius),
                                  _c(jid, np.int32), _c(jmass), _c(jradius), float(rmin), out)
    return tuple(int(v) for v in out)


def initial_timestep(system_time, eta, acc, jerk):
    """jdata::set_initial_timestep restatement (jdata.cc:503-548)."""
    n = len(acc)
    out = np.zeros(n)
    L = lib()
    L.oracle_initial_timestep.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
    L.oracle_initial_timestep(n, float(system_time), float(eta), _c(acc), _c(jerk), out)
    return out


def ref_evolve(mass, pos, vel, eps2, eta, t_end, ids=None, use_gpu=False, libname="libph4ref.so", max_block_steps=0):
    """Run the reference Hermite integrator (CPU mode, or g6-ABI mode with
    libname='libph4ref_gpu.so' built by ``make -C oracle refgpu``)."""
    n = len(mass)
    L = ref(libname)
    out = np.zeros(8)
    ids_ = _c(ids, np.int32) if ids is not None else None
    L.ph4ref_evolve_steps.argtypes = [C.c_int, C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                      C.c_long, _dp, C.c_void_p, C.c_void_p]
    L.ph4ref_evolve_steps(n, ids_.ctypes.data if ids is not None else None, _c(mass), _c(pos), _c(vel),
                          float(eps2), float(eta), float(t_end), int(use_gpu), int(max_block_steps), out, None, None)
    return dict(E0=out[0], E1=out[1], block_steps=int(out[2]), particle_steps=int(out[3]),
                seconds=out[4], t=out[5])
