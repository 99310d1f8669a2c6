"""ctypes binding of the B200 g6 library (``amuse_b200/csrc/libsapporo.so``).

This is the call surface a g6 client sees: the same fourteen Fortran-convention
symbols ph4 binds in ``src/amuse_ph4/src/grape.h:6-120`` (every argument a
pointer), plus the ``g6x_`` batched/device-resident extensions declared in
``include/g6_b200.h``.  The class below only marshals numpy arrays to those C
calls; it contains no arithmetic and there is no CPU fallback -- if the CUDA
library is missing or no GPU is present the calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G6_B200_LIB", os.path.join(HERE, "csrc", "libsapporo.so"))

# Every symbol include/g6_b200.h declares (checked by tests/test_abi.py).
G6_SYMBOLS = [
    "g6_open_", "g6_close_", "g6_npipes_", "g6_set_tunit_", "g6_set_xunit_", "g6_set_ti_",
    "g6_set_j_particle_", "g6calc_firsthalf_", "g6calc_lasthalf_", "g6calc_lasthalf2_",
    "g6_initialize_jp_buffer_", "g6_flush_jp_buffer_", "g6_reset_", "g6_reset_fofpga_",
    "g6_read_neighbour_list_", "g6_get_neighbour_list_", "get_device_count",
    "g6_open", "g6_close", "g6_npipes", "g6_set_tunit", "g6_set_xunit", "g6_set_ti", "g6_set_j_particle",
    "g6calc_firsthalf", "g6calc_lasthalf", "g6calc_lasthalf2", "g6_initialize_jp_buffer",
    "g6_flush_jp_buffer", "g6_reset", "g6_reset_fofpga", "g6_reinitialize", "g6_get_number_of_pipelines",
    "g6_read_neighbour_list", "g6_get_neighbour_list", "g6_set_neighbour_list_sort_mode",
    "g6_get_neighbour_list_sort_mode", "g6_set_overflow_flag_test_mode", "force_j_particle_send", "get_j_part_data",
    "g6x_version", "g6x_set_stream", "g6x_set_refine", "g6x_set_j_offset", "g6x_set_j_particles", "g6x_predict",
    "g6x_calc_device", "g6x_device_chunk", "g6x_resolve_nn", "g6x_synchronize", "g6x_launch_count", "g6x_get_predicted",
    "g6x_read_predicted", "g6x_time_predictor", "g6x_set_variant", "g6x_fp32_peak",
    "g6x_peer_handle_bytes", "g6x_peer_alloc", "g6x_peer_attach", "g6x_peer_detach", "g6x_peer_error",
    "g6x_calc_device_allreduce", "g6x_hermite_init", "g6x_hermite_step", "g6x_hermite_evolve", "g6x_hermite_get_state", "g6x_hermite_set_shard", "g6x_latency_probe",
    "g6x_set_close_factor", "g6x_order_rebuilds", "g6x_device_count_open", "g6x_block_stats", "g6x_set_j_window",
]

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)

_lib = None


def load():
    """dlopen the library; raises if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "amuse_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.g6_open_.argtypes = [_pi]
    L.g6_close_.argtypes = [_pi]
    L.g6_set_ti_.argtypes = [_pi, _pd]
    L.g6_set_j_particle_.argtypes = [_pi, _pi, _pi, _pd, _pd, _pd, _dp, _dp, _dp, _dp, _dp]
    L.g6calc_firsthalf_.argtypes = [_pi, _pi, _pi, _ip, _dp, _dp, _dp, _dp, _dp, _pd, _dp]
    L.g6calc_firsthalf_.restype = None
    L.g6calc_lasthalf_.argtypes = [_pi, _pi, _pi, _ip, _dp, _dp, _pd, _dp, _dp, _dp, _dp]
    L.g6calc_lasthalf2_.argtypes = [_pi, _pi, _pi, _ip, _dp, _dp, _pd, _dp, _dp, _dp, _dp, _ip]
    L.g6_read_neighbour_list_.argtypes = [_pi]
    L.g6_get_neighbour_list_.argtypes = [_pi, _pi, _pi, _pi, _ip]
    L.g6x_set_stream.argtypes = [C.c_void_p, C.c_int]
    L.g6x_set_refine.argtypes = [C.c_int]
    L.g6x_set_close_factor.argtypes = [C.c_double, C.c_double]
    L.g6x_order_rebuilds.restype = C.c_longlong
    L.g6x_block_stats.argtypes = [C.c_void_p]
    L.g6x_set_j_window.argtypes = [C.c_int, C.c_int]
    L.g6x_set_j_offset.argtypes = [C.c_int]
    L.g6x_set_j_particles.argtypes = [C.c_int, C.c_void_p, C.c_int, _ip, C.c_void_p, _dp, C.c_void_p,
                                      C.c_void_p, _dp, _dp]
    L.g6x_predict.argtypes = [C.c_int, C.c_double]
    L.g6x_calc_device.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.g6x_calc_device_allreduce.argtypes = L.g6x_calc_device.argtypes
    L.g6x_peer_alloc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.g6x_peer_attach.argtypes = [C.c_void_p]
    L.g6x_hermite_init.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
    L.g6x_hermite_step.argtypes = [C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_void_p,
                                   C.c_void_p]
    L.g6x_hermite_evolve.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_longlong, _dp]
    L.g6x_hermite_evolve.restype = C.c_longlong
    L.g6x_hermite_get_state.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.get_j_part_data.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    L.g6x_hermite_set_shard.argtypes = [C.c_int, C.c_int]
    L.g6x_latency_probe.argtypes = [C.c_int, C.c_int]
    L.g6x_latency_probe.restype = C.c_double
    L.g6x_device_chunk.argtypes = [C.c_int]
    L.g6x_resolve_nn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.g6x_launch_count.restype = C.c_longlong
    L.g6x_read_predicted.argtypes = [C.c_int, _dp, _dp]
    L.g6x_time_predictor.argtypes = [C.c_int, C.c_int]
    L.g6x_time_predictor.restype = C.c_double
    L.g6x_set_variant.argtypes = [C.c_int]
    L.g6x_fp32_peak.argtypes = [C.c_int]
    L.g6x_fp32_peak.restype = C.c_double
    _lib = L
    return L


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class G6:
    """One opened g6 device, driven exactly like ph4 drives it
    (``src/amuse_ph4/src/gpu.cc:36-92`` initialise, ``:302-492`` force)."""

    def __init__(self, device=0):
        self.L = load()
        self.cid = C.c_int(device)
        rc = self.L.g6_open_(C.byref(self.cid))
        if rc != 0:
            raise RuntimeError("g6_open_(%d) failed with %d" % (device, rc))
        self.npipes = self.L.g6_npipes_()
        self.nj = 0

    def close(self):
        self.L.g6_close_(C.byref(self.cid))

    # -- j side ------------------------------------------------------------
    def set_j_particle(self, address, index, tj, dtj, mass, k18, j6, a2, v, x):
        """One g6_set_j_particle_ call (jdata::initialize_gpu, gpu.cc:83-88)."""
        self.L.g6_set_j_particle_(C.byref(self.cid), C.byref(C.c_int(address)), C.byref(C.c_int(index)),
                                  C.byref(C.c_double(tj)), C.byref(C.c_double(dtj)), C.byref(C.c_double(mass)),
                                  _c(k18), _c(j6), _c(a2), _c(v), _c(x))
        self.nj = max(self.nj, address + 1)

    def set_j_particles(self, ids, mass, pos, vel, acc=None, jerk=None, tj=None, address0=0, address=None):
        """Batched upload (g6x_set_j_particles).  acc/jerk are the physical values;
        the ABI's a2 = acc/2, j6 = jerk/6 scaling is applied here like gpu.cc:76-81."""
        n = len(mass)
        a2 = _c(np.asarray(acc) / 2) if acc is not None else None
        j6 = _c(np.asarray(jerk) / 6) if jerk is not None else None
        t = _c(tj) if tj is not None else None
        ad = _c(address, np.int32) if address is not None else None
        self.L.g6x_set_j_particles(n, ad.ctypes.data if ad is not None else None, int(address0),
                                   _c(ids, np.int32), t.ctypes.data if t is not None else None, _c(mass),
                                   j6.ctypes.data if j6 is not None else None,
                                   a2.ctypes.data if a2 is not None else None, _c(vel), _c(pos))
        hi = (int(np.max(address)) + 1) if address is not None else address0 + n
        self.nj = max(self.nj, hi)

    def set_ti(self, ti):
        self.L.g6_set_ti_(C.byref(self.cid), C.byref(C.c_double(ti)))

    # -- i side ------------------------------------------------------------
    def calc(self, ids, xi, vi, eps2, h2=None, nj=None, want_nn=True):
        """firsthalf + lasthalf[2] over npipes-sized chunks, the loop of
        idata::get_partial_acc_and_jerk_on_gpu (gpu.cc:369-407).
        Returns dict(acc, jerk, pot, nn) with nn = id of the nearest j."""
        ids = _c(ids, np.int32)
        xi = _c(xi)
        vi = _c(vi)
        ni = len(ids)
        nj = self.nj if nj is None else nj
        h2 = _c(h2) if h2 is not None else np.zeros(ni)
        acc = np.zeros((ni, 3))
        jerk = np.zeros((ni, 3))
        pot = np.zeros(ni)
        nn = np.full(ni, -1, dtype=np.int32)
        zeros3 = np.zeros((self.npipes, 3))
        zeros1 = np.zeros(self.npipes)
        e = C.c_double(eps2)
        cnj = C.c_int(nj)
        for i0 in range(0, ni, self.npipes):
            n = min(self.npipes, ni - i0)
            cn = C.c_int(n)
            sl = slice(i0, i0 + n)
            # row slices of C-contiguous arrays are contiguous: the library reads the caller's
            # arrays and writes the results in place, like the C callers do (no staging copies here)
            idc, xc, vc, hc = ids[sl], xi[sl], vi[sl], h2[sl]
            a, j, p = acc[sl], jerk[sl], pot[sl]
            self.L.g6calc_firsthalf_(C.byref(self.cid), C.byref(cnj), C.byref(cn), idc, xc, vc, zeros3[:n],
                                     zeros3[:n], zeros1[:n], C.byref(e), hc)
            if want_nn:
                self.L.g6calc_lasthalf2_(C.byref(self.cid), C.byref(cnj), C.byref(cn), idc, xc, vc, C.byref(e),
                                         hc, a, j, p, nn[sl])
            else:
                self.L.g6calc_lasthalf_(C.byref(self.cid), C.byref(cnj), C.byref(cn), idc, xc, vc, C.byref(e),
                                        hc, a, j, p)
        return dict(acc=acc, jerk=jerk, pot=pot, nn=nn)

    def read_neighbour_list(self):
        return self.L.g6_read_neighbour_list_(C.byref(self.cid))

    def get_neighbour_list(self, ipipe, maxlength=4096):
        lst = np.empty(maxlength, dtype=np.int32)
        n = C.c_int(0)
        rc = self.L.g6_get_neighbour_list_(C.byref(self.cid), C.byref(C.c_int(ipipe)),
                                           C.byref(C.c_int(maxlength)), C.byref(n), lst)
        return rc, n.value, lst[:min(n.value, maxlength)].copy()

    # -- extensions --------------------------------------------------------
    def read_predicted(self, nj=None):
        nj = self.nj if nj is None else nj
        pos = np.empty((nj, 3))
        vel = np.empty((nj, 3))
        self.L.g6x_read_predicted(nj, pos, vel)
        return pos, vel

    def predict(self, ti, nj=None):
        self.L.g6x_predict(self.nj if nj is None else nj, float(ti))

    def set_variant(self, v):
        if self.L.g6x_set_variant(int(v)) != 0:
            raise ValueError("unknown force-kernel variant %r" % (v,))

    def set_close_factor(self, k_close=-1.0, far_factor=-1.0):
        """FP64 radius factor / FAR-block factor of the pair classification (negative: unchanged)."""
        self.L.g6x_set_close_factor(float(k_close), float(far_factor))

    def launch_count(self):
        return int(self.L.g6x_launch_count())

    def synchronize(self):
        self.L.g6x_synchronize()
