"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/g6_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from amuse_b200 import g6lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "g6_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g6x?_?\w*|g6calc_\w+|get_device_count|get_j_part_data|force_j_particle_send)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(g6lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return g6lib.LIB_PATH


def test_header_and_python_symbol_lists_agree():
    assert sorted(g6lib.G6_SYMBOLS) == _header_symbols()


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in _header_symbols():
        assert hasattr(lib, s), "missing export %s" % s


def test_reference_callers_link_symbols_are_covered(built_lib):
    # the 16 undefined g6 symbols of ph4's -DGPU objects (grape.h:8-108)
    need = ["g6_open_", "g6_close_", "g6_npipes_", "g6_set_tunit_", "g6_set_xunit_", "g6_set_ti_",
            "g6_set_j_particle_", "g6calc_firsthalf_", "g6calc_lasthalf_", "g6calc_lasthalf2_",
            "g6_initialize_jp_buffer_", "g6_flush_jp_buffer_", "g6_reset_", "g6_reset_fofpga_",
            "g6_read_neighbour_list_", "g6_get_neighbour_list_", "get_device_count"]
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib]).decode()
    have = set(l.split()[-1] for l in out.splitlines() if l.strip())
    assert not [s for s in need if s not in have]


def test_alias_libg6_exists(built_lib):
    assert os.path.exists(os.path.join(os.path.dirname(built_lib), "libg6.so"))


def test_library_is_sm100a_cuda_code(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_npipes_and_stubs_without_device(built_lib):
    lib = g6lib.load()
    assert lib.g6_npipes_() == int(os.environ.get("G6_B200_NPIPES", 16384))
    assert lib.g6x_version() >= 100
    assert lib.g6_set_tunit_(None) == 0 and lib.g6_set_xunit_(None) == 0


def test_open_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: without a device g6_open_ aborts the process with a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import ctypes,sys; sys.path.insert(0,%r); from amuse_b200 import g6lib; L=g6lib.load(); "
            "i=ctypes.c_int(0); L.g6_open_(ctypes.byref(i)); print('SURVIVED')" % ROOT)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout and "no CUDA device" in r.stderr


def test_install_lays_out_what_amuse_configure_looks_for(built_lib, tmp_path):
    """lib/sapporo_light/Makefile:28-101 + support/shared/m4/amuse_lib.m4:17-72: libraries under lib/, the
    header under include/ (also as g6lib.h), and the sapporo_light / g6lib pkg-config modules."""
    import subprocess
    prefix = str(tmp_path / "prefix")
    csrc = os.path.join(ROOT, "amuse_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "install", "PREFIX=" + prefix])
    for rel in ("lib/libsapporo.so", "lib/libg6.so", "include/g6_b200.h", "include/g6lib.h",
                "lib/pkgconfig/sapporo_light.pc", "lib/pkgconfig/g6lib.pc"):
        assert os.path.exists(os.path.join(prefix, rel)), rel
    pc = open(os.path.join(prefix, "lib/pkgconfig/sapporo_light.pc")).read()
    assert "Name: sapporo_light" in pc and "-lsapporo" in pc and ("prefix=" + prefix) in pc
    pc = open(os.path.join(prefix, "lib/pkgconfig/g6lib.pc")).read()
    assert "Name: g6lib" in pc and "-lg6" in pc
    lib = ctypes.CDLL(os.path.join(prefix, "lib", "libg6.so"))
    assert lib.g6_npipes_() == 16384
    subprocess.check_call(["make", "-s", "-C", csrc, "uninstall", "PREFIX=" + prefix])
    assert not os.path.exists(os.path.join(prefix, "lib/libsapporo.so"))
