"""Static checks of the built library's SASS (no GPU needed: cuobjdump reads the cubin).  They pin what DESIGN.md §3.1
claims about the hot loop of the bench's force kernel: the FAR pair block costs at most 30 FP32 lane-operations per
pair (the 60-flop convention's 100 % line), is packed arithmetic throughout, carries no mask instructions, and the
j tiles arrive by TMA bulk copies."""
import os
import re
import shutil
import subprocess
import sys
from collections import Counter

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "amuse_b200", "csrc", "libsapporo.so")
KERNEL = "force_fast_kernelILi2ELb1ELb1ELi2ELb1E"   # <IPT 2, NN, NR, 2 CTAs/SM, EPS0>: what bench.py launches


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(LIB) or shutil.which("cuobjdump") is None:
        pytest.skip("library not built or cuobjdump missing")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = [f for f in funcs if KERNEL in f.split("\n", 1)[0]]
    assert body, "bench kernel instantiation not found in the library"
    return body[0]


def _blocks(body):
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    targets = set()
    for a, t in ins:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and re.search(r"\b(BRA|BSSY|CALL)\b", t):
            targets.add(int(m.group(1), 16))
    blocks, cur = [], []
    for a, t in ins:
        if a in targets and cur:
            blocks.append(cur)
            cur = []
        cur.append(t)
        if re.search(r"\b(BRA|RET|EXIT|BRX|CALL)\b", t):
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    return blocks


def _op(t):
    return re.match(r"(@!?U?P\d\s+)?(\S+)", t).group(2).split(".")[0]


def test_far_and_near_pair_blocks(sass):
    blocks = _blocks(sass)
    stats = []
    for b in blocks:
        ops = Counter(_op(t) for t in b)
        packed = ops["FFMA2"] + ops["FADD2"] + ops["FMUL2"]
        if packed >= 500 and ops["MUFU"] == 64:      # a whole (32 j x 2 i per lane) group, fully unrolled
            scalar = ops["FFMA"] + ops["FADD"] + ops["FMUL"]
            stats.append(((2 * packed + scalar) / 64.0, ops, len(b)))
    assert len(stats) >= 2, "expected the unrolled FAR and NEAR group blocks"
    stats.sort(key=lambda s: s[0])
    far, near = stats[0], stats[-1]
    assert far[0] <= 30.0, "FAR block: %.1f FP32 lane-operations per pair" % far[0]
    assert near[0] <= 38.0, "NEAR block: %.1f FP32 lane-operations per pair" % near[0]
    # no masks and no neighbour search in the FAR block; broadcast shared-memory loads only
    for bad in ("FSEL", "FSETP", "ISETP", "FMNMX", "FMNMX3", "LDG", "ATOM", "ATOMS"):
        assert far[1][bad] == 0, "FAR block contains %s" % bad
    assert far[1]["LDS"] == 64            # A and C of 32 j (hi parts and velocities): the lo parts are not even loaded
    assert near[1]["LDS"] == 96           # A, B and C
    assert far[2] <= 1100, "FAR block has %d instructions for 64 pairs per lane" % far[2]


def test_tiles_arrive_by_tma_bulk_copies(sass):
    assert "UBLKCP" in sass, "no TMA bulk copy in the force kernel"
    assert "SYNCS" in sass, "no mbarrier instructions in the force kernel"
    assert "ATOMS" in sass, "the per-stage release counter (shared-memory atomic) is missing"


if __name__ == "__main__":
    sys.exit(pytest.main([__file__, "-q"]))
