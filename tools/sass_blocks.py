"""Straight-line blocks of one kernel's SASS with the most packed-FP32 arithmetic: the fully unrolled FAR and NEAR
pair blocks of force_fast_kernel have no backward branch of their own, so tools/sass_loop.py cannot isolate them.
Prints each block's instruction mix and lane-operations per pair, and the listing of the first one.
Usage: python tools/sass_blocks.py <mangled-name-substring> [--lib path] [--pairs-per-block 64] [--list N] [--which K]"""
import re
import subprocess
import sys
from collections import Counter

lib = "amuse_b200/csrc/libsapporo.so"
args = sys.argv[1:]
if "--lib" in args:
    lib = args[args.index("--lib") + 1]
pairs = int(args[args.index("--pairs-per-block") + 1]) if "--pairs-per-block" in args else 64
nlist = int(args[args.index("--list") + 1]) if "--list" in args else 120
which = int(args[args.index("--which") + 1]) if "--which" in args else 0
name = args[0]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = [f for f in funcs if name in f.split("\n", 1)[0]][0]
print("function:", body.split("\n", 1)[0])
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
    cur.append((a, t))
    if re.search(r"\b(BRA|RET|EXIT|BRX|CALL)\b", t):
        blocks.append(cur)
        cur = []
if cur:
    blocks.append(cur)


def opname(t):
    m = re.match(r"(@!?U?P\d\s+)?(\S+)", t)
    return m.group(2).split(".")[0]


def packed(b):
    return sum(1 for _, t in b if opname(t) in ("FFMA2", "FADD2", "FMUL2"))


best = sorted(blocks, key=lambda b: -packed(b))[:3]
for k, b in enumerate(best):
    ops = Counter(opname(t) for _, t in b)
    fp = ops["FFMA2"] + ops["FADD2"] + ops["FMUL2"]
    sc = ops["FFMA"] + ops["FADD"] + ops["FMUL"]
    print("block 0x%x..0x%x: %d instructions; packed FP32 %d, scalar FP32 %d, MUFU %d, LDS %d, other %d" % (
        b[0][0], b[-1][0], len(b), fp, sc, ops["MUFU"], ops["LDS"], len(b) - fp - sc - ops["MUFU"] - ops["LDS"]))
    print("   per pair (%d pairs per lane and block): %.1f FP32 lane-operations (packed counted twice) + %.2f MUFU, %.1f issue slots"
          % (pairs, (2 * fp + sc) / pairs, ops["MUFU"] / pairs, len(b) / pairs))
    print("   mix:", dict(ops.most_common()))
print("\nlisting of block %d (first %d instructions):" % (which, nlist))
for a, t in best[which][:nlist]:
    print("  /*%05x*/ %s" % (a, t))
