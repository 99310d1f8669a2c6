"""Dump the SASS of one kernel of the built library and summarise its hottest loop (the innermost
backward branch with the most FFMA2): instruction mix, register words read per instruction class
(.reuse operands are free), and the resulting operand-bandwidth estimate.
Usage: python tools/sass_loop.py <mangled-name-substring> [--lib path] [--print]"""
import re
import subprocess
import sys
from collections import Counter

lib = "amuse_b200/csrc/libsapporo.so"
args = sys.argv[1:]
if "--lib" in args:
    lib = args[args.index("--lib") + 1]
name = args[0]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = [f for f in funcs if name in f.split("\n", 1)[0]]
assert body, "no function matching " + name
body = body[0]
print("function:", body.split("\n", 1)[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: k for k, (a, _) in enumerate(ins)}
loops = []
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)* (?:!?U?P\d, )?0x([0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_idx:
            seg = ins[addr_idx[tgt]:k + 1]
            loops.append((sum("FFMA2" in x[1] for x in seg), tgt, a, seg))
loops.sort(reverse=True)
# innermost: the smallest loop among those with the most FFMA2 per instruction
for l in sorted(loops, key=lambda l: l[1]):
    print("  loop 0x%x..0x%x  %d instr, %d FFMA2, %d FSEL" % (l[1], l[2], len(l[3]), l[0], sum("FSEL" in x[1] for x in l[3])))
sel = [a for a in args if a.startswith("--loop=")]
if sel:
    best = [l for l in loops if l[1] == int(sel[0][7:], 16)][0]
else:
    best = max(loops, key=lambda l: (l[0] / max(len(l[3]), 1), l[0]))
nf, tgt, a, seg = best
print("hot loop 0x%x..0x%x: %d instructions" % (tgt, a, len(seg)))
mix = Counter()
words = Counter()
for _, t in seg:
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0].split(".")[0]
    mix[op] += 1
    ops = t2.split(None, 1)[1] if " " in t2 else ""
    parts = [x.strip() for x in ops.split(",")]
    srcs = parts[1:] if op not in ("BRA", "BAR", "STS", "ST", "STG") else parts
    w = 0
    seen = set()
    for sreg in srcs:
        m = re.match(r"[-|~!]?(R\d+)((?:\.\w+)*)", sreg)
        if not m:
            continue
        r, mods = m.group(1), m.group(2)
        if ".reuse" in mods or (r, mods) in seen:
            continue
        seen.add((r, mods))
        w += 2 if "F32x2" in mods or ".64" in mods else 1
    words[op] += w
hist = Counter()
model = 0.0
for _, t in seg:
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0].split(".")[0]
    ops = t2.split(None, 1)[1] if " " in t2 else ""
    parts = [x.strip() for x in ops.split(",")]
    srcs = parts[1:]
    w = 0
    seen = set()
    for sreg in srcs:
        m = re.match(r"[-|~!]?(R\d+)((?:\.\w+)*)", sreg)
        if not m:
            continue
        r, mods = m.group(1), m.group(2)
        if ".reuse" in mods or (r, mods) in seen:
            continue
        seen.add((r, mods))
        w += 2 if "F32x2" in mods or ".64" in mods else 1
    if op in ("FFMA2", "FADD2", "FMUL2"):
        hist[(op, w)] += 1
        model += max(2.0, w / 2.0)
    elif op in ("FFMA", "FADD", "FMUL"):
        model += max(1.0, w / 2.0)
    elif op not in ("BRA", "LDS", "STS") and not op.startswith("U"):
        model += w / 2.0
print("packed ops by register words read:", dict(sorted(hist.items())))
print("serial operand-fetch model (max(pipe cycles, words/2) per instruction): %.0f cycles per loop iteration" % model)
tot_w = sum(words.values())
print("mix:", dict(mix.most_common()))
print("register words read:", dict(words.most_common()), "total", tot_w)
fma = sum(v for k, v in mix.items() if k in ("FFMA2", "FADD2", "FMUL2"))
fma1 = sum(v for k, v in mix.items() if k in ("FFMA", "FADD", "FMUL"))
print("FMA-pipe cycles (2 per packed op, 1 per scalar): %d; issue slots %d; operand cycles at 2 words/clk: %.0f" % (
    2 * fma + fma1, len(seg), tot_w / 2))
if "--print" in args:
    for a_, t in seg:
        print("  %05x  %s" % (a_, t))
