"""Opcode histogram per kernel of libdynam3d_b200.so (cuobjdump -sass): evidence that the tcgen05 / TMEM / TMA path is what is compiled.
    python tools/sass_hist.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynam3d_b200", "libdynam3d_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "HMMA", "LDGSTS", "SYNCS", "MUFU", "FFMA2", "FADD2",
       "LDSM", "STS", "LDS", "LDG", "STG", "ATOMG", "RED", "DFMA", "F2FP")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        mods = m.group(2)
        hist[kern][op + (".2CTA" if ".2CTA" in mods else "")] += 1
print(f"# {os.path.basename(SO)}: SASS opcode counts per kernel (static instruction counts; sm_100a)")
tot = collections.Counter()
for k in sorted(hist):
    h = hist[k]
    n = sum(h.values())
    sel = {op: c for op, c in h.items() if any(op.startswith(p) for p in KEY)}
    for op, c in sel.items():
        tot[op] += c
    print(f"{k}  [{n} instr]  " + "  ".join(f"{op}:{c}" for op, c in sorted(sel.items(), key=lambda kv: -kv[1])))
print("# totals: " + "  ".join(f"{op}:{c}" for op, c in sorted(tot.items(), key=lambda kv: -kv[1])))
