#!/usr/bin/env python
"""Opcode histogram of the built library's SASS: proof of what the hot kernels are made of (tcgen05 MMA kinds, TMEM
loads / stores, TMA loads / stores) without disassembling by hand.  usage: tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "universal-metal-flash-attention_b200", "lib", "libMFAFFI.so")
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCCP", "SYNCS", "MUFU", "FFMA2", "FADD2",
        "F2FP", "I2FP", "IMAD", "IDP", "PRMT", "FMNMX3", "LDL", "STL", "LDG", "STG", "LDS", "STS", "ATOMG", "REDG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {m.group(1): m.group(2) for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+ STACK:\d+)", res)}
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
print("# SASS opcode counts per kernel of lib/libMFAFFI.so (cuobjdump -sass; sm_100a).  UTCHMMA = tcgen05.mma kind::f16, UTCIMMA = kind::i8,")
print("# UTCQMMA = kind::f8f6f4, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA bulk tensor load / store, SYNCS = mbarrier ops.")
tot = collections.Counter()
for name, h in hist.items():
    tot.update(h)
    d = demangle(name)
    d = re.sub(r"mfa::\(anonymous namespace\)::", "", d)
    if not any(h[k] for k in ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTMALDG")) and sum(h.values()) < 3000:
        continue
    keys = " ".join(f"{k}={h[k]}" for k in KEYS if h[k])
    print(f"{d[:110]:110s} | {usage.get(name, '')} | n={sum(h.values())} | {keys}")
print("# whole library: " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
