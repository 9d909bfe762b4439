"""Opcode counts per kernel of the in-tree libcnf.so (cuobjdump -sass): the mnemonics that prove which hardware path a
kernel uses (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, FFMA2 = packed fp32 FMA).
usage: python tools/sass_summary.py [lib] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "arl_conditional_normalizing_flows_b200", "libcnf.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "HMMA", "DFMA",
        "DADD", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "MUFU", "BAR"]
PREFIX = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "LDGSTS", "UBLKCP", "SYNCS")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
mangled = re.findall(r"Function : (\S+)", sass)
names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")
counts, order, cur, i = {}, [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[i]).replace("void ", "")
        i += 1
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for w in WANT:
            if op == w or (w in PREFIX and op.startswith(w)):
                counts[cur][w] += 1
print(f"# SASS opcode counts per kernel, {os.path.basename(lib)} (sm_100a); made by tools/sass_summary.py")
print("# columns: total instructions, then the non-zero counts among: " + " ".join(WANT))
tot = collections.Counter()
for k in order:
    c = counts[k]
    tot.update(c)
    print(f"{k[:150]}\n    total={c['_total']} " + " ".join(f"{w}={c[w]}" for w in WANT if c[w]))
print("\n# whole library: " + " ".join(f"{w}={tot[w]}" for w in WANT if tot[w]))
