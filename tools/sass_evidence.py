"""tools/sass_evidence.py -- static evidence from the built objects (no GPU needed): per kernel the registers, stack
(spills) and shared memory ptxas assigned (cuobjdump -res-usage), and the SASS mnemonics that prove the Blackwell paths
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, fp64 tensor pipe -> DMMA).
   python tools/sass_evidence.py > profiles/sass_evidence_r01.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gpc_b200", "csrc")
WANT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|DMMA|HMMA|IMMA|SYNCS|ELECT|LDGSTS|"
                  r"DFMA|DADD|DMUL)\b")


def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout


_dm = {}


def demangle(n):
    if n not in _dm:
        _dm[n] = re.sub(r"\(.*", "", run("c++filt", n).strip() or n).replace("void ", "")
    return _dm[n]


print("static evidence from gpc_b200/csrc/*.o (nvcc -gencode arch=compute_100a,code=sm_100a): python tools/sass_evidence.py\n")
for src in ("ozaki", "dense", "gpkern", "api", "lapack_api", "dist"):
    obj = os.path.join(CSRC, src + ".o")
    if not os.path.exists(obj):
        continue
    res = run("cuobjdump", "-res-usage", obj)
    kernels = re.findall(r"Function (\S+):\s*\n\s*(.*)", res)
    if not kernels:
        continue
    sass = run("cuobjdump", "-sass", obj)
    cur, per = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur:
            for w in WANT.findall(line):
                per[cur][w] += 1
    print("== %s.cu" % src)
    print("%-84s %5s %6s %8s  %s" % ("kernel", "regs", "stack", "smem(B)", "SASS mnemonics (count)"))
    for name, usage in kernels:
        g = lambda k: (re.search(k + r":(\d+)", usage) or [0, "0"])[1]
        c = per.get(name, {})
        mn = "  ".join("%s=%d" % kv for kv in sorted(c.items(), key=lambda x: -x[1]))
        print("%-84s %5s %6s %8s  %s" % (demangle(name)[:84], g("REG"), g("STACK"), g("SHARED"), mn))
    print()
