"""Static evidence from the built library (no GPU needed): per-kernel resource usage (cuobjdump -res-usage) and counts of the
SASS mnemonics that prove the tcgen05 / TMEM / TMA path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor
load, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM allocation, REDG = fp32 reductions of the weight-gradient epilogue).
usage: python tools/sass_summary.py [round]  ->  profiles/sass_<round>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vaenar_tts_b200", "libvaenar_sm100.so")
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r1"
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "REDG", "ATOMG", "SYNCS", "BAR.SYNC", "HMMA", "FFMA", "DFMA", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
names = demangle(list(counts))


def short(n):
    n = re.sub(r"\(.*", "", names.get(n, n)).replace("void ", "").replace("vb::", "")
    return n[:72]


rows = []
for k, c in counts.items():
    u = usage.get(k, {})
    rows.append((short(k), u.get("REG", "?"), u.get("STACK", "?"), c["_all"], [c[m] for m in MN]))
rows.sort(key=lambda r: (-r[4][0], -r[3]))
with open(os.path.join(ROOT, "profiles", f"sass_{ROUND}.md"), "w") as f:
    f.write(f"# Static SASS / resource summary of libvaenar_sm100.so (sm_100a) -- {ROUND}\n\n"
            "`python tools/sass_summary.py` (cuobjdump -res-usage / -sass; no GPU).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, "
            "UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = TMEM alloc/dealloc, REDG = "
            "red.global.add (weight-gradient epilogue / bias-gradient column sums), HMMA = legacy mma.sync (must be 0).\n\n"
            "| kernel | regs | stack | SASS instrs | " + " | ".join(MN) + " |\n|---|---|---|---|" + "---|" * len(MN) + "\n")
    for name, reg, stack, n, cs in rows:
        f.write(f"| {name} | {reg} | {stack} | {n} | " + " | ".join(str(x) for x in cs) + " |\n")
    tot = [sum(r[4][i] for r in rows) for i in range(len(MN))]
    f.write(f"| **total ({len(rows)} kernels)** | | | {sum(r[3] for r in rows)} | " + " | ".join(str(x) for x in tot) + " |\n")
print("wrote profiles/sass_%s.md" % ROUND, len(rows), "kernels")
