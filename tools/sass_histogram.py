"""SASS opcode evidence for the in-tree library: per kernel, counts of the Blackwell-specific mnemonics (tcgen05 MMA =
UTCHMMA / UTCQMMA..., TMEM loads/stores = LDTM / STTM, TMA = UTMALDG / UBLKCP / UTMASTG, mbarrier = SYNCS, warp MMA =
HMMA / *MMA.16816 ...) and the ten most frequent opcodes.  Runs on the CPU (cuobjdump only).

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.md
"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ctrlsim_b200", "lib", "libctrlsim_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS",
        "HMMA", "IMMA", "DMMA", "LDGSTS", "LDSM", "FFMA", "DFMA", "MUFU"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
kern, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        per[kern][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode histogram of `{os.path.relpath(LIB, ROOT)}` (`cuobjdump -sass`, sm_100a)\n")
print("Per kernel: instruction count, the tensor-core / TMEM / TMA / mbarrier mnemonics present, and the top opcodes.\n")
tot = collections.Counter()
print("| kernel | instr | " + " | ".join(KEYS) + " | top opcodes |")
print("|---|---|" + "---|" * len(KEYS) + "---|")
for (k, c), name in zip(per.items(), names):
    tot.update(c)
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ctrlsim::", "")
    top = ", ".join(f"{o} {n}" for o, n in c.most_common(6))
    print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + f" | {top} |")
print("\nLibrary totals: " + ", ".join(f"{x} {tot[x]}" for x in KEYS if tot[x]))
