"""Per-kernel SASS mnemonic counts of the built library (B200_PROFILING.md, "What proves a Blackwell-native kernel"):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, HMMA = legacy mma.sync (none),
FFMA for the fp32 kernels.  Runs without a GPU:  python profiles/sass_evidence.py > profiles/<tag>_sass_mnemonics.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "realpdebench_b200", "lib", "libb200fno.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "FFMA2", "MUFU", "LDGSTS",
        "LDG", "STG", "instructions"]
counts, kernel = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kernel = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        kernel = re.sub(r"\(anonymous namespace\)::", "", kernel)
        n, base = 1, kernel
        while kernel in counts:  # template instantiations demangle to the same prefix only if truncated
            n += 1
            kernel = f"{base}#{n}"
        counts[kernel] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kernel:
        op = m.group(1)
        c = counts[kernel]
        c["instructions"] += 1
        if re.match(r"UTC[A-Z]*MMA", op):
            c["UTC*MMA"] += 1
        elif op in KEYS:
            c[op] += 1
print("kernel," + ",".join(KEYS))
for k, c in counts.items():
    print('"' + k.replace("b200fno::", "") + '",' + ",".join(str(c[x]) for x in KEYS))
