"""Per-kernel counts of the Blackwell-native SASS opcodes in the shipped library (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
UTCBAR (tcgen05.commit), MUFU.EX2.  Usage: python tools/sass_table.py > profiles/rN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genomics-lm_b200", "codonlm_b200", "libcgpt_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "MUFU.EX2", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
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
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in OPS:
                if op == key or op.startswith(key + "."):
                    counts[cur][key] += 1
            counts[cur]["_n"] += 1
    names = list(counts)
    if names:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): opcode counts per kernel")
    print(f"{'kernel':86s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in OPS))
    tot = collections.Counter()
    for fn, c in counts.items():
        name = demangle.get(fn, fn)
        name = re.sub(r"cgpt::\(anonymous namespace\)::|cgpt::<unnamed>::|cgpt::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{name[:86]:86s} {c['_n']:7d} " + " ".join(f"{c[k]:8d}" for k in OPS))
        tot.update(c)
    print(f"{'TOTAL':86s} {tot['_n']:7d} " + " ".join(f"{tot[k]:8d}" for k in OPS))


if __name__ == "__main__":
    sys.exit(main())
