"""SASS opcode summary of libdgp_b200.so (cuobjdump -sass): per kernel, the counts of the instructions that prove the
Blackwell data path (tcgen05 MMA / TMEM / TMA / bulk copies / packed fp32 math).  CPU only.

    python tools/sass_summary.py > profiles/rNN_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepgraphpose_b200", "libdgp_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG.4D.IM2COL.2CTA", "UTMALDG.4D.IM2COL", "UTMALDG.4D", "UTMALDG.2D",
        "UTMASTG", "LDTM", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "UCGABAR", "REDUX", "FFMA2", "FADD2",
        "F2FP", "HMNMX2", "VHMNMX", "SHFL", "MUFU.EX2", "ATOMG", "BAR.SYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|dgp::", "", name).split("(")[0]
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[name][k] += 1
                    break
    print("# SASS opcode summary of libdgp_b200.so (cuobjdump -sass, code objects: %s)\n" % ", ".join(archs))
    print("Counts of the instructions that matter per kernel (static occurrences in the binary, not executions).  UTCHMMA = tcgen05.mma,")
    print("`.2CTA` = cta_group::2; UTMALDG / UTMASTG = TMA tensor loads / stores (`.IM2COL` = im2col mode); LDTM = tcgen05.ld;")
    print("UTCBAR = tcgen05.commit (`.MULTICAST` to both CTAs of a pair); UBLKCP = cp.async.bulk; FFMA2 / FADD2 = packed fp32x2.\n")
    used = [k for k in KEYS if any(c[k] for c in kernels.values())]
    print("| kernel | SASS instr | " + " | ".join(used) + " |")
    print("|---|---|" + "---|" * len(used))
    for n, c in kernels.items():
        if c["_total"] < 40:
            continue
        print("| `%s` | %d | " % (n[:70], c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in used) + " |")


if __name__ == "__main__":
    sys.exit(main())
