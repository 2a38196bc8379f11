"""Opcode evidence for the shipped library: per kernel, how many tcgen05 / TMEM / TMA / cluster instructions its SASS holds.
    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
Runs on the build machine (cuobjdump needs no GPU). Mnemonics: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), UBLKCP = cp.async.bulk (untiled bulk copy,
the resizing preprocess), UCGABAR = barrier.cluster,
SYNCS = mbarrier, UTCATOMSWS = tcgen05.alloc / dealloc, LDGMC = multimem.ld_reduce (NVLS; multimem.st lowers to a plain STG.E.STRONG.SYS on the multicast address) - /opt/skills/guides/B200_PROFILING.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unet-lane-detection_b200", "libunet_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UCGABAR", "SYNCS", "UTCATOMSWS", "LDGMC", "HMMA", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            base = op.split(".")[0]
            if base.startswith("UCGABAR"):
                base = "UCGABAR"
            kernels[cur][base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines() if kernels else []
    for k, n in zip(kernels, names):
        demangle[k] = re.sub(r"\(.*", "", n.replace("(int)", "").replace("(bool)", ""))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (sm_100a)")
    print("# " + " ".join(f"{o:>12s}" for o in OPS) + "  kernel")
    tot = collections.Counter()
    for k, c in kernels.items():
        print("  " + " ".join(f"{c.get(o, 0):12d}" for o in OPS) + "  " + demangle.get(k, k))
        tot.update({o: c.get(o, 0) for o in OPS})
    print("# " + " ".join(f"{tot[o]:12d}" for o in OPS) + "  TOTAL")
    return 0


if __name__ == "__main__":
    sys.exit(main())
