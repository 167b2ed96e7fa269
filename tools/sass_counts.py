"""Counts of the Blackwell-specific SASS mnemonics per kernel of libhonerf_b200.so (cuobjdump -sass): the evidence that the hot
kernels are tcgen05 / TMEM / bulk-copy code and not a recompiled mma.sync path.  Writes profiles/r02_sass_counts.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ho-nerf_b200", "libhonerf_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "IMMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        short = re.sub(r"^void ", "", short)
        rows.append((short, c))
    dst = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02_sass_counts.txt")
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass ho-nerf_b200/libhonerf_b200.so (sm_100a): instruction counts per kernel\n")
        f.write("# UTCHMMA = tcgen05.mma (kind::f16 / tf32), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,\n")
        f.write("# UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier ops, HMMA / IMMA = legacy mma.sync (expected: 0)\n")
        f.write("%-64s " % "kernel" + " ".join("%8s" % k for k in MNEMONICS) + "\n")
        for short, c in sorted(rows):
            if not any(c[k] for k in MNEMONICS):
                continue
            f.write("%-64s " % short[:64] + " ".join("%8d" % c[k] for k in MNEMONICS) + "\n")
        tot = collections.Counter()
        for _, c in rows:
            tot.update(c)
        f.write("%-64s " % "TOTAL" + " ".join("%8d" % tot[k] for k in MNEMONICS) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
