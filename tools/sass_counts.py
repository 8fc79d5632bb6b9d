"""Per-kernel SASS evidence for the shipped library: counts of the Blackwell-relevant mnemonics in every
sm_100a kernel of libhssb200.so (cuobjdump -sass).  FP64 has no tcgen05 kind, so the tensor path is DMMA
(mma.sync m8n8k4 f64); TMA shows up as UBLKCP (1-D bulk copy) and UTMALDG (tiled tensor load); SYNCS are the
mbarrier operations.  Usage: python tools/sass_counts.py > profiles/sass_r02.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "hssmatrices.jl_b200", "lib", "libhssb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = ("DMMA", "DFMA", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDS", "STG", "LDG", "BAR")
kern, counts, arch = None, collections.OrderedDict(), set()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("hssb::", "").replace("(int)", "").replace("(bool)", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if kern:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for n in names:
                if op.startswith(n):
                    counts[kern][n] += 1
print("libhssb200.so: cubins for", ", ".join(sorted(arch)), "-", len(counts), "kernels")
print("%-64s" % "kernel", " ".join("%7s" % n for n in names))
tot = collections.Counter()
for k, c in counts.items():
    print("%-64s" % k[:64], " ".join("%7d" % c[n] for n in names))
    tot.update(c)
print("%-64s" % "TOTAL", " ".join("%7d" % tot[n] for n in names))
