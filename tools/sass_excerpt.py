"""Writes profiles/rNN/sass_conv_tc_excerpt.txt: per kernel of libsrdiff_b200.so the count of the SASS mnemonics that
prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md), plus the first occurrence of each, from `cuobjdump -sass`.
    python tools/sass_excerpt.py profiles/r02/sass_conv_tc_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dgm_img_super_resolution_b200", "libsrdiff_b200.so")
out = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "SYNCS", "UCGABAR", "UTCATOMSWS", "ELECT", "MUFU.EX2", "MUFU.RCP"]
cur, stats, first, arch = None, collections.OrderedDict(), {}, set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None or "/*" not in line:
        continue
    for mn in MN:
        if re.search(r"\b" + re.escape(mn), line):
            stats[cur][mn] += 1
            first.setdefault((cur, mn), line.split("*/")[1].split(";")[0].strip() if "*/" in line else line.strip())
    stats[cur]["_lines"] += 1
with open(out, "w") as f:
    f.write("cuobjdump -sass dgm_img_super_resolution_b200/libsrdiff_b200.so   (arch: %s)\n" % ", ".join(sorted(arch)))
    f.write("mnemonic counts per kernel (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / "
            "store, UTCBAR = tcgen05.commit, UTMAPF = prefetch.tensormap, SYNCS = mbarrier ops)\n\n")
    for k, c in stats.items():
        if not any(c[m] for m in MN[:5]):
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
        f.write("%s\n  SASS instructions: %d\n  " % (name, c["_lines"]))
        f.write("  ".join("%s %d" % (m, c[m]) for m in MN if c[m]) + "\n")
        for m in MN[:6]:
            if (k, m) in first:
                f.write("    first %-8s: %s\n" % (m, first[(k, m)]))
        f.write("\n")
    others = [k for k, c in stats.items() if not any(c[m] for m in MN[:5])]
    f.write("kernels without tensor-core / TMA instructions (elementwise, converters, validation kernels): %d\n" % len(others))
print(open(out).read()[:3000])
