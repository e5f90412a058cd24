"""Per-source-line histogram of one kernel from `ncu -i rep --page source --csv --print-source cuda,sass` (gzipped ok):
executed warp instructions and stall samples by (file, line), heaviest first.
    python tools/ncu_source_hist.py src.csv[.gz] [top=40]"""
import csv
import gzip
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
cur_file, cur_line, cur_src = None, None, ""
inst = collections.Counter()
samp = collections.Counter()
srcs = {}
ops = collections.defaultdict(collections.Counter)
for row in csv.reader(f):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] in ("Function Name", "Line No", "Kernel Name"):
        continue
    if row[0]:                      # a source line
        cur_line = int(row[0])
        srcs[(cur_file, cur_line)] = row[1].strip()[:90]
        continue
    if len(row) >= 8 and row[2].startswith("0x"):
        try:
            n, s = int(row[7]), int(row[4])
        except ValueError:
            continue
        key = (cur_file, cur_line)
        inst[key] += n
        samp[key] += s
        toks = row[3].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[key][op.split(".")[0]] += n
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print("total executed warp instructions %d, stall samples %d" % (tot_i, tot_s))
print("%8s %6s %8s %6s  %s" % ("inst", "%", "samples", "%", "line"))
for key, n in inst.most_common(top):
    o = ", ".join("%s %d%%" % (k, 100 * v // max(n, 1)) for k, v in ops[key].most_common(4))
    print("%8.2fM %5.1f%% %8d %5.1f%%  %s:%d  %s   [%s]" % (n / 1e6, 100.0 * n / tot_i, samp[key], 100.0 * samp[key] / max(tot_s, 1),
                                                        key[0], key[1], srcs.get(key, ""), o))
