"""Join an ncu SASS source page (csv) with nvdisasm --print-line-info of the same cubin: executed warp instructions, lanes per
instruction (thread instructions / warp instructions) and stall samples per source line.
Usage: sass_lines.py NCU_SOURCE.csv[.gz] ALL.sass KERNEL_SUBSTRING [TOP]"""
import gzip
import collections
import csv
import re
import sys

src_csv, sass, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: address (offset in function) -> (file, line, inlined-at chain)
addr2line = {}
infunc = False
cur = None
for line in open(sass):
    if line.startswith("//--------------------- .text."):
        infunc = kern in line
        cur = None
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", line)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(gzip.open(src_csv, "rt") if src_csv.endswith(".gz") else open(src_csv)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
it = h.index("Thread Instructions Executed") if "Thread Instructions Executed" in h else None
base = None
per = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
for r in rows[hi + 1:]:
    try:
        a = int(r[ia], 16); n = int(r[ie]); s = int(r[isamp] or 0)
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    k = addr2line.get(a - base)
    per[k] += n; samp[k] += s
    if it is not None:
        try:
            thr[k] += int(r[it])
        except ValueError:
            pass
tot, tots = sum(per.values()), sum(samp.values())
print("total warp instructions %d, lanes per instruction %.1f, samples %d" % (tot, sum(thr.values()) / max(tot, 1), tots))
for k, n in per.most_common(top):
    print("%10d %5.1f%%  lanes %4.1f  samples %5.1f%%  %s" % (n, 100.0 * n / tot, thr[k] / max(n, 1), 100.0 * samp[k] / max(tots, 1), k))
