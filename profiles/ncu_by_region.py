"""Dynamic instruction counts / stall samples per source region: joins `ncu --page source --csv` (SASS rows, in
program order) with `nvdisasm -g -c` of the same build (source line per instruction).
   python profiles/ncu_by_region.py SRC.csv FILE.sass FUNCTION_SUBSTRING [n_systems]"""
import csv
import re
import sys
from collections import Counter

src_csv, sass, fun = sys.argv[1:4]
nsys = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
lines = []
inside, cur = False, ("?", 0)
for ln in open(sass):
    if ln.startswith(".text."):
        inside = fun in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*)", ln)
    if m:
        lines.append((cur, m.group(1)))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
body = rows[2:]
assert len(body) == len(lines), (len(body), len(lines))
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ex, smp, st = Counter(), Counter(), {}
fine = len(sys.argv) > 5
for ((f, l), _ins), r in zip(lines, body):
    key = f"{f}:{l}" if fine else f"{f}:{l // 20 * 20}"
    ex[key] += int(r[iex]); smp[key] += int(r[ismp])
    d = st.setdefault(key, Counter())
    for i in stall_cols:
        if r[i] not in ("", "-", "0"):
            d[hdr[i]] += int(r[i])
tot, tots = sum(ex.values()), sum(smp.values())
print(f"total executed {tot}  per system {tot / nsys:.0f}   samples {tots}")
for k, v in smp.most_common(45):
    top = ", ".join(f"{a[6:]} {b}" for a, b in st[k].most_common(3))
    print(f"{ex[k] / nsys:9.1f} inst {100 * ex[k] / tot:5.1f}%   samples {100 * smp[k] / tots:5.1f}%  {k:32s} {top}")
