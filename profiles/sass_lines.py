"""Static SASS instruction count per source region of one kernel:
   python profiles/sass_lines.py FILE.sass FUNCTION_SUBSTRING   (FILE.sass = `nvdisasm -g -c x.cubin`)"""
import re
import sys
from collections import Counter

path, fun = sys.argv[1], sys.argv[2]
REGIONS = [("plb_device.cuh", 232, 411, "setup_consts"), ("plb_device.cuh", 454, 798, "lane_eval"),
           ("plb_device.cuh", 120, 220, "grp primitives"), ("plb_device.cuh", 800, 1200, "factor/solve iso"),
           ("plb_device.cuh", 1200, 1800, "factor/solve th")]
cnt, ops = Counter(), Counter()
inside, cur = False, ("?", 0)
for ln in open(path):
    if ln.startswith(".text."):
        inside = fun in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        f, l = cur
        name = f
        for rf, a, b, nm in REGIONS:
            if f == rf and a <= l <= b:
                name = nm
        if f == "plb_variant.cuh" or f == "plb_tick.cuh" or f == "plb_integrator.cuh":
            name = f"{f}:{l // 10 * 10}"
        cnt[name] += 1
        ops[m.group(2).split(".")[0]] += 1
tot = sum(cnt.values())
print("total", tot)
for k, v in cnt.most_common(40):
    print(f"{v:7d} {100 * v / tot:5.1f}%  {k}")
print(" ".join(f"{k}:{v}" for k, v in ops.most_common(25)))
