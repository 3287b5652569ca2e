"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares,
and the kernels of ONE PVCNN2Prior forward.   python tools/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = list(csv.DictReader(lines))


def us(x):
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)


def short(n):
    return re.sub(r"\(.*", "", n).replace("lion::", "").replace("void ", "")


idx = [i for i, x in enumerate(recs) if "k_make_coords" in x["Kernel Name"]]
# one U-Net forward = from one k_make_coords to the next k_gp_posemb / k_make_coords
s = idx[len(idx) // 2]
e = next((i for i in range(s + 1, len(recs)) if "k_make_coords" in recs[i]["Kernel Name"] or "k_gp_posemb" in recs[i]["Kernel Name"]
          or "k_pack" in recs[i]["Kernel Name"]), len(recs))
agg = collections.OrderedDict()
tot = 0.0
for x in recs[s:e]:
    n = short(x["Kernel Name"])
    if "k_ddpm" in n or "distribution" in n:
        continue
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us(x)
    tot += us(x)
print("one PVCNN2Prior forward: %d kernels, %.1f us summed (serialised, cold cache)" % (e - s, tot))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-34s n=%4d  total=%9.1f us  avg=%8.1f  share=%5.1f%%" % (k[:34], n, t, t / n, 100 * t / tot))
g = [us(x) for x in recs if "k_gp_linear" in x["Kernel Name"]]
if g:
    print("global prior: k_gp_linear avg %.1f us x %d per step" % (sum(g) / len(g), 35))
