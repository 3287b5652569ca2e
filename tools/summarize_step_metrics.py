"""Per-kernel-class table of one profiled denoising step (tools/profile_step.py under ncu --metrics ... --csv):
time share, achieved HBM GB/s (dram bytes / duration), tensor-pipe % -- for profiles/.

    python tools/summarize_step_metrics.py gpurun_out/step.csv > profiles/r02_step_kernel_classes.txt"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))          # (id) -> metric -> value
    names = {}
    for r in rows[1:]:
        try:
            kid = r[col["ID"]]
            names[kid] = r[col["Kernel Name"]]
            v = float(r[col["Metric Value"]].replace(",", ""))
            per[kid][r[col["Metric Name"]]] = v
            per[kid]["unit:" + r[col["Metric Name"]]] = r[col["Metric Unit"]]
        except (ValueError, KeyError):
            pass
    cls = defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0, tensor_w=0.0))
    tot = 0.0
    for kid, m in per.items():
        name = re.sub(r"\(.*", "", names[kid])
        name = re.sub(r"^void ", "", name).replace("lion::", "")
        dur = m.get("gpu__time_duration.sum", 0.0)
        u = m.get("unit:gpu__time_duration.sum", "ns")
        us = dur / 1000.0 if u in ("ns", "nsecond") else (dur if u in ("us", "usecond") else dur * 1000.0)
        def bytes_of(k):
            v = m.get(k, 0.0)
            u = str(m.get("unit:" + k, "byte")).lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        c = cls[name]
        c["n"] += 1; c["us"] += us; c["rd"] += bytes_of("dram__bytes_read.sum"); c["wr"] += bytes_of("dram__bytes_write.sum")
        c["tensor_w"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
        tot += us
    print("one global-prior + one PVCNN2Prior forward, B=32: %d kernels, %.1f us summed (ncu: serialised, cold cache; compare SHARES)" % (len(per), tot))
    print("%-44s %5s %10s %7s %9s %9s %9s %8s" % ("kernel", "n", "total us", "share", "avg us", "DRAM MB", "HBM GB/s", "tensor%"))
    for name, c in sorted(cls.items(), key=lambda kv: -kv[1]["us"]):
        mb = (c["rd"] + c["wr"]) / 1e6
        gbs = (c["rd"] + c["wr"]) / 1e9 / (c["us"] * 1e-6) if c["us"] > 0 else 0.0
        print("%-44s %5d %10.1f %6.1f%% %9.1f %9.1f %9.0f %8.1f" % (name[:44], c["n"], c["us"], 100 * c["us"] / tot, c["us"] / c["n"], mb, gbs,
                                                                 c["tensor_w"] / c["us"] if c["us"] else 0.0))


if __name__ == "__main__":
    main(sys.argv[1])
