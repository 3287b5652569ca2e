"""Summarise one kernel of an .ncu-rep (ncu --set full) into a small JSON for profiles/:
python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_uniform", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tc", "sm__pipe_tc", "tensor")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    res = []
    for v in vals:
        d = {}
        for h, u, x in zip(hdr, units, v):
            if any(k in h for k in KEEP):
                d[h] = [x, u]
        res.append(d)
    json.dump(res if len(res) > 1 else res[0], sys.stdout, indent=1)


if __name__ == "__main__":
    main()
