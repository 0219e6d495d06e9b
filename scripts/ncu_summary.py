#!/usr/bin/env python3
"""Condense an .ncu-rep (via `ncu -i X --page raw --csv`) into one line per profiled launch with the metrics the
roofline discussion needs."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma_%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1_gld_MB", 1e-6),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    for r in data:
        parts = [r[name_i][:60]]
        for k, label, scale in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                try:
                    v = float(r[idx[k]].replace(",", ""))
                    u = units[idx[k]]
                    if label.endswith("_us") and u in ("ns", "nsecond"):
                        v = v * 1e-3
                    elif label.endswith("_us") and u in ("us", "usecond"):
                        pass
                    elif label.endswith("_us") and u in ("ms", "msecond"):
                        v = v * 1e3
                    elif label.endswith("_MB"):
                        mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, scale)
                        v = v * mult
                    parts.append("%s=%.4g" % (label, v))
                except ValueError:
                    pass
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
