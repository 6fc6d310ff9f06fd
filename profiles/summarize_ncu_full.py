#!/usr/bin/env python
"""Markdown table of the headline metrics of an `ncu --set full` report: python summarize_ncu_full.py report.ncu-rep
(also prints the mean DRAM bytes per launch, the `traffic` figure of bench.py's roofline object)."""
import csv
import io
import subprocess
import sys

COLS = ["Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    print("| " + " | ".join(hdr[i] for i in idx) + " |")
    print("|" + "---|" * len(idx))
    print("| " + " | ".join(units[i] for i in idx) + " |")
    for r in data:
        print("| " + " | ".join(r[i] for i in idx) + " |")
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot = sum(float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]] for r in data)
    print("\nmean DRAM bytes per launch (read + write): %.0f over %d launches" % (tot / len(data), len(data)))


if __name__ == "__main__":
    main(sys.argv[1])
