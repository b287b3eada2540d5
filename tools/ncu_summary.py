#!/usr/bin/env python
"""Markdown bullets of the metrics DESIGN.md / profiles/ quote, from an .ncu-rep brought back in gpurun_out/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra_metric ...]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_ltcfabric.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second"]


def main():
    rep, extra = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for w in WANT + extra:
            for i, h in enumerate(hdr):
                if h == w:
                    print(f"* `{w}` = {r[i]} {units[i]}".rstrip())
        print()


if __name__ == "__main__":
    main()
