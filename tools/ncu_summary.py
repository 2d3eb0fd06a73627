#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline report needs."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), "| id", d.get("ID"))
        for h, u in zip(hdr, units):
            if any(h.endswith(k) for k in KEYS):
                print(f"  {h} [{u}] = {d[h]}")


if __name__ == "__main__":
    main(sys.argv[1])
