#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/gemm_r1.ncu-rep profiles/gemm_r1_summary.csv"""
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [i for i, h in enumerate(hdr) if h in KEEP]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch_%d" % i for i in range(len(data))])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    print("wrote", out, "launches", len(data))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
