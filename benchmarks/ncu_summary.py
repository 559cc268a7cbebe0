"""Turns an `.ncu-rep` (read here, on the CPU box, with `ncu -i ... --page raw --csv`) into the small
JSON summaries committed under profiles/: one object per captured launch with the metrics the
roofline discussion in DESIGN.md section 4 uses.

    python benchmarks/ncu_summary.py gpurun_out/r2_lut.ncu-rep > profiles/r2_ncu_....json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = (
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
)


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"Kernel Name": r[header.index("Kernel Name")]}
        for k in header:
            if k in KEEP or k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("_not_issued"):
                i = header.index(k)
                if r[i] not in ("", "0"):
                    d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    json.dump(res if len(res) != 1 else res[0], sys.stdout, indent=1)


if __name__ == "__main__":
    main()
