"""Summarise an .ncu-rep (from `ncu --set full`) into a small text file for profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]:>18s} {units[i]}")
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h.split('warp_issue_stalled_')[1].split('_per_warp')[0]:30s} {v:8.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])
