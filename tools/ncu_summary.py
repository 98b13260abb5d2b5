#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few metrics the roofline argument uses.

    python tools/ncu_summary.py gpurun_out/prefill_full.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_active.avg",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"   # per-reason warp stall (warp-state section)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {name[:110]}")
        print(f"   grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]} id={r[hdr.index('ID')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k} [{units[i]}] = {r[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL_PREFIX) and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h[len(STALL_PREFIX):-len("_per_warp_active.pct")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("   top warp stall reasons (% of active warps): " + ", ".join(f"{n}={v:.1f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
