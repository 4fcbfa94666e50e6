"""Summarise an ncu report (--set full) into the handful of numbers DESIGN.md / profiles/ cite.

    python scripts/summarize_ncu.py gpurun_out/conv_umma_full.ncu-rep > profiles/rNN_conv_umma_full.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "HMMA subpipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.sum.per_cycle_elapsed", "warp instructions / cycle (all SMs)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__cluster_dim_x", "cluster x"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall: long scoreboard %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall: barrier %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall: math pipe throttle %"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    print(f"# {path}: {len(rows) - hdr - 2} kernel instance(s); ncu --set full --clock-control none (cold-cache, serialised replays)")
    for r in rows[hdr + 2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        kn = d.get("Kernel Name", "?")
        print(f"\n## {kn[:150]}")
        for key, label in KEYS:
            hits = [n for n in names if n.endswith(key)]
            if hits and d.get(hits[0], "") != "":
                print(f"  {label:48s} {d[hits[0]]:>16s} {u[hits[0]]}")


if __name__ == "__main__":
    main(sys.argv[1])
