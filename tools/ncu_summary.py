#!/usr/bin/env python
"""Curated per-kernel summary of an .ncu-rep (read with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [-o profiles/name.md]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (lts) throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("lts__t_sectors_op_red.sum", "L2 sectors reduced (red)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput % (max of sub-units)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "  .. shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "  .. of which LDS"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  .. bank-conflict wavefronts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests (warp instr)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "global red sectors"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR request cycles active %"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "L1->XBAR write bytes (stores+reds)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "XBAR->L1 read bytes (fills)"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "LSU writeback active %"),
    ("l1tex__t_set_conflicts_pipe_lsu.sum", "L1 tag set conflicts"),
    ("l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed", "LSU-in requests %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe instr %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    dst = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    hdr, units, kernels = load(rep)
    lines = [f"# ncu summary of `{rep}`", "",
             "`ncu --set full --clock-control none --import-source on`; per-launch values; cold-ish caches (ncu replays each kernel ~40x).", ""]
    for r in kernels:
        d = dict(zip(hdr, r))
        lines.append(f"## {d.get('Kernel Name', '?')}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        for k, label in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append(f"| {label} (`{k}`) | {d[k]} | {units[hdr.index(k)]} |")
        lines.append("")
    text = "\n".join(lines)
    if dst:
        open(dst, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
