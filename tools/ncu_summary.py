"""Summarise an .ncu-rep (read here, no GPU needed) into the counters BASELINE.json asks for:
SM / FP32-pipe issue utilisation, warp execution efficiency, L1/L2 hit rates, DRAM traffic, stalls.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (busiest pipe/issue)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / warp instruction"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "warp execution efficiency %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "FMA-heavy pipe %"),
    ("sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "FMA-lite pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU/convert) pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU pipe %"),
    ("sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "CBU (branch) pipe %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("sass__inst_executed_local_loads", "local-memory load instructions"),
    ("sass__inst_executed_local_stores", "local-memory store instructions"),
    ("sass__inst_executed_shared_loads", "shared-memory load instructions"),
    ("sass__inst_executed_global_loads", "global load instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__warps_active.avg.per_cycle_active", "active warps / cycle / SMSP"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    path = sys.argv[1]
    recs, units = load(path)
    lines = []
    for rec in recs:
        lines.append(f"### {rec.get('Kernel Name', '?')}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        lines.append("")
        lines.append("| counter | value | unit |")
        lines.append("|---|---|---|")
        for key, label in KEYS:
            if key in rec and rec[key] != "":
                lines.append(f"| {label} (`{key}`) | {rec[key]} | {units.get(key, '')} |")
        stalls = sorted(((float(v), k) for k, v in rec.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v),
                        reverse=True)
        lines.append("")
        lines.append("Top warp stall reasons (warps stalled per issue-active cycle): " +
                     ", ".join(f"{k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, k in stalls[:7]))
        lines.append("")
    text = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
