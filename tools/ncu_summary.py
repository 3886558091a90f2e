#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture of the fused kernel for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_prof_fused_final_summary.txt [--json profiles/fused_ncu.json]

Writes the headline raw metrics, the per-pipeline-stage stall totals (the kernel's warp roles are its hot
loops, found by their execution counts) and the most-sampled instructions.  With --json it also refreshes the
numbers bench.py quotes in `roofline.traffic`.
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "smsp__sass_inst_executed_op_tmem_stt.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_not_selected", "stall_selected",
          "stall_branch_resolving", "stall_dispatch", "stall_math", "stall_mio", "stall_no_inst"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = []
    raw = ncu(rep, "raw")
    head, units, row = raw[0], raw[1], raw[2]
    vals = dict(zip(head, row))
    unit = dict(zip(head, units))
    lines.append("kernel: %s   (ncu --set full --clock-control none --import-source on; one launch = both views of one pair)"
                 % vals.get("Kernel Name", "?"))
    for k in head:
        if k in KEYS or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")) or "TriageCompute.l1tex__data_pipe" in k \
                or "pipe_tensor_cycles_active" in k:
            lines.append("%-92s %-16s %s" % (k, unit.get(k, ""), vals[k]))

    src = ncu(rep, "source")
    h, rows = src[1], src[2:]
    ix = {k: i for i, k in enumerate(h)}
    base = int(rows[0][0], 16)
    total = sum(int(x[ix["# Samples"]]) for x in rows)
    lines.append("")
    lines.append("instructions in the kernel: %d, warp-state samples: %d" % (len(rows), total))
    # hot loops = runs of instructions executed about once per pipeline iteration
    import collections
    exs = [int(x[ix["Instructions Executed"]]) for x in rows]
    mode_ex = max(collections.Counter(exs).items(), key=lambda kv: kv[0] * kv[1])[0]  # the row-loop trip count
    segs, cur = [], None
    for x in rows:
        off = int(x[0], 16) - base
        if 0.75 * mode_ex <= int(x[ix["Instructions Executed"]]) <= 1.1 * mode_ex:
            if cur is None:
                cur = [off, off, []]
            cur[1] = off
            cur[2].append(x)
        elif cur is not None and off - cur[1] > 0x200:
            segs.append(cur)
            cur = None
    if cur:
        segs.append(cur)
    lines.append("hot loops (one per warp role; `selected` ~ instructions issued, the rest is where the role's warps wait):")
    for lo, hi, sel in segs:
        if len(sel) < 40:
            continue
        ops = " ".join(x[1] for x in sel)
        # warp roles by what only they execute: FHFMA = first-stage accumulation, FSETP without shuffles = the merge
        # warps, half->float conversions with shuffles = q, LDTM + STTM with FFMA = coefficients and rings
        if "UTCHMMA" in ops:
            role = "MMA issue warp"
        elif "UBLKCP" in ops:
            role = "TMA producer warp"
        elif "HMNMX2" in ops and "UTCHMMA" in vals.get("Kernel Name", "") + "".join(x[1] for x in rows[:0]) or ("HMNMX2" in ops and "k_fused_mma" in vals.get("Kernel Name", "")):
            role = "role A (cost, exact fp16 pieces)"
        elif "FHADD" in ops and "k_fused_mma" in vals.get("Kernel Name", ""):
            role = "role B (box sums, a, b, hi/lo split)"
        elif "FSETP" in ops and "k_fused_mma" in vals.get("Kernel Name", ""):
            role = "role C (q, winner-take-all)"
        elif "FHFMA" in ops:
            role = "stage 0 (cost, first box filter)"
        elif "FSETP" in ops and "SHFL" not in ops:
            role = "stage 3 (merge warps)"
        elif "FSETP" in ops:
            role = "stage 2 (q, merge)"
        elif "HADD2.F32" in ops or "STS.128" in ops and "LDTM" in ops and "STTM" not in ops:
            role = "stage 2 (second box filter, q)"
        else:
            role = "stage 1 (a, b, rings)"
        lines.append("  [0x%04x, 0x%04x] %-34s instr %4d  samples %6d  %s" % (
            lo, hi, role, len(sel), sum(int(x[ix["# Samples"]]) for x in sel),
            " ".join("%s:%d" % (c[6:], sum(int(x[ix[c]]) for x in sel)) for c in STALLS)))
    lines.append("")
    lines.append("most-sampled instructions (offset, samples, share, executions, SASS, dominant stalls):")
    for x in sorted(rows, key=lambda x: -int(x[ix["# Samples"]]))[:16]:
        n = int(x[ix["# Samples"]])
        lines.append("  %5x %7d %5.1f%% ex=%9s  %-58s %s" % (
            int(x[0], 16) - base, n, 100.0 * n / total, x[ix["Instructions Executed"]], x[1].strip()[:58],
            " ".join("%s:%s" % (c[6:], x[ix[c]]) for c in STALLS if int(x[ix[c]]) > 0.2 * n)))
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))

    if "--json" in sys.argv:
        jp = sys.argv[sys.argv.index("--json") + 1]
        def scaled(k):  # ncu prints bytes in the unit of the second header row
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[k]]
            return float(vals[k]) * mult
        rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
        t = float(vals["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit["gpu__time_duration.sum"]]
        json.dump({
            "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
            "src": "ncu --set full --clock-control none, %s (1080p D=256, one launch = both views of one pair)" % dst,
            "l1tex_data_pipe_pct": float(vals["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
            "issue_active_pct": float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            "fma_pipe_pct": float(vals["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
            "alu_pipe_pct": float(vals["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]),
            "tensor_pipe_pct": next((float(vals[k]) for k in head if "pipe_tensor_cycles_active_realtime.avg.pct" in k), None),
            "smem_pipe_pct": {"lsu": float(vals.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "nan")),
                              "tensor_operands": float(vals.get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "nan"))},
            "kernel": vals.get("Kernel Name", "?"),
            "kernel_ms_under_ncu": t, "registers": float(vals["launch__registers_per_thread"]),
        }, open(jp, "w"), indent=1)


if __name__ == "__main__":
    main()
