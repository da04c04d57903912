#!/usr/bin/env python
"""profiles/<tag>_ncu_linear_tc.md from an `ncu --set full -k regex:linear_tc_kernel` capture of `tools/linear_tc_probe.py ncu`."""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration us"), ("launch__grid_size", "CTAs"), ("launch__registers_per_thread", "regs"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory pipe active %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->TC wavefronts % of peak"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp insts")]
SHAPES = ["NT fwd layer 0: 16157 x 602 -> 512 (ld 608)", "NT fwd layer 1: 8689 x 1024 -> 512", "NT dX layer 1: 8689 x 512 -> 1024",
          "TN dW layer 0: M 16157, 512 x 602", "TN dW layer 1: M 8689, 512 x 1024"]


def main():
    tag, rep = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full, round {tag}: `tc::linear_tc_kernel` (tcgen05 3xTF32 linears), one launch per production shape",
           "", "`ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -c 5 python tools/linear_tc_probe.py ncu`",
           "(cold, serialised replays: read ratios, not absolute times; timed numbers are in profiles/r2_linear_tc.md).", "",
           "| metric | " + " | ".join(SHAPES[:len(body)]) + " |", "|---|" + "---:|" * len(body)]
    for key, label in KEYS:
        if key not in idx:
            continue
        vals = []
        for r in body:
            v, u = r[idx[key]], units[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.1f}" if abs(f) < 1e4 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip() if u not in ("%", "", "inst", "register/thread") else v)
        out.append(f"| {label} | " + " | ".join(vals) + " |")
    print("\n".join(out))


if __name__ == "__main__":
    main()
