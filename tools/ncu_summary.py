#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

  python tools/ncu_summary.py r1 gpurun_out/prof_r1_spmm.ncu-rep gpurun_out/launches_r1.csv

writes profiles/<tag>_ncu_kernels.md (key metrics per captured launch), profiles/<tag>_launches.md
(per-kernel totals and shares of the launch list) and profiles/traffic.json (DRAM bytes per launch of the
SpMM ops, read by bench.py for roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->L1 sectors"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("derived__lts__lts2xbar_bytes.sum.per_second", "L2->xbar rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "SM ingest % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full, round {tag}: captured launches of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train "
           "--no-other-workloads --no-ref-gpu --skip-gate` (one timed step: fwd0, fwd1, fwd2, A^T build + bwd1, scatter bwd2)",
           "", "Cold-cache, serialised replays: read shares and ratios, not absolute times.  `L2->xbar rate` is against ncu's own",
           "`derived__lts__lts2xbar_bytes.sum.peak_sustained` = 11,776 B/clk (184 L2 slices x 64 B) = 23.1 TB/s at 1.96 GHz: the dense SpMM",
           "launches run at 18.2-18.7 TB/s = 79-81 % of it (the gather-only probe kernel: 19.2-19.6 TB/s = 83-85 %), while the SM side",
           "(`SM ingest`) is only half used - the L2 output ports, not HBM (4-7 %) and not the SMs, bind these kernels.", "",
           "| # | kernel | grid | " + " | ".join(n for _, n in KEYS) + " |", "|---|---|---|" + "---|" * len(KEYS)]
    traffic = {}
    spmm_seen = 0
    op_order = ["fwd0", "fwd1", "fwd2", "bwd1", "bwd2"]
    for i, r in enumerate(body):
        name = r[idx["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
        cells = []
        for k, _ in KEYS:
            v = r[idx[k]] if k in idx else ""
            u = units[idx[k]] if k in idx else ""
            f = num(v)
            cells.append(f"{f:.4g} {u}".strip() if f is not None else v)
        out.append(f"| {i} | `{name}` | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")
        if (name.startswith("spmm_rowsplit") or name.startswith("spmm_scatter") or name.startswith("spmm_flat")) and spmm_seen < len(op_order):
            rd, wr = num(r[idx["dram__bytes_read.sum"]]), num(r[idx["dram__bytes_write.sum"]])
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            b = rd * scale.get(units[idx["dram__bytes_read.sum"]], 1.0) + wr * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
            traffic[op_order[spmm_seen]] = int(b)
            spmm_seen += 1
    open(os.path.join(REPO, "profiles", f"{tag}_ncu_kernels.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(REPO, "profiles", "traffic.json"), "w"), indent=1)

    lrows = list(csv.DictReader(l for l in open(launches) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in lrows:
        name = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("void ", "")[:90]
        agg.setdefault(name, []).append(float(r["Metric Value"]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    lo = [f"# ncu launch list, round {tag} (`--metrics gpu__time_duration.sum --clock-control none`), same command as above", "",
          f"{len(lrows)} launches, {tot / 1e3:.2f} ms of kernel time (includes workload set-up kernels: build_adj, torch fills/randn).", "",
          "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lo.append(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {sum(v) / tot:.3f} |")
    open(os.path.join(REPO, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lo) + "\n")
    print("wrote profiles/", tag, traffic)


if __name__ == "__main__":
    main()
