"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> compact per-launch table of the metrics the roofline needs.
Run ON the GPU box right after a capture so that only the small summary travels back (gpurun_out/ is capped at 64 MiB)."""
import csv
import sys

KEYS = [
    ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
]


GLUE_KEYS = [
    ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "us"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
    ("smsp__inst_executed.sum", "inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
    ("sm__inst_executed_pipe_tensor.sum", "inst_tensor"),
]
if "--glue" in sys.argv:
    sys.argv.remove("--glue")
    KEYS = GLUE_KEYS


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("kernel | " + " | ".join(lbl for _, lbl in KEYS) + "\n")
tot_dram, n = 0.0, 0
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    vals = []
    for k, lbl in KEYS:
        if k not in col:
            vals.append("-")
            continue
        v, u = r[col[k]], units[col[k]]
        if "bytes" in k:
            b = to_bytes(v, u)
            vals.append(f"{b / 1e6:.2f}MB")
            if lbl in ("dram_rd", "dram_wr"):
                tot_dram += b
        else:
            vals.append(v if u != "ns" else f"{float(v.replace(',', '')) / 1e3:.2f}")
    n += 1
    out.write(r[col["Kernel Name"]][:46] + " | " + " | ".join(vals) + "\n")
out.write(f"# {n} launches, mean DRAM bytes (read+write) per launch: {tot_dram / max(n, 1):.0f}\n")
