"""Turns ncu output into the summaries kept under profiles/ (run here, after gpurun brought the files back).

  python tools/ncu_summary.py launches gpurun_out/X_launches.csv profiles/X_launch_shares.md "<command profiled>"
      launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ...`)
      -> per-kernel time and share of one MSM step (the last complete k_hist .. k_sum_encode run) and of the whole list.
  python tools/ncu_summary.py full gpurun_out/X.ncu-rep profiles/X_ncu_summary.json "<capture command>"
      `ncu --set full` report -> the metrics DESIGN.md quotes, per captured kernel.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "launch__grid_size",
    "launch__block_size",
]


def short(name: str) -> str:
    name = name.split("(")[0].strip()
    for pre in ("void ", "bpg::"):
        if name.startswith(pre):
            name = name[len(pre):]
    return name.replace("bpg::", "")


def read_launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = list(csv.reader(rows))
    hdr = rd[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    out = []
    for r in rd[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v / 1e3)
        out.append((short(r[ki]), v))
    return out


def launches(src, dst, command):
    ls = read_launches(src)
    # the last complete MSM step: from the last k_hist that is followed by a k_sum_encode
    ends = [i for i, (n, _) in enumerate(ls) if n.startswith("k_sum_encode") or n.startswith("k_exchange_sum_encode")]
    step = []
    for e in reversed(ends):
        starts = [i for i in range(e) if ls[i][0] in ("k_hist", "k_rs_hist")]
        if starts:
            # all k_hist launches that belong to this step (a piecewise upload launches several)
            s = starts[-1]
            while s > 0 and ls[s - 1][0] in ("k_hist", "k_rs_hist"):
                s -= 1
            step = ls[s : e + 1]
            break
    lines = [f"# ncu launch list: `{command}`", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised, so compare SHARES.",
             f"Raw list: `{src.replace('gpurun_out/', 'profiles/')}` ({len(ls)} launches).", ""]
    if step:
        tot = sum(v for _, v in step)
        lines += ["## One MSM step (last complete k_hist .. k_sum_encode run)", "", "| kernel | us | share of step |", "|---|---|---|"]
        agg = {}
        order = []
        for n, v in step:
            if n not in agg:
                order.append(n)
                agg[n] = [0.0, 0]
            agg[n][0] += v
            agg[n][1] += 1
        for n in order:
            v, c = agg[n]
            lines.append(f"| {n}{' x' + str(c) if c > 1 else ''} | {v:.1f} | {100 * v / tot:.1f}% |")
        lines += [f"| **total** | {tot:.1f} | |", ""]
    tot = sum(v for _, v in ls)
    agg = {}
    for n, v in ls:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    lines += ["## Whole list", "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"| {n} | {c} | {v:.1f} | {100 * v / tot:.1f}% |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


def full(src, dst, command):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    ki = hdr.index("Kernel Name")
    kernels = []
    for r in rd[2:]:
        k = {"kernel": short(r[ki])}
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                k[m] = f"{r[i]} {units[i]}".strip()
        kernels.append(k)
    json.dump({"capture": command, "source_report": src + " (not committed)", "kernels": kernels}, open(dst, "w"), indent=1)
    for k in kernels:
        print(k["kernel"], k.get("gpu__time_duration.sum"), k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    command = sys.argv[4] if len(sys.argv) > 4 else ""
    (launches if mode == "launches" else full)(src, dst, command)
