"""Summarise ncu reports under gpurun_out/ into the text files kept under profiles/ (run in the build container):
    python tools/ncu_summary.py full gpurun_out/prof_pw1dw.ncu-rep "title"     -> key metrics of one --set full capture
    python tools/ncu_summary.py list gpurun_out/fwd_launches.csv "title"       -> per-kernel share of a launch list
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(title)
    print("Kernel Name |  |", vals[ix["Kernel Name"]])
    rd = wr = None
    for k in KEYS:
        if k in ix:
            print(f"{k} | {units[ix[k]]} | {vals[ix[k]]}")
            if k == "dram__bytes_read.sum":
                rd = (float(vals[ix[k]].replace(",", "")), units[ix[k]])
            if k == "dram__bytes_write.sum":
                wr = (float(vals[ix[k]].replace(",", "")), units[ix[k]])
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd and wr:
        tot = rd[0] * mult[rd[1]] + wr[0] * mult[wr[1]]
        print(f"dram traffic per launch = {tot / 1e6:.1f} MB  ({tot:.0f} B)")


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    acc, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        u = r[ix["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("sty::", "")
        name = re.sub(r"\(int\)", "", name)
        acc[name] += ms
        cnt[name] += 1
    tot = sum(acc.values())
    print(title)
    print(f"launches {sum(cnt.values())}, total {tot:.1f} ms (cold-cache, serialised: shares, not absolutes)")
    for k, v in acc.most_common(40):
        print(f"{100 * v / tot:6.2f}%  n={cnt[k]:5d}  {v:10.3f} ms  {k}")


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
