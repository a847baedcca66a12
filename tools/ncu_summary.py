"""Summarise ncu outputs into small text files for profiles/ (the raw .ncu-rep / launch csv stay in gpurun_out/).
  python tools/ncu_summary.py rep <file.ncu-rep>            -> key metrics of each captured launch
  python tools/ncu_summary.py compact <file.ncu-rep>        -> one line per captured launch (time, tensor %, L2->SM, DRAM)
  python tools/ncu_summary.py list <launches.csv>           -> per-kernel time shares of the profiled command"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[name_i])
        for i, h in enumerate(hdr):
            if any(h == k or h.endswith("." + k) or k in h for k in KEYS):
                print("  %-95s %s %s" % (h, r[i], units[i]))


def compact(path):
    """One line per captured launch with the metrics the conv / GEMM / attention analysis uses."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    want = [("ms", "gpu__time_duration.sum"), ("sm_ghz", "sm__cycles_elapsed.avg.per_second"),
            ("tensor_pct", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            ("tensor_mem_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            ("l2_to_sm_GB", "l1tex__m_xbar2l1tex_read_bytes.sum"),
            ("l2_to_sm_pct", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed"),
            ("lts_pct", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
            ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
            ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size")]
    col = {}
    for key, name in want:
        for i, h in enumerate(hdr):
            if h == name or h.endswith("." + name):
                col[key] = i
                break
    name_i = hdr.index("Kernel Name")
    print("units: " + ", ".join("%s[%s]" % (k, units[i]) for k, i in col.items()))
    for r in rows[2:]:
        print("%-3s %-48s " % (r[0], r[name_i].split("(")[0][-48:]) + " ".join(
            "%s=%s" % (k, r[i][:8] if r[i] not in ("", "no data") else "-") for k, i in col.items()))


def launch_list(path):
    tot = defaultdict(float)
    cnt = defaultdict(int)
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3,
                 "second": 1e3}.get(unit, 1e-6)
        k = r["Kernel Name"].split("(")[0]
        tot[k] += v * scale
        cnt[k] += 1
    total = sum(tot.values())
    print("total device time of profiled launches: %.1f ms over %d launches" % (total, sum(cnt.values())))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-60s launches %6d  time %10.2f ms  share %6.2f%%  avg %9.4f ms" % (k[:60], cnt[k], v, 100 * v / total, v / cnt[k]))


if __name__ == "__main__":
    {"rep": rep, "list": launch_list, "compact": compact}[sys.argv[1]](sys.argv[2])
