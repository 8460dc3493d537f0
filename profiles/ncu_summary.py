"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics, top stall lines, per-source-line instruction share.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_issued.min", "smsp__inst_issued.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "launch__shared_mem_per_block_dynamic"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2 + which]))
    print("kernel:", d["Kernel Name"][:100])
    for k in KEYS:
        if k in d:
            print("  %-62s %s %s" % (k, d[k], units[hdr.index(k)]))
    stalls = []
    for h in hdr:
        if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
            try:
                stalls.append((float(d[h]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1
    print("  stall samples:", ", ".join("%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:8]))
    # per source line
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    cur = None
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0].isdigit() and len(r) > 8 and cur:
            key = (cur, int(r[0]))
            try:
                inst[key] += int(r[7]); samp[key] += int(r[6]); text[key] = r[1]
            except ValueError:
                pass
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print("  top source lines by stall samples (all profiled launches):")
    for k, v in samp.most_common(14):
        print("   %5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * v / ts, 100 * inst[k] / ti, k[0], k[1], text[k].strip()[:90]))
    print("  top source lines by instructions:")
    for k, v in inst.most_common(14):
        print("   %5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100 * v / ti, 100 * samp[k] / ts, k[0], k[1], text[k].strip()[:90]))


if __name__ == "__main__":
    main()
