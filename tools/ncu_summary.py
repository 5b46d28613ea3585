"""Turn gpurun_out/*.csv / *.ncu-rep into small tracked text summaries under profiles/.
usage: python tools/ncu_summary.py launches <launches.csv> <out.txt>
       python tools/ncu_summary.py full <file.ncu-rep> <out.txt>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu", "smsp__warp_issue_stalled", "sm__inst_executed_pipe_tc"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        agg.setdefault(r[kn].split("(")[0][-60:], []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s\n" % src)
        for k, v in agg.items():
            f.write("%-62s launches=%4d  avg=%10.1f ns  share=%5.1f%%\n" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none; source: %s\n" % src)
        for r in rows[2:]:
            f.write("== %s\n" % r[hdr.index("Kernel Name")][:90])
            for i, h in enumerate(hdr):
                if any(h.startswith(k) for k in KEYS) and r[i] not in ("", "n/a"):
                    f.write("  %-88s %s %s\n" % (h, r[i], units[i]))
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
