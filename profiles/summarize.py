"""Turn the ncu artefacts gpurun brought back (gpurun_out/, scratch) into the small tracked summaries
under profiles/.   python profiles/summarize.py r01
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("tmg::", "")


def launches(tag):
    path = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        t = float(r["Metric Value"].replace(",", "")) / 1e3   # us
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += t
        total += t
    with open(os.path.join(ROOT, "profiles", tag + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)\n")
        f.write("# %d launches captured, %.1f us total\n" % (sum(a[0] for a in agg.values()), total))
        f.write("%-70s %8s %12s %8s\n" % ("kernel", "launches", "total_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-70s %8d %12.1f %7.1f%%\n" % (k[:70], n, t, 100 * t / total))
    print(open(os.path.join(ROOT, "profiles", tag + "_launches_summary.txt")).read())


def rep(tag, which):
    path = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, which))
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (tag, which)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; selected raw metrics per captured launch\n")
        for d in data:
            f.write("\n== %s  grid %s block %s\n" % (short(d[idx["Kernel Name"]])[:90], d[idx["Grid Size"]], d[idx["Block Size"]]))
            for k in KEYS:
                if k in idx:
                    f.write("   %-75s %-12s %s\n" % (k, units[idx[k]], d[idx[k]]))
    print(open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (tag, which))).read()[:6000])
    # dram traffic per launch of the captured kernel, for bench.py's roofline.traffic
    if data and "dram__bytes_read.sum" in idx:
        def num(d, k):
            v, u = float(d[idx[k]].replace(",", "")), units[idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return num(data[0], "dram__bytes_read.sum") + num(data[0], "dram__bytes_write.sum")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    import json
    traffic = {}
    for which in sys.argv[2:] or ["conv", "pointwise"]:
        t = rep(tag, which)
        if t is not None:
            traffic[{"step": "flow_step_fused", "gate": "conv_lstm_gates"}.get(which, which)] = t
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        print("traffic.json:", traffic)
