"""Summarise ncu output for profiles/: either a `--set full` report (key metrics of each kernel) or a
launch-list CSV (`--metrics gpu__time_duration.sum[,dram__bytes_*] --csv --log-file`): time share per kernel.
    python tools/ncu_summary.py report.ncu-rep > profiles/rNN/..._summary.txt
    python tools/ncu_summary.py launches.csv   > profiles/rNN/..._summary.txt"""
import csv
import os
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("Kernel Name".ljust(92), d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(k.ljust(92), units[hdr.index(k)].ljust(16), d[k])
        print()


def launches(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        m, v = r["Metric Name"], float(r["Metric Value"].replace(",", ""))
        e = per.setdefault(name, {"n": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
        e["n"].add(r["ID"])
        unit = r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            e["us"] += v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            e["rd" if "read" in m else "wr"] += v * scale
    tot = sum(e["us"] for e in per.values())
    print("%-44s %8s %12s %7s %12s %12s" % ("kernel", "launches", "total_us", "share", "dram_rd_MB", "dram_wr_MB"))
    for name, e in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        print("%-44s %8d %12.1f %6.1f%% %12.1f %12.1f" % (name[:44], len(e["n"]), e["us"], 100 * e["us"] / tot,
                                                         e["rd"] / 1e6, e["wr"] / 1e6))
    print("%-44s %8s %12.1f" % ("total", "", tot))


def pass_traffic(path, out_json):
    """One Unet pass = the conv launches between the last two p_sample launches of the list: serialized time and
    DRAM bytes, per launch and summed -> out_json (read by bench.py as roofline.traffic)."""
    import json
    text = open(path).read()
    rows = list(csv.DictReader(io.StringIO(text[text.index('"ID"'):])))
    L = OrderedDict()
    for r in rows:
        e = L.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})
        v, unit = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            e["rd" if "read" in r["Metric Name"] else "wr"] = v * scale
    ids = list(L)
    ps = [i for i, k in enumerate(ids) if "p_sample" in L[k]["name"]]
    a, b = ps[-2], ps[-1]
    convs = [L[ids[i]] for i in range(a + 1, b)]
    res = {"conv_launches_per_unet_pass": len(convs), "serialized_cold_us": sum(c["us"] for c in convs),
           "dram_bytes_read": sum(c.get("rd", 0.0) for c in convs), "dram_bytes_write": sum(c.get("wr", 0.0) for c in convs),
           "per_launch": [{"us": round(c["us"], 2), "rd_MB": round(c.get("rd", 0.0) / 1e6, 2),
                           "wr_MB": round(c.get("wr", 0.0) / 1e6, 2)} for c in convs],
           "source": os.path.join(os.path.basename(os.path.dirname(os.path.abspath(out_json))), os.path.basename(path)) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                     "dram__bytes_write.sum --clock-control none, --cache-control none, python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn)"}
    json.dump(res, open(out_json, "w"), indent=1)
    print("pass: %d launches, %.1f us serialized, %.2f GB read + %.2f GB written -> %s" % (
        res["conv_launches_per_unet_pass"], res["serialized_cold_us"], res["dram_bytes_read"] / 1e9,
        res["dram_bytes_write"] / 1e9, out_json))


if __name__ == "__main__":
    p = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "--pass-json":
        pass_traffic(p, sys.argv[3])
    else:
        (report if p.endswith(".ncu-rep") else launches)(p)
