#!/usr/bin/env python
"""Summarise ncu outputs into small text files that can be committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_<kernel>.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            v_us = v / 1000.0 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1000.0
            rows.append((short(r["Kernel Name"]), v_us, r["Grid Size"]))
    agg = OrderedDict()
    for k, us, _ in rows:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    total = sum(t for _, t in agg.values())
    print(f"# ncu launch list: {len(rows)} launches, {total:.1f} us of kernel time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / total:.1f}% |")
    print("\n## ten longest launches\n\n| kernel | grid | us |\n|---|---:|---:|")
    for k, us, g in sorted(rows, key=lambda r: -r[1])[:10]:
        print(f"| `{k}` | {g} | {us:.1f} |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full capture: {path}\n")
    for r in rows[2:]:
        print(f"## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        print()


def _num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return float("nan")


def _scaled(v, unit, table):
    return _num(v) * table.get(unit, 1.0)


def variants(path, order_path):
    """One row per captured launch of tools/ncu_targets.py: label and algorithmic GB/s from the order file, the rest
    from the report."""
    import json
    order = json.load(open(order_path))
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in hdr}
    t_unit = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}
    b_unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print(f"# ncu --set full, one launch per kernel variant ({path})\n")
    print("`alg GB/s` = algorithmic bytes / gpu__time_duration of the profiled (cold, serialised) launch; `traffic/alg` = "
          "(dram read + write) / algorithmic bytes.\n")
    print("`instr/elem` = 32 x smsp__inst_executed.sum / elements (thread-level instructions per tensor element); `smem conflicts` = "
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum per 1000 elements.\n")
    print("| # | variant | kernel | us | alg GB/s | dram R GB | dram W GB | traffic/alg | dram % peak | issue active % | warps active % | regs | instr/elem | smem conflicts |")
    print("|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    data = rows[2:]
    for i, r in enumerate(data):
        lab = order[i]["label"] if i < len(order) else "?"
        alg = order[i]["algorithmic_bytes"] if i < len(order) else float("nan")
        ne = order[i].get("elements", float("nan")) if i < len(order) else float("nan")
        us = _scaled(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]], t_unit)
        rd = _scaled(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]], b_unit)
        wr = _scaled(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]], b_unit)
        g = lambda k: r[col[k]] if k in col else ""  # noqa: E731
        print(f"| {i} | {lab} | `{short(r[col['Kernel Name']])[:60]}` | {us:.1f} | {alg / us / 1e3:.0f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | "
              f"{(rd + wr) / alg:.3f} | {_num(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
              f"{_num(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
              f"{_num(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {g('launch__registers_per_thread')} | "
              f"{32 * _num(g('smsp__inst_executed.sum')) / ne:.1f} | "
              f"{1000 * _num(g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')) / ne:.1f} |")


def traffic(path):
    """Launch list with dram bytes (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum):
    per-kernel totals as JSON (used for bench.py's roofline.traffic)."""
    import json
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        key = (r["ID"], short(r["Kernel Name"]))
        d = per.setdefault(key, {})
        v = _num(r["Metric Value"])
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            d["us"] = v * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3}.get(unit, 1.0)
        elif name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d[name] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    agg = OrderedDict()
    for (_, k), d in per.items():
        a = agg.setdefault(k, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1
        a["us"] += d.get("us", 0.0)
        a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    for a in agg.values():
        a["dram_bytes_per_launch"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / max(a["launches"], 1)
    print(json.dumps(agg, indent=1))


if __name__ == "__main__":
    {"launches": launches, "report": report, "variants": variants, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
