#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_full.csv   > profiles/rNN_launches.txt
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep        > profiles/rNN_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for row in r:
        if row[mi] != "gpu__time_duration.sum":       # the same pass may carry the DRAM byte counters too
            continue
        name = row[ki].split("(")[0][:64]
        v = float(row[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row[ui], 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"# total {tot / 1e6:.3f} ms over {sum(n for n, _ in agg.values())} launches")
    print(f"{'total ms':>10} {'n':>4} {'ms/launch':>10} {'share':>6}  kernel")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / 1e6:10.3f} {n:4d} {v / n / 1e6:10.3f} {100 * v / tot:5.1f}%  {k}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def full(path):
    if path.endswith(".csv"):      # `ncu -i rep --page raw --csv` already exported on the GPU box
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = csv.reader(out.splitlines())
    hdr = next(r)
    units = next(r)
    print("# ncu --set full --clock-control none --import-source on; per launch (first launch of each distinct kernel)")
    seen = set()
    for row in r:
        name = row[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        print("\n== " + name[:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:80s} {row[i]:>16s} {units[i]}")
        try:
            rd = float(row[hdr.index("dram__bytes_read.sum")])
            wr = float(row[hdr.index("dram__bytes_write.sum")])
            t = float(row[hdr.index("gpu__time_duration.sum")])
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            ts = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}
            b = rd * sc[units[hdr.index("dram__bytes_read.sum")]] + wr * sc[units[hdr.index("dram__bytes_write.sum")]]
            sec = t * ts[units[hdr.index("gpu__time_duration.sum")]]
            print(f"   {'=> DRAM traffic':80s} {b / 1e9:16.3f} GB   ({b / sec / 1e9:.0f} GB/s under the profiler)")
        except (ValueError, KeyError):
            pass


if __name__ == "__main__" and sys.argv[1] in ("launches", "full"):
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])


def traffic(path, out_json, **cfg):
    """ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum CSV -> per-kernel DRAM
    bytes per launch (max over the launches seen) as JSON for bench.py's roofline.traffic."""
    import json
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = {}
    for row in r:
        d = per.setdefault(row[ii], {"name": row[ki].split("(")[0]})
        v = float(row[vi].replace(",", ""))
        u = row[ui]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
        d[row[mi]] = v * scale
    out = {}
    for d in per.values():
        b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        k = out.setdefault(d["name"], {"dram_bytes_per_launch": 0.0, "ms": 0.0, "launches": 0})
        k["launches"] += 1
        if b > k["dram_bytes_per_launch"]:
            k["dram_bytes_per_launch"], k["ms"] = b, d.get("gpu__time_duration.sum", 0.0)
    json.dump({"how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                      "--clock-control none, largest launch per kernel", "config": cfg, "kernels": out},
              open(out_json, "w"), indent=1)
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"]):
        print(f"{v['dram_bytes_per_launch'] / 1e9:10.3f} GB {v['ms']:9.3f} ms  {k}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "traffic":
    traffic(sys.argv[2], sys.argv[3], nodes=2_000_000, slices=32, pairs=10_000_000, rho=0.9, band=10, feat=128,
            generator="global")
