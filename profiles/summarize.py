#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed_pipe_xu.sum", "sm__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)  total {tot:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.3f} ms {v[0]:5d}x  {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0]:8.4f} ms  {k[:120]}")


def raw_page(path):
    """The `--page raw --csv` text of a capture: from the .ncu-rep, or a .csv already exported on the GPU box (reports with
    --import-source run to 100+ MB, more than gpurun brings back)."""
    if path.endswith(".csv"):
        return open(path).read()
    return subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)


def full(path):
    raw = raw_page(path)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP or h == "Kernel Name"]
    for r in rows[2:]:
        print("----")
        for i in idx:
            print(f"  {hdr[i]:85s} {r[i][:100]} {units[i]}")


def traffic(*paths):
    """dram bytes (read + write) per launch for each trunk kernel class of one or more full captures -> JSON on stdout
    (committed as profiles/ncu_traffic.json and read by bench.py's roofline.traffic).  `conv`: the mean over every captured
    conv_tcgen05_kernel launch -- capture ALL conv launches of one tokenizer pass (encode + decode of one 32-image chunk) so that
    the mean matches the mean launch bench.py --workload tokenizer times."""
    import json
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kinds = {"gemm2_bf16_tcgen05_kernel<5>": "gemm_qkv", "gemm2_bf16_tcgen05_kernel<6>": "gemm_up", "attention_tc_kernel": "attention",
             "conv_tcgen05_kernel": "conv"}
    acc = collections.defaultdict(list)
    seen7 = 0
    allrows = []
    for path in paths:
        raw = raw_page(path)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            allrows.append((r, ki, ri, wi, units))
    for r, ki, ri, wi, units in allrows:
        name = re.sub(r"\(int\)", "", r[ki])
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        kind = next((v for k, v in kinds.items() if k in name), None)
        if kind is None and "gemm2_bf16_tcgen05_kernel<7>" in name:   # out-projection and MLP-down alternate within a layer
            kind = "gemm_out" if seen7 % 2 == 0 else "gemm_down"
            seen7 += 1
        if kind:
            acc[kind].append(b)
    out = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in acc.items()}
    out["_source"] = f"ncu --set full --clock-control none captures {', '.join(paths)} (profiles/run_profile_r02.sh); trunk: B=256 -> 512 sequences per forward; conv: one 32-image tokenizer pass"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
