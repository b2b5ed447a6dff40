"""Turn the raw ncu artefacts brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r01b.csv > profiles/r01_launches_summary.txt
    python profiles/summarize.py kernel gpurun_out/r01_flash_ws.ncu-rep > profiles/r01_flash_ws.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# per-launch device time, summed by kernel ({sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms); cold-cache and")
    print("# serialised under ncu: compare SHARES with bench.py's live CUDA-event share, not absolutes")
    print(f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {n:8d} {t / 1e3:10.3f} {t / n:9.1f} {100 * t / tot:6.1f}%")


def kernel(rep, top=18):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr = r[0]
    print(f"# {rep}: ncu --set full --clock-control none --import-source on (one launch)")
    print("kernel:", r[2][hdr.index("Kernel Name")] if len(r) > 2 else "?")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"{m:72s} {r[2][i]:>16s} {r[1][i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return
    h = rows[1]
    ci = {n: i for i, n in enumerate(h)}
    data = []
    for x in rows[2:]:  # a report with several launches repeats the header: keep the first launch only
        if len(x) != len(h):
            continue
        if x == h:
            break
        data.append(x)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(x[ci["# Samples"]] or 0) for x in data)
    agg = {s: sum(int(x[ci[s]] or 0) for x in data) for s in stalls}
    print(f"\n# warp-state samples: {tot}")
    print("  " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print(f"\n# top {top} SASS instructions by samples")
    for x in sorted(data, key=lambda x: -int(x[ci["# Samples"]] or 0))[:top]:
        st = {s[6:]: int(x[ci[s]] or 0) for s in stalls if int(x[ci[s]] or 0)}
        print(f"{x[ci['# Samples']]:>7s}  {x[ci['Source']][:80]:80s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])}")
    ops = collections.Counter()
    for x in data:
        mn = x[ci["Source"]].split()
        if mn:
            op = mn[1] if mn[0].startswith("@") and len(mn) > 1 else mn[0]
            if re.match(r"UTC|LDTM|STTM|UTMA|UBLKCP|SYNCS|LDGSTS|HMMA|FFMA|MUFU", op):
                ops[op.split(".")[0]] += 1
    print("\n# SASS mnemonics present (static count):", dict(ops))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else kernel)(sys.argv[2])
