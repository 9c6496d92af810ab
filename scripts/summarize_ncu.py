"""Turns ncu output into the markdown summaries kept under profiles/.

  launches: python scripts/summarize_ncu.py launches <csv from `ncu --metrics gpu__time_duration.sum --csv --log-file`> "<command line>"
  full    : python scripts/summarize_ncu.py full <file.ncu-rep | raw.csv from `ncu -i rep --page raw --csv`> [kernel-name-substring ...]
  traffic : python scripts/summarize_ncu.py traffic <raw.csv> <out.json> "<source note>"
            per kernel (by short name) the dram bytes read / written and the duration of its LARGEST launch: the
            file bench.py reads its roofline.traffic from (profiles/kernel_traffic.json)
"""
import json
import re
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads",
        "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


def launches(path, cmd):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        d[r[ik]].append(float(r[iv].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in d.values())
    print(f"# command: {cmd}")
    print(f"# total GPU time of all launches: {total:.1f} us\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:100]}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {100 * sum(v) / total:.1f}% |")


def raw_rows(path):
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(l for l in out.splitlines() if l.startswith('"')))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def traffic(path, out_path, note):
    rows = raw_rows(path)
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    ir, iw, it = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    best = {}
    for r in rows[2:]:
        m = re.search(r"(\w+?)(<|\(|$)", r[ik].replace("void ", "").replace("vqseg::", ""))
        name = m.group(1) if m else r[ik]
        rec = {"kernel": r[ik][:120], "dram_bytes_read": to_bytes(r[ir], units[ir]), "dram_bytes_write": to_bytes(r[iw], units[iw]),
               "duration_us": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1e-3),
               "grid": r[h.index("launch__grid_size")] if "launch__grid_size" in h else None}
        if name not in best or rec["dram_bytes_read"] + rec["dram_bytes_write"] > best[name]["dram_bytes_read"] + best[name]["dram_bytes_write"]:
            best[name] = rec
    json.dump({"source": note, "kernels": best}, open(out_path, "w"), indent=1)
    print("wrote", out_path, "with", len(best), "kernels")


def full(path, names):
    rows = raw_rows(path)
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    for r in rows[2:]:
        if names and not any(n in r[ik] for n in names):
            continue
        print(f"### {r[ik][:110]}")
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"  {w:84s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu --set full")
    else:
        full(sys.argv[2], sys.argv[3:])
