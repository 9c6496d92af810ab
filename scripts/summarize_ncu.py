"""Turns ncu output into the markdown summaries kept under profiles/.

  launches: python scripts/summarize_ncu.py launches <csv from `ncu --metrics gpu__time_duration.sum --csv --log-file`> "<command line>"
  full    : python scripts/summarize_ncu.py full <file.ncu-rep> [kernel-name-substring ...]
"""
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


def full(path, names):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
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
    else:
        full(sys.argv[2], sys.argv[3:])
