"""Summarises ncu outputs into the text files committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv
      per-kernel share of the step from a `--metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summary.py full gpurun_out/prof_rows_r1.ncu-rep
      the headline counters of every launch in a `--set full` capture
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max",
]


def launches(path):
  rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
  hdr = None
  agg = collections.OrderedDict()
  order = []
  for r in rows:
    if r[0] == "ID":
      hdr = r
      continue
    if hdr is None:
      continue
    d = dict(zip(hdr, r))
    try:
      v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
      continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
    name = d["Kernel Name"]
    order.append((name, v))
    a = agg.setdefault(name, [0, 0.0, float("inf"), 0.0])
    a[0] += 1
    a[1] += v
    a[2] = min(a[2], v)
    a[3] = max(a[3], v)
  tot = sum(a[1] for a in agg.values()) or 1.0
  print(f"# {path}: {len(order)} launches, {tot / 1e6:.3f} ms total device time (cold-cache, serialised)")
  print(f"{'ms':>10} {'share':>6} {'n':>5} {'avg us':>10} {'min us':>10} {'max us':>10}  kernel")
  for k, (n, v, lo, hi) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e6:10.3f} {100 * v / tot:5.1f}% {n:5d} {v / n / 1e3:10.1f} {lo / 1e3:10.1f} {hi / 1e3:10.1f}  {k[:110]}")


def full(path):
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  print(f"# {path}")
  for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"== launch {d.get('ID')}: {d.get('Kernel Name', '')[:100]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for k in KEYS:
      if k in d:
        print(f"  {k:70s} {d[k]:>18s} {units[hdr.index(k)]}")
    try:
      rd = float(d["dram__bytes_read.sum"].replace(",", ""))
      wr = float(d["dram__bytes_write.sum"].replace(",", ""))
      mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
      rd *= mul.get(units[hdr.index("dram__bytes_read.sum")], 1)
      wr *= mul.get(units[hdr.index("dram__bytes_write.sum")], 1)
      t = float(d["gpu__time_duration.sum"].replace(",", ""))
      t *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[hdr.index("gpu__time_duration.sum")], 1e-9)
      print(f"  traffic = {rd + wr:.0f} B per launch; {(rd + wr) / t / 1e9:.1f} GB/s under the profiler")
    except (KeyError, ValueError):
      pass


if __name__ == "__main__":
  {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
