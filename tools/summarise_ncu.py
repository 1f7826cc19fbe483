"""Turn the ncu outputs of tools/gpu_round.sh into the text summaries kept under profiles/.

  python tools/summarise_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.txt
  python tools/summarise_ncu.py full     gpurun_out/prof_conv.ncu-rep > profiles/rNN_ncu_<kernel>_summary.csv
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

CMD = "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
FULL_METRICS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.strip()


def launches(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        tot[k] += us
        cnt[k] += 1
    total = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400: {CMD}")
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES")
    print(f"{'kernel':<70} {'launches':>8} {'total_us':>12} {'share':>7}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k:<70} {cnt[k]:>8} {tot[k]:>12.1f} {100 * tot[k] / total:>6.1f}%")


def full(path, note=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in FULL_METRICS if m in hdr]
    print(f"# {note or 'ncu --set full --clock-control none --import-source on'} : {CMD}")
    print(",".join(f"{hdr[c]} [{units[c]}]" for c in cols))
    for r in body:
        print(",".join(('"' + r[c] + '"') if c == cols[0] else r[c] for c in cols))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
