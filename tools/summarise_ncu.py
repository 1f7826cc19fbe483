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




def batch(path, bench_json=None):
    """profiles/rNN_kernels.csv: one row per kernel of a whole batch (ncu --metrics ... --csv log of tools/gpu_ncu_r2.sh),
    joined with bench.py's live `kernels` table (CUDA-event time, algorithmic FLOPs / bytes) when given."""
    import json
    rows = [l for l in open(path) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = {}
    order = []
    for r in rd:
        key = (r["ID"], short(r["Kernel Name"]))
        if key not in per:
            per[key] = {}
            order.append(key)
        per[key][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        per[key]["unit:" + r["Metric Name"]] = r.get("Metric Unit", "")
    agg = defaultdict(lambda: defaultdict(float))
    for key in order:
        m = per[key]
        k = key[1]
        t = m.get("gpu__time_duration.sum", 0.0)
        u = m.get("unit:gpu__time_duration.sum", "ns")
        us = t / 1e3 if u in ("ns", "nsecond") else (t if u in ("us", "usecond") else t * 1e3)

        def byt(name):
            v, un = m.get(name, 0.0), m.get("unit:" + name, "byte")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
        a = agg[k]
        a["n"] += 1
        a["us"] += us
        a["tensor_x_us"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us
        a["dram"] += byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum")
        a["dram_pct_x_us"] += m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * us
        a["regs"] = m.get("launch__registers_per_thread", 0.0)
    total = sum(a["us"] for a in agg.values())
    live = {}
    if bench_json:
        d = json.load(open(bench_json))
        for row in (d.get("roofline", {}).get("kernels") or []):
            live.setdefault(row["kernel"], []).append(row)
    print(f"# four consecutive batches of 444 tiles of 512^2 (130 launches, steady state) inside: {CMD} --no-library --no-kernel-table ; ncu --metrics ... --clock-control none -s 2200 -c 130")
    print("# ncu times are cold-cache and serialised (compare SHARES); live_* columns come from bench.py's traced step (CUDA events, warm)")
    print("kernel,launches,ncu_us_total,share_pct,tensor_pipe_active_pct,dram_bytes_per_launch_MB,dram_throughput_pct,registers,live_ms_sum,live_alg_tflops,live_alg_gbs")
    conv_t = conv_tx = 0.0
    for k in sorted(agg, key=lambda k: -agg[k]["us"]):
        a = agg[k]
        tens = a["tensor_x_us"] / a["us"] if a["us"] else 0.0
        if "conv_" in k or "stem" in k:
            conv_t += a["us"]
            conv_tx += a["tensor_x_us"]
        def sig(n):          # (kernel base name, first template integer)
            m = re.search(r"([A-Za-z_0-9]+)\s*<\s*(\d+)", n)
            return (m.group(1), m.group(2)) if m else (re.sub(r"^void\s+", "", n).strip(), "")
        lv = None
        for name, rws in live.items():
            if sig(name) == sig(k):
                lv = (lv or []) + rws
        lms = sum(r["ms"] for r in lv) if lv else ""
        ltf = (sum(r["tflops"] * r["ms"] for r in lv) / max(sum(r["ms"] for r in lv), 1e-9)) if lv else ""
        lgb = (sum(r["gbs"] * r["ms"] for r in lv) / max(sum(r["ms"] for r in lv), 1e-9)) if lv else ""
        fmt = lambda v: f"{v:.1f}" if isinstance(v, float) else str(v)
        print(f"\"{k}\",{int(a['n'])},{a['us']:.1f},{100 * a['us'] / total:.1f},{tens:.1f},{a['dram'] / a['n'] / 1e6:.1f},{a['dram_pct_x_us'] / a['us']:.1f},{int(a['regs'])},{fmt(lms)},{fmt(ltf)},{fmt(lgb)}")
    print(f"# time-weighted tensor-pipe activity over the conv kernels (stem included): {conv_tx / max(conv_t, 1e-9):.1f} % of peak sustained active; conv kernels = {100 * conv_t / total:.1f} % of the batch")


if __name__ == "__main__":
    {"launches": launches, "full": full, "batch": batch}[sys.argv[1]](*sys.argv[2:])
