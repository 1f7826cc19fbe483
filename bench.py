#!/usr/bin/env python
"""bench.py — slide megapixels/s of the sliding-window whole-slide inference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (gather+normalise -> U-Net-R18 on tcgen05 -> overlap
stitch -> softmax/argmax/heatmap) over one synthetic slide.  N=1 runs BASELINE.json configs[1]
(20k x 20k slide, U-Net seg, 512 px tiles, stride 128, bf16).  N>1 is weak scaling: the slide is
20k wide and 20k*N tall, partitioned into N row bands with halo rows, one process per
GPU, no data-path collective, and one final NCCL gather of the u8 band masks + heatmaps to rank 0
inside the timed region.

`value`  : whole-job slide-Mpx/s with the raster resident in HBM and outputs left in HBM.
`e2e`    : the same metric through the C-ABI with HOST buffers (pinned raster in, u8 mask+heatmap
           out), host<->device copies inside the timed region.
`roofline`: the dominant kernel (conv_igemm, tcgen05 implicit GEMM): algorithmic conv FLOPs of the
           timed steps / its device time from CUDA events on the launching stream.
`cpu_baseline`: the CPU oracle (port of the reference loop, torch fp32) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TILE, STRIDE = 512, 128
SLIDE_W, BAND_H = 20000, 20000
UNET_GFLOP_PER_TILE_MPX = 164.58      # SURVEY 8d (algorithmic, 2*MAC, conv only)
SEED = 1234


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(threads: int, sample_hw=(1024, 1536)):
    """The CPU oracle (port of utils/eval.py:155-229 + the restated U-Net) on a crop of the same
    synthetic slide.  Returns (tiles/s, seconds, n_tiles)."""
    import torch
    from oracle import wsi_oracle as O
    from wsi_segmentation_pipeline_b200 import synth
    torch.set_num_threads(threads)
    if os.environ.get("WSI_BENCH_TINY"):        # tests/test_host_cpu.py: contract check only
        sample_hw = (640, 768)
    h, w = sample_hw
    raster = synth.synth_slide(h, w, SEED)
    mask = np.ones((h, w), np.uint8)
    sd = O.random_state_dict("unet", 0)
    tiles = O.plan_tiles(h, w, TILE, TILE, STRIDE, STRIDE)
    t0 = time.perf_counter()
    O.predict_tumorbed(sd, "unet_seg", raster, mask, TILE, TILE, STRIDE, STRIDE, "seg", batch=4, tiles=tiles)
    dt = time.perf_counter() - t0
    return len(tiles) / dt, dt, len(tiles)


def full_geometry(n_gpus: int):
    from wsi_segmentation_pipeline_b200 import capi
    ih, iw = BAND_H * n_gpus, SLIDE_W
    tiles = capi.plan_tiles(ih, iw, TILE, TILE, STRIDE, STRIDE)
    return ih, iw, tiles


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the Python
    reference cannot travel to the GPU box) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ih, iw, tiles = full_geometry(args.gpus)
    mpx_per_tile = ih * iw / 1e6 / len(tiles)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(threads, (640, 768))
    rates, secs = [], 0.0
    for _ in range(args.steps):
        tps, dt, nt = cpu_sample(threads)
        rates.append(tps)
        secs += dt
    v = float(np.mean(rates)) * mpx_per_tile
    sample = (f"{args.steps} x (1024x1536 crop of the synthetic slide, {nt} tiles of 512/128, U-Net-R18 fp32 torch CPU); "
              f"tiles/s scaled to slide-Mpx/s by the full slide's tile density ({len(tiles)} tiles / {ih * iw / 1e6:.0f} Mpx)")
    line = {"impl": "reference", "metric": "slide megapixels/sec", "value": v, "unit": "Mpx/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, ih, iw, len(tiles)),
            "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(n_gpus, ih, iw, n_tiles):
    return {"workload": f"BASELINE configs[1]: U-Net-R18 (smp Unet restated) dense seg on a synthetic {iw}x{ih} H&E slide, "
                        f"{TILE}px tiles stride {STRIDE}, all-foreground mask, random-init calibrated weights",
            "tiles": int(n_tiles), "tile": TILE, "stride": STRIDE, "slide_wh": [iw, ih],
            "parallelism": f"row-bands x{n_gpus} (halo = tile overlap, no data-path collective, final NCCL gather of u8 outputs)",
            "l2_policy": "inputs larger than L2 (band raster 1.2 GB, canvas 6.4 GB per GPU)"}


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version at
    NCCL_DEBUG=WARN, torchrun children inherit the fd), so file descriptor 1 is pointed at stderr for the whole run
    and the JSON line is written to the saved descriptor at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-tiles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from oracle import wsi_oracle as O          # weights only (random_state_dict); never on the measured path
    from wsi_segmentation_pipeline_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    assert n_gpus == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"

    ih, iw, tiles = full_geometry(n_gpus)
    bands = capi.band_partition(ih, TILE, STRIDE, n_gpus)
    own0, own1, row0, row1 = (int(v) for v in bands[rank])
    idx = capi.band_tiles(tiles, TILE, 1.0, own0, own1)
    my_tiles = np.ascontiguousarray(tiles[idx])

    ctx = capi.Context(local_rank)
    ctx.load_state_dict(capi.ARCH_UNET_R18, O.random_state_dict("unet", 0))
    ctx.set_option("batch_tiles", args.batch_tiles)
    ctx.set_option("stage_timing", 1)
    raster = ctx.synth_slide(ih, iw, SEED, row0, row1)                 # band + halo, resident in HBM
    rows = own1 - own0
    dev_out = {"classes": torch.empty((rows, iw), dtype=torch.uint8, device="cuda"),
               "heatmap": torch.empty((rows, iw), dtype=torch.uint8, device="cuda")}
    max_rows = int((bands[:, 1] - bands[:, 0]).max())
    gather_buf = [torch.empty((2, max_rows, iw), dtype=torch.uint8, device="cuda") for _ in range(n_gpus)] if (rank == 0 and n_gpus > 1) else None
    send_buf = torch.zeros((2, max_rows, iw), dtype=torch.uint8, device="cuda") if n_gpus > 1 else None

    def step_device():
        sl = ctx.slide_desc(raster, ih, iw, TILE, TILE, row0=row0, rows=row1 - row0, own0=own0, own1=own1)
        ctx.run_slide(sl, my_tiles, capi.HEAD_SEG, device_out=True, out=dev_out)
        if n_gpus > 1:      # the only collective: gather the u8 band outputs on rank 0 over NVLink
            send_buf[0, :rows].copy_(dev_out["classes"])
            send_buf[1, :rows].copy_(dev_out["heatmap"])
            dist.gather(send_buf, gather_buf, dst=0)

    def barrier():
        torch.cuda.synchronize()
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if n_gpus > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    ctx.stage_reset()
    l0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launches - l0
    stats = ctx.stage_stats()
    ms_step = ms_total / args.steps
    mpx = ih * iw / 1e6
    value = mpx / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region -------------
    e2e = None
    if not args.no_e2e:
        host_raster = torch.empty(raster.shape, dtype=torch.uint8, pin_memory=True)
        host_raster.copy_(raster)
        del raster
        torch.cuda.empty_cache()
        host_out = {"classes": torch.empty((rows, iw), dtype=torch.uint8, pin_memory=True),
                    "heatmap": torch.empty((rows, iw), dtype=torch.uint8, pin_memory=True)}
        hsend = torch.zeros((2, max_rows, iw), dtype=torch.uint8, device="cuda") if n_gpus > 1 else None

        def step_host():
            sl = ctx.slide_desc(host_raster, ih, iw, TILE, TILE, row0=row0, rows=row1 - row0, own0=own0, own1=own1)
            ctx.run_slide(sl, my_tiles, capi.HEAD_SEG, device_out=False, out=host_out)
            if n_gpus > 1:
                hsend[0, :rows].copy_(host_out["classes"], non_blocking=True)
                hsend[1, :rows].copy_(host_out["heatmap"], non_blocking=True)
                dist.gather(hsend, gather_buf, dst=0)

        step_host()
        e_steps = max(1, min(args.steps, 3))
        e_ms = timed(step_host, e_steps) / e_steps
        e2e = {"value": mpx / (e_ms * 1e-3), "unit": "Mpx/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(host_raster.numel()) * n_gpus, "d2h_bytes_per_step": int(2 * ih * iw)}

    if rank == 0:
        hbm, tf_sus, tf_burst, which = _peaks()
        conv = stats["conv"]
        conv_tflops = conv["work"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        per_launch_flops = conv["work"] / max(conv["launches"], 1)
        roofline = {"kernel": "conv stage: conv_halo_pair / conv_igemm(_pair) / conv_rowstream(2) / conv_upstream kernels (tcgen05 implicit-GEMM convs + BN/ReLU/residual epilogues)", "bound": "tensor",
                    "achieved": conv_tflops, "peak": tf_sus, "unit": "TFLOP/s", "frac": conv_tflops / tf_sus,
                    "peak_source": f"{which} bf16_tflops_sustained (kernel timed inside a long step)", "traffic": None,
                    "launches": conv["launches"], "avg_launch_ms": conv["ms"] / max(conv["launches"], 1),
                    "algorithmic_flops_per_launch": per_launch_flops,
                    "share_of_step": conv["ms"] / max(ms_total, 1e-9),
                    "other_stages": {k: {"ms_per_step": v["ms"] / args.steps,
                                         "achieved": (v["work"] / (v["ms"] * 1e-3) / (1e12 if k == "stem" else 1e9)) if v["ms"] > 0 else 0.0,
                                         "unit": "TFLOP/s" if k == "stem" else "GB/s",
                                         "frac": ((v["work"] / (v["ms"] * 1e-3) / 1e12 / tf_sus) if k == "stem" else
                                                  (v["work"] / (v["ms"] * 1e-3) / 1e9 / hbm)) if v["ms"] > 0 else 0.0}
                                     for k, v in stats.items() if k not in ("conv",) and v["launches"] > 0}}
        cpu_baseline = None
        if not args.no_cpu_baseline and n_gpus == 1:
            threads = os.cpu_count() or 1
            tps, dt, nt = cpu_sample(threads)
            cpu_baseline = {"value": tps * mpx / len(tiles), "unit": "Mpx/s", "cores": threads, "kind": "port",
                            "sample": f"1024x1536 crop, {nt} tiles of 512/128 in {dt:.1f} s (U-Net-R18 fp32, torch CPU oracle); tiles/s scaled by "
                                      f"the full slide's tile density ({len(tiles)} tiles / {mpx:.0f} Mpx)"}
        line = {"metric": "slide megapixels/sec", "value": value, "unit": "Mpx/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(n_gpus, ih, iw, len(tiles)),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "tile_mpx_per_s": value * len(tiles) * TILE * TILE / (ih * iw),
                "unet_tflops_whole_step": len(tiles) * TILE * TILE / 1e6 * UNET_GFLOP_PER_TILE_MPX / 1e3 / (ms_step * 1e-3)}
        _emit(line)
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
