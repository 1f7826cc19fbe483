#!/usr/bin/env python
"""bench.py — slide megapixels/s of the sliding-window whole-slide inference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (gather+normalise -> U-Net-R18 on tcgen05 -> fused overlap-stitch + softmax /
argmax / heatmap) over one synthetic slide.

Configs (BASELINE.json `configs`):
  c2 (default)  20k x 20k slide, U-Net seg, 512 px tiles, stride 128, bf16 — the config the metric is quoted on.
                N > 1: WEAK scaling, the slide is 20k wide and 20k*N tall.
  c3            100k x 80k slide, same model / tiles, STRONG scaling over N row bands (2/4/8 GPUs; N = 1 also fits).
  c5            50k x 50k slide, stride 256 / 128 / 64, U-Net seg (stitch-heavy) and ResNet-18 cls (stitch-light):
                one line with a `sweep` table (per-stage time shares and roofline fractions).
  c4            regression head over 256 tiles of 512 x 512, 4-view TTA: bf16 vs the fp32-emulated precision
                (max-abs on the clamped scalar), tiles/s.
  c1            2048 x 2048 slide, ResNet-18 cls, 256 px tiles stride 128 (the reference's CPU-runnable case).

`value`   whole-job slide-Mpx/s, raster resident in HBM, outputs gathered on rank 0 in HBM (final NCCL exchange inside
          the timed region for N > 1).
`e2e`     the same through the C-ABI with HOST buffers: pinned raster in, u8 mask + heatmap out, copies inside the timed
          region.  N > 1: every rank uploads its band over its own PCIe link and downloads its rows into one host
          buffer shared by all ranks (POSIX shared memory, page-locked in every process).
`roofline` the conv stage (tcgen05 implicit GEMMs): algorithmic FLOPs / device time from CUDA events on the launching
          stream, plus a per-kernel table from a separate traced step (`kernels`).
`cpu_baseline` the CPU oracle (port of the reference loop, torch fp32) on a bounded sample.
`library_baseline` the same network as plain torch eager modules (cuDNN) on the same GPU, forward only, bf16 autocast
          and fp32/TF32 — the "library bar" of SURVEY 2a / BASELINE.md 3; outside the timed region.
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
R18_GFLOP_PER_TILE_MPX = 72.288
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


# ------------------------------------------------------------------------------------------------------------------
# geometry (pure Python: the reference arm must not load the CUDA library)
# ------------------------------------------------------------------------------------------------------------------
def n_tiles(ih, iw, tile, stride):
    """T = nx*ny + nx + ny of the reference enumeration, all-foreground mask (utils/dataset.py:147-166)."""
    nx, ny = len(range(1, iw - 1 - tile, stride)), len(range(1, ih - 1 - tile, stride))
    return nx * ny + nx + ny


def config_geometry(cfg: str, n_gpus: int):
    """-> (ih, iw, tile, stride, scaling)"""
    if cfg == "c2":
        return BAND_H * n_gpus, SLIDE_W, TILE, STRIDE, "weak"
    if cfg == "c3":
        return 80000, 100000, TILE, STRIDE, "strong"
    if cfg == "c5":
        return 50000, 50000, TILE, STRIDE, "strong"
    if cfg == "c1":
        return 2048, 2048, 256, 128, "strong"
    raise SystemExit(f"unknown config {cfg}")


def workload_config(cfg, n_gpus, ih, iw, tile, stride, T, scaling):
    names = {"c2": "BASELINE configs[1]", "c3": "BASELINE configs[2]", "c5": "BASELINE configs[4]", "c1": "BASELINE configs[0]"}
    model = "ResNet-18 4-class patch classifier (resnets_shift trunk + fc0)" if cfg == "c1" else "U-Net-R18 (smp Unet restated: decoder parity unpinned) dense seg"
    return {"workload": f"{names[cfg]}: {model} on a synthetic {iw}x{ih} H&E slide, {tile}px tiles stride {stride}, "
                        f"all-foreground mask, random-init calibrated weights",
            "tiles": int(T), "tile": tile, "stride": stride, "slide_wh": [iw, ih],
            "parallelism": f"row-bands x{n_gpus} ({scaling} scaling; halo = tile overlap, no data-path collective; how the u8 band outputs reach rank 0: see `gather`)",
            "l2_policy": "inputs larger than L2 (band raster >= 1.2 GB, logit ring >= 3 GB per GPU)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample(threads: int, sample_hw=(1024, 1536), tile=TILE, stride=STRIDE, arch="unet_seg"):
    """The CPU oracle (port of utils/eval.py:155-229 + the restated U-Net) on a crop of the same synthetic slide.
    Returns (tiles/s, seconds, n_tiles)."""
    import torch
    from oracle import wsi_oracle as O
    from wsi_segmentation_pipeline_b200 import synth        # pure numpy generator, no CUDA library involved
    torch.set_num_threads(threads)
    if os.environ.get("WSI_BENCH_TINY"):        # tests/test_host_cpu.py: contract check only
        sample_hw = (640, 768)
    h, w = sample_hw
    raster = synth.synth_slide(h, w, SEED)
    mask = np.ones((h, w), np.uint8)
    sd = O.random_state_dict("unet" if arch.startswith("unet") else "resnet18", 0)
    tiles = O.plan_tiles(h, w, tile, tile, stride, stride)
    t0 = time.perf_counter()
    O.predict_tumorbed(sd, arch, raster, mask, tile, tile, stride, stride, "seg" if arch == "unet_seg" else "cls", batch=4, tiles=tiles)
    dt = time.perf_counter() - t0
    return len(tiles) / dt, dt, len(tiles)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the Python reference cannot
    travel to the GPU box) on all host cores, bounded sample per step.  Pure Python + torch CPU: libwsi_b200.so is
    never loaded in this arm."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = args.config if args.config in ("c1", "c2", "c3", "c5") else "c2"
    ih, iw, tile, stride, scaling = config_geometry(cfg, args.gpus)
    T = n_tiles(ih, iw, tile, stride)
    mpx_per_tile = ih * iw / 1e6 / T
    arch = "resnet18_cls" if cfg == "c1" else "unet_seg"
    sample_hw = (2048, 2048) if cfg == "c1" else (1024, 1536)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(threads, (640, 768), tile, stride, arch)
    rates, secs = [], 0.0
    for _ in range(args.steps):
        tps, dt, nt = cpu_sample(threads, sample_hw, tile, stride, arch)
        rates.append(tps)
        secs += dt
    v = float(np.mean(rates)) * mpx_per_tile
    sample = (f"{args.steps} x ({sample_hw[0]}x{sample_hw[1]} crop of the synthetic slide, {nt} tiles of {tile}/{stride}, {arch} fp32 torch CPU); "
              f"tiles/s scaled to slide-Mpx/s by the full slide's tile density ({T} tiles / {ih * iw / 1e6:.0f} Mpx)")
    line = {"impl": "reference", "metric": "slide megapixels/sec", "value": v, "unit": "Mpx/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args.gpus, ih, iw, tile, stride, T, scaling),
            "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


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


# ------------------------------------------------------------------------------------------------------------------
# synthetic weights (bench-local: nothing under oracle/ is imported by the GPU arm)
# ------------------------------------------------------------------------------------------------------------------
def bench_state_dict(arch: str, seed: int = 0):
    """Random-init weights with the reference's state_dict keys and BatchNorm statistics calibrated on synthetic H&E
    tiles (see wsi_segmentation_pipeline_b200/weights.py)."""
    from wsi_segmentation_pipeline_b200 import weights
    return weights.random_state_dict(arch, seed)


# ------------------------------------------------------------------------------------------------------------------
# library bar: the same network as plain torch eager modules (cuDNN) on the same GPU
# ------------------------------------------------------------------------------------------------------------------
def library_baseline(sd, arch: str, tile: int, n: int, mpx_per_tile: float, iters: int = 5):
    """Forward only (no gather / stitch / finalise) of the reference's module graph in torch eager: F.conv2d +
    F.batch_norm + ReLU + max-pool + nearest-upsample + cat, channels_last, (a) bf16 autocast, (b) fp32 with TF32 allowed.
    Returns tile-Mpx/s and the slide-Mpx/s the forward alone would allow."""
    import torch
    import torch.nn.functional as F
    dev = torch.device("cuda")
    w = {k: v.to(dev) for k, v in sd.items() if hasattr(v, "dtype") and v.dtype.is_floating_point}

    def bn(p, x):
        return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"], False, 0.0, 1e-5)

    def block(p, x, stride):
        idt = x
        out = F.relu(bn(p + ".bn1", F.conv2d(x, w[p + ".conv1.weight"], None, stride, 1)))
        out = bn(p + ".bn2", F.conv2d(out, w[p + ".conv2.weight"], None, 1, 1))
        if (p + ".downsample.0.weight") in w:
            idt = bn(p + ".downsample.1", F.conv2d(x, w[p + ".downsample.0.weight"], None, stride, 0))
        return F.relu(out + idt)

    def forward(x):
        e = "encoder." if arch == "unet" else ""
        x0 = F.relu(bn(e + "bn1", F.conv2d(x, w[e + "conv1.weight"], None, 2, 3)))
        cur = F.max_pool2d(x0, 3, 2, 1)
        feats = []
        for li in (1, 2, 3, 4):
            for b in range(2):
                cur = block(f"{e}layer{li}.{b}", cur, 2 if (li > 1 and b == 0) else 1)
            feats.append(cur)
        if arch != "unet":
            return F.linear(torch.flatten(F.adaptive_avg_pool2d(cur, 1), 1), w["fc0.weight"], w["fc0.bias"])
        xd = feats[3]
        for i, skip in enumerate([feats[2], feats[1], feats[0], x0, None], start=1):
            xd = F.interpolate(xd, scale_factor=2, mode="nearest")
            if skip is not None:
                xd = torch.cat([xd, skip], 1)
            for j in range(2):
                q = f"decoder.layer{i}.block.{j}.block"
                xd = F.relu(bn(q + ".1", F.conv2d(xd, w[q + ".0.weight"], None, 1, 1)))
        return F.conv2d(xd, w["decoder.final_conv.weight"], w["decoder.final_conv.bias"])

    out = {}
    x = torch.randn(n, 3, tile, tile, device=dev).contiguous(memory_format=torch.channels_last)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name in ("bf16_autocast", "tf32"):
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(name == "bf16_autocast")):
                for _ in range(3):
                    forward(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    forward(x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tile_mpx_s = n * tile * tile / 1e6 / (ms * 1e-3)
            out[name] = {"ms_per_batch": ms, "tile_mpx_per_s": tile_mpx_s,
                         "slide_mpx_per_s_forward_only": tile_mpx_s / (tile * tile / 1e6) * mpx_per_tile}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
        del x
        torch.cuda.empty_cache()
    out["what"] = (f"torch {torch.__version__} eager (cuDNN {torch.backends.cudnn.version()}), forward only of the same module graph and weights, "
                   f"batch {n} x {tile}^2, channels_last, cudnn.benchmark; no gather / stitch / finalise")
    return out


# ------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Runner:
    """One slide geometry on this rank's band: device-resident and host-buffer steps."""

    def __init__(self, ctx, rank, world, local_rank, ih, iw, tile, stride, head):
        import torch
        from wsi_segmentation_pipeline_b200 import capi
        self.torch, self.capi, self.ctx = torch, capi, ctx
        self.rank, self.world, self.ih, self.iw, self.tile, self.stride, self.head = rank, world, ih, iw, tile, stride, head
        tiles = capi.plan_tiles(ih, iw, tile, tile, stride, stride)
        self.T = len(tiles)
        self.bands = capi.band_partition(ih, tile, stride, world)
        self.own0, self.own1, self.row0, self.row1 = (int(v) for v in self.bands[rank])
        self.my_tiles = np.ascontiguousarray(tiles[capi.band_tiles(tiles, tile, 1.0, self.own0, self.own1)])
        self.rows = self.own1 - self.own0
        self.raster = ctx.synth_slide(ih, iw, SEED, self.row0, self.row1)            # band + halo, resident in HBM
        self.peer = None
        self._flag = torch.zeros(1, device="cuda")
        if world > 1 and not os.environ.get("WSI_BENCH_NO_PEER"):
            # the result lives in rank 0's HBM and is mapped into every rank (CUDA IPC): each rank's fused stitch + finalise
            # kernel stores its u8 rows straight into it over NVLink — no gather pass
            from wsi_segmentation_pipeline_b200 import eval as ev
            try:
                self.peer = ev.PeerResult(ctx, ih, iw, rank, world)
                self.dev_out = self.peer.band(self.own0, self.own1)
            except capi.WsiError as e:
                if rank == 0:
                    print(f"[bench] peer-mapped result unavailable ({e}); falling back to the NCCL exchange", file=sys.stderr)
                self.peer = None
        if self.peer is not None:
            pass
        elif rank == 0 and world > 1:          # rank 0 holds the whole result; its own band is written in place
            self.full = {k: torch.empty((ih, iw), dtype=torch.uint8, device="cuda") for k in ("classes", "heatmap")}
            self.dev_out = {k: v[self.own0:self.own1] for k, v in self.full.items()}
        else:
            self.dev_out = {k: torch.empty((self.rows, iw), dtype=torch.uint8, device="cuda") for k in ("classes", "heatmap")}
        self.host_raster = None
        self.shared_host = False

    def slide(self, raster):
        return self.ctx.slide_desc(raster, self.ih, self.iw, self.tile, self.tile, row0=self.row0, rows=self.row1 - self.row0,
                                   own0=self.own0, own1=self.own1)

    def exchange(self):
        """The path's only collective: every rank's u8 band outputs go to rank 0 over NVLink (ragged bands -> grouped
        NCCL send/recv straight from / into the output tensors, no staging copies)."""
        import torch.distributed as dist
        if self.world == 1:
            return
        if self.peer is not None:          # rows already sit in rank 0's memory: the "gather" is a closing stream-ordered barrier
            dist.all_reduce(self._flag)     # every rank's finalise kernels precede its contribution on the stream
            return
        ops = []
        if self.rank == 0:
            for k in range(1, self.world):
                o0, o1 = int(self.bands[k][0]), int(self.bands[k][1])
                ops += [dist.P2POp(dist.irecv, self.full["classes"][o0:o1], k), dist.P2POp(dist.irecv, self.full["heatmap"][o0:o1], k)]
        else:
            ops += [dist.P2POp(dist.isend, self.dev_out["classes"], 0), dist.P2POp(dist.isend, self.dev_out["heatmap"], 0)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def step_device(self):
        self.ctx.run_slide(self.slide(self.raster), self.my_tiles, self.head, device_out=True, out=self.dev_out)
        self.exchange()

    # ---- host buffers -------------------------------------------------------------------------
    def prepare_host(self):
        torch = self.torch
        self.host_raster = torch.empty(self.raster.shape, dtype=torch.uint8, pin_memory=True)
        self.host_raster.copy_(self.raster)
        self.raster = None
        torch.cuda.empty_cache()
        if self.world == 1:
            self.host_out = {k: torch.empty((self.rows, self.iw), dtype=torch.uint8, pin_memory=True) for k in ("classes", "heatmap")}
            return
        # one host result shared by all ranks: POSIX shared memory, page-locked in every process.  When /dev/shm cannot
        # hold it (container limit) or cannot be page-locked, every rank keeps its rows in its own pinned buffer instead.
        import shutil
        import torch.distributed as dist
        nbytes = 2 * self.ih * self.iw
        name = [None]
        if self.rank == 0:
            try:
                if shutil.disk_usage("/dev/shm").free > nbytes + (1 << 30):
                    name = [f"/dev/shm/wsi_b200_bench_{os.getpid()}"]
                    with open(name[0], "wb") as f:
                        f.truncate(nbytes)
            except OSError:
                name = [None]
        dist.broadcast_object_list(name, src=0)
        self._shm_path, self._shm, ok, self._regs = name[0], None, 0, []
        if self._shm_path is not None:
            try:
                self._shm = torch.from_file(self._shm_path, shared=True, size=nbytes, dtype=torch.uint8)
                full = self._shm.view(2, self.ih, self.iw)
                views = {"classes": full[0, self.own0:self.own1], "heatmap": full[1, self.own0:self.own1]}
                ok = 1
                for v in views.values():           # every rank page-locks ITS rows of both planes (one registration each)
                    if self.capi.host_register(v.data_ptr(), v.numel()):
                        self._regs.append(v.data_ptr())
                    else:
                        ok = 0
                        break
            except Exception:
                ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.shared_host = bool(int(flag.item()))
        if self.shared_host:
            self.host_out = views
            return
        for ptr in self._regs:
            self.capi.host_unregister(ptr)
        self._regs, self._shm = [], None
        dist.barrier()
        if self.rank == 0 and self._shm_path and os.path.exists(self._shm_path):
            os.unlink(self._shm_path)
        self.host_out = {k: torch.empty((self.rows, self.iw), dtype=torch.uint8, pin_memory=True) for k in ("classes", "heatmap")}

    def step_host(self):
        self.ctx.run_slide(self.slide(self.host_raster), self.my_tiles, self.head, device_out=False, out=self.host_out)

    def release_host(self):
        if self.world > 1 and getattr(self, "_shm", None) is not None:
            import torch.distributed as dist
            for ptr in self._regs:
                self.capi.host_unregister(ptr)
            del self.host_out, self._shm
            dist.barrier()
            if self.rank == 0:
                os.unlink(self._shm_path)


def stage_table(stats, steps, hbm, tf_sus):
    out = {}
    for k, v in stats.items():
        if v["launches"] <= 0 and v["ms"] <= 0:
            continue
        flops = k in ("conv", "stem")
        rate = v["work"] / (v["ms"] * 1e-3) if v["ms"] > 0 else 0.0
        out[k] = {"ms_per_step": v["ms"] / steps, "achieved": rate / (1e12 if flops else 1e9), "unit": "TFLOP/s" if flops else "GB/s",
                  "frac": rate / ((tf_sus * 1e12) if flops else (hbm * 1e9)), "launches_per_step": v["launches"] / steps}
    return out


def kernel_table(ctx, hbm, tf_burst):
    """Per-conv-kernel evidence from a traced step: algorithmic FLOPs and bytes of one launch / its mean device time."""
    rows = []
    for o in ctx.op_stats():
        if not o["count"] or o["ms"] <= 0:
            continue
        t = o["ms"] * 1e-3
        tf, gb = o["flops"] / t / 1e12, o["bytes"] / t / 1e9
        rows.append({"op": o["desc"].strip(), "kernel": o["kernel"], "ms": round(o["ms"], 4), "tflops": round(tf, 1), "gbs": round(gb, 1),
                     "frac_tensor_burst": round(tf / tf_burst, 3), "frac_hbm": round(gb / hbm, 3),
                     "bound": "hbm" if gb / hbm > tf / tf_burst else "tensor"})
    return rows


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels captured with `ncu --set full`
    (profiles/r02_ncu_traffic.json, written by tools/summarise_ncu.py from the committed capture); None when absent."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--batch-tiles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
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
    from wsi_segmentation_pipeline_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    assert n_gpus == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"

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

    hbm, tf_sus, tf_burst, which = _peaks()
    ctx = capi.Context(local_rank)
    ctx.set_option("batch_tiles", args.batch_tiles)

    if args.config == "c4":
        return bench_regression(args, ctx, rank, n_gpus, timed)
    if args.config == "c5":
        return bench_stride_sweep(args, ctx, rank, n_gpus, timed, hbm, tf_sus, which)

    cfg = args.config
    ih, iw, tile, stride, scaling = config_geometry(cfg, n_gpus)
    arch = "resnet18" if cfg == "c1" else "unet"
    head = capi.HEAD_CLS if cfg == "c1" else capi.HEAD_SEG
    sd = bench_state_dict(arch, 0)
    ctx.load_state_dict(capi.ARCH_RESNET18 if arch == "resnet18" else capi.ARCH_UNET_R18, sd)
    ctx.set_option("stage_timing", 1)
    R = Runner(ctx, rank, n_gpus, local_rank, ih, iw, tile, stride, head)
    T = R.T

    for _ in range(args.warmup):
        R.step_device()
    ctx.stage_reset()
    l0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(R.step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launches - l0
    stats = ctx.stage_stats()
    ms_step = ms_total / args.steps
    mpx = ih * iw / 1e6
    value = mpx / (ms_step * 1e-3)

    # ---- per-kernel table: one extra traced step, outside the timed region ---------------------
    kernels = None
    if not args.no_kernel_table:
        ctx.set_option("stage_timing", 0)
        ctx.set_option("op_trace", 1)
        R.step_device()
        barrier()
        if rank == 0:
            kernels = kernel_table(ctx, hbm, tf_burst)
        ctx.set_option("op_trace", 0)

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        # pinned band raster + (N > 1) the shared pinned result: skip e2e rather than exhaust the box's host memory
        import psutil
        need = (R.row1 - R.row0) * iw * 3 * n_gpus + 2 * ih * iw
        if need > 0.4 * psutil.virtual_memory().available:
            args.no_e2e = True
            e2e = {"skipped": f"host buffers of {need / 1e9:.1f} GB exceed 40 % of the available host memory"}
    if not args.no_e2e:
        # a failure here (host memory, shared mapping) must not lose the device-resident line: every rank takes the same
        # decision, the error is reported in the line
        err = None
        try:
            ctx.set_option("stage_timing", 0)
            R.prepare_host()
            R.step_host()
        except Exception as ex:       # noqa: BLE001
            err = repr(ex)[:300]
        bad = torch.tensor([1 if err else 0], device="cuda")
        if n_gpus > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            e2e = {"error": err or "another rank failed to prepare its host buffers"}
        else:
            e_ms = timed(R.step_host, args.steps) / args.steps
            h2d = torch.tensor([int(R.host_raster.numel())], device="cuda", dtype=torch.int64)
            if n_gpus > 1:
                dist.all_reduce(h2d)
            e2e = {"value": mpx / (e_ms * 1e-3), "unit": "Mpx/s", "ms_per_step": e_ms, "steps": args.steps,
                   "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(2 * ih * iw),
                   "how": "pinned band raster -> wsi_run_slide(WSI_MEM_HOST) -> u8 mask + heatmap in host memory"
                          + ("" if n_gpus == 1 else ("; every rank writes its rows into one POSIX-shared, page-locked host buffer over its own PCIe link"
                                                     if R.shared_host else "; every rank keeps its rows in its own pinned host buffer (a shared result could not be page-locked)"))}
            R.release_host()

    if rank == 0:
        conv = stats["conv"]
        conv_tflops = conv["work"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        traffic = ncu_traffic()
        roofline = {"kernel": "conv stage = every tcgen05 implicit-GEMM conv kernel except the stem (conv_halo_pair / conv_igemm(_pair) / "
                              "conv_rowstream(2) / conv_upstream, BN/ReLU/residual epilogues fused); per-kernel rows in `kernels`",
                    "bound": "tensor", "achieved": conv_tflops, "peak": tf_sus, "unit": "TFLOP/s", "frac": conv_tflops / tf_sus,
                    "peak_source": f"{which} bf16_tflops_sustained (kernels timed inside a long step)",
                    "traffic": (traffic or {}).get("top_kernel_dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                    "launches": conv["launches"], "avg_launch_ms": conv["ms"] / max(conv["launches"], 1),
                    "algorithmic_flops_per_launch": conv["work"] / max(conv["launches"], 1),
                    "share_of_step": conv["ms"] / max(ms_total, 1e-9),
                    "other_stages": {k: v for k, v in stage_table(stats, args.steps, hbm, tf_sus).items() if k != "conv"},
                    "kernels": kernels}
        cpu_baseline = None
        if not args.no_cpu_baseline and n_gpus == 1:
            threads = os.cpu_count() or 1
            o_arch = "resnet18_cls" if cfg == "c1" else "unet_seg"
            tps, dt, nt = cpu_sample(threads, (2048, 2048) if cfg == "c1" else (1024, 1536), tile, stride, o_arch)
            cpu_baseline = {"value": tps * mpx / T, "unit": "Mpx/s", "cores": threads, "kind": "port",
                            "sample": f"{nt} tiles of {tile}/{stride} in {dt:.1f} s ({o_arch} fp32, torch CPU oracle); tiles/s scaled by "
                                      f"the full slide's tile density ({T} tiles / {mpx:.0f} Mpx)"}
        library = None
        if not args.no_library and n_gpus == 1:
            try:
                library = library_baseline(sd, arch, tile, 74 if tile == 512 else 256, mpx / T)
            except Exception as e:      # the library bar must never take the bench line down
                library = {"error": repr(e)[:200]}
        gflop = R18_GFLOP_PER_TILE_MPX if cfg == "c1" else UNET_GFLOP_PER_TILE_MPX
        line = {"metric": "slide megapixels/sec", "value": value, "unit": "Mpx/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(cfg, n_gpus, ih, iw, tile, stride, T, scaling),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "library_baseline": library,
                "tile_mpx_per_s": value * T * tile * tile / (ih * iw),
                "net_tflops_whole_step": T * tile * tile / 1e6 * gflop / 1e3 / (ms_step * 1e-3)}
        line["config"]["gather"] = ("none (single GPU)" if n_gpus == 1 else
                                    "peer stores: every rank's stitch+finalise kernel writes its u8 rows into rank 0's HBM over NVLink (CUDA IPC mapping), closing all-reduce barrier"
                                    if R.peer is not None else "grouped NCCL send/recv of the u8 band outputs to rank 0")
        _emit(line)
    if n_gpus > 1:
        if R.peer is not None:
            R.peer.close()
        dist.barrier()
        dist.destroy_process_group()


def bench_stride_sweep(args, ctx, rank, n_gpus, timed, hbm, tf_sus, which):
    """BASELINE configs[4]: 50k x 50k slide, stride 256 / 128 / 64, stitch-heavy (U-Net seg) vs stitch-light (ResNet-18
    cls) — where does the time go as the overlap factor grows 4x -> 16x -> 64x?"""
    import torch
    from wsi_segmentation_pipeline_b200 import capi
    assert n_gpus == 1, "c5 is a single-GPU sweep"
    ih = iw = 50000
    tile = TILE
    ctx.set_option("stage_timing", 1)
    sweep = []
    strides = [int(s) for s in os.environ.get("WSI_C5_STRIDES", "256,128,64").split(",")]
    for arch, head, gflop in (("unet", capi.HEAD_SEG, UNET_GFLOP_PER_TILE_MPX), ("resnet18", capi.HEAD_CLS, R18_GFLOP_PER_TILE_MPX)):
        ctx.load_state_dict(capi.ARCH_UNET_R18 if arch == "unet" else capi.ARCH_RESNET18, bench_state_dict(arch, 0))
        for stride in strides:
            R = Runner(ctx, 0, 1, 0, ih, iw, tile, stride, head)
            R.step_device()                                  # warm-up (plan build, ring allocation)
            ctx.stage_reset()
            ms = timed(R.step_device, 1)
            st = stage_table(ctx.stage_stats(), 1, hbm, tf_sus)
            mpx = ih * iw / 1e6
            tot = sum(v["ms_per_step"] for k, v in st.items() if k not in ("h2d", "d2h"))
            sweep.append({"model": "U-Net-R18 seg" if arch == "unet" else "ResNet-18 cls", "stride": stride, "tiles": R.T,
                          "coverage": R.T * tile * tile / (ih * iw), "ms_per_step": ms, "slide_mpx_per_s": mpx / (ms * 1e-3),
                          "net_tflops": R.T * tile * tile / 1e6 * gflop / 1e3 / (ms * 1e-3),
                          "stages": {k: {"ms": round(v["ms_per_step"], 2), "share": round(v["ms_per_step"] / max(tot, 1e-9), 4),
                                         "achieved": round(v["achieved"], 1), "unit": v["unit"], "frac": round(v["frac"], 3)} for k, v in st.items()},
                          "bound": "conv (tensor pipe)" if st.get("conv", {}).get("ms_per_step", 0) > 0.5 * tot else "memory-bound stages"})
            del R
            torch.cuda.empty_cache()
    ref = next(s for s in sweep if s["model"].startswith("U-Net") and s["stride"] == (128 if 128 in strides else strides[0]))
    line = {"metric": "slide megapixels/sec", "value": ref["slide_mpx_per_s"], "unit": "Mpx/s", "n_gpus": 1, "steps": 1, "warmup": 1,
            "ms_per_step": ref["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: tile-overlap sweep on a synthetic 50000x50000 H&E slide, 512px tiles, stride 256/128/64, "
                                   "U-Net-R18 seg (stitch-heavy) and ResNet-18 cls (stitch-light); value = U-Net seg at stride 128",
                       "tile": tile, "slide_wh": [iw, ih], "l2_policy": "inputs larger than L2"},
            "sweep": sweep, "peak_source": which}
    _emit(line)


def bench_regression(args, ctx, rank, n_gpus, timed):
    """BASELINE configs[3]: cellularity regression (smp encoder + Regressor, utils/eval.py:288-409) over 512 x 512 patches,
    batch 256, 4-view TTA: bf16 vs the fp32-emulated precision vs the CPU oracle (fp32) on a subset."""
    import torch
    from wsi_segmentation_pipeline_b200 import capi
    assert n_gpus == 1, "c4 is a single-GPU config"
    sd = bench_state_dict("unet", 0)
    ctx.load_state_dict(capi.ARCH_UNET_R18, sd)
    n, hw = 256, 512
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(n, 3, hw, hw, generator=g)
    xd = x.cuda()
    res = {}
    for name, prec, sub in (("bf16", capi.PRECISION_BF16, 64), ("fp32_emulated", capi.PRECISION_FP32, 32)):
        ctx.set_precision(prec)
        outs = []

        def step():
            outs.clear()
            for i in range(0, n, sub):
                outs.append(ctx.forward_batch_tta(xd[i:i + sub], capi.HEAD_REG))
        step()
        ms = timed(step, max(1, args.steps if prec == capi.PRECISION_BF16 else 1))
        ms /= max(1, args.steps if prec == capi.PRECISION_BF16 else 1)
        raw = torch.cat(outs).view(-1).cpu()
        res[name] = {"raw": raw, "pred": raw.clamp(0, 1), "ms": ms, "patches_per_s": n / (ms * 1e-3), "sub_batch": sub}
    ctx.set_precision(capi.PRECISION_BF16)
    d16 = (res["bf16"]["pred"] - res["fp32_emulated"]["pred"]).abs().max().item()
    d16_raw = (res["bf16"]["raw"] - res["fp32_emulated"]["raw"]).abs().max().item()
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import wsi_oracle as O
        k = 4
        t0 = time.perf_counter()
        ref_raw = torch.from_numpy(O.predict_reg_tta(sd, x[:k])).view(-1)
        ref = ref_raw.clamp(0, 1)
        dt = time.perf_counter() - t0
        cpu = {"patches": k, "seconds": dt, "patches_per_s": k / dt, "cores": os.cpu_count(),
               "max_abs_fp32_emulated_vs_cpu_fp32": (res["fp32_emulated"]["pred"][:k] - ref).abs().max().item(),
               "max_abs_bf16_vs_cpu_fp32": (res["bf16"]["pred"][:k] - ref).abs().max().item(),
               "unclamped": {"cpu_fp32_range": [ref_raw.min().item(), ref_raw.max().item()],
                             "max_abs_fp32_emulated_vs_cpu_fp32": (res["fp32_emulated"]["raw"][:k] - ref_raw).abs().max().item(),
                             "max_abs_bf16_vs_cpu_fp32": (res["bf16"]["raw"][:k] - ref_raw).abs().max().item()}}
    line = {"metric": "patches/sec (512x512, 4-view TTA regression)", "value": res["bf16"]["patches_per_s"], "unit": "patches/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 1, "ms_per_step": res["bf16"]["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: cellularity regression (ResNet-18 encoder + Regressor) over 256 patches of 512x512, 4-view TTA "
                                   "(predict_reg / predict_breastpathq), random-init calibrated weights"},
            "fp32_emulated": {"patches_per_s": res["fp32_emulated"]["patches_per_s"], "ms_per_step": res["fp32_emulated"]["ms"]},
            "tolerance_check": {"max_abs_bf16_vs_fp32_emulated_on_clamped_scalar": d16, "max_abs_bf16_vs_fp32_emulated_unclamped": d16_raw,
                                "prediction_range_unclamped": [res["fp32_emulated"]["raw"].min().item(), res["fp32_emulated"]["raw"].max().item()],
                                "note": "random-init regressor: where the unclamped outputs fall outside [0, 1] the clamped comparison is vacuous — read the unclamped rows",
                                "cpu_fp32_subset": cpu}}
    _emit(line)


if __name__ == "__main__":
    main()
