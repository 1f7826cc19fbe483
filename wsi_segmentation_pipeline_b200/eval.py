"""Host-side mirror of the reference's inference entry points (utils/eval.py).

``predict_tumorbed(model, dataset, ep, mode)`` keeps the reference's signature and side effects
(heatmap / overlay PNGs under ``{val_save_pth}/{ep}/``, ``dataset.wsis[key] = None`` when done,
``model.train()`` on exit) but the loop body — DataLoader, ``model.encoder``/``decoder``,
``.cpu().numpy()``, the per-tile numpy ``+=`` and ``threshold_probs`` (utils/eval.py:190-228) —
is ONE call into libwsi_b200 per slide (``wsi_run_slide``).  No CPU fallback.

Multi-GPU: ``predict_tumorbed_banded`` partitions a slide into row bands (one per rank, halo =
tile overlap), runs each band locally and gathers the u8 band outputs on rank 0 — the only
collective on the path.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import capi
from .dataset import DotDict

DEFAULT_ARGS = DotDict(num_classes=4, class_probs=[0.0, 0.0, 0.0, 0.0], scan_level=2, scan_resize=1, val_save_pth=None,
                       tile_stride_w=128, save_overlay=True)


def _engine_of(model) -> capi.Context:
    if isinstance(model, capi.Context):
        return model
    if hasattr(model, "_engine"):
        return model._engine()
    raise TypeError("model must be a wsi_segmentation_pipeline_b200.models module (or a capi.Context with weights loaded)")


def _merge_args(args):
    a = DotDict(DEFAULT_ARGS)
    if args is not None:
        for k in DEFAULT_ARGS:
            v = args.get(k) if isinstance(args, dict) else getattr(args, k, None)
            if v is not None:
                a[k] = v
    return a


def _scan_resize(a, params) -> int:
    """myargs.py:115 (type=int).  The Dataset params must have been built as the reference's scripts build them:
    ph, pw = tile_h * scan_resize, tile_w * scan_resize (eval_tumorbed.py:39-40)."""
    rs = int(a.scan_resize)
    if rs < 1:
        raise ValueError(f"scan_resize must be a positive integer (got {a.scan_resize})")
    if rs != 1 and (params.ph % rs or params.pw % rs):
        raise ValueError(f"scan_resize={rs}: params.ph/pw ({params.ph}, {params.pw}) must be tile_h/tile_w * scan_resize")
    return rs


def run_slide(ctx: capi.Context, entry: dict, params, mode: str, scan_resize: int = 1, **kw) -> dict:
    """One slide through the CUDA path.  entry: a ``Dataset_wsis.wsis[key]`` dict.  scan_resize (myargs.py:115):
    ``params.ph/pw`` are tile * scan_resize (eval_tumorbed.py:39-40); the windows are resized like PIL does and the
    logits re-interpolated (nearest) inside ``wsi_run_slide``."""
    it = entry["iterator"]
    scan = entry["scan"]
    raster = it.raster()
    ih, iw = raster.shape[:2]
    W2, H2 = scan.level_dimensions[2]                          # utils/eval.py:182
    mask = entry.get("mask")
    mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    sl = ctx.slide_desc(raster, ih, iw, params.ph, params.pw, m=it.m, H2=H2, W2=W2, mask=mask, resize=scan_resize)
    return ctx.run_slide(sl, it.tiles, capi.HEAD_SEG if mode == "seg" else capi.HEAD_CLS, **kw)


def predict_tumorbed(model, dataset, ep, mode: str = "seg", args=None, return_outputs: bool = True):
    """utils/eval.py:155-286.  Returns {key: {'classes': u8 [H2,W2], 'heatmap': u8 [H2,W2]}}."""
    assert mode in ("seg", "cls")
    a = _merge_args(args)
    rs = _scan_resize(a, dataset.params)
    if rs != 1 and mode == "cls":
        # the reference fails here too: F.interpolate on the [B, C] classifier output (utils/eval.py:202-206)
        raise ValueError("scan_resize != 1 with mode='cls': F.interpolate needs (N, C, d1, d2) input (utils/eval.py:202-206 raises)")
    ctx = _engine_of(model)
    ctx.set_class_probs(a.class_probs)
    if hasattr(model, "eval"):
        model.eval()
    save = a.val_save_pth is not None
    if save:
        os.makedirs(f"{a.val_save_pth}/{ep}", exist_ok=True)
    outputs = {}
    for key in list(dataset.wsis):
        entry = dataset.wsis[key]
        if entry is None:
            continue
        r = run_slide(ctx, entry, dataset.params, mode, scan_resize=rs)
        heat, classes = r["heatmap"].numpy(), r["classes"].numpy()
        if save:
            from PIL import Image
            Image.fromarray(heat).save(f"{a.val_save_pth}/{ep}/{key}_{a.tile_stride_w}_heatmap.png")     # :229
            if a.save_overlay:                                                                            # :262-267
                scan = entry["scan"]
                if isinstance(scan, capi.TiffSlide):                                  # thumbnail decoded on the GPU, blended there
                    over = ctx.overlay_heat(scan.read_level(ctx, 2), r["heatmap"].cuda()).cpu().numpy()
                else:
                    img = np.ascontiguousarray(np.asarray(scan.read_region((0, 0), 2, scan.level_dimensions[2]).convert("RGB")), dtype=np.uint8)
                    over = ctx.overlay_heat(img, np.ascontiguousarray(heat))        # img * 0.75 + 255 * (heat > 255 * 0.99) * 0.25, on the device
                Image.fromarray(over).save(f"{a.val_save_pth}/{ep}/{key}_{a.tile_stride_w}_overlay.png")
        if return_outputs:
            outputs[key] = {"classes": classes, "heatmap": heat}
        dataset.wsis[key] = None                                                                          # :282
    if hasattr(model, "train"):
        model.train()                                                                                     # :286
    return outputs


# the reference's three tumour-bed rules (evaluated per u8 level by numpy, exactly as the reference evaluates them per pixel)
RULE_CLASSES_GE2 = lambda v: v >= 2                                                     # utils/eval.py:91  (p >= 2)
RULE_HEAT_099 = lambda v: v >= 0.99 * 255                                               # paper_tools/check_for_false_positives.py:64
RULE_HEAT_090 = lambda v: v / 255 >= 0.9                                                # paper_tools/overlay_tb_wsi.py:48


def tumor_bed(model, src, rule=RULE_CLASSES_GE2, open_k: int = 20, dilate_k: int = 20, **kw) -> dict:
    """The tumour-bed chain that follows the argmax in predict_wsis (utils/eval.py:90-96) and the heatmap in
    paper_tools/: threshold rule -> cv2 MORPH_OPEN -> convex hull image -> perimeter -> dilate, on the device
    (``wsi_tumor_bed``).  ``src``: u8 class mask or heatmap (numpy or CUDA tensor)."""
    return _engine_of(model).tumor_bed(src, rule, open_k, dilate_k, **kw)


def slide_has_tumor(model, heatmap, open_k: int = 50, cancer_thresh: float = 0.0) -> bool:
    """paper_tools/check_for_false_positives.py:62-72: heatmap >= 0.99 * 255, MORPH_OPEN 50 x 50,
    count_nonzero / size > cancer_thresh."""
    r = _engine_of(model).tumor_bed(heatmap, RULE_HEAT_099, open_k, 0, want=())
    return r["n_open"] / float(heatmap.shape[0] * heatmap.shape[1]) > cancer_thresh


def predict_wsis(model, dataset, ep, args=None):
    """utils/eval.py:22-81 up to the argmax: dense seg logits summed on a canvas at SCAN-LEVEL resolution
    (``pred[:, y:y+ph, x:x+pw] += model(batch)``, :46-60), per-class ``cv2.resize`` to the level-2 size (:66-71) and
    ``np.argmax`` (:81).  Returns {key: {'classes': u8 [H2,W2], 'pred': f32 [4,H2,W2]}}.

    Not mirrored: the ground-truth scores, tumour-bed morphology and colour-mask PNG after the
    argmax — ``pred_to_mask`` (utils/preprocessing.py:186-189) raises in the reference itself."""
    a = _merge_args(args)
    rs = _scan_resize(a, dataset.params)
    ctx = _engine_of(model)
    ctx.set_class_probs([0.0, 0.0, 0.0, 0.0])
    if hasattr(model, "eval"):
        model.eval()
    outputs = {}
    for key in list(dataset.wsis):
        entry = dataset.wsis[key]
        if entry is None:
            continue
        it, scan = entry["iterator"], entry["scan"]
        raster = it.raster()
        ih, iw = raster.shape[:2]
        W2, H2 = scan.level_dimensions[2]
        # the canvas lives at scan-level resolution: tiles land at (x, y) unscaled, no foreground mask inside the loop
        sl = ctx.slide_desc(raster, ih, iw, dataset.params.ph, dataset.params.pw, m=1.0, H2=ih, W2=iw, mask=None, resize=rs)
        r = ctx.run_slide(sl, it.tiles, capi.HEAD_SEG, device_out=True, want_canvas=True)
        classes, pred = ctx.resize_argmax(r["canvas"], H2, W2)
        outputs[key] = {"classes": classes.cpu().numpy(), "pred": pred.cpu().numpy()}
    if hasattr(model, "train"):
        model.train()
    return outputs


def predict_reg(model, dataset, ep, clamp: bool = False) -> np.ndarray:
    """utils/eval.py:288-352 (predict_reg) and, with ``clamp``, the per-image arithmetic of predict_breastpathq
    (:384-409): the regression head averaged over the 4 test-time-augmentation views of every (square) tile, one
    ``wsi_forward_batch_tta`` call per batch.  ``dataset`` yields tuples whose first element is the normalised image
    batch f32 [n,3,h,h].  Returns the predictions f32 [N]; the reference's overlay PNGs and printed l1/mse are left to
    the caller."""
    ctx = _engine_of(model)
    if hasattr(model, "eval"):
        model.eval()
    preds = []
    for batch in dataset:
        image = batch[0] if isinstance(batch, (tuple, list)) else batch
        p = ctx.forward_batch_tta(image.cuda() if hasattr(image, "cuda") else image, capi.HEAD_REG).view(-1)
        preds.append(p.cpu().numpy())
    if hasattr(model, "train"):
        model.train()
    out = np.concatenate(preds) if preds else np.zeros((0,), np.float32)
    return np.minimum(np.maximum(out, 0.0), 1.0) if clamp else out


def band_plan(ih: int, ph: int, sh: int, tiles: np.ndarray, m: float, world: int):
    """Per rank: (own0, own1, row0, row1, tile indices).  SURVEY 8e: boundaries on the tile grid;
    a rank evaluates every tile intersecting its band, so nothing is exchanged during compute."""
    bands = capi.band_partition(ih, ph, sh, world)
    return [(int(o0), int(o1), int(r0), int(r1), capi.band_tiles(tiles, ph, m, int(o0), int(o1))) for o0, o1, r0, r1 in bands]


def gather_bands(parts, rows_per_rank, W2: int, rank: int, world: int, device=None):
    """The path's only collective: gather u8 [2, rows_k, W2] (classes, heatmap) band outputs on
    rank 0.  ``parts``: this rank's (classes, heatmap) tensors.  Backend-agnostic (NCCL over NVLink
    on the GPU box, gloo in CPU tests); bands are ragged, so they are padded to the tallest band."""
    import torch
    import torch.distributed as dist
    max_rows = int(max(rows_per_rank))
    cls, heat = parts
    dev = device if device is not None else cls.device
    send = torch.zeros((2, max_rows, W2), dtype=torch.uint8, device=dev)
    send[0, :cls.shape[0]].copy_(cls)
    send[1, :heat.shape[0]].copy_(heat)
    if world == 1:
        return send[0, :rows_per_rank[0]], send[1, :rows_per_rank[0]]
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, bufs, dst=0)
    if rank != 0:
        return None
    classes = torch.cat([bufs[k][0, :rows_per_rank[k]] for k in range(world)])
    heatmap = torch.cat([bufs[k][1, :rows_per_rank[k]] for k in range(world)])
    return classes, heatmap


class PeerResult:
    """The slide's u8 result [2, H2, W2] (classes, heatmap) in rank 0's HBM, mapped into every rank's address space
    through CUDA IPC (wsi_ipc_alloc / wsi_ipc_open).  Each rank hands ``band(own0, own1)`` to ``run_slide`` as its device
    outputs: the fused stitch + finalise kernel stores its rows straight into rank 0's memory over NVLink, so the final
    "gather" of SURVEY 8e costs no extra pass — only ``finish()``, a stream sync + barrier.  Needs ``torch.distributed``
    initialised with one process per GPU on one node; raises ``capi.WsiError`` when peer mapping is unavailable (callers
    fall back to ``gather_bands``)."""

    def __init__(self, ctx: capi.Context, H2: int, W2: int, rank: int, world: int):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.H2, self.W2 = ctx, rank, world, int(H2), int(W2)
        self._ptr, self._owner = None, rank == 0
        handle, err = [None], None
        if rank == 0:
            self._ptr, handle[0] = ctx.ipc_alloc(2 * self.H2 * self.W2)
        if world > 1:
            dist.broadcast_object_list(handle, src=0)
            if rank != 0:
                try:
                    self._ptr = ctx.ipc_open(handle[0])
                except capi.WsiError as e:
                    err = e
            ok = torch.tensor([0 if err else 1], device=torch.device("cuda", ctx.device))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:                       # every rank takes the same decision
                self.close()
                raise err or capi.WsiError(-2, "a peer rank could not map the result buffer")
        self.full = capi.device_u8_tensor(self._ptr, (2, self.H2, self.W2), ctx.device)

    def band(self, own0: int, own1: int) -> dict:
        return {"classes": self.full[0, own0:own1], "heatmap": self.full[1, own0:own1]}

    def finish(self):
        """All ranks: wait for the local kernels, then a barrier; rank 0 gets (classes, heatmap)."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.ctx.device)
        if self.world > 1:
            dist.barrier()
        return (self.full[0], self.full[1]) if self.rank == 0 else None

    def close(self):
        """Collective: every rank calls it (also after a failed mapping)."""
        import torch.distributed as dist
        if getattr(self, "_closed", False):
            return
        self._closed = True
        self.full = None
        if self._ptr is not None and not self._owner:
            self.ctx.ipc_close(self._ptr)
        if self.world > 1 and dist.is_initialized():
            dist.barrier()
        if self._ptr is not None and self._owner:
            self.ctx.ipc_free(self._ptr)
        self._ptr = None


def predict_tumorbed_banded(model, raster_rows_fn, ih: int, iw: int, params, mask: Optional[np.ndarray], rank: int, world: int,
                            mode: str = "seg", peer: Optional[bool] = None):
    """Row-band sharded slide (m == 1).  ``raster_rows_fn(row0, row1)`` returns the u8 [row1-row0, iw, 3]
    raster rows this rank needs (numpy or torch, host or device).  Returns (classes, heatmap) on
    rank 0, None elsewhere.  peer (default: on for world > 1): every rank's finalise kernel writes its rows straight
    into rank 0's result over NVLink (``PeerResult``); otherwise, or when peer mapping is unavailable, one NCCL
    gather of the band outputs (``gather_bands``)."""
    ctx = _engine_of(model)
    tiles = capi.plan_tiles(ih, iw, params.ph, params.pw, params.sh, params.sw, mask, 1.0)
    plan = band_plan(ih, params.ph, params.sh, tiles, 1.0, world)
    own0, own1, row0, row1, idx = plan[rank]
    band = raster_rows_fn(row0, row1)
    bmask = None if mask is None else np.ascontiguousarray(mask[own0:own1])
    sl = ctx.slide_desc(band, ih, iw, params.ph, params.pw, mask=bmask, row0=row0, rows=row1 - row0, own0=own0, own1=own1)
    head = capi.HEAD_SEG if mode == "seg" else capi.HEAD_CLS
    import torch
    if peer is None:
        peer = world > 1 and torch.cuda.is_available()
    if peer:
        try:
            pr = PeerResult(ctx, ih, iw, rank, world)
        except capi.WsiError:
            pr = None                                  # no peer mapping between these processes: NCCL gather below
        if pr is not None:
            ctx.run_slide(sl, tiles[idx], head, device_out=True, out=pr.band(own0, own1))
            res = pr.finish()
            out = None if res is None else (res[0].clone(), res[1].clone())
            pr.close()
            return out
    r = ctx.run_slide(sl, tiles[idx], head, device_out=True)
    rows = [p[1] - p[0] for p in plan]
    return gather_bands((r["classes"], r["heatmap"]), rows, iw, rank, world)
