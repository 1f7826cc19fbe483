"""CPU ORACLE — test infrastructure only.  NOT a product path, NOT a fallback.

A numpy / torch-fp32 restatement of the reference's sliding-window whole-slide inference path
(acproject/wsi-segmentation-pipeline).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package (``wsi_segmentation_pipeline_b200``) never does and fails loudly when its CUDA
library is missing.

Every function cites the reference lines it restates (paths relative to /root/reference).

Pinning status
--------------
* tile planner, normalisation, ResNet-18 trunk + heads, overlap-add, threshold_probs and the
  heatmap finalisation are pinned against the *unmodified reference code executed in the build
  container* (``oracle/ref_harness.py`` -> ``tests/golden/make_golden.py`` ->
  ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
* the U-Net encoder/decoder arithmetic lives in the third-party package
  ``segmentation_models_pytorch`` (no version pin in the reference — it ships no requirements
  file; the call sites ``eval_tumorbed.py:21-28`` use the pre-0.1 API: ``encoder.out_shapes``,
  callable ``activation``).  That package is absent from /root/reference and from this image,
  so ``unet_*`` below restates its published architecture (smp 0.0.x ``Unet('resnet18')``:
  decoder channels (256,128,64,32,16), nearest x2 upsample, concat [x, skip],
  2x(conv3x3 no-bias + BN + ReLU), final 1x1 conv with bias).  **Parity for the U-Net model is
  UNPINNED** against smp itself; it is pinned only for the loop around it (the reference's own
  ``predict_tumorbed`` is run with this restated model to make the goldens).
* ``pil_resize`` (the ``scan_resize != 1`` tile resize, utils/dataset.py:180-181) restates Pillow's
  antialiased bicubic resample (third-party, not vendored, no version pinned by the reference; this
  image has Pillow 12.2, whose ``Image.resize`` default for RGB is BICUBIC).  Pinned against the
  installed Pillow itself (``tests/test_oracle_golden.py``) and against the unmodified reference run
  with ``scan_resize = 2`` (``tests/golden/seg_resize2.npz``, ``wsis_resize2.npz``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # myargs.py:127-128
IMAGENET_STD = (0.229, 0.224, 0.225)    # myargs.py:129-130
BN_EPS = 1e-5                           # torch.nn.BatchNorm2d default used by resnets_shift.py:37


# --------------------------------------------------------------------------------------------
# A1: tile planner  (utils/dataset.py:143-166, utils/preprocessing.py:60-71)
# --------------------------------------------------------------------------------------------
def isforeground(arr: np.ndarray, thresh: float = 0.05) -> bool:
    """utils/preprocessing.py:60-71 — raises ZeroDivisionError on an empty window, as the
    reference does (python int / python int)."""
    return int(np.count_nonzero(arr)) / int(arr.size) >= thresh


def plan_tiles(ih, iw, ph, pw, sh, sw, mask=None, m=1.0):
    """Tile origins (x, y) in scan-level pixels, in the reference's enumeration order:
    main grid, right column, bottom row, never the corner (utils/dataset.py:147-166).
    ``mask`` None means all-foreground.  ``m`` = downsample[scan_level]/downsample[2] (:144)."""
    dx, dy = int(pw * m), int(ph * m)

    def fg(xpos, ypos):
        if mask is None:
            return True
        yp, xp = int(ypos * m), int(xpos * m)
        return isforeground(mask[yp:yp + dy, xp:xp + dx])

    out = []
    for ypos in range(1, ih - 1 - ph, sh):
        for xpos in range(1, iw - 1 - pw, sw):
            if fg(xpos, ypos):
                out.append((xpos, ypos))
    xpos = iw - 1 - pw
    for ypos in range(1, ih - 1 - ph, sh):
        if fg(xpos, ypos):
            out.append((xpos, ypos))
    ypos = ih - 1 - ph
    for xpos in range(1, iw - 1 - pw, sw):
        if fg(xpos, ypos):
            out.append((xpos, ypos))
    return out


def find_nuclei_hsv(rgb: np.ndarray, mu_percent: float = 0.1) -> np.ndarray:
    """utils/preprocessing.py:74-110 with the defaults mode='hsv', fill_mask=False:
    ``color.rgb2hsv(np.asarray(wsi))[..., 1] > mu_percent`` as uint8.

    scikit-image is NOT installed in this image and the reference pins no version: **parity unpinned** against
    skimage itself.  Restated from skimage's published source (identical in 0.14 ... 0.25): ``rgb2hsv`` first converts
    the uint8 image with ``img_as_float`` = ``np.multiply(image, 1. / 255, dtype=float64)`` (util/dtype.py ``_convert``:
    a multiplication by the reciprocal, not a division), then ``out_v = arr.max(-1)``, ``delta = arr.ptp(-1)``,
    ``out_s = delta / out_v`` with ``out_s[delta == 0.] = 0.``.  The formula is pinned against two independent
    witnesses in tests/test_oracle_golden.py: stdlib ``colorsys.rgb_to_hsv`` (same (max - min) / max in float64, all
    256 x 256 (max, min) pairs) and ``cv2.cvtColor(..., COLOR_RGB2HSV)`` on float32 input (away from the threshold)."""
    arr = np.multiply(np.asarray(rgb)[..., :3], 1.0 / 255, dtype=np.float64)
    out_v = arr.max(-1)
    delta = arr.max(-1) - arr.min(-1)          # np.ptp
    with np.errstate(divide="ignore", invalid="ignore"):
        out_s = delta / out_v
    out_s[delta == 0.0] = 0.0
    return (out_s > mu_percent).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# A2/A3: tile gather + normalise  (utils/dataset.py:171-185, utils/preprocessing.py:206-212)
# --------------------------------------------------------------------------------------------
def normalise_tile(rgb_u8: np.ndarray) -> torch.Tensor:
    """ToTensor (u8 HWC -> f32 CHW, /255) then Normalize (sub mean, div std) in fp32, the op
    order torchvision uses (F.to_tensor: ``.to(float32).div(255)``; F.normalize: ``sub_().div_()``)."""
    t = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1).contiguous()
    t = t.to(torch.float32).div(255)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1, 1)
    return t.sub_(mean).div_(std)


def _bicubic_weight(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resample_coeffs(in_size: int, out_size: int):
    """One axis of Pillow's resample (libImaging/Resample.c: precompute_coeffs with the bicubic filter, support 2,
    then normalize_coeffs_8bpc): per output sample (first input sample, count) and 22-bit fixed-point weights."""
    import math
    bits = 32 - 8 - 2
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / fs
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic_weight((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        if ww != 0.0:
            k = [v / ww for v in k]
        for x, v in enumerate(k):
            kk[xx, x] = int(-0.5 + v * (1 << bits)) if v < 0 else int(0.5 + v * (1 << bits))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h))`` for u8 [H,W,C] (utils/dataset.py:180-181: default filter,
    BICUBIC for RGB in the Pillow of this image): horizontal pass then vertical pass, u8 intermediate,
    out = clip8((2^21 + sum(k * in)) >> 22)."""
    bits = 32 - 8 - 2

    def one_axis(a, out_n):                                       # resample axis 1 of a [R, N, C]
        b, k = pil_resample_coeffs(a.shape[1], out_n)
        out = np.empty((a.shape[0], out_n, a.shape[2]), np.uint8)
        a64 = a.astype(np.int64)
        for xx in range(out_n):
            x0, n = int(b[xx, 0]), int(b[xx, 1])
            acc = (1 << (bits - 1)) + np.tensordot(a64[:, x0:x0 + n, :], k[xx, :n].astype(np.int64), axes=([1], [0]))
            out[:, xx, :] = np.clip(acc >> bits, 0, 255)
        return out

    tmp = one_axis(np.ascontiguousarray(img), out_w)
    return one_axis(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2).copy()


def gather_tiles(raster: np.ndarray, tiles, ph, pw, resize: int = 1) -> torch.Tensor:
    """read_region((ds*x, ds*y), level, (pw, ph)) on an in-memory scan-level raster, (scan_resize != 1:
    ``image.resize((tile_w, tile_h))`` with tile = window / scan_resize, utils/dataset.py:180-181), then the
    eval augmentor.  Tiles never leave the raster (the planner keeps x+pw <= iw-1)."""
    if resize != 1:
        return torch.stack([normalise_tile(pil_resize(raster[y:y + ph, x:x + pw], pw // resize, ph // resize)) for (x, y) in tiles])
    return torch.stack([normalise_tile(raster[y:y + ph, x:x + pw]) for (x, y) in tiles])


# --------------------------------------------------------------------------------------------
# A5: ResNet-18 trunk  (resnets_shift.py:30-65 BasicBlock, :111-217 ResNet)
# --------------------------------------------------------------------------------------------
# bf16 emulation (tests only): when enabled, conv operands (activations and weights) are rounded to
# bf16 exactly where the CUDA path stores bf16 — accumulation, BN, residual add and ReLU stay
# fp32 — so a comparison against the kernels isolates arithmetic bugs from bf16 rounding noise.
_EMULATE_BF16 = False


class bf16_emulation:
    def __enter__(self):
        global _EMULATE_BF16
        self._old, _EMULATE_BF16 = _EMULATE_BF16, True

    def __exit__(self, *a):
        global _EMULATE_BF16
        _EMULATE_BF16 = self._old


def _q(t):
    return t.to(torch.bfloat16).to(torch.float32) if _EMULATE_BF16 else t


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _basic_block(sd, p, x, stride):
    """resnets_shift.py:49-65."""
    idt = x
    out = _q(F.relu(_bn(sd, p + ".bn1", F.conv2d(x, _q(sd[p + ".conv1.weight"]), None, stride, 1))))
    out = _bn(sd, p + ".bn2", F.conv2d(out, _q(sd[p + ".conv2.weight"]), None, 1, 1))
    if (p + ".downsample.0.weight") in sd:
        idt = _q(_bn(sd, p + ".downsample.1", F.conv2d(x, _q(sd[p + ".downsample.0.weight"]), None, stride, 0)))
    return _q(F.relu(out + idt))


def resnet18_stages(sd, x, prefix=""):
    """Returns [x4(512,/32), x3, x2, x1(64,/4), x0(64,/2)] — the order smp's encoder returns
    (SURVEY §8a A7); x4 alone is what resnets_shift.ResNet.forward feeds its heads (:196-204)."""
    p = prefix
    x0 = _q(F.relu(_bn(sd, p + "bn1", F.conv2d(_q(x), _q(sd[p + "conv1.weight"]), None, 2, 3))))
    x1 = F.max_pool2d(x0, 3, 2, 1)
    for b in range(2):
        x1 = _basic_block(sd, f"{p}layer1.{b}", x1, 1)
    feats = [x1]
    cur = x1
    for li in (2, 3, 4):
        for b in range(2):
            cur = _basic_block(sd, f"{p}layer{li}.{b}", cur, 2 if b == 0 else 1)
        feats.append(cur)
    x1, x2, x3, x4 = feats
    return [x4, x3, x2, x1, x0]


def resnet_multipatch_forward(sd, xs):
    """resnets_shift.py:189-217: xs [B,P,3,H,W] -> (cat(y_list,0) [P*B,4] patch-major, fc(features) [B,4])."""
    B, P = xs.shape[:2]
    xs = xs.transpose(0, 1)
    x_list, y_list = [], []
    for ij in range(P):
        f = resnet18_stages(sd, xs[ij])[0]
        f = torch.flatten(F.adaptive_avg_pool2d(f, 1), 1)
        y_list.append(F.linear(f, sd["fc0.weight"], sd["fc0.bias"]))
        x_list.append(f)
    feats = torch.cat(x_list, 1).view(B, -1)
    out = F.linear(F.relu(F.linear(feats, sd["fc.0.weight"], sd["fc.0.bias"])), sd["fc.2.weight"], sd["fc.2.bias"])
    return torch.cat(y_list, 0), out


# --------------------------------------------------------------------------------------------
# A6: heads  (models/models.py:20-38 Classifier, :41-58 Regressor)
# --------------------------------------------------------------------------------------------
def classifier_head(sd, x4, prefix="classifier."):
    f = torch.flatten(F.adaptive_avg_pool2d(x4, 1), 1)
    return F.linear(f, sd[prefix + "fc.0.weight"], sd[prefix + "fc.0.bias"])


def regressor_head(sd, x4, prefix="regressor."):
    f = torch.flatten(F.adaptive_avg_pool2d(x4, 1), 1)
    f = F.relu(F.linear(f, sd[prefix + "fc.0.weight"], sd[prefix + "fc.0.bias"]))
    return F.linear(f, sd[prefix + "fc.2.weight"], sd[prefix + "fc.2.bias"])


def fc0_head(sd, x4):
    """Per-patch head of resnets_shift.ResNet.forward (:206-209), used by the config-1 adapter."""
    f = torch.flatten(F.adaptive_avg_pool2d(x4, 1), 1)
    return F.linear(f, sd["fc0.weight"], sd["fc0.bias"])


# --------------------------------------------------------------------------------------------
# A7: U-Net decoder — RESTATED from the published smp 0.0.x architecture (UNPINNED, see header)
# --------------------------------------------------------------------------------------------
def _conv_bn_relu(sd, p, x, last=False):
    y = F.relu(_bn(sd, p + ".block.1", F.conv2d(x, _q(sd[p + ".block.0.weight"]), None, 1, 1)))
    return y if last else _q(y)      # the CUDA path feeds the last block's fp32 result to the fused 1x1 head


def unet_decoder(sd, feats, prefix="decoder."):
    """feats = [x4, x3, x2, x1, x0]; call site utils/eval.py:200."""
    x = feats[0]
    skips = list(feats[1:]) + [None]
    for i, skip in enumerate(skips, start=1):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        x = _conv_bn_relu(sd, f"{prefix}layer{i}.block.0", x)
        x = _conv_bn_relu(sd, f"{prefix}layer{i}.block.1", x, last=(i == 5))
    return F.conv2d(x, sd[prefix + "final_conv.weight"], sd[prefix + "final_conv.bias"])


def model_forward(sd, arch: str, x: torch.Tensor) -> torch.Tensor:
    """One batch through the model the way predict_tumorbed does (utils/eval.py:196-200).
    arch: 'resnet18_cls' (config-1 adapter: trunk + fc0), 'unet_cls' (encoder + Classifier),
    'unet_seg' (encoder + decoder), 'unet_reg' (encoder + Regressor)."""
    with torch.no_grad():
        if arch == "resnet18_cls":
            return fc0_head(sd, resnet18_stages(sd, x)[0])
        feats = resnet18_stages(sd, x, "encoder.")
        if arch == "unet_cls":
            return classifier_head(sd, feats[0])
        if arch == "unet_reg":
            return regressor_head(sd, feats[0])
        if arch == "unet_seg":
            return unet_decoder(sd, feats)
    raise ValueError(arch)


# --------------------------------------------------------------------------------------------
# A8/A9: overlap-add  (utils/eval.py:179-215 predict_tumorbed, :42-60 predict_wsis)
# --------------------------------------------------------------------------------------------
def stitch(canvas: np.ndarray, tiles, logits: np.ndarray, ph, pw, m=1.0, up: int = 1):
    """canvas [C,H2,W2] f64 += per-tile logits; seg logits [T,C,ph,pw] or cls logits [T,C]
    broadcast over the tile rectangle.  numpy slicing clips silently at the canvas edge."""
    dx, dy = int(m * pw), int(m * ph)
    src = logits
    if up != 1:                                        # F.interpolate(pred_src, (tile_h*r, tile_w*r)), default 'nearest' (utils/eval.py:202-206)
        src = np.repeat(np.repeat(src, up, axis=-2), up, axis=-1)
    while canvas.ndim >= src.ndim:
        src = np.expand_dims(src, -1)
    for j, (x, y) in enumerate(tiles):
        tx, ty = int(m * float(x)), int(m * float(y))
        canvas[:, ty:ty + dy, tx:tx + dx] += src[j]
    return canvas


def coverage_counts(shape_hw, tiles, ph, pw, m=1.0) -> np.ndarray:
    """Overlap count map derived from the tile list (the reference holds no count array;
    SURVEY §0): ``+= 1`` over exactly the rectangles ``stitch`` touches."""
    cnt = np.zeros(shape_hw, np.int32)
    dx, dy = int(m * pw), int(m * ph)
    for (x, y) in tiles:
        tx, ty = int(m * float(x)), int(m * float(y))
        cnt[ty:ty + dy, tx:tx + dx] += 1
    return cnt


def cv2_resize_linear(src: np.ndarray, W2: int, H2: int) -> np.ndarray:
    """cv2.resize(src, (W2, H2)) with the default INTER_LINEAR for a CV_64F plane, restated from OpenCV's
    resize.cpp as opencv-python 4.13 (the version in this image; the reference pins none) evaluates it — checked
    against cv2 itself in tests/test_oracle_golden.py to a few ulp (OpenCV's build contracts a*b+c into FMAs):
    source coordinate ``(d + 0.5) * scale - 0.5`` in double, floor + fraction, the horizontal border clamp resets
    the fraction, weights (1 - f, f), horizontal pass then vertical pass, all in double.  An exact 2x2 decimation
    is computed by OpenCV as INTER_AREA (the mean of the 2x2 block)."""
    src = np.asarray(src, np.float64)
    H, W = src.shape
    if (H2, W2) == (H, W):
        return src.copy()
    if H == 2 * H2 and W == 2 * W2:
        return (src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2]) * 0.25
    sx_scale, sy_scale = 1.0 / (W2 / W), 1.0 / (H2 / H)

    def taps(n_dst, n_src, scale, clamp_resets):
        f = (np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5
        s = np.floor(f).astype(np.int64)
        f = f - s
        if clamp_resets:
            lo, hi = s < 0, s >= n_src - 1
            f[lo | hi] = 0
            s[lo] = 0
            s[hi] = n_src - 1
        i0 = np.clip(s, 0, n_src - 1)
        i1 = np.clip(s + 1, 0, n_src - 1)
        return i0, i1, 1.0 - f, f

    x0, x1, a0, a1 = taps(W2, W, sx_scale, True)
    y0, y1, b0, b1 = taps(H2, H, sy_scale, False)
    rows = src[:, x0] * a0[None, :] + src[:, x1] * a1[None, :]                # hresize
    return rows[y0, :] * b0[:, None] + rows[y1, :] * b1[:, None]              # vresize


def predict_wsis_scores(canvas: np.ndarray, W2: int, H2: int):
    """utils/eval.py:66-71,81: per-class cv2.resize of the scan-level canvas to the level-2 size, then argmax."""
    pred = np.stack([cv2_resize_linear(canvas[c], W2, H2) for c in range(canvas.shape[0])])
    return np.argmax(pred, 0).astype(np.uint8), pred


# --------------------------------------------------------------------------------------------
# A10/A11: softmax over summed logits, floor, argmax, heatmap  (utils/preprocessing.py:156-172,
#          utils/eval.py:217-228)
# --------------------------------------------------------------------------------------------
def threshold_probs(pred: np.ndarray, class_probs=(0.0, 0.0, 0.0, 0.0)):
    p = torch.softmax(torch.from_numpy(pred), dim=0)
    for cj in range(p.shape[0]):
        p[cj, p[cj, ...] < class_probs[cj]] = 0
    p = p.numpy()
    return np.argmax(p, axis=0).astype(np.uint8), p


def finalise_heatmap(probs: np.ndarray, mask: np.ndarray, mode: str) -> np.ndarray:
    h = probs[1] if mode == "cls" else probs[2] + probs[3]
    return np.uint8(255 * (mask * h))


def predict_tumorbed(sd, arch, raster, mask, ph, pw, sh, sw, mode, batch=16, m=1.0, tiles=None,
                     class_probs=(0.0, 0.0, 0.0, 0.0), scan_resize: int = 1):
    """End-to-end restatement of utils/eval.py:155-229 for one slide whose scan-level raster is
    ``raster`` u8 [ih,iw,3] and whose level-2 canvas is [int(ih*m), int(iw*m)] (== mask.shape)."""
    ih, iw = raster.shape[:2]
    if tiles is None:
        tiles = plan_tiles(ih, iw, ph, pw, sh, sw, mask, m)
    C = 4
    canvas = np.zeros((C,) + tuple(mask.shape), np.float64)
    all_logits = []
    for i in range(0, len(tiles), batch):
        chunk = tiles[i:i + batch]
        x = gather_tiles(raster, chunk, ph, pw, scan_resize)           # ph, pw = tile * scan_resize (eval_tumorbed.py:39-40)
        y = model_forward(sd, arch, x).numpy()
        all_logits.append(y)
        stitch(canvas, chunk, y, ph, pw, m, scan_resize)
    classes, probs = threshold_probs(canvas, class_probs)
    heat = finalise_heatmap(probs, mask, mode)
    return {"tiles": tiles, "canvas": canvas, "classes": classes, "probs": probs, "heatmap": heat,
            "counts": coverage_counts(mask.shape, tiles, ph, pw, m),
            "logits": np.concatenate(all_logits) if all_logits else np.zeros((0, C), np.float32)}


def predict_wsis(sd, raster, mask, ph, pw, sh, sw, m=1.0, batch=16, tiles=None, scan_resize: int = 1):
    """utils/eval.py:22-81 for one slide, up to the argmax: the canvas lives at SCAN-LEVEL resolution (:44-47,
    tiles land unscaled, :56-60), is resized per class to the level-2 size == mask.shape (:66-71), then argmax (:81).
    ``m`` only enters the tile plan (foreground test of Dataset_wsi against the level-2 mask)."""
    ih, iw = raster.shape[:2]
    if tiles is None:
        tiles = plan_tiles(ih, iw, ph, pw, sh, sw, mask, m)
    canvas = np.zeros((4, ih, iw), np.float64)
    for i in range(0, len(tiles), batch):
        chunk = tiles[i:i + batch]
        y = model_forward(sd, "unet_seg", gather_tiles(raster, chunk, ph, pw, scan_resize)).numpy()
        stitch(canvas, chunk, y, ph, pw, 1.0, scan_resize)
    H2, W2 = mask.shape
    classes, pred = predict_wsis_scores(canvas, W2, H2)
    return {"tiles": tiles, "canvas": canvas, "pred": pred, "classes": classes}


def predict_reg_tta(sd, x: "torch.Tensor", arch: str = "unet_reg") -> np.ndarray:
    """utils/eval.py:303-334 (predict_reg) / :384-405 (predict_breastpathq): the regressor evaluated on the 4 views
    [x, x.transpose(2,3), x.flip(2), x.transpose(2,3).flip(3)], summed in that order and divided by 4."""
    views = [x, x.transpose(2, 3), x.flip(2), x.transpose(2, 3).flip(3)]
    acc = None
    for v in views:
        p = model_forward(sd, arch, v.contiguous())
        p = p.reshape(p.shape[0], -1)
        acc = p if acc is None else acc + p
    return (acc / len(views)).numpy()


# --------------------------------------------------------------------------------------------
# tumour-bed post-processing (utils/eval.py:90-96,262-267; paper_tools/overlay_tb_wsi.py; check_for_false_positives.py)
#
# cv2 IS installed: the morphology oracle is cv2 itself (tests call cv2.morphologyEx / cv2.dilate directly).
# scikit-image and mahotas are NOT: convex_hull_image and bwperim are restated from their published sources —
# **parity unpinned** against the packages themselves.
# --------------------------------------------------------------------------------------------
def convex_hull_image(image: np.ndarray) -> np.ndarray:
    """skimage.morphology.convex_hull_image(image) with its defaults (offset_coordinates=True, tolerance=1e-10,
    include_borders=True), restated: candidate points = first / last set pixel of every row and column
    (``possible_hull``), each replaced by the midpoints of its four pixel edges (``_offsets_diamond``), duplicates
    removed, ``scipy.spatial.ConvexHull`` (qhull — the same library skimage calls), then every pixel centre inside or on
    the hull polygon (``grid_points_in_poly(..., binarize=False) >= 1``).  Coordinates are multiples of 0.5, so the
    half-plane tests below are exact in float64."""
    from scipy.spatial import ConvexHull
    img = np.asarray(image) != 0
    out = np.zeros(img.shape, bool)
    if not img.any():
        return out
    pts = set()
    for r in np.nonzero(img.any(1))[0]:
        c = np.nonzero(img[r])[0]
        pts.add((int(r), int(c[0]))); pts.add((int(r), int(c[-1])))
    for c in np.nonzero(img.any(0))[0]:
        r = np.nonzero(img[:, c])[0]
        pts.add((int(r[0]), int(c))); pts.add((int(r[-1]), int(c)))
    coords = np.array(sorted(pts), np.float64)
    offsets = np.array([[-0.5, 0.0], [0.5, 0.0], [0.0, -0.5], [0.0, 0.5]])
    coords = np.unique((coords[:, None, :] + offsets).reshape(-1, 2), axis=0)
    hull = ConvexHull(coords)
    v = hull.points[hull.vertices]                   # polygon vertices in (row, col); orientation taken from the signed area
    rr, cc = np.mgrid[0:img.shape[0], 0:img.shape[1]]
    inside = np.ones(img.shape, bool)
    n = len(v)
    area2 = sum(v[i, 0] * v[(i + 1) % n, 1] - v[(i + 1) % n, 0] * v[i, 1] for i in range(n))
    orient = 1.0 if area2 > 0 else -1.0
    for i in range(n):
        a, b = v[i], v[(i + 1) % n]
        cross = (b[0] - a[0]) * (cc - a[1]) - (b[1] - a[1]) * (rr - a[0])
        inside &= (orient * cross >= 0)
    return inside


def bwperim(bw: np.ndarray) -> np.ndarray:
    """mahotas.bwperim(bw, n=4) restated: a set pixel belongs to the perimeter iff at least one of its 4 neighbours is
    unset; pixels outside the image count as unset."""
    b = np.asarray(bw) != 0
    p = np.pad(b, 1, constant_values=False)
    full = p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
    return b & ~full


def tumor_bed(src: np.ndarray, rule, open_k: int = 20, dilate_k: int = 20) -> dict:
    """utils/eval.py:90-96 with cv2 for the morphology (the reference's own calls) and the restatements above."""
    import cv2
    tb = np.asarray(rule(src)).astype(np.uint8)
    tb = cv2.morphologyEx(tb, cv2.MORPH_OPEN, np.ones((open_k, open_k), dtype=np.uint8))
    hull = convex_hull_image(tb)
    per = bwperim(hull).astype(np.uint8)
    outline = cv2.dilate(per, np.ones((dilate_k, dilate_k), dtype=np.uint8), iterations=1) if dilate_k > 1 else per
    return {"opened": tb, "hull": hull.astype(np.uint8), "outline": outline, "n_open": int(np.count_nonzero(tb))}


# --------------------------------------------------------------------------------------------
# random-init weights with the reference's state_dict keys: the synthetic-checkpoint generator lives in the
# package (wsi_segmentation_pipeline_b200/weights.py — test / bench DATA, shared by bench.py's GPU arm, which must not
# import anything under oracle/); re-exported here for the tests.
# --------------------------------------------------------------------------------------------
from wsi_segmentation_pipeline_b200.weights import calibrate_bn, random_resnet18_trunk, random_state_dict  # noqa: E402,F401
