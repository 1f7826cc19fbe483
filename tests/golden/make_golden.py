"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU.

Run from the repo root in the build container:  python tests/golden/make_golden.py
(the GPU box has no /root/reference; it only reads the committed .npz files).

Scenarios
  plan_*      tile lists from utils.dataset.Dataset_wsi for a grid of geometries / masks
  cls_small   predict_tumorbed(mode='cls'), reference resnets_shift.ResNet via the config-1 adapter
  cls_m4      same with scan_level=1 (m = 0.25: int(m*x) truncation, level-2 canvas)
  seg_small   predict_tumorbed(mode='seg'), reference ResNet encoder + restated smp decoder
  reg_tta     predict_reg: 4-view TTA mean of the regression head (captured from the function's frame)
  wsis_l2/l1  predict_wsis up to the argmax (scan_level 2: resize is the identity; scan_level 1: cv2.resize 4x down)
  resnet_fwd  resnets_shift.ResNet.forward (multi-patch) on a [2,16,3,64,64] batch
  normalise   standard_augmentor(True) on a PIL tile
  threshold_floor   preprocessing.threshold_probs with non-zero class_probs (the per-class floor)
  seg_resize2/3, wsis_resize2   predict_tumorbed / predict_wsis with scan_resize = 2 / 3 (PIL resize of the tiles,
              nearest re-interpolation of the logits); non-square tile for scan_resize 3
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H          # noqa: E402
from oracle import wsi_oracle as O           # noqa: E402
from wsi_segmentation_pipeline_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_mask(h, w, seed=1234, zoom=8):
    return np.ascontiguousarray(synth.synth_mask(h * zoom, w * zoom, seed)[::zoom, ::zoom])


def box_down(a, f):
    h, w = a.shape[0] // f, a.shape[1] // f
    return a[:h * f, :w * f].reshape(h, f, w, f, 3).astype(np.float64).mean(axis=(1, 3)).astype(np.uint8)


def plans():
    R = H.ref_modules()
    cases = [
        # ih, iw, ph, pw, sh, sw, scan_level, mask kind
        (2048, 2048, 256, 256, 128, 128, 2, "ones"),      # BASELINE config 1 -> 224 tiles
        (256, 256, 64, 64, 32, 32, 2, "ones"),            # SURVEY probe: 48 tiles
        (300, 421, 64, 48, 32, 40, 2, "synth"),           # stride does not divide extent
        (200, 333, 64, 64, 96, 80, 2, "synth"),           # stride > tile
        (352, 416, 64, 64, 32, 32, 2, "synth"),
        (640, 768, 128, 128, 64, 64, 1, "synth"),         # m = 0.25
        (131, 140, 64, 64, 32, 32, 2, "ones"),            # few grid rows
        (70, 500, 64, 64, 32, 32, 2, "ones"),             # ih-1-ph <= 1+... : single grid row
    ]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for ci, (ih, iw, ph, pw, sh, sw, lvl, mk) in enumerate(cases):
            m = 16.0 / 16.0 if lvl == 2 else 4.0 / 16.0
            h2, w2 = (ih, iw) if lvl == 2 else (ih // 4, iw // 4)
            mask = np.ones((h2, w2), np.uint8) if mk == "ones" else small_mask(h2, w2, 1234 + ci)
            levels = {2: np.zeros((h2, w2, 3), np.uint8)}
            if lvl != 2:
                levels[lvl] = np.zeros((ih, iw, 3), np.uint8)
            # drive Dataset_wsi directly
            from PIL import Image
            svs = os.path.join(td, f"p{ci}.svs")
            open(svs, "wb").close()
            H._SLIDES[os.path.abspath(svs)] = levels
            R.args.wsi_mask_pth = td
            R.args.scan_level = lvl
            Image.fromarray(mask).save(os.path.join(td, f"p{ci}.svs.png"))
            p = R.preprocessing.DotDict({"ph": ph, "pw": pw, "sh": sh, "sw": sw})
            d = R.dataset.Dataset_wsi(svs, p)
            tiles = np.array(d.datalist, np.int32).reshape(-1, 2)
            out[f"case{ci}_geom"] = np.array([ih, iw, ph, pw, sh, sw, lvl], np.int32)
            out[f"case{ci}_mask"] = mask
            out[f"case{ci}_tiles"] = tiles
            print(f"plan case {ci}: {ih}x{iw} tile {ph}x{pw} stride {sh}x{sw} lvl {lvl} {mk}: {len(tiles)} tiles")
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "plan.npz"), **out)


def predict(name, arch, ih, iw, ph, pw, sh, sw, mode, scan_level=2, seed=0, scan_resize=1):
    sd = O.random_state_dict("resnet18" if arch == "resnet18_cls" else "unet", seed)
    if arch == "resnet18_cls":
        model = H.ResNetClsAdapter(H.make_reference_resnet(sd))
    else:
        model = H.UnetAdapter(sd)
    raster = synth.synth_slide(ih, iw, 1234)
    if scan_level == 2:
        levels = {2: raster}
        mask = small_mask(ih, iw, 77)
    else:
        levels = {scan_level: raster, 2: box_down(raster, 4)}
        mask = small_mask(ih // 4, iw // 4, 77)
    with tempfile.TemporaryDirectory() as td:
        r = H.run_reference_predict_tumorbed(model, levels, mask, td, ph=ph, pw=pw, sh=sh, sw=sw, mode=mode,
                                             scan_level=scan_level, batch=7, scan_resize=scan_resize)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        geom=np.array([ih, iw, ph, pw, sh, sw, scan_level], np.int32), seed=np.array(seed), scan_resize=np.array(scan_resize),
        mask=mask, tiles=np.array(r["tiles"], np.int32).reshape(-1, 2),
        canvas=r["canvas"].astype(np.float32), classes=r["classes"], heatmap=r["heatmap"],
        probs=r["probs"].astype(np.float32))
    print(name, "tiles", len(r["tiles"]), "canvas", r["canvas"].shape, "heat range", r["heatmap"].min(), r["heatmap"].max(),
          "classes hist", np.bincount(r["classes"].ravel(), minlength=4))


def predict_wsis(name, ih, iw, ph, pw, sh, sw, scan_level, seed, scan_resize=1):
    """utils/eval.py:22-81 predict_wsis up to the argmax: canvas at scan-level resolution, cv2.resize to level 2."""
    sd = O.random_state_dict("unet", seed)
    model = H.UnetAdapter(sd)
    raster = synth.synth_slide(ih, iw, 4321)
    if scan_level == 2:
        levels = {2: raster}
        mask = small_mask(ih, iw, 78)
    else:
        levels = {scan_level: raster, 2: box_down(raster, 4)}
        mask = small_mask(ih // 4, iw // 4, 78)
    with tempfile.TemporaryDirectory() as td:
        r = H.run_reference_predict_wsis(model, levels, mask, td, ph=ph, pw=pw, sh=sh, sw=sw, scan_level=scan_level, batch=5,
                                         scan_resize=scan_resize)
    pred = r["pred"]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        geom=np.array([ih, iw, ph, pw, sh, sw, scan_level], np.int32), seed=np.array(seed), mask=mask, scan_resize=np.array(scan_resize),
        tiles=np.array(r["tiles"], np.int32).reshape(-1, 2), pred=pred.astype(np.float32),
        classes=np.argmax(pred, 0).astype(np.uint8))
    print(name, "tiles", len(r["tiles"]), "pred", pred.shape, "classes hist", np.bincount(np.argmax(pred, 0).ravel(), minlength=4))


def reg_tta():
    """predict_reg (utils/eval.py:288-352): 4-view TTA mean of the Regressor head, 7 synthetic 96x96 tiles."""
    sd = O.random_state_dict("unet", 8)
    model = H.UnetAdapter(sd)
    raster = synth.synth_slide(300, 400, 31)
    tiles = [(3, 5), (100, 20), (200, 60), (290, 150), (17, 190), (120, 101), (250, 7)]
    x = O.gather_tiles(raster, tiles, 96, 96)
    with tempfile.TemporaryDirectory() as td:
        preds = H.run_reference_predict_reg(model, x, td, batch=3)
    np.savez_compressed(os.path.join(OUT, "reg_tta.npz"), tiles=np.array(tiles, np.int32), preds=preds.astype(np.float32), seed=np.array(8))
    print("reg_tta preds", preds)


def resnet_fwd():
    sd = O.random_state_dict("resnet18", 3, with_fc=True)
    net = H.make_reference_resnet(sd)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(2, 16, 3, 64, 64, generator=g)
    with torch.no_grad():
        y, out = net(xs)
        feats = H._trunk(net, xs[:, 0])
    np.savez_compressed(os.path.join(OUT, "resnet_fwd.npz"), y=y.numpy(), out=out.numpy(),
                        x4=feats[0].numpy(), x1_mean=np.array([f.mean().item() for f in feats]),
                        x_absmean=np.array([f.abs().mean().item() for f in feats]))
    print("resnet_fwd y", y.shape, "out", out.shape)


def normalise():
    R = H.ref_modules()
    from PIL import Image
    tile = synth.synth_slide(96, 80, 99)
    t = R.preprocessing.standard_augmentor(True)(Image.fromarray(tile))
    np.savez_compressed(os.path.join(OUT, "normalise.npz"), tile=tile, out=t.numpy())


def threshold_floor():
    """preprocessing.threshold_probs with non-zero args.class_probs (the per-class probability floor, myargs.py:15) on a
    random float64 canvas — the floor branch the predict_tumorbed goldens (class_probs = 0) never take."""
    R = H.ref_modules()
    rng = np.random.default_rng(11)
    canvas = rng.standard_normal((4, 48, 60)) * 2.5
    canvas[:, :6] = 0.0                                   # uncovered pixels: softmax 0.25 each, below some floors
    out = {"canvas": canvas}
    for i, cp in enumerate([(0.0, 0.3, 0.25, 0.5), (0.3, 0.3, 0.3, 0.3), (0.9, 0.0, 0.0, 0.0)]):
        R.args.num_classes, R.args.class_probs = 4, list(cp)
        classes, probs = R.preprocessing.threshold_probs(canvas.copy())
        out[f"cp{i}"], out[f"classes{i}"], out[f"probs{i}"] = np.array(cp), classes, probs
    R.args.class_probs = [0.0, 0.0, 0.0, 0.0]
    out["n"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "threshold_floor.npz"), **out)
    print("threshold_floor: classes hist", [np.bincount(out[f"classes{i}"].ravel(), minlength=4).tolist() for i in range(3)])


def resize_only():
    """The scan_resize goldens alone (ph, pw = tile * scan_resize as eval_tumorbed.py:39-40 sets them)."""
    predict("seg_resize2", "unet_seg", 288, 352, 128, 128, 32, 32, "seg", seed=6, scan_resize=2)
    predict_wsis("wsis_resize2", 288, 352, 128, 128, 64, 32, 2, 7, scan_resize=2)
    predict("seg_resize3", "unet_seg", 300, 330, 96, 192, 48, 64, "seg", seed=9, scan_resize=3)


if __name__ == "__main__":
    if "--resize-only" in sys.argv:
        resize_only()
        sys.exit(0)
    if "--threshold-only" in sys.argv:
        threshold_floor()
        sys.exit(0)
    plans()
    normalise()
    resnet_fwd()
    predict("cls_small", "resnet18_cls", 352, 416, 64, 64, 32, 32, "cls")
    predict("cls_m4", "resnet18_cls", 640, 768, 128, 128, 64, 64, "cls", scan_level=1, seed=1)
    predict("seg_small", "unet_seg", 160, 192, 64, 64, 32, 32, "seg", seed=2)
    reg_tta()
    predict_wsis("wsis_l2", 160, 192, 64, 64, 32, 32, 2, 4)
    predict_wsis("wsis_l1", 256, 320, 64, 64, 32, 32, 1, 5)
    resize_only()
    threshold_floor()
