"""GPU parity tests of the scan_resize != 1 branch (myargs.py:115; utils/dataset.py:180-181; utils/eval.py:52-55, :202-206):
tile windows of (tile_h * r) x (tile_w * r) pixels are resized to the tile exactly as ``PIL.Image.resize`` does at its
default filter (bit-exact, integer arithmetic), the network runs on the resized tiles and the SEG logits are
nearest-upsampled x r before the slice-add.  Checked against Pillow itself, the oracle, and goldens produced by the
unmodified reference run with scan_resize = 2 and 3 (tests/golden/make_golden.py --resize-only)."""
import os

import numpy as np
import pytest
import torch

from oracle import wsi_oracle as O
from wsi_segmentation_pipeline_b200 import capi, dataset as ds, eval as ev, models, synth

pytestmark = pytest.mark.gpu

PROB_TOL_FP32 = 1e-4   # north_star: fp32 probabilities within 1e-4 max-abs
ARGMAX_AGREE = 0.999


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("ih,iw,ph,pw,r,device_raster", [
    (300, 400, 128, 128, 2, False),
    (300, 400, 96, 192, 3, True),         # non-square, odd factor
    (600, 700, 256, 128, 4, True),
    (1100, 1200, 1024, 1024, 2, True),    # the headline tile (512) at scan_resize 2
    (200, 260, 70, 50, 5, False),         # tile 14 x 10: windows clipped on every side by the tile edge
])
def test_gather_resize_is_pil_bit_exact(ctx, ih, iw, ph, pw, r, device_raster):
    """K0r + K0: resized, normalised tiles == standard_augmentor(True)(PIL.resize(window)) — bit for bit."""
    from PIL import Image
    rng = np.random.default_rng(ih + r)
    raster = synth.synth_slide(ih, iw, 99)
    raster[: ih // 4] = rng.integers(0, 2, (ih // 4, iw, 3), dtype=np.uint8) * 255      # saturated noise: bicubic overshoot must clip like PIL's
    tiles = np.array([(0, 0), (iw - pw, ih - ph), (7, 3), ((iw - pw) // 2 + 1, (ih - ph) // 3)], np.int32)
    src = torch.from_numpy(raster).cuda() if device_raster else raster
    sl = ctx.slide_desc(src, ih, iw, ph, pw, resize=r)
    norm, padded = ctx.debug_gather(sl, tiles, want_padded=True)
    th, tw = ph // r, pw // r
    assert norm.shape == (len(tiles), 3, th, tw)
    for i, (x, y) in enumerate(tiles):
        pil = np.asarray(Image.fromarray(raster[y:y + ph, x:x + pw]).resize((tw, th)))
        np.testing.assert_array_equal(O.pil_resize(raster[y:y + ph, x:x + pw], tw, th), pil)
        ref = O.normalise_tile(pil)
        assert torch.equal(norm[i].cpu(), ref), f"tile {i}"
        pad = padded[i].float().cpu()
        assert torch.equal(pad[3:3 + th, 3:3 + tw, :3], ref.permute(1, 2, 0).bfloat16().float())
        assert pad[:3].abs().sum() == 0 and pad[:, :3].abs().sum() == 0 and pad[..., 3].abs().sum() == 0


@pytest.mark.parametrize("name", ["seg_resize2", "seg_resize3"])
def test_run_slide_resize_matches_reference_golden(ctx, golden_dir, name):
    """wsi_run_slide with slide.resize = r against the unmodified reference (predict_tumorbed, scan_resize = r), at
    WSI_PRECISION_FP32 with the north-star fp32 tolerance un-relaxed; bf16 mode bounded by the bf16-emulating oracle."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    r = int(g["scan_resize"])
    sd = O.random_state_dict("unet", int(g["seed"]))
    ctx.load_state_dict(capi.ARCH_UNET_R18, sd)
    raster = synth.synth_slide(ih, iw, 1234)
    mask = np.ascontiguousarray(g["mask"])
    tiles = capi.plan_tiles(ih, iw, ph, pw, sh, sw, mask, 1.0)
    np.testing.assert_array_equal(tiles, g["tiles"])
    sl = ctx.slide_desc(raster, ih, iw, ph, pw, mask=mask, resize=r)
    ctx.set_precision(capi.PRECISION_FP32)
    try:
        f = ctx.run_slide(sl, tiles, capi.HEAD_SEG, want_canvas=True, want_probs=True, want_counts=True)
    finally:
        ctx.set_precision(capi.PRECISION_BF16)
    np.testing.assert_array_equal(f["counts"].numpy(), O.coverage_counts(mask.shape, [tuple(t) for t in tiles], ph, pw, 1.0))
    perr = np.abs(f["probs"].numpy() - g["probs"]).max()
    cerr = np.abs(f["canvas"].numpy() - g["canvas"]).max()
    agree = (f["classes"].numpy() == g["classes"]).mean()
    hd = np.abs(f["heatmap"].numpy().astype(int) - g["heatmap"].astype(int)).max()
    print(f"{name} @fp32-emulated: prob max-abs {perr:.2e}, summed-logit max-abs {cerr:.2e}, argmax agreement {agree:.5f}, heatmap max diff {hd}")
    assert perr <= PROB_TOL_FP32 and agree >= ARGMAX_AGREE and hd <= 1
    # bf16 mode: no further from the reference than bf16 operand rounding alone (oracle emulation, no kernel involved)
    b = ctx.run_slide(sl, tiles, capi.HEAD_SEG, want_probs=True)
    with O.bf16_emulation():
        emu = O.predict_tumorbed(sd, "unet_seg", raster, mask, ph, pw, sh, sw, "seg", batch=16, scan_resize=r)
    noise = np.abs(emu["probs"] - g["probs"])
    err = np.abs(b["probs"].numpy() - g["probs"])
    print(f"{name} @bf16: prob max-abs {err.max():.2e} (bf16 emulation {noise.max():.2e}), argmax agreement "
          f"{(b['classes'].numpy() == g['classes']).mean():.5f} (emulation {(emu['classes'] == g['classes']).mean():.5f})")
    assert err.max() <= 1.5 * noise.max() + 5e-3
    assert (b["classes"].numpy() == g["classes"]).mean() >= (emu["classes"] == g["classes"]).mean() - 5e-3


def test_run_slide_resize_order_band_and_batch_invariance(ctx):
    """The nearest re-interpolation in the fused stitch: shuffled tile order, row bands and a different batch size give
    byte-identical outputs; every canvas pixel of a tile rectangle reads logit (oy // r, ox // r)."""
    ih, iw, ph, pw, sh, sw, r = 700, 520, 128, 128, 64, 96, 2
    sd = O.random_state_dict("unet", 4)
    ctx.load_state_dict(capi.ARCH_UNET_R18, sd)
    raster = torch.from_numpy(synth.synth_slide(ih, iw, 5)).cuda()
    tiles = capi.plan_tiles(ih, iw, ph, pw, sh, sw)
    sl = ctx.slide_desc(raster, ih, iw, ph, pw, resize=r)
    full = ctx.run_slide(sl, tiles, capi.HEAD_SEG, want_canvas=True)
    # the summed logits rebuilt on the host from the per-tile network outputs of the same engine
    y = ctx.forward_tiles(sl, tiles, capi.HEAD_SEG).numpy()
    assert y.shape == (len(tiles), 4, ph // r, pw // r)
    canvas = O.stitch(np.zeros((4, ih, iw), np.float64), [tuple(t) for t in tiles], y, ph, pw, 1.0, up=r)
    np.testing.assert_allclose(full["canvas"].numpy(), canvas.astype(np.float32), rtol=1e-6, atol=1e-6)   # float64 sums in a different order
    perm = np.random.default_rng(0).permutation(len(tiles))
    ctx.set_option("batch_tiles", 5)
    try:
        shuf = ctx.run_slide(sl, tiles[perm], capi.HEAD_SEG)
    finally:
        ctx.set_option("batch_tiles", 0)
    assert torch.equal(shuf["classes"], full["classes"]) and torch.equal(shuf["heatmap"], full["heatmap"])
    parts = []
    for own0, own1, row0, row1 in capi.band_partition(ih, ph, sh, 3):
        idx = capi.band_tiles(tiles, ph, 1.0, int(own0), int(own1))
        bsl = ctx.slide_desc(raster[row0:row1], ih, iw, ph, pw, row0=int(row0), rows=int(row1 - row0), own0=int(own0), own1=int(own1), resize=r)
        parts.append(ctx.run_slide(bsl, tiles[idx], capi.HEAD_SEG))
    assert torch.equal(torch.cat([p["classes"] for p in parts]), full["classes"])
    assert torch.equal(torch.cat([p["heatmap"] for p in parts]), full["heatmap"])


def test_predict_mirrors_with_scan_resize(golden_dir, tmp_path):
    """eval.predict_tumorbed / predict_wsis with args.scan_resize = 2 (Dataset params built like eval_tumorbed.py:39-40)
    against the reference's outputs; mode='cls' raises as the reference does (F.interpolate on a [B, C] tensor)."""
    g = np.load(os.path.join(golden_dir, "seg_resize2.npz"))
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    sd = O.random_state_dict("unet", int(g["seed"]))
    net = models.unet_resnet18()
    net.load_state_dict(sd, strict=False)
    net = net.cuda()
    net._engine().set_precision(capi.PRECISION_FP32)
    mk = lambda raster, mask: ds.Dataset_wsis({"slide0.svs": ds.ArraySlide({2: raster})}, {"ph": ph, "pw": pw, "sh": sh, "sw": sw},
                                              masks={"slide0.svs": mask})
    args = ds.DotDict(val_save_pth=str(tmp_path), tile_stride_w=sw, class_probs=[0.0] * 4, scan_resize=2)
    data = mk(synth.synth_slide(ih, iw, 1234), g["mask"])
    np.testing.assert_array_equal(data.wsis["slide0.svs"]["iterator"].tiles, g["tiles"])
    out = ev.predict_tumorbed(net, data, 0, "seg", args=args)["slide0.svs"]
    assert (out["classes"] == g["classes"]).mean() >= ARGMAX_AGREE
    assert np.abs(out["heatmap"].astype(int) - g["heatmap"].astype(int)).max() <= 1
    with pytest.raises(ValueError):
        ev.predict_tumorbed(net, mk(synth.synth_slide(ih, iw, 1234), g["mask"]), 0, "cls", args=args)
    with pytest.raises(ValueError):
        ev.predict_tumorbed(net, mk(synth.synth_slide(ih, iw, 1234), g["mask"]), 0, "seg", args=ds.DotDict(args, scan_resize=3))

    w = np.load(os.path.join(golden_dir, "wsis_resize2.npz"))
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in w["geom"])
    sd = O.random_state_dict("unet", int(w["seed"]))
    net.load_state_dict(sd, strict=False)
    net = net.cuda()
    net._engine().set_precision(capi.PRECISION_FP32)
    data = mk(synth.synth_slide(ih, iw, 4321), w["mask"])
    np.testing.assert_array_equal(data.wsis["slide0.svs"]["iterator"].tiles, w["tiles"])
    r = ev.predict_wsis(net, data, 0, args=ds.DotDict(scan_resize=2))["slide0.svs"]
    err = np.abs(r["pred"] - w["pred"]).max() / np.abs(w["pred"]).max()
    agree = (r["classes"] == w["classes"]).mean()
    print(f"predict_wsis scan_resize=2 @fp32-emulated: summed logits rel err {err:.2e}, argmax agreement {agree:.5f}")
    assert err <= 5e-5 and agree >= ARGMAX_AGREE
