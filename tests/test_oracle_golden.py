"""Pins the CPU oracle (oracle/wsi_oracle.py) against fixtures produced by executing the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import wsi_oracle as O
from wsi_segmentation_pipeline_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_plan_matches_reference_enumeration(golden_dir):
    g = _load(golden_dir, "plan")
    for ci in range(int(g["n_cases"])):
        ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g[f"case{ci}_geom"])
        m = 1.0 if lvl == 2 else 4.0 / 16.0
        tiles = O.plan_tiles(ih, iw, ph, pw, sh, sw, g[f"case{ci}_mask"], m)
        np.testing.assert_array_equal(np.array(tiles, np.int32).reshape(-1, 2), g[f"case{ci}_tiles"])


def test_config1_tile_count():
    assert len(O.plan_tiles(2048, 2048, 256, 256, 128, 128)) == 224
    assert len(O.plan_tiles(20000, 20000, 512, 512, 128, 128)) == 23715


def test_normalise_bit_exact(golden_dir):
    g = _load(golden_dir, "normalise")
    out = O.normalise_tile(g["tile"]).numpy()
    np.testing.assert_array_equal(out, g["out"])


def test_resnet_multipatch_forward(golden_dir):
    g = _load(golden_dir, "resnet_fwd")
    sd = O.random_state_dict("resnet18", 3, with_fc=True)
    xs = torch.randn(2, 16, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y, out = O.resnet_multipatch_forward(sd, xs)
        x4 = O.resnet18_stages(sd, xs[:, 0])[0]
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x4.numpy(), g["x4"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name,arch,mode", [("cls_small", "resnet18_cls", "cls"),
                                            ("cls_m4", "resnet18_cls", "cls"),
                                            ("seg_small", "unet_seg", "seg"),
                                            ("seg_resize2", "unet_seg", "seg"),
                                            ("seg_resize3", "unet_seg", "seg")])
def test_predict_tumorbed_matches_reference(golden_dir, name, arch, mode):
    g = _load(golden_dir, name)
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    m = 1.0 if lvl == 2 else 0.25
    sd = O.random_state_dict("resnet18" if arch == "resnet18_cls" else "unet", int(g["seed"]))
    raster = synth.synth_slide(ih, iw, 1234)
    rs = int(g["scan_resize"]) if "scan_resize" in g else 1
    r = O.predict_tumorbed(sd, arch, raster, g["mask"], ph, pw, sh, sw, mode, batch=16, m=m, scan_resize=rs)
    np.testing.assert_array_equal(np.array(r["tiles"], np.int32).reshape(-1, 2), g["tiles"])
    np.testing.assert_allclose(r["canvas"], g["canvas"], rtol=1e-4, atol=2e-4)
    # probabilities within the north-star fp32 tolerance, layout identical
    np.testing.assert_allclose(r["probs"], g["probs"], atol=1e-4)
    assert r["classes"].shape == g["classes"].shape and r["classes"].dtype == np.uint8
    assert (r["classes"] == g["classes"]).mean() >= 0.999
    assert np.abs(r["heatmap"].astype(int) - g["heatmap"].astype(int)).max() <= 1
    # uncovered pixels: summed logits 0 -> argmax 0, seg heatmap uint8(255*0.5)*mask
    unc = r["counts"] == 0
    assert (r["classes"][unc] == 0).all()
    if mode == "seg":
        assert (r["heatmap"][unc] == 127 * g["mask"][unc]).all()


def test_threshold_probs_floor_matches_reference(golden_dir):
    """threshold_probs with non-zero class_probs (utils/preprocessing.py:156-172): the reference run on a random float64
    canvas — classes bit-exact, probabilities to the last ulp of torch's float64 softmax."""
    g = _load(golden_dir, "threshold_floor")
    for i in range(int(g["n"])):
        classes, probs = O.threshold_probs(g["canvas"].copy(), tuple(float(v) for v in g[f"cp{i}"]))
        np.testing.assert_array_equal(classes, g[f"classes{i}"])
        np.testing.assert_allclose(probs, g[f"probs{i}"], rtol=0, atol=1e-15)
        assert classes.dtype == np.uint8 and probs.dtype == np.float64


@pytest.mark.parametrize("shape", [(256, 320, 64, 80), (100, 130, 33, 47), (64, 64, 32, 32), (40, 60, 80, 120), (97, 61, 13, 9), (37, 53, 37, 53)])
def test_cv2_resize_restatement(shape):
    """A9: the oracle restates cv2.resize(INTER_LINEAR, CV_64F) — pinned against cv2 itself (same image on the GPU box)."""
    cv2 = pytest.importorskip("cv2")
    H, W, H2, W2 = shape
    a = np.random.default_rng(H * W2).standard_normal((H, W)) * 5
    np.testing.assert_allclose(O.cv2_resize_linear(a, W2, H2), cv2.resize(a, (W2, H2)), rtol=0, atol=2e-14)


@pytest.mark.parametrize("shape", [(128, 128, 64, 64), (96, 192, 32, 64), (100, 150, 50, 50), (64, 64, 64, 64), (60, 90, 20, 30), (1024, 40, 512, 20),
                                   (50, 70, 17, 31), (33, 47, 66, 94)])
def test_pil_resize_restatement(shape):
    """scan_resize (utils/dataset.py:180-181): the oracle restates PIL.Image.resize at its default filter — pinned, bit for
    bit, against the Pillow of this image (the same one on the GPU box), and so are the host tables the CUDA path uses."""
    from PIL import Image
    from wsi_segmentation_pipeline_b200 import capi
    H, W, oh, ow = shape
    a = np.random.default_rng(H * ow).integers(0, 256, (H, W, 3), dtype=np.uint8)
    a[: H // 3] = (a[: H // 3] // 128) * 255                                # saturated edges: overshoot must clip like PIL's
    np.testing.assert_array_equal(O.pil_resize(a, ow, oh), np.asarray(Image.fromarray(a).resize((ow, oh))))
    for n_in, n_out in ((W, ow), (H, oh)):
        b, k = capi.resample_coeffs(n_in, n_out)
        b0, k0 = O.pil_resample_coeffs(n_in, n_out)
        np.testing.assert_array_equal(b, b0)
        np.testing.assert_array_equal(k, k0)


@pytest.mark.parametrize("name", ["wsis_l2", "wsis_l1", "wsis_resize2"])
def test_predict_wsis_matches_reference(golden_dir, name):
    """A9: predict_wsis (utils/eval.py:22-81) run unmodified up to pred_to_mask by oracle/ref_harness.py."""
    g = _load(golden_dir, name)
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    m = 1.0 if lvl == 2 else 0.25
    sd = O.random_state_dict("unet", int(g["seed"]))
    raster = synth.synth_slide(ih, iw, 4321)
    r = O.predict_wsis(sd, raster, g["mask"], ph, pw, sh, sw, m=m, batch=16, scan_resize=int(g["scan_resize"]) if "scan_resize" in g else 1)
    np.testing.assert_array_equal(np.array(r["tiles"], np.int32).reshape(-1, 2), g["tiles"])
    assert r["pred"].shape == g["pred"].shape == (4,) + g["mask"].shape
    np.testing.assert_allclose(r["pred"], g["pred"], rtol=1e-4, atol=2e-4)
    assert (r["classes"] == g["classes"]).mean() >= 0.999


def test_predict_reg_tta_matches_reference(golden_dir):
    """predict_reg (utils/eval.py:288-352): the oracle's 4-view TTA mean equals the reference's `preds`."""
    g = _load(golden_dir, "reg_tta")
    sd = O.random_state_dict("unet", int(g["seed"]))
    x = O.gather_tiles(synth.synth_slide(300, 400, 31), [tuple(t) for t in g["tiles"]], 96, 96)
    np.testing.assert_allclose(O.predict_reg_tta(sd, x).ravel(), g["preds"], rtol=0, atol=1e-6)


def test_synth_checksum_is_stable():
    s = synth.synth_slide(2048, 2048, 1234, y0=100, y1=164)
    assert s.shape == (64, 2048, 3)
    full = synth.synth_slide(256, 2048, 1234)
    np.testing.assert_array_equal(full[100:164], s)
    assert synth.checksum(synth.synth_slide(128, 160, 7)) == synth.checksum(synth.synth_slide(128, 160, 7))


def test_find_nuclei_hsv_pinned_by_colorsys_and_cv2():
    """A12: skimage is absent, so the restated saturation rule is pinned against two independent
    witnesses.  (1) stdlib colorsys.rgb_to_hsv — s = (maxc - minc) / maxc in float64 — on ALL 256 x 256 (max, min)
    pairs, fed the same float64 image skimage's img_as_float produces (u8 * (1/255)): identical masks for three
    thresholds.  (2) cv2.cvtColor(COLOR_RGB2HSV) on the float32 image: identical wherever S is further than 1e-5 from
    the threshold (cv2 computes in float32)."""
    import colorsys
    import cv2
    mx, mn = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    lo = np.minimum(mx, mn)
    pairs = np.stack([mx, lo, (mx.astype(int) + lo) // 2], -1).astype(np.uint8)      # max = R, min = G, B in between
    arr = np.multiply(pairs, 1.0 / 255, dtype=np.float64)
    sat = np.array([[colorsys.rgb_to_hsv(*arr[i, j])[1] for j in range(256)] for i in range(256)])
    for mu in (0.1, 0.05, 0.5):
        np.testing.assert_array_equal(O.find_nuclei_hsv(pairs, mu), (sat > mu).astype(np.uint8))
        np.testing.assert_array_equal(O.find_nuclei_hsv(np.ascontiguousarray(pairs[..., [2, 0, 1]]), mu), (sat > mu).astype(np.uint8))
    hsv = cv2.cvtColor(arr.astype(np.float32), cv2.COLOR_RGB2HSV)
    clear = np.abs(hsv[..., 1].astype(np.float64) - 0.1) > 1e-5
    np.testing.assert_array_equal(O.find_nuclei_hsv(pairs)[clear], (hsv[..., 1] > 0.1).astype(np.uint8)[clear])
    # hand-checked pixels: background grey, eosin pink, black (max == 0 -> S = 0), barely-saturated
    rgb = np.array([[[240, 240, 240], [200, 100, 180], [0, 0, 0], [100, 91, 100]]], np.uint8)
    np.testing.assert_array_equal(O.find_nuclei_hsv(rgb), [[0, 1, 0, 0]])
