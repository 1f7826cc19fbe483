"""GPU parity tests of the tumour-bed post-processing through the C-ABI: morphology bit-exact with cv2
(the reference's own calls), the hull / perimeter chain against the oracle's restatements, the overlays against the
reference's numpy expressions."""
import cv2
import numpy as np
import pytest
import torch

from oracle import wsi_oracle as O
from wsi_segmentation_pipeline_b200 import capi, eval as ev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _blobs(H, W, seed, n=6):
    """A heatmap-like u8 image: a few smooth blobs saturating at 255, plus speckle that the opening must remove."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W))
    for _ in range(n):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(15, min(H, W) / 4)
        img = np.maximum(img, 330.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
    img = np.clip(img, 0, 255)
    speck = rng.random((H, W)) < 0.01
    img[speck] = 255
    return img.astype(np.uint8)


@pytest.mark.parametrize("k", [1, 2, 3, 20, 30, 50, 51])
@pytest.mark.parametrize("shape", [(97, 131), (200, 1500), (301, 64)])
def test_morphology_bit_exact_with_cv2(ctx, k, shape):
    H, W = shape
    gray = _blobs(H, W, k + H)
    binary = (gray >= 253).astype(np.uint8)
    kern = np.ones((k, k), np.uint8)
    for img in (binary, gray):
        for op, cvop in ((capi.MORPH_ERODE, cv2.MORPH_ERODE), (capi.MORPH_DILATE, cv2.MORPH_DILATE), (capi.MORPH_OPEN, cv2.MORPH_OPEN),
                         (capi.MORPH_CLOSE, cv2.MORPH_CLOSE)):
            ref = cv2.morphologyEx(img, cvop, kern)
            np.testing.assert_array_equal(ctx.morph(img, op, k), ref, err_msg=f"op {op} k {k} host")
            dev = ctx.morph(torch.from_numpy(img).cuda(), op, k)
            np.testing.assert_array_equal(dev.cpu().numpy(), ref, err_msg=f"op {op} k {k} device")
    np.testing.assert_array_equal(ctx.morph(binary, capi.MORPH_DILATE, k), cv2.dilate(binary, kern, iterations=1))     # utils/eval.py:96


@pytest.mark.parametrize("rule,open_k,src_kind", [(ev.RULE_CLASSES_GE2, 20, "classes"), (ev.RULE_HEAT_099, 50, "heat"), (ev.RULE_HEAT_090, 30, "heat")])
def test_tumor_bed_chain_matches_oracle(ctx, rule, open_k, src_kind):
    H, W = 420, 533
    heat = _blobs(H, W, 7)
    src = heat if src_kind == "heat" else np.digitize(heat, [60, 150, 230]).astype(np.uint8)      # classes 0..3
    ref = O.tumor_bed(src, rule, open_k, 20)
    assert ref["n_open"] > 0
    for dev in (False, True):
        r = ctx.tumor_bed(torch.from_numpy(src).cuda() if dev else src, rule, open_k, 20)
        get = (lambda k: r[k].cpu().numpy()) if dev else (lambda k: r[k])
        np.testing.assert_array_equal(get("opened"), ref["opened"])           # cv2.MORPH_OPEN, bit-exact
        np.testing.assert_array_equal(get("hull"), ref["hull"])               # convex_hull_image (restated)
        np.testing.assert_array_equal(get("outline"), ref["outline"])         # cv2.dilate(bwperim(hull))
        assert r["n_open"] == ref["n_open"]
    # nothing survives the opening -> empty hull (skimage returns zeros for an empty image)
    empty = ctx.tumor_bed(np.zeros((64, 80), np.uint8), rule, open_k, 20)
    assert empty["n_open"] == 0 and not empty["hull"].any() and not empty["outline"].any()


def test_slide_level_tumor_call(ctx):
    """paper_tools/check_for_false_positives.py:62-72."""
    heat = _blobs(300, 400, 3)
    im = cv2.morphologyEx(np.uint8(heat >= 0.99 * 255), cv2.MORPH_OPEN, kernel=np.ones((50, 50)))
    assert ev.slide_has_tumor(ctx, heat) == (np.count_nonzero(im) / im.size > 0.0)
    assert not ev.slide_has_tumor(ctx, np.full((200, 200), 120, np.uint8))


def test_overlays_match_reference_expressions(ctx):
    rng = np.random.default_rng(5)
    H, W = 123, 217
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    heat = _blobs(H, W, 11)
    ref = np.uint8(img * 0.75 + 255 * np.repeat(np.expand_dims(heat > 255 * 0.99, -1), repeats=3, axis=-1) * 0.25)     # utils/eval.py:266
    np.testing.assert_array_equal(ctx.overlay_heat(img, heat), ref)
    np.testing.assert_array_equal(ctx.overlay_heat(torch.from_numpy(img).cuda(), torch.from_numpy(heat).cuda()).cpu().numpy(), ref)
    # paper_tools/overlay_tb_wsi.py:48-72
    im = cv2.morphologyEx(np.uint8(heat / 255 >= 0.9), cv2.MORPH_OPEN, kernel=np.ones((30, 30)))
    hm = np.repeat((heat * im)[..., np.newaxis], 3, 2)
    perim = cv2.morphologyEx(O.bwperim(O.convex_hull_image(im)).astype(np.uint8), cv2.MORPH_DILATE, kernel=np.ones((20, 20)))
    overlay = 0.65 * img + 0.35 * hm
    yy, xx = np.where(perim)
    overlay[yy, xx, ...] = 0
    np.testing.assert_array_equal(ctx.overlay_bed(img, heat, im, perim), np.uint8(overlay))
