"""GPU tests of slide ingestion: JPEG-compressed TIFF / SVS-like pyramids decoded by nvJPEG into the device
raster, against libtiff + libjpeg (PIL) reading the same file.  Decoders differ in their IDCT and chroma upsampling, so the
comparison is a tolerance (stated per case); geometry — which pixel lands where, clipping of edge tiles, row bands — is exact."""
import numpy as np
import pytest
import torch
from PIL import Image

from tiff_fixtures import write_stripped, write_tiled_pyramid
from wsi_segmentation_pipeline_b200 import capi, dataset as ds, eval as ev, synth, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _cmp(got, ref, max_abs, mean_abs, what):
    d = np.abs(got.astype(int) - ref.astype(int))
    print(f"{what}: max abs diff {d.max()}, mean {d.mean():.3f}, identical {100 * (d == 0).mean():.1f} %")
    assert got.shape == ref.shape and d.max() <= max_abs and d.mean() <= mean_abs, what


@pytest.mark.parametrize("ycbcr", [False, True])
def test_stripped_tiff_decodes_like_libjpeg(ctx, tmp_path, ycbcr):
    rgb = synth.synth_slide(300, 421, 5)
    p = str(tmp_path / "s.tif")
    write_stripped(p, rgb, ycbcr=ycbcr)
    ref = np.asarray(Image.open(p).convert("RGB"))
    ts = capi.TiffSlide(p)
    got = ts.read_level(ctx, 0).cpu().numpy()
    _cmp(got, ref, 3 if not ycbcr else 6, 0.6, f"stripped ycbcr={ycbcr}")
    band = ts.read_level(ctx, 0, row0=101, rows=77).cpu().numpy()          # a row band that starts and ends inside strips
    np.testing.assert_array_equal(band, got[101:178])
    ts.close()


@pytest.mark.parametrize("subsampling", [0, 2])
def test_tiled_pyramid_levels_and_edge_tiles(ctx, tmp_path, subsampling):
    lv0 = synth.synth_slide(500, 700, 9)
    levels = [lv0, np.ascontiguousarray(lv0[::2, ::2]), np.ascontiguousarray(lv0[::4, ::4])]
    p = str(tmp_path / "pyr.tif")
    write_tiled_pyramid(p, levels, tile=240, subsampling=subsampling)
    ts = capi.TiffSlide(p)
    for lv in range(3):
        Image.MAX_IMAGE_PIXELS = None
        im = Image.open(p)
        im.seek(lv)
        ref = np.asarray(im.convert("RGB"))
        got = ts.read_level(ctx, lv).cpu().numpy()
        _cmp(got, ref, 6 if subsampling == 0 else 40, 0.8 if subsampling == 0 else 2.0, f"tiled level {lv} subsampling {subsampling}")
        _cmp(got, levels[lv], 255, 4.0, f"tiled level {lv} vs the uncompressed source")         # lossy, but every tile in its place
    ts.close()


def test_predict_tumorbed_from_tiff_slide(ctx, tmp_path):
    """The whole path from a file: TiffSlide -> Dataset_wsis (GPU find_nuclei + plan) -> predict_tumorbed == the same call on the
    decoded raster held in memory."""
    full = synth.synth_slide(512, 640, 1234)
    lv2 = np.ascontiguousarray(np.concatenate([full[:256, :320:1][:, :200], np.full((256, 120, 3), 240, np.uint8)], 1))   # tissue + blank margin
    lv0 = np.ascontiguousarray(np.repeat(np.repeat(lv2, 4, 0), 4, 1))
    lv1 = np.ascontiguousarray(np.repeat(np.repeat(lv2, 2, 0), 2, 1))
    p = str(tmp_path / "slide.svs")
    write_tiled_pyramid(p, [lv0, lv1, lv2], tile=256)
    ctx.load_state_dict(capi.ARCH_UNET_R18, weights.random_state_dict("unet", 1))
    ts = capi.TiffSlide(p)
    assert ts.level_downsamples == (1.0, 2.0, 4.0)
    params = {"ph": 64, "pw": 64, "sh": 32, "sw": 32}
    d_file = ds.Dataset_wsis({"slide.svs": ts}, params, scan_level=2, engine=ctx)
    decoded = ts.read_level(ctx, 2).cpu().numpy()
    arr = ds.ArraySlide({2: decoded}, level_downsamples=(1.0, 2.0, 4.0))
    d_mem = ds.Dataset_wsis({"slide.svs": arr}, params, scan_level=2, engine=ctx)
    assert 0 < len(d_file.wsis["slide.svs"]["iterator"].tiles) < capi.plan_tiles(256, 320, 64, 64, 32, 32).shape[0]       # the blank margin is filtered out
    np.testing.assert_array_equal(d_file.wsis["slide.svs"]["iterator"].tiles, d_mem.wsis["slide.svs"]["iterator"].tiles)
    a = ev.predict_tumorbed(ctx, d_file, 0, mode="seg", args={"val_save_pth": str(tmp_path / "out")})
    b = ev.predict_tumorbed(ctx, d_mem, 0, mode="seg")
    np.testing.assert_array_equal(a["slide.svs"]["classes"], b["slide.svs"]["classes"])
    np.testing.assert_array_equal(a["slide.svs"]["heatmap"], b["slide.svs"]["heatmap"])
    assert (tmp_path / "out" / "0" / "slide.svs_128_heatmap.png").exists() and (tmp_path / "out" / "0" / "slide.svs_128_overlay.png").exists()
