"""CPU tests of the host-side mirror (dataset / eval helpers) and of the N>1 gather with gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wsi_oracle as O
from wsi_segmentation_pipeline_b200 import capi, dataset as ds, eval as ev, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dataset_wsis_matches_reference_plan(golden_dir):
    g = np.load(os.path.join(golden_dir, "cls_m4.npz"))
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    raster = synth.synth_slide(ih, iw, 1234)
    lvl2 = np.zeros((ih // 4, iw // 4, 3), np.uint8)
    scan = ds.ArraySlide({1: raster, 2: lvl2})
    assert scan.level_dimensions[1] == (iw, ih) and scan.level_dimensions[2] == (iw // 4, ih // 4)
    d = ds.Dataset_wsis({"s.svs": scan}, {"ph": ph, "pw": pw, "sh": sh, "sw": sw}, scan_level=1, masks={"s.svs": g["mask"]})
    it = d.wsis["s.svs"]["iterator"]
    np.testing.assert_array_equal(it.tiles, g["tiles"])
    assert d.params.iw == iw and d.params.ih == ih and it.m == 0.25
    assert it.raster().shape == (ih, iw, 3)
    tile = np.asarray(scan.read_region((4 * 10, 4 * 20), 1, (pw, ph)).convert("RGB"))
    np.testing.assert_array_equal(tile, raster[20:20 + ph, 10:10 + pw])


def test_empty_slide_is_dropped_like_the_reference():
    scan = ds.ArraySlide({2: np.zeros((200, 200, 3), np.uint8)})
    d = ds.Dataset_wsis({"e.svs": scan}, {"ph": 64, "pw": 64, "sh": 32, "sw": 32}, masks={"e.svs": np.zeros((200, 200), np.uint8)})
    assert "e.svs" not in d.wsis                      # GenerateIterator_wsi returns None (utils/dataset.py:198-201)
    short = ds.ArraySlide({0: np.zeros((64, 64, 3), np.uint8)}, level_downsamples=(1.0,))
    assert len(ds.Dataset_wsi(short, ds.DotDict(ph=8, pw=8, sh=8, sw=8), None, scan_level=2)) == 0   # :123-124


def test_scan_resize_argument_checks(golden_dir):
    """myargs.py:115 scan_resize: the Dataset params carry tile * scan_resize (eval_tumorbed.py:39-40); the mirror checks that
    before anything reaches the device, and the tile plan for such windows equals the reference's (golden of the unmodified
    reference run with scan_resize = 2)."""
    P = ds.DotDict(ph=128, pw=128, sh=32, sw=32)
    assert ev._scan_resize(ds.DotDict(scan_resize=1), P) == 1 and ev._scan_resize(ds.DotDict(scan_resize=2), P) == 2
    for bad in (0, -1, 3):                                   # 128 is not a multiple of 3
        with pytest.raises(ValueError):
            ev._scan_resize(ds.DotDict(scan_resize=bad), P)
    g = np.load(os.path.join(golden_dir, "seg_resize2.npz"))
    ih, iw, ph, pw, sh, sw, lvl = (int(v) for v in g["geom"])
    data = ds.Dataset_wsis({"s.svs": ds.ArraySlide({2: synth.synth_slide(ih, iw, 1234)})}, {"ph": ph, "pw": pw, "sh": sh, "sw": sw},
                           masks={"s.svs": g["mask"]})
    np.testing.assert_array_equal(data.wsis["s.svs"]["iterator"].tiles, g["tiles"])
    # the host tables of the device resize equal the oracle's restatement of Pillow's (also for an upscale and odd sizes)
    for n_in, n_out in ((128, 64), (96, 32), (1024, 512), (50, 17), (33, 66)):
        b, k = capi.resample_coeffs(n_in, n_out)
        b0, k0 = O.pil_resample_coeffs(n_in, n_out)
        np.testing.assert_array_equal(b, b0)
        np.testing.assert_array_equal(k, k0)
        assert (k.sum(axis=1) - (1 << 22)).__abs__().max() <= k.shape[1]      # weights sum to 1 in 22-bit fixed point (rounding)


def test_no_cpu_find_nuclei_in_the_product_package():
    """find_nuclei runs on the GPU only: without an engine and without a cached mask the dataset refuses (no CPU fallback)."""
    assert not hasattr(ds, "find_nuclei_hsv")
    scan = ds.ArraySlide({2: np.full((200, 200, 3), 128, np.uint8)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ds.Dataset_wsi(scan, ds.DotDict(ph=64, pw=64, sh=32, sw=32), None, scan_level=2)


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ih, iw, p, s = 700, 96, 64, 32
        tiles = capi.plan_tiles(ih, iw, p, p, s, s)
        plan = ev.band_plan(ih, p, s, tiles, 1.0, world)
        own0, own1, row0, row1, idx = plan[rank]
        rows = [pl[1] - pl[0] for pl in plan]
        # stand-in band outputs: a function of the absolute canvas row, so stitching errors show
        yy = torch.arange(own0, own1).view(-1, 1).expand(-1, iw)
        cls = (yy % 4).to(torch.uint8).contiguous()
        heat = (yy % 251).to(torch.uint8).contiguous()
        out = ev.gather_bands((cls, heat), rows, iw, rank, world)
        if rank == 0:
            c, h = out
            yy = torch.arange(0, ih).view(-1, 1).expand(-1, iw)
            ok = torch.equal(c, (yy % 4).to(torch.uint8)) and torch.equal(h, (yy % 251).to(torch.uint8))
            q.put(("ok" if ok else "mismatch", rows))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_gather_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + world
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    status, rows = q.get(timeout=10)
    assert status == "ok" and sum(rows) == 700 and len(set(rows)) > 1      # ragged bands


def test_bench_reference_arm_prints_contract_line():
    import json
    import subprocess
    env = dict(os.environ, WSI_BENCH_TINY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpx/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
