"""Host-side mirror of the reference's whole-slide datasets (utils/dataset.py:83-201).

Same names, same ``.params`` / ``.wsis`` layout, same tile enumeration (the C planner
``wsi_plan_tiles`` is bit-exact with ``Dataset_wsi.__init__``) — but no DataLoader workers: the
"iterator" of a slide is the planned tile list plus the decoded scan-level raster, which the
CUDA gather kernel cuts tiles from (replacing ``__getitem__``'s read_region + ToTensor + Normalize,
utils/dataset.py:171-185).
"""
from __future__ import annotations

import glob
import os
from typing import Mapping, Optional

import numpy as np

from . import capi


class DotDict(dict):
    """utils/preprocessing.py:50-57."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


class ArraySlide:
    """OpenSlide-like view of in-memory rasters: ``levels`` maps level -> u8 [h, w, 3].
    ``level_dimensions`` are (w, h); ``read_region`` takes level-0 coordinates and returns RGBA
    (openslide semantics, utils/dataset.py:175-178)."""

    def __init__(self, levels: Mapping[int, np.ndarray], level_downsamples=(1.0, 4.0, 16.0)):
        self._levels = dict(levels)
        self.level_downsamples = tuple(float(d) for d in level_downsamples)
        ref_level = max(self._levels)
        h, w = self._levels[ref_level].shape[:2]
        ds_ref = self.level_downsamples[ref_level]
        dims = []
        for lv, d in enumerate(self.level_downsamples):
            if lv in self._levels:
                dims.append((self._levels[lv].shape[1], self._levels[lv].shape[0]))
            else:
                dims.append((int(w * ds_ref / d), int(h * ds_ref / d)))
        self.level_dimensions = tuple(dims)

    def level_array(self, level: int) -> Optional[np.ndarray]:
        return self._levels.get(level)

    def read_region(self, location, level, size):
        from PIL import Image
        ds = self.level_downsamples[level]
        x, y = int(location[0] / ds), int(location[1] / ds)
        w, h = size
        src = self._levels[level]
        out = np.zeros((h, w, 4), np.uint8)
        out[..., 3] = 255
        ys, xs = max(y, 0), max(x, 0)
        ye, xe = min(y + h, src.shape[0]), min(x + w, src.shape[1])
        if ye > ys and xe > xs:
            out[ys - y:ye - y, xs - x:xe - x, :3] = src[ys:ye, xs:xe]
        return Image.fromarray(out, "RGBA")


def _level_raster(scan, level: int, engine=None):
    """Scan-level raster u8 [ih, iw, 3] of an OpenSlide-like object (numpy), or — for an on-disk ``capi.TiffSlide`` — a
    CUDA tensor decoded by nvJPEG straight into HBM (``wsi_tiff_read_rows``; needs ``engine``)."""
    if isinstance(scan, capi.TiffSlide):
        if engine is None:
            raise RuntimeError("a TiffSlide is decoded on the GPU (nvJPEG): pass engine=capi.Context(...) — no CPU decode path")
        return scan.read_level(engine, level)
    if hasattr(scan, "level_array"):
        a = scan.level_array(level)
        if a is not None:
            return np.ascontiguousarray(a[..., :3], dtype=np.uint8)
    img = scan.read_region((0, 0), level, scan.level_dimensions[level]).convert("RGB")
    return np.ascontiguousarray(np.asarray(img), dtype=np.uint8)


class Dataset_wsi:
    """One slide: the tile plan (``datalist``: list of (x, y) in scan-level pixels, reference
    order) and lazy access to the scan-level raster."""

    def __init__(self, scan, params, mask: Optional[np.ndarray], scan_level: int = 2, engine=None):
        """engine: a capi.Context — the foreground mask (find_nuclei, utils/preprocessing.py:74-110) and the foreground
        test of the tile plan run on the GPU (wsi_find_nuclei / wsi_plan_tiles_gpu).  With engine None a mask must be
        given (the reference's cached mask PNG, utils/dataset.py:131-134) and the host C++ planner enumerates the tiles;
        there is no CPU implementation of find_nuclei in this package."""
        self.scan, self.params, self.scan_level, self._engine = scan, params, scan_level, engine
        self.datalist, self.tiles = [], np.zeros((0, 2), np.int32)
        self.mask = mask
        if len(scan.level_dimensions) - 1 < scan_level:     # utils/dataset.py:123-124: slide silently skipped
            return
        self.params.iw, self.params.ih = scan.level_dimensions[scan_level]
        if self.mask is None:                                # :131-134
            thumb = _level_raster(scan, 2, engine)
            if engine is None:
                raise RuntimeError("no foreground mask for this slide and no engine: find_nuclei runs on the GPU only "
                                   "(pass engine=capi.Context(...) or a cached mask) — no CPU fallback")
            self.mask = engine.find_nuclei(thumb)
        self.m = scan.level_downsamples[scan_level] / scan.level_downsamples[2]
        plan = engine.plan_tiles if engine is not None else capi.plan_tiles
        self.tiles = plan(self.params.ih, self.params.iw, self.params.ph, self.params.pw, self.params.sh, self.params.sw, self.mask, self.m)
        self.datalist = [tuple(int(v) for v in t) for t in self.tiles]
        self._raster = None

    def __len__(self):
        return len(self.datalist)

    def raster(self) -> np.ndarray:
        if self._raster is None:
            self._raster = _level_raster(self.scan, self.scan_level, self._engine)
        return self._raster


class Dataset_wsis:
    """All validation slides (utils/dataset.py:83-107).  ``svs_pth`` is either a directory that is
    globbed for ``Case*/*.svs`` (needs ``openslide``) or a mapping ``key -> OpenSlide-like``
    (e.g. ``ArraySlide``); ``masks`` optionally maps key -> level-2 foreground mask u8 {0,1}
    (the reference caches these as PNGs under ``args.wsi_mask_pth``)."""

    def __init__(self, svs_pth, params, bs: int = 30, scan_level: int = 2, masks: Optional[Mapping] = None,
                 wsi_mask_pth: Optional[str] = None, engine=None):
        self.params = DotDict(params)
        self.wsis = {}
        self.scan_level = scan_level
        self.bs = bs
        if isinstance(svs_pth, Mapping):
            scans = {k: (v, k) for k, v in svs_pth.items()}
        else:
            # the reference opens these with openslide (utils/dataset.py:90-96); here the TIFF / SVS container is parsed by the
            # library and the JPEG tiles are decoded on the GPU (capi.TiffSlide, wsi_tiff_*)
            scans = {os.path.basename(p): (capi.TiffSlide(p), p) for p in sorted(glob.glob(f"{svs_pth}/Case*/*.svs"))}
        for key, (scan, path) in scans.items():
            mask = None if masks is None else masks.get(key)
            msk_pth = f"{wsi_mask_pth}/{key}.png" if wsi_mask_pth else None
            if mask is None and msk_pth and os.path.exists(msk_pth):
                from PIL import Image
                mask = np.asarray(Image.open(msk_pth).convert("L"))
            itr = Dataset_wsi(scan, DotDict(self.params), mask, scan_level, engine=engine)
            if len(itr) > 0:                                  # GenerateIterator_wsi returns None for empty slides (:198-201)
                self.params.iw, self.params.ih = itr.params.iw, itr.params.ih
                self.wsis[key] = {"iterator": itr, "wsipath": path, "scan": scan, "maskpath": msk_pth, "mask": itr.mask}
