"""ctypes binding of libwsi_b200.so (include/wsi_b200.h) — the only way the Python host reaches
the CUDA path.  There is NO CPU fallback: a missing library or a missing CUDA device raises.

Host-only entry points (tile planner, band partition) work without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Mapping, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WSI_B200_LIB") or os.path.join(HERE, "libwsi_b200.so")

WSI_OK = 0
ERR_NAMES = {-1: "WSI_ERR_INVALID", -2: "WSI_ERR_CUDA", -3: "WSI_ERR_NOMODEL", -4: "WSI_ERR_DEGENERATE",
             -5: "WSI_ERR_UNSUPPORTED", -6: "WSI_ERR_NOMEM"}
ARCH_RESNET18, ARCH_UNET_R18 = 0, 1
HEAD_SEG, HEAD_CLS, HEAD_REG, HEAD_FEATURES = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
PRECISION_BF16, PRECISION_FP32 = 0, 1
STAGES = ("gather", "stem", "maxpool", "conv", "head", "stitch", "finalise", "h2d", "d2h")

# every symbol include/wsi_b200.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "wsi_ctx_create", "wsi_ctx_destroy", "wsi_last_error", "wsi_version", "wsi_set_option",
    "wsi_set_class_probs", "wsi_kernel_launches", "wsi_model_load", "wsi_plan_tiles", "wsi_free",
    "wsi_band_partition", "wsi_band_tiles", "wsi_run_slide", "wsi_forward_batch", "wsi_forward_tiles",
    "wsi_synth_slide", "wsi_debug_conv", "wsi_debug_gather", "wsi_debug_stem", "wsi_debug_maxpool",
    "wsi_stage_stats", "wsi_stage_reset", "wsi_resize_argmax",
    "wsi_find_nuclei", "wsi_plan_tiles_gpu", "wsi_forward_patches",
    "wsi_forward_batch_tta", "wsi_debug_umma_shift", "wsi_check", "wsi_debug_conv_f32", "wsi_op_stats",
    "wsi_ipc_alloc", "wsi_ipc_open", "wsi_ipc_close", "wsi_ipc_free",
    "wsi_morph", "wsi_tumor_bed", "wsi_overlay", "wsi_hull_rows", "wsi_host_register", "wsi_host_unregister",
    "wsi_tiff_open", "wsi_tiff_close", "wsi_tiff_last_error", "wsi_tiff_levels", "wsi_tiff_level_info", "wsi_tiff_unit_stream",
    "wsi_tiff_read_rows", "wsi_resample_ksize", "wsi_resample_coeffs",
)


class WsiError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(status, status)}: {msg}")
        self.status = status


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class SlideDesc(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("row_stride", C.c_int64), ("rgb_mem", C.c_int32),
                ("ih", C.c_int64), ("iw", C.c_int64), ("row0", C.c_int64), ("rows", C.c_int64),
                ("ph", C.c_int32), ("pw", C.c_int32), ("m", C.c_double),
                ("H2", C.c_int64), ("W2", C.c_int64), ("own0", C.c_int64), ("own1", C.c_int64),
                ("mask", C.c_void_p), ("mask_mem", C.c_int32), ("resize", C.c_int32)]


class OutDesc(C.Structure):
    _fields_ = [("mem", C.c_int32), ("classes", C.c_void_p), ("heatmap", C.c_void_p), ("canvas", C.c_void_p),
                ("probs", C.c_void_p), ("counts", C.c_void_p), ("tile_logits", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    """Load libwsi_b200.so or fail loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing — build it with `python -m wsi_segmentation_pipeline_b200.build` "
            "(or __graft_entry__.build()).  This package has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "wsi_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "wsi_ctx_destroy": (C.c_int, [vp]),
        "wsi_last_error": (C.c_char_p, [vp]),
        "wsi_version": (C.c_char_p, []),
        "wsi_set_option": (C.c_int, [vp, C.c_char_p, i64]),
        "wsi_set_class_probs": (C.c_int, [vp, C.POINTER(C.c_float), C.c_int]),
        "wsi_kernel_launches": (i64, [vp]),
        "wsi_model_load": (C.c_int, [vp, C.c_int, C.POINTER(TensorDesc), C.c_int, C.c_int]),
        "wsi_plan_tiles": (C.c_int, [i64, i64, i32, i32, i32, i32, vp, i64, i64, dbl, C.POINTER(C.POINTER(i32)), C.POINTER(i64)]),
        "wsi_free": (None, [vp]),
        "wsi_band_partition": (C.c_int, [i64, i32, i32, i32, C.POINTER(i64)]),
        "wsi_band_tiles": (C.c_int, [vp, i64, i32, dbl, i64, i64, C.POINTER(C.POINTER(i64)), C.POINTER(i64)]),
        "wsi_run_slide": (C.c_int, [vp, C.POINTER(SlideDesc), vp, i64, C.c_int, C.POINTER(OutDesc), vp]),
        "wsi_forward_batch": (C.c_int, [vp, vp, i64, i32, i32, C.c_int, vp, C.c_int, vp]),
        "wsi_forward_tiles": (C.c_int, [vp, C.POINTER(SlideDesc), vp, i64, C.c_int, vp, C.c_int, vp]),
        "wsi_synth_slide": (C.c_int, [vp, i64, i64, C.c_uint32, i64, i64, vp, vp, i64, vp, vp]),
        "wsi_debug_conv": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                     vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp]),
        "wsi_debug_conv_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp]),
        "wsi_check": (C.c_int, [vp, vp]),
        "wsi_tiff_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
        "wsi_tiff_close": (C.c_int, [vp]),
        "wsi_tiff_last_error": (C.c_char_p, [vp]),
        "wsi_tiff_levels": (C.c_int, [vp]),
        "wsi_tiff_level_info": (C.c_int, [vp, C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "wsi_tiff_unit_stream": (C.c_int, [vp, C.c_int, i64, vp, i64, C.POINTER(i64)]),
        "wsi_tiff_read_rows": (C.c_int, [vp, vp, C.c_int, i64, i64, vp, i64, vp]),
        "wsi_host_register": (C.c_int, [vp, i64]),
        "wsi_host_unregister": (C.c_int, [vp]),
        "wsi_morph": (C.c_int, [vp, vp, i64, i64, C.c_int, C.c_int, vp, C.c_int, vp]),
        "wsi_tumor_bed": (C.c_int, [vp, vp, i64, i64, vp, C.c_int, C.c_int, vp, vp, vp, C.POINTER(i64), C.c_int, vp]),
        "wsi_overlay": (C.c_int, [vp, vp, vp, i64, i64, C.c_int, vp, vp, vp, vp, C.c_int, vp]),
        "wsi_hull_rows": (C.c_int, [vp, vp, i64, vp, vp]),
        "wsi_ipc_alloc": (C.c_int, [vp, i64, C.POINTER(vp), C.c_char_p]),
        "wsi_ipc_open": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "wsi_ipc_close": (C.c_int, [vp, vp]),
        "wsi_ipc_free": (C.c_int, [vp, vp]),
        "wsi_op_stats": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl),
                                   C.POINTER(i64)]),
        "wsi_debug_gather": (C.c_int, [vp, C.POINTER(SlideDesc), vp, C.c_int, vp, vp, vp]),
        "wsi_debug_stem": (C.c_int, [vp, C.POINTER(SlideDesc), vp, C.c_int, vp, vp, vp, vp, vp]),
        "wsi_debug_maxpool": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
        "wsi_stage_stats": (C.c_int, [vp, C.c_char_p, C.POINTER(dbl), C.POINTER(i64), C.POINTER(dbl)]),
        "wsi_stage_reset": (C.c_int, [vp]),
        "wsi_resize_argmax": (C.c_int, [vp, vp, i64, i64, i64, i64, vp, vp, C.c_int, vp]),
        "wsi_debug_umma_shift": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
        "wsi_forward_batch_tta": (C.c_int, [vp, vp, i64, i32, i32, C.c_int, vp, C.c_int, vp]),
        "wsi_forward_patches": (C.c_int, [vp, vp, i64, i32, i32, i32, vp, vp, C.c_int, vp]),
        "wsi_find_nuclei": (C.c_int, [vp, vp, i64, C.c_int, i64, i64, dbl, vp, C.c_int, vp]),
        "wsi_resample_ksize": (C.c_int, [i32, i32]),
        "wsi_resample_coeffs": (C.c_int, [i32, i32, vp, vp]),
        "wsi_plan_tiles_gpu": (C.c_int, [vp, i64, i64, i32, i32, i32, i32, vp, C.c_int, i64, i64, dbl, C.POINTER(C.POINTER(i32)),
                                         C.POINTER(i64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(status: int, ctx=None):
    if status != WSI_OK:
        msg = lib().wsi_last_error(ctx)
        raise WsiError(status, (msg or b"").decode("utf-8", "replace"))


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------
# host-only: planner and band partition
# ------------------------------------------------------------------------------------------
def plan_tiles(ih, iw, ph, pw, sh, sw, mask: Optional[np.ndarray] = None, m: float = 1.0) -> np.ndarray:
    """int32 [T,2] (x,y) in the reference's enumeration order (utils/dataset.py:143-166)."""
    L = lib()
    xy = C.POINTER(C.c_int32)()
    n = C.c_int64(0)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        mh, mw = mask.shape
    else:
        mh = mw = 0
    _check(L.wsi_plan_tiles(ih, iw, ph, pw, sh, sw, _np_ptr(mask), mh, mw, float(m), C.byref(xy), C.byref(n)))
    try:
        out = np.ctypeslib.as_array(xy, shape=(max(n.value, 1) * 2,))[: n.value * 2].copy().reshape(-1, 2)
    finally:
        L.wsi_free(xy)
    return out


def resample_coeffs(in_size: int, out_size: int):
    """One axis of the scan_resize tile resize (PIL.Image.resize default = antialiased bicubic, utils/dataset.py:180-181):
    (bounds int32 [out,2] = (first input sample, count), kk int32 [out,ksize] 22-bit fixed-point weights)."""
    L = lib()
    ks = L.wsi_resample_ksize(int(in_size), int(out_size))
    if ks <= 0:
        raise WsiError(-1, "bad resample sizes")
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks), np.int32)
    _check(L.wsi_resample_coeffs(int(in_size), int(out_size), _np_ptr(bounds), _np_ptr(kk)))
    return bounds, kk


def band_partition(ih, ph, sh, nranks) -> np.ndarray:
    """int64 [nranks,4]: own0, own1 (canvas rows owned), row0, row1 (raster rows needed)."""
    bands = (C.c_int64 * (4 * nranks))()
    _check(lib().wsi_band_partition(ih, ph, sh, nranks, bands))
    return np.array(bands, dtype=np.int64).reshape(nranks, 4)


def band_tiles(xy: np.ndarray, ph, m, own0, own1) -> np.ndarray:
    xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
    idx = C.POINTER(C.c_int64)()
    n = C.c_int64(0)
    L = lib()
    _check(L.wsi_band_tiles(_np_ptr(xy), xy.shape[0], ph, float(m), own0, own1, C.byref(idx), C.byref(n)))
    try:
        out = np.ctypeslib.as_array(idx, shape=(max(n.value, 1),))[: n.value].copy()
    finally:
        L.wsi_free(idx)
    return out


def host_register(ptr: int, nbytes: int) -> bool:
    """Page-lock caller-owned host memory; False when the platform refuses (nothing stays locked)."""
    return lib().wsi_host_register(C.c_void_p(ptr), int(nbytes)) == WSI_OK


def host_unregister(ptr: int):
    lib().wsi_host_unregister(C.c_void_p(ptr))


def hull_rows(xmin: np.ndarray, xmax: np.ndarray):
    """Host part of convex_hull_image: per-row first / last set column (-1 for empty rows) -> inclusive hull range per row."""
    xmin = np.ascontiguousarray(xmin, dtype=np.int32)
    xmax = np.ascontiguousarray(xmax, dtype=np.int32)
    H = xmin.shape[0]
    xl, xr = np.empty(H, np.int32), np.empty(H, np.int32)
    _check(lib().wsi_hull_rows(_np_ptr(xmin), _np_ptr(xmax), H, _np_ptr(xl), _np_ptr(xr)))
    return xl, xr


MORPH_ERODE, MORPH_DILATE, MORPH_OPEN, MORPH_CLOSE = 0, 1, 2, 3
OVERLAY_HEAT, OVERLAY_BED = 0, 1


def rule_lut(rule) -> np.ndarray:
    """A per-level threshold rule of the reference, evaluated by numpy for the 256 u8 levels exactly as the reference
    evaluates it per pixel (e.g. ``lambda v: v >= 0.99 * 255``) -> u8 [256] of {0, 1}."""
    return np.ascontiguousarray(np.asarray(rule(np.arange(256, dtype=np.uint8))).astype(np.uint8))


# ------------------------------------------------------------------------------------------
# device context
# ------------------------------------------------------------------------------------------
def _ptr_and_mem(t):
    """(pointer, mem_kind, keepalive) for a numpy array or a torch tensor (CPU or CUDA)."""
    if t is None:
        return None, MEM_HOST, None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return C.c_void_p(t.ctypes.data), MEM_HOST, t
    import torch
    assert isinstance(t, torch.Tensor) and t.is_contiguous()
    return C.c_void_p(t.data_ptr()), (MEM_DEVICE if t.is_cuda else MEM_HOST), t


def _stream_ptr(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class Context:
    """One wsi_ctx: bound to one CUDA device, not thread-safe (one per GPU / process)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._lib = lib()
        _check(self._lib.wsi_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self.arch = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.wsi_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- knobs -------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        _check(self._lib.wsi_set_option(self._h, key.encode(), int(value)), self._h)

    def set_precision(self, precision: int):
        """PRECISION_BF16 (throughput) or PRECISION_FP32 (fp32 emulated on the bf16 tensor cores; include/wsi_b200.h)."""
        self.set_option("precision", int(precision))

    def check(self, stream=None):
        """Synchronise and raise if an earlier asynchronous (device-output) call failed on the device."""
        _check(self._lib.wsi_check(self._h, _stream_ptr(stream)), self._h)

    # ---- peer-mapped result buffer (multi-GPU) ---------------------------------------------
    def ipc_alloc(self, nbytes: int):
        """Device buffer on this GPU + the 64-byte CUDA IPC handle other processes open it with."""
        ptr, h = C.c_void_p(), C.create_string_buffer(64)
        _check(self._lib.wsi_ipc_alloc(self._h, int(nbytes), C.byref(ptr), h), self._h)
        return int(ptr.value), bytes(h.raw)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        _check(self._lib.wsi_ipc_open(self._h, C.create_string_buffer(handle, 64), C.byref(ptr)), self._h)
        return int(ptr.value)

    def ipc_close(self, ptr: int):
        _check(self._lib.wsi_ipc_close(self._h, C.c_void_p(ptr)), self._h)

    def ipc_free(self, ptr: int):
        _check(self._lib.wsi_ipc_free(self._h, C.c_void_p(ptr)), self._h)

    def set_class_probs(self, probs):
        arr = (C.c_float * 4)(*[float(p) for p in probs])
        _check(self._lib.wsi_set_class_probs(self._h, arr, 4), self._h)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.wsi_kernel_launches(self._h))

    def stage_stats(self) -> dict:
        out = {}
        for s in STAGES:
            ms, n, w = C.c_double(0), C.c_int64(0), C.c_double(0)
            _check(self._lib.wsi_stage_stats(self._h, s.encode(), C.byref(ms), C.byref(n), C.byref(w)), self._h)
            out[s] = {"ms": ms.value, "launches": n.value, "work": w.value}
        return out

    def op_stats(self) -> list:
        """Per-conv evidence of the current plan (option op_trace = 1): [{desc, kernel, ms, flops, bytes, count}]."""
        out, i = [], 0
        while True:
            d, k = C.create_string_buffer(200), C.create_string_buffer(100)
            ms, fl, by, n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
            if self._lib.wsi_op_stats(self._h, i, d, 200, k, 100, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)) != WSI_OK:
                break
            out.append({"desc": d.value.decode(), "kernel": k.value.decode(), "ms": ms.value, "flops": fl.value, "bytes": by.value, "count": n.value})
            i += 1
        return out

    def stage_reset(self):
        _check(self._lib.wsi_stage_reset(self._h), self._h)

    # ---- model -------------------------------------------------------------------------
    def load_state_dict(self, arch: int, sd: Mapping, num_classes: int = 4):
        """sd: name -> torch.Tensor / ndarray with the reference's checkpoint keys
        (utils/networks.py:6-10).  Integer entries (num_batches_tracked) are skipped."""
        keep, descs = [], []
        for name, v in sd.items():
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            if a.dtype.kind != "f" or a.ndim > 4:
                continue
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append((name.encode(), a))
        arr = (TensorDesc * len(keep))()
        for i, (nm, a) in enumerate(keep):
            arr[i].name = nm
            arr[i].data = a.ctypes.data
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
            descs.append(arr[i])
        _check(self._lib.wsi_model_load(self._h, int(arch), arr, len(keep), int(num_classes)), self._h)
        self.arch = int(arch)

    # ---- descriptors -------------------------------------------------------------------
    @staticmethod
    def slide_desc(raster, ih, iw, ph, pw, m=1.0, H2=None, W2=None, mask=None, row0=0, rows=0, own0=0, own1=0,
                   row_stride=None, resize=1):
        """raster: u8 [rows, iw, 3] (numpy / torch CPU / torch CUDA); mask: u8 [own rows, W2] or None.
        resize: myargs scan_resize — ph x pw = (tile_h * resize) x (tile_w * resize) windows, PIL-resized to the tile."""
        d = SlideDesc()
        p, mem, k1 = _ptr_and_mem(raster)
        d.rgb, d.rgb_mem = p, mem
        d.row_stride = int(row_stride if row_stride is not None else 3 * iw)
        d.ih, d.iw, d.row0, d.rows = int(ih), int(iw), int(row0), int(rows)
        d.ph, d.pw, d.m = int(ph), int(pw), float(m)
        d.H2 = int(H2 if H2 is not None else int(ih * m))
        d.W2 = int(W2 if W2 is not None else int(iw * m))
        d.own0, d.own1 = int(own0), int(own1)
        p, mem, k2 = _ptr_and_mem(mask)
        d.mask, d.mask_mem = p, mem
        d.resize = int(resize)
        d._keep = (k1, k2)
        return d

    def run_slide(self, slide: SlideDesc, tiles_xy: np.ndarray, head: int, *, device_out: bool = False,
                  want_canvas=False, want_probs=False, want_counts=False, want_tile_logits=False, stream=None,
                  out: Optional[dict] = None) -> dict:
        """The hot path (utils/eval.py:190-228 for one slide / band).  Returns dict with
        classes/heatmap u8 [rows, W2] (+ canvas/probs f32 [4, rows, W2], counts i32, tile_logits f32 [T,4])."""
        import torch
        tiles_xy = np.ascontiguousarray(tiles_xy, dtype=np.int32).reshape(-1, 2)
        T = tiles_xy.shape[0]
        own0 = slide.own0
        own1 = slide.own1 if (slide.own0 or slide.own1) else slide.H2
        rows, W2 = own1 - own0, slide.W2
        dev = torch.device("cuda", self.device) if device_out else torch.device("cpu")

        def buf(name, shape, dtype):
            if out is not None and name in out:
                t = out[name]
                assert tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.is_contiguous()
                return t
            if dev.type == "cpu":
                return torch.empty(shape, dtype=dtype, pin_memory=torch.cuda.is_available())
            return torch.empty(shape, dtype=dtype, device=dev)

        res = {"classes": buf("classes", (rows, W2), torch.uint8), "heatmap": buf("heatmap", (rows, W2), torch.uint8)}
        if want_canvas:
            res["canvas"] = buf("canvas", (4, rows, W2), torch.float32)
        if want_probs:
            res["probs"] = buf("probs", (4, rows, W2), torch.float32)
        if want_counts:
            res["counts"] = buf("counts", (rows, W2), torch.int32)
        if want_tile_logits:
            res["tile_logits"] = buf("tile_logits", (T, 4), torch.float32)
        o = OutDesc()
        o.mem = MEM_DEVICE if device_out else MEM_HOST
        for k in ("classes", "heatmap", "canvas", "probs", "counts", "tile_logits"):
            setattr(o, k, C.c_void_p(res[k].data_ptr()) if k in res else None)
        _check(self._lib.wsi_run_slide(self._h, C.byref(slide), _np_ptr(tiles_xy), T, int(head), C.byref(o),
                                       _stream_ptr(stream)), self._h)
        return res

    def forward_batch(self, x, head: int, stream=None):
        """x: f32 [n,3,h,w] normalised (torch CPU or CUDA).  Returns a tensor on the same device."""
        import torch
        x = x.contiguous().float()
        n, _, h, w = x.shape
        shape = {HEAD_SEG: (n, 4, h, w), HEAD_CLS: None, HEAD_REG: (n, 1), HEAD_FEATURES: (n, 512)}[head]
        if shape is None:
            shape = (n, 4)
        y = torch.empty(shape, dtype=torch.float32, device=x.device)
        mem = MEM_DEVICE if x.is_cuda else MEM_HOST
        _check(self._lib.wsi_forward_batch(self._h, C.c_void_p(x.data_ptr()), n, h, w, int(head), C.c_void_p(y.data_ptr()),
                                           mem, _stream_ptr(stream)), self._h)
        return y

    def find_nuclei(self, rgb, mu_percent: float = 0.1, device_out: bool = False, stream=None):
        """find_nuclei(mode='hsv') (utils/preprocessing.py:74-110) on the device.  rgb: u8 [H,W,3] numpy array or CUDA
        tensor; returns the u8 {0,1} mask as a numpy array, or as a CUDA tensor when device_out."""
        import torch
        if isinstance(rgb, np.ndarray):
            rgb = np.ascontiguousarray(rgb[..., :3], dtype=np.uint8)
            H, W = rgb.shape[:2]
            src, stride, mem = rgb.ctypes.data, 3 * W, MEM_HOST
        else:
            rgb = rgb.contiguous()
            assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.shape[-1] == 3
            H, W = int(rgb.shape[0]), int(rgb.shape[1])
            src, stride, mem = rgb.data_ptr(), 3 * W, MEM_DEVICE
        if device_out:
            mask = torch.empty((H, W), dtype=torch.uint8, device=torch.device("cuda", self.device))
            dst, dmem = mask.data_ptr(), MEM_DEVICE
        else:
            mask = np.empty((H, W), np.uint8)
            dst, dmem = mask.ctypes.data, MEM_HOST
        _check(self._lib.wsi_find_nuclei(self._h, C.c_void_p(src), stride, mem, H, W, float(mu_percent), C.c_void_p(dst), dmem,
                                         _stream_ptr(stream)), self._h)
        return mask

    def plan_tiles(self, ih, iw, ph, pw, sh, sw, mask, m: float = 1.0, stream=None) -> np.ndarray:
        """wsi_plan_tiles_gpu: the reference's tile list with the foreground test on the device.  mask: u8 [mh,mw] numpy
        array or CUDA tensor (required; use capi.plan_tiles for the all-foreground case)."""
        if isinstance(mask, np.ndarray):
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            ptr, mem = mask.ctypes.data, MEM_HOST
        else:
            mask = mask.contiguous()
            ptr, mem = mask.data_ptr(), MEM_DEVICE
        mh, mw = int(mask.shape[0]), int(mask.shape[1])
        xy, n = C.POINTER(C.c_int32)(), C.c_int64()
        _check(self._lib.wsi_plan_tiles_gpu(self._h, ih, iw, ph, pw, sh, sw, C.c_void_p(ptr), mem, mh, mw, float(m), C.byref(xy),
                                            C.byref(n), _stream_ptr(stream)), self._h)
        try:
            return np.ctypeslib.as_array(xy, shape=(max(n.value, 1), 2))[:n.value].copy() if n.value else np.zeros((0, 2), np.int32)
        finally:
            self._lib.wsi_free(xy)

    # ---- tumour-bed post-processing ------------------------------------------
    def _u8_like(self, ref, shape):
        import torch
        if isinstance(ref, np.ndarray):
            return np.empty(shape, np.uint8)
        return torch.empty(shape, dtype=torch.uint8, device=ref.device)

    def morph(self, src, op: int, k: int, stream=None):
        """cv2.erode / dilate / morphologyEx(OPEN | CLOSE) with np.ones((k, k)) — bit-exact.  src: u8 [H,W] numpy or CUDA tensor."""
        p, mem, _ = _ptr_and_mem(src)
        H, W = int(src.shape[0]), int(src.shape[1])
        dst = self._u8_like(src, (H, W))
        q, _, _ = _ptr_and_mem(dst)
        _check(self._lib.wsi_morph(self._h, p, H, W, int(op), int(k), q, mem, _stream_ptr(stream)), self._h)
        return dst

    def tumor_bed(self, src, rule, open_k: int = 20, dilate_k: int = 20, want=("opened", "hull", "outline"), stream=None) -> dict:
        """utils/eval.py:90-96: tb = rule(src); MORPH_OPEN(open_k); convex_hull_image; dilate(bwperim(.), dilate_k).
        src: u8 [H,W] (class mask or heatmap; numpy or CUDA tensor); rule: callable on the u8 levels (see rule_lut).
        Returns {'opened', 'hull', 'outline' (each u8 {0,1} [H,W], as requested), 'n_open': count_nonzero(opened)}."""
        p, mem, _ = _ptr_and_mem(src)
        H, W = int(src.shape[0]), int(src.shape[1])
        lut = rule_lut(rule)
        outs = {k: self._u8_like(src, (H, W)) for k in ("opened", "hull", "outline") if k in want}
        ptr = lambda k: _ptr_and_mem(outs[k])[0] if k in outs else None
        n_open = C.c_int64(0)
        _check(self._lib.wsi_tumor_bed(self._h, p, H, W, _np_ptr(lut), int(open_k), int(dilate_k), ptr("opened"), ptr("hull"), ptr("outline"),
                                       C.byref(n_open), mem, _stream_ptr(stream)), self._h)
        outs["n_open"] = int(n_open.value)
        return outs

    def overlay_heat(self, rgb, heat, rule=lambda v: v > 255 * 0.99, stream=None):
        """utils/eval.py:262-267: np.uint8(img * 0.75 + 255 * (heat > 255 * 0.99) * 0.25).  rgb u8 [H,W,3], heat u8 [H,W]."""
        p, mem, _ = _ptr_and_mem(rgb)
        q, mem2, _ = _ptr_and_mem(heat)
        assert mem == mem2
        H, W = int(heat.shape[0]), int(heat.shape[1])
        out = self._u8_like(rgb, (H, W, 3))
        lut = rule_lut(rule)
        _check(self._lib.wsi_overlay(self._h, p, q, H, W, OVERLAY_HEAT, _np_ptr(lut), None, None, _ptr_and_mem(out)[0], mem, _stream_ptr(stream)), self._h)
        return out

    def overlay_bed(self, rgb, heat, im=None, perim=None, stream=None):
        """paper_tools/overlay_tb_wsi.py:56-72: np.uint8(0.65 * img + 0.35 * (heat * im)), outline pixels 0."""
        p, mem, _ = _ptr_and_mem(rgb)
        H, W = int(heat.shape[0]), int(heat.shape[1])
        out = self._u8_like(rgb, (H, W, 3))
        _check(self._lib.wsi_overlay(self._h, p, _ptr_and_mem(heat)[0], H, W, OVERLAY_BED, None, _ptr_and_mem(im)[0], _ptr_and_mem(perim)[0],
                                     _ptr_and_mem(out)[0], mem, _stream_ptr(stream)), self._h)
        return out

    def resize_argmax(self, canvas, H2: int, W2: int, want_pred: bool = True, stream=None):
        """predict_wsis tail (utils/eval.py:66-81): canvas f32 [4,H,W] (torch CPU or CUDA) -> classes u8 [H2,W2]
        (argmax of the per-class cv2.resize) and, optionally, the resized pred f32 [4,H2,W2], on the same device."""
        import torch
        canvas = canvas.contiguous().float()
        assert canvas.dim() == 3 and canvas.shape[0] == 4
        H, W = int(canvas.shape[1]), int(canvas.shape[2])
        classes = torch.empty((H2, W2), dtype=torch.uint8, device=canvas.device)
        pred = torch.empty((4, H2, W2), dtype=torch.float32, device=canvas.device) if want_pred else None
        mem = MEM_DEVICE if canvas.is_cuda else MEM_HOST
        _check(self._lib.wsi_resize_argmax(self._h, C.c_void_p(canvas.data_ptr()), H, W, int(H2), int(W2), C.c_void_p(classes.data_ptr()),
                                           C.c_void_p(pred.data_ptr()) if want_pred else None, mem, _stream_ptr(stream)), self._h)
        return classes, pred

    def forward_batch_tta(self, x, head: int, stream=None):
        """Mean of the REG / CLS head over the 4 TTA views of predict_reg (utils/eval.py:303-334).  x: f32 [n,3,h,h]."""
        import torch
        x = x.contiguous().float()
        n, _, h, w = x.shape
        y = torch.empty((n, 1 if head == HEAD_REG else 4), dtype=torch.float32, device=x.device)
        _check(self._lib.wsi_forward_batch_tta(self._h, C.c_void_p(x.data_ptr()), n, h, w, int(head), C.c_void_p(y.data_ptr()),
                                               MEM_DEVICE if x.is_cuda else MEM_HOST, _stream_ptr(stream)), self._h)
        return y

    def forward_patches(self, xs_patch_major, B: int, P: int, stream=None):
        """resnets_shift.ResNet.forward (:189-217).  xs_patch_major: f32 [P*B,3,h,w] (torch CPU or CUDA).  Returns
        (y [P*B,4] per-patch fc0 logits, ens [B,4] ensemble head) on the same device."""
        import torch
        x = xs_patch_major.contiguous().float()
        assert x.dim() == 4 and x.shape[0] == P * B and x.shape[1] == 3
        h, w = int(x.shape[2]), int(x.shape[3])
        y = torch.empty((P * B, 4), dtype=torch.float32, device=x.device)
        ens = torch.empty((B, 4), dtype=torch.float32, device=x.device)
        _check(self._lib.wsi_forward_patches(self._h, C.c_void_p(x.data_ptr()), B, P, h, w, C.c_void_p(y.data_ptr()),
                                             C.c_void_p(ens.data_ptr()), MEM_DEVICE if x.is_cuda else MEM_HOST, _stream_ptr(stream)), self._h)
        return y, ens

    def forward_tiles(self, slide: SlideDesc, tiles_xy: np.ndarray, head: int, device_out=False, stream=None):
        import torch
        tiles_xy = np.ascontiguousarray(tiles_xy, dtype=np.int32).reshape(-1, 2)
        n = tiles_xy.shape[0]
        r = max(1, slide.resize)
        shape = {HEAD_SEG: (n, 4, slide.ph // r, slide.pw // r), HEAD_CLS: (n, 4), HEAD_REG: (n, 1), HEAD_FEATURES: (n, 512)}[head]
        y = torch.empty(shape, dtype=torch.float32, device=torch.device("cuda", self.device) if device_out else "cpu")
        _check(self._lib.wsi_forward_tiles(self._h, C.byref(slide), _np_ptr(tiles_xy), n, int(head), C.c_void_p(y.data_ptr()),
                                           MEM_DEVICE if device_out else MEM_HOST, _stream_ptr(stream)), self._h)
        return y

    def synth_slide(self, ih, iw, seed=1234, y0=0, y1=None, with_mask=False, stream=None):
        """Synthetic H&E raster rows [y0,y1) generated on the device (twin of synth.synth_slide)."""
        import torch
        from . import synth
        y1 = ih if y1 is None else y1
        dev = torch.device("cuda", self.device)
        rgb = torch.empty((y1 - y0, iw, 3), dtype=torch.uint8, device=dev)
        mask = torch.empty((y1 - y0, iw), dtype=torch.uint8, device=dev) if with_mask else None
        lut = np.ascontiguousarray(synth.synth_lut())
        _check(self._lib.wsi_synth_slide(self._h, ih, iw, seed, y0, y1, _np_ptr(lut), C.c_void_p(rgb.data_ptr()), 3 * iw,
                                         C.c_void_p(mask.data_ptr()) if with_mask else None, _stream_ptr(stream)), self._h)
        return (rgb, mask) if with_mask else rgb

    # ---- per-kernel debug entry points (tests) --------------------------------------------
    def debug_conv(self, x_nhwc, weight, stride=1, pad=1, scale=None, bias=None, res=None, relu=False, up2=False, skip=None):
        """x_nhwc: bf16 CUDA [n,h,w,cin]; weight: f32 CPU OIHW.  Returns bf16 CUDA NHWC."""
        import torch
        n, h, w, cin = x_nhwc.shape
        cout, _, k, _ = weight.shape
        hin, win = (2 * h, 2 * w) if up2 else (h, w)
        oh, ow = (hin + 2 * pad - k) // stride + 1, (win + 2 * pad - k) // stride + 1
        y = torch.empty((n, oh, ow, cout), dtype=torch.bfloat16, device=x_nhwc.device)
        wt = np.ascontiguousarray(weight.detach().cpu().numpy(), dtype=np.float32)
        sc = None if scale is None else np.ascontiguousarray(scale.detach().cpu().numpy(), dtype=np.float32)
        bi = None if bias is None else np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
        _check(self._lib.wsi_debug_conv(
            self._h, C.c_void_p(x_nhwc.data_ptr()), n, h, w, cin, _np_ptr(wt), cout, k, stride, pad, _np_ptr(sc), _np_ptr(bi),
            C.c_void_p(res.data_ptr()) if res is not None else None, int(relu), int(up2),
            C.c_void_p(skip.data_ptr()) if skip is not None else None, 0 if skip is None else skip.shape[-1],
            C.c_void_p(y.data_ptr()), _stream_ptr(None)), self._h)
        return y

    def debug_conv_f32(self, x_nhwc, weight, stride=1, pad=1, scale=None, bias=None, res=None, relu=False, up2=False, skip=None):
        """The fp32-emulated conv (PRECISION_FP32).  x_nhwc / res / skip: f32 CUDA NHWC; weight f32 CPU OIHW.  Returns f32 CUDA NHWC."""
        import torch
        n, h, w, cin = x_nhwc.shape
        cout, _, k, _ = weight.shape
        hin, win = (2 * h, 2 * w) if up2 else (h, w)
        oh, ow = (hin + 2 * pad - k) // stride + 1, (win + 2 * pad - k) // stride + 1
        y = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=x_nhwc.device)
        wt = np.ascontiguousarray(weight.detach().cpu().numpy(), dtype=np.float32)
        sc = None if scale is None else np.ascontiguousarray(scale.detach().cpu().numpy(), dtype=np.float32)
        bi = None if bias is None else np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
        _check(self._lib.wsi_debug_conv_f32(
            self._h, C.c_void_p(x_nhwc.data_ptr()), n, h, w, cin, _np_ptr(wt), cout, k, stride, pad, _np_ptr(sc), _np_ptr(bi),
            C.c_void_p(res.data_ptr()) if res is not None else None, int(relu), int(up2),
            C.c_void_p(skip.data_ptr()) if skip is not None else None, 0 if skip is None else skip.shape[-1],
            C.c_void_p(y.data_ptr()), _stream_ptr(None)), self._h)
        return y

    def debug_gather(self, slide: SlideDesc, tiles_xy, want_padded=False, want_norm=True):
        """want_norm=False runs the kernel exactly as the engine does (padded bf16 operand only: the staged fast path)."""
        import torch
        tiles_xy = np.ascontiguousarray(tiles_xy, dtype=np.int32).reshape(-1, 2)
        n = tiles_xy.shape[0]
        dev = torch.device("cuda", self.device)
        r = max(1, slide.resize)
        norm = torch.empty((n, 3, slide.ph // r, slide.pw // r), dtype=torch.float32, device=dev) if want_norm else None
        padded = torch.empty((n, slide.ph // r + 6, slide.pw // r + 8, 4), dtype=torch.bfloat16, device=dev) if want_padded else None
        _check(self._lib.wsi_debug_gather(self._h, C.byref(slide), _np_ptr(tiles_xy), n, C.c_void_p(norm.data_ptr()) if want_norm else None,
                                          C.c_void_p(padded.data_ptr()) if want_padded else None, _stream_ptr(None)), self._h)
        if not want_norm:
            return padded
        return (norm, padded) if want_padded else norm

    def debug_stem(self, slide: SlideDesc, tiles_xy, weight, scale=None, bias=None):
        import torch
        tiles_xy = np.ascontiguousarray(tiles_xy, dtype=np.int32).reshape(-1, 2)
        n = tiles_xy.shape[0]
        y = torch.empty((n, slide.ph // 2, slide.pw // 2, 64), dtype=torch.bfloat16, device=torch.device("cuda", self.device))
        wt = np.ascontiguousarray(weight.detach().cpu().numpy(), dtype=np.float32)
        sc = None if scale is None else np.ascontiguousarray(scale.detach().cpu().numpy(), dtype=np.float32)
        bi = None if bias is None else np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
        _check(self._lib.wsi_debug_stem(self._h, C.byref(slide), _np_ptr(tiles_xy), n, _np_ptr(wt), _np_ptr(sc), _np_ptr(bi),
                                        C.c_void_p(y.data_ptr()), _stream_ptr(None)), self._h)
        return y

    def debug_maxpool(self, x_nhwc):
        import torch
        n, h, w, c = x_nhwc.shape
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=torch.bfloat16, device=x_nhwc.device)
        _check(self._lib.wsi_debug_maxpool(self._h, C.c_void_p(x_nhwc.data_ptr()), n, h, w, c, C.c_void_p(y.data_ptr()),
                                           _stream_ptr(None)), self._h)
        return y


class _DevPtr:
    """Raw device memory as a __cuda_array_interface__ object (zero-copy torch.as_tensor)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_u8_tensor(ptr: int, shape, device: int):
    """A torch u8 view of device memory the library (or a peer process) owns."""
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape), device=torch.device("cuda", device))


class TiffSlide:
    """An on-disk slide (JPEG-compressed TIFF / Aperio SVS) with OpenSlide's attribute names (``level_dimensions``,
    ``level_downsamples``, ``level_count``; utils/dataset.py:121-126).  Pixels are decoded on the GPU by nvJPEG straight into a
    device raster (``read_level``); there is no CPU decode path in this package."""

    def __init__(self, path: str):
        self._lib = lib()
        self._h = C.c_void_p()
        st = self._lib.wsi_tiff_open(os.fsencode(path), C.byref(self._h))
        if st != WSI_OK:
            raise WsiError(st, (self._lib.wsi_last_error(None) or b"").decode("utf-8", "replace"))
        self.path = path
        self.level_count = int(self._lib.wsi_tiff_levels(self._h))
        self.levels = []
        for lv in range(self.level_count):
            W, H = C.c_int64(), C.c_int64()
            tw, th, comp, ph = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            self._check(self._lib.wsi_tiff_level_info(self._h, lv, C.byref(W), C.byref(H), C.byref(tw), C.byref(th), C.byref(comp), C.byref(ph)))
            self.levels.append({"W": W.value, "H": H.value, "tile_w": tw.value, "tile_h": th.value, "compression": comp.value, "photometric": ph.value})
        self.level_dimensions = tuple((lv["W"], lv["H"]) for lv in self.levels)
        w0 = self.levels[0]["W"]
        self.level_downsamples = tuple(float(w0) / lv["W"] for lv in self.levels)

    def _check(self, st):
        if st != WSI_OK:
            raise WsiError(st, (self._lib.wsi_tiff_last_error(self._h) or b"").decode("utf-8", "replace"))

    def unit_stream(self, level: int, unit: int) -> bytes:
        """Host-only: the spliced JPEG stream of tile / strip ``unit`` (what nvJPEG is handed)."""
        n = C.c_int64(0)
        self._check(self._lib.wsi_tiff_unit_stream(self._h, level, unit, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._check(self._lib.wsi_tiff_unit_stream(self._h, level, unit, buf, n.value, C.byref(n)))
        return bytes(buf.raw[:n.value])

    def read_level(self, ctx: "Context", level: int, row0: int = 0, rows: Optional[int] = None, stream=None):
        """Rows [row0, row0 + rows) of ``level`` as a CUDA u8 tensor [rows, W, 3] (nvJPEG on ``ctx``'s device)."""
        import torch
        W, H = self.level_dimensions[level]
        rows = H - row0 if rows is None else rows
        out = torch.empty((rows, W, 3), dtype=torch.uint8, device=torch.device("cuda", ctx.device))
        self._check(self._lib.wsi_tiff_read_rows(ctx._h, self._h, level, row0, rows, C.c_void_p(out.data_ptr()), 3 * W, _stream_ptr(stream)))
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.wsi_tiff_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
