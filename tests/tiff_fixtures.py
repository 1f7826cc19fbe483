"""Test fixtures for slide ingestion: JPEG-compressed TIFFs written here (no sample slides ship with the reference).

* ``write_stripped(path, rgb, ycbcr)``: PIL + libtiff, compression 7, shared JPEGTables (tag 347), photometric RGB or YCbCr.
* ``write_tiled_pyramid(path, levels, tile)``: a minimal classic-TIFF writer for an Aperio-SVS-like file — one directory per
  pyramid level, square JPEG tiles (each a self-contained YCbCr JFIF stream, as many SVS writers store them)."""
import io
import struct

import numpy as np
from PIL import Image


def write_stripped(path, rgb: np.ndarray, ycbcr: bool = False, quality: int = 90):
    im = Image.fromarray(rgb)
    if ycbcr:
        im = im.convert("YCbCr")
    im.save(path, format="TIFF", compression="jpeg", quality=quality)


def write_tiled_pyramid(path, levels, tile: int = 240, quality: int = 90, subsampling: int = 0):
    """levels: list of u8 [H, W, 3] arrays, full resolution first."""
    blobs, dirs = [], []
    for rgb in levels:
        H, W = rgb.shape[:2]
        ty, tx = -(-H // tile), -(-W // tile)
        streams = []
        for j in range(ty):
            for i in range(tx):
                t = np.zeros((tile, tile, 3), np.uint8)
                src = rgb[j * tile:(j + 1) * tile, i * tile:(i + 1) * tile]
                t[:src.shape[0], :src.shape[1]] = src
                b = io.BytesIO()
                Image.fromarray(t).save(b, format="JPEG", quality=quality, subsampling=subsampling)
                streams.append(b.getvalue())
        dirs.append((W, H, streams))
    out = bytearray(b"II" + struct.pack("<HI", 42, 0))       # first IFD offset patched below

    def align():
        while len(out) % 2:
            out.append(0)

    ifd_pos = []
    for (W, H, streams) in dirs:
        offs = []
        for s in streams:
            align()
            offs.append(len(out))
            out.extend(s)
        align()
        off_arr = len(out)
        out.extend(struct.pack(f"<{len(offs)}I", *offs))
        cnt_arr = len(out)
        out.extend(struct.pack(f"<{len(streams)}I", *[len(s) for s in streams]))
        bits_arr = len(out)
        out.extend(struct.pack("<3H", 8, 8, 8))
        align()
        n = len(streams)
        entries = [
            (256, 4, 1, W), (257, 4, 1, H), (258, 3, 3, bits_arr), (259, 3, 1, 7), (262, 3, 1, 6), (277, 3, 1, 3), (284, 3, 1, 1),
            (322, 4, 1, tile), (323, 4, 1, tile),
            (324, 4, n, off_arr if n > 1 else offs[0]), (325, 4, n, cnt_arr if n > 1 else len(streams[0])),
        ]
        ifd_pos.append(len(out))
        out.extend(struct.pack("<H", len(entries)))
        for tag, typ, cnt, val in sorted(entries):
            out.extend(struct.pack("<HHII", tag, typ, cnt, val))
        out.extend(struct.pack("<I", 0))                       # next IFD, patched below
    struct.pack_into("<I", out, 4, ifd_pos[0])
    for k in range(len(ifd_pos) - 1):
        struct.pack_into("<I", out, ifd_pos[k] + 2 + 12 * 11, ifd_pos[k + 1])
    with open(path, "wb") as f:
        f.write(out)
