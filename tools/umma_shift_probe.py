"""Hardware probe: does a SWIZZLE_128B K-major A operand work when it starts at an arbitrary 128-byte row of a
TMA-written halo tile (descriptor base_offset) and its 8-row groups are `pitch` rows apart?  See conv_igemm.cu."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from wsi_segmentation_pipeline_b200 import capi  # noqa: E402

ctx = capi.Context(0)
L = capi.lib()
g = torch.Generator().manual_seed(0)
for pitch in (16, 10, 12):
    A = torch.randint(-4, 5, (18, pitch, 64), generator=g).float()
    B = torch.randint(-3, 4, (16, 64), generator=g).float()
    Ad, Bd = A.bfloat16().cuda().contiguous(), B.bfloat16().cuda().contiguous()
    for use_bo in (1, 0):
        ok = []
        for r in range(3):
            for s in range(3):
                D = torch.zeros(128, 16, device="cuda")
                rc = L.wsi_debug_umma_shift(ctx._h, C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()), r, s, pitch, use_bo,
                                            C.c_void_p(D.data_ptr()), None)
                assert rc == 0, L.wsi_last_error(ctx._h)
                win = A[r:r + 16, s:s + 8, :].reshape(128, 64)           # m = y * 8 + x
                ref = win @ B.t()
                err = (D.cpu() - ref).abs().max().item()
                ok.append(err == 0)
        print(f"pitch {pitch:2d} base_offset {'on ' if use_bo else 'off'}: exact for (r,s) =",
              [(i // 3, i % 3) for i, v in enumerate(ok) if v], "| wrong for", [(i // 3, i % 3) for i, v in enumerate(ok) if not v])
