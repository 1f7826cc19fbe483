"""Quick device-side timing of the hot path on a synthetic slide (not the bench; a probe)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from wsi_segmentation_pipeline_b200 import weights as O   # noqa: E402  (synthetic checkpoint)
from wsi_segmentation_pipeline_b200 import capi        # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    arch = sys.argv[4] if len(sys.argv) > 4 else "unet"
    batch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    resize = int(sys.argv[6]) if len(sys.argv) > 6 else 1        # scan_resize: windows of tile * resize pixels, stride as given
    ctx = capi.Context(0)
    sd = O.random_state_dict(arch, 0)
    ctx.load_state_dict(capi.ARCH_UNET_R18 if arch == "unet" else capi.ARCH_RESNET18, sd)
    head = capi.HEAD_SEG if arch == "unet" else capi.HEAD_CLS
    ctx.set_option("stage_timing", 1)
    ctx.set_option("batch_tiles", batch)
    rgb = ctx.synth_slide(size, size, 1234)
    tiles = capi.plan_tiles(size, size, tile * resize, tile * resize, stride, stride)
    sl = ctx.slide_desc(rgb, size, size, tile * resize, tile * resize, resize=resize)
    for it in range(3):
        ctx.stage_reset()
        torch.cuda.synchronize()
        t0 = time.time()
        r = ctx.run_slide(sl, tiles, head, device_out=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = ctx.stage_stats()
        print(f"iter {it}: {dt*1e3:.1f} ms wall, {size*size/1e6/dt:.1f} slide-Mpx/s, tiles {len(tiles)}")
        for k, v in st.items():
            if v["launches"] or v["ms"]:
                rate = v["work"] / (v["ms"] * 1e-3) if v["ms"] > 0 else 0
                unit = "TFLOP/s" if k in ("conv", "stem") else "GB/s"
                scale = 1e12 if k in ("conv", "stem") else 1e9
                print(f"   {k:9s} {v['ms']:9.2f} ms  {v['launches']:6d} launches  {rate/scale:9.1f} {unit}")
    print(json.dumps({"classes_hist": np.bincount(r["classes"].cpu().numpy().ravel(), minlength=4).tolist()}))


if __name__ == "__main__":
    main()
