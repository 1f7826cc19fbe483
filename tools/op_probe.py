"""Per-op times of one traced slide (option op_trace): which conv kernel takes what — for same-box elimination experiments
(debug-switch builds: WSI_STREAM_DBG / WSI_UP_DBG make results garbage but time the remaining roles)."""
import sys

import torch

sys.path.insert(0, ".")
from wsi_segmentation_pipeline_b200 import weights as W   # noqa: E402
from wsi_segmentation_pipeline_b200 import capi          # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    ctx = capi.Context(0)
    ctx.load_state_dict(capi.ARCH_UNET_R18, W.random_state_dict("unet", 0))
    ctx.set_option("op_trace", 1)
    rgb = ctx.synth_slide(size, size, 1234)
    tiles = capi.plan_tiles(size, size, 512, 512, 128, 128)
    sl = ctx.slide_desc(rgb, size, size, 512, 512)
    for _ in range(2):
        ctx.run_slide(sl, tiles, capi.HEAD_SEG, device_out=True)
        torch.cuda.synchronize()
    rows = ctx.op_stats()
    tot = sum(r["ms"] for r in rows)
    out = []
    for r in rows:
        if pat and not any(p in r["kernel"] + r["desc"] for p in pat.split(",")):
            continue
        per = r["ms"] / max(r["count"], 1)
        out.append(f"{r['kernel'][:34]:34s} {r['desc'][:44]:44s} {per * 1e3:8.1f} us/launch")
    print(f"tiles {len(tiles)}, traced conv total {tot:.1f} ms")
    print("\n".join(out))


if __name__ == "__main__":
    main()
