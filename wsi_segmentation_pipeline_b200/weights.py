"""Synthetic checkpoints: random-init weights with the reference's state_dict keys.

Test / bench DATA, like ``synth.py`` — not an inference path: nothing here is called by ``capi`` / ``eval`` / ``models``.
``bench.py`` (GPU arm), ``__graft_entry__.smoke()`` and the tests load these into the engine; the CPU oracle re-exports
``random_state_dict`` so that both sides of every comparison hold identical weights.

The one torch-CPU pass below only CALIBRATES BatchNorm running statistics (what training would converge to): a freshly
constructed network has identity BN statistics (resnets_shift.py:155-157), which would hide epilogue bugs, while
independent random statistics make activations grow without bound through the residual stages (summed logits of +-1000:
every comparison would measure softmax saturation, not arithmetic).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _kaiming(shape, g):
    """nn.init.kaiming_normal_(mode='fan_out', nonlinearity='relu') — resnets_shift.py:152-154."""
    fan_out = shape[0] * shape[2] * shape[3]
    return torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5


def _kaiming_uniform_fan_in(shape, g):
    """smp's decoder init (initialization.py: kaiming_uniform_, fan_in, relu)."""
    fan_in = shape[1] * shape[2] * shape[3]
    bound = (6.0 / fan_in) ** 0.5
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _xavier_uniform(shape, g):
    fan_in, fan_out = shape[1] * shape[2] * shape[3], shape[0] * shape[2] * shape[3]
    bound = (6.0 / (fan_in + fan_out)) ** 0.5
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _linear(out_f, in_f, g):
    bound = 1.0 / in_f ** 0.5
    return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound,
            (torch.rand(out_f, generator=g) * 2 - 1) * bound)


def _bn_keys(sd, p, n, g):
    sd[p + ".weight"] = torch.rand(n, generator=g) + 0.5
    sd[p + ".bias"] = torch.randn(n, generator=g) * 0.1
    sd[p + ".running_mean"] = torch.zeros(n)
    sd[p + ".running_var"] = torch.ones(n)
    sd[p + ".num_batches_tracked"] = torch.tensor(0)


def random_resnet18_trunk(g, prefix=""):
    sd = {}
    p = prefix
    sd[p + "conv1.weight"] = _kaiming((64, 3, 7, 7), g)
    _bn_keys(sd, p + "bn1", 64, g)
    cin = 64
    for li, cout in zip((1, 2, 3, 4), (64, 128, 256, 512)):
        for b in range(2):
            q = f"{p}layer{li}.{b}"
            sd[q + ".conv1.weight"] = _kaiming((cout, cin if b == 0 else cout, 3, 3), g)
            _bn_keys(sd, q + ".bn1", cout, g)
            sd[q + ".conv2.weight"] = _kaiming((cout, cout, 3, 3), g)
            _bn_keys(sd, q + ".bn2", cout, g)
            if b == 0 and li > 1:
                sd[q + ".downsample.0.weight"] = _kaiming((cout, cin, 1, 1), g)
                _bn_keys(sd, q + ".downsample.1", cout, g)
        cin = cout
    return sd


def _calib_bn(sd, p, x):
    """Set BN ``p``'s running statistics to those of its input ``x`` (what training converges to)."""
    sd[p + ".running_mean"] = x.mean(dim=(0, 2, 3)).clone()
    sd[p + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-4).clone()
    return _bn(sd, p, x)


def calibrate_bn(sd, arch: str, x: torch.Tensor):
    """One fp32 pass in network order; every BN sees the activations produced with the already
    calibrated layers before it.  arch: 'resnet18' (prefix '') or 'unet' (prefix 'encoder.')."""
    p = "" if arch == "resnet18" else "encoder."
    with torch.no_grad():
        x0 = F.relu(_calib_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3)))
        cur = F.max_pool2d(x0, 3, 2, 1)
        feats = []
        for li in (1, 2, 3, 4):
            for b in range(2):
                q = f"{p}layer{li}.{b}"
                stride = 2 if (li > 1 and b == 0) else 1
                out = F.relu(_calib_bn(sd, q + ".bn1", F.conv2d(cur, sd[q + ".conv1.weight"], None, stride, 1)))
                out = _calib_bn(sd, q + ".bn2", F.conv2d(out, sd[q + ".conv2.weight"], None, 1, 1))
                idt = cur
                if (q + ".downsample.0.weight") in sd:
                    idt = _calib_bn(sd, q + ".downsample.1", F.conv2d(cur, sd[q + ".downsample.0.weight"], None, stride, 0))
                cur = F.relu(out + idt)
            feats.append(cur)
        if arch == "unet":
            xd = feats[3]
            skips = [feats[2], feats[1], feats[0], x0, None]
            for i, skip in enumerate(skips, start=1):
                xd = F.interpolate(xd, scale_factor=2, mode="nearest")
                if skip is not None:
                    xd = torch.cat([xd, skip], dim=1)
                for j in range(2):
                    q = f"decoder.layer{i}.block.{j}.block"
                    xd = F.relu(_calib_bn(sd, q + ".1", F.conv2d(xd, sd[q + ".0.weight"], None, 1, 1)))
    return sd


def _calibration_batch(seed: int) -> torch.Tensor:
    """8 normalised 128x128 tiles of the synthetic H&E slide (integer-only generator => identical
    bytes on every machine)."""
    from . import synth
    import numpy as np
    raster = synth.synth_slide(512, 1024, 4321 + seed)
    tiles = [(x, y) for y in (0, 256) for x in (0, 256, 512, 768)]
    mean = torch.tensor((0.485, 0.456, 0.406), dtype=torch.float32).view(3, 1, 1)      # myargs.py:127-130
    std = torch.tensor((0.229, 0.224, 0.225), dtype=torch.float32).view(3, 1, 1)
    out = []
    for (x, y) in tiles:                                                                # ToTensor + Normalize (utils/preprocessing.py:209-212)
        t = torch.from_numpy(np.ascontiguousarray(raster[y:y + 128, x:x + 128])).permute(2, 0, 1).contiguous()
        out.append(t.to(torch.float32).div(255).sub_(mean).div_(std))
    return torch.stack(out)


def random_state_dict(arch: str, seed: int = 0, num_classes: int = 4, with_fc: bool = False):
    """arch 'resnet18' -> resnets_shift.ResNet keys (fc.* only when with_fc: 33.6 M params);
    arch 'unet' -> smp.Unet('resnet18') keys + classifier/regressor heads (eval_tumorbed.py:21-28).
    BatchNorm statistics are calibrated (see the section comment)."""
    g = torch.Generator().manual_seed(seed)
    if arch == "resnet18":
        sd = random_resnet18_trunk(g)
        sd["fc0.weight"], sd["fc0.bias"] = _linear(4, 512, g)
        if with_fc:
            sd["fc.0.weight"], sd["fc.0.bias"] = _linear(4096, 8192, g)
            sd["fc.2.weight"], sd["fc.2.bias"] = _linear(4, 4096, g)
        return calibrate_bn(sd, arch, _calibration_batch(seed))
    if arch == "unet":
        sd = random_resnet18_trunk(g, "encoder.")
        ins = (768, 384, 192, 128, 32)
        outs = (256, 128, 64, 32, 16)
        for i, (ci, co) in enumerate(zip(ins, outs), start=1):
            for j, c_in in enumerate((ci, co)):
                q = f"decoder.layer{i}.block.{j}.block"
                sd[q + ".0.weight"] = _kaiming_uniform_fan_in((co, c_in, 3, 3), g)
                _bn_keys(sd, q + ".1", co, g)
        sd["decoder.final_conv.weight"] = _xavier_uniform((num_classes, 16, 1, 1), g)
        sd["decoder.final_conv.bias"] = torch.randn(num_classes, generator=g) * 0.1
        sd["classifier.fc.0.weight"], sd["classifier.fc.0.bias"] = _linear(num_classes, 512, g)
        sd["regressor.fc.0.weight"], sd["regressor.fc.0.bias"] = _linear(128, 512, g)
        sd["regressor.fc.2.weight"], sd["regressor.fc.2.bias"] = _linear(1, 128, g)
        return calibrate_bn(sd, arch, _calibration_batch(seed))
    raise ValueError(arch)
