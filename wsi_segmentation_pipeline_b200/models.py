"""Drop-in ``torch.nn.Module`` surface of the models on the hot path.

The modules below hold parameters under exactly the reference's ``state_dict`` keys so that
``model.load_state_dict(torch.load(path)['state_dict'])`` (utils/networks.py:6-10) works
unchanged, but their ``forward`` runs on libwsi_b200 (tcgen05 implicit-GEMM convolutions),
never on PyTorch/cuDNN.  There is no CPU fallback: calling them without a CUDA device raises.

* ``resnet18()``      -> resnets_shift.ResNet(BasicBlock,[2,2,2,2]) (resnets_shift.py:111-225):
                         ``forward(xs[B,P,3,H,W]) -> (cat(y_list,0)[P*B,4], fc(features)[B,4])``
* ``unet_resnet18()`` -> smp.Unet('resnet18', classes=C) + ``.classifier`` / ``.regressor`` heads
                         as eval_tumorbed.py:21-28 builds them: ``.encoder(x) -> [x4..x0]``,
                         ``.decoder(feats) -> [B,C,H,W]``, ``.classifier(x4) -> [B,C]``,
                         ``.regressor(x4) -> [B,1]``, ``model(x) == decoder(encoder(x))``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import capi

HR_NUM_SAMPLES = 16     # dhr.HR_NUM_CNT_SAMPLES + dhr.HR_NUM_PERIM_SAMPLES (utils/dataset_hr.py:14-18)


# ------------------------------------------------------------------------------------------
# parameter containers (same module tree => same state_dict keys as the reference)
# ------------------------------------------------------------------------------------------
class _BasicBlockParams(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _TrunkParams(nn.Module):
    """conv1/bn1/layer1..4 of resnets_shift.ResNet (:122-130) == smp's ResNetEncoder."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, cout in zip((1, 2, 3, 4), (64, 128, 256, 512)):
            setattr(self, f"layer{li}", nn.Sequential(_BasicBlockParams(cin, cout, 1 if li == 1 else 2),
                                                      _BasicBlockParams(cout, cout, 1)))
            cin = cout
        for m in self.modules():       # resnets_shift.py:152-157
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _EngineBound:
    """Mixin: lazily creates the wsi_ctx for the module's device and (re)uploads the weights
    whenever they changed (load_state_dict, .cuda(), in-place edits bump tensor versions)."""

    _arch = None

    def _engine(self) -> capi.Context:
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("wsi_b200 models run on CUDA only (call .cuda()); there is no CPU fallback")
        dev = p.device.index if p.device.index is not None else torch.cuda.current_device()
        ctx = self.__dict__.get("_ctx")
        if ctx is None or ctx.device != dev:
            ctx = capi.Context(dev)
            self.__dict__["_ctx"] = ctx
            self.__dict__["_sig"] = None
        sig = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self.__dict__.get("_sig") != sig:
            ctx.load_state_dict(self._arch, self.state_dict(), num_classes=getattr(self, "num_classes", 4))
            self.__dict__["_sig"] = sig
        return ctx

    @property
    def ctx(self) -> capi.Context:
        return self._engine()


class _Fn(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.__dict__["_fn"] = fn

    def forward(self, *a):
        return self.__dict__["_fn"](*a)


class EncodedBatch(list):
    """What ``.encoder(x)`` returns: a list ``[x4, x3, x2, x1, x0]`` whose entries are produced
    lazily, and which remembers the input batch so that ``.decoder`` / ``.classifier`` /
    ``.regressor`` run the fused network in one pass instead of re-entering from feature maps."""

    def __init__(self, owner, x):
        super().__init__([None] * 5)
        self.owner, self.x = owner, x

    def __getitem__(self, i):
        v = super().__getitem__(i)
        if v is None:
            if i not in (0, -5):
                raise NotImplementedError("only encoding[0] (x4) is materialised as a tensor; pass the whole "
                                          "EncodedBatch to .decoder for dense prediction")
            v = _Handle(self)
            super().__setitem__(i, v)
        return v


class _Handle:
    """encoding[0]: stands for x4 of a batch; heads consume it through the engine."""

    def __init__(self, enc):
        self.enc = enc


# ------------------------------------------------------------------------------------------
# resnets_shift.ResNet drop-in
# ------------------------------------------------------------------------------------------
class ResNet(_EngineBound, _TrunkParams):
    _arch = capi.ARCH_RESNET18

    def __init__(self, num_classes=1000):
        super().__init__()
        n = 512 * HR_NUM_SAMPLES
        self.fc = nn.Sequential(nn.Linear(n, n // 2), nn.ReLU(True), nn.Linear(n // 2, 4))   # :133-139
        self.fc0 = nn.Linear(512, 4)                                                          # :140
        self.fc1 = nn.Sequential(nn.Linear(512, 16), nn.ReLU(True))                           # :143-146
        self.fc2 = nn.Sequential(nn.Linear(16 * HR_NUM_SAMPLES, 4))                           # :147-150

    def forward(self, xs):
        """resnets_shift.py:189-217, one call into the engine (wsi_forward_patches): the conv trunk, average pool and the
        per-patch head fc0 for all P*B patches at once (patch-major order, as the reference concatenates), then the
        ensemble head ``fc`` (8192 -> 4096 -> 4 at P = 16: 33.6 M fp32 parameters, weight-bandwidth bound) on the pooled
        features.  Returns (cat(y_list, 0) [P*B, 4], fc(features) [B, 4])."""
        ctx = self._engine()
        B, P = xs.shape[:2]
        flat = xs.transpose(0, 1).reshape(P * B, *xs.shape[2:]).contiguous()
        return ctx.forward_patches(flat, B, P)

    # the predict_tumorbed protocol (SURVEY 8b config-1 adapter)
    @property
    def encoder(self):
        return _Fn(lambda x: EncodedBatch(self, x))

    @property
    def classifier(self):
        return _Fn(lambda h: self._engine().forward_batch(h.enc.x, capi.HEAD_CLS))

    @property
    def decoder(self):
        return nn.Identity()

    @property
    def regressor(self):
        return nn.Identity()


def resnet18(pretrained=False, progress=True, **kwargs):
    """resnets_shift.resnet18 (:219-225); ``pretrained`` needs network access and is refused."""
    if pretrained:
        raise ValueError("pretrained ImageNet weights are not available offline; load a checkpoint instead")
    return ResNet(**kwargs)


# ------------------------------------------------------------------------------------------
# smp.Unet('resnet18') + Classifier/Regressor drop-in
# ------------------------------------------------------------------------------------------
class _ConvBnRelu(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.block = nn.Sequential(_ConvBnRelu(cin, cout), _ConvBnRelu(cout, cout))


class _DecoderParams(nn.Module):
    def __init__(self, classes):
        super().__init__()
        for i, (ci, co) in enumerate(zip((768, 384, 192, 128, 32), (256, 128, 64, 32, 16)), start=1):
            setattr(self, f"layer{i}", _DecoderBlock(ci, co))
        self.final_conv = nn.Conv2d(16, classes, 1)


class _HeadParams(nn.Module):
    def __init__(self, *dims):
        super().__init__()
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:
                layers.append(nn.ReLU(True))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Sequential(*layers)


class _Encoder(_TrunkParams):
    out_shapes = (512, 256, 128, 64, 64)        # smp 0.0.x API used at eval_tumorbed.py:27

    def forward(self, x):
        return EncodedBatch(self.__dict__["_owner"], x)


class _Decoder(_DecoderParams):
    def forward(self, enc):
        if not isinstance(enc, EncodedBatch):
            raise TypeError(".decoder expects what .encoder returned")
        return enc.owner._engine().forward_batch(enc.x, capi.HEAD_SEG)


class _Classifier(_HeadParams):
    def forward(self, h):
        return h.enc.owner._engine().forward_batch(h.enc.x, capi.HEAD_CLS)


class _Regressor(_HeadParams):
    def forward(self, h):
        return h.enc.owner._engine().forward_batch(h.enc.x, capi.HEAD_REG)


class Unet(_EngineBound, nn.Module):
    """smp.Unet('resnet18', classes=C) with the reference's extra heads."""
    _arch = capi.ARCH_UNET_R18

    def __init__(self, encoder_name="resnet18", encoder_weights=None, classes=4, activation=None):
        super().__init__()
        if encoder_name != "resnet18":
            raise ValueError("only the resnet18 encoder is on the hot path (myargs.py arch_encoder)")
        if encoder_weights not in (None, "none"):
            raise ValueError("ImageNet weights need network access; load a checkpoint instead")
        self.num_classes = classes
        self.encoder = _Encoder()
        self.decoder = _Decoder(classes)
        self.classifier = _Classifier(512, classes)          # eval_tumorbed.py:27
        self.regressor = _Regressor(512, 128, 1)             # eval_tumorbed.py:28
        self.encoder.__dict__["_owner"] = self
        self.activation = activation                          # only smp's .predict() applies it

    def forward(self, x):
        return self.decoder(self.encoder(x))


def unet_resnet18(classes=4):
    return Unet("resnet18", None, classes)
