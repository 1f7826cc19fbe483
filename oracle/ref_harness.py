"""Reference harness — test infrastructure only (see oracle/wsi_oracle.py header).

Runs the UNMODIFIED reference code from /root/reference (``utils.dataset.Dataset_wsis``,
``utils.eval.predict_tumorbed``, ``resnets_shift.ResNet``) on CPU against an in-memory
synthetic slide, following the shim recipe of SURVEY.md §8(c).  It exists to (1) pin the oracle
restatement and (2) generate the golden fixtures under tests/golden/ (make_golden.py).

It only works where /root/reference exists (the build container).  Nothing under tests/ marked
``gpu``, ``__graft_entry__.smoke()`` or ``bench.py`` imports this at run time on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("WSI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "eval.py"))


# ------------------------------------------------------------------------------------------
# fake openslide serving numpy rasters: {level: u8 [h, w, 3]}; level-0 coordinates on input
# ------------------------------------------------------------------------------------------
_SLIDES: dict[str, dict[int, np.ndarray]] = {}
_DOWNSAMPLES = (1.0, 4.0, 16.0)


class _FakeOpenSlide:
    def __init__(self, path):
        from PIL import Image  # noqa: F401
        self._levels = _SLIDES[os.path.abspath(path)]
        h2, w2 = self._levels[2].shape[:2]
        self.level_downsamples = _DOWNSAMPLES
        self.level_dimensions = tuple((int(w2 * 16 / d), int(h2 * 16 / d)) for d in _DOWNSAMPLES)

    def read_region(self, location, level, size):
        from PIL import Image
        ds = _DOWNSAMPLES[level]
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


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_INSTALLED = False


def install_shim():
    """Idempotent.  Mutates sys.argv/sys.modules/np/torch.Tensor.cuda — call only from golden
    generation or the CPU reference baseline, never from product code."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.argv = ["x"]                                   # myargs.py:139 parses at import
    if not hasattr(np, "float"):
        np.float = float                               # utils/eval.py:46,183
    if not hasattr(np, "int"):
        np.int = int
    _stub("openslide", OpenSlide=_FakeOpenSlide)
    _stub("mahotas", bwperim=lambda a, *k, **kw: np.zeros_like(a))
    sk = _stub("skimage")
    sk.color = _stub("skimage.color")
    sk.morphology = _stub("skimage.morphology")
    sk.morphology.convex_hull = _stub("skimage.morphology.convex_hull", convex_hull_image=lambda a: a)
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot")
    _stub("adabound", AdaBound=torch.optim.Adam)
    torch.Tensor.cuda = lambda self, *a, **k: self      # CPU run of `.cuda()` call sites
    nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    _INSTALLED = True


def ref_modules():
    install_shim()
    import myargs
    import resnets_shift
    import utils.dataset as rds
    import utils.eval as rval
    import utils.preprocessing as rpre
    from models.models import Classifier, Regressor
    return types.SimpleNamespace(args=myargs.args, resnets_shift=resnets_shift, dataset=rds,
                                 eval=rval, preprocessing=rpre, Classifier=Classifier, Regressor=Regressor)


# ------------------------------------------------------------------------------------------
# models presented to the reference loop
# ------------------------------------------------------------------------------------------
class _Fn(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def forward(self, *a):
        return self._fn(*a)


def make_reference_resnet(sd):
    """The reference's own resnets_shift.ResNet loaded with ``sd`` (missing fc/fc1/fc2 keys keep
    their constructor init — they are not on the sliding-window path)."""
    R = ref_modules()
    torch.manual_seed(0)
    net = R.resnets_shift.resnet18()
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.split(".")[0] in ("fc", "fc1", "fc2") for k in missing), missing
    return net.eval()


def _trunk(net, x):
    x0 = net.relu(net.bn1(net.conv1(x)))
    x1 = net.layer1(net.maxpool(x0))
    x2 = net.layer2(x1)
    x3 = net.layer3(x2)
    x4 = net.layer4(x3)
    return [x4, x3, x2, x1, x0]


class ResNetClsAdapter(nn.Module):
    """resnets_shift.ResNet does not itself satisfy the predict_tumorbed protocol;
    the config-1 adapter exposes encoder/classifier built from the reference's own sub-modules."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.encoder = _Fn(lambda x: _trunk(net, x))
        self.classifier = _Fn(lambda e: net.fc0(torch.flatten(net.avgpool(e), 1)))
        self.decoder = nn.Identity()
        self.regressor = nn.Identity()


class UnetAdapter(nn.Module):
    """encoder = the reference's resnets_shift modules (same arithmetic as smp's ResNetEncoder);
    decoder = the oracle's restated smp decoder (UNPINNED); heads = reference Classifier/Regressor."""

    def __init__(self, sd):
        super().__init__()
        from oracle import wsi_oracle as O
        R = ref_modules()
        enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        torch.manual_seed(0)
        net = R.resnets_shift.resnet18()
        net.load_state_dict(enc_sd, strict=False)
        net.eval()
        self.net = net
        self.encoder = _Fn(lambda x: _trunk(net, x))
        self.decoder = _Fn(lambda feats: O.unet_decoder(sd, feats))
        self.classifier = R.Classifier(512, 4)
        self.classifier.load_state_dict({k[len("classifier."):]: v for k, v in sd.items() if k.startswith("classifier.")})
        self.regressor = R.Regressor(512, 1)
        self.regressor.load_state_dict({k[len("regressor."):]: v for k, v in sd.items() if k.startswith("regressor.")})

    def forward(self, x):
        """predict_wsis calls the model itself (utils/eval.py:51): smp.Unet.forward = decoder(encoder(x))."""
        return self.decoder(self.encoder(x))


# ------------------------------------------------------------------------------------------
# run the reference loop on one synthetic slide
# ------------------------------------------------------------------------------------------
def run_reference_predict_tumorbed(model, levels: dict, mask: np.ndarray, workdir: str, *, ph, pw, sh, sw,
                                   mode, scan_level=2, batch=16, shuffle_seed=0, key="slide0.svs", scan_resize=1):
    """levels: {2: raster} (and {scan_level: raster} when scan_level != 2).  Returns dict with the
    reference's tile list, f64 canvas, classes, probs, heatmap (as saved to PNG).  ph, pw: the Dataset params
    (= tile * scan_resize, eval_tumorbed.py:39-40); args.tile_h / tile_w get ph / scan_resize."""
    from PIL import Image
    R = ref_modules()
    a = R.args
    root = os.path.join(workdir, "wsi")
    case = os.path.join(root, "Case_1")
    os.makedirs(case, exist_ok=True)
    maskdir = os.path.join(workdir, "masks")
    os.makedirs(maskdir, exist_ok=True)
    svs = os.path.join(case, key)
    open(svs, "wb").close()
    _SLIDES[os.path.abspath(svs)] = levels
    Image.fromarray(mask.astype(np.uint8)).save(os.path.join(maskdir, key + ".png"))

    a.wsi_mask_pth = maskdir
    a.val_save_pth = os.path.join(workdir, "out")
    a.scan_level = scan_level
    a.scan_resize = scan_resize
    a.workers = 0
    a.num_classes = 4
    a.tile_h, a.tile_w, a.tile_stride_h, a.tile_stride_w = ph // scan_resize, pw // scan_resize, sh, sw
    a.class_probs = [0.0, 0.0, 0.0, 0.0]

    ds = R.dataset.Dataset_wsis(root, {"ph": ph, "pw": pw, "sh": sh, "sw": sw}, bs=batch)
    tiles = list(ds.wsis[key]["iterator"].dataset.datalist)

    captured = {}
    orig = R.preprocessing.threshold_probs

    def capture(pred):
        captured["canvas"] = np.array(pred, copy=True)
        classes, probs = orig(pred)
        captured["classes"], captured["probs"] = classes, probs
        return classes, probs

    R.preprocessing.threshold_probs = capture
    try:
        torch.manual_seed(shuffle_seed)
        R.eval.predict_tumorbed(model, ds, 0, mode)
    finally:
        R.preprocessing.threshold_probs = orig
    heat = np.array(Image.open(os.path.join(a.val_save_pth, "0", f"{key}_{sw}_heatmap.png")))
    captured["heatmap"] = heat
    captured["tiles"] = tiles
    return captured


class _StopAfterScores(Exception):
    pass


def run_reference_predict_wsis(model, levels: dict, mask: np.ndarray, workdir: str, *, ph, pw, sh, sw, scan_level=2,
                               batch=16, shuffle_seed=0, key="slide0.svs", scan_resize=1):
    """Runs the reference's predict_wsis (utils/eval.py:22-60 loop, :66-71 cv2.resize to level 2) on one synthetic
    slide and returns its tile list, the f64 scan-level canvas is not observable, so: `pred` = the RESIZED [C,H2,W2]
    array handed to `preprocessing.pred_to_mask` (:139), where the run is stopped — pred_to_mask itself has the tuple
    bug of SURVEY §8c and the rest of the function (scores, tumour-bed morphology, colour mask) is out of scope."""
    from PIL import Image
    R = ref_modules()
    a = R.args
    root = os.path.join(workdir, "wsi")
    case = os.path.join(root, "Case_1")
    os.makedirs(case, exist_ok=True)
    maskdir = os.path.join(workdir, "masks")
    os.makedirs(maskdir, exist_ok=True)
    svs = os.path.join(case, key)
    open(svs, "wb").close()
    _SLIDES[os.path.abspath(svs)] = levels
    Image.fromarray(mask.astype(np.uint8)).save(os.path.join(maskdir, key + ".png"))           # Dataset_wsi foreground mask
    Image.fromarray(mask.astype(np.uint8)).save(svs + "_find_nuclei.png")                      # utils/eval.py:64

    a.wsi_mask_pth = maskdir
    a.val_save_pth = os.path.join(workdir, "out")
    a.scan_level = scan_level
    a.scan_resize = scan_resize
    a.workers = 0
    a.num_classes = 4
    a.tile_h, a.tile_w, a.tile_stride_h, a.tile_stride_w = ph // scan_resize, pw // scan_resize, sh, sw

    ds = R.dataset.Dataset_wsis(root, {"ph": ph, "pw": pw, "sh": sh, "sw": sw}, bs=batch)
    tiles = list(ds.wsis[key]["iterator"].dataset.datalist)
    captured = {}
    orig = R.preprocessing.pred_to_mask

    def capture(pred, *k, **kw):
        captured["pred"] = np.array(pred, copy=True)
        raise _StopAfterScores()

    R.preprocessing.pred_to_mask = capture
    try:
        torch.manual_seed(shuffle_seed)
        try:
            R.eval.predict_wsis(model, ds, 0)
        except _StopAfterScores:
            pass
    finally:
        R.preprocessing.pred_to_mask = orig
    captured["tiles"] = tiles
    return captured


def run_reference_predict_reg(model, images: "torch.Tensor", workdir: str, batch=3):
    """Runs the reference's predict_reg (utils/eval.py:288-352: 4-view TTA of the regressor) on normalised tiles
    ``images`` f32 [n,3,h,h] and returns its local ``preds`` list — the function only prints l1/mse, so the value is
    read from the frame at return time with sys.settrace (no reference code is edited).  It also writes overlay PNGs
    under ./data/cell_seg, so it is run from ``workdir``."""
    import sys
    R = ref_modules()
    n = images.shape[0]
    data = []
    for i in range(0, n, batch):
        im = images[i:i + batch]
        z = torch.zeros(im.shape[0])
        data.append((im, z, z, z, z, torch.zeros(im.shape[0])))
    captured = {}

    def tracer(frame, event, arg):
        if event == "call" and frame.f_code.co_name == "predict_reg":
            def local(fr, ev, a):
                if ev == "return":
                    captured["preds"] = np.array(fr.f_locals["preds"], np.float64)
                return local
            return local
        return None

    cwd = os.getcwd()
    os.makedirs(os.path.join(workdir, "data", "cell_seg"), exist_ok=True)
    os.chdir(workdir)
    sys.settrace(tracer)
    try:
        R.eval.predict_reg(model, data, 0)
    finally:
        sys.settrace(None)
        os.chdir(cwd)
    return captured["preds"]
