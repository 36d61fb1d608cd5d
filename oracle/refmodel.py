"""TEST INFRASTRUCTURE (BASELINE.json configs 4-5): import the UNMODIFIED reference model code
(oracle/_ref/reflib, made by oracle/make_reflib.py) under today's torch / numpy, with this package's ops aliased in
at the reference's import paths.  Nothing under detectron/ imports this module.

The shims below are exactly the list SURVEY.md 8c verified; they are applied from OUTSIDE (module attributes, sys.modules
entries) -- no reference file is edited.
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFLIB = os.path.join(HERE, "_ref", "reflib")
_state = {}


def available():
    return os.path.isdir(os.path.join(REFLIB, "lib", "modeling"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_compat():
    import torch
    import yaml
    six = _stub("torch._six", string_classes=(str, bytes), int_classes=(int,))
    torch._six = six
    import torch.utils.data.dataloader as dl
    if not hasattr(dl, "numpy_type_map"):
        dl.numpy_type_map = {}
    collections.Sequence = collections.abc.Sequence
    collections.Mapping = collections.abc.Mapping
    collections.Iterable = collections.abc.Iterable
    for k, v in (("float", float), ("int", int), ("bool", bool), ("object", object)):
        if k not in np.__dict__:
            setattr(np, k, v)
    # pycocotools: only needed to import modeling/*; masks are rasterised with cv2 where the train step needs them
    if "pycocotools" not in sys.modules:
        try:
            import pycocotools  # noqa: F401
        except ImportError:
            def _fr(polys, h, w):
                import cv2
                out = []
                for poly in polys:
                    m = np.zeros((h, w), dtype=np.uint8)
                    pts = np.asarray(poly, dtype=np.float64).reshape(-1, 2)
                    cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
                    out.append(m)
                return out
            mask = _stub("pycocotools.mask", frPyObjects=_fr, decode=lambda rles: np.stack(rles, axis=2) if isinstance(rles, list) else rles,
                         merge=lambda rles: np.maximum.reduce(rles), area=lambda r: float(np.sum(r)), iou=None, encode=lambda m: m)
            coco = _stub("pycocotools.coco", COCO=object)
            cocoeval = _stub("pycocotools.cocoeval", COCOeval=object)
            _stub("pycocotools", mask=mask, coco=coco, cocoeval=cocoeval)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            plt = _stub("matplotlib.pyplot")
            patches = _stub("matplotlib.patches", Polygon=object)
            _stub("matplotlib", use=lambda *a, **k: None, pyplot=plt, patches=patches)
    if not getattr(yaml, "_b200_patched", False):
        _load = yaml.load
        yaml.load = lambda stream, Loader=None: _load(stream, Loader=Loader or yaml.SafeLoader)
        yaml._b200_patched = True


def setup(use_b200_ops=True, proposals=False):
    """Returns the reference's `cfg` after making `lib/` importable.  use_b200_ops=False leaves the op import paths to
    the caller (who registers oracle/_ref-backed stand-ins before importing modeling.model_builder)."""
    if _state.get("ready"):
        return _state["cfg"]
    if not available():
        raise RuntimeError("oracle/_ref/reflib missing: run `python oracle/make_reflib.py` in the build container")
    _install_compat()
    lib = os.path.join(REFLIB, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    import utils.cython_bbox  # noqa: F401  (built from the reference's .pyx)
    import utils.cython_nms  # noqa: F401
    if use_b200_ops:
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(0, root)
        import detectron.pytorch_b200 as pkg
        pkg.install_reference_aliases(proposals=proposals)
    from core.config import cfg
    import utils.blob as blob_utils
    import utils.net as net_utils
    # numpy 2 removed the binary mode of np.fromstring (lib/utils/blob.py:165-169)
    import pickle
    blob_utils.serialize = lambda obj: np.frombuffer(pickle.dumps(obj), dtype=np.uint8).astype(np.float32)
    # .view(-1) on a sliced (non-contiguous) RPN target (lib/utils/net.py:31, lib/modeling/FPN.py:430-436): hand the
    # reference's own function contiguous copies instead of editing it
    _sl1 = net_utils.smooth_l1_loss
    net_utils.smooth_l1_loss = lambda p, t, wi, wo, beta=1.0: _sl1(p.contiguous(), t.contiguous(), wi.contiguous(), wo.contiguous(), beta)
    _state.update(ready=True, cfg=cfg, blob_utils=blob_utils, net_utils=net_utils)
    return cfg


def build_model(cfg_name, num_classes=81, seed=0):
    """Generalized_RCNN for a baseline yaml (random init, no pretrained weights)."""
    import torch
    cfg = setup()
    from core.config import cfg_from_file, assert_and_infer_cfg
    cfg_from_file(os.path.join(REFLIB, "configs", "baselines", cfg_name))
    cfg.MODEL.NUM_CLASSES = num_classes
    cfg.MODEL.LOAD_IMAGENET_PRETRAINED_WEIGHTS = False
    cfg.RESNETS.IMAGENET_PRETRAINED_WEIGHTS = ""
    assert_and_infer_cfg(make_immutable=False)
    torch.manual_seed(seed)
    from modeling.model_builder import Generalized_RCNN
    return Generalized_RCNN()
