"""CPU-only: the reference-shaped Python surface (names, constructor signatures, call protocol,
error behaviour) -- SURVEY.md 8(b)."""
import inspect
import sys

import pytest
import torch


def test_reference_import_paths_and_signatures():
    import detectron.pytorch_b200 as pkg
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("model", "modeling")}
    try:
        pkg.install_reference_aliases()
        from modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction          # model_builder.py:13
        from modeling.roi_xfrom.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
        from model.roi_align.functions.roi_align import RoIAlignFunction as LegacyFn
        from model.roi_align.modules.roi_align import RoIAlignAvg as LegacyAvg
        from model.roi_pooling.functions.roi_pool import RoIPoolFunction                       # model_builder.py:11
        from model.roi_pooling.modules.roi_pool import _RoIPooling
        from model.roi_crop.functions.roi_crop import RoICropFunction                          # model_builder.py:12
        from model.roi_crop.modules.roi_crop import _RoICrop
        from model.nms.nms_gpu import nms_gpu
        from model.nms.nms_wrapper import nms
        assert list(inspect.signature(RoIAlignFunction.__init__).parameters)[1:] == ["aligned_height", "aligned_width", "spatial_scale", "sampling_ratio"]
        assert list(inspect.signature(LegacyFn.__init__).parameters)[1:] == ["aligned_height", "aligned_width", "spatial_scale"]
        assert list(inspect.signature(RoIPoolFunction.__init__).parameters)[1:] == ["pooled_height", "pooled_width", "spatial_scale"]
        assert list(inspect.signature(nms_gpu).parameters) == ["dets", "thresh"]
        assert list(inspect.signature(nms).parameters) == ["dets", "thresh", "force_cpu"]
        f = RoIAlignFunction(7.0, 7, "0.25" and 0.25, 2)
        assert (f.aligned_height, f.aligned_width, f.spatial_scale, f.sampling_ratio) == (7, 7, 0.25, 2)
        assert f.rois is None and f.feature_size is None
        for m in (RoIAlign(7, 7, 0.25, 2), RoIAlignAvg(7, 7, 0.25, 2), RoIAlignMax(7, 7, 0.25, 2), LegacyAvg(7, 7, 0.25),
                  _RoIPooling(7, 7, 0.25), _RoICrop()):
            assert isinstance(m, torch.nn.Module)
        assert RoICropFunction().input1 is None
        from model.roi_crop.functions.gridgen import AffineGridGenFunction                       # helpers of the RoICrop mode
        from model.roi_crop.modules.gridgen import _AffineGridGen
        from model.roi_crop.functions.crop_resize import RoICropFunction as CropResize
        assert issubclass(CropResize, RoICropFunction) and CropResize().device == -1          # the reference's second RoICrop front end (records its device)
        assert _AffineGridGen(3, 4).f.height == 3 and AffineGridGenFunction(3, 4).width == 4
        assert "modeling.generate_proposals" not in sys.modules                                  # opt-in only
        pkg.install_reference_aliases(proposals=True)
        from modeling.generate_proposals import GenerateProposalsOp
        from modeling.collect_and_distribute_fpn_rpn_proposals import CollectAndDistributeFpnRpnProposalsOp
        assert list(inspect.signature(GenerateProposalsOp.__init__).parameters)[1:3] == ["anchors", "spatial_scale"]
        assert isinstance(CollectAndDistributeFpnRpnProposalsOp(), torch.nn.Module)
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("model", "modeling")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_cpu_tensors_are_rejected_like_the_reference():
    from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    from detectron.pytorch_b200.model.roi_align.functions.roi_align import RoIAlignFunction as LegacyFn
    from detectron.pytorch_b200.model.roi_pooling.functions.roi_pool import RoIPoolFunction
    from detectron.pytorch_b200.model.roi_crop.functions.roi_crop import RoICropFunction
    from detectron.pytorch_b200.model.nms.nms_wrapper import nms
    feats, rois = torch.zeros(1, 2, 8, 8), torch.zeros(3, 5)
    with pytest.raises(NotImplementedError):        # functions/roi_align.py:28-29
        RoIAlignFunction(7, 7, 0.25, 2)(feats, rois)
    with pytest.raises(NotImplementedError):
        LegacyFn(7, 7, 0.25)(feats, rois)
    with pytest.raises(NotImplementedError):
        RoIPoolFunction(7, 7, 0.25)(feats, rois)
    with pytest.raises(NotImplementedError):
        RoICropFunction()(feats, torch.zeros(1, 7, 7, 2))
    assert nms(torch.zeros(0, 5), 0.7) == []        # nms_wrapper.py:13-14, before any device work
    f = RoIAlignFunction(7, 7, 0.25, 2)
    with pytest.raises(AssertionError):             # functions/roi_align.py:35 `assert ... grad_output.is_cuda`
        f.backward(torch.zeros(3, 2, 7, 7))


def test_no_silent_fallback_when_library_missing(monkeypatch, tmp_path):
    from detectron.pytorch_b200 import _lib, build
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(build, "find_nvcc", lambda: None)
    with pytest.raises(ImportError):
        _lib.load()


def test_generate_proposals_op_mirrors_the_reference_configuration_surface():
    """Host-side behaviour of the device proposal layer: cfg[TRAIN|TEST].RPN_* lookup at call time like the reference
    (generate_proposals.py:108-113), keyword overrides, and no silent CPU path."""
    import numpy as np
    import pytest
    import torch
    from detectron.pytorch_b200.modeling.generate_proposals import GenerateProposalsOp

    class _Mode(object):
        def __init__(self, pre, post, thr, ms):
            self.RPN_PRE_NMS_TOP_N, self.RPN_POST_NMS_TOP_N, self.RPN_NMS_THRESH, self.RPN_MIN_SIZE = pre, post, thr, ms

    cfg = {"TRAIN": _Mode(12000, 2000, 0.7, 0), "TEST": _Mode(1000, 300, 0.5, 4)}
    anchors = np.array([[-8., -8., 8., 8.]])
    op = GenerateProposalsOp(anchors, 1.0 / 16, cfg=cfg)
    assert op._mode_params() == (12000, 2000, 0.7, 0)
    op.eval()
    assert op._mode_params() == (1000, 300, 0.5, 4)
    cfg["TEST"].RPN_POST_NMS_TOP_N = 50                      # read at call time, not at construction
    assert op._mode_params()[1] == 50
    op2 = GenerateProposalsOp(anchors, 1.0 / 16, test=dict(RPN_PRE_NMS_TOP_N=77))
    op2.eval()
    assert op2._mode_params() == (77, 1000, 0.7, 0)          # the reference's TEST defaults for the rest
    assert op2._feat_stride == 16.0 and op2._num_anchors == 1
    with pytest.raises(NotImplementedError):
        op2(torch.zeros(1, 1, 2, 2), torch.zeros(1, 4, 2, 2), torch.tensor([[32., 32., 1.]]))


def test_module_shells_call_their_function_with_the_reference_arguments():
    """The nn.Module shells (detectron.pytorch_b200._modules) on CPU, with a recording stand-in for the CUDA function:
    constructor arguments and attributes as in the reference, Avg / Max enlarge the aligned size by one and pool 2x2."""
    import pytest
    import torch
    from detectron.pytorch_b200._modules import roi_module
    from detectron.pytorch_b200.model.roi_align.modules import roi_align as legacy
    from detectron.pytorch_b200.model.roi_crop.modules.roi_crop import _RoICrop
    from detectron.pytorch_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.modules import roi_align as xfrom

    calls = []

    class Recorder(object):
        def __init__(self, *ctor):
            self.ctor = ctor

        def __call__(self, features, rois):
            calls.append(self.ctor)
            return torch.arange(float(rois.shape[0] * 2 * self.ctor[0] * self.ctor[1])).reshape(rois.shape[0], 2, self.ctor[0], self.ctor[1])

    for mod, extra in ((xfrom, (2,)), (legacy, ())):
        for cls, grow in ((mod.RoIAlign, 0), (mod.RoIAlignAvg, 1), (mod.RoIAlignMax, 1)):
            probe = roi_module(cls.__name__, Recorder, cls._fields, grow=cls._grow, epilogue=cls._epilogue)
            m = probe(7, 5, 0.25, *extra)
            assert (m.aligned_height, m.aligned_width, m.spatial_scale) == (7, 5, 0.25)
            y = m(torch.zeros(1, 2, 8, 8), torch.zeros(3, 5))
            assert calls[-1] == (7 + grow, 5 + grow, 0.25) + extra
            assert y.shape == (3, 2, 7, 5)
        assert issubclass(mod.RoIAlignAvg, mod.RoIAlign) and issubclass(mod.RoIAlignMax, mod.RoIAlign)
    ref = torch.arange(3 * 2 * 8 * 6, dtype=torch.float32).reshape(3, 2, 8, 6)
    m = roi_module("RoIAlignMax", Recorder, xfrom.RoIAlignMax._fields, grow=xfrom.RoIAlignMax._grow, epilogue="max")(7, 5, 1.0, 2)
    assert torch.equal(m(torch.zeros(1, 2, 4, 4), torch.zeros(3, 5)), torch.nn.functional.max_pool2d(ref, 2, 1))
    # the real classes reach the CUDA functions, which refuse CPU tensors like the reference does
    for m in (xfrom.RoIAlignAvg(7, 7, 0.25, 2), legacy.RoIAlign(7, 7, 0.25), _RoIPooling(7, 7, 0.0625)):
        with pytest.raises(NotImplementedError):
            m(torch.zeros(1, 2, 8, 8), torch.zeros(1, 5))
    assert _RoICrop().layout == 'BHWD' and _RoIPooling(7, 6, 0.5).pooled_width == 6
    with pytest.raises(TypeError):
        xfrom.RoIAlign(7, 7, 0.25)
