"""SURVEY.md row X1 (BASELINE.json configs[3]): the reference's own Generalized_RCNN, imported unmodified from
oracle/_ref/reflib, runs its e2e_faster_rcnn_R-50-FPN_1x forward on the GPU with this package's ops at the reference's
import paths, and every roi_feature_transform call equals the same call backed by the reference's own CUDA kernels."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reference_tree_importable_with_aliases_on_cpu():
    """CPU part: the reference's lib/ imports under today's torch / numpy with the shims of oracle/refmodel.py, and
    model_builder's op names resolve to this package (no compute here)."""
    from oracle import refmodel
    if not refmodel.available():
        pytest.skip("oracle/_ref/reflib not built (python oracle/make_reflib.py in the build container)")
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import refmodel\n"
            "import detectron.pytorch_b200 as pkg\n"
            "refmodel.setup(); pkg.install_reference_aliases(nms=True)\n"
            "import modeling.model_builder as mb, utils.boxes as bu\n"
            "from detectron.pytorch_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction\n"
            "from detectron.pytorch_b200.utils.boxes import nms\n"
            "assert mb.RoIAlignFunction is RoIAlignFunction and bu.nms is nms\n"
            "m = refmodel.build_model('e2e_faster_rcnn_R-50-FPN_1x.yaml')\n"
            "assert sum(p.numel() for p in m.parameters()) == 41757156\n"
            "print('ok')\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:]


@pytest.mark.gpu
def test_cfg4_faster_rcnn_fpn_forward_matches_reference_kernels():
    from oracle import gpu_ref as G
    from oracle import refmodel
    if not refmodel.available() or not G.available():
        pytest.skip("oracle/_ref (reflib + reference kernels) not built")
    import subprocess
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "x1_cfg4.py"), "--iters", "2"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    import json
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["roi_feature_transform_calls"] >= 1 and res["rois"] > 0
    for c in res["parity_vs_reference_kernels"]:
        assert c["max_abs_diff"] <= 1e-6, c
        assert c["frac_bit_equal"] > 0.999, c


def test_cfg5_train_step_harness_two_ranks_gloo_cpu():
    """SURVEY 8e / row X2, host logic only: the unmodified reference Mask R-CNN runs one TRAINING step (forward, losses,
    backward, SGD) per rank under torchrun with world_size 2 on the gloo backend; gradients and the loss dict are
    all-reduced as on the GPU box.  RoIAlign is torchvision's here (no GPU in this container): the harness, the compat
    shims and the collective pattern are what is under test."""
    from oracle import refmodel
    if not refmodel.available():
        pytest.skip("oracle/_ref/reflib not built (python oracle/make_reflib.py in the build container)")
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "x2_cfg5.py"), "--device", "cpu", "--stub-ops", "--steps", "1",
           "--warmup", "0", "--images", "1"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == 2 and res["row"] == "X2"
    assert res["losses"]["total"] > 0 and all(v == v for v in res["losses"].values())       # finite
    assert "loss_mask" in res["losses"] and "loss_rpn_cls_fpn2" in res["losses"]
