"""Make the reference's PYTHON tree importable on the GPU box (test infrastructure for SURVEY.md X1 / BASELINE cfg 4-5).

    python oracle/make_reflib.py            # build container only (needs /root/reference)

Writes oracle/_ref/reflib/ -- git-ignored, never part of the product, but it travels to the GPU box with the
snapshot exactly like oracle/_ref/*.so:
  * lib/**/*.py of the reference, copied verbatim (no file is edited);
  * utils/cython_bbox, utils/cython_nms built from the reference's .pyx (cython_nms with the 3-token numpy-2 dtype
    patch SURVEY.md 8c describes, applied to a temp copy);
  * the two baseline yaml files the harness builds models from.
The reference's compiled op extensions (_ext/, cffi) are NOT built: that is exactly the seam where this package's
ops are aliased in (detectron.pytorch_b200.install_reference_aliases).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "reflib")
CFGS = ["configs/baselines/e2e_faster_rcnn_R-50-FPN_1x.yaml", "configs/baselines/e2e_mask_rcnn_R-50-FPN_1x.yaml",
        "configs/baselines/e2e_faster_rcnn_R-50-C4_1x.yaml"]


def main():
    if not os.path.isdir(os.path.join(REF, "lib")):
        raise SystemExit("make_reflib: %s not present (run in the build container)" % REF)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    n = 0
    for dirpath, dirnames, files in os.walk(os.path.join(REF, "lib")):
        dirnames[:] = [d for d in dirnames if d not in ("_ext", "src", "build", "__pycache__")]
        rel = os.path.relpath(dirpath, os.path.join(REF, "lib"))
        for f in files:
            if f.endswith(".py"):
                dst = os.path.join(OUT, "lib", rel)
                os.makedirs(dst, exist_ok=True)
                shutil.copy(os.path.join(dirpath, f), dst)
                n += 1
    for c in CFGS:
        dst = os.path.join(OUT, os.path.dirname(c))
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(REF, c), dst)
    tmp = tempfile.mkdtemp(prefix="ref_cython_")
    for f in ("cython_nms.pyx", "cython_bbox.pyx"):
        shutil.copy(os.path.join(REF, "lib", "utils", f), tmp)
    p = os.path.join(tmp, "cython_nms.pyx")
    s = open(p).read().replace("np.int_t", "np.intp_t")
    s = re.sub(r"np\.int\b", "np.intp", s)
    open(p, "w").write(s)
    open(os.path.join(tmp, "setup.py"), "w").write(
        "from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy as np\n"
        "setup(ext_modules=cythonize([Extension('cython_nms', ['cython_nms.pyx'], include_dirs=[np.get_include()]),\n"
        "                             Extension('cython_bbox', ['cython_bbox.pyx'], include_dirs=[np.get_include()])], language_level=2))\n")
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in os.listdir(tmp):
        if f.endswith(".so"):
            shutil.copy(os.path.join(tmp, f), os.path.join(OUT, "lib", "utils"))
    shutil.rmtree(tmp, ignore_errors=True)
    print("make_reflib: %d .py files + 2 cython modules + %d configs -> %s" % (n, len(CFGS), OUT))


if __name__ == "__main__":
    main()
