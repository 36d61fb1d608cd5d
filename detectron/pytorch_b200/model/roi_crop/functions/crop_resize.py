"""Alternative import path of the RoICrop function; mirrors lib/model/roi_crop/functions/crop_resize.py:8-37
(reference), whose CUDA branch calls the same BilinearSamplerBHWD launchers as functions/roi_crop.py."""
from .roi_crop import RoICropFunction  # noqa: F401
