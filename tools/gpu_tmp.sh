python - <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from detectron.pytorch_b200.roi_data import fast_rcnn as FR
G = np.load("tests/golden/targets.npz")
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for c in ["b", "c"]:
    ncls, batch, agn = [int(v) for v in G[c + "_cfg"]]
    boxes = np.concatenate([G[c + "_gt"], G[c + "_prop"]]).astype(np.float32)
    b = FR.sample_rois(dev(boxes), dev(G[c + "_gt"]), dev(G[c + "_gt_classes"]), 1.5, 1, ncls, batch_size_per_im=batch,
                       cls_agnostic_bbox_reg=bool(agn), fg_choice=G[c + "_fg_choice"], bg_choice=G[c + "_bg_choice"])
    got = b["bbox_targets"].cpu().numpy(); ref = G[c + "_bbox_targets"]
    bad = np.argwhere((got == 0) != (ref == 0))
    print(c, "mismatching zero pattern entries:", len(bad))
    for i, j in bad[:6]:
        k = int(b["keep_inds"][i]); print("  row", i, "col", j, "got", got[i, j], "ref", ref[i, j], "keep", k, "box", boxes[k], "label", int(b["labels_int32"][i]))
PY
