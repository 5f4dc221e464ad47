"""Golden vectors for the CAM -> pseudo-GT path from the OpenCV calls the reference makes (oracle/cam_boxes.py restates
engine.get_pseudo_label / cams_deit.resize_cam / cams_deit.get_bboxes with the same cv2 functions).  Also re-checks the two facts the
CUDA kernels rely on: (1) cv2.resize == fma(s1 - s0, w, s0) with double-precision source coordinates, horizontal then vertical pass;
(2) cv2.contourArea(outer border) == #full 2x2 cells + 1/2 #3-pixel cells of the hole-filled component.

    python tests/golden/make_cam_fixture.py        (needs cv2 + scipy; run in the build container)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cam_boxes as OC  # noqa: E402


def synth_cams(seed, B, C, h, w, structured=False):
    import cv2
    rng = np.random.default_rng(seed)
    cams = np.zeros((B, C, h, w), np.float32)
    for b in range(B):
        for c in range(C):
            kind = (b * C + c) % 4
            z = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), [3.0, 1.5, 5.0, 2.0][kind])
            if kind == 1:                                   # ring-like maps: holes and islands
                yy, xx = np.mgrid[0:h, 0:w]
                r = np.hypot(yy - h / 2, xx - w / 2)
                z = z * 0.3 + np.exp(-((r - h / 4) ** 2) / 8.0).astype(np.float32) + 0.6 * np.exp(-(r ** 2) / 4.0).astype(np.float32)
            if kind == 2 and structured:                    # several blobs of similar size: multi-box selection
                yy, xx = np.mgrid[0:h, 0:w]
                z = np.zeros((h, w), np.float32)
                for _ in range(4):
                    cy, cx, sg = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w, rng.uniform(0.06, 0.1) * min(h, w)
                    z += (rng.uniform(0.8, 1.0) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg))).astype(np.float32)
            if kind == 3 and structured:                    # annulus with a detached island in the hole: hole border + nested outer border
                yy, xx = np.mgrid[0:h, 0:w]
                r = np.hypot((yy - h / 2) / h, (xx - w / 2) / w)
                z = (np.exp(-((r - 0.3) ** 2) / 0.004) + 0.9 * np.exp(-(r ** 2) / 0.003)).astype(np.float32) + 0.02 * z
            cams[b, c] = z
    return torch.from_numpy(cams)


def main():
    import cv2
    cases = []
    for seed, (B, C, h, w, H, W) in enumerate([(2, 6, 40, 40, 640, 640), (2, 5, 14, 14, 224, 224), (1, 6, 24, 32, 384, 512), (1, 4, 50, 83, 800, 1333),
                                               (2, 8, 40, 40, 640, 640), (1, 8, 32, 24, 512, 384)]):
        cams = synth_cams(seed, B, C, h, w, structured=seed >= 4)
        g = torch.Generator().manual_seed(seed)
        lab = (torch.rand(B, C, generator=g) < 0.7).float()
        lab[:, 0] = 1
        for thr in (0.2, 0.35):
            pl, raw = OC.pseudo_labels(cams, lab, (H, W), cam_thr=thr)
            case = {"cams": cams, "img_label": lab, "image_size": (H, W), "cam_thr": thr,
                    "boxes": [p["boxes"] for p in pl], "labels": [p["labels"] for p in pl], "xyxy": raw, "multi": {}}
            for ratio in (0.5, 0.15):                    # engine.get_pseudo_label_multi_boxes: args.multi_box_ratio default 0.5
                plm, rawm, areas = OC.pseudo_labels_multi(cams, lab, (H, W), cam_thr=thr, area_ratio=ratio)
                case["multi"][ratio] = {"boxes": [p["boxes"] for p in plm], "labels": [p["labels"] for p in plm], "xyxy": rawm, "areas": areas}
            cases.append(case)
    torch.save({"cases": cases, "cv2": cv2.__version__}, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cam_boxes.pt"))
    print("cam_boxes.pt: %d cases, %d boxes, %d multi boxes" % (len(cases), sum(len(x) for c in cases for x in c["xyxy"]),
                                                                 sum(len(x) for c in cases for r in c["multi"].values() for x in r["xyxy"])))


if __name__ == "__main__":
    main()
