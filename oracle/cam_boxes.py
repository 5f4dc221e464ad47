"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's CAM -> pseudo ground-truth path with the same OpenCV calls.

  resize_cam      cams_deit.py:9-14     cv2.resize(cam, (size[0], size[1])); cam -= min; cam /= max
  get_bboxes      cams_deit.py:34-58    (cam * 255).astype(uint8); threshold TOZERO at int(cam_thr * max); findContours(RETR_TREE,
                                         CHAIN_APPROX_SIMPLE); max by contourArea; boundingRect -> [x, y, x + w, y + h] (else [0, 0, 1, 1])
  get_pseudo_label engine.py:310-352    per image, per class with img_label > 0: the box of cams_cls[b, c] -> cxcywh / [w, h, w, h]
  get_multi_bboxes cams_deit.py:61-97   all contours with area >= area_ratio * max, by decreasing area
  get_pseudo_label_multi_boxes engine.py:356-398   the multi-box variant the refine training loops call

OpenCV (cv2 4.13 here, the reference pins nothing) is the un-vendored third-party dependency of this path; the product kernels
(spe_b200/csrc/cam_boxes.cu) restate its arithmetic.  Only tests/ may import this module."""
import numpy as np
import torch


def resize_cam(cam, size):
    import cv2
    cam = cv2.resize(cam, (size[0], size[1]))
    cam = cam - cam.min()
    cam = cam / cam.max()
    return cam


def get_bboxes(cam, cam_thr=0.2):
    import cv2
    cam = (cam * 255.).astype(np.uint8)
    map_thr = cam_thr * np.max(cam)
    _, thr = cv2.threshold(cam, int(map_thr), 255, cv2.THRESH_TOZERO)
    contours, _ = cv2.findContours(thr, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) != 0:
        c = max(contours, key=cv2.contourArea)
        x, y, w, h = cv2.boundingRect(c)
        return [x, y, x + w, y + h]
    return [0, 0, 1, 1]


def get_multi_bboxes(cam, cam_thr=0.2, area_ratio=0.5):
    """cams_deit.py:61-97: every contour with area >= area_ratio * the largest, by decreasing area (stable)."""
    import cv2
    cam = (cam * 255.).astype(np.uint8)
    map_thr = cam_thr * np.max(cam)
    _, thr = cv2.threshold(cam, int(map_thr), 255, cv2.THRESH_TOZERO)
    contours, _ = cv2.findContours(thr, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) != 0:
        out = []
        areas = list(map(cv2.contourArea, contours))
        area_idx = sorted(range(len(areas)), key=areas.__getitem__, reverse=True)
        for idx in area_idx:
            if areas[idx] >= areas[area_idx[0]] * area_ratio:
                x, y, w, h = cv2.boundingRect(contours[idx])
                out.append([x, y, x + w, y + h])
        return out, [areas[i] for i in area_idx if areas[i] >= areas[area_idx[0]] * area_ratio]
    return [[0, 0, 1, 1]], [0.0]


def pseudo_labels_multi(cams_cls, img_labels, image_size, cam_thr=0.2, area_ratio=0.5):
    """engine.get_pseudo_label_multi_boxes (engine.py:356-398).  Returns (list of {'boxes','labels'}, integer boxes per (image, class)
    pair, contour areas per pair)."""
    out, raw, ars = [], [], []
    H, W = image_size
    for b in range(cams_cls.shape[0]):
        boxes, labels = [], []
        for c in range(cams_cls.shape[1]):
            if img_labels[b][c] > 0:
                cam = cams_cls[b, [c]].mean(0, keepdim=True).numpy().transpose(1, 2, 0)
                cam = resize_cam(cam, size=(H, W))
                bb, areas = get_multi_bboxes(cam, cam_thr=cam_thr, area_ratio=area_ratio)
                bb = torch.tensor(bb)
                raw.append(bb)
                ars.append(areas)
                x0, y0, x1, y1 = bb[..., 0], bb[..., 1], bb[..., 2], bb[..., 3]
                boxes.append(torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1))
                labels += [c + 1] * bb.shape[0]
        if boxes:
            bx = torch.cat(boxes, dim=0) / torch.tensor([W, H, W, H], dtype=torch.float32)
            out.append({"boxes": bx, "labels": torch.tensor(labels)})
        else:
            out.append({"boxes": torch.zeros(0, 4), "labels": torch.zeros(0, dtype=torch.int64)})
    return out, raw, ars


def pseudo_labels(cams_cls, img_labels, image_size, cam_thr=0.2):
    """cams_cls f32 [B,C,h,w] (cpu), img_labels [B,C] (>0 = class present), image_size = (H, W) = samples.tensors.shape[-2:].
    Returns the list of {'boxes' f32 [k,4] cxcywh normalised, 'labels' int64 [k] (class + 1)} of engine.get_pseudo_label, plus the
    integer xyxy boxes."""
    out, raw = [], []
    H, W = image_size
    for b in range(cams_cls.shape[0]):
        boxes, labels, ints = [], [], []
        for c in range(cams_cls.shape[1]):
            if img_labels[b][c] > 0:
                cam = cams_cls[b, [c]].mean(0, keepdim=True).numpy().transpose(1, 2, 0)
                cam = resize_cam(cam, size=(H, W))
                bb = torch.tensor(get_bboxes(cam, cam_thr=cam_thr))
                ints.append(bb)
                x0, y0, x1, y1 = bb
                boxes.append(torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1))
                labels.append(c + 1)
        if boxes:
            bx = torch.stack(boxes) / torch.tensor([W, H, W, H], dtype=torch.float32)
            out.append({"boxes": bx, "labels": torch.tensor(labels)})
            raw.append(torch.stack(ints))
        else:
            out.append({"boxes": torch.zeros(0, 4), "labels": torch.zeros(0, dtype=torch.int64)})
            raw.append(torch.zeros(0, 4, dtype=torch.int64))
    return out, raw
