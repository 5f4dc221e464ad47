"""engine.get_pseudo_label (engine.py:310-352) and engine.get_pseudo_label_multi_boxes (:356-398) of the reference on the device: the class-activation maps never leave the
GPU.  One batched sequence of launches (csrc/cam_boxes.cu) handles every (image, present class) pair: bilinear resize to the image
size, min-max normalisation, uint8 quantisation, threshold, connected components + contour areas, bounding box of the largest
contour -- bit-exact with the cv2 calls the reference makes (tests/test_cam_boxes_gpu.py)."""
import torch

from ._lib import check, lib, ptr, stream


def cam_boxes(cams_cls, pairs, image_size, cam_thr=0.2, return_xyxy=False):
    """cams_cls f32 [B,C,h,w] (cuda); pairs: int tensor / list [[b, c], ...]; image_size = (H, W).
    -> boxes f32 [npairs,4] (cxcywh / [W,H,W,H]) on the device (+ integer [x, y, x+w, y+h] boxes)."""
    if not cams_cls.is_cuda:
        raise RuntimeError("spe_b200 CAM boxes need CUDA tensors (no CPU fallback exists)")
    dev = cams_cls.device
    cams = cams_cls.detach().float().contiguous()
    B, C, h, w = cams.shape
    pairs = torch.as_tensor(pairs, dtype=torch.int32).reshape(-1, 2).to(dev).contiguous()
    n = int(pairs.shape[0])
    H, W = int(image_size[0]), int(image_size[1])
    boxes = torch.empty((n, 4), dtype=torch.float32, device=dev)
    xyxy = torch.empty((n, 4), dtype=torch.int32, device=dev)
    if n == 0:
        return (boxes, xyxy) if return_xyxy else boxes
    rows, cols = W, H                      # cams_deit.resize_cam hands (H, W) to cv2.resize as dsize = (width, height)
    thr = int(cam_thr * 255)               # get_bboxes: int(cam_thr * np.max(cam_u8)); the normalised maximum is always 255
    nbytes = int(lib().spe_cam_boxes_workspace_bytes(n, rows, cols))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib().spe_cam_boxes(ptr(cams), B, C, h, w, ptr(pairs), n, rows, cols, thr, float(W), float(H), ptr(boxes), ptr(xyxy), ptr(ws), nbytes, stream()))
    return (boxes, xyxy) if return_xyxy else boxes


def cam_boxes_multi(cams_cls, pairs, image_size, cam_thr=0.2, area_ratio=0.5, max_boxes=16, return_xyxy=False):
    """cams_deit.get_multi_bboxes for every (image, class) pair: boxes f32 [npairs, max_boxes, 4] (cxcywh / [W,H,W,H], by decreasing
    contour area), counts i32 [npairs] (+ integer boxes) -- all on the device."""
    if not cams_cls.is_cuda:
        raise RuntimeError("spe_b200 CAM boxes need CUDA tensors (no CPU fallback exists)")
    dev = cams_cls.device
    cams = cams_cls.detach().float().contiguous()
    B, C, h, w = cams.shape
    pairs = torch.as_tensor(pairs, dtype=torch.int32).reshape(-1, 2).to(dev).contiguous()
    n = int(pairs.shape[0])
    H, W = int(image_size[0]), int(image_size[1])
    boxes = torch.zeros((n, max_boxes, 4), dtype=torch.float32, device=dev)
    xyxy = torch.zeros((n, max_boxes, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros((n,), dtype=torch.int32, device=dev)
    if n:
        rows, cols = W, H
        nbytes = int(lib().spe_cam_boxes_workspace_bytes(n, rows, cols))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib().spe_cam_boxes_multi(ptr(cams), B, C, h, w, ptr(pairs), n, rows, cols, int(cam_thr * 255), float(W), float(H), float(area_ratio),
                                        int(max_boxes), ptr(boxes), ptr(xyxy), ptr(counts), ptr(ws), nbytes, stream()))
    return (boxes, counts, xyxy) if return_xyxy else (boxes, counts)


@torch.no_grad()
def get_pseudo_label_multi_boxes(outputs, samples, targets, args=None, cam_thr=None, area_ratio=None, max_boxes=16, label_key="img_label"):
    """Same call as the reference's engine.get_pseudo_label_multi_boxes (engine.py:356-398), the variant its refine training loops use:
    per image {'boxes': cxcywh normalised [k,4], 'labels': class + 1 [k]} with every contour above args.multi_box_ratio of the largest.
    One D2H read of the per-pair box counts (the result is ragged); maps and boxes stay on the device."""
    cams = outputs["cams_cls"]
    tens = samples.tensors if hasattr(samples, "tensors") else samples
    H, W = int(tens.shape[-2]), int(tens.shape[-1])
    thr = cam_thr if cam_thr is not None else float(getattr(args, "cam_thr", 0.2))
    ratio = area_ratio if area_ratio is not None else float(getattr(args, "multi_box_ratio", 0.5))
    ncls = int(getattr(args, "num_classes", cams.shape[1])) if args is not None else cams.shape[1]
    labels = torch.stack([torch.as_tensor(t[label_key]).reshape(-1)[:ncls] for t in targets]).cpu()
    pairs = torch.nonzero(labels > 0)
    boxes, counts = cam_boxes_multi(cams, pairs, (H, W), thr, ratio, max_boxes)
    cnt = counts.cpu()
    keep = torch.arange(max_boxes)[None, :] < cnt[:, None]                       # [npairs, max_boxes]
    flat = boxes[keep.to(boxes.device)]                                           # pairs in order, boxes by decreasing area
    cls = (pairs[:, 1] + 1).repeat_interleave(cnt.long()).to(cams.device)
    per_img = torch.zeros(len(targets), dtype=torch.int64).index_add_(0, pairs[:, 0], cnt.long()).tolist()
    out, o = [], 0
    for n in per_img:
        out.append({"boxes": flat[o:o + n], "labels": cls[o:o + n]})
        o += n
    return out


@torch.no_grad()
def get_pseudo_label_multi_boxes_voc(outputs, samples, targets, args=None, cam_thr=None, max_boxes=16):
    """engine.get_pseudo_label_multi_boxes_voc (engine.py:402-442): present classes from targets[b]['label'], get_multi_bboxes with its
    default area ratio 0.5."""
    return get_pseudo_label_multi_boxes(outputs, samples, targets, args, cam_thr=cam_thr, area_ratio=0.5, max_boxes=max_boxes, label_key="label")


@torch.no_grad()
def get_pseudo_label(outputs, samples, targets, args=None, cam_thr=None):
    """Same call as the reference's engine.get_pseudo_label(outputs, samples, targets, args): list (one per image) of
    {'boxes': cxcywh normalised, 'labels': class + 1} on the device.  The present classes come from targets[b]['img_label'] (read on
    the host, as the reference does); the maps stay on the device."""
    cams = outputs["cams_cls"]
    tens = samples.tensors if hasattr(samples, "tensors") else samples
    H, W = int(tens.shape[-2]), int(tens.shape[-1])
    thr = cam_thr if cam_thr is not None else float(getattr(args, "cam_thr", 0.2))
    ncls = int(getattr(args, "num_classes", cams.shape[1])) if args is not None else cams.shape[1]
    labels = torch.stack([t["img_label"].reshape(-1)[:ncls] for t in targets]).cpu()
    pairs = torch.nonzero(labels > 0)                     # row-major: images in order, classes ascending -- the reference's loop order
    boxes = cam_boxes(cams, pairs, (H, W), thr)
    counts = torch.bincount(pairs[:, 0], minlength=len(targets)).tolist()
    cls = (pairs[:, 1] + 1).to(cams.device)
    out, o = [], 0
    for n in counts:
        out.append({"boxes": boxes[o:o + n], "labels": cls[o:o + n]})
        o += n
    return out
