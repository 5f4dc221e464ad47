"""util/box_ops.py of the reference (:18-74) over the sm_100a pairwise kernel (spe_box_iou_pairwise).
box_iou / generalized_box_iou run on the GPU and carry no autograd (the criterion's fused loss kernels
produce the box gradients); format conversions are views/arithmetics kept as in the reference."""
import torch

from .._lib import check, lib, ptr, stream


def box_cxcywh_to_xyxy(x):
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([(x_c - 0.5 * w), (y_c - 0.5 * h), (x_c + 0.5 * w), (y_c + 0.5 * h)], dim=-1)


def box_xyxy_to_cxcywh(x):
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


def _pairwise(boxes1, boxes2, want_giou):
    if not boxes1.is_cuda:
        raise RuntimeError("spe_b200.util.box_ops needs CUDA tensors (no CPU fallback exists)")
    a = boxes1.detach().float().contiguous()
    b = boxes2.detach().float().contiguous()
    n, m = a.shape[0], b.shape[0]
    iou = torch.empty((n, m), dtype=torch.float32, device=a.device)
    uni = torch.empty((n, m), dtype=torch.float32, device=a.device)
    giou = torch.empty((n, m), dtype=torch.float32, device=a.device) if want_giou else None
    if n and m:
        check(lib().spe_box_iou_pairwise(ptr(a), n, ptr(b), m, ptr(iou), ptr(uni), ptr(giou), stream()))
    return iou, uni, giou


def box_iou(boxes1, boxes2):
    iou, uni, _ = _pairwise(boxes1, boxes2, False)
    return iou, uni


def generalized_box_iou(boxes1, boxes2):
    """[N,4] x [M,4] (xyxy) -> [N,M]; asserts well-formed boxes like the reference (:64-65)."""
    assert (boxes1[:, 2:] >= boxes1[:, :2]).all(), boxes1
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all(), boxes2
    return _pairwise(boxes1, boxes2, True)[2]
