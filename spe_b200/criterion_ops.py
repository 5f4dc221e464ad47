"""Device-side matcher + set-criterion plumbing over the C ABI (spe_match_cost, spe_lsap_batched,
spe_focal_loss, spe_box_loss, spe_bce_logits).  Ragged targets are packed once per step into
(labels, boxes, offsets) arrays on the device; the assignment stays on the device as a
dense query->gt map, so the criterion needs no host synchronisation.  Index lists in the reference's
format (CPU int64 pairs, matcher.py:87) are produced only on request (HungarianMatcher.forward)."""
import torch

from ._lib import check, lib, ptr, stream


class PackedTargets:
    """labels i32 [sumG], boxes f32 [sumG,4], offsets i32 [B+1], optional scores f32 [sumG] (all on device)."""

    def __init__(self, targets, device):
        sizes = [int(t["labels"].shape[0]) for t in targets]
        self.sizes = sizes
        self.B = len(targets)
        self.total = sum(sizes)
        self.max_g = max(sizes) if sizes else 0
        off = [0]
        for s in sizes:
            off.append(off[-1] + s)
        pin = torch.cuda.is_available()
        self.offsets_host = torch.tensor(off, dtype=torch.int32)
        self.counts_host = torch.tensor(sizes, dtype=torch.int32)
        if self.total > 0:
            labels = torch.cat([t["labels"].reshape(-1) for t in targets]).to(torch.int32)
            boxes = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(torch.float32)
        else:
            labels = torch.zeros(1, dtype=torch.int32)
            boxes = torch.zeros(1, 4, dtype=torch.float32)
        self.labels = labels.to(device, non_blocking=True).contiguous()
        self.boxes = boxes.to(device, non_blocking=True).contiguous()
        self.offsets = self.offsets_host.to(device, non_blocking=True)
        self.counts = self.counts_host.to(device, non_blocking=True)
        self.scores = None
        if targets and all("scores" in t for t in targets) and self.total > 0:
            self.scores = torch.cat([t["scores"].reshape(-1) for t in targets]).to(torch.float32).to(device, non_blocking=True).contiguous()
        self.device = device
        self._inv_num_boxes = None
        self.img_label = None
        if targets and all("img_label" in t for t in targets):
            self.img_label = torch.stack([t["img_label"] for t in targets]).to(device, non_blocking=True).float()

    def inv_num_boxes(self):
        """1 / clamp(all_reduce(num_boxes)/world, 1) as a device scalar (conditional_detr.py:436-440), no .item()."""
        if self._inv_num_boxes is None:
            nb = torch.tensor([float(self.total)], dtype=torch.float32).to(self.device, non_blocking=True)
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                torch.distributed.all_reduce(nb)
                nb = nb / torch.distributed.get_world_size()
            self._inv_num_boxes = 1.0 / torch.clamp(nb, min=1.0)
        return self._inv_num_boxes


class StaticTargets(PackedTargets):
    """Fixed-capacity target buffers at fixed device addresses (what a captured CUDA graph needs): `update(targets)` refills
    them in place from pinned staging memory, outside the graph.  Layout is the same compact CSR (labels, boxes, offsets);
    `max_g` (the cost-matrix pitch) is the static per-image capacity, the real counts live on the device."""

    def __init__(self, batch, max_gt, device, with_scores=False, img_classes=0):
        self.B = int(batch)
        self.cap = int(max_gt)
        self.max_g = self.cap
        self.device = device
        ct = max(self.B * self.cap, 1)
        self.cap_total = ct
        pin = torch.cuda.is_available()
        self._h_labels = torch.zeros(ct, dtype=torch.int32, pin_memory=pin)
        self._h_boxes = torch.zeros(ct, 4, dtype=torch.float32, pin_memory=pin)
        self._h_scores = torch.ones(ct, dtype=torch.float32, pin_memory=pin) if with_scores else None
        self._h_off = torch.zeros(self.B + 1, dtype=torch.int32, pin_memory=pin)
        self._h_cnt = torch.zeros(self.B, dtype=torch.int32, pin_memory=pin)
        self._h_inv = torch.ones(1, dtype=torch.float32, pin_memory=pin)
        self.labels = torch.zeros(ct, dtype=torch.int32, device=device)
        self.boxes = torch.zeros(ct, 4, dtype=torch.float32, device=device)
        self.scores = torch.ones(ct, dtype=torch.float32, device=device) if with_scores else None
        self.offsets = torch.zeros(self.B + 1, dtype=torch.int32, device=device)
        self.counts = torch.zeros(self.B, dtype=torch.int32, device=device)
        self._inv_num_boxes = torch.ones(1, dtype=torch.float32, device=device)
        self.img_label = torch.zeros(self.B, img_classes, dtype=torch.float32, device=device) if img_classes else None
        self.sizes, self.total = [0] * self.B, 0
        # the pinned staging buffers are reused every step: the previous step's non_blocking copies must have been consumed
        # before the host overwrites them (the host runs ahead of the graph replays)
        self._staged = torch.cuda.Event() if pin else None
        self._staged_pending = False

    def update(self, targets, sync_num_boxes=True):
        """sync_num_boxes=False: the caller all-reduces num_boxes for several StaticTargets at once (sync_num_boxes below)."""
        assert len(targets) == self.B, "StaticTargets: batch size changed"
        sizes = [int(t["labels"].shape[0]) for t in targets]
        if max(sizes, default=0) > self.cap:
            raise ValueError("StaticTargets: %d targets in one image exceed the static capacity %d" % (max(sizes), self.cap))
        self.sizes, self.total = sizes, sum(sizes)
        if self._staged_pending:
            self._staged.synchronize()
            self._staged_pending = False
        tot, o = self.total, 0
        for b, n in enumerate(sizes):
            self._h_off[b] = o
            self._h_cnt[b] = n
            o += n
        self._h_off[self.B] = o
        if tot:
            live = [t for t, n in zip(targets, sizes) if n]
            on_dev = live[0]["labels"].is_cuda

            def fill(dst, host, parts):
                src = torch.cat(parts)
                if on_dev:                                   # already resident: device-side pack, no host round trip
                    dst[:tot].copy_(src)
                else:
                    host[:tot].copy_(src)
                    dst[:tot].copy_(host[:tot], non_blocking=True)

            fill(self.labels, self._h_labels, [t["labels"].reshape(-1) for t in live])
            fill(self.boxes, self._h_boxes, [t["boxes"].reshape(-1, 4) for t in live])
            if self.scores is not None:
                fill(self.scores, self._h_scores, [t["scores"].reshape(-1) for t in live])
        self.offsets.copy_(self._h_off, non_blocking=True)
        self.counts.copy_(self._h_cnt, non_blocking=True)
        if self.img_label is not None and targets and "img_label" in targets[0]:
            self.img_label.copy_(torch.stack([t["img_label"] for t in targets]).float(), non_blocking=True)
        self.set_num_boxes(sync_num_boxes, staged=True)
        if self._staged is not None:
            self._staged.record()
            self._staged_pending = True
        return self

    def set_num_boxes(self, sync_num_boxes=True, staged=False):
        """1 / clamp(all_reduce(num_boxes) / world, 1)   (conditional_detr.py:436-440) from the host-side count self.total, refreshed in
        place.  staged: called from update(), which guards the pinned staging word with its event; other callers copy from a fresh
        pageable scalar."""
        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        if dist_on and not sync_num_boxes:
            pass
        elif dist_on:
            nb = torch.tensor([float(self.total)], dtype=torch.float32, device=self.device)
            torch.distributed.all_reduce(nb)
            self._inv_num_boxes.copy_(1.0 / torch.clamp(nb / torch.distributed.get_world_size(), min=1.0))
        else:
            inv = 1.0 / max(float(self.total), 1.0)
            if staged:
                self._h_inv[0] = inv
                self._inv_num_boxes.copy_(self._h_inv, non_blocking=True)
            else:
                self._inv_num_boxes.copy_(torch.tensor([inv], dtype=torch.float32))


class JitterRng:
    """Device-side generator state of spe_gt_jitter_repeat: u64 {seed, launch counter} + the kernel's ticket word."""

    def __init__(self, device, seed=None):
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # follows torch.manual_seed like the reference's uniform_ draws
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)


def jitter_repeat(T, ratio, jitter, rng, n_try=1000, iou_thr=0.7, out=None):
    """conditional_detr.py:410-431 on packed device targets: every GT box -> `ratio` rows (up to ratio-1 accepted jittered copies, then
    the box itself), labels / scores repeated.  Returns a PackedTargets over new device arrays (`out`: write into an existing object's
    arrays -- StaticTargets).  No host synchronisation."""
    if not T.boxes.is_cuda:
        raise RuntimeError("spe_b200 GT jitter needs CUDA tensors (no CPU fallback exists)")
    dev = T.boxes.device
    cap = getattr(T, "cap_total", None) or max(T.total, 1)
    if out is None:
        out = PackedTargets.__new__(PackedTargets)
        out.B, out.device = T.B, dev
        out.labels = torch.zeros(cap * ratio, dtype=torch.int32, device=dev)
        out.boxes = torch.zeros(cap * ratio, 4, dtype=torch.float32, device=dev)
        out.scores = torch.ones(cap * ratio, dtype=torch.float32, device=dev) if T.scores is not None else None
        out.offsets = torch.zeros(T.B + 1, dtype=torch.int32, device=dev)
        out.counts = torch.zeros(T.B, dtype=torch.int32, device=dev)
        out.img_label = T.img_label
        out._inv_num_boxes = None
    out.sizes = [n * ratio for n in T.sizes]
    out.total = T.total * ratio
    out.max_g = (T.cap * ratio) if isinstance(T, StaticTargets) else (max(out.sizes) if out.sizes else 0)
    out.offsets_host = None
    out.counts_host = None
    check(lib().spe_gt_jitter_repeat(ptr(T.boxes), ptr(T.labels), ptr(T.scores) if T.scores is not None else None, ptr(T.offsets), T.B, cap, int(ratio),
                                     float(jitter), int(n_try), float(iou_thr), ptr(rng.state), ptr(out.boxes), ptr(out.labels),
                                     ptr(out.scores) if out.scores is not None else None, ptr(out.offsets), ptr(out.counts), ptr(rng.ticket), stream()))
    return out


def sync_num_boxes(static_targets):
    """ONE all-reduce for the num_boxes normalisers (conditional_detr.py:436-440) of all the step's criteria."""
    ts = [t for t in static_targets if t is not None]
    if not ts or not (torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
        return
    nb = torch.tensor([float(t.total) for t in ts], dtype=torch.float32, device=ts[0].device)
    torch.distributed.all_reduce(nb)
    inv = 1.0 / torch.clamp(nb / torch.distributed.get_world_size(), min=1.0)
    for i, t in enumerate(ts):
        t._inv_num_boxes.copy_(inv[i:i + 1])


def pack_targets(targets, device):
    return targets if isinstance(targets, PackedTargets) else PackedTargets(targets, device)


def match_cost(logits, boxes, T, weights):
    """cost f32 [B,Q,ld] (block-diagonal: image b uses its own G_b columns), matcher.py:62-82."""
    if not logits.is_cuda:
        raise RuntimeError("spe_b200 matcher needs CUDA tensors (no CPU fallback exists)")
    B, Q, C = logits.shape
    ld = max(T.max_g, 1)
    cost = torch.empty((B, Q, ld), dtype=torch.float32, device=logits.device)
    w_class, w_bbox, w_giou = weights
    check(lib().spe_match_cost(ptr(logits.detach().float().contiguous()), ptr(boxes.detach().float().contiguous()), ptr(T.labels), ptr(T.boxes),
                               ptr(T.offsets), B, Q, C, float(w_class), float(w_bbox), float(w_giou), ptr(cost), ld, 0, stream()))
    return cost


def lsap_raw(cost, ncols):
    """cost f32 [B,nr,ld]; ncols i32 [B] device or None -> row_to_col i32 [B,nr]."""
    B, nr, ld = cost.shape
    out = torch.empty((B, nr), dtype=torch.int32, device=cost.device)
    check(lib().spe_lsap_batched(ptr(cost), B, nr, ld, ptr(ncols), ld, ptr(out), 0, stream()))
    return out


def lsap(cost, T):
    if T.max_g == 0:
        return torch.full(cost.shape[:2], -1, dtype=torch.int32, device=cost.device)
    return lsap_raw(cost, T.counts)


def match(logits, boxes, T, weights):
    """dense query->gt assignment i32 [B,Q] (-1 = unmatched), on device."""
    return lsap(match_cost(logits, boxes, T, weights), T)


def match_levels(logits_levels, boxes_levels, T, weights):
    """logits [L,B,Q,C], boxes [L,B,Q,4] -> dense assignment i32 [L,B,Q]; one cost + one LSAP launch for all levels
    (L*B independent problems; problem p uses the targets of image p % B)."""
    L, B, Q, C = logits_levels.shape
    if T.max_g == 0:
        return torch.full((L, B, Q), -1, dtype=torch.int32, device=logits_levels.device)
    ld = T.max_g
    cost = torch.empty((L * B, Q, ld), dtype=torch.float32, device=logits_levels.device)
    w_class, w_bbox, w_giou = weights
    check(lib().spe_match_cost(ptr(logits_levels.detach().float().contiguous()), ptr(boxes_levels.detach().float().contiguous()), ptr(T.labels),
                               ptr(T.boxes), ptr(T.offsets), L * B, Q, C, float(w_class), float(w_bbox), float(w_giou), ptr(cost), ld, B, stream()))
    return lsap_raw(cost, T.counts.repeat(L)).view(L, B, Q)


def indices_from_dense(r2g_cpu):
    """reference format: list of (idx_pred int64, idx_gt int64), predictions ascending (= scipy's order, App. C)."""
    out = []
    for row in r2g_cpu:
        i = torch.nonzero(row >= 0).flatten()
        out.append((i.to(torch.int64), row[i].to(torch.int64)))
    return out


class FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, r2g, T, alpha, gamma, use_scores):
        B, Q, C = logits.shape
        lg = logits.contiguous()
        out = torch.empty(3, dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(lg)
        sc = T.scores if use_scores else None
        ws = torch.empty(lib().spe_focal_loss_workspace_bytes(B, Q), dtype=torch.uint8, device=logits.device)
        check(lib().spe_focal_loss(ptr(lg), ptr(r2g), ptr(T.labels), ptr(T.offsets), ptr(sc), ptr(T.inv_num_boxes()), B, Q, C, float(alpha),
                                   float(gamma), ptr(out), ptr(dl), ptr(ws), stream()))
        ctx.save_for_backward(dl)
        return out

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g[0], None, None, None, None, None


class BoxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes, r2g, T, use_scores):
        B, Q, _ = boxes.shape
        bx = boxes.contiguous()
        out = torch.empty(2, dtype=torch.float32, device=boxes.device)
        d1, d2 = torch.empty_like(bx), torch.empty_like(bx)
        sc = T.scores if use_scores else None
        check(lib().spe_box_loss(ptr(bx), ptr(r2g), ptr(T.boxes), ptr(T.offsets), ptr(sc), ptr(T.inv_num_boxes()), B, Q, ptr(out), ptr(d1), ptr(d2),
                                 stream()))
        ctx.save_for_backward(d1, d2)
        return out

    @staticmethod
    def backward(ctx, g):
        d1, d2 = ctx.saved_tensors
        return d1 * g[0] + d2 * g[1], None, None, None


class BceLogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        xc = x.contiguous()
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(xc)
        check(lib().spe_bce_logits(ptr(xc), ptr(y.contiguous()), xc.numel(), ptr(out), ptr(dx), stream()))
        ctx.save_for_backward(dx)
        return out

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g[0], None


def set_losses(logits, boxes, T, match_weights, alpha, gamma, refine=False, losses=("labels", "boxes", "cardinality"), log=True, r2g=None):
    """One decoder level: match + labels/boxes/cardinality losses; returns dict of 0-dim tensors (+ '_r2g')."""
    if r2g is None:
        r2g = match(logits, boxes, T, match_weights)
    if refine and T.scores is None and T.total > 0:
        # the reference indexes t['scores'] unconditionally (conditional_detr.py:523): a refine criterion without pseudo-label
        # scores is a caller error, not "weight 1"
        raise KeyError("scores: SetCriterionRefine needs a 'scores' entry in every target")
    res = {"_r2g": r2g}
    if "labels" in losses or "cardinality" in losses:
        f = FocalLossFn.apply(logits.float(), r2g, T, alpha, gamma, refine)
        if "labels" in losses:
            res["loss_ce"] = f[0]
            if log:
                res["class_error"] = f[1].detach()
        if "cardinality" in losses:
            res["cardinality_error"] = f[2].detach()
    if "boxes" in losses:
        b = BoxLossFn.apply(boxes.float(), r2g, T, refine)
        res["loss_bbox"] = b[0]
        res["loss_giou"] = b[1]
    return res
