"""The training iteration of the reference's engine (engine.py:train_one_epoch_refine, :100-170: forward -> criterion(outputs[0]) +
criterion_refine(outputs[1]) -> weighted sum -> backward) as ONE object, optionally captured into a CUDA graph.

Eager mode is what the reference's loop does line by line.  Graph mode (graph=True) records the same work once -- for a fixed
image shape and a fixed per-image target capacity -- and replays it: the ~3200 kernel launches of a step then cost the host one
cudaGraphLaunch, which is what keeps the step device-bound (eager, the Python/launch path needs ~57 ms per step next to
~70 ms of device work).  Inputs are copied into static buffers before each replay; the matcher, the losses and the gradient
accumulation all run on the device with no host synchronisation, so the captured step is the complete step.

Gradient all-reduce (N > 1): the flat buffer is reduced in TWO buckets issued from inside the step (and captured with it): the
detector-side slice (transformer, heads, query embeddings: 138 MB of 331 MB at cfg2) as soon as the backward pass reaches the backbone
tap -- it then runs on NCCL's stream under the backbone's backward --, the backbone slice after the last kernel.  SPE_AR_OVERLAP=0
falls back to one all-reduce after the step."""
import os

import torch

from . import criterion_ops as CO
from . import ops
from .dp import FlatGradBuffer
from .util.misc import NestedTensor


class TrainStep:
    def __init__(self, model, criterion, criterion_refine=None, weight_dict=None, grad_buffer=None, graph=False, max_gt=None,
                 refresh_shadows=True, optimizer=None):
        """optimizer: an optim.FlatAdamW built on `grad_buffer` BEFORE this object (it re-points p.data into its flat parameter
        buffer): clip_grad_norm_ + AdamW then run inside the step -- and inside its CUDA graph -- after the gradient all-reduce
        (engine.py:161-165 of the reference: zero_grad, backward, clip, step)."""
        self.optimizer = optimizer
        self.model, self.criterion, self.criterion_refine = model, criterion, criterion_refine
        self.weight_dict = dict(weight_dict if weight_dict is not None else criterion.weight_dict)
        self.gbuf = grad_buffer if grad_buffer is not None else FlatGradBuffer(model.parameters())
        self.graph = bool(graph)
        self.max_gt = max_gt
        self.shadows = ops.ShadowSet(model.parameters()) if refresh_shadows else None
        self._g = {}            # (B, H, W, cap) -> captured state
        self._side = None       # second stream for the refine criterion's matching
        # bucketed all-reduce: the flat buffer must hold every non-backbone parameter before the first backbone parameter
        self._split = None
        if os.environ.get("SPE_AR_OVERLAP", "1") != "0" and self.gbuf.mode == "views":
            names = [n for n, p in model.named_parameters() if p.requires_grad]
            first_bb = next((i for i, n in enumerate(names) if n.startswith("backbone")), None)
            if first_bb not in (None, 0) and all(n.startswith("backbone") for n in names[first_bb:]):
                pbb = dict(model.named_parameters())[names[first_bb]]
                self._split = self.gbuf.offset_of(pbb)
        self._works = []

    # ---- the step body (identical in both modes) ----
    def _dist(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1

    def _reduce_bucket(self, lo, hi):
        flat = self.gbuf.flat
        op = torch.distributed.ReduceOp.AVG if flat.is_cuda else torch.distributed.ReduceOp.SUM
        self._works.append(torch.distributed.all_reduce(flat[lo:hi], op=op, async_op=True))

    def _body(self, images, T, T_refine):
        self.gbuf.zero_()
        if self.shadows is not None:
            self.shadows.refresh()
        bucketed = self._split is not None and self._dist() and self.gbuf.flat.is_cuda
        self._works = []
        hooked = bucketed and os.environ.get("SPE_AR_OVERLAP", "1") != "2"       # "2": both buckets after the backward pass (no hook)
        ops.set_grad_milestone("detector_grads_done", (lambda: self._reduce_bucket(0, self._split)) if hooked else None)
        for X in (T, T_refine):                       # device-side GT jitter / repeat of the raw static targets (graph mode, training criteria)
            if X is not None and getattr(X, "_expand", None) is not None:
                X._expand()
        out = self.model(images)
        m1 = m2 = None
        packed = isinstance(T, CO.PackedTargets) and (T_refine is None or isinstance(T_refine, CO.PackedTargets))
        if packed and self.criterion_refine is not None:
            # the two criteria's Hungarian matchings are independent and each keeps only a few dozen CTAs busy: run them side by side
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                m2 = self.criterion_refine.match_all_levels(out[1], T_refine)
            m1 = self.criterion.match_all_levels(out[0], T)
            cur.wait_stream(self._side)
            if not torch.cuda.is_current_stream_capturing():
                m2.record_stream(cur)
        ld = self.criterion(out[0], T, _matches=m1)
        wd = self.weight_dict
        loss = sum(ld[k] * wd[k] for k in ld if k in wd)
        ld2 = None
        if self.criterion_refine is not None:
            ld2 = self.criterion_refine(out[1], T_refine, _matches=m2)
            loss = loss + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
        loss.backward()
        ops.set_grad_milestone("detector_grads_done", None)
        if bucketed:
            if not self._works:                       # the milestone did not fire (model without the tap): reduce everything now
                self._reduce_bucket(0, self._split)
            self._reduce_bucket(self._split, self.gbuf.flat.numel())
            for w in self._works:
                w.wait()                              # the step's stream waits for NCCL's stream (captured with the step in graph mode)
            self._works = []
            self._reduced = True
        else:
            self._reduced = False
        self._stepped = False
        if self.optimizer is not None and (self._reduced or not self._dist()):
            if getattr(self, "_warming", False):      # capture warm-up passes must not move the parameters / Adam state:
                if self.optimizer.shadows is not None:    # only settle the one-time host work of the step (shadow descriptor table)
                    self.optimizer.shadows.refresh()
            else:
                self.optimizer.step()                 # gradients are final here: part of the (captured) step
            self._stepped = True
        return loss.detach(), {k: v.detach() for k, v in ld.items()}, (None if ld2 is None else {k: v.detach() for k, v in ld2.items()})

    def __call__(self, samples, targets, targets_refine=None):
        """samples: float tensor [B,3,H,W] (or NestedTensor / list, eager only); targets: list of dicts as the reference's
        data loader yields them; targets_refine: the pseudo-label targets of the refine stage (default: targets).
        Returns (loss, loss_dict, loss_dict_refine) -- device scalars; gradients are in the flat buffer, all-reduced."""
        if targets_refine is None:
            targets_refine = targets
        if not self.graph or not torch.is_tensor(samples):
            # jitter / repeat ONCE (conditional_detr.py:410-431), then pack: criterion.forward prepares plain lists itself and
            # would expand them a second time (every GT 25x instead of 5x with match_ratio 5)
            dev = next(self.model.parameters()).device
            tg = self.criterion.prepare_packed(targets, dev)
            tr = self.criterion_refine.prepare_packed(targets_refine, dev) if self.criterion_refine is not None else None
            res = self._body(samples, tg, tr)
        else:
            res = self._replay(samples, targets, targets_refine)
        if not getattr(self, "_reduced", False):
            self.gbuf.all_reduce_mean()
        if self.optimizer is not None and not self._stepped:
            self.optimizer.step()                     # single all-reduce after the step (SPE_AR_OVERLAP=0): the update follows it
        return res

    def prefetch(self, images):
        """Start the host-to-device copy of the NEXT step's image batch (a pinned host tensor) on a copy stream, so that it overlaps
        the step that is running; the next __call__ with the same tensor object picks the staged copy up (graph mode).  Input
        pipelining, what a data loader with pin_memory + non_blocking does for the reference loop."""
        if not (torch.is_tensor(images) and not images.is_cuda and images.is_pinned()):
            return
        dev = next(self.model.parameters()).device
        if self.__dict__.get("_copy_stream") is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [None, None]
            self._stage_i = 0
        i = self._stage_i = self._stage_i ^ 1
        if self._stage[i] is None or self._stage[i].shape != images.shape:
            self._stage[i] = torch.empty(images.shape, dtype=torch.float32, device=dev)
        ev = torch.cuda.Event()
        if self.__dict__.get("_stage_read") is not None:
            self._copy_stream.wait_event(self._stage_read)
        with torch.cuda.stream(self._copy_stream):
            self._stage[i].copy_(images, non_blocking=True)
            ev.record()
        self._prefetched = (images, self._stage[i], ev)

    def close(self):
        """Drop the captured graphs.  With N > 1 they hold the gradient all-reduce: NCCL keeps a communicator alive (and
        destroy_process_group() waiting) for as long as a captured collective exists, so call this before tearing the group down."""
        self._g.clear()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    # ---- graph mode ----
    def _device_jitter(self):
        """training-mode criteria with device_jitter: the GT jitter / repeat (conditional_detr.py:410-431) runs inside the step (and its
        graph) on raw targets; otherwise the targets are prepared on the host before they are packed."""
        cs = [c for c in (self.criterion, self.criterion_refine) if c is not None]
        return all(c.training and getattr(c, "device_jitter", False) for c in cs)

    def _replay(self, images, targets, targets_refine):
        dj = self._device_jitter()
        crits = [self.criterion] + ([self.criterion_refine] if self.criterion_refine is not None else [])
        if dj:
            tg, tr = targets, (targets_refine if self.criterion_refine is not None else None)
            rs = [c.hung_match_ratio for c in crits]
            need = max([len(t["labels"]) * rs[0] for t in tg] + ([len(t["labels"]) * rs[-1] for t in tr] if tr is not None else []) + [1])
        else:
            tg = self.criterion.prepare_targets(targets)
            tr = self.criterion_refine.prepare_targets(targets_refine) if self.criterion_refine is not None else None
            rs = [1, 1]
            need = max([len(t["labels"]) for t in tg] + ([len(t["labels"]) for t in tr] if tr is not None else []) + [1])
        cap = self.max_gt if self.max_gt is not None else need
        if need > cap:
            raise ValueError("TrainStep(graph=True): %d targets in one image exceed max_gt=%d" % (need, cap))
        key = (tuple(images.shape), cap, dj, tuple(rs))
        st = self._g.get(key)
        if st is None:
            st = self._capture(images, tg, tr, cap, rs if dj else None)
            self._g[key] = st
        pre = self.__dict__.pop("_prefetched", None)
        if pre is not None and pre[0] is images and pre[1].shape == st["images"].shape:
            # the batch was staged by prefetch() on the copy stream while the previous step ran: a device-to-device copy is left
            self._prefetch_hits = self.__dict__.get("_prefetch_hits", 0) + 1
            torch.cuda.current_stream().wait_event(pre[2])
            st["images"].copy_(pre[1], non_blocking=True)
            self._stage_read = torch.cuda.Event()
            self._stage_read.record()                 # a later prefetch may overwrite the staging buffers only after this copy
        else:
            st["images"].copy_(images, non_blocking=True)
        for name, raw, t, r in (("T", "Traw", tg, rs[0]), ("Tr", "Trraw", tr, rs[-1])):
            if st[name] is None:
                continue
            if dj:                                           # raw targets in; the captured jitter kernel fills st[name] every replay
                st[raw].update(t, sync_num_boxes=False)
                st[name].sizes, st[name].total = [n * r for n in st[raw].sizes], st[raw].total * r
                st[name].set_num_boxes(sync_num_boxes=False)
            else:
                st[name].update(t, sync_num_boxes=False)
        CO.sync_num_boxes([st["T"], st["Tr"]])          # one 8-byte all-reduce for both criteria (N > 1)
        st["graph"].replay()
        return st["out"]

    def _capture(self, images, tg, tr, cap, ratios=None):
        dev = next(self.model.parameters()).device
        B = images.shape[0]
        ncls = tg[0]["img_label"].numel() if tg and "img_label" in tg[0] else 0
        st = {"images": torch.empty(images.shape, dtype=torch.float32, device=dev),
              "T": CO.StaticTargets(B, cap, dev, with_scores=all("scores" in t for t in tg), img_classes=ncls),
              "Tr": None, "Traw": None, "Trraw": None}
        if tr is not None:
            st["Tr"] = CO.StaticTargets(B, cap, dev, with_scores=all("scores" in t for t in tr), img_classes=ncls)
        st["images"].copy_(images)
        if ratios is None:
            st["T"].update(tg)
            if st["Tr"] is not None:
                st["Tr"].update(tr)
        else:
            for name, raw, t, r, crit in (("T", "Traw", tg, ratios[0], self.criterion), ("Tr", "Trraw", tr, ratios[-1], self.criterion_refine)):
                if st[name] is None:
                    continue
                R = CO.StaticTargets(B, max(cap // r, 1), dev, with_scores=st[name].scores is not None, img_classes=ncls)
                R.update(t)
                st[raw] = R
                T = st[name]
                T.img_label = R.img_label
                T._expand = (lambda R=R, T=T, r=r, crit=crit: CO.jitter_repeat(R, r, crit.box_jitter, crit.jitter_rng(dev), out=T))
                T.sizes, T.total = [n * r for n in R.sizes], R.total * r
                T.set_num_boxes()
        # warm-up on a side stream (lazy one-time initialisation: shadows, kernel attributes, autograd buffers), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._warming = True
            try:
                for _ in range(2):
                    self._body(st["images"], st["T"], st["Tr"])
            finally:
                self._warming = False
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st["out"] = self._body(st["images"], st["T"], st["Tr"])
        st["graph"] = g
        return st
