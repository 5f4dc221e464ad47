"""Optimizer step of the reference loop on flat buffers:  clip_grad_norm_(model.parameters(), max_norm)  (engine.py:163-164)
followed by  torch.optim.AdamW  with the reference's three parameter groups (main.py:177-190):
    group 0  every parameter without "backbone" in its name          lr
    group 1  backbone parameters except blocks_token_only            lr_backbone
    group 2  backbone.*.blocks_token_only.*                          lr_cls_head
The gradients already are one fp32 buffer (dp.FlatGradBuffer); FlatAdamW lays parameters, exp_avg and exp_avg_sq out the same way
(each p.data becomes a view of the flat parameter buffer) so a step is 4 kernel launches instead of ~10 per parameter tensor, with no
host synchronisation -- it can be captured in the CUDA graph of engine.TrainStep(optimizer=...)."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream


class FlatAdamW:
    def __init__(self, model, gbuf, lr=1e-4, lr_backbone=1e-5, lr_cls_head=None, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 clip_max_norm=0.1, write_clipped_grads=False, bf16_shadow=False, refresh_shadows=True):
        """refresh_shadows: re-cast the bf16 weight shadows the GEMMs read (ops.shadow) right after the update, one launch -- the
        in-place kernel does not bump the parameters' autograd version counters, so the lazy refresh in ops.shadow cannot see it."""
        assert gbuf.mode == "views", "FlatAdamW needs FlatGradBuffer(mode='views')"
        names = {id(p): n for n, p in model.named_parameters()}
        self.gbuf = gbuf
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.clip_max_norm = float(clip_max_norm) if clip_max_norm else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        lr_cls_head = lr_backbone if lr_cls_head is None else lr_cls_head
        self.param_groups = [{"lr": float(lr), "initial_lr": float(lr), "weight_decay": float(weight_decay), "name": "detector"},
                             {"lr": float(lr_backbone), "initial_lr": float(lr_backbone), "weight_decay": float(weight_decay), "name": "backbone"},
                             {"lr": float(lr_cls_head), "initial_lr": float(lr_cls_head), "weight_decay": float(weight_decay), "name": "blocks_token_only"}]
        flat = gbuf.flat
        dev = flat.device
        n = flat.numel()
        assert n % 4 == 0
        # parameters into one buffer with the gradient buffer's layout; p.data re-pointed at the slices
        self.pflat = torch.zeros_like(flat)
        segs = []                                   # (end offset, group) runs
        for p, gv in zip(gbuf.params, gbuf.views):
            off = (gv.data_ptr() - flat.data_ptr()) // 4
            name = names.get(id(p), "")
            grp = 0 if "backbone" not in name else (2 if "blocks_token_only" in name else 1)
            pv = self.pflat[off:off + p.numel()].view_as(p)
            pv.copy_(p.data)
            p.data = pv
            p.__dict__.pop("_spe_shadow", None)     # bf16 shadows are keyed by storage offset: rebuilt on first use
            if segs and segs[-1][1] == grp:
                segs[-1][0] = off + p.numel()
            else:
                if segs:
                    segs[-1][0] = off               # the alignment padding before `off` belongs to the previous run
                segs.append([off + p.numel(), grp])
        segs[-1][0] = n
        assert len(segs) <= 16, "more than 16 parameter-group runs in the flat buffer"
        assert all(e % 4 == 0 for e, _ in segs)
        self.segments = [(int(e), int(g)) for e, g in segs]
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=dev) if bf16_shadow else None
        self.state = torch.tensor([1.0, 1.0, 1.0, 0.0], dtype=torch.float32, device=dev)        # beta1^t, beta2^t, clip coef, t
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws = torch.empty(int(_lib.lib().spe_sumsq_workspace_floats()), dtype=torch.float32, device=dev)
        self._lr_dev = torch.zeros(3, dtype=torch.float32, device=dev)
        self._wd_dev = torch.zeros(3, dtype=torch.float32, device=dev)
        self._lr_host = None
        self._push_hparams()
        a = _lib.AdamwArgs()
        a.p, a.g, a.m, a.v, a.n = ptr(self.pflat), ptr(flat), ptr(self.exp_avg), ptr(self.exp_avg_sq), n
        a.nseg = len(self.segments)
        for i, (e, g) in enumerate(self.segments):
            a.seg_end[i], a.seg_group[i] = e, g
        a.lr, a.wd = ptr(self._lr_dev), ptr(self._wd_dev)
        a.beta1, a.beta2, a.eps = self.betas[0], self.betas[1], self.eps
        a.state = ptr(self.state)
        a.g_out = ptr(flat) if self.write_clipped_grads else None
        a.shadow_bf16 = ptr(self.shadow) if self.shadow is not None else None
        self._args = a
        from .ops import ShadowSet
        self.shadows = ShadowSet(gbuf.params) if refresh_shadows else None

    # learning rates / weight decays live on the device (graph replays read them); pushed when the host copy changed (StepLR etc.)
    def _push_hparams(self):
        cur = ([g["lr"] for g in self.param_groups], [g["weight_decay"] for g in self.param_groups])
        if cur != self._lr_host:
            self._lr_dev.copy_(torch.tensor(cur[0], dtype=torch.float32))
            self._wd_dev.copy_(torch.tensor(cur[1], dtype=torch.float32))
            self._lr_host = cur

    def step_lr(self, epoch, lr_drop, gamma=0.1):
        """torch.optim.lr_scheduler.StepLR(optimizer, lr_drop) (main.py:191): lr = initial_lr * gamma ** (epoch // lr_drop)."""
        for g in self.param_groups:
            g["lr"] = g["initial_lr"] * gamma ** (epoch // lr_drop)

    def zero_grad(self, set_to_none=False):
        self.gbuf.zero_()

    @torch.no_grad()
    def step(self):
        """clip (if clip_max_norm > 0) + AdamW on the current stream.  Returns the device scalar sum of squared gradients before
        clipping (total_norm ** 2) or None."""
        L = _lib.lib()
        if not torch.cuda.is_current_stream_capturing():
            self._push_hparams()
        flat = self.gbuf.flat
        clip = self.clip_max_norm > 0
        if clip:
            check(L.spe_sumsq_f32(ptr(flat), flat.numel(), ptr(self._sumsq), ptr(self._ws), stream()))
        check(L.spe_adamw_tick(ptr(self.state), self.betas[0], self.betas[1], ptr(self._sumsq) if clip else None, self.clip_max_norm, stream()))
        check(L.spe_adamw_flat(C.byref(self._args), stream()))
        if self.shadows is not None:
            self.shadows.refresh()
        return self._sumsq if clip else None

    def total_norm(self):
        """gradient norm seen by the last step (before clipping) -- synchronises"""
        return float(self._sumsq.sqrt())

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "state": self.state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.state.copy_(sd["state"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        self._lr_host = None
        self._push_hparams()


@torch.no_grad()
def clip_grad_norm_(gbuf, max_norm, _cache={}):
    """torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) (engine.py:163-164) on the flat gradient buffer: 4 launches, no
    host synchronisation.  Returns total_norm as a device scalar."""
    L = _lib.lib()
    flat = gbuf.flat
    key = (flat.device, flat.data_ptr())
    if key not in _cache:
        _cache[key] = (torch.zeros(1, dtype=torch.float32, device=flat.device),
                       torch.empty(int(L.spe_sumsq_workspace_floats()), dtype=torch.float32, device=flat.device),
                       torch.zeros(4, dtype=torch.float32, device=flat.device))
    sumsq, ws, state = _cache[key]
    check(L.spe_sumsq_f32(ptr(flat), flat.numel(), ptr(sumsq), ptr(ws), stream()))
    state.fill_(1.0)
    check(L.spe_adamw_tick(ptr(state), 1.0, 1.0, ptr(sumsq), float(max_norm), stream()))
    check(L.spe_scale_by_clip_coef(ptr(flat), flat.numel(), ptr(state), stream()))
    return sumsq.sqrt()
