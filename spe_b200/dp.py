"""Data-parallel host logic: one flat fp32 gradient buffer whose slices ARE the parameters' .grad,
so a training step needs exactly one all-reduce (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class FlatGradBuffer:
    """mode 'views'  : p.grad ARE slices of the flat buffer.  With fused_accumulate (default) the spe_b200 backward kernels add
                       their results straight into the slices (wgrad GEMM reduce-add, atomics) and autograd sees None for those
                       parameters; otherwise autograd accumulates (one temporary + one add kernel per parameter);
       mode 'gather' : autograd produces free-standing gradients (no per-parameter add), `gather()` packs them into the flat
                       buffer with ONE multi-tensor copy before the all-reduce, and re-points p.grad at the reduced slices."""

    def __init__(self, params, align=8, mode="views", fused_accumulate=True):
        self.params = [p for p in params if p.requires_grad]
        self.mode = mode
        sizes = [(p.numel() + align - 1) // align * align for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p, s in zip(self.params, sizes):
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += s
        if mode == "views":
            for p, v in zip(self.params, self.views):
                p.grad = v
                p._spe_accum = fused_accumulate     # ops.grad_sink: backward kernels accumulate straight into the slice

    def _repoint(self, fold):
        """'views' mode: p.grad must BE the flat-buffer slice.  optimizer.zero_grad() (set_to_none=True is torch's default and what
        the reference loop calls, engine.py:77/161) drops it; with `fold`, gradients that autograd then created free-standing are
        added back into the slice (otherwise the all-reduce would ship zeros and the ranks would silently stop syncing)."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None or g.data_ptr() != v.data_ptr():
                if fold and g is not None:
                    v.add_(g.to(v.dtype).view_as(v))
                p.grad = v

    def offset_of(self, param):
        """element offset of `param`'s slice in the flat buffer (None if it is not in the buffer)"""
        for p, v in zip(self.params, self.views):
            if p is param:
                return (v.data_ptr() - self.flat.data_ptr()) // self.flat.element_size()
        return None

    def zero_(self):
        if self.mode == "views":
            self.flat.zero_()
            self._repoint(fold=False)
        else:
            for p in self.params:
                p.grad = None

    def gather(self):
        if self.mode != "gather":
            return
        src, dst, missing = [], [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                missing.append(v)
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad); dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        if missing:
            torch._foreach_zero_(missing)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self):
        """grad <- mean over ranks (what DDP does, main.py:171-173).  No-op when not distributed."""
        if self.mode == "views":
            self._repoint(fold=True)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.gather()
            if self.flat.is_cuda:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self.flat)
                self.flat.div_(dist.get_world_size())
        return self.flat
