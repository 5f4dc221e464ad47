"""Data-parallel host logic: one flat fp32 gradient buffer whose slices ARE the parameters' .grad,
so a training step needs exactly one all-reduce (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, params, align=8):
        self.params = [p for p in params if p.requires_grad]
        sizes = [(p.numel() + align - 1) // align * align for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        for p, s in zip(self.params, sizes):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += s

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """grad <- mean over ranks (what DDP does, main.py:171-173).  No-op when not distributed."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if self.flat.is_cuda:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self.flat)
                self.flat.div_(dist.get_world_size())
        return self.flat
