"""HBM bandwidth for the access patterns of the talking-heads kernels: [B,H,N,ld] fp16 tensors touched row by row (b, q) over all heads
(8 pieces of ld*2 bytes, N*ld*2 bytes apart) versus contiguous streaming."""
import torch
dev = torch.device("cuda")
B, H, N = 8, 8, 1600
x = torch.randn(B, H, N, N, device=dev).half()
y = torch.empty_like(x)
xp = torch.empty(B, N, H, N, device=dev, dtype=torch.float16)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gb = 2 * x.numel() * 2 / 1e9
for name, fn in [("contiguous copy", lambda: y.copy_(x)),
                 ("strided read  -> contiguous write  ([B,H,N,ld] -> [B,N,H,ld])", lambda: xp.copy_(x.permute(0, 2, 1, 3))),
                 ("contiguous read -> strided write", lambda: y.permute(0, 2, 1, 3).copy_(xp)),
                 ("fp16 -> bf16 elementwise (x.to)", lambda: x.to(torch.bfloat16))]:
    ms = t(fn)
    print("%-70s %.3f ms  %.2f TB/s" % (name, ms, gb / ms))
big = torch.empty(1 << 30, device=dev, dtype=torch.bfloat16); big2 = torch.empty_like(big)
ms = t(lambda: big2.copy_(big), 10)
print("%-70s %.3f ms  %.2f TB/s" % ("2 GiB contiguous copy", ms, 2 * big.numel() * 2 / 1e9 / ms))
