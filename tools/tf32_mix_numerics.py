"""Design input for the fused talking-heads kernel: error of the pre-softmax head mix when the logits S and the H x H
weights are fed to the tensor core as tf32 (10-bit mantissa, truncated) -- single pass and hi/lo split -- against the error floor that the
bf16 Q / K inputs already put on S.  CPU / numpy only.   python tools/tf32_mix_numerics.py"""
import numpy as np

rng = np.random.default_rng(0)
H, N = 8, 1600


def trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


for sigma in (1.0, 3.0, 6.0):
    errs = {"tf32 single pass": [], "tf32 hi/lo": [], "bf16-input floor": []}
    for _ in range(20):
        S = rng.normal(0, sigma, (H, N))
        W = np.eye(H) + 0.3 * rng.normal(0, 1, (H, H)) / np.sqrt(H)
        V = rng.normal(0, 1, (N,))
        L = W @ S
        P = np.exp(L - L.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
        O = P @ V

        def run(Sx, Wx):
            Lx = Wx.astype(np.float64) @ Sx.astype(np.float64)
            Px = np.exp(Lx - Lx.max(1, keepdims=True)); Px /= Px.sum(1, keepdims=True)
            return np.abs(Px @ V - O).max() / np.abs(O).max()

        errs["tf32 single pass"].append(run(trunc_tf32(S), trunc_tf32(W)))
        Sh = trunc_tf32(S); Sl = trunc_tf32(S.astype(np.float32) - Sh)
        errs["tf32 hi/lo"].append(run(Sh.astype(np.float64) + Sl.astype(np.float64), W.astype(np.float32)))
        errs["bf16-input floor"].append(run(S * (1 + rng.uniform(-1, 1, S.shape) * 2 ** -9), W))
    print("logit std %.0f:" % sigma, {k: "%.1e" % np.mean(v) for k, v in errs.items()})
