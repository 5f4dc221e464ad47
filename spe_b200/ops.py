"""Host-side operator layer: torch.autograd Functions whose forward AND backward are calls into the
C ABI of libspe_b200.so (sm_100a kernels).  PyTorch is used for device memory, streams and the
autograd tape only -- there is no eager/PyTorch compute fallback: every op raises if the native
library is missing or the tensors are not CUDA tensors.

Layout conventions (batch-first, token-major):
  residual streams / LayerNorm inputs : fp32 [B, N, D]
  GEMM operands (activations)         : bf16 [B, N, D]   (row stride may exceed D for packed views)
  attention logits S                  : fp32 [B, H, Lq, ld]   ld = Lk rounded up to 8
  attention probabilities             : bf16 [B, H, Lq, ld]
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream

MAJOR_K, MAJOR_MN = 0, 1
DT_BF16, DT_F32, DT_F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU, ACT_RELU_GRAD, ACT_GELU_GRAD = 0, 1, 2, 3, 4
_ACT = {None: ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU}
_ACT_GRAD = {"relu": ACT_RELU_GRAD, "gelu": ACT_GELU_GRAD}


_spe_gemm = None


def _bind_gemm():
    global _spe_gemm
    _spe_gemm = lib().spe_gemm
    return _spe_gemm


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spe_b200 ops need CUDA tensors (no CPU fallback exists)")


def rup(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# bf16 shadows of fp32 parameters (H4: reference-named fp32 nn.Parameters stay the source of truth)
# ------------------------------------------------------------------------------------------------
def shadow(p):
    """bf16 copy of an fp32 parameter (or of a slice view of one), refreshed when the parameter changes.
    The cache lives ON the (base) parameter object, so it can never be confused with another tensor that later
    reuses the same device address."""
    base = p._base if p._base is not None else p
    cache = base.__dict__.get("_spe_shadow")
    if cache is None:
        cache = {}
        try:
            base._spe_shadow = cache
        except Exception:
            pass
    key = (p.storage_offset(), tuple(p.shape), tuple(p.stride()))
    ver = base._version
    ent = cache.get(key)
    if ent is not None and ent[0] == ver and ent[1].device == p.device:
        return ent[1]
    src = p.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    out = ent[1] if ent is not None and ent[1].device == p.device else torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
    axpby_cast(src, None, 1.0, 0.0, out_bf16=out)
    cache[key] = (ver, out)
    return out


class ShadowSet:
    """All bf16 shadows of a set of parameters, refreshed by ONE multi-tensor cast launch (spe_cast_f32_to_bf16_multi) --
    what a training loop does right after optimizer.step() instead of ~400 single casts on first use.  Inside a captured
    CUDA graph the refresh is part of the graph, so replays always see the current fp32 masters."""

    def __init__(self, params):
        self.params = [p for p in params]
        self._sig = None
        self._desc = None
        self._entries = []
        self._max_n = 1

    def _collect(self):
        ents = []
        for p in self.params:
            cache = p.__dict__.get("_spe_shadow")
            if not cache:
                continue
            for key, (ver, out) in cache.items():
                off, shape, stride = key
                src = p.detach().as_strided(shape, stride, off)
                if src.is_contiguous() and out.is_contiguous() and out.device == p.device:
                    ents.append((p, key, src, out))
        return ents

    def refresh(self):
        """re-cast every known shadow from its fp32 master (no version check: one launch covers them all)."""
        ents = self._collect()
        if not ents:
            return 0
        sig = tuple((e[2].data_ptr(), e[3].data_ptr()) for e in ents)
        if sig != self._sig:
            import numpy as np
            tab = np.zeros((len(ents), 3), dtype=np.int64)
            for i, (_, _, src, out) in enumerate(ents):
                tab[i] = (src.data_ptr(), out.data_ptr(), src.numel())
            self._desc = torch.from_numpy(tab).to(ents[0][3].device)
            self._max_n = int(tab[:, 2].max())
            self._sig = sig
        self._entries = ents
        check(lib().spe_cast_f32_to_bf16_multi(ptr(self._desc), len(ents), self._max_n, stream()))
        for p, key, _, out in ents:
            p._spe_shadow[key] = (p._version, out)
        return len(ents)


def clear_shadows():
    """kept for API stability: shadows are owned by their parameters and die with them."""
    return None


# ------------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# ------------------------------------------------------------------------------------------------
def gemm(a, b, c, M, N, K, *, a_major=MAJOR_K, lda=None, b_major=MAJOR_K, ldb=None, ldc=None, batch=(1, 1),
         a_sb=(0, 0), b_sb=(0, 0), c_sb=(0, 0), alpha=1.0, bias=None, act=ACT_NONE, aux_in=None, aux_out=None, ld_aux=0,
         gamma=None, residual=None, ldr=0, r_sb=(0, 0), split=0, split_stride=0):
    if not (a.is_cuda and b.is_cuda and c.is_cuda):
        _need_cuda(a, b, c)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    # positional init in field order (one C call instead of ~35 attribute stores)
    g = GemmArgs(M, N, K, batch[0], batch[1],
                 a.data_ptr(), a_major, lda, a_sb[0], a_sb[1],
                 b.data_ptr(), b_major, ldb, b_sb[0], b_sb[1],
                 c.data_ptr(), (DT_F32 if c.dtype == torch.float32 else (DT_F16 if c.dtype == torch.float16 else DT_BF16)), ldc, c_sb[0], c_sb[1],
                 alpha, ptr(bias), act, ptr(aux_in), ptr(aux_out), ld_aux,
                 ptr(gamma), ptr(residual), ldr, r_sb[0], r_sb[1],
                 split, split_stride)
    rc = (_spe_gemm or _bind_gemm())(C.byref(g), stream())
    if rc:
        check(rc)
    return c


def axpby_cast(x, y, a, b, out_bf16=None, out_f32=None):
    _need_cuda(x)
    check(lib().spe_axpby_cast(ptr(x), ptr(y), a, b, x.numel(), ptr(out_bf16), ptr(out_f32), stream()))


def to_bf16(x):
    """fp32 -> bf16 cast kernel (contiguous)."""
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    axpby_cast(x, None, 1.0, 0.0, out_bf16=out)
    return out


def colsum_bf16(x2d, out):
    rows, n = x2d.shape
    check(lib().spe_colsum_bf16(ptr(x2d), rows, n, x2d.stride(0), ptr(out), stream()))


def _rows(t):
    return t.numel() // t.shape[-1]


# y[M,N] = x[M,K] @ w[N,K]^T (+ epilogue)
def _linear_fwd(x, w16, bias, out, **kw):
    M, K = _rows(x), x.shape[-1]
    N = w16.shape[0]
    return gemm(x, w16, out, M, N, K, lda=x.stride(-2) if x.dim() > 1 else K, ldb=w16.stride(0), ldc=kw.pop("ldc", N), bias=bias, **kw)


# dx[M,K] = dy[M,N] @ w[N,K]
def _linear_dgrad(dy, w16, out, **kw):
    M, N = _rows(dy), dy.shape[-1]
    K = w16.shape[1]
    return gemm(dy, w16, out, M, K, N, lda=dy.stride(-2), b_major=MAJOR_MN, ldb=w16.stride(0), ldc=kw.pop("ldc", K), **kw)


# dw[N,K] = dy[M,N]^T @ x[M,K]   (fp32 out)
def _linear_wgrad(dy, x, out, accumulate=False):
    M, N = _rows(dy), dy.shape[-1]
    K = x.shape[-1]
    return gemm(dy, x, out, N, K, M, a_major=MAJOR_MN, lda=dy.stride(-2), b_major=MAJOR_MN, ldb=x.stride(-2), ldc=K,
                residual=out if accumulate else None, ldr=K)


def grad_sink(p):
    """The fp32 gradient accumulator of parameter `p` when its owner opted in (dp.FlatGradBuffer 'views' mode sets
    p._spe_accum): backward kernels then accumulate straight into p.grad (wgrad GEMM reduce-adds its tiles, the column-sum
    kernels use atomics) and hand autograd None for p -- no temporary, no zero-fill, no AccumulateGrad add kernel.
    Anything else (no .grad yet, views of a packed parameter, foreign dtype) takes the ordinary autograd route."""
    if p is None or not p.is_leaf or not getattr(p, "_spe_accum", False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device or g.shape != p.shape:
        return None
    return g


def _grad_out(p, shape, dev):
    """(buffer to accumulate into, value to return to autograd)."""
    g = grad_sink(p)
    if g is not None:
        return g, None
    z = torch.zeros(shape, dtype=torch.float32, device=dev)
    return z, z


_SIDE = {}
_OVERLAP = None


def _overlap_wgrad():
    global _OVERLAP
    if _OVERLAP is None:
        import os
        _OVERLAP = os.environ.get("SPE_WGRAD_OVERLAP", "1") == "1"
    return _OVERLAP


class _SideStream:
    """Run the weight-gradient GEMM (+ bias column sum) of a layer on a second stream, next to the data-gradient GEMM of the same
    layer: the two are independent, each is a persistent kernel with one CTA per SM (or fewer: the decoder's GEMMs have 19-57
    tiles), so on one stream every launch pays its own ramp-up and tail (~8 us of a 15-25 us kernel); side by side the second
    kernel's CTAs start on the SMs the first one has already left.  Works eagerly and inside CUDA-graph capture (fork / join)."""

    def __init__(self):
        self.cur = torch.cuda.current_stream()
        dev = self.cur.device
        side = _SIDE.get(dev)
        if side is None:
            side = _SIDE[dev] = torch.cuda.Stream(device=dev)
        self.side = side
        self.ctx = None

    def __enter__(self):
        self.side.wait_stream(self.cur)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        self.cur.wait_stream(self.side)
        if not torch.cuda.is_current_stream_capturing():
            for t in tensors:
                if t is not None:
                    t.record_stream(self.cur)


def _wgrad_into(p, dy, x):
    g = grad_sink(p)
    if g is not None:
        _linear_wgrad(dy, x, g, accumulate=True)
        return None
    dw = torch.empty(p.shape, dtype=torch.float32, device=x.device)
    _linear_wgrad(dy, x, dw)
    return dw


def _pad_cols(t2d):
    """TMA needs a 16-byte row pitch: return a view of t2d [M,N] whose row stride is N rounded up to 8."""
    M, N = t2d.shape
    if N % 8 == 0 and t2d.stride(0) % 8 == 0:
        return t2d
    buf = torch.zeros((M, rup(N, 8)), dtype=t2d.dtype, device=t2d.device)
    buf[:, :N].copy_(t2d)
    return buf[:, :N]


# ------------------------------------------------------------------------------------------------
# autograd Functions
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """nn.Linear on bf16 activations:  out = [residual +] [gamma *] act(x W^T + b).
    `residual` (fp32) has the shape of the output, or the output shape without its leading batch dim
    (broadcast over the batch, e.g. the batch-invariant query_pos projections of transformer.py:368-372).
    Output is fp32 when out_f32 (always when gamma is given), else bf16.
    Replaces cait.py:376,390 / transformer.py projections / conditional_detr.py heads / MLP layers."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, gamma, out_f32, act):
        _need_cuda(x, weight)
        assert x.dtype == torch.bfloat16 and x.stride(-1) == 1
        x = x if x.is_contiguous() else x.contiguous()
        w16 = shadow(weight)
        N, Kd = weight.shape
        f32 = bool(out_f32) or gamma is not None
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32 if f32 else torch.bfloat16, device=x.device)
        y = None
        bcast = residual is not None and residual.dim() == x.dim() - 1
        if residual is not None:
            residual = residual.contiguous()
            assert residual.dtype == torch.float32
        if gamma is not None:
            assert not bcast and act is None
            y = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)   # pre-LayerScale branch (for dgamma)
        if bcast:
            Bt = x.shape[0]
            Mb = _rows(x) // Bt
            gemm(x, w16, out, Mb, N, Kd, lda=Kd, ldb=Kd, ldc=N, batch=(Bt, 1), a_sb=(Mb * Kd, 0), b_sb=(0, 0), c_sb=(Mb * N, 0), bias=bias,
                 act=_ACT[act], residual=residual, ldr=N, r_sb=(0, 0))
        else:
            _linear_fwd(x, w16, bias, out, gamma=gamma, residual=residual, ldr=N, aux_out=y, ld_aux=N, act=_ACT[act])
        ctx.save_for_backward(x, weight, gamma, y, out if act is not None else None, bias)
        ctx.has_bias = bias is not None
        ctx.res = None if residual is None else ("b" if bcast else "f")
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, gamma, y, h, bias = ctx.saved_tensors
        w16 = shadow(weight)
        N, K = weight.shape
        dout = dout.contiguous()
        dgamma = dbias = None
        dgamma_b = dbias_b = None
        dres = None
        late_colsum = False
        if ctx.res is not None:
            d32 = dout if dout.dtype == torch.float32 else _to_f32(dout)
            dres = d32 if ctx.res == "f" else d32.sum(0)
        if gamma is not None:
            dy = torch.empty(dout.shape, dtype=torch.bfloat16, device=dout.device)
            dgamma_b, dgamma = _grad_out(gamma, gamma.shape, dout.device)
            if ctx.has_bias:
                dbias_b, dbias = _grad_out(bias, (N,), dout.device)
            check(lib().spe_layerscale_bwd(ptr(dout), ptr(y), ptr(gamma), _rows(dout), N, ptr(dy), ptr(dgamma_b), ptr(dbias_b), stream()))
        else:
            dy = dout if dout.dtype == torch.bfloat16 else to_bf16(dout)
            if ctx.act is not None:
                assert ctx.act == "relu"
                h16 = h if h.dtype == torch.bfloat16 else to_bf16(h)
                dpre = torch.empty_like(dy)
                check(lib().spe_relu_bwd_bf16(ptr(dy), ptr(h16), ptr(dpre), dy.numel(), stream()))
                dy = dpre
            if ctx.has_bias:
                dbias_b, dbias = _grad_out(bias, (N,), dout.device)
                late_colsum = ctx.needs_input_grad[0] and _overlap_wgrad()
                if not late_colsum:
                    colsum_bf16(_pad_cols(dy.view(-1, N)), dbias_b)
        dy = _pad_cols(dy.view(-1, N))
        dx = None
        if ctx.needs_input_grad[0] and _overlap_wgrad():
            with _SideStream() as ss:                      # bias column sum + wgrad next to dgrad
                if late_colsum:
                    colsum_bf16(dy, dbias_b)
                dw = _wgrad_into(weight, dy, x)
            dx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            _linear_dgrad(dy, w16, dx)
            ss.join(dw, dbias)
            return dx, dw, dbias, dres, dgamma, None, None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            _linear_dgrad(dy, w16, dx)
        dw = _wgrad_into(weight, dy, x)
        return dx, dw, dbias, dres, dgamma, None, None


def linear(x, weight, bias=None, residual=None, gamma=None, out_f32=False, act=None):
    return LinearFn.apply(x, weight, bias, residual, gamma, out_f32, act)


def _to_f32(x16):
    out = torch.empty(x16.shape, dtype=torch.float32, device=x16.device)
    check(lib().spe_cast_bf16_to_f32(ptr(x16.contiguous()), ptr(out), x16.numel(), stream()))
    return out


class FfnFn(torch.autograd.Function):
    """out_f32 = residual + [gamma *] (act(x W1^T + b1) W2^T + b2).   timm Mlp (GELU, cait.py:415) and the DETR
    FFN (ReLU, transformer.py:285-286, 424).  The activation derivative is fused into the dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, gamma, act):
        _need_cuda(x, w1, w2)
        w1_16, w2_16 = shadow(w1), shadow(w2)
        Fh, D = w1.shape
        lead = x.shape[:-1]
        a = torch.empty(lead + (Fh,), dtype=torch.bfloat16, device=x.device) if act == "gelu" else None
        h = torch.empty(lead + (Fh,), dtype=torch.bfloat16, device=x.device)
        # one GEMM: C = act(x W1^T + b1); aux_out = the pre-activation (needed by gelu')
        _linear_fwd(x, w1_16, b1, h, act=_ACT[act], aux_out=a, ld_aux=Fh)
        out = torch.empty(lead + (w2.shape[0],), dtype=torch.float32, device=x.device)
        y = torch.empty(lead + (w2.shape[0],), dtype=torch.bfloat16, device=x.device) if gamma is not None else None
        _linear_fwd(h, w2_16, b2, out, gamma=gamma, residual=residual, ldr=w2.shape[0], aux_out=y, ld_aux=w2.shape[0])
        ctx.save_for_backward(x, w1, w2, gamma, a, h, y, b1, b2)
        ctx.act = act
        ctx.has_res = residual is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, w2, gamma, a, h, y, b1, b2 = ctx.saved_tensors
        w1_16, w2_16 = shadow(w1), shadow(w2)
        Fh, D = w1.shape
        Do = w2.shape[0]
        dout = dout.contiguous()
        dev = dout.device
        dgamma = None
        db2_b, db2 = _grad_out(b2, (Do,), dev)
        if gamma is not None:
            dy = torch.empty(dout.shape, dtype=torch.bfloat16, device=dev)
            dgamma_b, dgamma = _grad_out(gamma, gamma.shape, dev)
            check(lib().spe_layerscale_bwd(ptr(dout), ptr(y), ptr(gamma), _rows(dout), Do, ptr(dy), ptr(dgamma_b), ptr(db2_b), stream()))
        else:
            dy = to_bf16(dout)
            colsum_bf16(dy.view(-1, Do), db2_b)
        aux = a if ctx.act == "gelu" else h
        if _overlap_wgrad():
            with _SideStream() as ss:                      # dW2 next to the fc2 data gradient
                dw2 = _wgrad_into(w2, dy, h)
            da = torch.empty(h.shape, dtype=torch.bfloat16, device=dev)
            _linear_dgrad(dy, w2_16, da, act=_ACT_GRAD[ctx.act], aux_in=aux, ld_aux=Fh)
            ss.join(dw2)
            db1_b, db1 = _grad_out(b1, (Fh,), dev)
            with _SideStream() as ss:                      # db1, dW1 next to the fc1 data gradient
                colsum_bf16(da.view(-1, Fh), db1_b)
                dw1 = _wgrad_into(w1, da, x)
            dx = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
            _linear_dgrad(da, w1_16, dx)
            ss.join(dw1, db1)
            return dx, dw1, db1, dw2, db2, (dout if ctx.has_res else None), dgamma, None
        dw2 = _wgrad_into(w2, dy, h)
        da = torch.empty(h.shape, dtype=torch.bfloat16, device=dev)
        _linear_dgrad(dy, w2_16, da, act=_ACT_GRAD[ctx.act], aux_in=aux, ld_aux=Fh)
        db1_b, db1 = _grad_out(b1, (Fh,), dev)
        colsum_bf16(da.view(-1, Fh), db1_b)
        dw1 = _wgrad_into(w1, da, x)
        dx = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
        _linear_dgrad(da, w1_16, dx)
        return dx, dw1, db1, dw2, db2, (dout if ctx.has_res else None), dgamma, None


def ffn(x, w1, b1, w2, b2, residual=None, gamma=None, act="relu"):
    return FfnFn.apply(x, w1, b1, w2, b2, residual, gamma, act)


class LayerNormFn(torch.autograd.Function):
    """LayerNorm(x_f32) -> (y_f32 or None, y_bf16).  Backward sums the gradients of both outputs."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, want_f32):
        _need_cuda(x, weight, bias)
        x = x.contiguous()
        D = x.shape[-1]
        rows = _rows(x)
        y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        y32 = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_f32 else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(lib().spe_layernorm_fwd(ptr(x), ptr(weight), ptr(bias), eps, rows, D, ptr(y16), ptr(y32), ptr(mean), ptr(rstd), stream()))
        ctx.save_for_backward(x, weight, mean, rstd, bias)
        ctx.want_f32 = want_f32
        if want_f32:
            return y32, y16
        return y16

    @staticmethod
    def backward(ctx, *grads):
        x, weight, mean, rstd, bias = ctx.saved_tensors
        if ctx.want_f32:
            d32, d16 = grads
        else:
            d32, d16 = None, grads[0]
        D = x.shape[-1]
        if d32 is None and d16 is None:
            return None, None, None, None, None
        d32 = d32.contiguous() if d32 is not None else None
        d16 = d16.contiguous() if d16 is not None else None
        dx = torch.empty_like(x)
        dw_b, dw = _grad_out(weight, (D,), x.device)
        db_b, db = _grad_out(bias, (D,), x.device)
        check(lib().spe_layernorm_bwd(ptr(d16), ptr(d32), 0, ptr(x), ptr(weight), ptr(mean), ptr(rstd), _rows(x), D, ptr(dx), ptr(dw_b), ptr(db_b),
                                      stream()))
        return dx, dw, db, None, None


class LayerNormResFn(torch.autograd.Function):
    """Pre-norm residual block entry (cait.py:414-415: x + gamma * f(LN(x))):  x_f32 -> (LN(x) bf16, x).  The second output is x
    itself, handed on to the block's residual add; in the backward its gradient is added to the LayerNorm input gradient INSIDE
    the LayerNorm backward kernel (`dres`), instead of by a separate autograd accumulation kernel over the whole residual stream."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _need_cuda(x, weight, bias)
        x = x.contiguous()
        D = x.shape[-1]
        rows = _rows(x)
        y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(lib().spe_layernorm_fwd(ptr(x), ptr(weight), ptr(bias), eps, rows, D, ptr(y16), 0, ptr(mean), ptr(rstd), stream()))
        ctx.save_for_backward(x, weight, mean, rstd, bias)
        return y16, x.view(x.shape)

    @staticmethod
    def backward(ctx, d16, dres):
        x, weight, mean, rstd, bias = ctx.saved_tensors
        if d16 is None:
            return dres, None, None, None
        D = x.shape[-1]
        d16 = d16.contiguous()
        dres = dres.contiguous() if dres is not None else None
        dx = torch.empty_like(x)
        dw_b, dw = _grad_out(weight, (D,), x.device)
        db_b, db = _grad_out(bias, (D,), x.device)
        check(lib().spe_layernorm_bwd(ptr(d16), 0, ptr(dres), ptr(x), ptr(weight), ptr(mean), ptr(rstd), _rows(x), D, ptr(dx), ptr(dw_b), ptr(db_b),
                                      stream()))
        return dx, dw, db, None


def layernorm_res(x, weight, bias, eps):
    """returns (LN(x) bf16, x): use the second value as the residual operand of the block (see LayerNormResFn)."""
    return LayerNormResFn.apply(x, weight, bias, eps)


def layernorm(x, weight, bias, eps, want_f32=False):
    """returns y_bf16, or (y_f32, y_bf16) when want_f32."""
    return LayerNormFn.apply(x, weight, bias, eps, want_f32)


class AddCastFn(torch.autograd.Function):
    """(a_f32 + b_f32) -> bf16   (with_pos_embed, transformer.py:272-273)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        assert a.shape == b.shape
        out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
        axpby_cast(a, b, 1.0, 1.0, out_bf16=out)
        return out

    @staticmethod
    def backward(ctx, d):
        d32 = torch.empty(d.shape, dtype=torch.float32, device=d.device)
        check(lib().spe_cast_bf16_to_f32(ptr(d.contiguous()), ptr(d32), d.numel(), stream()))
        return d32, d32


class CastBf16Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        return to_bf16(a)

    @staticmethod
    def backward(ctx, d):
        d32 = torch.empty(d.shape, dtype=torch.float32, device=d.device)
        check(lib().spe_cast_bf16_to_f32(ptr(d.contiguous()), ptr(d32), d.numel(), stream()))
        return d32


class CastF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        out = torch.empty(a.shape, dtype=torch.float32, device=a.device)
        check(lib().spe_cast_bf16_to_f32(ptr(a.contiguous()), ptr(out), a.numel(), stream()))
        return out

    @staticmethod
    def backward(ctx, d):
        return to_bf16(d)


def add_cast(a, b):
    return AddCastFn.apply(a, b)


def cast_bf16(a):
    return CastBf16Fn.apply(a)


def cast_f32(a):
    return CastF32Fn.apply(a)


# ---- backward milestones -------------------------------------------------------------------------
# A model marks a tensor whose gradient becomes ready at a known point of the backward pass (e.g. the backbone tap: once its gradient
# exists, every detector-side parameter gradient is final); engine.TrainStep hangs the first bucket of the gradient all-reduce there.
_MILESTONES = {}


def set_grad_milestone(name, fn):
    if fn is None:
        _MILESTONES.pop(name, None)
    else:
        _MILESTONES[name] = fn


def grad_milestone(x, name):
    fn = _MILESTONES.get(name)
    if fn is not None and x.requires_grad:
        def _hook(g, _fn=fn):
            _fn()
            return g
        x.register_hook(_hook)
    return x


# ---- dropout (cait.py:36-38,294,449; transformer.py:268-270,333-337; attention.py:371) -----------------------------------------
# The BASELINE configurations run p = 0; the reference's training scripts do not (scripts/run_coco17.py:30-32: backbone_drop_rate
# 0.07, drop_path_rate 0.2, drop_attn_rate 0.05; main.py:73: decoder dropout 0.1).  With a non-zero rate in train() mode the modules
# take un-fused routes (GEMM -> dropout -> residual) built from the same kernels; the masks come from torch's Philox generator
# (elementwise glue, reproducible with torch.manual_seed, capturable in CUDA graphs).  Eval mode and p = 0 keep the fused epilogues.
def dropout(x, p, training=True):
    if not training or p <= 0.0:
        return x
    return torch.nn.functional.dropout(x, p, True)


def drop_path(x, p, training=True):
    """timm DropPath (stochastic depth per sample): x [B, ...] * bernoulli(1 - p) / (1 - p)."""
    if not training or p <= 0.0:
        return x
    keep = 1.0 - p
    m = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
    return x * (m / keep)


def _drop_mask(shape, p, device):
    """boolean keep-mask of an attention-probability tensor (True = kept), drawn from torch's CUDA generator"""
    return torch.rand(shape, dtype=torch.float32, device=device) >= p


def _apply_drop_(t16, keep, p):
    """in place on a bf16 [B,H,Lq,ld] tensor: t * keep / (1 - p)   (the scalar multiply runs in fp32 op-math)"""
    t16.mul_(1.0 / (1.0 - p))
    t16.masked_fill_(~keep, 0.0)
    return t16


# ---- attention ---------------------------------------------------------------------------------
def _qk_logits(q, k, H, alpha, S, ldS, q2=None, k2=None):
    """S[b,h] = alpha * (q_h k_h^T [+ q2_h k2_h^T])   q [B,Lq,H*d] (row stride free), S f32 [B,H,Lq,ldS]."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    d = E // H
    gemm(q, k, S, Lq, Lk, d, lda=q.stride(1), ldb=k.stride(1), ldc=ldS, batch=(B, H), a_sb=(q.stride(0), d), b_sb=(k.stride(0), d),
         c_sb=(H * Lq * ldS, Lq * ldS), alpha=alpha)
    if q2 is not None:
        d2 = q2.shape[2] // H
        gemm(q2, k2, S, Lq, Lk, d2, lda=q2.stride(1), ldb=k2.stride(1), ldc=ldS, batch=(B, H), a_sb=(q2.stride(0), d2), b_sb=(k2.stride(0), d2),
             c_sb=(H * Lq * ldS, Lq * ldS), alpha=alpha, residual=S, ldr=ldS, r_sb=(H * Lq * ldS, Lq * ldS))


def _pv(P, v, H, out, Lq, Lk, ldP):
    """out[b,:,h*dv:(h+1)*dv] = P[b,h] @ v_h   (v consumed MN-major straight from [B,Lk,H*dv])."""
    B = v.shape[0]
    dv = v.shape[2] // H
    gemm(P, v, out, Lq, dv, Lk, lda=ldP, a_sb=(H * Lq * ldP, Lq * ldP), b_major=MAJOR_MN, ldb=v.stride(1), b_sb=(v.stride(0), dv),
         ldc=out.stride(1), c_sb=(out.stride(0), dv), batch=(B, H))


def _attn_bwd_common(dO, P, v, H, Lq, Lk, ldP):
    """dP[b,h] = dO_h v_h^T (bf16 [B,H,Lq,ldP]);  dV = P^T dO  (bf16 [B,Lk,H*dv])."""
    B = v.shape[0]
    dv = v.shape[2] // H
    dP = torch.empty((B, H, Lq, ldP), dtype=torch.bfloat16, device=v.device)
    gemm(dO, v, dP, Lq, Lk, dv, lda=dO.stride(1), a_sb=(dO.stride(0), dv), ldb=v.stride(1), b_sb=(v.stride(0), dv), ldc=ldP,
         c_sb=(H * Lq * ldP, Lq * ldP), batch=(B, H))
    dV = torch.empty((B, Lk, H * dv), dtype=torch.bfloat16, device=v.device)
    gemm(P, dO, dV, Lk, dv, Lq, a_major=MAJOR_MN, lda=ldP, a_sb=(H * Lq * ldP, Lq * ldP), b_major=MAJOR_MN, ldb=dO.stride(1),
         b_sb=(dO.stride(0), dv), ldc=H * dv, c_sb=(Lk * H * dv, dv), batch=(B, H))
    return dP, dV


def _dq_dk(dS, q, k, H, alpha, Lq, Lk, ldP, dq_out=None, dk_out=None):
    """dQ = alpha dS K ; dK = alpha dS^T Q  (both bf16, laid out like q / k)."""
    B = q.shape[0]
    d = q.shape[2] // H
    dq = dq_out if dq_out is not None else torch.empty((B, Lq, H * d), dtype=torch.bfloat16, device=q.device)
    dk = dk_out if dk_out is not None else torch.empty((B, Lk, H * d), dtype=torch.bfloat16, device=q.device)
    gemm(dS, k, dq, Lq, d, Lk, lda=ldP, a_sb=(H * Lq * ldP, Lq * ldP), b_major=MAJOR_MN, ldb=k.stride(1), b_sb=(k.stride(0), d),
         ldc=dq.stride(1), c_sb=(dq.stride(0), d), batch=(B, H), alpha=alpha)
    gemm(dS, q, dk, Lk, d, Lq, a_major=MAJOR_MN, lda=ldP, a_sb=(H * Lq * ldP, Lq * ldP), b_major=MAJOR_MN, ldb=q.stride(1),
         b_sb=(q.stride(0), d), ldc=dk.stride(1), c_sb=(dk.stride(0), d), batch=(B, H), alpha=alpha)
    return dq, dk


_UNFUSED_ATTN = None


def _fused_attention_ok(q, k, v, q2, k2, H):
    """shapes / layouts the fused forward kernel takes (everything on the detector path); SPE_ATTN_UNFUSED=1 forces the
    GEMM -> softmax -> GEMM pipeline (A/B timing)."""
    global _UNFUSED_ATTN
    if _UNFUSED_ATTN is None:
        import os
        _UNFUSED_ATTN = os.environ.get("SPE_ATTN_UNFUSED", "0") == "1"
    if _UNFUSED_ATTN:
        return False
    d, dv = q.shape[2] // H, v.shape[2] // H
    ok = d % 16 == 0 and d <= 64 and dv % 16 == 0 and dv <= 64 and k.shape[1] <= 8192
    if q2 is not None:
        d2 = q2.shape[2] // H
        ok = ok and d2 % 16 == 0 and d2 <= 64
    for t in (q, k, v, q2, k2):
        if t is not None:
            ok = ok and t.stride(2) == 1 and t.stride(1) % 8 == 0 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0
    return ok


def fused_attention_fwd(q, k, v, q2, k2, mask_u8, H, scale, out, P=None, lse=None):
    B, Lq, E = q.shape
    Lk = k.shape[1]
    a = _lib.AttentionArgs(B, H, Lq, Lk, E // H, (q2.shape[2] // H if q2 is not None else 0), v.shape[2] // H,
                           q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1), v.stride(0),
                           ptr(q2), (q2.stride(1) if q2 is not None else 0), (q2.stride(0) if q2 is not None else 0),
                           ptr(k2), (k2.stride(1) if k2 is not None else 0), (k2.stride(0) if k2 is not None else 0),
                           ptr(mask_u8), scale, out.data_ptr(), out.stride(1), out.stride(0), ptr(P), (P.shape[3] if P is not None else 0), ptr(lse))
    check(lib().spe_attention_fwd(C.byref(a), stream()))
    return out


def _fused_bwd_ok(q, k, dO, H):
    if not _fused_attention_ok(q, k, dO, None, None, H):
        return False
    return True


def fused_attention_bwd_gemms(dS, P, q, k, dO, H, alpha, dq, dk, dv, delta=None):
    """dV = P^T dO, dK = alpha dS^T q, dQ = alpha dS k in one pass over dS / P (attn_fused.cu).  dv None: dQ, dK only.
    delta (f32 [B,H,Lq]) given: `dS` holds dP and the softmax backward P o (dP - delta) is applied in shared memory (needs P)."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    d = E // H
    ws = torch.empty(int(lib().spe_attention_bwd_gemms_workspace(B, H, Lq, d)), dtype=torch.float32, device=q.device)
    a = _lib.AttentionBwdArgs(B, H, Lq, Lk, d, (dO.shape[2] // H if dO is not None else 0),
                              dS.data_ptr(), ptr(P), dS.shape[3], q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                              ptr(dO), (dO.stride(1) if dO is not None else 0), (dO.stride(0) if dO is not None else 0), alpha,
                              dq.data_ptr(), dq.stride(1), dq.stride(0), dk.data_ptr(), dk.stride(1), dk.stride(0),
                              ptr(dv), (dv.stride(1) if dv is not None else 0), (dv.stride(0) if dv is not None else 0), ws.data_ptr(), ptr(delta))
    check(lib().spe_attention_bwd_gemms(C.byref(a), stream()))


class AttentionFn(torch.autograd.Function):
    """softmax(scale * (q k^T [+ q2 k2^T]) + key_padding_mask) v, heads packed along the feature dim.
    nn.MultiheadAttention core (transformer.py:280) and models/attention.py:345-378 (d_qk != d_v, the
    conditional cross-attention's [content | position] concat expressed as two accumulated QK^T GEMMs).
    Optionally returns the head-mean attention map (cait.py:658-667)."""

    @staticmethod
    def forward(ctx, q, k, v, q2, k2, mask_u8, H, scale, want_mean, drop_p=0.0):
        _need_cuda(q, k, v)
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        ld = rup(Lk, 8)
        ctx.drop_p = float(drop_p)
        if not want_mean and drop_p <= 0.0 and _fused_attention_ok(q, k, v, q2, k2, H):
            # fused forward (attn_fused.cu): logits stay in TMEM.  With the recomputing backward only the row statistics are kept
            # (nothing N^2 reaches HBM in either direction); SPE_ATTN_BWD=gemm keeps P (bf16) for the GEMM-based backward.
            out = torch.empty((B, Lq, v.shape[2]), dtype=torch.bfloat16, device=q.device)
            if _recompute_bwd():
                lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
                fused_attention_fwd(q, k, v, q2, k2, mask_u8, H, scale, out, lse=lse)
                ctx.save_for_backward(q, k, v, q2, k2, lse, out, mask_u8)
                ctx.recompute = True
            else:
                P = torch.empty((B, H, Lq, ld), dtype=torch.bfloat16, device=q.device)
                fused_attention_fwd(q, k, v, q2, k2, mask_u8, H, scale, out, P=P)
                ctx.save_for_backward(q, k, v, q2, k2, P, out)
                ctx.recompute = False
            ctx.H, ctx.scale, ctx.ld = H, scale, ld
            return out
        ctx.recompute = False
        S = torch.empty((B, H, Lq, ld), dtype=torch.float32, device=q.device)
        _qk_logits(q, k, H, scale, S, ld, q2, k2)
        P = torch.empty((B, H, Lq, ld), dtype=torch.bfloat16, device=q.device)
        want_probs = want_mean == "probs"              # per-head probabilities (bf16 [B,H,Lq,ld]) instead of their head mean
        want_mean = bool(want_mean) and not want_probs
        pmean = torch.empty((B, Lq, Lk), dtype=torch.float32, device=q.device) if want_mean else None
        check(lib().spe_softmax_fwd(ptr(S), ptr(P), ptr(mask_u8), B, H, Lq, Lk, ld, ld, ptr(pmean), stream()))
        del S
        out = torch.empty((B, Lq, v.shape[2]), dtype=torch.bfloat16, device=q.device)
        if drop_p > 0.0:
            # attention dropout (attention.py:371 / nn.MultiheadAttention): P V runs on the dropped probabilities; the softmax
            # backward needs the clean ones, so both P and the keep-mask are saved
            keep = _drop_mask(P.shape, drop_p, q.device)
            _pv(_apply_drop_(P.clone(), keep, drop_p), v, H, out, Lq, Lk, ld)
            ctx.save_for_backward(q, k, v, q2, k2, P, out, keep)
        else:
            _pv(P, v, H, out, Lq, Lk, ld)
            ctx.save_for_backward(q, k, v, q2, k2, P, out)
        ctx.H, ctx.scale, ctx.ld = H, scale, ld
        if want_probs:
            Pv = P.detach()
            ctx.mark_non_differentiable(Pv)
            return out, Pv
        if want_mean:
            ctx.mark_non_differentiable(pmean)
            return out, pmean
        return out

    @staticmethod
    def backward(ctx, dO, *unused):
        if ctx.recompute:
            return AttentionFn._backward_recompute(ctx, dO) + (None,)
        if ctx.drop_p > 0.0:
            q, k, v, q2, k2, P, out, keep = ctx.saved_tensors
            H, scale, ld = ctx.H, ctx.scale, ctx.ld
            B, Lq, _ = q.shape
            Lk = k.shape[1]
            dO = dO.contiguous()
            dP, dV = _attn_bwd_common(dO, _apply_drop_(P.clone(), keep, ctx.drop_p), v, H, Lq, Lk, ld)
            _apply_drop_(dP, keep, ctx.drop_p)                  # d(dropout): the same mask and scale
            check(lib().spe_softmax_bwd(ptr(P), ptr(dP), ptr(dP), B, H, Lq, Lk, ld, stream()))
            dq, dk = _dq_dk(dP, q, k, H, scale, Lq, Lk, ld)
            dq2 = dk2 = None
            if q2 is not None:
                dq2, dk2 = _dq_dk(dP, q2, k2, H, scale, Lq, Lk, ld)
            return dq, dk, dV, dq2, dk2, None, None, None, None, None
        q, k, v, q2, k2, P, out = ctx.saved_tensors
        H, scale, ld = ctx.H, ctx.scale, ctx.ld
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        dO = dO.contiguous()
        if _fused_bwd_ok(q, k, dO, H) and (q2 is None or _fused_bwd_ok(q2, k2, dO, H)):
            dv_h = v.shape[2] // H
            dP = torch.empty((B, H, Lq, ld), dtype=torch.bfloat16, device=v.device)
            gemm(dO, v, dP, Lq, Lk, dv_h, lda=dO.stride(1), a_sb=(dO.stride(0), dv_h), ldb=v.stride(1), b_sb=(v.stride(0), dv_h), ldc=ld,
                 c_sb=(H * Lq * ld, Lq * ld), batch=(B, H))
            # softmax backward folded into the fused kernel: dS = P o (dP - delta) is formed on the staged tiles, delta = rowsum(dO o O)
            delta = torch.empty((B, H, Lq), dtype=torch.float32, device=v.device)
            check(lib().spe_attention_delta(ptr(dO), ptr(out), B, H, Lq, dv_h, dO.stride(1), dO.stride(0), out.stride(1), out.stride(0), ptr(delta), stream()))
            dq, dk = torch.empty_like(q, memory_format=torch.contiguous_format), torch.empty_like(k, memory_format=torch.contiguous_format)
            dV = torch.empty((B, Lk, v.shape[2]), dtype=torch.bfloat16, device=v.device)
            fused_attention_bwd_gemms(dP, P, q, k, dO, H, scale, dq, dk, dV, delta=delta)       # one pass over dP and P: dQ, dK, dV
            dq2 = dk2 = None
            if q2 is not None:
                dq2, dk2 = torch.empty_like(q2, memory_format=torch.contiguous_format), torch.empty_like(k2, memory_format=torch.contiguous_format)
                fused_attention_bwd_gemms(dP, P, q2, k2, None, H, scale, dq2, dk2, None, delta=delta)
            return dq, dk, dV, dq2, dk2, None, None, None, None, None
        dP, dV = _attn_bwd_common(dO, P, v, H, Lq, Lk, ld)
        check(lib().spe_softmax_bwd(ptr(P), ptr(dP), ptr(dP), B, H, Lq, Lk, ld, stream()))
        dq, dk = _dq_dk(dP, q, k, H, scale, Lq, Lk, ld)
        dq2 = dk2 = None
        if q2 is not None:
            dq2, dk2 = _dq_dk(dP, q2, k2, H, scale, Lq, Lk, ld)
        return dq, dk, dV, dq2, dk2, None, None, None, None, None


def _attention_backward_recompute(ctx, dO):
    """fully fused backward (attn_bwd_kernel): dq, dk, dv (+ dq2, dk2) from q, k, v, dO and the saved row statistics."""
    q, k, v, q2, k2, lse, out, mask_u8 = ctx.saved_tensors
    H, scale = ctx.H, ctx.scale
    B, Lq, E = q.shape
    Lk = k.shape[1]
    d, dv_h = E // H, v.shape[2] // H
    d2 = q2.shape[2] // H if q2 is not None else 0
    dO = dO.contiguous()
    delta = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
    check(lib().spe_attention_delta(ptr(dO), ptr(out), B, H, Lq, dv_h, dO.stride(1), dO.stride(0), out.stride(1), out.stride(0), ptr(delta), stream()))
    cf = torch.contiguous_format
    dq, dk, dV = torch.empty_like(q, memory_format=cf), torch.empty_like(k, memory_format=cf), torch.empty_like(v, memory_format=cf)
    dq2 = torch.empty_like(q2, memory_format=cf) if q2 is not None else None
    dk2 = torch.empty_like(k2, memory_format=cf) if k2 is not None else None
    ws = torch.empty(B * Lq * H * max(d, d2), dtype=torch.float32, device=q.device)
    st = lambda t: (t.data_ptr(), t.stride(1), t.stride(0)) if t is not None else (0, 0, 0)
    a = _lib.AttentionBwd2Args(B, H, Lq, Lk, d, d2, dv_h, *st(q), *st(k), *st(v), *st(q2), *st(k2), *st(dO), ptr(mask_u8), scale, ptr(lse), ptr(delta),
                               *st(dq), *st(dk), *st(dV), *st(dq2), *st(dk2), ws.data_ptr())
    check(lib().spe_attention_bwd(C.byref(a), stream()))
    return dq, dk, dV, dq2, dk2, None, None, None, None


AttentionFn._backward_recompute = staticmethod(_attention_backward_recompute)

_RECOMPUTE = None


def _recompute_bwd():
    global _RECOMPUTE
    if _RECOMPUTE is None:
        import os
        _RECOMPUTE = os.environ.get("SPE_ATTN_BWD", "recompute") != "gemm"
    return _RECOMPUTE


def attention(q, k, v, H, scale, mask_u8=None, q2=None, k2=None, want_mean=False, drop_p=0.0):
    """want_mean: False | True (also return the head-mean map f32 [B,Lq,Lk]) | "probs" (also return P bf16 [B,H,Lq,ld]).
    drop_p > 0: dropout on the attention probabilities (training), un-fused route."""
    return AttentionFn.apply(q, k, v, q2, k2, mask_u8, H, scale, want_mean, float(drop_p))


def cam_std_reweight(P, q0, C, k0, N):
    """TSCAM_cait_two_branch.std_reweighting (cait.py:801-806) on per-head probabilities P bf16 [B,H,Lq,ld] -> f32 [B,C,N]."""
    _need_cuda(P)
    B, H, Lq, ld = P.shape
    out = torch.empty((B, C, N), dtype=torch.float32, device=P.device)
    check(lib().spe_cam_std_reweight(ptr(P), B, H, Lq, ld, q0, C, k0, N, ptr(out), stream()))
    return out


_TH_FUSED = None
_TH_S16 = None


def _talking_s16():
    """SPE_TH_S16 (default 1): keep the talking-heads logits S in fp16 in HBM (unfused pipeline)."""
    global _TH_S16
    if _TH_S16 is None:
        import os
        _TH_S16 = os.environ.get("SPE_TH_S16", "1") != "0"
    return _TH_S16


def _talking_fused_mode():
    """SPE_TH_FUSED selects the talking-heads implementation where the head geometry is supported by csrc/talking_fused.cu:
      1    fused kernels in both directions: no [B,H,N,N] tensor in HBM (saved activations of a block: qkv + lse2 instead of
           S f32 + A bf16: ~0.98 GB -> ~30 MB per block at cfg2), DRAM traffic ~4.4 GB -> ~0.2 GB per layer;
      0    the unfused GEMM -> talking-softmax -> GEMM pipeline;
      auto (default) fused forward when no gradient is required (inference: 6.7 vs 11.4 ms per cfg2 step), unfused for
           training -- measured on B200 the three recomputing backward kernels are tensor-pipe bound (64-row x 16-column
           tcgen05.mma re-read their A operand from shared memory per instruction) and a fused training step takes 58.7 ms
           against 51.4 ms."""
    global _TH_FUSED
    if _TH_FUSED is None:
        import os
        _TH_FUSED = os.environ.get("SPE_TH_FUSED", "auto")
    return _TH_FUSED


def _talking_fused_ok(qkv, H, needs_grad):
    mode = _talking_fused_mode()
    if mode == "0" or (mode != "1" and needs_grad):
        return False
    B, N, D3 = qkv.shape
    D = D3 // 3
    return (D % H == 0 and lib().spe_talking_fused_supported(H, D // H) == 1 and qkv.stride(2) == 1
            and qkv.stride(1) % 8 == 0 and qkv.stride(0) % 8 == 0 and qkv.data_ptr() % 16 == 0 and D % 8 == 0)


def _fp32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class TalkingHeadsFusedFn(torch.autograd.Function):
    """Attention_talking_head core (cait.py:377-389), fused: logits in TMEM, head mixes on mma.sync, nothing of size N^2 in HBM.
    Saves qkv, the output-independent log2 normalisers lse2 [B,H,N] and the four small parameters; the backward recomputes."""

    @staticmethod
    def forward(ctx, qkv, Wl, bl, Ww, bw, H):
        _need_cuda(qkv, Wl)
        B, N, D3 = qkv.shape
        D = D3 // 3
        dh = D // H
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        Wl32, bl32, Ww32, bw32 = _fp32c(Wl), _fp32c(bl), _fp32c(Ww), _fp32c(bw)
        out = torch.empty((B, N, D), dtype=torch.bfloat16, device=qkv.device)
        lse2 = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
        ws = torch.empty(int(lib().spe_talking_fused_fwd_workspace(B, H, N, dh)), dtype=torch.uint8, device=qkv.device)
        a = _lib.TalkingFusedArgs(B, H, N, dh, q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                  v.data_ptr(), v.stride(1), v.stride(0), Wl32.data_ptr(), bl32.data_ptr(), Ww32.data_ptr(), bw32.data_ptr(),
                                  dh ** -0.5, out.data_ptr(), out.stride(1), out.stride(0), lse2.data_ptr(), ws.data_ptr(), ws.numel())
        check(lib().spe_talking_fused_fwd(C.byref(a), stream()))
        ctx.save_for_backward(qkv, lse2, Wl, bl, Ww, bw)
        ctx.H = H
        return out

    @staticmethod
    def backward(ctx, dO):
        qkv, lse2, Wl, bl, Ww, bw = ctx.saved_tensors
        H = ctx.H
        B, N, D3 = qkv.shape
        D = D3 // 3
        dh = D // H
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        dO = dO.contiguous()
        if dO.dtype != torch.bfloat16:
            dO = dO.to(torch.bfloat16)
        dev = qkv.device

        # exact bias gradient: dbw[g] = sum_b colsum_i(dO[b,:,g]) . colsum_j(V[b,:,g]);  dbl = 0 (softmax is shift invariant)
        def _dbw():
            cs = torch.zeros((2, B, D), dtype=torch.float32, device=dev)
            check(lib().spe_colsum_bf16_batched(ptr(dO), B, N, D, dO.stride(1), dO.stride(0), ptr(cs[0]), stream()))
            check(lib().spe_colsum_bf16_batched(ptr(v), B, N, D, v.stride(1), v.stride(0), ptr(cs[1]), stream()))
            return (cs[0] * cs[1]).view(B, H, dh).sum((0, 2))

        ss = None
        if _overlap_wgrad():
            with _SideStream() as ss:
                dbw = _dbw()
        Wl32, bl32, Ww32, bw32 = _fp32c(Wl), _fp32c(bl), _fp32c(Ww), _fp32c(bw)
        dqkv = torch.empty((B, N, D3), dtype=torch.bfloat16, device=dev)
        dWl_b, dWl = _grad_out(Wl, Wl.shape, dev)
        dWw_b, dWw = _grad_out(Ww, Ww.shape, dev)
        ws = torch.empty(int(lib().spe_talking_fused_bwd_workspace(B, H, N, dh)), dtype=torch.uint8, device=dev)
        a = _lib.TalkingFusedBwdArgs(B, H, N, dh, q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                     v.data_ptr(), v.stride(1), v.stride(0), dO.data_ptr(), dO.stride(1), dO.stride(0),
                                     Wl32.data_ptr(), bl32.data_ptr(), Ww32.data_ptr(), bw32.data_ptr(), dh ** -0.5, lse2.data_ptr(),
                                     dqkv.data_ptr(), dqkv.stride(1), dWl_b.data_ptr(), dWw_b.data_ptr(), ws.data_ptr(), ws.numel())
        check(lib().spe_talking_fused_bwd(C.byref(a), stream()))
        if ss is not None:
            ss.join(dbw)
        else:
            dbw = _dbw()
        dbl = None if grad_sink(bl) is not None else torch.zeros_like(bl)
        return dqkv, dWl, dbl, dWw, dbw, None


class TalkingHeadsAttentionFn(torch.autograd.Function):
    """Attention_talking_head core (cait.py:377-389) on a packed qkv [B,N,3D] bf16:
    S = scale q k^T -> proj_l over heads -> softmax -> proj_w over heads -> @ v.
    Unfused pipeline (batched GEMMs around the row-staged talking-softmax kernels): head geometries the fused kernels do not take."""

    @staticmethod
    def forward(ctx, qkv, Wl, bl, Ww, bw, H, drop_p=0.0):
        _need_cuda(qkv, Wl)
        B, N, D3 = qkv.shape
        D = D3 // 3
        dh = D // H
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        ld = rup(N, 8)
        scale = dh ** -0.5
        ctx.drop_p = float(drop_p)
        # logits in fp16 where the row-staged kernels apply: the head mix rounds S to fp16 anyway, half the N^2 logit traffic / memory
        s16 = _talking_s16() and lib().spe_talking_s16_supported(H, N, ld, ld) == 1
        S = torch.empty((B, H, N, ld), dtype=torch.float16 if s16 else torch.float32, device=qkv.device)
        _qk_logits(q, k, H, scale, S, ld)
        A = torch.empty((B, H, N, ld), dtype=torch.bfloat16, device=qkv.device)
        stats = torch.empty((B * N, H), dtype=torch.float32, device=qkv.device)
        fwd = lib().spe_talking_softmax_fwd_s16 if s16 else lib().spe_talking_softmax_fwd
        check(fwd(ptr(S), ptr(A), ptr(Wl), ptr(bl), ptr(Ww), ptr(bw), ptr(stats), B, H, N, N, ld, ld, stream()))
        out = torch.empty((B, N, D), dtype=torch.bfloat16, device=qkv.device)
        keep = None
        if drop_p > 0.0:                   # attn_drop on the post-mix probabilities (cait.py:387): A is saved dropped, + the mask
            keep = _drop_mask(A.shape, drop_p, qkv.device)
            _apply_drop_(A, keep, drop_p)
        _pv(A, v, H, out, N, N, ld)
        ctx.save_for_backward(qkv, S, A, Wl, bl, Ww, bw, stats, keep)
        ctx.H, ctx.ld = H, ld
        return out

    @staticmethod
    def backward(ctx, dO):
        qkv, S, A, Wl, bl, Ww, bw, stats, keep = ctx.saved_tensors
        H, ld = ctx.H, ctx.ld
        B, N, D3 = qkv.shape
        D = D3 // 3
        dh = D // H
        scale = dh ** -0.5
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
        dO = dO.contiguous()
        dqkv = torch.empty_like(qkv)

        # exact bias gradients (the kernel's own sums of bf16 dA over B*N*N keys cancel catastrophically):
        #   dbw[o] = sum_{b,i,j} dA[b,o,i,j] = sum_b colsum_i(dO[b,:,o]) . colsum_j(V[b,:,o]);   dbl = 0 (softmax is shift invariant)
        def _dbw():
            cs = torch.zeros((2, B, D), dtype=torch.float32, device=qkv.device)
            check(lib().spe_colsum_bf16_batched(ptr(dO), B, N, D, dO.stride(1), dO.stride(0), ptr(cs[0]), stream()))
            check(lib().spe_colsum_bf16_batched(ptr(v), B, N, D, v.stride(1), v.stride(0), ptr(cs[1]), stream()))
            return (cs[0] * cs[1]).view(B, H, dh).sum((0, 2))

        ss = None
        if _overlap_wgrad():                                  # needs dO and v only: runs beside the N^2 chain on the second stream
            with _SideStream() as ss:
                dbw = _dbw()
        # dA = dO v^T ; dV = A^T dO (written straight into the v third of dqkv)
        dA = torch.empty((B, H, N, ld), dtype=torch.bfloat16, device=qkv.device)
        gemm(dO, v, dA, N, N, dh, lda=dO.stride(1), a_sb=(dO.stride(0), dh), ldb=v.stride(1), b_sb=(v.stride(0), dh), ldc=ld,
             c_sb=(H * N * ld, N * ld), batch=(B, H))
        dv = dqkv[:, :, 2 * D:]
        fused = _fused_bwd_ok(q, k, dO, H)
        if not fused:
            gemm(A, dO, dv, N, dh, N, a_major=MAJOR_MN, lda=ld, a_sb=(H * N * ld, N * ld), b_major=MAJOR_MN, ldb=dO.stride(1), b_sb=(dO.stride(0), dh),
                 ldc=dv.stride(1), c_sb=(dv.stride(0), dh), batch=(B, H))
        dbw_drop = None
        if keep is not None:               # d(dropout): dA * keep / (1 - p); dV above used the dropped A, as the forward did
            _apply_drop_(dA, keep, ctx.drop_p)
            dbw_drop = dA.sum((0, 2, 3), dtype=torch.float32)      # the bias reaches only the kept positions
        dWl_b, dWl = _grad_out(Wl, Wl.shape, qkv.device)
        dWw_b, dWw = _grad_out(Ww, Ww.shape, qkv.device)
        junk = torch.zeros((2, H), dtype=torch.float32, device=qkv.device)      # the kernel's own bias sums (not used, see below)
        nws = lib().spe_talking_softmax_bwd_workspace(B, H, N, N)
        ws = torch.empty(nws, dtype=torch.float32, device=qkv.device)
        bwd = lib().spe_talking_softmax_bwd_s16 if S.dtype == torch.float16 else lib().spe_talking_softmax_bwd
        check(bwd(ptr(S), ptr(dA), ptr(dA), ptr(Wl), ptr(bl), ptr(Ww), ptr(bw), ptr(stats), B, H, N, N, ld, ld, ptr(dWl_b), ptr(junk[0]),
                  ptr(dWw_b), ptr(junk[1]), ptr(ws), nws, stream()))
        if fused:       # dQ, dK, dV from one pass over dS (= dA, in place) and A
            fused_attention_bwd_gemms(dA, A, q, k, dO, H, scale, dqkv[:, :, :D], dqkv[:, :, D:2 * D], dv)
        else:
            _dq_dk(dA, q, k, H, scale, N, N, ld, dq_out=dqkv[:, :, :D], dk_out=dqkv[:, :, D:2 * D])
        if ss is not None:
            ss.join(dbw)
        else:
            dbw = _dbw()
        if dbw_drop is not None:
            dbw = dbw_drop
        dbl = None if grad_sink(bl) is not None else torch.zeros_like(bl)
        return dqkv, dWl, dbl, dWw, dbw, None, None


def talking_heads_attention(qkv, Wl, bl, Ww, bw, H, drop_p=0.0):
    """drop_p > 0 (training): attn_drop on the post-mix probabilities (cait.py:387) -- un-fused route."""
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (qkv, Wl, bl, Ww, bw))
    if drop_p <= 0.0 and _talking_fused_ok(qkv, H, needs_grad):
        return TalkingHeadsFusedFn.apply(qkv, Wl, bl, Ww, bw, H)
    return TalkingHeadsAttentionFn.apply(qkv, Wl, bl, Ww, bw, H, float(drop_p))


# ---- patch embedding / position encodings ---------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """Conv2d(k=s=p) as im2col + tcgen05 GEMM, with the (bicubic-resized) position embedding added as the
    fp32 residual of the epilogue (cait.py:527, :598-613, :623-624).  pos_tokens f32 [N, D]."""

    @staticmethod
    def forward(ctx, img, weight, bias, pos_tokens, p):
        _need_cuda(img, weight)
        B, Cc, Hh, Ww = img.shape
        h, w = Hh // p, Ww // p
        D = weight.shape[0]
        Kc = Cc * p * p
        cols = torch.empty((B * h * w, Kc), dtype=torch.bfloat16, device=img.device)
        check(lib().spe_im2col_patch(ptr(img.contiguous()), B, Hh, Ww, p, ptr(cols), stream()))
        w16 = shadow(weight).view(D, Kc)
        out = torch.empty((B, h * w, D), dtype=torch.float32, device=img.device)
        # residual = pos tokens broadcast over the batch: batch1 = B with residual batch stride 0
        gemm(cols, w16, out, h * w, D, Kc, lda=Kc, ldb=Kc, ldc=D, batch=(B, 1), a_sb=(h * w * Kc, 0), b_sb=(0, 0), c_sb=(h * w * D, 0),
             bias=bias, residual=pos_tokens, ldr=D, r_sb=(0, 0))
        ctx.save_for_backward(cols, weight)
        ctx.shape = (B, h * w, D, Kc)
        return out

    @staticmethod
    def backward(ctx, dout):
        cols, weight = ctx.saved_tensors
        B, N, D, Kc = ctx.shape
        dout = dout.contiguous()
        d16 = to_bf16(dout)
        db = torch.zeros(D, dtype=torch.float32, device=dout.device)
        colsum_bf16(d16.view(-1, D), db)
        dw = _wgrad_into(weight, d16.view(-1, D), cols)
        dpos = dout.sum(0) if B > 1 else dout[0]
        return None, dw, db, dpos, None


class BicubicTokensFn(torch.autograd.Function):
    """F.interpolate(mode='bicubic', align_corners=False) of the pos_embed grid (cait.py:588-600), token-major."""

    @staticmethod
    def forward(ctx, src, sh, sw, dh, dw):
        _need_cuda(src)
        D = src.shape[-1]
        src2 = src.reshape(sh * sw, D).contiguous()
        dst = torch.empty((dh * dw, D), dtype=torch.float32, device=src.device)
        check(lib().spe_bicubic_tokens_fwd(ptr(src2), sh, sw, D, ptr(dst), dh, dw, stream()))
        ctx.dims = (sh, sw, dh, dw, D, tuple(src.shape))
        return dst

    @staticmethod
    def backward(ctx, d):
        sh, sw, dh, dw, D, shp = ctx.dims
        dsrc = torch.zeros((sh * sw, D), dtype=torch.float32, device=d.device)
        check(lib().spe_bicubic_tokens_bwd(ptr(d.contiguous()), dh, dw, D, ptr(dsrc), sh, sw, stream()))
        return dsrc.view(shp), None, None, None, None


def sine_pos_2d(mask_u8, D):
    """mask u8 [B,h,w] -> (pos f32 [B,h*w,D], pos bf16) (position_encoding.py:37-57). No gradient."""
    _need_cuda(mask_u8)
    B, h, w = mask_u8.shape
    pos = torch.empty((B, h * w, D), dtype=torch.float32, device=mask_u8.device)
    pos16 = torch.empty((B, h * w, D), dtype=torch.bfloat16, device=mask_u8.device)
    check(lib().spe_sine_pos_2d(ptr(mask_u8.contiguous()), B, h, w, D, ptr(pos), ptr(pos16), stream()))
    return pos, pos16


class QuerySineFn(torch.autograd.Function):
    """gen_sineembed_for_position (transformer.py:35-49): ref f32 [...,2] -> f32 [...,D]."""

    @staticmethod
    def forward(ctx, ref, D):
        _need_cuda(ref)
        ref = ref.contiguous()
        n = ref.numel() // 2
        emb = torch.empty(ref.shape[:-1] + (D,), dtype=torch.float32, device=ref.device)
        check(lib().spe_query_sine_fwd(ptr(ref), n, D, ptr(emb), stream()))
        ctx.save_for_backward(ref)
        ctx.D = D
        return emb

    @staticmethod
    def backward(ctx, d):
        (ref,) = ctx.saved_tensors
        dref = torch.empty_like(ref)
        check(lib().spe_query_sine_bwd(ptr(ref), ptr(d.contiguous()), ref.numel() // 2, ctx.D, ptr(dref), stream()))
        return dref, None


def query_sine_embed(ref, D):
    return QuerySineFn.apply(ref, D)
