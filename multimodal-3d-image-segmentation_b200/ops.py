"""Thin PyTorch wrappers (tensor allocation + autograd wiring) over the C ABI.  All compute is in the .so.

Tensor convention: activations are (B, C, D, P) "planar" tensors (P = plane pitch >= H*W) or ordinary dense
(B, C, D, H, W) tensors, which are the same memory layout with P == H*W.
"""
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .plan import get_crop_plan, get_interp_tables

_ws_cache = {}
_ws_retired = []  # outgrown scratch buffers stay alive: captured CUDA graphs have their addresses baked in


def _require_cuda(t, name):
    if t.device.type != 'cuda':
        raise RuntimeError(f'hno_b200: {name} must be a CUDA tensor (got {t.device}); this package has no CPU path')
    if t.dtype != torch.float32:
        raise RuntimeError(f'hno_b200: {name} must be float32 (got {t.dtype})')
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream (the reference sets it once, experiments/run.py:38-40)
        raise RuntimeError(f'hno_b200: {name} lives on {t.device} but the current CUDA device is '
                           f'cuda:{torch.cuda.current_device()}; call torch.cuda.set_device({t.device.index}) first')


def workspace(nbytes, device, tag):
    """Grow-only scratch buffer per (device, stream, tag).  Kernels on one stream are serialised, so reuse within a
    stream is safe; different streams get different buffers.  A buffer that is outgrown is retired, not freed:
    parallel.Trainer replays CUDA graphs that captured its address (a later, larger request -- validation at another
    batch size, a super-resolution grid -- must not hand that memory back to the caching allocator)."""
    dev = torch.device(device)
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream if dev.type == 'cuda' else 0, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = _ws_cache[key] = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev)
    return buf


# ------------------------------------------------------------------------------------------ DHT
def _geom(x, hw=None):
    """(nslab, D, pitch, dense_hw) of a planar (B,C,D,P) or dense (B,C,D,H,W) tensor."""
    if x.ndim == 5:
        return x.shape[0] * x.shape[1], x.shape[2], x.shape[3] * x.shape[4]
    assert x.ndim == 4
    return x.shape[0] * x.shape[1], x.shape[2], x.shape[3]


def dht3_forward(x, plan, scale):
    """z[B,C,Ld,Lh,Lw] = scale * C x."""
    _require_cuda(x, 'x')
    x = x.contiguous()
    nslab, D, pitch = _geom(x)
    assert D == plan.spatial[0] and pitch >= plan.spatial[1] * plan.spatial[2]
    z = torch.empty((x.shape[0], x.shape[1]) + plan.modes_shape, dtype=torch.float32, device=x.device)
    ws = workspace(plan.workspace_bytes(pitch, nslab), x.device, 'dht')
    call('hno_dht3_forward', plan.host.data_ptr(), plan.dev.data_ptr(), ptr(x), pitch, D * pitch, ptr(z), ptr(ws),
         nslab, float(scale), stream_ptr())
    return z


def dht3_adjoint(z, plan, scale, epilogue=0, out=None, pitch=None):
    """x (op)= scale * C^T z.  epilogue 0 store / 1 accumulate into `out` / 2 SELU / 3 out = selu(out + .).  Returns a dense (B,C,D,H,W)
    tensor when pitch is None, else a planar (B,C,D,pitch) one."""
    _require_cuda(z, 'z')
    z = z.contiguous()
    D, H, W = plan.spatial
    B, C = z.shape[:2]
    if out is None:
        assert epilogue not in (1, 3)
        shape = (B, C, D, H, W) if pitch is None else (B, C, D, pitch)
        out = torch.empty(shape, dtype=torch.float32, device=z.device)
    nslab, D2, p = _geom(out)
    assert D2 == D and nslab == B * C and out.is_contiguous()
    ws = workspace(plan.workspace_bytes(p, nslab), z.device, 'dht')
    call('hno_dht3_adjoint', plan.host.data_ptr(), plan.dev.data_ptr(), ptr(z), ptr(out), p, D * p, ptr(ws), nslab,
         float(scale), int(epilogue), stream_ptr())
    return out


def dht3_chain_eligible(x, plan, C, L):
    """The fused 'DHT -> shared-weight mixes -> inverse DHT' call applies (planar 16-byte aligned activations, 8 / 24 channels,
    grids the tensor-core H stage handles)."""
    nslab, D, pitch = _geom(x)
    return bool(_lib.load().hno_dht3_chain_eligible(plan.host.data_ptr(), ptr(x), pitch, D * pitch, x.shape[0], int(C),
                                                    int(L)))


def dht3_chain_forward(x, plan, weights, scale_in, epilogue=2, save=True, out=None):
    """out = EPI(C^T chain(scale_in C x)) for one HNO-XS block (reference nets/hnosegxs.py:260-268).  Returns (out, z_all):
    z_all (L + 1, B, C, Ld, Lh, Lw) holds z_0 .. z_L for the backward (None when save is False)."""
    _require_cuda(x, 'x')
    assert x.is_contiguous() and x.ndim == 4
    B, C = x.shape[:2]
    nslab, D, pitch = _geom(x)
    ws_ = [w.contiguous() for w in weights]
    L = len(ws_)
    z_all = torch.empty((L + 1, B, C) + plan.modes_shape, dtype=torch.float32, device=x.device) if save else None
    if out is None:
        assert epilogue in (0, 2)
        out = torch.empty_like(x)
    ws = workspace(plan.workspace_bytes(pitch, nslab), x.device, 'dht')
    call('hno_dht3_chain_forward', plan.host.data_ptr(), plan.dev.data_ptr(), ptr(x), ptr(out), pitch, D * pitch,
         _ptr_array(ws_), ptr(z_all), ptr(ws), B, C, L, float(scale_in), int(epilogue), stream_ptr())
    return out, z_all


def dht3_chain_backward(dt, plan, z_all, weights, scale_out, out, epilogue=1, dweights=None, accumulate=False):
    """out (+)= scale_out C^T chain_bwd(C dt); returns [dW_1 .. dW_L] (written into `dweights` when given)."""
    B, C = dt.shape[:2]
    nslab, D, pitch = _geom(dt)
    ws_ = [w.contiguous() for w in weights]
    L = len(ws_)
    if dweights is None:
        dweights = [torch.empty_like(w) for w in ws_]
    ws = workspace(plan.workspace_bytes(pitch, nslab), dt.device, 'dht')
    part = workspace(_lib.load().hno_dht3_chain_partials_bytes(plan.host.data_ptr(), C, L, B), dt.device, 'mc')
    call('hno_dht3_chain_backward', plan.host.data_ptr(), plan.dev.data_ptr(), ptr(dt.contiguous()), ptr(out), pitch,
         D * pitch, _ptr_array(ws_), _ptr_array(dweights), ptr(z_all), ptr(ws), ptr(part), B, C, L, float(scale_out),
         int(epilogue), int(bool(accumulate)), stream_ptr())
    return dweights


class TruncatedDHT(torch.autograd.Function):
    """TransformCrop: z = (1/N) C x; backward dx = (1/N) C^T dz."""

    @staticmethod
    def forward(ctx, x, plan):
        ctx.plan = plan
        return dht3_forward(x, plan, 1.0 / plan.n_voxels)

    @staticmethod
    def backward(ctx, dz):
        return dht3_adjoint(dz, ctx.plan, 1.0 / ctx.plan.n_voxels), None


class TruncatedIDHT(torch.autograd.Function):
    """PadInverse: x = C^T z (unnormalised); backward dz = C dx."""

    @staticmethod
    def forward(ctx, z, plan):
        ctx.plan = plan
        return dht3_adjoint(z, plan, 1.0)

    @staticmethod
    def backward(ctx, dx):
        return dht3_forward(dx, ctx.plan, 1.0), None


class AddIDHTSelu(torch.autograd.Function):
    """y = selu(t + C^T z): the spectral branch of an FNO / HNO block added to its 1x1x1 conv branch and activated
    (reference nets/architectures.py:521-536), evaluated by the epilogue of the adjoint-DHT's last stage."""

    @staticmethod
    def forward(ctx, t, z, plan):
        y = t.detach().clone()
        dht3_adjoint(z.detach(), plan, 1.0, epilogue=3, out=y)
        ctx.plan = plan
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dt = selu_backward(dy.contiguous(), y)
        return dt, dht3_forward(dt, ctx.plan, 1.0), None


# ------------------------------------------------------------------------------------------ pointwise conv
def _flat_s(t):
    s = 1
    for v in t.shape[2:]:
        s *= v
    return s


def pwconv_forward(in1, in2, weight, bias, act, residual=False):
    _require_cuda(in1, 'in1')
    in1 = in1.contiguous()
    in2 = in2.contiguous() if in2 is not None else None
    B, ci1 = in1.shape[:2]
    ci2 = in2.shape[1] if in2 is not None else 0
    co = weight.shape[0]
    out = torch.empty((B, co) + tuple(in1.shape[2:]), dtype=torch.float32, device=in1.device)
    call('hno_pwconv_forward', ptr(in1), ptr(in2), ptr(weight.contiguous()), ptr(bias), ptr(out), B, ci1, ci2, co,
         _flat_s(in1), int(act), int(bool(residual)), stream_ptr())
    return out


def pwconv_backward(dy, y, in1, in2, weight, act, residual=False, hw=None, need_in1=True, need_in2=True,
                    in1_is_selu=False, din1=None, din2=None, dweight=None, dbias=None, has_bias=True,
                    accumulate_w=False):
    """Returns (din1, din2, dweight, dbias).  `hw`: (P, HW) when the tensors are planar with padding columns."""
    B, ci1 = in1.shape[:2]
    ci2 = in2.shape[1] if in2 is not None else 0
    co = weight.shape[0]
    S = _flat_s(in1)
    P, HW = hw if hw is not None else (S, S)
    dev = in1.device
    flags = 0
    if din1 is not None:
        flags |= 1
    elif need_in1:
        din1 = torch.empty_like(in1)
    if in2 is not None:
        if din2 is not None:
            flags |= 2
        elif need_in2:
            din2 = torch.empty_like(in2)
    if in1_is_selu:
        flags |= 4
    if accumulate_w:
        flags |= 8
    if dweight is None:
        dweight = torch.empty((co, ci1 + ci2), dtype=torch.float32, device=dev)
    if dbias is None and has_bias:
        dbias = torch.empty((co,), dtype=torch.float32, device=dev)
    ws = workspace(_lib.load().hno_pwconv_backward_workspace_bytes(ci1, ci2, co), dev, 'pw')
    call('hno_pwconv_backward', ptr(dy.contiguous()), ptr(y), ptr(in1), ptr(in2), ptr(weight.contiguous()),
         ptr(din1), ptr(din2), ptr(dweight), ptr(dbias), ptr(ws), B, ci1, ci2, co, S, P, HW, int(act),
         int(bool(residual)), flags, stream_ptr())
    return din1, din2, dweight, dbias


class PointwiseConv(torch.autograd.Function):
    """y = act(W [in1; in2] + b (+ in1)); weight (CO, CI) or (CO, CI, 1, 1, 1)."""

    @staticmethod
    def forward(ctx, in1, in2, weight, bias, act, residual):
        w2 = weight.reshape(weight.shape[0], -1)
        in1 = in1.contiguous()
        in2 = in2.contiguous() if in2 is not None else None
        y = pwconv_forward(in1, in2, w2, bias, act, residual)
        ctx.save_for_backward(y, in1, in2, w2, bias)
        ctx.cfg = (act, residual, weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, in1, in2, w2, bias = ctx.saved_tensors
        act, residual, wshape = ctx.cfg
        din1, din2, dw, db = pwconv_backward(dy, y, in1, in2, w2, act, residual, need_in1=ctx.needs_input_grad[0],
                                             need_in2=in2 is not None and ctx.needs_input_grad[1],
                                             has_bias=bias is not None)
        return din1, din2, dw.reshape(wshape), db, None, None


# ------------------------------------------------------------------------------------------ shared-weight mode chain
def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def modechain_forward(z0, weights):
    """z_l = selu(W_l z_{l-1} + z_{l-1}) for every layer in ONE launch.  Returns a (L, B, C, ...) tensor of z_1..z_L."""
    _require_cuda(z0, 'z0')
    z0 = z0.contiguous()
    ws = [w.contiguous() for w in weights]
    B, C = z0.shape[:2]
    M = _flat_s(z0)
    zs = torch.empty((len(ws),) + tuple(z0.shape), dtype=torch.float32, device=z0.device)
    call('hno_modechain_forward', ptr(z0), _ptr_array(ws), ptr(zs), B, C, M, len(ws), stream_ptr())
    return zs


def modechain_backward(dzL, z0, zs, weights, dweights=None, accumulate=False):
    """Returns (dz0, [dW_1..dW_L]); `dweights`: optional destination tensors (views of a flat gradient buffer)."""
    ws = [w.contiguous() for w in weights]
    L = len(ws)
    B, C = z0.shape[:2]
    M = _flat_s(z0)
    if dweights is None:
        dweights = [torch.empty_like(w) for w in ws]
    dz0 = torch.empty_like(z0)
    wsp = workspace(_lib.load().hno_modechain_backward_workspace_bytes(B, C, M, L), z0.device, 'mc')
    call('hno_modechain_backward', ptr(dzL.contiguous()), ptr(z0), ptr(zs), _ptr_array(ws), ptr(dz0),
         _ptr_array(dweights), ptr(wsp), B, C, M, L, int(bool(accumulate)), stream_ptr())
    return dz0, dweights


class ModeChain(torch.autograd.Function):
    """The n_XS shared-weight NeuralOperatorBlocks of one HNO-XS block (reference nets/hnosegxs.py:261-262)."""

    @staticmethod
    def forward(ctx, z0, *weights):
        z0 = z0.contiguous()
        zs = modechain_forward(z0, weights)
        ctx.save_for_backward(z0, zs, *weights)
        return zs[-1]

    @staticmethod
    def backward(ctx, dz):
        z0, zs, *weights = ctx.saved_tensors
        dz0, dws = modechain_backward(dz, z0, zs, weights)
        return (dz0,) + tuple(dws)


def modechain_supported(C, L):
    return bool(_lib.load().hno_modechain_supported(int(C))) and 1 <= L <= 8


# ------------------------------------------------------------------------------------------ individual-weight mixing
class HartleyConv(torch.autograd.Function):
    """out(k) = 1/2 [W(k)(X(k)+X(~k)) + W(~k)(X(k)-X(~k))]  (reference nets/hartley_operator.py:293-317)."""

    @staticmethod
    def forward(ctx, x, weight, residual_selu=False):
        x = x.contiguous()
        weight = weight.contiguous()
        out = hartley_conv_forward(x, weight, residual_selu)
        ctx.save_for_backward(x, weight, out if residual_selu else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, y = ctx.saved_tensors
        dx, dw = hartley_conv_backward(dout, y, x, weight, need_dx=ctx.needs_input_grad[0],
                                       need_dw=ctx.needs_input_grad[1])
        return dx, dw, None


def hartley_conv_forward(x, weight, residual_selu=False):
    _require_cuda(x, 'x')
    B, ci = x.shape[:2]
    co = weight.shape[0]
    n0, n1, n2 = x.shape[2:]
    assert tuple(weight.shape[2:]) == (n0, n1, n2), 'individual weights must match the cropped mode block'
    out = torch.empty((B, co, n0, n1, n2), dtype=torch.float32, device=x.device)
    call('hno_hartley_conv_forward', ptr(x), ptr(weight), ptr(out), B, ci, co, n0, n1, n2, int(bool(residual_selu)),
         stream_ptr())
    return out


def hartley_conv_backward(dout, y, x, weight, need_dx=True, need_dw=True, dw=None, accumulate=False):
    B, ci = x.shape[:2]
    co = weight.shape[0]
    n0, n1, n2 = x.shape[2:]
    dx = torch.empty_like(x) if need_dx else None
    if dw is None and need_dw:
        dw = torch.empty_like(weight)
    call('hno_hartley_conv_backward', ptr(dout.contiguous()), ptr(y), ptr(x), ptr(weight), ptr(dx), ptr(dw), B, ci,
         co, n0, n1, n2, int(accumulate), stream_ptr())
    return dx, dw


class HartleyConvFull(torch.autograd.Function):
    """selu(1/2 [W(j)(X(j) + X(r j)) + W(~j)(X(j) - X(r j))]) with the partner r(j) taken in the FULL spectrum: the mixing of
    HartleyOperator(use_transform=True, weights_type='individual') (reference nets/hartley_operator.py:196-241, 267).
    x_ext (B, CI, E0, E1, E2): retained modes plus frequency +m per axis; rtab: int32 device tensor of partner positions."""

    @staticmethod
    def forward(ctx, x_ext, weight, rtab, act):
        _require_cuda(x_ext, 'x_ext')
        x_ext, weight = x_ext.contiguous(), weight.contiguous()
        B, ci = x_ext.shape[:2]
        co = weight.shape[0]
        n = tuple(weight.shape[2:])
        e = tuple(x_ext.shape[2:])
        out = torch.empty((B, co) + n, dtype=torch.float32, device=x_ext.device)
        call('hno_hartley_conv_full_forward', ptr(x_ext), ptr(weight), ptr(rtab), ptr(out), B, ci, co, *n, *e, int(act),
             stream_ptr())
        ctx.save_for_backward(x_ext, weight, rtab, out)
        ctx.act = int(act)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_ext, weight, rtab, out = ctx.saved_tensors
        B, ci = x_ext.shape[:2]
        co = weight.shape[0]
        dx = torch.empty_like(x_ext) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        call('hno_hartley_conv_full_backward', ptr(dout.contiguous()), ptr(out) if ctx.act else None, ptr(x_ext), ptr(weight),
             ptr(rtab), ptr(dx), ptr(dw), B, ci, co, *weight.shape[2:], *x_ext.shape[2:], stream_ptr())
        return dx, dw, None, None


class ComplexModeMix(torch.autograd.Function):
    """(a + i b)(k) = sum_i (wr + i wi)(o, i, k) (re + i im)(i, k): per-mode complex weights of FourierOperator
    (reference nets/fourier_operator.py:165-187, weights_type='individual') on real / imaginary mode tensors."""

    @staticmethod
    def forward(ctx, re, im, wr, wi):
        _require_cuda(re, 're')
        re, im, wr, wi = re.contiguous(), im.contiguous(), wr.contiguous(), wi.contiguous()
        B, ci = re.shape[:2]
        co = wr.shape[0]
        M = _flat_s(re)
        if tuple(wr.shape) != (co, ci) + tuple(re.shape[2:]) or wi.shape != wr.shape or im.shape != re.shape:
            raise ValueError(f'individual Fourier weights {tuple(wr.shape)} do not match the retained modes '
                             f'{tuple(re.shape)}')
        a = torch.empty((B, co) + tuple(re.shape[2:]), dtype=torch.float32, device=re.device)
        b = torch.empty_like(a)
        call('hno_complex_modemix_forward', ptr(re), ptr(im), ptr(wr), ptr(wi), ptr(a), ptr(b), B, ci, co, M,
             stream_ptr())
        ctx.save_for_backward(re, im, wr, wi)
        return a, b

    @staticmethod
    def backward(ctx, da, db):
        re, im, wr, wi = ctx.saved_tensors
        B, ci = re.shape[:2]
        co = wr.shape[0]
        need_x = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dre = torch.empty_like(re) if need_x else None
        dim = torch.empty_like(im) if need_x else None
        dwr = torch.empty_like(wr) if need_w else None
        dwi = torch.empty_like(wi) if need_w else None
        call('hno_complex_modemix_backward', ptr(da.contiguous()), ptr(db.contiguous()), ptr(re), ptr(im), ptr(wr),
             ptr(wi), ptr(dre), ptr(dim), ptr(dwr), ptr(dwi), B, ci, co, _flat_s(re), 0, stream_ptr())
        return dre, dim, dwr, dwi


def fourier_mix_forward(z, wr, wi, lin_k, lin_n, ck):
    """hno_fourier_mix_forward: z [B, ci, MS] -> hp [B, co, MS] (see FourierMixShared)."""
    _require_cuda(z, 'z')
    z, wr, wi = z.contiguous(), wr.contiguous(), wi.contiguous()
    B, ci, MS = z.shape
    co = wr.shape[0]
    MK = lin_k.numel()
    individual = wr.ndim > 2  # per-mode weights [co, ci, *K grid] (config_fno.ini) instead of one [co, ci] matrix
    if (tuple(wr.shape[:2]) != (co, ci) or wi.shape != wr.shape or lin_n.numel() != MK or ck.numel() != MK
            or (individual and wr[0, 0].numel() != MK)):
        raise ValueError(f'fourier_mix: weights {tuple(wr.shape)} / index tables do not match z {tuple(z.shape)}')
    if lin_k.dtype != torch.int32 or lin_n.dtype != torch.int32 or ck.dtype != torch.float32:
        raise TypeError('fourier_mix: lin_k / lin_n must be int32 and ck float32')
    hp = torch.empty((B, co, MS), dtype=torch.float32, device=z.device)
    call('hno_fourier_mix_forward', ptr(z), ptr(wr), ptr(wi), ptr(lin_k), ptr(lin_n), ptr(ck), ptr(hp), B, ci, co, MK, MS,
         int(individual), stream_ptr())
    return hp


def fourier_mix_backward(dhp, z, wr, wi, lin_k, lin_n, ck, need_x=True, need_w=True):
    """hno_fourier_mix_backward -> (dz, dw_real, dw_imag); entries not asked for are None."""
    z, wr, wi = z.contiguous(), wr.contiguous(), wi.contiguous()
    B, ci, MS = z.shape
    co = wr.shape[0]
    MK = lin_k.numel()
    dz = torch.empty_like(z) if need_x else None
    dwr = torch.empty_like(wr) if need_w else None
    dwi = torch.empty_like(wi) if need_w else None
    individual = wr.ndim > 2
    ws = (workspace(_lib.load().hno_fourier_mix_workspace_bytes(ci, co, MK, B), z.device, 'fmix')
          if need_w and not individual else None)
    call('hno_fourier_mix_backward', ptr(dhp.contiguous()), ptr(z), ptr(wr), ptr(wi), ptr(lin_k), ptr(lin_n), ptr(ck),
         ptr(dz), ptr(dwr), ptr(dwi), ptr(ws), B, ci, co, MK, MS, int(individual), 0, stream_ptr())
    return dz, dwr, dwi


class FourierMixShared(torch.autograd.Function):
    """Mode-domain step of FourierOperator with shared complex weights (reference nets/fourier_operator.py:155, 165-209) on
    the Hartley coefficients z [B, ci, MS] of the symmetric mode set: Re / Im split over the (k, N - k) pairs, complex
    channel mix, c_k-weighted re-assembly -> hp [B, co, MS] (hno_fourier_mix_forward / _backward)."""

    @staticmethod
    def forward(ctx, z, wr, wi, lin_k, lin_n, ck):
        hp = fourier_mix_forward(z, wr, wi, lin_k, lin_n, ck)
        ctx.save_for_backward(z, wr, wi, lin_k, lin_n, ck)
        return hp

    @staticmethod
    def backward(ctx, dhp):
        z, wr, wi, lin_k, lin_n, ck = ctx.saved_tensors
        dz, dwr, dwi = fourier_mix_backward(dhp, z, wr, wi, lin_k, lin_n, ck, ctx.needs_input_grad[0],
                                            ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        return dz, dwr, dwi, None, None, None


# ------------------------------------------------------------------------------------------ stem
def stem_out_shape(spatial):
    return tuple(s // 2 + 1 for s in spatial)


def stem_forward(x, weight, bias, pitch=None):
    _require_cuda(x, 'x')
    x = x.contiguous()
    B, cin, Dx, Hx, Wx = x.shape
    f = weight.shape[0]
    D, H, W = stem_out_shape((Dx, Hx, Wx))
    shape = (B, f, D, H, W) if pitch is None else (B, f, D, pitch)
    P = H * W if pitch is None else pitch
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    call('hno_stem_forward', ptr(x), ptr(weight.contiguous()), ptr(bias), ptr(out), B, cin, f, Dx, Hx, Wx, P,
         stream_ptr())
    return out


def stem_backward(dpre, x, f, pitch=None, dweight=None, dbias=None, accumulate=False):
    B, cin, Dx, Hx, Wx = x.shape
    D, H, W = stem_out_shape((Dx, Hx, Wx))
    P = H * W if pitch is None else pitch
    if dweight is None:
        dweight = torch.empty((f, cin, 2, 2, 2), dtype=torch.float32, device=x.device)
    if dbias is None:
        dbias = torch.empty((f,), dtype=torch.float32, device=x.device)
    ws = workspace(_lib.load().hno_stem_backward_workspace_bytes(cin, f), x.device, 'pw')
    call('hno_stem_backward', ptr(dpre.contiguous()), ptr(x), ptr(dweight), ptr(dbias), ptr(ws), B, cin, f, Dx, Hx,
         Wx, P, int(accumulate), stream_ptr())
    return dweight, dbias


def stem_backward_input(dpre, weight, image, pitch=None):
    """dx (B, cin, Dx, Hx, Wx): gradient of the stem w.r.t. the image from the gradient of its pre-activation."""
    B, f = dpre.shape[:2]
    cin = weight.shape[1]
    Dx, Hx, Wx = image
    D, H, W = stem_out_shape(image)
    P = H * W if pitch is None else pitch
    dx = torch.empty((B, cin, Dx, Hx, Wx), dtype=torch.float32, device=dpre.device)
    call('hno_stem_backward_input', ptr(dpre.contiguous()), ptr(weight.contiguous()), ptr(dx), B, cin, f, Dx, Hx, Wx, P,
         stream_ptr())
    return dx


class StemConv(torch.autograd.Function):
    """selu(Conv3d(k=2, s=2, p=1)(x)); the image gradient (transposed convolution) is computed only when asked for."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = stem_forward(x, weight, bias)
        ctx.save_for_backward(x, y, weight)
        ctx.f = weight.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight = ctx.saved_tensors
        dpre = selu_backward(dy, y)
        dw, db = stem_backward(dpre, x, ctx.f)
        dx = stem_backward_input(dpre, weight, tuple(x.shape[2:])) if ctx.needs_input_grad[0] else None
        return dx, dw, db


def selu_backward(dy, y):
    """dy * selu'(.) from the SELU output (tiny torch expression; only used by the stand-alone stem module)."""
    scale, alpha = 1.0507009873554805, 1.6732632423543772
    return dy * torch.where(y > 0, torch.full_like(y, scale), y + scale * alpha)


def permute_spatial(t, perm):
    """t.permute(<leading dims>, spatial axes in order `perm`).contiguous() for the permutations that move one spatial axis to
    the end -- (1, 2, 0) and (0, 2, 1) -- and the inverse of the first, (2, 0, 1), each as ONE batched 2-D transpose
    (hno_transpose2d); t: (..., D, H, W) contiguous, uint8 / int16 / float32."""
    perm = tuple(perm)
    if t.device.type != 'cuda':
        raise RuntimeError('hno_b200: permute_spatial needs a CUDA tensor; this package has no CPU path')
    t = t.contiguous()
    lead = tuple(t.shape[:-3])
    D, H, W = t.shape[-3:]
    n = 1
    for v in lead:
        n *= v
    if perm == (1, 2, 0):
        out = torch.empty(lead + (H, W, D), dtype=t.dtype, device=t.device)
        R, C = D, H * W
    elif perm == (0, 2, 1):
        out = torch.empty(lead + (D, W, H), dtype=t.dtype, device=t.device)
        n, R, C = n * D, H, W
    elif perm == (2, 0, 1):
        out = torch.empty(lead + (W, D, H), dtype=t.dtype, device=t.device)
        R, C = D * H, W
    else:
        raise ValueError(f'permute_spatial handles (1, 2, 0), (0, 2, 1) and (2, 0, 1), got {perm}')
    if t.element_size() not in (1, 2, 4):
        raise TypeError(f'permute_spatial: unsupported element type {t.dtype}')
    call('hno_transpose2d', ptr(t), ptr(out), t.element_size(), n, R, C, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------ head / losses
def head_forward(logits_low, tables, pitch, activation=1):
    B, C = logits_low.shape[:2]
    probs = torch.empty((B, C) + tuple(tables.hi), dtype=torch.float32, device=logits_low.device)
    call('hno_head_forward', tables.host.data_ptr(), tables.dev.data_ptr(), ptr(logits_low), ptr(probs), B, C, pitch,
         int(activation), stream_ptr())
    return probs


def head_argmax(logits_low, tables, pitch):
    """uint8 label map (B, Dx, Hx, Wx): argmax over the classes of the up-sampled logits (reference: np.argmax of the
    probabilities on the host, experiments/train_test.py:402-408)."""
    B, C = logits_low.shape[:2]
    labels = torch.empty((B,) + tuple(tables.hi), dtype=torch.uint8, device=logits_low.device)
    call('hno_head_argmax', tables.host.data_ptr(), tables.dev.data_ptr(), ptr(logits_low), ptr(labels), B, C, pitch,
         stream_ptr())
    return labels


def head_backward(dprobs, probs, tables, pitch, activation=1):
    B, C = dprobs.shape[:2]
    D, H, W = tables.lo
    shape = (B, C, D, pitch) if pitch != H * W else (B, C, D, H, W)
    dll = torch.empty(shape, dtype=torch.float32, device=dprobs.device)
    ws = workspace(tables.head_backward_workspace_bytes(B, C), dprobs.device, 'head')
    call('hno_head_backward', tables.host.data_ptr(), tables.dev.data_ptr(), ptr(dprobs.contiguous()), ptr(probs),
         ptr(dll), ptr(ws), B, C, pitch, int(activation), stream_ptr())
    return dll


def head_direct_forward(logits, spatial, activation=1):
    """Full-resolution head of the use_resize=False models: planar logits (B, C, D, P) -> dense output (B, C, D, H, W) with
    the softmax over the classes (activation=1) or as they are (0).  No interpolation (reference nets/hnosegxs.py:174-180
    with use_resize False)."""
    _require_cuda(logits, 'logits')
    logits = logits.contiguous()
    B, C = logits.shape[:2]
    D, H, W = spatial
    P = _geom(logits)[2]
    probs = torch.empty((B, C, D, H, W), dtype=torch.float32, device=logits.device)
    call('hno_head_direct_forward', ptr(logits), ptr(probs), B, C, D, H, W, P, int(activation), stream_ptr())
    return probs


def head_direct_argmax(logits, spatial):
    """uint8 label map (B, D, H, W) of planar full-resolution logits (first maximum, like np.argmax)."""
    _require_cuda(logits, 'logits')
    logits = logits.contiguous()
    B, C = logits.shape[:2]
    D, H, W = spatial
    labels = torch.empty((B, D, H, W), dtype=torch.uint8, device=logits.device)
    call('hno_head_direct_argmax', ptr(logits), ptr(labels), B, C, D, H, W, _geom(logits)[2], stream_ptr())
    return labels


def head_direct_backward(dprobs, probs, pitch, activation=1):
    """Gradient of the planar logits (B, C, D, pitch); padding columns are zero."""
    B, C, D, H, W = dprobs.shape
    dll = torch.empty((B, C, D, pitch), dtype=torch.float32, device=dprobs.device)
    call('hno_head_direct_backward', ptr(dprobs.contiguous()), ptr(probs), ptr(dll), B, C, D, H, W, pitch,
         int(activation), stream_ptr())
    return dll


class HeadUpsample(torch.autograd.Function):
    """probs = softmax(trilinear(logits_low)) (activation=1) or just the interpolation (activation=0)."""

    @staticmethod
    def forward(ctx, logits_low, tables, activation):
        logits_low = logits_low.contiguous()
        pitch = _geom(logits_low)[2]
        probs = head_forward(logits_low, tables, pitch, activation)
        ctx.save_for_backward(probs)
        ctx.cfg = (tables, pitch, activation)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        (probs,) = ctx.saved_tensors
        tables, pitch, activation = ctx.cfg
        return head_backward(dprobs, probs, tables, pitch, activation), None, None


# moment-based losses (one pass, five sums per (sample, label)); ExpDiceLoss carries its exponent as `param`
LOSS_KINDS = {'DiceLoss': 0, 'PCCLoss': 1, 'ExpDiceLoss': 2}
LOSS_DEFAULT_PARAM = {'DiceLoss': 0.0, 'PCCLoss': 0.0, 'ExpDiceLoss': 0.3}


class ProbabilityLoss(torch.autograd.Function):
    """DiceLoss / PCCLoss / ExpDiceLoss on probabilities and one-hot float targets (reference nets/custom_losses.py)."""

    @staticmethod
    def forward(ctx, y_pred, y_true, kind, param=0.0):
        y_pred = y_pred.contiguous()
        y_true = y_true.contiguous().to(torch.float32)
        loss, coef = prob_loss_forward(y_pred, y_true, kind, param)
        ctx.save_for_backward(y_pred, y_true, coef)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise RuntimeError('hno_b200: losses do not provide a gradient w.r.t. y_true')
        y_pred, y_true, coef = ctx.saved_tensors
        return prob_loss_backward(y_pred, y_true, coef, g), None, None, None


def prob_loss_forward(y_pred, y_true, kind, param=0.0):
    """(loss[1], coef[B*C*3]) of DiceLoss / PCCLoss / ExpDiceLoss on contiguous fp32 probabilities and one-hot targets."""
    _require_cuda(y_pred, 'y_pred')
    B, C = y_pred.shape[:2]
    N = _flat_s(y_pred)
    loss = torch.empty((1,), dtype=torch.float32, device=y_pred.device)
    coef = torch.empty((B * C * 3,), dtype=torch.float32, device=y_pred.device)
    ws = workspace(_lib.load().hno_loss_workspace_bytes(B, C), y_pred.device, 'loss')
    call('hno_loss_forward', ptr(y_pred), ptr(y_true), ptr(loss), ptr(coef), ptr(ws), B, C, N, int(kind),
         float(param), stream_ptr())
    return loss, coef


def prob_loss_backward(y_pred, y_true, coef, g):
    B, C = y_pred.shape[:2]
    N = _flat_s(y_pred)
    dyp = torch.empty_like(y_pred)
    if g is not None:  # None: d loss / d loss = 1
        g = g.reshape(1).to(torch.float32).contiguous()
    call('hno_loss_backward', ptr(y_pred), ptr(y_true), ptr(coef), ptr(g), ptr(dyp), B, C, N, stream_ptr())
    return dyp


def _ce_check(y_pred, y_true, labels):
    _require_cuda(y_pred, 'y_pred')
    if (y_true is None) == (labels is None):
        raise ValueError('cross entropy: pass exactly one of y_true (one-hot floats) and labels (uint8)')
    if not y_pred.is_contiguous():
        raise ValueError('cross entropy: y_pred must be a dense NCDHW tensor')
    if y_true is not None:
        _require_cuda(y_true, 'y_true')
        if y_true.shape != y_pred.shape or not y_true.is_contiguous():
            raise ValueError('cross entropy: y_true must be a dense tensor of the shape of y_pred')
    else:
        if labels.dtype != torch.uint8 or not labels.is_contiguous() or labels.device != y_pred.device or \
                tuple(labels.shape) != (y_pred.shape[0],) + tuple(y_pred.shape[2:]):
            raise ValueError('cross entropy: labels must be dense uint8 class indices of shape (B, *spatial) on the '
                             'device of y_pred')


def ce_loss_forward(y_pred, y_true=None, labels=None):
    """Cross entropy on probabilities (torch.nn.CrossEntropyLoss() as experiments/run.py:105-110 + train_test.py:159-160
    use it) against one-hot float targets OR uint8 labels [B][N]: returns loss[1]."""
    _ce_check(y_pred, y_true, labels)
    B, C = y_pred.shape[:2]
    N = _flat_s(y_pred)
    loss = torch.empty((1,), dtype=torch.float32, device=y_pred.device)
    ws = workspace(_lib.load().hno_ce_loss_workspace_bytes(B), y_pred.device, 'ce')
    call('hno_ce_loss_forward', ptr(y_pred), ptr(y_true), ptr(labels), ptr(loss), ptr(ws), B, C, N, stream_ptr())
    return loss


def ce_loss_backward(y_pred, y_true=None, labels=None, grad_loss=None):
    _ce_check(y_pred, y_true, labels)
    B, C = y_pred.shape[:2]
    dyp = torch.empty_like(y_pred)
    call('hno_ce_loss_backward', ptr(y_pred), ptr(y_true), ptr(labels), ptr(grad_loss), ptr(dyp), B, C, _flat_s(y_pred),
         stream_ptr())
    return dyp


class CrossEntropyOnProbabilities(torch.autograd.Function):
    """loss_fn(y_pred, y_true) for loss_name = CrossEntropyLoss; y_true is one-hot floats (or any class probabilities)."""

    @staticmethod
    def forward(ctx, y_pred, y_true):
        _require_cuda(y_pred, 'y_pred')
        y_pred = y_pred.contiguous()
        y_true = y_true.contiguous().to(torch.float32)
        if y_true.shape != y_pred.shape:
            raise ValueError(f'CrossEntropyLoss: y_true {tuple(y_true.shape)} must match y_pred {tuple(y_pred.shape)} '
                             '(class-probability targets, as experiments/train_test.py:152 builds them)')
        loss = ce_loss_forward(y_pred, y_true=y_true)
        ctx.save_for_backward(y_pred, y_true)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise RuntimeError('hno_b200: losses do not provide a gradient w.r.t. y_true')
        y_pred, y_true = ctx.saved_tensors
        g = g.reshape(1).to(torch.float32).contiguous()
        return ce_loss_backward(y_pred, y_true=y_true, grad_loss=g), None


def head_loss_forward(logits_low, labels, tables, pitch, kind, param=0.0):
    """Fused head + loss on uint8 labels: returns (loss[1], coef)."""
    B, C = logits_low.shape[:2]
    dev = logits_low.device
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    coef = torch.empty((B * C * 3,), dtype=torch.float32, device=dev)
    ws = workspace(tables.head_backward_workspace_bytes(B, C), dev, 'head')
    call('hno_head_loss_forward', tables.host.data_ptr(), tables.dev.data_ptr(), ptr(logits_low), ptr(labels),
         ptr(loss), ptr(coef), ptr(ws), B, C, pitch, int(kind), float(param), stream_ptr())
    return loss, coef


def head_loss_backward(logits_low, labels, coef, grad_loss, tables, pitch):
    B, C = logits_low.shape[:2]
    dll = torch.empty_like(logits_low)
    ws = workspace(tables.head_backward_workspace_bytes(B, C), logits_low.device, 'head')
    call('hno_head_loss_backward', tables.host.data_ptr(), tables.dev.data_ptr(), ptr(logits_low), ptr(labels),
         ptr(coef), ptr(grad_loss), ptr(dll), ptr(ws), B, C, pitch, stream_ptr())
    return dll


# ------------------------------------------------------------------------------------------ deep-supervision conv
def _int_array(vals):
    import ctypes
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def dsconv_forward(sources, weight, bias, act=1, weight2=None):
    """out = act(bias + sum_i W_i in_i) over a LIST of (B, C_i, D, P) planar tensors (conv_ds over the virtual
    concatenation of all block outputs, reference nets/architectures.py:306-311, 339-343).  weight (CO, sum C_i).
    weight2 (CO, CO): the bias-free conv_out that follows; returns (out, weight2 * out) then, else (out, None)."""
    _require_cuda(sources[0], 'sources[0]')
    srcs = [t.contiguous() for t in sources]
    B = srcs[0].shape[0]
    S = _flat_s(srcs[0])
    co = weight.shape[0]
    out = torch.empty((B, co) + tuple(srcs[0].shape[2:]), dtype=torch.float32, device=srcs[0].device)
    out2 = torch.empty_like(out) if weight2 is not None else None
    call('hno_dsconv_forward', _ptr_array(srcs), _int_array([t.shape[1] for t in srcs]), len(srcs),
         ptr(weight.contiguous()), ptr(bias), ptr(out), ptr(weight2.contiguous()) if weight2 is not None else None,
         ptr(out2), B, co, S, int(act), stream_ptr())
    return out, out2


def dsconv_backward(dy, y, sources, weight, hw, act=1, need_din=True, has_bias=True, dweight=None, dbias=None,
                    weight2=None, dweight2=None):
    """Returns ([din_i], dweight, dbias, dweight2).  hw = (P, HW): plane pitch and valid columns of the planar layout;
    with weight2, dy is the gradient of the second output."""
    import ctypes
    srcs = [t.contiguous() for t in sources]
    B = srcs[0].shape[0]
    S = _flat_s(srcs[0])
    co = weight.shape[0]
    ctot = sum(t.shape[1] for t in srcs)
    dev = srcs[0].device
    dins = [torch.empty_like(t) for t in srcs] if need_din else None
    if dweight is None:
        dweight = torch.empty((co, ctot), dtype=torch.float32, device=dev)
    if dbias is None and has_bias:
        dbias = torch.empty((co,), dtype=torch.float32, device=dev)
    ws = workspace(_lib.load().hno_dsconv_backward_workspace_bytes(ctot, co, B, S), dev, 'ds')
    din_arr = _ptr_array(dins) if dins is not None else (ctypes.c_void_p * len(srcs))()
    P, HW = hw if hw is not None else (S, S)
    if weight2 is not None and dweight2 is None:
        dweight2 = torch.empty((co, co), dtype=torch.float32, device=dev)
    call('hno_dsconv_backward', _ptr_array(srcs), din_arr, _int_array([t.shape[1] for t in srcs]), len(srcs),
         ptr(weight.contiguous()), ptr(dy.contiguous()), ptr(y),
         ptr(weight2.contiguous()) if weight2 is not None else None, ptr(dweight), ptr(dbias), ptr(dweight2), ptr(ws), B,
         co, S, int(P), int(HW), int(act), stream_ptr())
    return dins, dweight, dbias, dweight2


# ------------------------------------------------------------------------------------------ Hartley multi-head attention
def _round_up(v, m):
    return (v + m - 1) // m * m


def mha_feature_pitch(features):
    """Features per head padded for the GEMM tiles: a multiple of 32, and of 128 beyond 256."""
    fp = _round_up(features, 32)
    return fp if fp <= 256 else _round_up(features, 128)


class MhaGeometry:
    """Token / feature geometry of HartleyMultiHeadAttention on a cropped mode block (reference nets/hartley_mha.py:
    179-196: grouping3d turns every patch of pd*ph*pw modes into one token with channels*pd*ph*pw features)."""

    def __init__(self, modes_shape, patch):
        self.L = tuple(int(v) for v in modes_shape)
        self.patch = tuple(int(v) for v in patch) if patch is not None else (1, 1, 1)
        if any(l % p for l, p in zip(self.L, self.patch)):
            raise AssertionError(f'retained modes {self.L} must be divisible by the patch size {self.patch}')
        self.P = self.patch[0] * self.patch[1] * self.patch[2]
        self.T = (self.L[0] // self.patch[0]) * (self.L[1] // self.patch[1]) * (self.L[2] // self.patch[2])
        self.Tp = _round_up(self.T, 128)

    def args(self):
        return self.L + self.patch + (self.Tp,)


def mha_project_forward(z, w, bias, geom):
    """z (B, cin, Ld, Lh, Lw), w (H, cd, cin) -> x_tok (B*H, Tp, Fp), x_chan (B*H, Fp, Tp)."""
    _require_cuda(z, 'z')
    z = z.contiguous()
    B, cin = z.shape[:2]
    H, cd = w.shape[:2]
    fp = mha_feature_pitch(cd * geom.P)
    x_tok = torch.empty((B * H, geom.Tp, fp), dtype=torch.float32, device=z.device)
    x_chan = torch.empty((B * H, fp, geom.Tp), dtype=torch.float32, device=z.device)
    call('hno_mha_project_forward', ptr(z), ptr(w.contiguous()), ptr(bias), ptr(x_tok), ptr(x_chan), B, H, cin, cd,
         *geom.args(), fp, stream_ptr())
    return x_tok, x_chan


def mha_project_backward(dx_tok, z, w, geom, dz=None, has_bias=False):
    """Returns (dz, dw, dbias); `dz` given -> accumulated into."""
    B, cin = z.shape[:2]
    H, cd = w.shape[:2]
    acc = dz is not None
    if dz is None:
        dz = torch.empty_like(z)
    dw = torch.empty_like(w)
    db = torch.empty((H, cd), dtype=torch.float32, device=z.device) if has_bias else None
    ws = workspace(_lib.load().hno_mha_wgrad_workspace_bytes(B, H, cin, cd, geom.T), z.device, 'mhaw')
    call('hno_mha_project_backward', ptr(dx_tok), ptr(z), ptr(w.contiguous()), ptr(dz), ptr(dw), ptr(db), ptr(ws), B, H, cin,
         cd, *geom.args(), dx_tok.shape[2], int(acc), stream_ptr())
    return dz, dw, db


class _MhaSaved:
    pass


def hartley_attention_forward(zq, zk, zv, wq, wk, wv, wo, bq=None, bk=None, bv=None, bo=None, patch=None, activation=1,
                              save=True):
    """The frequency-domain part of HartleyMultiHeadAttention (reference nets/hartley_mha.py:165-216 between the transforms):
    per-head Q / K / V projections + grouping, att = act(Q^T K / sqrt(F)), att V, ungrouping, output projection.
    zk / zv None: shared with the previous input (self-attention computes ONE transform).  Returns (y, saved)."""
    geom = MhaGeometry(zq.shape[2:], patch)
    zk_ = zq if zk is None else zk
    zv_ = zk_ if zv is None else zv
    B = zq.shape[0]
    H, kd = wq.shape[:2]
    vd = wv.shape[1]
    q_tok, q_chan = mha_project_forward(zq, wq, bq, geom)
    k_tok, k_chan = mha_project_forward(zk_, wk, bk, geom)
    v_tok, v_chan = mha_project_forward(zv_, wv, bv, geom)
    fq, fv = q_tok.shape[2], v_tok.shape[2]
    scale = 1.0 / float(kd * geom.P) ** 0.5  # att / sqrt(key.shape[2]) with the grouped channel count (:199)
    dev = zq.device
    P = torch.empty((B * H, geom.Tp, geom.Tp), dtype=torch.float32, device=dev)
    PT = torch.empty_like(P) if save else None
    o_tok = torch.empty((B * H, geom.Tp, fv), dtype=torch.float32, device=dev)
    call('hno_mha_attention_forward', ptr(q_tok), ptr(k_tok), ptr(v_chan), ptr(P), ptr(PT), ptr(o_tok), B * H, geom.Tp, fq,
         fv, scale, int(activation), stream_ptr())
    co = wo.shape[0]
    y = torch.empty((B, co) + geom.L, dtype=torch.float32, device=dev)
    call('hno_mha_output_forward', ptr(o_tok), ptr(wo.contiguous()), ptr(bo), ptr(y), B, H, co, vd, *geom.args(), fv,
         stream_ptr())
    if not save:
        return y, None
    S = _MhaSaved()
    S.geom, S.scale, S.activation = geom, scale, int(activation)
    S.z = (zq, zk, zv)
    S.w = (wq, wk, wv, wo)
    S.has_bias = (bq is not None, bk is not None, bv is not None, bo is not None)
    S.q_chan, S.k_chan, S.v_tok, S.P, S.PT, S.o_tok = q_chan, k_chan, v_tok, P, PT, o_tok
    return y, S


def hartley_attention_backward(dy, S):
    """Returns (dzq, dzk, dzv, dwq, dwk, dwv, dwo, dbq, dbk, dbv, dbo); dzk / dzv are None where the input was shared
    (their contribution is accumulated into the gradient of the tensor they share)."""
    geom = S.geom
    zq, zk, zv = S.z
    wq, wk, wv, wo = S.w
    B = zq.shape[0]
    H, kd = wq.shape[:2]
    vd = wv.shape[1]
    co = wo.shape[0]
    dev = zq.device
    fq, fv = S.q_chan.shape[1], S.v_tok.shape[2]
    dy = dy.contiguous()
    do_tok = torch.empty((B * H, geom.Tp, fv), dtype=torch.float32, device=dev)
    do_chan = torch.empty((B * H, fv, geom.Tp), dtype=torch.float32, device=dev)
    dwo = torch.empty_like(wo)
    dbo = torch.empty((co,), dtype=torch.float32, device=dev) if S.has_bias[3] else None
    ws = workspace(_lib.load().hno_mha_wgrad_workspace_bytes(B, H, co, vd, geom.T), dev, 'mhaw')
    call('hno_mha_output_backward', ptr(dy), ptr(S.o_tok), ptr(wo.contiguous()), ptr(do_tok), ptr(do_chan), ptr(dwo),
         ptr(dbo), ptr(ws), B, H, co, vd, *geom.args(), fv, stream_ptr())
    tt = B * H * geom.Tp * geom.Tp * 4
    scratch = workspace(2 * tt, dev, 'mha')
    dS = scratch[:tt].view(torch.float32)
    dST = scratch[tt:2 * tt].view(torch.float32)
    dq_tok = torch.empty((B * H, geom.Tp, fq), dtype=torch.float32, device=dev)
    dk_tok = torch.empty_like(dq_tok)
    dv_tok = torch.empty((B * H, geom.Tp, fv), dtype=torch.float32, device=dev)
    call('hno_mha_attention_backward', ptr(do_tok), ptr(do_chan), ptr(S.q_chan), ptr(S.k_chan), ptr(S.v_tok), ptr(S.P),
         ptr(S.PT), ptr(dS), ptr(dST), ptr(dq_tok), ptr(dk_tok), ptr(dv_tok), B * H, geom.Tp, fq, fv, S.scale,
         S.activation, stream_ptr())
    dzq, dwq, dbq = mha_project_backward(dq_tok, zq, wq, geom, None, S.has_bias[0])
    if zk is None:
        _, dwk, dbk = mha_project_backward(dk_tok, zq, wk, geom, dzq, S.has_bias[1])
        dzk, zk_ = None, zq
        dzk_acc = dzq
    else:
        dzk, dwk, dbk = mha_project_backward(dk_tok, zk, wk, geom, None, S.has_bias[1])
        zk_, dzk_acc = zk, dzk
    if zv is None:
        _, dwv, dbv = mha_project_backward(dv_tok, zk_, wv, geom, dzk_acc, S.has_bias[2])
        dzv = None
    else:
        dzv, dwv, dbv = mha_project_backward(dv_tok, zv, wv, geom, None, S.has_bias[2])
    return dzq, dzk, dzv, dwq, dwk, dwv, dwo, dbq, dbk, dbv, dbo


class HartleyAttention(torch.autograd.Function):
    """Autograd wrapper of hartley_attention_forward / _backward for the stand-alone module."""

    @staticmethod
    def forward(ctx, zq, zk, zv, wq, wk, wv, wo, bq, bk, bv, bo, patch, activation):
        y, S = hartley_attention_forward(zq.detach(), None if zk is None else zk.detach(),
                                         None if zv is None else zv.detach(), wq.detach(), wk.detach(), wv.detach(),
                                         wo.detach(), bq, bk, bv, bo, patch, activation, save=True)
        ctx.S = S
        return y

    @staticmethod
    def backward(ctx, dy):
        g = hartley_attention_backward(dy, ctx.S)
        ctx.S = None
        return g + (None, None)


__all__ = ['dht3_forward', 'dht3_adjoint', 'TruncatedDHT', 'TruncatedIDHT', 'AddIDHTSelu', 'pwconv_forward', 'pwconv_backward',
           'PointwiseConv', 'HartleyConv', 'ComplexModeMix', 'FourierMixShared', 'fourier_mix_forward', 'fourier_mix_backward', 'stem_forward', 'stem_backward', 'StemConv', 'head_forward',
           'head_backward', 'HeadUpsample', 'ProbabilityLoss', 'CrossEntropyOnProbabilities', 'ce_loss_forward',
           'ce_loss_backward', 'LOSS_DEFAULT_PARAM', 'head_loss_forward', 'head_loss_backward',
           'get_crop_plan', 'get_interp_tables', 'workspace', 'LOSS_KINDS', 'HartleyAttention', 'hartley_attention_forward',
           'hartley_attention_backward', 'dsconv_forward', 'dsconv_backward', 'head_argmax']
