"""Whole-network execution of HNOSeg-XS on the CUDA kernels: one autograd node for the entire model.

Replaces the call stack of reference nets/hnosegxs.py:145-182 (HNOSegXS.forward -> 8 x HNOXSBlock.forward
-> TransformCrop / NeuralOperatorBlock / PadInverse) and, in the fused variant, the loss of
experiments/train_test.py:152-160.  Activations live in planar tensors (B, C, D, P) with the plane pitch P
rounded up to 128 bytes; the concat skips are virtual (two pointers); U-Net skip gradients are accumulated
in place by the producing kernels, so no stand-alone add / cat / zeros kernels run.
"""
import torch

from . import ops
from .plan import get_crop_plan, get_interp_tables, plane_pitch


def _w2(conv):
    return conv.weight.reshape(conv.weight.shape[0], -1)


class _Saved:
    pass


def preferred_axis_perm(model, spatial):
    """Spatial permutation a volume of shape `spatial` should run on, or None.  HNOSeg-XS with shared weights is equivariant
    under permutations of the volume's axes (XSEngine.run_forward), and the transform's L2-resident stages are cheapest with
    the shortest axis LAST: real BraTS tensors are (155, 240, 240) (SimpleITK's z, y, x order) and run 13 % slower than the
    same volume stored (240, 240, 155).  HNO_AXIS_PERM: '0' off, '1' (default) large volumes only -- below ~2 M voxels the
    transposing copies cost more than they save --, 'force' any size (tests)."""
    import os
    mode = os.environ.get('HNO_AXIS_PERM', '1')
    if mode == '0':
        return None
    if not getattr(model, 'use_resize', True) or getattr(model, 'weights_type', 'shared') != 'shared':
        return None
    d, h, w = spatial
    if (d * h * w < (1 << 21) and mode != 'force') or w <= min(d, h):
        return None
    k = 0 if d <= h else 1  # the shortest axis goes last, the other two keep their order
    return tuple(i for i in range(3) if i != k) + (k,)


def _entry_forward(m, x, S, w_in=None):
    """First activation of the network and the grid the blocks run on.  use_resize=True: the stride-2 stem (reference
    nets/hnosegxs.py:102-105, 150-151).  use_resize=False (:102-109 skipped): the image itself as a planar tensor -- a view
    when the channel count is a multiple of 4 and the plane needs no padding (4 x 240 x 240 x 155: both hold), else a
    zero-padded copy; conv1's weight gets zero columns for the padding channels (S.w1)."""
    image = tuple(x.shape[2:])
    B, C = x.shape[:2]
    w1 = _w2(m.conv1.op)
    if m.use_resize:
        D, H, W = ops.stem_out_shape(image)
        pitch = plane_pitch(H, W)
        a0 = ops.stem_forward(x, m.conv_in.op.weight if w_in is None else w_in, m.conv_in.op.bias, pitch)
    else:
        D, H, W = image
        pitch = plane_pitch(H, W)
        cp = (C + 3) // 4 * 4
        if cp == C and pitch == H * W:
            a0 = x.view(B, C, D, pitch)
        else:
            a0 = x.new_zeros((B, cp, D, pitch))
            a0[:, :C, :, :H * W] = x.view(B, C, D, H * W)
            if cp != C:
                w1 = torch.nn.functional.pad(w1, (0, cp - C))
    S.x, S.geom, S.a0, S.w1 = x, (D, H, W, pitch), a0, w1
    S.a1 = ops.pwconv_forward(a0, None, w1, m.conv1.op.bias, 1, False)
    return S.a1


def _entry_backward(m, S, dcur, hw, dst_w1=None, dst_b1=None, dst_win=None, dst_bin=None, need_dx=False):
    """Gradients of conv1 and (use_resize=True) the stem: [g_win, g_bin,] g_w1, g_b1 in named_slots() order.  `need_dx`: also
    the gradient w.r.t. the image, left in S.dx (training never asks for it: the image is data)."""
    op = m.conv1.op
    S.dx = None
    if m.use_resize:
        dpre0, _, g_w1, g_b1 = ops.pwconv_backward(dcur, S.a1, S.a0, None, S.w1, 1, False, hw=hw, in1_is_selu=True,
                                                   dweight=dst_w1, dbias=dst_b1)
        perm = getattr(S, 'perm', None)
        if perm is not None:  # the network ran on permuted axes: the stem's tap axes go back to the parameter's order
            assert not need_dx
            inv = [perm.index(i) for i in range(3)]
            g_win, g_bin = ops.stem_backward(dpre0, S.x, m.filters, S.geom[3], dbias=dst_bin)
            g_win = g_win.permute(0, 1, *[2 + i for i in inv])
            if dst_win is not None:
                dst_win.copy_(g_win.reshape(dst_win.shape))
                g_win = dst_win
            return [g_win, g_bin, g_w1.reshape(op.weight.shape), g_b1]
        g_win, g_bin = ops.stem_backward(dpre0, S.x, m.filters, S.geom[3], dweight=dst_win, dbias=dst_bin)
        if need_dx:
            S.dx = ops.stem_backward_input(dpre0, m.conv_in.op.weight, tuple(S.x.shape[2:]), S.geom[3])
        return [g_win, g_bin, g_w1.reshape(op.weight.shape), g_b1]
    cin = m.in_channels
    padded = S.w1.shape[1] != cin
    din, _, g_w1, g_b1 = ops.pwconv_backward(dcur, S.a1, S.a0, None, S.w1, 1, False, hw=hw, need_in1=need_dx,
                                             dweight=None if padded else dst_w1, dbias=dst_b1)
    if need_dx:  # planar (B, cp, D, pitch) -> the image's dense layout
        B, _, D, H, W = S.x.shape
        S.dx = din[:, :cin, :, :H * W].reshape(B, cin, D, H, W)
    if padded:
        g_w1 = g_w1[:, :cin]
        if dst_w1 is not None:
            dst_w1.copy_(g_w1)
            g_w1 = dst_w1
    return [g_w1.reshape(op.weight.shape), g_b1]


def _head_forward(m, S, image, act):
    """Probabilities (or logits, act 0) at the image resolution from the low-resolution logits S.ll."""
    D, H, W, pitch = S.geom
    if m.use_resize:
        return ops.head_forward(S.ll, S.tables, pitch, act)
    return ops.head_direct_forward(S.ll, (D, H, W), act)


def _head_backward(m, S, dprobs):
    if m.use_resize:
        return ops.head_backward(dprobs, S.probs, S.tables, S.geom[3], S.act)
    return ops.head_direct_backward(dprobs, S.probs, S.geom[3], S.act)


class XSEngine:
    def __init__(self, model):
        self.model = model

    # ------------------------------------------------------------------------------------------ parameters
    def named_slots(self):
        """[(parameter, kind)] in a fixed order; the backward returns gradients in the same order."""
        m = self.model
        slots = [m.conv_in.op.weight, m.conv_in.op.bias] if m.use_resize else []
        slots += [m.conv1.op.weight, m.conv1.op.bias]
        for layer in m.layers:
            if layer.mapping_conv is not None:
                slots += [layer.mapping_conv.op.weight, layer.mapping_conv.op.bias]
            slots += [blk.op.weight for blk in layer.conv_blocks]
            if layer.conv_concat is not None:
                slots += [layer.conv_concat.op.weight, layer.conv_concat.op.bias]
        slots.append(m.conv_out.weight)
        return slots

    # ------------------------------------------------------------------------------------------ public entry points
    def forward(self, x):
        params = self.named_slots()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _XSFunction.apply(self, x, *params)
        return self.run_forward(x, save=False)[0]

    def loss(self, x, labels, loss_name='DiceLoss', param=None):
        """Fused training objective on integer labels: the probabilities are never materialised.
        Numerically the same as loss_fn(model(x), to_categorical(labels)) of experiments/train_test.py:152-160.
        `param` is the loss's constructor argument where it has one (ExpDiceLoss: exp, default 0.3).
        CrossEntropyLoss (torch.nn fall-through of run.py:105-110) is not of the five-moment form: it runs the head kernel,
        then the one-pass cross-entropy kernels on the uint8 labels, then the head backward."""
        if loss_name == 'CrossEntropyLoss':
            return _XSCrossEntropyFunction.apply(self, x, labels, *self.named_slots())
        if param is None:
            param = ops.LOSS_DEFAULT_PARAM[loss_name]
        return _XSLossFunction.apply(self, x, labels, ops.LOSS_KINDS[loss_name], float(param), *self.named_slots())

    # ------------------------------------------------------------------------------------------ forward
    def run_forward(self, x, save=True, head=True, perm=None):
        """`perm` (fused-loss training path only, head=False): run the network on the volume with its spatial axes permuted,
        x' = x.permute(0, 1, 2 + perm[0], 2 + perm[1], 2 + perm[2]).  Every operator of HNOSeg-XS with shared weights is
        equivariant under such a permutation once the mode counts and the 2x2x2 stem taps are permuted with it (pointwise
        convolutions, separable transform and interpolation, voxel-wise loss sums), so loss and gradients are those of the
        un-permuted volume; the point is speed: the transform contracts D -> H -> W and its L2-resident stages are cheapest
        with the SHORTEST axis last (155 x 240 x 240 as 240 x 240 x 155: DESIGN 8.3)."""
        m = self.model
        if x.ndim != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f'HNOSegXS expects (B, {m.in_channels}, D, H, W) input, got {tuple(x.shape)}')
        modes, w_in = m.num_modes, None
        if perm is not None:
            assert not head and m.use_resize and m.weights_type == 'shared' and sorted(perm) == [0, 1, 2]
            dims = [2 + p for p in perm]
            x = ops.permute_spatial(x, perm)
            modes = tuple(m.num_modes[p] for p in perm)
            w_in = m.conv_in.op.weight.permute(0, 1, *dims).contiguous()
        x = x.contiguous()
        dev = x.device
        image = tuple(x.shape[2:])
        S = _Saved()
        S.perm = perm
        a1 = _entry_forward(m, x, S, w_in)
        D, H, W, pitch = S.geom
        plan = get_crop_plan((D, H, W), modes, dev)
        inv_n = 1.0 / plan.n_voxels
        shared = m.weights_type == 'shared'
        nb = len(m.layers)
        S.plan = plan
        cur = a1
        stash = {}
        S.blocks = []
        S.ds = [a1] if m.use_deep_supervision else None
        for i, layer in enumerate(m.layers):
            rec = _Saved()
            rec.map_in = None
            if layer.mapping_conv is not None:
                enc = stash[nb - 1 - i]
                xin = ops.pwconv_forward(cur, enc, _w2(layer.mapping_conv.op), layer.mapping_conv.op.bias, 1, False)
                rec.map_in = (cur, enc)
            else:
                xin = cur
            nmix = len(layer.conv_blocks)
            rec.chain = rec.zall = None
            if shared and nmix and ops.dht3_chain_eligible(xin, plan, xin.shape[1], nmix):
                # transform -> n_XS mixes -> inverse transform + SELU as five launches (the W stages, the recombination and
                # the mixes in ONE kernel, csrc/spectral_core.cu)
                u, rec.zall = ops.dht3_chain_forward(xin, plan, [blk.op.weight for blk in layer.conv_blocks], inv_n,
                                                     epilogue=2, save=save)
                zs = None
            else:
                zs = [ops.dht3_forward(xin, plan, inv_n)]
                if shared and nmix and ops.modechain_supported(zs[0].shape[1], nmix):
                    # all n_XS mixes of the block in one launch (reference nets/hnosegxs.py:261-262)
                    rec.chain = ops.modechain_forward(zs[0], [blk.op.weight for blk in layer.conv_blocks])
                    zs += [rec.chain[j] for j in range(nmix)]
                else:
                    for blk in layer.conv_blocks:
                        if shared:
                            zs.append(ops.pwconv_forward(zs[-1], None, blk.op.weight, None, 1, True))
                        else:
                            zs.append(ops.hartley_conv_forward(zs[-1], blk.op.weight, True))
                u = ops.dht3_adjoint(zs[-1], plan, 1.0, epilogue=2, pitch=pitch)  # selu(PadInverse(z))
            if layer.conv_concat is not None:
                y = ops.pwconv_forward(u, xin, _w2(layer.conv_concat.op), layer.conv_concat.op.bias, 1, False)
            else:
                y = u + xin
            rec.xin, rec.zs, rec.u, rec.y = xin, zs, u, y
            if save:
                S.blocks.append(rec)
            cur = y
            if m.use_unet_skip and i < nb // 2:
                stash[i] = y
            if m.use_deep_supervision:
                S.ds.append(y)
        S.last = cur
        if m.use_deep_supervision:
            # conv_out over the concatenation of conv1's and every block's output (reference nets/hnosegxs.py:154-172):
            # one pass over the list, the 216-channel concatenation is never formed
            S.ll = ops.dsconv_forward(S.ds, _w2(m.conv_out), None, act=0)[0]
        else:
            S.ll = ops.pwconv_forward(cur, None, _w2(m.conv_out), None, 0, False)
        S.tables = get_interp_tables((D, H, W), image, dev) if m.use_resize else None
        S.act = 1 if m.output_activation == 'softmax' else 0
        probs = None
        if head:
            probs = _head_forward(m, S, image, S.act)
            S.probs = probs
        return probs, S

    # ------------------------------------------------------------------------------------------ backward
    def run_backward(self, S, dprobs=None, fused=None, dst=None, need_dx=False):
        """Returns the gradients in named_slots() order (need_dx: the image gradient is left in S.dx).  Either `dprobs` (drop-in autograd) or
        fused=(labels_u8, coef, grad_loss) (fused head + loss) drives the head.  `dst`: optional list of
        destination tensors in named_slots() order (views of a flat gradient buffer); they are overwritten."""
        m = self.model
        slot = {id(p): i for i, p in enumerate(self.named_slots())}

        def out_w(conv_or_param):
            """(dweight 2-D view, dbias) destinations for a conv / parameter, or (None, None)."""
            if dst is None:
                return None, None
            if isinstance(conv_or_param, torch.nn.Parameter):
                w = dst[slot[id(conv_or_param)]]
                return (w.view(w.shape[0], -1) if w.ndim == 2 or w.ndim == 5 and w.shape[2:] == (1, 1, 1) else w), None
            w = dst[slot[id(conv_or_param.weight)]]
            b = dst[slot[id(conv_or_param.bias)]] if conv_or_param.bias is not None else None
            return w.view(w.shape[0], -1), b

        D, H, W, pitch = S.geom
        hw = (pitch, H * W)
        plan = S.plan
        inv_n = 1.0 / plan.n_voxels
        shared = m.weights_type == 'shared'
        nb = len(m.layers)
        F = m.filters
        if fused is not None:
            labels, coef, grad_loss = fused
            dll = ops.head_loss_backward(S.ll, labels, coef, grad_loss, S.tables, pitch)
        else:
            dll = _head_backward(m, S, dprobs)
        dw_, _ = out_w(m.conv_out)
        dstash = {}
        if m.use_deep_supervision:
            dins, g_out, _, _ = ops.dsconv_backward(dll, None, S.ds, _w2(m.conv_out), hw, act=0, has_bias=False,
                                                    dweight=dw_)
            dcur = dins[-1]
            for j in range(nb):  # gradient of source j (= input of block j) waits as an accumulation target
                dstash[j - 1] = dins[j]
        else:
            dcur, _, g_out, _ = ops.pwconv_backward(dll, None, S.last, None, _w2(m.conv_out), 0, False, hw=hw,
                                                    has_bias=False, dweight=dw_)
        block_grads = [None] * nb
        for i in reversed(range(nb)):
            layer = m.layers[i]
            rec = S.blocks[i]
            g = []
            has_map = rec.map_in is not None
            target = dstash.pop(i - 1) if (not has_map and (i - 1) in dstash) else None
            if layer.conv_concat is not None:
                op = layer.conv_concat.op
                dw_, db_ = out_w(op)
                dt, dxin, g_wc, g_bc = ops.pwconv_backward(dcur, rec.y, rec.u, rec.xin, _w2(op), 1, False, hw=hw,
                                                           in1_is_selu=True, din2=target, dweight=dw_, dbias=db_)
            else:
                dt = ops.selu_backward(dcur, rec.u)
                if target is not None:
                    target += dcur
                    dxin = target
                else:
                    dxin = dcur.clone()
            g_mix = []
            if rec.zall is not None:
                ws = [blk.op.weight for blk in layer.conv_blocks]
                dws = [out_w(w)[0] for w in ws] if dst is not None else None
                # dxin += (1/N) C^T chain_bwd(C dt): the same five launches as the forward
                g_mix = list(ops.dht3_chain_backward(dt, plan, rec.zall, ws, inv_n, dxin, epilogue=1, dweights=dws))
            else:
                dz = ops.dht3_forward(dt, plan, 1.0)
                if rec.chain is not None:
                    ws = [blk.op.weight for blk in layer.conv_blocks]
                    dws = [out_w(w)[0] for w in ws] if dst is not None else None
                    dz, g_mix = ops.modechain_backward(dz, rec.zs[0], rec.chain, ws, dweights=dws)
                    g_mix = list(g_mix)
                else:
                    for j in reversed(range(len(layer.conv_blocks))):
                        w = layer.conv_blocks[j].op.weight
                        dw_, _ = out_w(w)
                        if shared:
                            dz, _, gw, _ = ops.pwconv_backward(dz, rec.zs[j + 1], rec.zs[j], None, w, 1, True,
                                                               has_bias=False, dweight=dw_)
                        else:
                            dz, gw = ops.hartley_conv_backward(dz, rec.zs[j + 1], rec.zs[j], w, dw=dw_)
                        g_mix.append(gw)
                    g_mix.reverse()
                ops.dht3_adjoint(dz, plan, inv_n, epilogue=1, out=dxin)  # dxin += (1/N) C^T dz
            if has_map:
                prev, enc = rec.map_in
                op = layer.mapping_conv.op
                tgt = dstash.pop(i - 1) if (i - 1) in dstash else None
                dw_, db_ = out_w(op)
                k = nb - 1 - i
                dprev, denc, g_wm, g_bm = ops.pwconv_backward(dxin, rec.xin, prev, enc, _w2(op), 1, False, hw=hw,
                                                              din1=tgt, din2=dstash.get(k), dweight=dw_, dbias=db_)
                dstash[k] = denc  # accumulated in place when an entry (deep-supervision gradient) was waiting
                g += [g_wm.reshape(op.weight.shape), g_bm]
                dcur = dprev
            else:
                dcur = dxin
            g += g_mix
            if layer.conv_concat is not None:
                g += [g_wc.reshape(layer.conv_concat.op.weight.shape), g_bc]
            block_grads[i] = g
        assert not dstash, 'unconsumed skip / deep-supervision gradients'
        dw_, db_ = out_w(m.conv1.op)
        dwin_, dbin_ = out_w(m.conv_in.op) if m.use_resize else (None, None)
        grads = _entry_backward(m, S, dcur, hw, dw_, db_, dwin_, dbin_, need_dx=need_dx)
        if dst is not None:
            return dst
        for g in block_grads:
            grads += g
        grads.append(g_out.reshape(m.conv_out.weight.shape))
        return grads


def _labels_u8(labels, x):
    if labels.ndim == 5:
        if labels.shape[1] != 1:
            raise ValueError('labels must be (B, 1, D, H, W) or (B, D, H, W) integer class indices')
        labels = labels[:, 0]
    if tuple(labels.shape) != (x.shape[0],) + tuple(x.shape[2:]):
        raise ValueError(f'labels shape {tuple(labels.shape)} does not match the input volume {tuple(x.shape)}')
    return labels.to(device=x.device, dtype=torch.uint8).contiguous()


class _XSFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        probs, S = engine.run_forward(x, save=True)
        ctx.engine, ctx.S = engine, S
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        need_dx = ctx.needs_input_grad[1]
        grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs.contiguous(), need_dx=need_dx)
        dx = ctx.S.dx if need_dx else None
        ctx.S = None
        return (None, dx) + tuple(grads)


class _XSLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, labels, kind, param, *params):
        _, S = engine.run_forward(x, save=True, head=False)
        if S.act != 1:
            raise NotImplementedError('the fused loss needs output_activation="softmax"')
        lab = _labels_u8(labels, x)
        ctx.engine, ctx.S, ctx.lab = engine, S, lab
        if S.tables is None:
            # use_resize=False: no interpolation to fuse with; the probabilities are formed once at the image resolution and
            # the moment-based loss kernels run on them and the one-hot labels (hno_to_categorical + hno_loss_*)
            from .experiments.utils import to_categorical
            S.probs = _head_forward(engine.model, S, tuple(x.shape[2:]), 1)
            ctx.onehot = to_categorical(lab[:, None], S.probs.shape[1], validate=False)
            loss, ctx.coef = ops.prob_loss_forward(S.probs, ctx.onehot, kind, param)
            return loss[0]
        loss, coef = ops.head_loss_forward(S.ll, lab, S.tables, S.geom[3], kind, param)
        ctx.coef = coef
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(1).to(torch.float32).contiguous()
        need_dx = ctx.needs_input_grad[1]
        if ctx.S.tables is None:
            dprobs = ops.prob_loss_backward(ctx.S.probs, ctx.onehot, ctx.coef, g)
            grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs, need_dx=need_dx)
            ctx.onehot = None
        else:
            grads = ctx.engine.run_backward(ctx.S, fused=(ctx.lab, ctx.coef, g), need_dx=need_dx)
        dx = ctx.S.dx if need_dx else None
        ctx.S = None
        return (None, dx, None, None, None) + tuple(grads)


class _XSCrossEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, labels, *params):
        probs, S = engine.run_forward(x, save=True)
        lab = _labels_u8(labels, x)
        loss = ops.ce_loss_forward(probs, labels=lab)
        ctx.engine, ctx.S, ctx.lab, ctx.probs = engine, S, lab, probs
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(1).to(torch.float32).contiguous()
        dprobs = ops.ce_loss_backward(ctx.probs, labels=ctx.lab, grad_loss=g)
        need_dx = ctx.needs_input_grad[1]
        grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs, need_dx=need_dx)
        dx = ctx.S.dx if need_dx else None
        ctx.S = ctx.probs = None
        return (None, dx, None) + tuple(grads)


# =====================================================================================================================
# FNO-family engine: NeuralOperatorSeg(transform_type='Hartley', shared weights) = HNOSeg, and HartleyMHASeg
# =====================================================================================================================
class TransSegEngine:
    """Whole-network execution of the reference's `_TransSeg` models (nets/architectures.py:255-353) whose blocks are
    `_TransBlock`s (:511-548): spectral op with its own transform pair + 1x1x1 conv branch -> SELU -> concat-skip conv,
    optionally with deep supervision (conv_ds over all block outputs, :295-311, 330-343).  One autograd node; planar
    activations; per block

        t  = W_b x (+ b_b)                                     hno_pwconv_forward (no activation)
        z  = (1/N) C x                                         hno_dht3_forward
        z' = mix(z)                                            'hartley': selu(W z)   |   'mha': Hartley multi-head attention
        y  = selu(t + C^T z')                                  hno_dht3_adjoint, epilogue 3, in place in t
        out = selu(W_c [y; x] + b_c)                           hno_pwconv_forward over the virtual concat

    The backward runs the same kernels transposed; skip / deep-supervision gradients accumulate in place."""

    def __init__(self, model):
        self.model = model

    # ------------------------------------------------------------------------------------------ capability
    @staticmethod
    def block_kind(block):
        from .nets.hartley_mha import HartleyMultiHeadAttention
        from .nets.hartley_operator import HartleyOperator
        op = block.op
        if isinstance(op, HartleyOperator) and op.weights_type == 'shared' and op.use_transform and op.bias is None:
            return 'hartley'
        if isinstance(op, HartleyMultiHeadAttention) and op.use_transform:
            return 'mha'
        from .nets.fourier_operator import FourierOperator
        if (isinstance(op, FourierOperator) and op.bias is None
                and op.in_channels % 4 == 0 and op.out_channels % 4 == 0):
            return 'fourier'  # FNOSeg / FNO: hno_fourier_mix_* between the two transforms on the symmetric mode set
        return None

    @classmethod
    def supports(cls, model):
        for blk in model.layers:
            if cls.block_kind(blk) is None or blk.conv_branch is None or blk.conv_concat is None:
                return False
        return True

    # ------------------------------------------------------------------------------------------ parameters
    @staticmethod
    def _op_params(blk, kind):
        op = blk.op
        if kind == 'hartley':
            return [op.weight]
        if kind == 'fourier':
            return [op.weight_real, op.weight_imag]
        ps = [op.weight_query, op.weight_key, op.weight_value, op.weight_out]
        if op.use_bias:
            ps += [op.bias_query, op.bias_key, op.bias_value, op.bias_out]
        return ps

    def named_slots(self):
        m = self.model
        slots = [m.conv_in.op.weight, m.conv_in.op.bias] if m.use_resize else []
        slots += [m.conv1.op.weight, m.conv1.op.bias]
        for blk in m.layers:
            slots.append(blk.conv_branch.weight)
            if blk.conv_branch.bias is not None:
                slots.append(blk.conv_branch.bias)
            slots += self._op_params(blk, self.block_kind(blk))
            slots += [blk.conv_concat.op.weight, blk.conv_concat.op.bias]
        if m.conv_ds is not None:
            slots += [m.conv_ds.op.weight, m.conv_ds.op.bias]
        slots.append(m.conv_out.weight)
        return slots

    def forward(self, x):
        params = self.named_slots()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _TransSegFunction.apply(self, x, *params)
        return self.run_forward(x, save=False)[0]

    # ------------------------------------------------------------------------------------------ forward
    def run_forward(self, x, save=True, head=True):
        m = self.model
        if x.ndim != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f'{type(m).__name__} expects (B, {m.in_channels}, D, H, W) input, got {tuple(x.shape)}')
        x = x.contiguous()
        dev = x.device
        image = tuple(x.shape[2:])
        S = _Saved()
        a1 = _entry_forward(m, x, S)
        D, H, W, pitch = S.geom
        ds = [a1] if m.conv_ds is not None else None
        cur = a1
        S.blocks = []
        for blk in m.layers:
            kind = self.block_kind(blk)
            rec = _Saved()
            rec.kind = kind
            op = blk.op
            if kind == 'mha':
                assert all(s >= 2 * mm for s, mm in zip((D, H, W), op.num_modes))  # reference hartley_mha.py:165-172
            if kind == 'fourier':
                plan, _, _, _, ls, _, rec.tables = op._geometry((D, H, W), dev)
            else:
                plan = get_crop_plan((D, H, W), op.num_modes, dev)
            rec.plan = plan
            t = ops.pwconv_forward(cur, None, _w2(blk.conv_branch), blk.conv_branch.bias, 0, False)
            z = ops.dht3_forward(cur, plan, 1.0 / plan.n_voxels)
            if kind == 'hartley':
                zmix = ops.pwconv_forward(z, None, op.weight, None, 1, False)  # mix + SELU on the retained modes
                rec.z, rec.zmix = z, zmix
            elif kind == 'fourier':
                rec.z = z.reshape(z.shape[0], z.shape[1], -1)
                zmix = ops.fourier_mix_forward(rec.z, op.weight_real, op.weight_imag, *rec.tables)
                zmix = zmix.reshape((z.shape[0], op.out_channels) + tuple(ls))
            else:
                def flat(b):
                    return None if b is None else b.reshape(b.shape[1], b.shape[2]).contiguous()
                bo = None if op.bias_out is None else op.bias_out.reshape(-1).contiguous()
                zmix, rec.att = ops.hartley_attention_forward(z, None, None, op.weight_query, op.weight_key,
                                                              op.weight_value, op.weight_out, flat(op.bias_query),
                                                              flat(op.bias_key), flat(op.bias_value), bo, op.patch_size,
                                                              op._act, save=save)
            y = ops.dht3_adjoint(zmix, plan, 1.0, epilogue=3, out=t)  # y = selu(t + C^T z'), in place
            out = ops.pwconv_forward(y, cur, _w2(blk.conv_concat.op), blk.conv_concat.op.bias, 1, False)
            rec.x, rec.y, rec.out = cur, y, out
            if save:
                S.blocks.append(rec)
            cur = out
            if ds is not None:
                ds.append(out)
        S.last, S.ds = cur, ds
        if ds is not None:
            S.d, S.ll = ops.dsconv_forward(ds, _w2(m.conv_ds.op), m.conv_ds.op.bias, act=1, weight2=_w2(m.conv_out))
        else:
            S.ll = ops.pwconv_forward(cur, None, _w2(m.conv_out), None, 0, False)
        S.tables = get_interp_tables((D, H, W), image, dev) if m.use_resize else None
        S.act = 1 if m.output_activation_name == 'softmax' else 0
        probs = None
        if head:
            probs = _head_forward(m, S, image, S.act)
            S.probs = probs
        return probs, S

    # ------------------------------------------------------------------------------------------ backward
    def run_backward(self, S, dprobs=None, fused=None, need_dx=False):
        """Gradients in named_slots() order.  `dprobs` (drop-in autograd) or fused=(labels_u8, coef, grad_loss); need_dx: the
        image gradient is left in S.dx."""
        m = self.model
        D, H, W, pitch = S.geom
        hw = (pitch, H * W)
        F = m.filters
        if fused is not None:
            labels, coef, grad_loss = fused
            dll = ops.head_loss_backward(S.ll, labels, coef, grad_loss, S.tables, pitch)
        else:
            dll = _head_backward(m, S, dprobs)
        nb = len(m.layers)
        tail = []
        if S.ds is not None:
            dins, g_wds, g_bds, g_out = ops.dsconv_backward(dll, S.d, S.ds, _w2(m.conv_ds.op), hw, act=1,
                                                            weight2=_w2(m.conv_out))
            dcur = dins[-1]
            tail = [g_wds.reshape(m.conv_ds.op.weight.shape), g_bds]
        else:
            dins = None
            dcur, _, g_out, _ = ops.pwconv_backward(dll, None, S.last, None, _w2(m.conv_out), 0, False, hw=hw,
                                                    has_bias=False)
        block_grads = [None] * nb
        for i in reversed(range(nb)):
            blk, rec = m.layers[i], S.blocks[i]
            opc = blk.conv_concat.op
            target = dins[i] if dins is not None else None  # deep-supervision gradient of this block's input
            dpre, dx, g_wc, g_bc = ops.pwconv_backward(dcur, rec.out, rec.y, rec.x, _w2(opc), 1, False, hw=hw,
                                                       in1_is_selu=True, din2=target)
            dzmix = ops.dht3_forward(dpre, rec.plan, 1.0)
            if rec.kind == 'hartley':
                dz, _, g_w, _ = ops.pwconv_backward(dzmix, rec.zmix, rec.z, None, blk.op.weight, 1, False, has_bias=False)
                g_op = [g_w]
            elif rec.kind == 'fourier':
                op = blk.op
                dz, g_wr, g_wi = ops.fourier_mix_backward(dzmix.reshape(dzmix.shape[0], dzmix.shape[1], -1), rec.z,
                                                          op.weight_real, op.weight_imag, *rec.tables)
                dz = dz.reshape((dz.shape[0], dz.shape[1]) + tuple(dzmix.shape[2:]))
                g_op = [g_wr, g_wi]
            else:
                g = ops.hartley_attention_backward(dzmix, rec.att)
                dz = g[0]
                g_op = [g[3], g[4], g[5], g[6]]
                if blk.op.use_bias:
                    op = blk.op
                    g_op += [g[7].reshape(op.bias_query.shape), g[8].reshape(op.bias_key.shape),
                             g[9].reshape(op.bias_value.shape), g[10].reshape(op.bias_out.shape)]
            ops.dht3_adjoint(dz, rec.plan, 1.0 / rec.plan.n_voxels, epilogue=1, out=dx)  # dx += (1/N) C^T dz
            cb = blk.conv_branch
            _, _, g_wb, g_bb = ops.pwconv_backward(dpre, None, rec.x, None, _w2(cb), 0, False, hw=hw, din1=dx,
                                                   has_bias=cb.bias is not None)
            g = [g_wb.reshape(cb.weight.shape)]
            if cb.bias is not None:
                g.append(g_bb)
            g += g_op + [g_wc.reshape(opc.weight.shape), g_bc]
            block_grads[i] = g
            dcur = dx
        grads = _entry_backward(m, S, dcur, hw, need_dx=need_dx)
        for g in block_grads:
            grads += g
        grads += tail
        grads.append(g_out.reshape(m.conv_out.weight.shape))
        return grads


class _TransSegFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        probs, S = engine.run_forward(x, save=True)
        ctx.engine, ctx.S = engine, S
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        need_dx = ctx.needs_input_grad[1]
        grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs.contiguous(), need_dx=need_dx)
        dx = ctx.S.dx if need_dx else None
        ctx.S = None
        return (None, dx) + tuple(grads)
