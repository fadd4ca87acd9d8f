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


class XSEngine:
    def __init__(self, model):
        self.model = model

    # ------------------------------------------------------------------------------------------ parameters
    def named_slots(self):
        """[(parameter, kind)] in a fixed order; the backward returns gradients in the same order."""
        m = self.model
        slots = [m.conv_in.op.weight, m.conv_in.op.bias, m.conv1.op.weight, m.conv1.op.bias]
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
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
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
    def run_forward(self, x, save=True, head=True):
        m = self.model
        if x.ndim != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f'HNOSegXS expects (B, {m.in_channels}, D, H, W) input, got {tuple(x.shape)}')
        x = x.contiguous()
        dev = x.device
        B = x.shape[0]
        image = tuple(x.shape[2:])
        D, H, W = ops.stem_out_shape(image)
        pitch = plane_pitch(H, W)
        plan = get_crop_plan((D, H, W), m.num_modes, dev)
        inv_n = 1.0 / plan.n_voxels
        shared = m.weights_type == 'shared'
        nb = len(m.layers)
        S = _Saved()
        S.x, S.geom, S.plan = x, (D, H, W, pitch), plan

        a0 = ops.stem_forward(x, m.conv_in.op.weight, m.conv_in.op.bias, pitch)
        a1 = ops.pwconv_forward(a0, None, _w2(m.conv1.op), m.conv1.op.bias, 1, False)
        S.a0, S.a1 = a0, a1
        cur = a1
        stash = {}
        S.blocks = []
        for i, layer in enumerate(m.layers):
            rec = _Saved()
            rec.map_in = None
            if layer.mapping_conv is not None:
                enc = stash[nb - 1 - i]
                xin = ops.pwconv_forward(cur, enc, _w2(layer.mapping_conv.op), layer.mapping_conv.op.bias, 1, False)
                rec.map_in = (cur, enc)
            else:
                xin = cur
            zs = [ops.dht3_forward(xin, plan, inv_n)]
            rec.chain = None
            nmix = len(layer.conv_blocks)
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
        S.last = cur
        S.ll = ops.pwconv_forward(cur, None, _w2(m.conv_out), None, 0, False)
        S.tables = get_interp_tables((D, H, W), image, dev)
        S.act = 1 if m.output_activation == 'softmax' else 0
        probs = None
        if head:
            probs = ops.head_forward(S.ll, S.tables, pitch, S.act)
            S.probs = probs
        return probs, S

    # ------------------------------------------------------------------------------------------ backward
    def run_backward(self, S, dprobs=None, fused=None, dst=None):
        """Returns the gradients in named_slots() order.  Either `dprobs` (drop-in autograd) or
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
            dll = ops.head_backward(dprobs, S.probs, S.tables, pitch, S.act)
        dw_, _ = out_w(m.conv_out)
        dcur, _, g_out, _ = ops.pwconv_backward(dll, None, S.last, None, _w2(m.conv_out), 0, False, hw=hw,
                                                has_bias=False, dweight=dw_)
        block_grads = [None] * nb
        dstash = {}
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
            dz = ops.dht3_forward(dt, plan, 1.0)
            g_mix = []
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
                dprev, denc, g_wm, g_bm = ops.pwconv_backward(dxin, rec.xin, prev, enc, _w2(op), 1, False, hw=hw,
                                                              din1=tgt, dweight=dw_, dbias=db_)
                k = nb - 1 - i
                if k in dstash:
                    dstash[k] += denc
                else:
                    dstash[k] = denc
                g += [g_wm.reshape(op.weight.shape), g_bm]
                dcur = dprev
            else:
                dcur = dxin
            g += g_mix
            if layer.conv_concat is not None:
                g += [g_wc.reshape(layer.conv_concat.op.weight.shape), g_bc]
            block_grads[i] = g
        dw_, db_ = out_w(m.conv1.op)
        dpre0, _, g_w1, g_b1 = ops.pwconv_backward(dcur, S.a1, S.a0, None, _w2(m.conv1.op), 1, False, hw=hw,
                                                   in1_is_selu=True, dweight=dw_, dbias=db_)
        dw_, db_ = out_w(m.conv_in.op)
        g_win, g_bin = ops.stem_backward(dpre0, S.x, F, pitch, dweight=dw_, dbias=db_)
        if dst is not None:
            return dst
        grads = [g_win, g_bin, g_w1.reshape(m.conv1.op.weight.shape), g_b1]
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
        if ctx.needs_input_grad[1]:
            raise RuntimeError('hno_b200: HNOSegXS does not provide a gradient w.r.t. its input volume')
        grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs.contiguous())
        ctx.S = None
        return (None, None) + tuple(grads)


class _XSLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, labels, kind, param, *params):
        _, S = engine.run_forward(x, save=True, head=False)
        if S.act != 1:
            raise NotImplementedError('the fused loss needs output_activation="softmax"')
        lab = _labels_u8(labels, x)
        loss, coef = ops.head_loss_forward(S.ll, lab, S.tables, S.geom[3], kind, param)
        ctx.engine, ctx.S, ctx.lab, ctx.coef = engine, S, lab, coef
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(1).to(torch.float32).contiguous()
        grads = ctx.engine.run_backward(ctx.S, fused=(ctx.lab, ctx.coef, g))
        ctx.S = None
        return (None, None, None, None, None) + tuple(grads)


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
        grads = ctx.engine.run_backward(ctx.S, dprobs=dprobs)
        ctx.S = ctx.probs = None
        return (None, None, None) + tuple(grads)
