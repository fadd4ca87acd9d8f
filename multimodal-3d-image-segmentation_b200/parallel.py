"""Data-parallel training plumbing: flat parameter / gradient buffers, ONE all-reduce per step, fused Adamax.

The reference is single-GPU (experiments/run.py:39).  HNOSeg-XS has 28,248 parameters (113 KB), every loss term
is a mean over (sample, label) pairs (nets/custom_losses.py:70,111), so batch-sharded data parallelism needs
exactly one SUM all-reduce of the flat gradient per step followed by a division by the world size; the
optimizer then runs replicated.  One process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr, stream_ptr


class FlatParameters:
    """Re-homes all parameters of a module into one contiguous buffer (and a matching gradient buffer)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters()]
        if not self.params:
            raise ValueError('module has no parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.data = torch.empty(n, dtype=dt, device=dev)
        self.grad = torch.zeros(n, dtype=dt, device=dev)
        self.views, self.grad_views = [], []
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.data[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.data[off:off + k].view(p.shape)
                gv = self.grad[off:off + k].view(p.shape)
                p.grad = gv
                self.views.append(p.data)
                self.grad_views.append(gv)
                off += k
        self.numel = n

    def grad_view_of(self, param):
        for p, g in zip(self.params, self.grad_views):
            if p is param:
                return g
        raise KeyError('parameter is not part of this flat buffer')

    def repack_grads_(self):
        """Makes `self.grad` hold every parameter's current gradient again and re-points p.grad at its view.

        `optimizer.zero_grad()` (set_to_none=True is the default, and it is what the reference loop calls at
        experiments/train_test.py:164) drops p.grad, so the next backward allocates fresh gradient tensors
        OUTSIDE the flat buffer.  A parameter whose .grad is None contributes zeros."""
        with torch.no_grad():
            for p, gv in zip(self.params, self.grad_views):
                g = p.grad
                if g is None:
                    gv.zero_()
                elif g.data_ptr() != gv.data_ptr() or g.shape != gv.shape:
                    gv.copy_(g.reshape(gv.shape))
                p.grad = gv
        return self.grad


def allreduce_mean_(flat_grad, group=None):
    """SUM all-reduce + 1/world scaling of the flat gradient: the only collective of a training step."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.mul_(1.0 / dist.get_world_size(group))
    return flat_grad


def attach_gradient_allreduce(optimizer, flat, group=None):
    """Drop-in hook for ANY torch.optim optimizer used by experiments/train_test.py:164-171: averages the flat
    gradient across ranks right before optimizer.step(), leaving the reference's training loop untouched."""
    def hook(opt, args, kwargs):
        flat.repack_grads_()  # survive optimizer.zero_grad(set_to_none=True): gradients may live outside the buffer
        allreduce_mean_(flat.grad, group)
    return optimizer.register_step_pre_hook(hook)


class FusedAdamax:
    """torch.optim.Adamax semantics (the reference's optimizer, config_hnoseg_xs.ini:53-55) as ONE kernel over
    the flat parameter vector; supports a per-step learning rate (cosine warm restarts are computed on the host)."""

    def __init__(self, flat, lr=5e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_inf = torch.zeros_like(flat.data)
        self.step_count = 0

    def step(self, lr=None):
        self.step_count += 1
        call('hno_adamax_step', ptr(self.flat.data), ptr(self.flat.grad), ptr(self.exp_avg), ptr(self.exp_inf),
             self.flat.numel, float(self.lr if lr is None else lr), float(self.betas[0]), float(self.betas[1]),
             float(self.eps), float(self.weight_decay), self.step_count, 1.0, stream_ptr())

    def state_dict(self):
        """Snapshot (cloned tensors) of the flat optimizer state."""
        return {'exp_avg': self.exp_avg.clone(), 'exp_inf': self.exp_inf.clone(), 'step': self.step_count,
                'lr': self.lr}

    def load_state_dict(self, sd):
        if 'state' in sd and 'param_groups' in sd:  # a torch.optim.Adamax checkpoint of the reference's loop
            return self.load_torch_state_dict(sd)
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_inf.copy_(sd['exp_inf'])
        self.step_count = int(sd['step'])
        self.lr = sd.get('lr', self.lr)

    # -- interchange with torch.optim.Adamax (the 'optimizer_state_dict' of experiments/train_test.py:262-286) --
    def torch_state_dict(self):
        """The same state in torch.optim.Adamax.state_dict() layout; parameter i = i-th of model.parameters()."""
        state = {}
        off = 0
        for i, p in enumerate(self.flat.params):
            k = p.numel()
            state[i] = {'step': torch.tensor(float(self.step_count)),
                        'exp_avg': self.exp_avg[off:off + k].view(p.shape).clone(),
                        'exp_inf': self.exp_inf[off:off + k].view(p.shape).clone()}
            off += k
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': self.weight_decay,
                 'foreach': None, 'maximize': False, 'differentiable': False, 'capturable': False,
                 'params': list(range(len(self.flat.params)))}
        return {'state': state, 'param_groups': [group]}

    def load_torch_state_dict(self, sd):
        group = sd['param_groups'][0]
        ids = list(group['params'])
        if len(ids) != len(self.flat.params):
            raise ValueError('optimizer checkpoint has a different number of parameters')
        off = 0
        step = 0
        for pid, p in zip(ids, self.flat.params):
            k = p.numel()
            st = sd['state'].get(pid)
            if st is None:
                self.exp_avg[off:off + k].zero_()
                self.exp_inf[off:off + k].zero_()
            else:
                self.exp_avg[off:off + k].copy_(st['exp_avg'].reshape(-1))
                self.exp_inf[off:off + k].copy_(st['exp_inf'].reshape(-1))
                step = max(step, int(float(st['step'])))
            off += k
        self.step_count = step
        self.lr = group.get('lr', self.lr)
        self.betas = tuple(group.get('betas', self.betas))
        self.eps = group.get('eps', self.eps)
        self.weight_decay = group.get('weight_decay', self.weight_decay)


def cosine_warm_restarts_lr(step, base_lr, T_0, T_mult=1, eta_min=0.0):
    """Learning rate of torch.optim.lr_scheduler.CosineAnnealingWarmRestarts after `step` calls of scheduler.step() -- the
    reference's schedule (experiments/run.py:96-103: T_0 = batches x epochs, stepped once per batch, train_test.py:173-174) --
    as a host scalar for FusedAdamax.step(lr=...) / Trainer.step(lr=...): step 0 is the lr of the first update."""
    import math
    step, T_0, T_mult = int(step), int(T_0), int(T_mult)
    if T_0 <= 0 or T_mult < 1:
        raise ValueError('T_0 must be positive and T_mult >= 1')
    if T_mult == 1:
        t_cur, t_i = step % T_0, T_0
    else:
        n = int(math.log(step / T_0 * (T_mult - 1) + 1, T_mult)) if step >= T_0 else 0
        t_cur = step - T_0 * (T_mult ** n - 1) // (T_mult - 1)
        t_i = T_0 * T_mult ** n
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


class Trainer:
    """The library's own training step for HNOSegXS (and the engine-backed NeuralOperatorSeg / HartleyMHASeg): forward, fused head+loss on integer labels, backward straight
    into the flat gradient buffer (no autograd graph), one gradient all-reduce, fused Adamax.

    Equivalent to the step body of experiments/train_test.py:146-171 with `to_categorical`, `loss_fn(model(x), y)`,
    `loss.backward()` and `optimizer.step()`; returns the loss as a 1-element device tensor (call .item() to
    reproduce the reference's per-step host sync).
    """

    def __init__(self, model, loss_name='DiceLoss', lr=5e-3, group=None, use_graph=None, loss_param=None):
        from . import ops
        # CUDA graph of forward + loss + backward (168 dependent launches per step): one graph per distinct pair of input
        # buffers, all sharing one memory pool; the all-reduce and the Adamax kernel (whose step count is a host
        # scalar) stay outside.  HNO_GRAPH=0 / use_graph=False launches kernel by kernel.
        self.use_graph = (os.environ.get('HNO_GRAPH', '1') != '0') if use_graph is None else bool(use_graph)
        self._graphs = {}
        self._pool = None
        self._xnorm = {}
        self.model = model
        self.engine = model.engine()
        # CrossEntropyLoss (kind None) is not of the five-moment form: head kernel + one-pass CE kernels + head backward
        self.kind = None if loss_name == 'CrossEntropyLoss' else ops.LOSS_KINDS[loss_name]
        self.loss_param = float(ops.LOSS_DEFAULT_PARAM.get(loss_name, 0.0) if loss_param is None else loss_param)
        self.flat = FlatParameters(model)
        self.slots = self.engine.named_slots()
        self.dst = [self.flat.grad_view_of(p) for p in self.slots]
        import inspect
        # XSEngine writes its gradients straight into the flat buffer; other engines return them and they are copied
        self._direct = 'dst' in inspect.signature(self.engine.run_backward).parameters
        self.optimizer = FusedAdamax(self.flat, lr=lr)
        self.group = group
        # SAMPLE STREAMS (HNO_SAMPLE_STREAMS=1, default off): the samples of a batch are independent until the loss is averaged
        # (every loss term is a mean over (sample, label) pairs), so each sample's forward + loss + backward can run as its own
        # chain of launches on one of two streams; sample b > 0 writes its gradient into a second flat buffer that is added
        # once at the end, the per-sample losses are averaged.  The idea: a quarter of a step is L2-resident, latency-bound work
        # (H stages, spectral core) during which HBM idles, and the other sample's HBM-bound kernels could fill those gaps.
        # MEASURED (profiles/r4b_bench_split.json vs r4b_bench_nosplit.json): 9.56 ms against 9.26 ms for the batched chain --
        # the persistent HBM-bound kernels occupy every SM's shared memory, so kernels of the other stream only start in their
        # tails, and twice the launches (27,100 per 100 steps against 13,600) each pay their prologue.  Kept as an A/B switch.
        self.sample_streams = (os.environ.get('HNO_SAMPLE_STREAMS', '0') == '1' and self._direct and self.kind is not None
                               and getattr(model, 'use_resize', True))
        self._streams = None
        self._flat2 = None

    def _split_state(self, device):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)]
            base = self.flat.grad.data_ptr()
            self._flat2 = torch.zeros_like(self.flat.grad)
            esz = self.flat.grad.element_size()
            self._dst2 = []
            for d in self.dst:
                off = (d.data_ptr() - base) // esz
                self._dst2.append(self._flat2[off:off + d.numel()].view(d.shape))
            self._grad_scale = {}
        return self._streams

    def _loss_and_grad_split(self, x, labels):
        """loss_and_grad of a 2-sample batch with one launch chain per sample on two streams (see __init__)."""
        from . import ops
        from .engine import _labels_u8
        assert x.shape[0] == 2
        streams = self._split_state(x.device)
        cur = torch.cuda.current_stream()
        scale = self._grad_scale.get(2)
        if scale is None:
            scale = self._grad_scale[2] = torch.full((1,), 0.5, dtype=torch.float32, device=x.device)
        losses = torch.empty((2,), dtype=torch.float32, device=x.device)
        lab = _labels_u8(labels, x)
        for b, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                xb, lb = x[b:b + 1], lab[b:b + 1]
                _, S = self.engine.run_forward(xb, save=True, head=False)
                loss_b, coef = ops.head_loss_forward(S.ll, lb, S.tables, S.geom[3], self.kind, self.loss_param)
                losses[b:b + 1].copy_(loss_b)
                self.engine.run_backward(S, dst=self.dst if b == 0 else self._dst2, fused=(lb, coef, scale))
                del S
        for s in streams:
            cur.wait_stream(s)
        self.flat.grad.add_(self._flat2)
        return losses.mean().reshape(1)

    def loss_and_grad(self, x, labels):
        from . import ops
        from .engine import _labels_u8
        if self.sample_streams and x.shape[0] == 2:
            with torch.no_grad():
                return self._loss_and_grad_split(x, labels)
        with torch.no_grad():
            lab = _labels_u8(labels, x)
            if self.kind is None:
                probs, S = self.engine.run_forward(x, save=True)
                loss = ops.ce_loss_forward(probs, labels=lab)
                self._backward(S, dprobs=ops.ce_loss_backward(probs, labels=lab))
                return loss
            if not getattr(self.model, 'use_resize', True):
                # full-resolution network: no interpolation to fuse the loss with; probabilities once, then the moment-based
                # loss kernels on them and the one-hot labels
                from .experiments.utils import to_categorical
                probs, S = self.engine.run_forward(x, save=True)
                onehot = to_categorical(lab[:, None], probs.shape[1], validate=False)
                loss, coef = ops.prob_loss_forward(probs, onehot, self.kind, self.loss_param)
                self._backward(S, dprobs=ops.prob_loss_backward(probs, onehot, coef, None))
                return loss
            perm = self._axis_perm(tuple(x.shape[2:]))
            if perm is not None:
                _, S = self.engine.run_forward(x, save=True, head=False, perm=perm)
                lab = ops.permute_spatial(lab, perm)
            else:
                _, S = self.engine.run_forward(x, save=True, head=False)
            loss, coef = ops.head_loss_forward(S.ll, lab, S.tables, S.geom[3], self.kind, self.loss_param)
            self._backward(S, fused=(lab, coef, None))
        return loss

    def _axis_perm(self, spatial):
        """Spatial permutation the step runs on, or None (engine.preferred_axis_perm; XSEngine only)."""
        from .engine import preferred_axis_perm
        return preferred_axis_perm(self.model, spatial) if self._direct else None

    def _backward(self, S, **kw):
        if self._direct:
            self.engine.run_backward(S, dst=self.dst, **kw)
        else:
            for d, g in zip(self.dst, self.engine.run_backward(S, **kw)):
                d.copy_(g.reshape(d.shape))

    def loss_and_grad_graphed(self, x, labels):
        """loss_and_grad replayed from a CUDA graph captured for exactly these two buffers (contents may change)."""
        global _graph_launches
        key = (x.data_ptr(), labels.data_ptr(), tuple(x.shape), tuple(labels.shape), labels.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 8:  # callers that allocate a fresh batch every step gain nothing from graphs
                self.use_graph = False
                return self.loss_and_grad(x, labels)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # eager pass first: plans, tables and workspaces are created here
                self.loss_and_grad(x, labels)
            cur.wait_stream(side)
            n0 = launches()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._pool):
                loss = self.loss_and_grad(x, labels)
            if self._pool is None:
                self._pool = graph.pool()
            n = launches() - n0
            _graph_launches -= n  # the capture itself launched nothing on the GPU
            ent = self._graphs[key] = (graph, loss, n, x, labels)  # x / labels kept alive: their addresses are baked in
        ent[0].replay()
        _graph_launches += ent[2]
        return ent[1]

    def step(self, x, labels, lr=None):
        loss = self.loss_and_grad_graphed(x, labels) if self.use_graph else self.loss_and_grad(x, labels)
        allreduce_mean_(self.flat.grad, self.group)
        self.optimizer.step(lr)
        return loss

    def step_raw(self, x_raw, labels, lr=None, mask_val=0, clip_val=None, augment=None):
        """step() on RAW modalities in their storage type: `x_raw` (B, C, D, H, W) int16 (or float32) un-normalised
        intensities as the reader returns them (experiments/utils.py:260-270).  The per-sample, per-modality z-scoring the
        reference runs in its loader workers (`x_processing = normalize_modalities(mask_val=0)`, experiments/run.py:52-55,
        data_io/dataset.py:49-50) happens here on the device, so a batch crosses PCIe at 2 bytes per voxel instead of 4.
        The normalised batch lives in one buffer owned by the trainer (its address is what the CUDA graph captured).
        `augment`: an experiments.data_io.ImageTransform; like the reference's loader (data_io/dataset.py:49-56: x_processing
        first, then the transform on image and labels) the NORMALISED batch and its label maps are augmented, one gather
        launch each, into two more trainer-owned buffers."""
        from .experiments.utils import normalize_rows
        key = (tuple(x_raw.shape), x_raw.device)
        buf = self._xnorm.get(key)
        if buf is None:
            buf = self._xnorm[key] = torch.empty(x_raw.shape, dtype=torch.float32, device=x_raw.device)
        normalize_rows(x_raw, x_raw.shape[0] * x_raw.shape[1], mask_val=mask_val, clip_val=clip_val, out=buf)
        if augment is not None:
            if labels.dtype not in (torch.uint8, torch.int16):
                labels = labels.to(torch.uint8)  # class indices; the gather kernel moves 1- / 2- / 4-byte elements
            akey = key + (tuple(labels.shape), labels.dtype)
            abuf = self._xnorm.get(akey)
            if abuf is None:
                abuf = self._xnorm[akey] = (torch.empty_like(buf), torch.empty_like(labels))
            buf, labels = augment.batch(buf, labels, out_x=abuf[0], out_y=abuf[1])
        return self.step(buf, labels, lr)


_graph_launches = 0  # kernels of libhno_b200.so launched through CUDA-graph replays (the library only counts direct ones)


def launches(reset=False):
    """Kernels of libhno_b200.so launched in this process so far (direct launches + graph replays)."""
    global _graph_launches
    n = int(_lib.load().hno_launch_count(1 if reset else 0)) + _graph_launches
    if reset:
        _graph_launches = 0
    return n
