"""CPU oracle for the HNOSeg-XS spectral hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain PyTorch CPU ops (fp32 or fp64), of what the reference
(IBM/multimodal-3d-image-segmentation, read-only at /root/reference in the build container) computes
on this path.  It exists so that the CUDA kernels can be checked on a GPU box where the reference is
not available.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import it; the product package never does (it has no CPU path at all).

Pinning: the reference ships no tests, golden vectors or fixtures for this path;
its only known answer is the 28,248 parameter count (README.md:57-63).  The oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports /root/reference/nets in the build
container, runs both, asserts agreement and writes small fixtures to tests/golden/, which
tests/test_oracle_golden.py replays everywhere.  The arithmetic underneath (torch.fft, conv) lives in
PyTorch (unpinned in the reference's pyproject.toml:20; 2.11.0+cu128 here).

Each function cites the reference lines it restates.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805


# ------------------------------------------------------------------------------------------------
# Discrete Hartley transform                                      nets/dht.py:16-36
# ------------------------------------------------------------------------------------------------
def dhtn(x, dims=(-3, -2, -1), inverse=False):
    """H = Re(F) - Im(F) of the complex DFT over `dims`; forward carries 1/N, inverse is unscaled."""
    spec = torch.fft.fftn(x, dim=dims, norm='backward' if inverse else 'forward')
    return spec.real - spec.imag


def cas_matrix(n, ks):
    """Dense 1-D kernel cos+sin(2*pi*k*i/n) for the listed frequencies, fp64 numpy [len(ks), n]."""
    k = np.asarray(ks, dtype=np.int64)[:, None]
    i = np.arange(n, dtype=np.int64)[None, :]
    ang = 2.0 * np.pi * ((k * i) % n) / n
    return np.cos(ang) + np.sin(ang)


def dht3_dense(x, klists, scale):
    """Direct O(N * modes) evaluation of the NON-separable 3-D DHT rows (definition in SURVEY.md 7.3):
    Z[kd,kh,kw] = scale * sum x[d,h,w] cas(2 pi (kd d/D + kh h/H + kw w/W)).  fp64 numpy, small sizes only.
    cas(a+b+c) is expanded through cos/sin of the per-axis angles."""
    x = np.asarray(x, dtype=np.float64)
    D, H, W = x.shape[-3:]
    cs = []
    for n, ks in zip((D, H, W), klists):
        k = np.asarray(ks, dtype=np.int64)[:, None]
        i = np.arange(n, dtype=np.int64)[None, :]
        ang = 2.0 * np.pi * ((k * i) % n) / n
        cs.append((np.cos(ang), np.sin(ang)))
    (cd, sd), (ch, sh), (cw, sw) = cs

    def proj(fd, fh, fw):
        return np.einsum('...dhw,ad,bh,cw->...abc', x, fd, fh, fw, optimize=True)

    cos_part = proj(cd, ch, cw) - proj(cd, sh, sw) - proj(sd, ch, sw) - proj(sd, sh, cw)
    sin_part = proj(sd, ch, cw) + proj(cd, sh, cw) + proj(cd, ch, sw) - proj(sd, sh, sw)
    return scale * (cos_part + sin_part)


def clamp_modes(modes, spatial):
    """nets/hnosegxs.py:382-387: a mode count that does not fit twice into the axis becomes size // 2."""
    return tuple(s // 2 if 2 * m > s else m for m, s in zip(modes, spatial))


def corner_indices(n, m):
    """Retained frequencies of one axis in output order: the low block then the high block
    (nets/hnosegxs.py:393-410 slices [:m] and [-m:] and concatenates low first)."""
    return list(range(m)) + list(range(n - m, n))


def transform_crop(x, modes):
    """TransformCrop.forward for 5-D input (nets/hnosegxs.py:378-410)."""
    spatial = x.shape[2:]
    modes = clamp_modes(modes, spatial)
    spec = dhtn(x)
    for axis, (n, m) in enumerate(zip(spatial, modes)):
        idx = torch.tensor(corner_indices(n, m), dtype=torch.long)
        spec = spec.index_select(2 + axis, idx)
    return spec


def pad_inverse(z, spatial):
    """PadInverse.forward (nets/hnosegxs.py:454-494): mode counts are inferred as shape // 2, the corners
    are scattered into a zero spectrum of the target size, then the unnormalised inverse DHT is taken."""
    modes = tuple(s // 2 for s in z.shape[2:])
    assert all(n >= 2 * m for n, m in zip(spatial, modes))
    full = z
    for axis, (n, m) in enumerate(zip(spatial, modes)):
        shape = list(full.shape)
        shape[2 + axis] = n
        grown = torch.zeros(shape, dtype=z.dtype)
        idx = torch.tensor(corner_indices(n, m), dtype=torch.long)
        grown.index_copy_(2 + axis, idx, full)
        full = grown
    return dhtn(full, inverse=True)


# ------------------------------------------------------------------------------------------------
# Hartley operator on already-cropped modes                      nets/hartley_operator.py:287-333
# ------------------------------------------------------------------------------------------------
def reverse_modes(t):
    """get_reverse (hartley_operator.py:320-333): index j -> (n - j) mod n on the last three axes."""
    for axis in (-3, -2, -1):
        n = t.shape[axis]
        idx = torch.tensor([(n - j) % n for j in range(n)], dtype=torch.long)
        t = t.index_select(axis, idx)
    return t


def hartley_mix(z, weight):
    """HartleyOperator._call3d_notransform (hartley_operator.py:287-299, 302-317)."""
    if weight.ndim == 2:  # shared: one (O, I) matrix for every mode
        return torch.einsum('oi,bidhw->bodhw', weight, z)
    zr, wr = reverse_modes(z), reverse_modes(weight)
    even = torch.einsum('oidhw,bidhw->bodhw', weight, z + zr)
    odd = torch.einsum('oidhw,bidhw->bodhw', wr, z - zr)
    return 0.5 * (even + odd)


def selu(x):
    return F.selu(x)


def pointwise(x, weight, bias=None):
    """Conv3d with a 1x1x1 kernel (nets/nets_utils.py:162-163); weight (O, I, 1, 1, 1) or (O, I)."""
    w = weight.reshape(weight.shape[0], weight.shape[1])
    y = torch.einsum('oi,bidhw->bodhw', w, x)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return y


def hartley_operator_with_transform(x, weight, modes, bias=None):
    """HartleyOperator._call3d, shared weights (hartley_operator.py:168-271): DHT, mix the corners, pad,
    (+bias), SELU in the frequency domain, inverse DHT."""
    assert weight.ndim == 2
    spatial = x.shape[2:]
    z = hartley_mix(transform_crop(x, modes), weight)
    modes = tuple(s // 2 for s in z.shape[2:])
    full = z
    for axis, (n, m) in enumerate(zip(spatial, modes)):
        shape = list(full.shape)
        shape[2 + axis] = n
        grown = torch.zeros(shape, dtype=z.dtype)
        grown.index_copy_(2 + axis, torch.tensor(corner_indices(n, m)), full)
        full = grown
    if bias is not None:
        full = full + bias
    return dhtn(selu(full), inverse=True)


def hartley_operator_with_transform_individual(x, weight, modes):
    """HartleyOperator._call3d with weights_type='individual', no bias (hartley_operator.py:179, 196-241, 243-271): the
    even/odd recombination pairs every retained mode with its reversal in the FULL spectrum (get_reverse of the
    transformed input, index j -> (n - j) mod n), but W with its reversal inside its own (2 m0, 2 m1, 2 m2) block; the
    weight's first / last m entries per axis serve the low / high corner.  Then pad, SELU, inverse DHT."""
    assert weight.ndim == 5
    spatial = tuple(x.shape[2:])
    assert all(n >= 2 * m for n, m in zip(spatial, modes))
    X = dhtn(x)
    Xr, Wr = reverse_modes(X), reverse_modes(weight)
    full = torch.zeros((x.shape[0], weight.shape[0]) + spatial, dtype=x.dtype)
    corners = [((slice(0, m), slice(0, m)), (slice(n - m, n), slice(m, 2 * m))) for n, m in zip(spatial, modes)]
    for sd_, wd_ in corners[0]:
        for sh_, wh_ in corners[1]:
            for sw_, ww_ in corners[2]:
                xs, xr = X[:, :, sd_, sh_, sw_], Xr[:, :, sd_, sh_, sw_]
                even = torch.einsum('oidhw,bidhw->bodhw', weight[:, :, wd_, wh_, ww_], xs + xr)
                odd = torch.einsum('oidhw,bidhw->bodhw', Wr[:, :, wd_, wh_, ww_], xs - xr)
                full[:, :, sd_, sh_, sw_] = 0.5 * (even + odd)
    return dhtn(selu(full), inverse=True)


# ------------------------------------------------------------------------------------------------
# Hartley multi-head attention (oracle only so far)   nets/hartley_mha.py:136-222
# ------------------------------------------------------------------------------------------------
def group_patches(t, patch):
    """grouping3d (hartley_mha.py:473-498): (B, Z, C, D, H, W) -> (B, Z, C*pd*ph*pw, D/pd, H/ph, W/pw); the new channel
    index is ((c*pd + i)*ph + j)*pw + k for the voxel (i, j, k) inside its patch."""
    pd, ph, pw = patch
    b, z, c, d, h, w = t.shape
    assert d % pd == 0 and h % ph == 0 and w % pw == 0
    u = t.unfold(3, pd, pd).unfold(4, ph, ph).unfold(5, pw, pw)        # (b, z, c, nd, nh, nw, pd, ph, pw)
    u = u.permute(0, 1, 2, 6, 7, 8, 3, 4, 5)                           # (b, z, c, pd, ph, pw, nd, nh, nw)
    return u.reshape(b, z, c * pd * ph * pw, d // pd, h // ph, w // pw)


def ungroup_patches(t, channels, patch):
    """ungrouping3d (hartley_mha.py:501-524): the inverse of group_patches."""
    pd, ph, pw = patch
    b, z, _, nd, nh, nw = t.shape
    u = t.reshape(b, z, channels, pd, ph, pw, nd, nh, nw).permute(0, 1, 2, 6, 3, 7, 4, 8, 5)
    return u.reshape(b, z, channels, nd * pd, nh * ph, nw * pw)


def hartley_mha(query, w_query, w_key, w_value, w_out, modes, patch=None, key=None, value=None, activation=selu):
    """HartleyMultiHeadAttention._call, 3-D, no bias (hartley_mha.py:136-222): DHT of every input, per-head channel mixing
    of the 8 retained corners (freq_conv3d :312-334 = TransformCrop followed by 'zoi,bidhw->bzodhw'), optional patch
    grouping, attention act(Q^T K / sqrt(features)) V over the retained (grouped) modes -- no softmax --, ungrouping, heads
    merged head-major, output projection, zero-padding and inverse DHT (inverse3d :365-400 = PadInverse).
    key defaults to query and value to key (:137-150).  Weights: (Z, Ck, Cin), (Z, Ck, Cin_k), (Z, Cv, Cin_v),
    (Cv, Z*Cv)."""
    spatial = tuple(query.shape[2:])
    assert all(n >= 2 * m for n, m in zip(spatial, modes))
    key = query if key is None else key
    value = key if value is None else value
    q = torch.einsum('zoi,bidhw->bzodhw', w_query, transform_crop(query, modes))
    k = torch.einsum('zoi,bidhw->bzodhw', w_key, transform_crop(key, modes))
    v = torch.einsum('zoi,bidhw->bzodhw', w_value, transform_crop(value, modes))
    if patch is not None:
        q, k, v = (group_patches(t, patch) for t in (q, k, v))
    grid = q.shape[3:]
    q, k, v = (t.flatten(3) for t in (q, k, v))                        # (B, Z, features, tokens)
    att = torch.einsum('bzcq,bzck->bzqk', q, k) / math.sqrt(k.shape[2])
    if activation is not None:
        att = activation(att)
    out = torch.einsum('bzqk,bzck->bzcq', att, v).reshape(v.shape[:3] + tuple(grid))
    if patch is not None:
        out = ungroup_patches(out, w_value.shape[1], patch)
    out = out.flatten(1, 2)                                            # heads x channels, head-major
    out = torch.einsum('oi,bidhw->bodhw', w_out, out)
    return pad_inverse(out, spatial)


# ------------------------------------------------------------------------------------------------
# HNOSeg-XS                                                       nets/hnosegxs.py
# ------------------------------------------------------------------------------------------------
def xs_block(x, sd, prefix, num_convs, modes):
    """HNOXSBlock.forward (hnosegxs.py:253-279) with NeuralOperatorBlock (:307-329) inlined."""
    key = prefix + 'mapping_conv.op.weight'
    if key in sd:
        x = selu(pointwise(x, sd[key], sd[prefix + 'mapping_conv.op.bias']))
    skip = x
    z = transform_crop(x, modes)
    for j in range(num_convs):
        z = selu(hartley_mix(z, sd[f'{prefix}conv_blocks.{j}.op.weight']) + z)
    y = selu(pad_inverse(z, x.shape[2:]))
    key = prefix + 'conv_concat.op.weight'
    if key in sd:
        y = selu(pointwise(torch.cat([y, skip], dim=1), sd[key], sd[prefix + 'conv_concat.op.bias']))
    else:
        y = y + skip
    return y


def center_padcrop(x, target):
    """spatial_padcrop (nets/nets_utils.py:22-99): symmetric pad/crop, the odd voxel goes to the upper side."""
    for axis, t in enumerate(target):
        n = x.shape[2 + axis]
        if t > n:
            lo = (t - n) // 2
            pad = [0, 0] * (x.ndim - 2)
            pos = (x.ndim - 3 - axis) * 2
            pad[pos], pad[pos + 1] = lo, t - n - lo
            x = F.pad(x, pad)
        elif t < n:
            lo = (n - t) // 2
            x = x.narrow(2 + axis, lo, t)
    return x


def hnosegxs_forward(sd, x, num_transform_blocks, num_modes, use_resize=True, use_unet_skip=True,
                     softmax=True, return_logits=False):
    """HNOSegXS.forward (hnosegxs.py:145-182) driven by a reference-layout state_dict."""
    image_size = x.shape[2:]
    if use_resize:  # Conv3d(k=2, s=2, p=1) + SELU                         hnosegxs.py:102-105,150-151
        x = selu(F.conv3d(x, sd['conv_in.op.weight'], sd['conv_in.op.bias'], stride=2, padding=1))
    x = selu(pointwise(x, sd['conv1.op.weight'], sd['conv1.op.bias']))
    nb = len(num_transform_blocks)
    stash = {}
    for i, n_convs in enumerate(num_transform_blocks):
        if use_unet_skip and i > nb // 2:                                   # hnosegxs.py:161-162
            x = torch.cat([x, stash[nb - 1 - i]], dim=1)
        x = xs_block(x, sd, f'layers.{i}.', n_convs, num_modes)
        if use_unet_skip and i < nb // 2:                                   # hnosegxs.py:168-169
            stash[i] = x
    if use_resize:
        x = F.interpolate(x, size=tuple(image_size), mode='trilinear')     # hnosegxs.py:174-176
    logits = center_padcrop(pointwise(x, sd['conv_out.weight']), image_size)  # :178-179
    out = torch.softmax(logits, dim=1) if softmax else logits
    return (out, logits) if return_logits else out


# ------------------------------------------------------------------------------------------------
# HNOSeg (NeuralOperatorSeg, transform_type='Hartley')            nets/architectures.py
# ------------------------------------------------------------------------------------------------
def fourier_operator_with_transform(x, weight_real, weight_imag, modes):
    """FourierOperator._call3d, no bias (fourier_operator.py:148-211): rfftn(norm='forward'), mix the four retained
    corners of the half-spectrum with the complex weight -- (O, I) shared, or (O, I, 2 m0, 2 m1, m2) 'individual' whose
    first / last m entries per axis belong to the low / high corner (:165-187) --, zero-pad, irfftn(norm='forward')."""
    s0, s1, s2 = x.shape[2:]
    individual = weight_real.ndim == 5
    if individual:  # :159 no clamping, the grid must hold the modes
        m0, m1, m2 = modes
        assert s0 >= 2 * m0 and s1 >= 2 * m1 and s2 >= 2 * m2
    else:
        m0, m1, m2 = (s // 2 if 2 * m > s else m for m, s in zip(modes, (s0, s1, s2)))
    f = torch.fft.rfftn(x, dim=(-3, -2, -1), norm='forward')
    w = torch.complex(weight_real, weight_imag)
    full = torch.zeros((x.shape[0], w.shape[0], s0, s1, m2), dtype=f.dtype)
    for sd_, wd_ in ((slice(0, m0), slice(0, m0)), (slice(s0 - m0, s0), slice(m0, 2 * m0))):
        for sh_, wh_ in ((slice(0, m1), slice(0, m1)), (slice(s1 - m1, s1), slice(m1, 2 * m1))):
            if individual:
                full[:, :, sd_, sh_, :] = torch.einsum('oidhw,bidhw->bodhw', w[:, :, wd_, wh_, :m2],
                                                       f[:, :, sd_, sh_, :m2])
            else:
                full[:, :, sd_, sh_, :] = torch.einsum('oi,bidhw->bodhw', w, f[:, :, sd_, sh_, :m2])
    return torch.fft.irfftn(full, s=(s0, s1, s2), dim=(-3, -2, -1), norm='forward')


def hno_block(x, sd, prefix, modes, use_block_skip=True, patch=None):
    """NeuralOperatorBlock via _TransBlock.forward (architectures.py:521-548, 551-608), shared weights, SELU:
    spectral layer with its own transform pair + 1x1x1 conv branch -> SELU -> concat skip conv (or additive skip)."""
    if prefix + 'op.weight_query' in sd:  # HartleyMHABlock (architectures.py:611-635)
        x1 = hartley_mha(x, sd[prefix + 'op.weight_query'], sd[prefix + 'op.weight_key'], sd[prefix + 'op.weight_value'],
                         sd[prefix + 'op.weight_out'], modes, patch)
    elif prefix + 'op.weight_real' in sd:  # transform_type='Fourier' (FNOSeg): no activation in the frequency domain
        x1 = fourier_operator_with_transform(x, sd[prefix + 'op.weight_real'], sd[prefix + 'op.weight_imag'], modes)
    elif sd[prefix + 'op.weight'].ndim == 5:  # weights_type='individual'
        x1 = hartley_operator_with_transform_individual(x, sd[prefix + 'op.weight'], modes)
    else:
        x1 = hartley_operator_with_transform(x, sd[prefix + 'op.weight'], modes)
    x2 = pointwise(x, sd[prefix + 'conv_branch.weight'], sd.get(prefix + 'conv_branch.bias'))
    y = selu(x1 + x2)
    if not use_block_skip:  # config_fno.ini: the original FNO block
        return y
    key = prefix + 'conv_concat.op.weight'
    if key in sd:
        return selu(pointwise(torch.cat([y, x], dim=1), sd[key], sd[prefix + 'conv_concat.op.bias']))
    return y + x


def hnoseg_forward(sd, x, num_transform_blocks, num_modes, softmax=True, return_logits=False, use_block_skip=True,
                   patch=None):
    """_TransSeg.forward (architectures.py:321-353) for NeuralOperatorSeg(..., 'Hartley'); use_resize and deep supervision
    are read off the state_dict (conv_in / conv_ds present or not, :286-289, :306-311)."""
    image_size = x.shape[2:]
    use_resize = 'conv_in.op.weight' in sd
    if use_resize:
        x = selu(F.conv3d(x, sd['conv_in.op.weight'], sd['conv_in.op.bias'], stride=2, padding=1))
    x = selu(pointwise(x, sd['conv1.op.weight'], sd['conv1.op.bias']))
    deep = 'conv_ds.op.weight' in sd  # use_deep_supervision (architectures.py:306-311, 330-343): every block output joins
    tensors = [x]
    for i in range(num_transform_blocks):
        x = hno_block(x, sd, f'layers.{i}.', num_modes, use_block_skip, patch)
        tensors.append(x)
    if deep:
        x = selu(pointwise(torch.cat(tensors, dim=1), sd['conv_ds.op.weight'], sd['conv_ds.op.bias']))
    if use_resize:
        x = F.interpolate(x, size=tuple(image_size), mode='trilinear')
    logits = center_padcrop(pointwise(x, sd['conv_out.weight']), image_size)
    out = torch.softmax(logits, dim=1) if softmax else logits
    return (out, logits) if return_logits else out


def hnoseg_train_step(sd, x, labels, num_transform_blocks, num_modes, loss='DiceLoss', use_block_skip=True, patch=None):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    probs = hnoseg_forward(params, x, num_transform_blocks, num_modes, use_block_skip=use_block_skip, patch=patch)
    value = LOSSES[loss](probs, to_categorical(labels, probs.shape[1]).to(probs.dtype))
    grads = torch.autograd.grad(value, list(params.values()))
    return value.detach(), dict(zip(params.keys(), grads))


# ------------------------------------------------------------------------------------------------
# Losses                                                          nets/custom_losses.py
# ------------------------------------------------------------------------------------------------
def to_categorical(y, num_classes):
    """experiments/utils.py:74-97: (B,1,D,H,W) integer labels -> (B,C,D,H,W) one-hot float32."""
    assert y.shape[1] == 1
    onehot = F.one_hot(y[:, 0].long(), num_classes)
    return onehot.movedim(-1, 1).to(torch.float32)


def normalize_modalities(data, mask_val=None, clip_val=None):
    """experiments/utils.py:25-71 in numpy: per modality (first axis) clip, masked mean / population std, (x - mean) / std,
    masked voxels (clipped value == mask_val) -> 0.  Statistics in float64 here (numpy's own float32 pairwise sums differ
    from this by ~1e-7 relative; make_golden.py checks the two against each other)."""
    out = []
    for da in np.asarray(data, dtype=np.float32):
        if clip_val is not None:
            da = np.clip(da, *clip_val)
        keep = np.ones(da.shape, bool) if mask_val is None else da != np.float32(mask_val)
        vals = da[keep].astype(np.float64)
        if vals.size == 0:
            out.append(np.zeros_like(da))
            continue
        mean, std = np.float32(vals.mean()), np.float32(vals.std())
        with np.errstate(divide='ignore', invalid='ignore'):
            out.append(np.where(keep, (da - mean) / std, np.float32(0)).astype(np.float32))
    return np.stack(out)


def dice_loss(y_pred, y_true):
    """custom_losses.py:73-111."""
    dims = tuple(range(2, y_true.ndim))
    inter = (y_true * y_pred).sum(dims)
    union = (y_true + y_pred).sum(dims)
    return (1 - 2.0 * inter / (union + 1e-7)).mean()


def pcc_loss(y_pred, y_true):
    """custom_losses.py:17-70."""
    dims = tuple(range(2, y_true.ndim))
    t = y_true - y_true.mean(dims, keepdim=True)
    p = y_pred - y_pred.mean(dims, keepdim=True)
    r = (t * p).sum(dims) / torch.sqrt((t * t).sum(dims) * (p * p).sum(dims) + 1e-7)
    return (1 - (r + 1) * 0.5).mean()


def exp_dice_loss(y_pred, y_true, exp=0.3):
    """custom_losses.py:114-133: mean((-log(clamp(dice, 1e-7, 1 - 1e-7))) ** exp)."""
    dims = tuple(range(2, y_true.ndim))
    dice = 2.0 * (y_true * y_pred).sum(dims) / ((y_true + y_pred).sum(dims) + 1e-7)
    return torch.pow(-torch.log(torch.clamp(dice, 1e-7, 1.0 - 1e-7)), exp).mean()


def cross_entropy_loss(y_pred, y_true):
    """torch.nn.CrossEntropyLoss() the way the reference reaches it: experiments/run.py:105-110 looks `loss_name` up
    in torch.nn when custom_losses has no such class, and train_test.py:159-160 calls loss_fn(model(x), one_hot).  So the
    log-softmax runs over the network's softmax OUTPUT and the target is class probabilities; 'mean' divides by the
    batch x voxel count.  Written out instead of calling F.cross_entropy."""
    lse = torch.logsumexp(y_pred, dim=1, keepdim=True)
    per_voxel = -(y_true * (y_pred - lse)).sum(1)
    return per_voxel.mean()


LOSSES = {'DiceLoss': dice_loss, 'PCCLoss': pcc_loss, 'ExpDiceLoss': exp_dice_loss,
          'CrossEntropyLoss': cross_entropy_loss}


# ------------------------------------------------------------------------------------------------
# Parameters                                                      nets/hnosegxs.py:95-143, nets_utils.py:102-117
# ------------------------------------------------------------------------------------------------
def init_state_dict(in_channels, out_channels, filters, num_transform_blocks, num_modes, weights_type='shared',
                    seed=0, dtype=torch.float32):
    """Random SNN-style parameters with the reference's key names and shapes (kaiming-normal 'linear'
    weights, bias ~ U(-1e-3, 1e-3)).  The draws are NOT RNG-compatible with the reference constructors;
    parity tests always copy one state_dict into both implementations."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, o, i, k=1, bias=True):
        fan_in = i * k ** 3
        sd[name + 'weight'] = torch.randn((o, i, k, k, k), generator=g, dtype=dtype) / math.sqrt(fan_in)
        if bias:
            sd[name + 'bias'] = (torch.rand((o,), generator=g, dtype=dtype) * 2 - 1) * 1e-3

    conv('conv_in.op.', filters, in_channels, 2)
    conv('conv1.op.', filters, filters)
    nb = len(num_transform_blocks)
    for i, n_convs in enumerate(num_transform_blocks):
        pre = f'layers.{i}.'
        if i > nb // 2:
            conv(pre + 'mapping_conv.op.', filters, 2 * filters)
        for j in range(n_convs):
            if weights_type == 'shared':
                shape, fan_in = (filters, filters), filters
            else:
                block = tuple(2 * m for m in num_modes)
                shape, fan_in = (filters, filters) + block, filters * int(np.prod(block))
            sd[f'{pre}conv_blocks.{j}.op.weight'] = torch.randn(shape, generator=g, dtype=dtype) / math.sqrt(fan_in)
        conv(pre + 'conv_concat.op.', filters, 2 * filters)
    sd['conv_out.weight'] = torch.randn((out_channels, filters, 1, 1, 1), generator=g, dtype=dtype) / math.sqrt(filters)
    return sd


def train_step(sd, x, labels, num_transform_blocks, num_modes, loss='DiceLoss'):
    """One step of experiments/train_test.py:146-171 without the optimizer: returns (loss, grads by key)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    probs = hnosegxs_forward(params, x, num_transform_blocks, num_modes, use_resize='conv_in.op.weight' in sd)
    value = LOSSES[loss](probs, to_categorical(labels, probs.shape[1]).to(probs.dtype))
    grads = torch.autograd.grad(value, list(params.values()))
    return value.detach(), dict(zip(params.keys(), grads))


# ------------------------------------------------------------------------------------------------
# Affine augmentation                                 experiments/data_io/dataset.py:63-245
# ------------------------------------------------------------------------------------------------
# PARITY UNPINNED for the resampler itself: the reference delegates it to SimpleITK (pyproject.toml:21, version not
# pinned; not installable here).  affine_resample_nn restates the PUBLISHED semantics of what the reference configures --
# itk::ResampleImageFilter with an itk::AffineTransform (y = A x + t about a zero centre), unit spacing, zero origin,
# identity direction, itk::NearestNeighborInterpolateImageFunction (ConvertContinuousIndexToNearestIndex = round half
# up, IsInsideBuffer = continuous index in [-0.5, size - 0.5)) and SetDefaultPixelValue(cval).  ITK walks a scan line by
# interpolating between its transformed end points, this evaluates A x + t per voxel: the two can differ in the last
# bit, i.e. only on exact rounding ties.  Everything the reference itself computes (which parameters are drawn from the
# generator and in which order, the matrix, the flips) IS pinned: oracle/make_golden.py runs the reference's
# ImageTransform with this resampler behind a stand-in `SimpleITK` module and records matrices and outputs
# (tests/golden/augment.npz).
def affine_resample_nn(x, matrix, offset, cval=0.0):
    """x (C, D, H, W) or (C, H, W) numpy; matrix (n, n), offset (n,) in SimpleITK's (x, y[, z]) order, mapping an OUTPUT
    index to the continuous INPUT index (dataset.py:220-224).  Returns the resampled array (same shape and dtype)."""
    x = np.asarray(x)
    nd = x.ndim - 1
    size_xyz = x.shape[1:][::-1]
    A = np.eye(3)
    t = np.zeros(3)
    A[:nd, :nd] = np.asarray(matrix, dtype=np.float64).reshape(nd, nd)
    t[:nd] = np.asarray(offset, dtype=np.float64)
    W, H = size_xyz[0], size_xyz[1]
    D = size_xyz[2] if nd == 3 else 1
    vol = x.reshape(x.shape[0], D, H, W)
    pz, py, px = np.meshgrid(np.arange(D, dtype=np.float64), np.arange(H, dtype=np.float64),
                             np.arange(W, dtype=np.float64), indexing='ij')
    idx = []
    for r in range(3):  # same operation order as the CUDA kernel, no fused multiply-add
        s = (A[r, 0] * px + A[r, 1] * py) + A[r, 2] * pz
        idx.append(np.floor((s + t[r]) + 0.5))
    inside = ((idx[0] >= 0) & (idx[0] < W) & (idx[1] >= 0) & (idx[1] < H) & (idx[2] >= 0) & (idx[2] < D))
    ix = np.where(inside, idx[0], 0).astype(np.int64)
    iy = np.where(inside, idx[1], 0).astype(np.int64)
    iz = np.where(inside, idx[2], 0).astype(np.int64)
    out = np.where(inside[None], vol[:, iz, iy, ix], np.asarray(cval).astype(x.dtype))
    return out.astype(x.dtype).reshape(x.shape)


def centred_affine(matrix, size_xyz):
    """dataset.py:195-202: the homogeneous matrix conjugated with the translation to size / 2 + 0.5."""
    n = matrix.shape[0]
    centre = np.asarray(size_xyz, dtype=np.float64) / 2.0 + 0.5
    to_c, from_c = np.eye(n), np.eye(n)
    to_c[:-1, -1] = centre
    from_c[:-1, -1] = -centre
    return to_c @ matrix @ from_c


def image_transform(x, y, rng, rotation_range=None, shift_range=None, zoom_range=None, flip=None, cval=0.0,
                    augmentation_probability=1.0):
    """ImageTransform.__call__ (dataset.py:94-181) on numpy arrays with `rng` = np.random.default_rng(seed): the same
    draws in the same order (binomial gate; one uniform per non-zero rotation / shift entry; zoom; one random() per
    enabled flip axis AFTER the resampling).  Returns (x', y', record) with record = dict(matrix, offset, flips)."""
    x = np.asarray(x)
    nd = x.ndim - 1
    rec = dict(matrix=None, offset=None, flips=[False] * nd)
    if not rng.binomial(1, augmentation_probability):
        return x, y, rec
    theta = None
    if rotation_range is not None:
        if np.isscalar(rotation_range):
            theta = np.pi / 180 * rng.uniform(-rotation_range, rotation_range) if rotation_range else 0
        else:
            theta = [np.pi / 180 * rng.uniform(-r, r) if r else 0 for r in rotation_range]
    shift = None
    if shift_range is not None:
        shift = [rng.uniform(-s, s) * x.shape[1 + i] if s else 0 for i, s in enumerate(shift_range)]
    zoom = rng.uniform(zoom_range[0], zoom_range[1]) if zoom_range is not None else None
    M = None
    if theta is not None:
        if np.isscalar(theta):
            if theta != 0:
                c, s = np.cos(theta), np.sin(theta)
                M = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        elif any(t != 0 for t in theta):
            a, b, g = theta[::-1]  # rotations about x, y, z
            ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
            M = np.array([[cb * cg, -ca * sg + sa * sb * cg, sa * sg + ca * sb * cg, 0],
                          [cb * sg, ca * cg + sa * sb * sg, -sa * cg + ca * sb * sg, 0],
                          [-sb, sa * cb, ca * cb, 0],
                          [0, 0, 0, 1]])
    if shift is not None and any(s != 0 for s in shift):
        S = np.eye(nd + 1)
        S[:-1, -1] = np.asarray(shift[::-1])
        M = S if M is None else S @ M
    if zoom is not None and zoom != 1:
        Z = np.eye(nd + 1)
        Z[:-1, :-1] *= zoom
        M = Z if M is None else Z @ M
    if M is not None:
        Mc = centred_affine(M, x.shape[1:][::-1])
        rec['matrix'], rec['offset'] = Mc[:-1, :-1].copy(), Mc[:-1, -1].copy()
        x = affine_resample_nn(x, rec['matrix'], rec['offset'], cval)
        if y is not None:
            y = affine_resample_nn(y, rec['matrix'], rec['offset'], cval)
    if flip is not None:
        for i, f in enumerate(flip):
            if f and rng.random() < 0.5:
                rec['flips'][i] = True
                x = np.flip(x, 1 + i)
                if y is not None:
                    y = np.flip(y, 1 + i)
    return x, y, rec


# the cases oracle/make_golden.py::case_augment records from the reference and tests/test_augment.py replays
AUGMENT_CASES = [
    # name, spatial, kwargs of ImageTransform, number of consecutive calls on one generator
    ('brats_ini', (9, 12, 10), dict(rotation_range=[30, 30, 30], shift_range=[0.2, 0.2, 0.2], zoom_range=[0.8, 1.2],
                                    augmentation_probability=0.8, seed=7), 6),  # [augmentation] of config_hnoseg_xs.ini shape
    ('flips', (7, 8, 9), dict(flip=[True, False, True], seed=3), 5),
    ('all', (8, 9, 11), dict(rotation_range=[20, 0, 45], shift_range=[0.1, 0, 0.3], zoom_range=[0.7, 1.3],
                             flip=[False, True, True], cval=-3.0, augmentation_probability=0.7, seed=11), 6),
    ('shift_only', (6, 7, 8), dict(shift_range=[0.25, 0.25, 0], seed=5), 3),
    ('zoom_only', (6, 7, 8), dict(zoom_range=[0.6, 1.5], seed=9), 3),
    ('planar', (10, 13), dict(rotation_range=25, shift_range=[0.2, 0.1], zoom_range=[0.8, 1.25], flip=[True, True],
                              seed=13), 5),
]
