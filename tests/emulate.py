"""numpy emulation of what the DHT kernels compute FROM THE PLAN TABLES (test infrastructure).

It mirrors the kernel structure stage by stage (folded analysis along D and H, plain analysis along W,
8-term recombination, and the exact transposes), so the C++ table builder and the decomposition are
verified on a CPU-only box; the GPU tests then only have to show that the kernels follow the same maths.
"""
import numpy as np


def _fold_analysis(x, fcos, fsin, JC, JS, axis):
    """out[j] = sum_i f_j(i) x[i] along `axis` with even/odd folding, like k_analysis_outer."""
    x = np.moveaxis(x, axis, 0).astype(np.float64)
    n = x.shape[0]
    nh = n // 2
    accC = fcos[0, :JC].reshape((JC,) + (1,) * (x.ndim - 1)) * x[0]
    accS = np.zeros((JS,) + x.shape[1:])
    for i in range(1, (n - 1) // 2 + 1):
        e, o = x[i] + x[n - i], x[i] - x[n - i]
        accC = accC + fcos[i, :JC].reshape((JC,) + (1,) * (x.ndim - 1)) * e
        accS = accS + fsin[i, :JS].reshape((JS,) + (1,) * (x.ndim - 1)) * o
    if n % 2 == 0 and n > 1:
        accC = accC + fcos[nh, :JC].reshape((JC,) + (1,) * (x.ndim - 1)) * x[nh]
    return np.moveaxis(np.concatenate([accC, accS], 0), 0, axis)


def _fold_synthesis(g, fcos, fsin, JC, JS, n, axis):
    g = np.moveaxis(g, axis, 0).astype(np.float64)
    C, S = g[:JC], g[JC:JC + JS]
    out = np.zeros((n,) + g.shape[1:])
    nh = n // 2
    out[0] = np.tensordot(fcos[0, :JC], C, 1)
    for i in range(1, (n - 1) // 2 + 1):
        e = np.tensordot(fcos[i, :JC], C, 1)
        o = np.tensordot(fsin[i, :JS], S, 1) if JS else 0.0
        out[i] = e + o
        out[n - i] = e - o
    if n % 2 == 0 and n > 1:
        out[nh] = np.tensordot(fcos[nh, :JC], C, 1)
    return np.moveaxis(out, 0, axis)


def forward(plan, x, scale):
    """x [..., D, H, W] -> z [..., Ld, Lh, Lw]"""
    ax = plan.axes
    g = _fold_analysis(x, plan.table(0, 'fcos'), plan.table(0, 'fsin'), ax[0]['JC'], ax[0]['JS'], -3)
    g = _fold_analysis(g, plan.table(1, 'fcos'), plan.table(1, 'fsin'), ax[1]['JC'], ax[1]['JS'], -2)
    T = np.einsum('...w,jw->...j', g, plan.table(2, 'full').astype(np.float64))
    kd, kh, kw = (plan.table(a, 'kdesc') for a in range(3))
    Z = np.zeros(x.shape[:-3] + plan.modes_shape)
    for a in range(len(kd)):
        cd, sd, gd = kd[a][:3]
        for b in range(len(kh)):
            ch, sh, gh = kh[b][:3]
            for c in range(len(kw)):
                cw, sw, gw = kw[c][:3]
                v = T[..., cd, ch, cw].copy()
                if sh >= 0 and sw >= 0:
                    v -= gh * gw * T[..., cd, sh, sw]
                if sd >= 0 and sw >= 0:
                    v -= gd * gw * T[..., sd, ch, sw]
                if sd >= 0 and sh >= 0:
                    v -= gd * gh * T[..., sd, sh, cw]
                if sd >= 0:
                    v += gd * T[..., sd, ch, cw]
                if sh >= 0:
                    v += gh * T[..., cd, sh, cw]
                if sw >= 0:
                    v += gw * T[..., cd, ch, sw]
                if sd >= 0 and sh >= 0 and sw >= 0:
                    v -= gd * gh * gw * T[..., sd, sh, sw]
                Z[..., a, b, c] = scale * v
    return Z


def adjoint(plan, z, scale):
    """z [..., Ld, Lh, Lw] -> x [..., D, H, W]  (transpose of forward, same scale semantics)"""
    ax = plan.axes
    jd, jh, jw = (plan.table(a, 'jdesc') for a in range(3))
    T = np.zeros(z.shape[:-3] + (ax[0]['J'], ax[1]['J'], ax[2]['J']))
    for a in range(len(jd)):
        for b in range(len(jh)):
            for c in range(len(jw)):
                nsin = jd[a][2] + jh[b][2] + jw[c][2]
                sign = -1.0 if nsin >= 2 else 1.0
                acc = 0.0
                for pa in range(2):
                    if jd[a][pa] < 0:
                        continue
                    fd = -1.0 if (jd[a][2] and pa == 1) else 1.0
                    for pb in range(2):
                        if jh[b][pb] < 0:
                            continue
                        fh = -fd if (jh[b][2] and pb == 1) else fd
                        for pc in range(2):
                            if jw[c][pc] < 0:
                                continue
                            fw = -fh if (jw[c][2] and pc == 1) else fh
                            acc = acc + fw * z[..., jd[a][pa], jh[b][pb], jw[c][pc]]
                T[..., a, b, c] = sign * scale * acc
    g = np.einsum('...j,jw->...w', T, plan.table(2, 'full').astype(np.float64))
    g = _fold_synthesis(g, plan.table(1, 'fcos'), plan.table(1, 'fsin'), ax[1]['JC'], ax[1]['JS'], ax[1]['n'], -2)
    return _fold_synthesis(g, plan.table(0, 'fcos'), plan.table(0, 'fsin'), ax[0]['JC'], ax[0]['JS'], ax[0]['n'], -3)


def interp_matrix(tables, a):
    """Dense [hi, lo] interpolation matrix of one axis from the tables."""
    i0, i1, l1, s, e = tables.axis(a)
    M = np.zeros((tables.hi[a], tables.lo[a]))
    for o in range(tables.hi[a]):
        M[o, i0[o]] += 1.0 - l1[o]
        M[o, i1[o]] += l1[o]
    return M


def stem_input_gradient(dpre, weight, image):
    """What k_stem_dx computes (csrc/stem_kernels.cu): with stride == kernel == 2 and padding 1 every image voxel z feeds exactly
    one output voxel (z + 1) >> 1 through tap (z + 1) & 1 per axis.  dpre (B, F, D, H, W), weight (F, C, 2, 2, 2)."""
    B, F = dpre.shape[:2]
    C = weight.shape[1]
    Dx, Hx, Wx = image
    zd, zh, zw = np.meshgrid(np.arange(Dx), np.arange(Hx), np.arange(Wx), indexing='ij')
    d, h, w = (zd + 1) >> 1, (zh + 1) >> 1, (zw + 1) >> 1
    kd, kh, kw = (zd + 1) & 1, (zh + 1) & 1, (zw + 1) & 1
    g = dpre[:, :, d, h, w].astype(np.float64)                       # (B, F, Dx, Hx, Wx)
    wt = weight[:, :, kd, kh, kw].astype(np.float64)                 # (F, C, Dx, Hx, Wx)
    return np.einsum('bfdhw,fcdhw->bcdhw', g, wt)


def resample_index_walk(N, W, H, V=4):
    """The (d, h, w) walk of k_affine_resample_nn<T, V> (csrc/input_kernels.cu): a thread decodes the first of its V consecutive
    linear indices with two 32-bit divisions and steps the rest with carries.  Returns an (N, 3) int array."""
    out = np.zeros((N, 3), dtype=np.int64)
    for g in range(N // V):
        i = g * V
        t = i // W
        w = i - t * W
        d = t // H
        h = t - d * H
        for e in range(V):
            out[i + e] = (d, h, w)
            w += 1
            if w == W:
                w = 0
                h += 1
                if h == H:
                    h = 0
                    d += 1
    return out
