import sys, os
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_3d_image_segmentation_b200 import ops
cuda = torch.device('cuda:0')
def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a - b).norm() / b.norm())
for (S, P, HW) in [(5120, 1024, 1000), (4100, 4100, 4100), (128 * 700, 1024, 1000)]:
    for ci2 in (0, 24):
        g = torch.Generator().manual_seed(7)
        B, ci1, co = 2, 24, 24
        in1 = F.selu(torch.randn(B, ci1, S, generator=g))
        in2 = torch.randn(B, ci2, S, generator=g) if ci2 else None
        w = torch.randn(co, ci1 + ci2, generator=g) / np.sqrt(ci1 + ci2)
        b = torch.randn(co, generator=g) * 0.1
        dy = torch.randn(B, co, S, generator=g)
        live = ((torch.arange(S) % P) < HW).double()
        r1 = in1.double().requires_grad_(True)
        r2 = in2.double().requires_grad_(True) if ci2 else None
        rw = w.double().requires_grad_(True); rb = b.double().requires_grad_(True)
        x = r1 if r2 is None else torch.cat([r1, r2], 1)
        yr = F.selu(torch.einsum('oi,bis->bos', rw, x) + rb.view(1, -1, 1))
        (yr * dy.double() * live).sum().backward()
        dev = lambda t: None if t is None else t.to(cuda)
        y = ops.pwconv_forward(dev(in1), dev(in2), dev(w), dev(b), 1, False)
        din1, din2, dw, db = ops.pwconv_backward(dev(dy), y, dev(in1), dev(in2), dev(w), 1, False, hw=(P, HW), has_bias=True)
        print(S, ci2, 'din1', rel(din1, r1.grad), 'din2', rel(din2, r2.grad) if ci2 else None, 'dw', rel(dw, rw.grad), 'db', rel(db, rb.grad))
        if ci2:
            e = (dw.cpu().double() - rw.grad).abs()
            print('   dw err by column block:', float(e[:, :24].max()), float(e[:, 24:].max()))
            e2 = (din2.cpu().double() - r2.grad).abs().amax(dim=(0, 1))
            bad = torch.nonzero(e2 > 1e-4).flatten()
            print('   din2 bad voxels:', bad.numel(), bad[:10].tolist(), bad[-5:].tolist())
