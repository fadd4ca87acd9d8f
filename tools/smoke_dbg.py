import os, sys, torch
sys.path.insert(0, '/root/repo')
from multimodal_3d_image_segmentation_b200 import nets
from oracle import hno_oracle as orc
dev = torch.device('cuda:0')
cfg = dict(in_channels=4, out_channels=4, filters=24, num_transform_blocks=[3] * 8, num_modes=(10, 14, 14))
sd = orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14), seed=0)
model = nets.HNOSegXS(**cfg, device=dev); model.load_state_dict(sd)
g = torch.Generator().manual_seed(1234)
shape = tuple(int(v) for v in os.environ.get("SHAPE", "48,44,40").split(","))
x = torch.randn(1, 4, *shape, generator=g); labels = torch.randint(0, 4, (1, 1) + shape, generator=g)
loss = model.loss(x.to(dev), labels.to(dev), 'DiceLoss'); loss.backward()
flat = torch.cat([p.grad.flatten().cpu() for _, p in model.named_parameters()])
o32_loss, o32 = orc.train_step(sd, x, labels, [3]*8, (10,14,14), 'DiceLoss')
sd64 = {k: v.double() for k, v in sd.items()}
o64_loss, o64 = orc.train_step(sd64, x.double(), labels, [3]*8, (10,14,14), 'DiceLoss')
f32 = torch.cat([o32[k].flatten() for k, _ in model.named_parameters()])
f64 = torch.cat([o64[k].flatten() for k, _ in model.named_parameters()])
r = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
print(os.environ.get('TAG', ''), 'cuda vs fp64', r(flat, f64), '| fp32 oracle vs fp64', r(f32, f64), '| cuda vs fp32 oracle', r(flat, f32), flush=True)
# per-parameter worst
worst = sorted(((r(p.grad.cpu(), o64[k]), k) for k, p in model.named_parameters()), reverse=True)[:4]
print(worst)
