"""Three untimed RND steps at the bench shape (for ncu launch lists / captures): python tools/rnd_step_once.py [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.trainer import RndStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T, dev = 128, "cuda"
torch.manual_seed(0)
m = RND(image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=0.25, gamma=0.99).to(dev)
step = RndStep(m, lr=1e-4)
g = torch.Generator(device=dev); g.manual_seed(1)
frames = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
done = torch.zeros(T, B, device=dev); value = torch.randn(T, B, device=dev, generator=g); boot = torch.randn(B, device=dev, generator=g)
mask = (torch.rand(T, B, device=dev, generator=g) > 0.25).float()
for _ in range(steps):
    loss = step(frames, done, torch.zeros(T, B, device=dev), value, boot, mask)[0]
torch.cuda.synchronize()
print("loss", float(loss))
