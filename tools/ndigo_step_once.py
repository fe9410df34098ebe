"""A few NDIGO steps at the bench shape (for ncu launch lists): python tools/ndigo_step_once.py [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curiosity_baselines_b200.curiosity.ndigo import NDIGO
from curiosity_baselines_b200.trainer import FlatAdam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T, A, dev = 128, 5, "cuda"
torch.manual_seed(0)
m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=3, feature_encoding="idf_maze", num_predictors=10).to(dev)
opt = FlatAdam(m.parameters(), lr=1e-4)
g = torch.Generator(device=dev); g.manual_seed(1)
obs = (torch.rand(T, B, 3, 5, 5, device=dev, generator=g) < 0.3).float()
oh = lambda: torch.nn.functional.one_hot(torch.randint(0, A, (T, B), device=dev, generator=g), A).float()
prev, act, valid = oh(), oh(), torch.ones(T, B, device=dev)
for _ in range(steps):
    r = m.compute_bonus(obs, prev, act)
    opt.grad.zero_()
    loss = m.compute_loss(obs, prev, act, valid)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
