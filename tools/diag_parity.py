"""Measured parity numbers on the GPU (diagnostic, prints only): per-element bonus errors, per-model prediction
errors of the Disagreement ensemble, and gradient cosine / norm ratios against the oracle's autograd at n >= 1024."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import recipes as R
from helpers import golden
from oracle import curiosity_oracle as O
from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement

DEV = "cuda"
torch.set_num_threads(os.cpu_count())


def rel(a, b):
    a, b = a.detach().cpu().double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    r = (a - b).abs() / b.abs().clamp_min(1e-12)
    return float(r.max()), float(r.mean())


def disagreement_case(T, B, A, seed, tag):
    p = R.make_params(R.disagreement_shapes(A, "idf_burda"), seed)
    m = Disagreement(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    m.load_state_dict(p)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    ref = O.disagreement_bonus(p, obs, nxt, act)
    bonus = m.compute_bonus(obs.to(DEV), nxt.to(DEV), act.to(DEV))
    print(tag, "bonus rel max/mean", rel(bonus, ref))
    with torch.no_grad():
        phi1, phi2, preds, _ = O.disagreement_forward(p, obs, nxt, act)
    out = m._cached_forward(obs.to(DEV), nxt.to(DEV), act.to(DEV), remember=False)
    phi1_b, phi2_f, pred = out[5], out[6], out[10]
    n, F = T * B, 512
    e1 = (phi1_b.float().cpu() - phi1.reshape(n, F)).norm() / phi1.norm()
    e2 = (phi2_f.cpu() - phi2.reshape(n, F)).norm() / phi2.norm()
    print(tag, "phi1_b rel-norm err", float(e1), "phi2_f", float(e2))
    for e in range(4):
        pe = pred[:, e * F:(e + 1) * F].cpu()
        pr = preds[e].reshape(n, F)
        print(tag, f"model {e}: pred rel-norm err {float((pe - pr).norm() / pr.norm()):.5f}  max abs {float((pe - pr).abs().max()):.4f} scale {float(pr.abs().mean()):.4f}")
    var_t = torch.var(torch.stack([pred[:, e * F:(e + 1) * F] for e in range(4)]), dim=0).mean(-1).cpu()
    print(tag, "torch.var of GPU preds vs oracle", rel(var_t, ref.reshape(-1)), " kernel vs torch.var", rel(bonus.reshape(-1), var_t))
    # non-grouped path
    m._grouped.grouped = False
    bonus2 = m.compute_bonus(obs.to(DEV), nxt.to(DEV), act.to(DEV))
    print(tag, "non-grouped bonus rel max/mean", rel(bonus2, ref))


def grad_case(cls_name, T, B, A, seed):
    shapes = R.icm_shapes(A, "idf_burda") if cls_name == "icm" else R.disagreement_shapes(A, "idf_burda")
    p = R.make_params(shapes, seed)
    cls = ICM if cls_name == "icm" else Disagreement
    m = cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    m.load_state_dict(p)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.3)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    t0 = time.time()
    if cls_name == "icm":
        b_ref = O.icm_bonus(p, obs, nxt, act)
        inv_r, fwd_r = O.icm_loss(pr, obs, nxt, act, valid)
    else:
        b_ref = O.disagreement_bonus(p, obs, nxt, act)
        inv_r, fwd_r = O.disagreement_loss(pr, obs, nxt, act, valid, dropout_masks=None)
    (inv_r + fwd_r).backward()
    print(cls_name, f"n={T*B} oracle {time.time()-t0:.1f}s")
    d = lambda t: t.to(DEV)
    print(cls_name, "bonus rel", rel(m.compute_bonus(d(obs), d(nxt), d(act)), b_ref))
    kw = {} if cls_name == "icm" else dict(dropout_masks=None)
    inv, fwd = m.compute_loss(d(obs), d(nxt), d(act), d(valid), **kw)
    (inv + fwd).backward()
    print(cls_name, "inv rel", abs(float(inv) - float(inv_r)) / float(inv_r), "fwd rel", abs(float(fwd) - float(fwd_r)) / float(fwd_r))
    worst = []
    for k, prm in m.named_parameters():
        gref, gg = pr[k].grad.flatten().double(), prm.grad.cpu().flatten().double()
        cos = float(gg @ gref / (gg.norm() * gref.norm()).clamp_min(1e-30))
        worst.append((cos, abs(float(gg.norm() / gref.norm()) - 1), k))
    worst.sort()
    for w in worst[:8]:
        print("   low cos", w)
    worst.sort(key=lambda x: -x[1])
    for w in worst[:8]:
        print("   norm err", w)


if __name__ == "__main__":
    g = golden("disagreement")
    disagreement_case(int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"]), "golden")
    disagreement_case(16, 16, 7, 77, "n256")
    grad_case("icm", 32, 32, 4, 500)
    grad_case("disagreement", 32, 32, 7, 600)
