"""GAE / discounted return / valid mask on the GPU vs the golden vectors (reference output),
the oracle on fresh seeded inputs, and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

import recipes as R
from helpers import golden
from oracle import curiosity_oracle as O
from curiosity_baselines_b200.algos.utils import (discount_return, generalized_advantage_estimation,
                                                  valid_from_done)

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: GAE/returns within 1e-5 in fp32 (we expect bit-exact)


@pytest.mark.parametrize("tag,done_fn", [("a", R.make_random_done), ("b", R.make_suffix_done),
                                         ("c", R.make_random_done)])
def test_against_reference_golden(tag, done_fn):
    g = golden(f"returns_{tag}")
    T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
    reward, value = R.make_normal(seed, T, B), R.make_normal(seed + 100, T, B)
    boot, done = R.make_normal(seed + 200, B), done_fn(seed + 300, T, B)
    adv, ret = generalized_advantage_estimation(reward.cuda(), value.cuda(), done.cuda(), boot.cuda(), 0.99, 0.95)
    np.testing.assert_allclose(adv.cpu().numpy(), g["advantage"], rtol=TOL, atol=TOL)
    np.testing.assert_allclose(ret.cpu().numpy(), g["return_"], rtol=TOL, atol=TOL)
    dr = discount_return(reward.cuda(), done.cuda(), boot.cuda(), 0.99)
    np.testing.assert_allclose(dr.cpu().numpy(), g["discount_return"], rtol=TOL, atol=TOL)
    assert np.array_equal(valid_from_done(done.cuda()).cpu().numpy(), g["valid"])


@pytest.mark.parametrize("T,B", [(1, 1), (2, 3), (128, 1024), (20, 4), (257, 130)])
def test_bit_exact_against_oracle(T, B):
    reward, value = R.make_normal(1, T, B), R.make_normal(2, T, B)
    boot, done = R.make_normal(3, B), R.make_random_done(4, T, B, p=0.05)
    adv, ret = generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95)   # CPU in -> CPU out
    assert adv.device.type == "cpu"
    a0, r0 = O.generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95)
    assert torch.equal(adv, a0) and torch.equal(ret, r0)
    assert torch.equal(discount_return(reward, done, boot, 0.99), O.discount_return(reward, done, boot, 0.99))
    assert torch.equal(valid_from_done(done), O.valid_from_done(done))


def test_bootstrap_row_vector_and_dest():
    T, B = 7, 5
    reward, value = R.make_normal(1, T, B), R.make_normal(2, T, B)
    boot, done = R.make_normal(3, 1, B), R.make_random_done(4, T, B) > 0.5      # [1,B], bool done
    adv_d, ret_d = torch.zeros(T, B), torch.zeros(T, B)
    adv, ret = generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95, adv_d, ret_d)
    assert adv is adv_d and ret is ret_d
    a0, _ = O.generalized_advantage_estimation(reward, value, done.float(), boot.reshape(-1), 0.99, 0.95)
    assert torch.equal(adv, a0)


def test_properties_full_size():
    """T=128, B=131072 (the bandwidth-sweep size): lambda=1 GAE equals discounted return minus
    value; done everywhere makes every step one-step; linearity in reward."""
    T, B = 128, 131072
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(0)
    reward = torch.randn(T, B, device=dev, generator=g)
    value = torch.randn(T, B, device=dev, generator=g)
    boot = torch.randn(B, device=dev, generator=g)
    done = (torch.rand(T, B, device=dev, generator=g) < 0.01).float()
    adv1, ret1 = generalized_advantage_estimation(reward, value, done, boot, 0.99, 1.0)
    dr = discount_return(reward, done, boot, 0.99)
    assert float((ret1 - dr).abs().max()) < 1e-3 * float(dr.abs().max())
    ones = torch.ones_like(done)
    adv_d, _ = generalized_advantage_estimation(reward, value, ones, boot, 0.99, 0.95)
    assert torch.equal(adv_d, reward - value)
    a2, _ = generalized_advantage_estimation(2 * reward, 2 * value, done, 2 * boot, 0.99, 0.95)
    a1, _ = generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95)
    assert torch.equal(a2, 2 * a1)          # scaling by 2 is exact in fp32
    v = valid_from_done(done)
    assert bool(((v[1:] <= v[:-1]).all())) and bool((v[0] == 1).all())
