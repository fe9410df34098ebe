"""Seeded recipes shared by the golden-vector generator (runs against /root/reference in
the build container) and by the tests (run anywhere).  Everything here is a pure
function of its integer seed on the CPU generator of this image's torch build, so the
big tensors (frames, weights) never need to be committed -- only seeds and outputs are.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


# ---------------------------------------------------------------- parameter shapes
def burda_shapes(prefix, c_in, out=512):
    return [(prefix + "0.weight", (32, c_in, 8, 8)), (prefix + "0.bias", (32,)),
            (prefix + "2.weight", (64, 32, 4, 4)), (prefix + "2.bias", (64,)),
            (prefix + "4.weight", (64, 64, 3, 3)), (prefix + "4.bias", (64,)),
            (prefix + "7.weight", (out, 3136)), (prefix + "7.bias", (out,))]


def universe_shapes(prefix, c_in):
    out = []
    for layer in range(5):
        out += [(prefix + f"{2 * layer}.weight", (32, c_in if layer == 0 else 32, 3, 3)),
                (prefix + f"{2 * layer}.bias", (32,))]
    return out


def maze_shapes(prefix, c_in):
    return [(prefix + "0.weight", (16, c_in, 3, 3)), (prefix + "0.bias", (16,)),
            (prefix + "2.weight", (16, 16, 3, 3)), (prefix + "2.bias", (16,)),
            (prefix + "5.weight", (256, 256)), (prefix + "5.bias", (256,))]


def encoder_shapes(feature_encoding, c_in):
    if feature_encoding == "idf_burda":
        return burda_shapes("encoder.model.", c_in), 512
    if feature_encoding == "idf":
        return universe_shapes("encoder.model.", c_in), 288
    if feature_encoding == "idf_maze":
        return maze_shapes("encoder.model.", c_in), 256
    raise ValueError(feature_encoding)


def res_forward_shapes(prefix, feat, act):
    names = ["lin_1"] + [f"res_block_{b}.lin_{i}" for b in range(1, 5) for i in (1, 2)] + ["lin_last"]
    out = []
    for n in names:
        out += [(prefix + n + ".weight", (feat, feat + act)), (prefix + n + ".bias", (feat,))]
    return out


def icm_shapes(action_size, feature_encoding="idf_burda", c_in=4):
    enc, feat = encoder_shapes(feature_encoding, c_in)
    inv = [("inverse_model.0.weight", (feat, 2 * feat)), ("inverse_model.0.bias", (feat,)),
           ("inverse_model.2.weight", (action_size, feat)), ("inverse_model.2.bias", (action_size,))]
    return enc + inv + res_forward_shapes("forward_model.", feat, action_size)


def disagreement_shapes(action_size, feature_encoding="idf_burda", c_in=4, n_models=4):
    enc, feat = encoder_shapes(feature_encoding, c_in)
    inv = [("inverse_model.0.weight", (feat, 2 * feat)), ("inverse_model.0.bias", (feat,)),
           ("inverse_model.2.weight", (action_size, feat)), ("inverse_model.2.bias", (action_size,))]
    out = enc + inv
    for e in range(1, n_models + 1):
        out += res_forward_shapes(f"forward_model_{e}.", feat, action_size)
    return out


def rnd_shapes():
    pred = burda_shapes("forward_model.", 1) + [
        ("forward_model.9.weight", (512, 512)), ("forward_model.9.bias", (512,)),
        ("forward_model.11.weight", (512, 512)), ("forward_model.11.bias", (512,))]
    return pred + burda_shapes("target_model.", 1)


def ndigo_shapes(action_size, c_in=3, hw=5, gru=128, num_predictors=10):
    enc, feat = encoder_shapes("idf_maze", c_in)
    out = enc + [("gru.weight_ih_l0", (3 * gru, feat + action_size)),
                 ("gru.weight_hh_l0", (3 * gru, gru)),
                 ("gru.bias_ih_l0", (3 * gru,)), ("gru.bias_hh_l0", (3 * gru,))]
    for k in range(1, num_predictors + 1):
        out += [(f"forward_model_{k}.model.0.weight", (64, gru + k * action_size)),
                (f"forward_model_{k}.model.0.bias", (64,)),
                (f"forward_model_{k}.model.2.weight", (c_in * hw * hw, 64)),
                (f"forward_model_{k}.model.2.bias", (c_in * hw * hw,))]
    return out


def conv2d_head_shapes(prefix, c_in, fc=512):
    """Conv2dHeadModel with the policy's defaults (models/conv2d.py:67-118, atari_lstm_model.py:94-112)."""
    return [(prefix + "conv.conv.0.weight", (16, c_in, 8, 8)), (prefix + "conv.conv.0.bias", (16,)),
            (prefix + "conv.conv.2.weight", (32, 16, 4, 4)), (prefix + "conv.conv.2.bias", (32,)),
            (prefix + "head.model.0.weight", (fc, 3200)), (prefix + "head.model.0.bias", (fc,))]


def policy_shapes(action_size, feature_encoding="idf_burda", c_in=4, lstm=512):
    """AtariLstmModel minus its curiosity_model (atari_lstm_model.py:80-114)."""
    if feature_encoding == "none":
        enc, feat = conv2d_head_shapes("encoder.", c_in), 512
    else:
        enc, feat = encoder_shapes(feature_encoding, c_in)
    enc = [(k.replace("encoder.", "conv.", 1), s) for k, s in enc]
    return enc + [("lstm.weight_ih_l0", (4 * lstm, feat + action_size)), ("lstm.weight_hh_l0", (4 * lstm, lstm)),
                  ("lstm.bias_ih_l0", (4 * lstm,)), ("lstm.bias_hh_l0", (4 * lstm,)),
                  ("pi.weight", (action_size, lstm)), ("pi.bias", (action_size,)),
                  ("value.weight", (1, lstm)), ("value.bias", (1,))]


def make_policy_params(action_size, seed, feature_encoding="idf_burda", conv_gain=0.65, c_in=4):
    """Policy weights whose encoder is damped (gain < 1 per layer) so that raw 0..255 pixels give O(1)
    features: the LSTM gates then work in their non-saturated range and parity is informative."""
    shapes = policy_shapes(action_size, feature_encoding, c_in=c_in)
    conv = [s for s in shapes if s[0].startswith("conv.")]
    rest = [s for s in shapes if not s[0].startswith("conv.")]
    out = make_params(conv, seed, gain=conv_gain, bias_scale=conv_gain)
    out.update(make_params(rest, seed + 1, gain=1.0))
    return out


def make_probs(seed, T, B, A):
    return torch.softmax(make_normal(seed, T, B, A), dim=-1)


def make_params(shapes, seed, gain=1.0, bias_scale=1.0):
    """U(-b, b) with b = gain / sqrt(fan_in) (torch's default Linear/Conv bound).
    Biases get the same bound so they are exercised (RND's own init zeroes them)."""
    g = _gen(seed)
    out = {}
    fan_in = 1
    for name, shape in shapes:
        if name.endswith("weight") or "weight_" in name:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = gain / math.sqrt(fan_in)
        else:
            bound = bias_scale / math.sqrt(fan_in)
        out[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


# ---------------------------------------------------------------- inputs
def make_frames(seed, T, B, C=4, H=84, W=84):
    """uint8 frame stacks [T,B,C,H,W]: a smooth per-env pattern plus noise so the
    per-pixel statistics are non-degenerate."""
    g = _gen(seed)
    base = torch.randint(0, 200, (1, B, C, H, W), generator=g)
    noise = torch.randint(0, 56, (T, B, C, H, W), generator=g)
    return (base + noise).to(torch.uint8)


def make_binary_frames(seed, T, B, C=3, H=5, W=5, p=0.3):
    g = _gen(seed)
    return (torch.rand((T, B, C, H, W), generator=g) < p).float()


def make_actions(seed, T, B, A):
    g = _gen(seed)
    return torch.randint(0, A, (T, B), generator=g)


def onehot(idx, A):
    return torch.nn.functional.one_hot(idx, A).float()


def make_suffix_done(seed, T, B, frac=0.4):
    """done[t,b]=1 for t >= first_done[b] for about ``frac`` of the envs (the layout the
    WaitReset collectors produce: once done, done stays 1 to the end of the batch)."""
    g = _gen(seed)
    first = torch.randint(1, T + 1, (B,), generator=g)
    has = torch.rand((B,), generator=g) < frac
    first = torch.where(has, first, torch.full_like(first, T))
    t = torch.arange(T).unsqueeze(1)
    return (t >= first.unsqueeze(0)).float()


def make_random_done(seed, T, B, p=0.1):
    g = _gen(seed)
    return (torch.rand((T, B), generator=g) < p).float()


def make_normal(seed, *shape):
    return torch.randn(shape, generator=_gen(seed))


def make_mask(seed, shape, keep_above):
    """{0,1} mask = (u > keep_above), the form of rnd.py:208-209 / F.dropout."""
    return (torch.rand(shape, generator=_gen(seed)) > keep_above).float()
