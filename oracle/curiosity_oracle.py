"""CPU oracle for the intrinsic-reward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 torch-CPU / float64 numpy restatement of the reference
algorithms on the hot path of Improbable-AI/curiosity_baselines.  It is the *checker*
for the CUDA product path, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  Nothing under ``curiosity_baselines_b200/`` imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from /root/reference (with mpi4py/gym/matplotlib stubbed), runs it on seeded
inputs and commits the input/output vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.

Weights are always passed as a flat ``dict`` using the reference's ``state_dict`` keys
(e.g. ``encoder.model.0.weight``) so an oracle call and a product call can share one
checkpoint.

Reference lines restated (all under rlpyt/):
  burda_head / universe_head / maze_head ... models/curiosity/encoders.py:7-119
  res_forward ........................... models/curiosity/icm.py:9-43
  icm_* ................................. models/curiosity/icm.py:105-145
  disagreement_* ........................ models/curiosity/disagreement.py:102-167
  RndOracle ............................. models/curiosity/rnd.py:127-212
  RunningMeanStd / RewardForwardFilter .. utils/averages.py:41-89
  ndigo_* ............................... models/curiosity/ndigo.py:114-274
  discount_return / gae / valid_from_done algos/utils.py:8-40,104-112
  process_returns ....................... algos/pg/base.py:53-108
  valid_mean ............................ utils/tensor.py:39-46
  conv2d_head ........................... models/conv2d.py:7-118 (the policy's own encoder, 'none')
  lstm_sequence / policy_forward ........ models/pg/atari_lstm_model.py:114-154
  ppo_loss .............................. algos/pg/ppo.py:192-242, distributions/categorical.py:33-46
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- optional operand rounding
# The product computes with bf16 GEMM / convolution OPERANDS and fp32 accumulation (north_star).  Against this
# fp32 oracle that shows up as (i) ~1e-3 relative noise on outputs and (ii) a few ReLU / LeakyReLU masks that flip
# for pre-activations within that noise of zero -- which no kernel can avoid.  ``operand_rounding(True)`` makes the
# oracle round the two operands of every Linear / convolution to bf16 (straight-through gradient), i.e. the same
# numerical model with fp32 everywhere else, so that a test can separate quantisation from implementation error.
# Default OFF: every parity gate quoted in DESIGN.md is against the plain fp32 oracle unless it says otherwise.
_ROUND = False


class operand_rounding:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _ROUND
        self.prev, _ROUND = _ROUND, self.on
        return self

    def __exit__(self, *exc):
        global _ROUND
        _ROUND = self.prev


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _bf(x):
    return _RoundSTE.apply(x) if _ROUND else x


def _linear(x, w, b=None):
    return F.linear(_bf(x), _bf(w), b)


def _conv2d(x, w, b=None, **kw):
    return F.conv2d(_bf(x), _bf(w), b, **kw)


# --------------------------------------------------------------------------- helpers
def valid_mean(x, valid=None):
    """utils/tensor.py:39-46 -- masked mean, sum(x*valid)/sum(valid)."""
    if valid is None:
        return x.mean()
    valid = valid.to(x.dtype)
    return (x * valid).sum() / valid.sum()


def _sub(p, prefix):
    """View of the weight dict with ``prefix`` stripped."""
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------- encoders
def _batch_norm(h, p, idx, training=True):
    """nn.BatchNorm2d(C) as the encoders place it (encoders.py:23-25, 87-96; eps 1e-5, affine): statistics of this
    batch in training mode (what agent.train_mode() selects for optimize_agent), the running ones otherwise.  The
    running statistics themselves are not updated here (functional restatement; the tests check that update
    against torch.nn.BatchNorm2d directly)."""
    w, b = p[f"model.{idx}.weight"], p[f"model.{idx}.bias"]
    if training:
        return F.batch_norm(h, None, None, w, b, training=True, eps=1e-5)
    return F.batch_norm(h, p[f"model.{idx}.running_mean"], p[f"model.{idx}.running_var"], w, b, training=False, eps=1e-5)


def burda_head(x, p, bn_training=True):
    """encoders.py:68-119.  x: [N,C,84,84] float -> [N,512].

    conv(C->32,k8,s4) lrelu [bn] conv(32->64,k4,s2) lrelu [bn] conv(64->64,k3,s1) lrelu [bn]
    flatten(CHW order) linear(3136->512); no activation after the linear.  With ``batch_norm`` the Sequential
    indices are conv 0/3/6, BatchNorm2d 2/5/8, Linear 10 (recognised by the keys present)."""
    bn = "model.2.running_mean" in p
    step = 3 if bn else 2
    h = x
    for i, stride in enumerate((4, 2, 1)):
        h = F.leaky_relu(_conv2d(h, p[f"model.{step * i}.weight"], p[f"model.{step * i}.bias"], stride=stride))
        if bn:
            h = _batch_norm(h, p, step * i + 2, bn_training)
    h = h.reshape(h.shape[0], -1)
    return _linear(h, p[f"model.{3 * step + 1}.weight"], p[f"model.{3 * step + 1}.bias"])


def universe_head(x, p, bn_training=True):
    """encoders.py:7-35.  5x [conv 3x3 s2 p1 -> 32, ELU, [BatchNorm2d]]; flatten -> [N,288]."""
    bn = "model.2.running_mean" in p
    step = 3 if bn else 2
    h = x
    for layer in range(5):
        h = F.elu(_conv2d(h, p[f"model.{step * layer}.weight"], p[f"model.{step * layer}.bias"],
                           stride=2, padding=1))
        if bn:
            h = _batch_norm(h, p, step * layer + 2, bn_training)
    return h.reshape(h.shape[0], -1)


def maze_head(x, p):
    """encoders.py:37-66.  conv3x3 s1 p1 relu conv3x3 s2 p2 relu flatten linear relu."""
    h = F.relu(_conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=1, padding=1))
    h = F.relu(_conv2d(h, p["model.2.weight"], p["model.2.bias"], stride=2, padding=2))
    h = h.reshape(h.shape[0], -1)
    return F.relu(_linear(h, p["model.5.weight"], p["model.5.bias"]))


def conv2d_head(x, p):
    """models/conv2d.py:67-118 with the policy's defaults (atari_lstm_model.py:94-112): conv(C->16,k8,s4) relu
    conv(16->32,k4,s2,p1) relu flatten(CHW) linear(->512) relu.  Keys conv.conv.{0,2}.*, head.model.0.*."""
    h = F.relu(_conv2d(x, p["conv.conv.0.weight"], p["conv.conv.0.bias"], stride=4))
    h = F.relu(_conv2d(h, p["conv.conv.2.weight"], p["conv.conv.2.bias"], stride=2, padding=1))
    h = h.reshape(h.shape[0], -1)
    return F.relu(_linear(h, p["head.model.0.weight"], p["head.model.0.bias"]))


ENCODERS = {"idf": (universe_head, 288), "idf_burda": (burda_head, 512),
            "idf_maze": (maze_head, 256), "none": (conv2d_head, 512)}


def encode(obs, p, feature_encoding):
    """Shared by icm.py:111-123 / disagreement.py:108-121: cast to float (raw 0..255,
    no scaling -- the obs_stats branch above it is dead code) and run the encoder on
    the [T*B,...] view."""
    T, B = obs.shape[:2]
    fn, _ = ENCODERS[feature_encoding]
    x = obs.to(torch.float32).reshape(T * B, *obs.shape[2:])
    return fn(x, _sub(p, "encoder.")).reshape(T, B, -1)


# --------------------------------------------------------------------------- ICM
def res_forward(phi, action, p):
    """icm.py:9-43.  Ten Linear(F+A -> F); the one-hot action is concatenated at
    every layer.  phi [T,B,F], action [T,B,A] -> [T,B,F]."""
    def lin(x, name):
        return _linear(torch.cat([x, action], 2), p[name + ".weight"], p[name + ".bias"])
    x = F.leaky_relu(lin(phi, "lin_1"))
    for blk in range(1, 5):
        r = F.leaky_relu(lin(x, f"res_block_{blk}.lin_1"))
        r = lin(r, f"res_block_{blk}.lin_2")
        x = r + x
    return lin(x, "lin_last")


def inverse_model(phi1, phi2, p):
    """icm.py:88-92: Linear(2F->F) ReLU Linear(F->A) on [phi1 | phi2]."""
    h = F.relu(_linear(torch.cat([phi1, phi2], 2), p["inverse_model.0.weight"],
                        p["inverse_model.0.bias"]))
    return _linear(h, p["inverse_model.2.weight"], p["inverse_model.2.bias"])


def icm_forward(p, obs, next_obs, action, feature_encoding="idf_burda"):
    """icm.py:105-129 -> (phi1, phi2, predicted_phi2, predicted_action)."""
    phi1 = encode(obs, p, feature_encoding)
    phi2 = encode(next_obs, p, feature_encoding)
    T, B = phi1.shape[:2]
    pred_a = inverse_model(phi1, phi2, p)
    pred_phi2 = res_forward(phi1.detach(), action.reshape(T, B, -1).detach(),
                            _sub(p, "forward_model."))
    return phi1, phi2, pred_phi2, pred_a


def icm_bonus(p, obs, next_obs, action, feature_encoding="idf_burda", prediction_beta=1.0):
    """icm.py:131-134: beta * sum_f (pred_phi2 - phi2)^2 / F  -> [T,B]."""
    with torch.no_grad():
        _, phi2, pred_phi2, _ = icm_forward(p, obs, next_obs, action, feature_encoding)
        return prediction_beta * ((pred_phi2 - phi2) ** 2).sum(-1) / phi2.shape[-1]


def _loss_weights(forward_loss_wt):
    """icm.py:70-75."""
    if forward_loss_wt == -1.0:
        return 1.0, 1.0
    return 1.0 - forward_loss_wt, forward_loss_wt


def icm_loss(p, obs, next_obs, action, valid, feature_encoding="idf_burda",
             forward_loss_wt=0.2):
    """icm.py:136-145 -> (inv_wt*inverse_loss, fwd_wt*forward_loss)."""
    if action.dim() == 2:
        action = action.unsqueeze(1)
    phi1, phi2, pred_phi2, pred_a = icm_forward(p, obs, next_obs, action, feature_encoding)
    T, B = phi1.shape[:2]
    idx = action.reshape(T * B, -1).argmax(1)
    inv = F.cross_entropy(pred_a.reshape(T * B, -1), idx, reduction="none").reshape(T, B)
    fwd = ((pred_phi2 - phi2.detach()) ** 2).sum(-1) / phi2.shape[-1]
    wi, wf = _loss_weights(forward_loss_wt)
    return wi * valid_mean(inv, valid.detach()), wf * valid_mean(fwd, valid.detach())


# --------------------------------------------------------------------------- Disagreement
def disagreement_forward(p, obs, next_obs, action, feature_encoding="idf_burda",
                         n_models=4):
    """disagreement.py:102-134.  Four hard-coded forward models (forward_model_1..4);
    the ``ensemble_size`` ctor argument is ignored by the reference."""
    phi1 = encode(obs, p, feature_encoding)
    phi2 = encode(next_obs, p, feature_encoding)
    T, B = phi1.shape[:2]
    pred_a = inverse_model(phi1, phi2, p)
    a = action.reshape(T, B, -1).detach()
    preds = [res_forward(phi1.detach(), a, _sub(p, f"forward_model_{e}."))
             for e in range(1, n_models + 1)]
    return phi1, phi2, preds, pred_a


def disagreement_bonus(p, obs, next_obs, action, feature_encoding="idf_burda",
                       prediction_beta=1.0, n_models=4):
    """disagreement.py:136-140: unbiased variance over the ensemble, mean over F."""
    with torch.no_grad():
        _, _, preds, _ = disagreement_forward(p, obs, next_obs, action, feature_encoding,
                                              n_models)
        var = torch.var(torch.stack(preds), dim=0)
        return prediction_beta * var.mean(-1)


def disagreement_loss(p, obs, next_obs, action, valid, feature_encoding="idf_burda",
                      forward_loss_wt=0.2, dropout_masks=None, n_models=4):
    """disagreement.py:142-167.  ``dropout_masks``: list of n_models {0,1} tensors
    [T,B,F] replaying F.dropout(p=0.2) (kept elements are scaled by 1/0.8); ``None``
    disables the dropout (mask of ones, scale 1) for deterministic parity."""
    if action.dim() == 2:
        action = action.unsqueeze(1)
    phi1, phi2, preds, pred_a = disagreement_forward(p, obs, next_obs, action,
                                                     feature_encoding, n_models)
    T, B = phi1.shape[:2]
    idx = action.reshape(T * B, -1).argmax(1)
    inv = F.cross_entropy(pred_a.reshape(T * B, -1), idx, reduction="none").reshape(T, B)
    inv = valid_mean(inv, valid)
    fwd = torch.zeros(())
    for e, pred in enumerate(preds):
        err = (pred - phi2.detach()) ** 2
        if dropout_masks is not None:
            err = err * dropout_masks[e] / 0.8
        fwd = fwd + valid_mean(err.sum(-1) / phi2.shape[-1], valid)
    wi, wf = _loss_weights(forward_loss_wt)
    return wi * inv, wf * fwd


# --------------------------------------------------------------------------- running stats
class RunningMeanStd:
    """utils/averages.py:41-68.  float64; mean 0, var 1, count 1e-4 at start."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, b_mean, b_var, b_count):
        delta = b_mean - self.mean
        tot = self.count + b_count
        m2 = (self.var * self.count + b_var * b_count
              + np.square(delta) * self.count * b_count / tot)
        self.mean = self.mean + delta * b_count / tot
        self.var = m2 / tot
        self.count = tot


class RewardForwardFilter:
    """utils/averages.py:70-89.  ``mask`` is a *not-done* mask (rnd.py:185-188); the
    first call seeds every env unmasked."""

    def __init__(self, gamma):
        self.rewems = None
        self.gamma = gamma

    def update(self, rews, mask=None):
        if self.rewems is None:
            self.rewems = rews.copy()
        elif mask is None:
            self.rewems = self.rewems * self.gamma + rews
        else:
            m = mask == 1.0
            self.rewems[m] = self.rewems[m] * self.gamma + rews[m]
        return self.rewems.copy()


# --------------------------------------------------------------------------- RND
def _rnd_net(x, p, predictor):
    """rnd.py:54-117: Burda-shaped stack on one channel; the predictor adds
    ReLU Linear ReLU Linear after the 3136->512 layer."""
    h = F.leaky_relu(_conv2d(x, p["0.weight"], p["0.bias"], stride=4))
    h = F.leaky_relu(_conv2d(h, p["2.weight"], p["2.bias"], stride=2))
    h = F.leaky_relu(_conv2d(h, p["4.weight"], p["4.bias"], stride=1))
    h = _linear(h.reshape(h.shape[0], -1), p["7.weight"], p["7.bias"])
    if predictor:
        h = _linear(F.relu(h), p["9.weight"], p["9.bias"])
        h = _linear(F.relu(h), p["11.weight"], p["11.bias"])
    return h


class RndOracle:
    """rnd.py:15-212.  Holds the three pieces of running state (obs_rms, rew_rms,
    rew_rff); weights come in as a dict with keys forward_model.* / target_model.*."""

    feature_size = 512

    def __init__(self, image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=1.0,
                 gamma=0.99):
        self.prediction_beta = prediction_beta
        self.drop_probability = drop_probability
        self.obs_rms = RunningMeanStd(shape=(1, 1, image_shape[1], image_shape[2]))
        self.rew_rms = RunningMeanStd()
        self.rew_rff = RewardForwardFilter(gamma)

    def forward(self, p, obs, done=None):
        """rnd.py:127-179.  Only the newest frame of the stack is used.  With ``done``
        the first n_b = sum_t(1-done[t,b]) frames of env b update obs_rms."""
        obs = obs[:, :, -1:, :, :]
        T, B = obs.shape[:2]
        if done is not None:
            n_valid = np.abs(done.numpy() - 1).sum(0)
            frames = obs.numpy()
            rows = [frames[: int(n_valid[b]), b] for b in range(B)]
            self.obs_rms.update(np.concatenate(rows))
        mean = torch.from_numpy(self.obs_rms.mean).float()
        var = torch.from_numpy(self.obs_rms.var).float()
        x = torch.clamp((obs - mean) / torch.sqrt(var), -5, 5).float()
        x = x.reshape(T * B, *x.shape[2:])
        phi = _rnd_net(x, _sub(p, "target_model."), False).reshape(T, B, -1)
        pred = _rnd_net(x, _sub(p, "forward_model."), True).reshape(T, B, -1)
        return phi, pred

    def compute_bonus(self, p, next_obs, done):
        """rnd.py:181-203."""
        with torch.no_grad():
            phi, pred = self.forward(p, next_obs, done=done)
            err = ((pred - phi) ** 2).sum(-1) / self.feature_size
            err_np = err.numpy().copy()
            nd = torch.abs(done - 1).numpy()
            T = err_np.shape[0]
            tot = np.array([self.rew_rff.update(err_np[t], nd[t]) for t in range(T)])
            mean_len = np.mean(nd.sum(0))
            self.rew_rms.update_from_moments(np.mean(tot), np.var(tot), mean_len)
            rew_var = torch.from_numpy(np.array(self.rew_rms.var)).float()
            r = err / torch.sqrt(rew_var) * torch.from_numpy(nd).float()
            return self.prediction_beta * r

    def compute_loss(self, p, obs, valid, mask=None):
        """rnd.py:205-212.  ``mask`` replays (rand(T,B) > drop_probability)."""
        phi, pred = self.forward(p, obs, done=None)
        err = ((pred - phi.detach()) ** 2).sum(-1) / self.feature_size
        if mask is None:
            mask = (torch.rand(err.shape) > self.drop_probability).float()
        return valid_mean(err * mask.detach(), valid.detach())


# --------------------------------------------------------------------------- NDIGO
def gru_sequence(x, p, hidden):
    """torch.nn.GRU single layer, zero initial state (ndigo.py:79,127).  x [T,B,I]."""
    T, B = x.shape[:2]
    w_ih, w_hh, b_ih, b_hh = (p["gru.weight_ih_l0"], p["gru.weight_hh_l0"],
                              p["gru.bias_ih_l0"], p["gru.bias_hh_l0"])
    h = torch.zeros(B, hidden)
    outs = []
    for t in range(T):
        gi = _linear(x[t], w_ih, b_ih)
        gh = _linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs)


def _action_windows(actions, k):
    """ndigo.py:154-163,236-240: row i is concat(actions[i], ..., actions[i+k-1])."""
    T = actions.shape[0]
    return torch.cat([actions[j:T - k + j] for j in range(k)], dim=2)


def _ndigo_predict(bel, seq, p, k):
    q = _sub(p, f"forward_model_{k}.model.")
    h = F.relu(_linear(torch.cat([bel, seq], 2), q["0.weight"], q["0.bias"]))
    return _linear(h, q["2.weight"], q["2.bias"])


def _bce(logits, target):
    """binary_cross_entropy(sigmoid(logits), target) with torch's log clamp at -100."""
    prob = torch.sigmoid(logits)
    return F.binary_cross_entropy(prob, target, reduction="none")


def ndigo_beliefs(p, obs, prev_actions, feature_encoding="idf_maze", gru_size=128):
    T, B = obs.shape[:2]
    fn, _ = ENCODERS[feature_encoding]
    enc = fn(obs.float().reshape(T * B, *obs.shape[2:]), _sub(p, "encoder.")).reshape(T, B, -1)
    return gru_sequence(torch.cat([enc, prev_actions], 2), p, gru_size)


def ndigo_bonus(p, obs, prev_actions, actions, horizon, feature_encoding="idf_maze",
                gru_size=128):
    """ndigo.py:132-217.  r[H:T-1] = L_{t-1} - L_t[1:], zero elsewhere."""
    with torch.no_grad():
        T, B = obs.shape[:2]
        H = horizon
        bel = ndigo_beliefs(p, obs, prev_actions, feature_encoding, gru_size)
        flat = obs.float().reshape(T, B, -1)
        logit_t = _ndigo_predict(bel[:T - H], _action_windows(actions, H), p, H)
        logit_tm1 = _ndigo_predict(bel[:T - H - 1], _action_windows(actions, H + 1), p, H + 1)
        loss_t = _bce(logit_t, flat[H:]).mean(-1)
        loss_tm1 = _bce(logit_tm1, flat[H:-1]).mean(-1)
        r = torch.zeros(T, B)
        r[H:T - 1] = loss_tm1 - loss_t[1:]
        return r


def ndigo_loss(p, obs, prev_actions, actions, valid=None, num_predictors=10,
               feature_encoding="idf_maze", gru_size=128):
    """ndigo.py:220-274.  Sum over k of mean BCE; ``valid`` is ignored."""
    if prev_actions.dim() == 2:
        prev_actions = prev_actions.unsqueeze(1)
    if actions.dim() == 2:
        actions = actions.unsqueeze(1)
    T, B = obs.shape[:2]
    bel = ndigo_beliefs(p, obs, prev_actions, feature_encoding, gru_size)
    flat = obs.float().reshape(T, B, -1)
    loss = 0.0
    for k in range(1, num_predictors + 1):
        logits = _ndigo_predict(bel[:T - k], _action_windows(actions, k).detach(), p, k)
        loss = loss + _bce(logits, flat[k:].detach()).mean()
    return loss


# --------------------------------------------------------------------------- returns
def discount_return(reward, done, bootstrap_value, discount):
    """algos/utils.py:8-21."""
    nd = 1 - done.to(reward.dtype)
    ret = torch.zeros_like(reward)
    ret[-1] = reward[-1] + discount * bootstrap_value * nd[-1]
    for t in range(reward.shape[0] - 2, -1, -1):
        ret[t] = reward[t] + ret[t + 1] * discount * nd[t]
    return ret


def generalized_advantage_estimation(reward, value, done, bootstrap_value, discount,
                                     gae_lambda):
    """algos/utils.py:24-40 -> (advantage, return_)."""
    nd = 1 - done.to(reward.dtype)
    adv = torch.zeros_like(reward)
    adv[-1] = reward[-1] + discount * bootstrap_value * nd[-1] - value[-1]
    for t in range(reward.shape[0] - 2, -1, -1):
        delta = reward[t] + discount * value[t + 1] * nd[t] - value[t]
        adv[t] = delta + discount * gae_lambda * nd[t] * adv[t + 1]
    return adv, adv + value


def valid_from_done(done):
    """algos/utils.py:104-112."""
    done = done.to(torch.float32)
    valid = torch.ones_like(done)
    valid[1:] = 1 - torch.clamp(torch.cumsum(done[:-1], dim=0), max=1)
    return valid


class ReturnsOracle:
    """algos/pg/base.py:53-108 after the bonus has been obtained: reward += r_int
    (in place), optional reward normalisation, GAE / discounted return, valid mask,
    optional advantage normalisation."""

    def __init__(self, discount=0.99, gae_lambda=0.95, normalize_reward=False,
                 normalize_advantage=False, recurrent=True, mid_batch_reset=False):
        self.discount, self.gae_lambda = discount, gae_lambda
        self.normalize_reward, self.normalize_advantage = normalize_reward, normalize_advantage
        self.use_valid = (not mid_batch_reset) or recurrent
        self.reward_ff = RewardForwardFilter(discount)
        self.reward_rms = RunningMeanStd()

    def __call__(self, reward, done, value, bootstrap_value, r_int=None):
        done = done.to(reward.dtype)
        if r_int is not None:
            reward += r_int
        if self.normalize_reward:
            rews = np.concatenate([self.reward_ff.update(r)
                                   for r in reward.clone().numpy()]).astype(np.float64)
            self.reward_rms.update_from_moments(np.mean(rews), np.var(rews), len(rews))
            reward = reward / math.sqrt(self.reward_rms.var)
        if self.gae_lambda == 1:
            ret = discount_return(reward, done, bootstrap_value, self.discount)
            adv = ret - value
        else:
            adv, ret = generalized_advantage_estimation(reward, value, done, bootstrap_value,
                                                        self.discount, self.gae_lambda)
        valid = valid_from_done(done) if self.use_valid else None
        if self.normalize_advantage:
            sel = adv[valid > 0] if valid is not None else adv
            adv = (adv - sel.mean()) / max(sel.std(), 1e-6)
        return ret, adv, valid


# --------------------------------------------------------------------------- policy net + PPO loss
def lstm_sequence(x, p, h0, c0):
    """torch.nn.LSTM, single layer (atari_lstm_model.py:112,142), gate order i,f,g,o.
    x [T,B,I]; h0, c0 [B,H] -> (outputs [T,B,H], h_T, c_T)."""
    w_ih, w_hh, b_ih, b_hh = (p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"],
                              p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"])
    h, c = h0, c0
    outs = []
    for t in range(x.shape[0]):
        a = _linear(x[t], w_ih, b_ih) + _linear(h, w_hh, b_hh)
        a_i, a_f, a_g, a_o = a.chunk(4, 1)
        c = torch.sigmoid(a_f) * c + torch.sigmoid(a_i) * torch.tanh(a_g)
        h = torch.sigmoid(a_o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs), h, c


def policy_forward(p, image, prev_action, init_rnn_state=None, feature_encoding="idf_burda"):
    """AtariLstmModel.forward (atari_lstm_model.py:117-154) for the curiosity configurations whose
    policy encoder is one of the curiosity heads (``self.conv = BurdaHead/UniverseHead/MazeHead``,
    :80-94): raw float pixels (no /255, :130-133) -> conv -> cat one-hot prev_action -> LSTM ->
    softmax(pi), value.  image [T,B,C,H,W]; prev_action one-hot float [T,B,A];
    init_rnn_state (h, c) each [1,B,H] or None.  Returns pi [T,B,A], v [T,B], (h_n, c_n) [1,B,H]."""
    T, B = image.shape[:2]
    fn, _ = ENCODERS[feature_encoding]
    x = image.to(torch.float32).reshape(T * B, *image.shape[2:])
    fc_out = fn(x, _sub(p, "conv."))
    lstm_in = torch.cat([fc_out.view(T, B, -1), prev_action.view(T, B, -1)], dim=2)
    H = p["lstm.weight_hh_l0"].shape[1]
    if init_rnn_state is None:
        h0, c0 = torch.zeros(B, H), torch.zeros(B, H)
    else:
        h0, c0 = init_rnn_state[0][0], init_rnn_state[1][0]
    out, hn, cn = lstm_sequence(lstm_in, p, h0, c0)
    flat = out.view(T * B, -1)
    pi = F.softmax(_linear(flat, p["pi.weight"], p["pi.bias"]), dim=-1)
    v = _linear(flat, p["value.weight"], p["value.bias"]).squeeze(-1)
    return pi.view(T, B, -1), v.view(T, B), (hn.unsqueeze(0), cn.unsqueeze(0))


def ppo_loss(pi, value, action, old_prob, return_, advantage, valid, ratio_clip=0.1,
             value_loss_coeff=1.0, entropy_loss_coeff=0.01):
    """PPO.loss without the curiosity terms (ppo.py:209-223) on the Categorical distribution
    (categorical.py:33-46, EPS=1e-8).  action int64 [T,B].  Returns
    (loss, pi_loss, value_loss, entropy_loss, entropy, perplexity)."""
    eps = 1e-8
    num = pi.gather(-1, action.unsqueeze(-1)).squeeze(-1)
    den = old_prob.gather(-1, action.unsqueeze(-1)).squeeze(-1)
    ratio = (num + eps) / (den + eps)
    surr_1 = ratio * advantage
    surr_2 = torch.clamp(ratio, 1.0 - ratio_clip, 1.0 + ratio_clip) * advantage
    pi_loss = -valid_mean(torch.min(surr_1, surr_2), valid)
    value_loss = value_loss_coeff * valid_mean(0.5 * (value - return_) ** 2, valid)
    ent = -torch.sum(pi * torch.log(pi + eps), dim=-1)
    entropy = valid_mean(ent, valid)
    perplexity = valid_mean(torch.exp(ent), valid)
    entropy_loss = -entropy_loss_coeff * entropy
    return pi_loss + value_loss + entropy_loss, pi_loss, value_loss, entropy_loss, entropy, perplexity
