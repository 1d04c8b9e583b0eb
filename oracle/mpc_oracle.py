"""CPU oracle for the MPC planning hot path of iclavera/learning_to_adapt.

TEST INFRASTRUCTURE ONLY.  Nothing under ``learning_to_adapt_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline, never as the product path.

It is a numpy *restatement* (not a copy) of the reference's algorithm in the reference's exact
dtype choreography: planner state and returns in float64, dense MLP in float32.
All ``file:line`` citations are relative to the upstream tree (``/root/reference``).

Pinning status (see DESIGN.md "Oracle"):
  * planner loops (random shooting, CEM)  -- PINNED: ``tests/golden/make_golden.py`` runs the
    verbatim reference ``MPCController`` (policies/mpc_controller.py) in the authoring container and the
    golden fixtures hold its outputs; ``tests/test_oracle_golden.py`` checks this oracle against them.
  * env ``reward`` closed forms            -- PINNED: fixtures come from the reference's own ``reward``
    method bodies, executed from the upstream source text.
  * ``predict`` / ``adapt`` host choreography (normalise, pad, task split, denormalise, delta add)
                                           -- PINNED: fixtures come from the reference's own
    ``MLPDynamicsModel.predict`` / ``MetaMLPDynamicsModel.{predict,_predict,_pad_inputs,adapt}`` bodies
    run with a stub in place of the TF session.
  * the TF1 graph itself (``tf.layers.dense`` forward, ``tf.gradients``)  -- PARITY UNPINNED:
    TensorFlow 1.13.1 is not installable here (no wheels for py3.12, no network); the dense-layer math
    ``act(x @ W + b)`` and the MSE gradient are restated from the published semantics and cross-checked
    against torch autograd in ``tests/test_oracle_golden.py``.
  * "ensemble = E" (BASELINE.json)         -- NO REFERENCE CODE exists; the aggregation
    rule (mean of the E predicted deltas) is this build's definition.
  * device candidate sampler (Philox4x32-10, throughput mode only; the reference draws with numpy's MT19937,
    which the parity mode keeps)             -- PINNED to Random123's published known-answer vectors.
  * numpy's legacy stream itself (MT19937 refill, ``uniform``, polar ``normal``; third-party: numpy, pinned 1.15.1 upstream)
                                           -- ``oracle/mt19937_oracle.py``, PINNED against numpy bit for bit
    (``tests/test_mt19937_oracle_cpu.py``); on the GPU the device generator is compared with numpy directly.
"""
from collections import OrderedDict

import numpy as np

EPS = 1e-10  # mlp_dynamics.py:265-270

REWARD_HALF_CHEETAH = 0
REWARD_ANT = 1
REWARD_ARM = 2
REWARD_KINDS = {"half_cheetah": REWARD_HALF_CHEETAH, "ant": REWARD_ANT, "arm_7dof": REWARD_ARM}


# --------------------------------------------------------------------------------------------------
# parameters / normalisation
# --------------------------------------------------------------------------------------------------
def param_keys(n_hidden):
    """Key order of the reference's param OrderedDict (core/utils.py:241-296, layers.py:142-171)."""
    keys = []
    for i in range(n_hidden):
        keys += ["hidden_%d/kernel" % i, "hidden_%d/bias" % i]
    keys += ["output/kernel", "output/bias"]
    return keys


def xavier_params(rng, in_dim, hidden_sizes, out_dim, out_scale=1.0):
    """Xavier-uniform kernels [in, out] fp32, zero biases (core/utils.py:81-82 defaults)."""
    sizes = [in_dim] + list(hidden_sizes) + [out_dim]
    params = OrderedDict()
    keys = param_keys(len(hidden_sizes))
    for l in range(len(sizes) - 1):
        lim = np.sqrt(6.0 / (sizes[l] + sizes[l + 1]))
        w = rng.uniform(-lim, lim, size=(sizes[l], sizes[l + 1])).astype(np.float32)
        if l == len(sizes) - 2:
            w = (w * np.float32(out_scale)).astype(np.float32)
        params[keys[2 * l]] = w
        params[keys[2 * l + 1]] = np.zeros(sizes[l + 1], np.float32)
    return params


def make_normalization(rng, obs_dim, act_dim, act_low, act_high):
    """Synthetic population stats in float64 (SURVEY.md 8(d) recipe)."""
    norm = OrderedDict()
    norm["obs"] = (rng.normal(0.0, 1.0, obs_dim), rng.uniform(0.5, 2.0, obs_dim))
    norm["delta"] = (rng.normal(0.0, 0.01, obs_dim), rng.uniform(0.05, 0.2, obs_dim))
    act_sigma = (np.asarray(act_high, np.float64) - np.asarray(act_low, np.float64)) / np.sqrt(12.0)
    norm["act"] = (rng.normal(0.0, 0.1, act_dim), act_sigma * np.ones(act_dim))
    return norm


def normalize(x, mean, std):
    return (x - mean) / (std + EPS)  # mlp_dynamics.py:265-266


def denormalize(x, mean, std):
    return x * (std + EPS) + mean  # mlp_dynamics.py:269-270


# --------------------------------------------------------------------------------------------------
# dense stack (tf.layers.dense semantics: act(x @ W + b), fp32)   core/utils.py:119-140, 241-296
# --------------------------------------------------------------------------------------------------
def mlp_forward(x32, params, keep_activations=False):
    """x32: [n, in] float32.  Hidden ReLU, linear output."""
    keys = list(params.keys())
    n_layers = len(keys) // 2
    h = np.ascontiguousarray(x32, dtype=np.float32)
    acts = [h]
    for l in range(n_layers):
        w = params[keys[2 * l]]
        b = params[keys[2 * l + 1]]
        h = h @ w + b
        if l < n_layers - 1:
            h = np.maximum(h, np.float32(0.0))
        if keep_activations:
            acts.append(h)
    return (h, acts) if keep_activations else h


def f_delta_pred(obs_n, act_n, params):
    """The TF placeholder feed casts float64 -> float32 (mlp_dynamics.py:63-68)."""
    x = np.concatenate([obs_n, act_n], axis=1).astype(np.float32)
    return mlp_forward(x, params)


# --------------------------------------------------------------------------------------------------
# one-step predict      mlp_dynamics.py:204-222 / meta_mlp_dynamics.py:276-306
# --------------------------------------------------------------------------------------------------
def predict(obs, act, params, norm):
    obs = np.asarray(obs, np.float64)
    act = np.asarray(act, np.float64)
    obs_n = normalize(obs, *norm["obs"])
    act_n = normalize(act, *norm["act"])
    delta = np.array(f_delta_pred(obs_n, act_n, params))  # float32
    delta = denormalize(delta, *norm["delta"])            # float32 * float64 -> float64
    return obs + delta


def predict_per_task(obs, act, param_sets, norm):
    """GrBAL post-adapt predict: row chunk k (equal chunks) uses weight set k
    (meta_mlp_dynamics.py:296-306 + inference graph 143-163).  Zero-padded tasks are not computed."""
    obs = np.asarray(obs, np.float64)
    act = np.asarray(act, np.float64)
    k = len(param_sets)
    assert obs.shape[0] % k == 0
    obs_n = normalize(obs, *norm["obs"])
    act_n = normalize(act, *norm["act"])
    chunks_o = np.split(obs_n, k, axis=0)
    chunks_a = np.split(act_n, k, axis=0)
    delta = np.concatenate([f_delta_pred(o, a, p) for o, a, p in zip(chunks_o, chunks_a, param_sets)], axis=0)
    delta = denormalize(delta, *norm["delta"])
    return obs + delta


def predict_ensemble_mean(obs, act, param_sets, norm):
    """'ensemble=E' reading (ii): every row goes through all E sets; the E denormalised deltas are
    averaged in float64, set order 0..E-1.  No reference code (see module docstring)."""
    obs = np.asarray(obs, np.float64)
    act = np.asarray(act, np.float64)
    obs_n = normalize(obs, *norm["obs"])
    act_n = normalize(act, *norm["act"])
    acc = np.zeros_like(obs)
    for p in param_sets:
        acc = acc + denormalize(np.array(f_delta_pred(obs_n, act_n, p)), *norm["delta"])
    return obs + acc / float(len(param_sets))


# --------------------------------------------------------------------------------------------------
# env reward closed forms
# --------------------------------------------------------------------------------------------------
def reward_fn(kind, dt):
    if kind == REWARD_HALF_CHEETAH:   # envs/half_cheetah_env.py:58-65
        def r(obs, action, next_obs):
            ctrl_cost = 1e-1 * 0.5 * np.sum(np.square(action), axis=1)
            return (next_obs[:, -3] - obs[:, -3]) / dt - ctrl_cost
    elif kind == REWARD_ANT:          # envs/ant_env.py:56-66
        def r(obs, action, next_obs):
            return (next_obs[:, -3] - obs[:, -3]) / dt - 0 + 0.05
    elif kind == REWARD_ARM:          # envs/arm_7dof_env.py:91-99
        def r(obs, action, next_obs):
            reward_dist = -np.linalg.norm(next_obs[:, -3:], axis=1)
            reward_ctrl = -np.sum(np.square(action), axis=1)
            return reward_dist + 0.01 * 0.5 * reward_ctrl
    else:
        raise ValueError(kind)
    return r


# --------------------------------------------------------------------------------------------------
# planner: random shooting       policies/mpc_controller.py:108-129
# --------------------------------------------------------------------------------------------------
def _step_fn(param_sets, norm, mode):
    if mode == "shared":
        return lambda o, a: predict(o, a, param_sets[0], norm)
    if mode == "per_env":
        return lambda o, a: predict_per_task(o, a, param_sets, norm)
    if mode == "ensemble":
        return lambda o, a: predict_ensemble_mean(o, a, param_sets, norm)
    raise ValueError(mode)


def rollout_returns(observations, actions, param_sets, norm, reward_kind, dt, discount=1.0, mode="shared"):
    """actions: [H, n*m, A] (row r belongs to env r // n).  Returns float64 [m, n]."""
    observations = np.asarray(observations, np.float64)
    actions = np.asarray(actions, np.float64)
    h, rows, _ = actions.shape
    m = observations.shape[0]
    n = rows // m
    assert n * m == rows
    step = _step_fn(param_sets, norm, mode)
    rew = reward_fn(reward_kind, dt)
    returns = np.zeros((rows,))
    observation = np.repeat(observations, n, axis=0)      # :119
    for t in range(h):                                    # :116
        next_observation = step(observation, actions[t])  # :120
        rewards = rew(observation, actions[t], next_observation)  # :125
        returns += discount ** t * rewards                # :126
        observation = next_observation                    # :127
    return returns.reshape(m, n)                          # :128


def rs_plan(observations, actions, param_sets, norm, reward_kind, dt, discount=1.0, mode="shared"):
    """Returns (chosen_actions [m, A] f64, best_idx [m], returns [m, n])."""
    returns = rollout_returns(observations, actions, param_sets, norm, reward_kind, dt, discount, mode)
    m, n = returns.shape
    cand_a = np.asarray(actions, np.float64)[0].reshape(m, n, -1)   # :118
    best = np.argmax(returns, axis=1)                                # :129 (first max wins ties)
    return cand_a[range(m), best], best, returns


# --------------------------------------------------------------------------------------------------
# planner: CEM, bug-compatible       policies/mpc_controller.py:71-106   
# --------------------------------------------------------------------------------------------------
def cem_plan(observations, z_per_iter, act_low, act_high, param_sets, norm, reward_kind, dt, horizon,
             percent_elites=0.1, alpha=0.1, discount=1.0, mode="shared", corrected=False):
    """z_per_iter: list of standard-normal draws, each [n, m, H*A] (mpc_controller.py:85).
    ``corrected=False`` reproduces the reference including the rank-mask defect (:101) and the use of
    unclipped samples in the rollout (:88-89); ``corrected=True`` selects the true top-k as elites.
    Returns (chosen_actions [m,A], best_idx [m], returns [m,n] of the last iteration, mean, std)."""
    observations = np.asarray(observations, np.float64)
    m = observations.shape[0]
    n = z_per_iter[0].shape[0]
    h = horizon
    act_dim = len(act_low)
    num_elites = max(int(n * percent_elites), 1)                 # :78
    mean = np.zeros((m, h * act_dim))                            # :79
    std = np.ones((m, h * act_dim))                              # :80
    clip_low = np.concatenate([np.asarray(act_low, np.float64)] * h)   # :81
    clip_high = np.concatenate([np.asarray(act_high, np.float64)] * h)  # :82
    returns = None
    cand_a = None
    for z in z_per_iter:                                         # :84
        a = mean + z * std                                       # :86
        a_stacked = np.clip(a, clip_low, clip_high)              # :87
        a = a.reshape((n * m, h, act_dim))                       # :88
        a = np.transpose(a, (1, 0, 2))                           # :89  -> [H, n*m, A], UNclipped
        cand_a = a[0].reshape((m, n, -1))                        # :94
        returns = rollout_returns(observations, a, param_sets, norm, reward_kind, dt, discount, mode)
        if corrected:
            ranks = (-returns).argsort(axis=-1).argsort(axis=-1)
            elites_idx = (ranks < num_elites).T
        else:
            elites_idx = ((-returns).argsort(axis=-1) < num_elites).T   # :101
        elites = a_stacked[elites_idx]                           # :102
        mean = mean * alpha + (1 - alpha) * np.mean(elites, axis=0)     # :103
        std = np.std(elites, axis=0)                             # :104
    best = np.argmax(returns, axis=1)
    return cand_a[range(m), best], best, returns, mean, std      # :106


# --------------------------------------------------------------------------------------------------
# GrBAL adapt: one SGD step on the M-row context, from the prior, per task
#   meta_mlp_dynamics.py:321-345 (host), 96-120 (graph), 409-421 (_adapt_sym)
# --------------------------------------------------------------------------------------------------
def adapt_one_task(x32, target32, params, inner_lr):
    """x32 [M, D+A] normalised input, target32 [M, D] normalised delta, both float32.
    loss = mean over M*D of (target - f(x))^2 (:118); theta' = theta - lr * dloss/dtheta (:409-421).
    Manual fp32 backprop (what tf.gradients computes for a dense/ReLU stack)."""
    keys = list(params.keys())
    n_layers = len(keys) // 2
    y, acts = mlp_forward(x32, params, keep_activations=True)
    mrows, dout = y.shape
    g = (np.float32(2.0) / np.float32(mrows * dout)) * (y - target32)      # dL/dy
    new_params = OrderedDict()
    for l in reversed(range(n_layers)):
        w = params[keys[2 * l]]
        h_in = acts[l]
        gw = h_in.T @ g
        gb = g.sum(axis=0)
        new_params[keys[2 * l]] = (w - np.float32(inner_lr) * gw).astype(np.float32)
        new_params[keys[2 * l + 1]] = (params[keys[2 * l + 1]] - np.float32(inner_lr) * gb).astype(np.float32)
        if l > 0:
            g = (g @ w.T) * (acts[l] > 0).astype(np.float32)
    return OrderedDict((k, new_params[k]) for k in keys)


def adapt(obs_list, act_list, obs_next_list, params, norm, inner_lr):
    """obs_list: K arrays [M, D] etc. (sampler.py:83-90).  Returns K adapted OrderedDicts."""
    out = []
    for ob, ac, ob_next in zip(obs_list, act_list, obs_next_list):
        ob = np.asarray(ob, np.float64)
        ac = np.asarray(ac, np.float64)
        ob_next = np.asarray(ob_next, np.float64)
        obs_n = normalize(ob, *norm["obs"])                                  # :334-339
        act_n = normalize(ac, *norm["act"])
        delta_n = normalize(ob_next - ob, *norm["delta"])
        x32 = np.concatenate([obs_n, act_n], axis=1).astype(np.float32)
        out.append(adapt_one_task(x32, delta_n.astype(np.float32), params, inner_lr))
    return out


# --------------------------------------------------------------------------------------------------
# seeded synthetic problem generator shared by tests / bench / golden generation
# --------------------------------------------------------------------------------------------------
ENV_SPECS = {
    # name: (obs_dim, act_dim, ctrl_limit, dt, reward kind)
    "half_cheetah": (20, 6, 1.0, 0.01, REWARD_HALF_CHEETAH),
    "ant": (41, 8, 150.0, 0.02, REWARD_ANT),
    "arm_7dof": (17, 7, 1.0, 0.02, REWARD_ARM),   # qpos 7 + qvel 7 + 3; ctrl range illustrative
}


def make_problem(env="half_cheetah", hidden_sizes=(512, 512), n_sets=1, m=1, out_scale=0.1, seed=0):
    obs_dim, act_dim, lim, dt, kind = ENV_SPECS[env]
    low = -lim * np.ones(act_dim)
    high = lim * np.ones(act_dim)
    sets = [xavier_params(np.random.RandomState(1000 * seed + e), obs_dim + act_dim, hidden_sizes, obs_dim,
                          out_scale=out_scale) for e in range(n_sets)]
    norm = make_normalization(np.random.RandomState(seed + 1), obs_dim, act_dim, low, high)
    rng = np.random.RandomState(seed + 2)
    obs0 = norm["obs"][0] + norm["obs"][1] * rng.normal(size=(m, obs_dim))
    return dict(env=env, obs_dim=obs_dim, act_dim=act_dim, low=low, high=high, dt=dt, reward_kind=kind,
                param_sets=sets, norm=norm, obs0=obs0, hidden_sizes=tuple(hidden_sizes))


def sample_rs_actions(seed, low, high, horizon, rows):
    """The reference's own draw: uniform(low, high, (H*rows, A)).reshape(H, rows, A) (mpc_controller.py:67-69,114)."""
    rng = np.random.RandomState(seed)
    return rng.uniform(low=low, high=high, size=(horizon * rows,) + low.shape).reshape((horizon, rows, -1))


def make_adapt_context(seed, prob, k, m_rows):
    rng = np.random.RandomState(seed)
    obs, act, nxt = [], [], []
    for _ in range(k):
        o = prob["norm"]["obs"][0] + prob["norm"]["obs"][1] * rng.normal(size=(m_rows, prob["obs_dim"]))
        a = rng.uniform(prob["low"], prob["high"], size=(m_rows, prob["act_dim"]))
        d = prob["norm"]["delta"][0] + prob["norm"]["delta"][1] * rng.normal(size=(m_rows, prob["obs_dim"]))
        obs.append(o)
        act.append(a)
        nxt.append(o + d)
    return obs, act, nxt


# --------------------------------------------------------------------------------------------------
# ReBAL: recurrent dynamics model (single-layer LSTM) and its planner   
#   dynamics/rnn_dynamics.py:233-252 (predict), dynamics/core/utils.py:145-238 (create_rnn: tf.nn.rnn_cell.LSTMCell +
#   tf.layers.dense output), policies/rnn_mpc_controller.py:57-65, 112-134, 165-187.
# TF 1.13 LSTMCell semantics (published; TensorFlow itself is not installable here -> PARITY UNPINNED for the cell math):
#   z = [x, h_prev] @ kernel + bias ;  i, j, f, o = split(z, 4) ;  c = sigmoid(f + 1.0) * c_prev + sigmoid(i) * act(j) ;
#   h = sigmoid(o) * act(c)            (forget_bias = 1.0, no peepholes / clipping / projection), act = tanh by default.
# --------------------------------------------------------------------------------------------------
def rnn_param_keys():
    return ["rnn/lstm_cell/kernel", "rnn/lstm_cell/bias", "output/kernel", "output/bias"]


def xavier_rnn_params(rng, in_dim, hidden, out_dim, out_scale=1.0):
    params = OrderedDict()
    lim = np.sqrt(6.0 / (in_dim + hidden + 4 * hidden))
    params["rnn/lstm_cell/kernel"] = rng.uniform(-lim, lim, size=(in_dim + hidden, 4 * hidden)).astype(np.float32)
    params["rnn/lstm_cell/bias"] = np.zeros(4 * hidden, np.float32)
    lim = np.sqrt(6.0 / (hidden + out_dim))
    params["output/kernel"] = (rng.uniform(-lim, lim, size=(hidden, out_dim)) * out_scale).astype(np.float32)
    params["output/bias"] = np.zeros(out_dim, np.float32)
    return params


def _sigmoid32(x):
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def lstm_step(x32, c_prev, h_prev, params):
    """One LSTMCell step in float32.  x32 [n, in], c_prev/h_prev [n, H]  ->  (delta_norm [n, D], c, h)."""
    hsz = h_prev.shape[1]
    z = np.concatenate([x32, h_prev], axis=1).astype(np.float32) @ params["rnn/lstm_cell/kernel"] + params["rnn/lstm_cell/bias"]
    i, j, f, o = z[:, :hsz], z[:, hsz:2 * hsz], z[:, 2 * hsz:3 * hsz], z[:, 3 * hsz:]
    c = _sigmoid32(f + np.float32(1.0)) * c_prev + _sigmoid32(i) * np.tanh(j)
    h = _sigmoid32(o) * np.tanh(c)
    y = h.astype(np.float32) @ params["output/kernel"] + params["output/bias"]
    return y.astype(np.float32), c.astype(np.float32), h.astype(np.float32)


def rnn_predict(obs, act, hidden, params, norm):
    """RNNDynamicsModel.predict (rnn_dynamics.py:233-252): hidden = (c [n,H], h [n,H]) float32.
    Returns (next_obs float64 [n, D], (c, h))."""
    obs = np.asarray(obs, np.float64)
    act = np.asarray(act, np.float64)
    obs_n = normalize(obs, *norm["obs"])
    act_n = normalize(act, *norm["act"])
    x32 = np.concatenate([obs_n, act_n], axis=1).astype(np.float32)
    y, c, h = lstm_step(x32, np.asarray(hidden[0], np.float32), np.asarray(hidden[1], np.float32), params)
    delta = denormalize(y, *norm["delta"])
    return obs + delta, (c, h)


def rnn_rollout_returns(observations, actions, hidden, params, norm, reward_kind, dt, discount=1.0):
    """RNNMPCController.get_rs_action loop (rnn_mpc_controller.py:112-134): hidden = (c [m,H], h [m,H]) is repeated n
    times per env (repeat_hidden :165-187)."""
    observations = np.asarray(observations, np.float64)
    actions = np.asarray(actions, np.float64)
    hh, rows, _ = actions.shape
    m = observations.shape[0]
    n = rows // m
    rew = reward_fn(reward_kind, dt)
    returns = np.zeros((rows,))
    observation = np.repeat(observations, n, axis=0)
    hid = (np.repeat(np.asarray(hidden[0], np.float32), n, axis=0), np.repeat(np.asarray(hidden[1], np.float32), n, axis=0))
    for t in range(hh):
        next_observation, hid = rnn_predict(observation, actions[t], hid, params, norm)
        rewards = rew(observation, actions[t], next_observation)
        returns += discount ** t * rewards
        observation = next_observation
    return returns.reshape(m, n)


def rnn_rs_plan(observations, actions, hidden, params, norm, reward_kind, dt, discount=1.0):
    """Returns (chosen [m,A], best [m], returns [m,n], new_hidden) -- new_hidden = one predict on the REAL observations with
    the chosen actions (rnn_mpc_controller.py:63)."""
    returns = rnn_rollout_returns(observations, actions, hidden, params, norm, reward_kind, dt, discount)
    m, n = returns.shape
    cand_a = np.asarray(actions, np.float64)[0].reshape(m, n, -1)
    best = np.argmax(returns, axis=1)
    chosen = cand_a[range(m), best]
    _, new_hidden = rnn_predict(np.asarray(observations, np.float64), chosen, hidden, params, norm)
    return chosen, best, returns, new_hidden


# --------------------------------------------------------------------------------------------------
# device candidate sampler (throughput mode of get_random_action): Philox4x32-10 restated in numpy
# --------------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) on arrays of
    counters [n, 4] uint32 with one key (k0, k1).  Returns [n, 4] uint32.  This is the published algorithm the CUDA kernel
    learning_to_adapt_b200/csrc/sample.cuh implements; the reference itself draws with numpy's MT19937
    (policies/mpc_controller.py:67-69), which the parity mode of the controller keeps."""
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    w0, w1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy(), c[:, 3].copy()
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2                       # 32 x 32 -> 64 bit products
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0, k1 = (k0 + w0) & mask, (k1 + w1) & mask
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def sample_rs_actions_device(seed, call_index, low, high, horizon, rows):
    """The candidate tensor [H, rows, A] float32 that l2a_plan_run / l2a_sample_uniform draw for (seed, call_index):
    element block b (4 consecutive elements of the flattened tensor) = Philox(counter = (b_lo, b_hi, call_lo, call_hi),
    key = (seed_lo, seed_hi)); u = (word >> 8) * 2^-24 in [0, 1); value = fma(high - low, u, low) in float32."""
    low32, high32 = np.asarray(low, np.float32), np.asarray(high, np.float32)
    A = low32.shape[0]
    total = horizon * rows * A
    nblk = (total + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([blk & np.uint64(0xFFFFFFFF), blk >> np.uint64(32),
                    np.full(nblk, call_index & 0xFFFFFFFF, np.uint64), np.full(nblk, (call_index >> 32) & 0xFFFFFFFF, np.uint64)], axis=1)
    words = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:total]
    u = (words >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)
    j = np.arange(total) % A
    span = (high32 - low32).astype(np.float32)           # the kernel forms hi - lo in float32
    val = span[j].astype(np.float64) * u + low32[j].astype(np.float64)   # exact product (24 x 24 bits), one rounding below = fma
    return val.astype(np.float32).reshape(horizon, rows, A)


# --------------------------------------------------------------------------------------------------
# training: the reference's fit procedures restated on the host
#   MLPDynamicsModel.fit      dynamics/mlp_dynamics.py:91-202   (Adam on the batch MSE, batches of consecutive rows)
#   MetaMLPDynamicsModel.fit  dynamics/meta_mlp_dynamics.py:96-140, 167-274, 353-383 (MAML: Adam on the post-update loss)
# tf.train.AdamOptimizer (TF 1.13, published update rule -- TensorFlow is not installable here, PARITY UNPINNED for the optimiser):
#   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  theta -= lr_t * m / (sqrt(v) + eps)
# The numpy draws (train/validation split, batch order, MAML windows) are the ones the product makes, in the same order, so
# a seeded product fit and a seeded oracle fit see identical batches.
# --------------------------------------------------------------------------------------------------
def mlp_loss_and_grads(x32, y32, params):
    """mean((y - f(x))^2) and its gradient w.r.t. every parameter, float32 (manual backprop of the dense/ReLU stack)."""
    keys = list(params.keys())
    n_layers = len(keys) // 2
    out, acts = mlp_forward(x32, params, keep_activations=True)
    rows, dout = out.shape
    loss = np.float32(np.mean((y32 - out) ** 2))
    g = (np.float32(2.0) / np.float32(rows * dout)) * (out - y32)
    grads = OrderedDict()
    for l in reversed(range(n_layers)):
        grads[keys[2 * l]] = (acts[l].T @ g).astype(np.float32)
        grads[keys[2 * l + 1]] = g.sum(axis=0).astype(np.float32)
        if l > 0:
            g = (g @ params[keys[2 * l]].T) * (acts[l] > 0).astype(np.float32)
    return loss, OrderedDict((k, grads[k]) for k in keys)


class AdamTF(object):
    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8, dtype=np.float32):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = OrderedDict((k, np.zeros_like(v, dtype=dtype)) for k, v in params.items())
        self.v = OrderedDict((k, np.zeros_like(v, dtype=dtype)) for k, v in params.items())

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in params:
            g = grads[k]
            self.m[k] = (self.b1 * self.m[k] + (1.0 - self.b1) * g).astype(g.dtype)
            self.v[k] = (self.b2 * self.v[k] + (1.0 - self.b2) * g * g).astype(g.dtype)
            params[k] = (params[k] - lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)).astype(g.dtype)


def train_test_split(obs, act, delta, test_split_ratio=0.2):
    """mlp_dynamics.py:273-285 (np.random.shuffle on the global stream)."""
    indices = np.arange(obs.shape[0])
    np.random.shuffle(indices)
    split_idx = int(obs.shape[0] * (1 - test_split_ratio))
    tr, te = indices[:split_idx], indices[split_idx:]
    return obs[tr], act[tr], delta[tr], obs[te], act[te], delta[te]


def fit_mlp(params, obs, act, obs_next, norm, epochs, batch_size, learning_rate, valid_split_ratio=0.2, persistency=0.99):
    """One MLPDynamicsModel.fit call from fresh optimiser state on raw (obs, act, obs_next); returns (params, valid losses)."""
    obs_n, act_n = normalize(obs, *norm["obs"]), normalize(act, *norm["act"])
    delta_n = normalize(obs_next - obs, *norm["delta"])
    o_tr, a_tr, d_tr, o_te, a_te, d_te = train_test_split(obs_n, act_n, delta_n, valid_split_ratio)
    xt, yt = np.concatenate([o_tr, a_tr], axis=1).astype(np.float32), d_tr.astype(np.float32)
    xv, yv = np.concatenate([o_te, a_te], axis=1).astype(np.float32), d_te.astype(np.float32)
    params = OrderedDict((k, v.copy()) for k, v in params.items())
    adam = AdamTF(params, learning_rate)
    avg = prev = None
    valid = []
    for epoch in range(epochs):
        starts = np.arange(0, xt.shape[0], batch_size)
        starts = starts[np.random.permutation(len(starts))]
        for s in starts:
            _, grads = mlp_loss_and_grads(xt[s:s + batch_size], yt[s:s + batch_size], params)
            adam.step(params, grads)
        vl = float(np.mean((yv - mlp_forward(xv, params)) ** 2))
        valid.append(vl)
        if avg is None:
            avg, prev = (vl / 1.5, vl / 2) if vl < 0 else (1.5 * vl, 2 * vl)
        avg = persistency * avg + (1.0 - persistency) * vl
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return params, valid


def fit_maml(params, obs, act, obs_next, norm, epochs, batch_size, meta_batch_size, learning_rate, inner_learning_rate,
             valid_split_ratio=0.2, persistency=0.99):
    """One MetaMLPDynamicsModel.fit call from fresh optimiser state; obs/act/obs_next [paths, T, dim].  The second-order
    gradient of the post-update loss is taken by torch autograd on the CPU in FLOAT64 (an independent execution of the math
    the product runs in float32 on the device)."""
    import torch
    obs_n, act_n = normalize(obs, *norm["obs"]), normalize(act, *norm["act"])
    delta_n = normalize(obs_next - obs, *norm["delta"])
    o_tr, a_tr, d_tr, o_te, a_te, d_te = train_test_split(obs_n, act_n, delta_n, valid_split_ratio)
    f32 = lambda a: torch.tensor(np.asarray(a, np.float32).astype(np.float64))      # the float32 feed, then float64 math
    xtr, ytr = f32(np.concatenate([o_tr, a_tr], axis=2)), f32(d_tr)
    xte, yte = f32(np.concatenate([o_te, a_te], axis=2)), f32(d_te)
    keys = list(params.keys())
    theta = [torch.tensor(params[k].astype(np.float64), requires_grad=True) for k in keys]
    np_theta = OrderedDict((k, params[k].astype(np.float64)) for k in keys)
    adam = AdamTF(np_theta, learning_rate, dtype=np.float64)

    def fwd(x, ps):
        h = x
        for l in range(len(ps) // 2):
            h = h @ ps[2 * l] + ps[2 * l + 1]
            if l < len(ps) // 2 - 1:
                h = torch.relu(h)
        return h

    def get_batch(x):
        num_paths, len_path = x.shape[:2]
        ip = np.random.randint(0, num_paths, size=meta_batch_size)
        ib = np.random.randint(batch_size, len_path - batch_size, size=meta_batch_size)
        return ip, ib

    steps_train = max(int(np.prod(xtr.shape[:2]) / (meta_batch_size * batch_size * 2)), 1)
    steps_test = max(int(np.prod(xte.shape[:2]) / (meta_batch_size * batch_size * 2)), 1)
    avg = prev = None
    valid = []
    for epoch in range(epochs):
        for _ in range(steps_train):
            ip, ib = get_batch(xtr)
            post = []
            for p, b in zip(ip, ib):
                xw, yw = xtr[p, b - batch_size:b + batch_size], ytr[p, b - batch_size:b + batch_size]
                pre_loss = torch.mean((yw[:batch_size] - fwd(xw[:batch_size], theta)) ** 2)
                g = torch.autograd.grad(pre_loss, theta, create_graph=True)
                adapted = [w - inner_learning_rate * gi for w, gi in zip(theta, g)]
                post.append(torch.mean((yw[batch_size:] - fwd(xw[batch_size:], adapted)) ** 2))
            grads = torch.autograd.grad(torch.stack(post).mean(), theta)
            adam.step(np_theta, OrderedDict((k, gi.numpy()) for k, gi in zip(keys, grads)))
            with torch.no_grad():
                for t, k in zip(theta, keys):
                    t.copy_(torch.tensor(np_theta[k]))
        vls = []
        with torch.no_grad():
            for _ in range(steps_test):
                ip, ib = get_batch(xte)
                xs = torch.cat([xte[p, b - batch_size:b + batch_size] for p, b in zip(ip, ib)])
                ys = torch.cat([yte[p, b - batch_size:b + batch_size] for p, b in zip(ip, ib)])
                vls.append(float(torch.mean((ys - fwd(xs, theta)) ** 2)))
        vl = float(np.mean(vls))
        valid.append(vl)
        if avg is None:
            avg, prev = (vl / 1.5, vl / 2) if vl < 0 else (1.5 * vl, 2 * vl)
        avg = persistency * avg + (1.0 - persistency) * vl
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return OrderedDict((k, np_theta[k].astype(np.float32)) for k in keys), valid
