"""Generate the golden fixtures in this directory by running the REFERENCE'S OWN CODE.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed from upstream, unmodified, imported from /root/reference:
  * ``learning_to_adapt.policies.mpc_controller.MPCController`` (random shooting and CEM) -- verbatim;
  * ``MLPDynamicsModel.predict`` / ``_normalize_data`` and ``MetaMLPDynamicsModel.predict`` /
    ``_predict`` / ``_pad_inputs`` / ``adapt`` / ``switch_to_pre_adapt`` -- the upstream method bodies,
    called on a stand-in ``self`` whose TF-session pieces (``f_delta_pred``, ``sess.run``) are
    replaced by a float32 numpy dense stack (TensorFlow 1.13.1 cannot be installed here);
  * the env ``reward`` methods -- their upstream source text is extracted with ``ast`` and executed
    (the env modules themselves import mujoco_py -> libmujoco131.so, which does not exist here).

The fixtures store inputs + the reference's outputs; tests/test_oracle_golden.py pins oracle/ to them and
the GPU parity tests compare the CUDA path against the same files.
"""
import ast
import os
import sys
import textwrap
import types
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


class _Anything(types.ModuleType):
    """A stand-in for the tensorflow module: every attribute / call yields another stand-in."""

    def __init__(self, name="tensorflow"):
        super().__init__(name)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Anything(self.__name__ + "." + item)
        object.__setattr__(self, item, child)          # cache: attributes set on a child persist
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


def import_reference():
    tf = _Anything("tensorflow")
    sys.modules["tensorflow"] = tf
    for name in ("gym", "gym.spaces", "pyprind", "mpi4py"):
        sys.modules.setdefault(name, _Anything(name))
    sys.path.insert(0, REF)
    from learning_to_adapt.policies.mpc_controller import MPCController
    from learning_to_adapt.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt.dynamics import meta_mlp_dynamics
    from learning_to_adapt.spaces.box import Box
    return tf, MPCController, MLPDynamicsModel, meta_mlp_dynamics, Box


def reference_reward(env_file, dt):
    """Extract ``def reward(self, obs, action, next_obs)`` from the upstream env source and bind it."""
    src = open(os.path.join(REF, "learning_to_adapt/envs", env_file)).read()
    tree = ast.parse(src)
    fn_src = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "reward":
            fn_src = ast.get_source_segment(src, node)
    assert fn_src is not None
    ns = {"np": np}
    exec(textwrap.dedent(fn_src), ns)
    holder = types.SimpleNamespace(dt=dt)
    return lambda obs, action, next_obs: ns["reward"](holder, obs, action, next_obs)


def main():
    from oracle import mpc_oracle as O

    tf, MPCController, MLPDynamicsModel, meta_mod, Box = import_reference()
    MetaMLPDynamicsModel = meta_mod.MetaMLPDynamicsModel
    out = {}

    # ------------------------------------------------------------------ rewards
    rng = np.random.RandomState(11)
    for env_name, env_file in (("half_cheetah", "half_cheetah_env.py"), ("ant", "ant_env.py"),
                               ("arm_7dof", "arm_7dof_env.py")):
        d, a, lim, dt, kind = O.ENV_SPECS[env_name]
        obs = rng.normal(size=(9, d))
        nxt = obs + 0.05 * rng.normal(size=(9, d))
        act = rng.uniform(-lim, lim, size=(9, a))
        r = reference_reward(env_file, dt)(obs, act, nxt)
        out["reward_%s_obs" % env_name] = obs
        out["reward_%s_act" % env_name] = act
        out["reward_%s_next" % env_name] = nxt
        out["reward_%s_out" % env_name] = r

    # ------------------------------------------------------------------ stand-in model objects
    def fake_mlp_model(prob, params):
        me = types.SimpleNamespace()
        me.obs_space_dims = prob["obs_dim"]
        me.action_space_dims = prob["act_dim"]
        me.normalize_input = True
        me.normalization = prob["norm"]
        # TF feed casts to float32; the dense stack is the only substituted piece
        me.f_delta_pred = lambda o, a: O.f_delta_pred(o, a, params)
        me._normalize_data = types.MethodType(MLPDynamicsModel._normalize_data, me)
        me.predict = types.MethodType(MLPDynamicsModel.predict, me)
        return me

    class FakeEnv(object):
        def __init__(self, prob, env_file):
            self.action_space = Box(prob["low"], prob["high"])
            self.observation_space = Box(-np.inf * np.ones(prob["obs_dim"]), np.inf * np.ones(prob["obs_dim"]))
            self._r = reference_reward(env_file, prob["dt"])
            self.step_rewards = []

        def reward(self, obs, action, next_obs):
            r = self._r(obs, action, next_obs)
            self.step_rewards.append(np.array(r))
            return r

    # ------------------------------------------------------------------ one-step predict (MLP)
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=3)
    model = fake_mlp_model(prob, prob["param_sets"][0])
    rng = np.random.RandomState(5)
    p_obs = prob["norm"]["obs"][0] + prob["norm"]["obs"][1] * rng.normal(size=(7, prob["obs_dim"]))
    p_act = rng.uniform(prob["low"], prob["high"], size=(7, prob["act_dim"]))
    out["predict_obs"] = p_obs
    out["predict_act"] = p_act
    out["predict_out"] = model.predict(p_obs, p_act)

    # ------------------------------------------------------------------ random shooting, verbatim planner
    def run_rs(tag, env_name, env_file, hidden, n, h, m, discount, seed):
        prob = O.make_problem(env_name, hidden_sizes=hidden, n_sets=1, m=m, seed=seed)
        env = FakeEnv(prob, env_file)
        model = fake_mlp_model(prob, prob["param_sets"][0])
        ctrl = MPCController("policy", env, model, discount=discount, n_candidates=n, horizon=h)
        np.random.seed(seed + 100)
        state = np.random.get_state()
        acts, info = ctrl.get_actions(prob["obs0"])
        np.random.set_state(state)
        drawn = ctrl.get_random_action(h * n * m).reshape((h, n * m, -1))
        out[tag + "_meta"] = np.array([n, h, m, seed], np.int64)
        out[tag + "_discount"] = np.array(discount)
        out[tag + "_actions"] = drawn
        out[tag + "_step_rewards"] = np.stack(env.step_rewards)          # [H, n*m]
        out[tag + "_chosen"] = acts

    run_rs("rs_hc", "half_cheetah", "half_cheetah_env.py", (32, 32), 64, 5, 3, 1.0, 7)
    run_rs("rs_hc_disc", "half_cheetah", "half_cheetah_env.py", (48,), 33, 4, 2, 0.9, 8)
    run_rs("rs_ant", "ant", "ant_env.py", (32, 32, 32), 40, 6, 2, 1.0, 9)
    run_rs("rs_arm", "arm_7dof", "arm_7dof_env.py", (32, 32), 50, 4, 1, 1.0, 10)

    # get_action (single obs -> [1, A])
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=12)
    env = FakeEnv(prob, "half_cheetah_env.py")
    ctrl = MPCController("policy", env, fake_mlp_model(prob, prob["param_sets"][0]), n_candidates=20, horizon=3)
    np.random.seed(112)
    act, _ = ctrl.get_action(prob["obs0"][0])
    out["get_action_shape"] = np.array(act.shape, np.int64)
    out["get_action_out"] = act

    # ------------------------------------------------------------------ CEM, verbatim planner (bug-compatible)
    def run_cem(tag, n, h, m, iters, pct, alpha, seed):
        prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=m, seed=seed)
        env = FakeEnv(prob, "half_cheetah_env.py")
        model = fake_mlp_model(prob, prob["param_sets"][0])
        ctrl = MPCController("policy", env, model, use_cem=True, n_candidates=n, horizon=h,
                             num_cem_iters=iters, percent_elites=pct, alpha=alpha)
        np.random.seed(seed + 100)
        acts, _ = ctrl.get_actions(prob["obs0"])
        out[tag + "_meta"] = np.array([n, h, m, iters, seed], np.int64)
        out[tag + "_pct_alpha"] = np.array([pct, alpha])
        out[tag + "_chosen"] = acts
        out[tag + "_last_step_rewards"] = np.stack(env.step_rewards[-h:])

    run_cem("cem_m1", 60, 4, 1, 3, 0.1, 0.1, 21)
    run_cem("cem_m2", 40, 3, 2, 2, 0.2, 0.1, 22)

    # ------------------------------------------------------------------ GrBAL adapt + per-task predict
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32, 32), n_sets=1, m=3, seed=31)
    theta = prob["param_sets"][0]
    K, M, MBS, LR = 3, 8, 5, 1e-2
    ctx_obs, ctx_act, ctx_next = O.make_adapt_context(41, prob, K, M)

    class FakeSession(object):
        """Stands in for tf.get_default_session(): evaluates the two fetches the upstream methods make."""

        def run(self, fetches, feed_dict=None):
            me = fake_self
            if fetches is me._adapted_params_marker or (isinstance(fetches, list) and fetches and
                                                        all(f in me._adapted_params for f in fetches)):
                obs = feed_dict[me.obs_ph]
                act = feed_dict[me.act_ph]
                delta = feed_dict[me.delta_ph]
                # graph 96-120: split into meta_batch_size tasks, each into (pre | post) halves
                x = np.concatenate([obs, act], axis=1).astype(np.float32)
                x_tasks = np.split(x, me.meta_batch_size, axis=0)
                d_tasks = np.split(delta.astype(np.float32), me.meta_batch_size, axis=0)
                res = []
                for f in fetches:
                    idx = me._adapted_params.index(f)
                    pre_x = np.split(x_tasks[idx], 2, axis=0)[0]
                    pre_d = np.split(d_tasks[idx], 2, axis=0)[0]
                    res.append(O.adapt_one_task(pre_x, pre_d, theta, me.inner_learning_rate))
                return res
            # post_update_delta[:K] with placeholders fed
            obs = feed_dict[me.obs_ph]
            act = feed_dict[me.act_ph]
            x_o = np.split(obs, me.meta_batch_size, axis=0)
            x_a = np.split(act, me.meta_batch_size, axis=0)
            res = []
            for f in fetches:
                idx = me.post_update_delta.index(f)
                params = OrderedDict((k, feed_dict[me.network_phs_meta_batch[idx][k]]) for k in theta.keys())
                res.append(O.f_delta_pred(x_o[idx], x_a[idx], params))
            return res

    tf.get_default_session = lambda: FakeSession()
    meta_mod.tf = tf

    fake_self = types.SimpleNamespace()
    fake_self.obs_space_dims = prob["obs_dim"]
    fake_self.action_space_dims = prob["act_dim"]
    fake_self.normalize_input = True
    fake_self.normalization = prob["norm"]
    fake_self.meta_batch_size = MBS
    fake_self.inner_learning_rate = LR
    fake_self.obs_ph, fake_self.act_ph, fake_self.delta_ph = "obs_ph", "act_ph", "delta_ph"
    fake_self._adapted_params = ["adapted_%d" % i for i in range(MBS)]
    fake_self._adapted_params_marker = object()
    fake_self.post_update_delta = ["post_delta_%d" % i for i in range(MBS)]
    fake_self.network_phs_meta_batch = [OrderedDict((k, "ph_%d_%s" % (i, k)) for k in theta.keys())
                                        for i in range(MBS)]
    fake_self._networks = [types.SimpleNamespace(get_param_values=lambda: theta, set_params=lambda p: None)]
    fake_self._prev_params = None
    fake_self._adapted_param_values = None
    fake_self._num_adapted_models = None
    fake_self.f_delta_pred = lambda o, a: O.f_delta_pred(o, a, theta)
    # network_params_feed_dict is a property upstream; evaluate its body against fake_self on demand
    prop = MetaMLPDynamicsModel.network_params_feed_dict.fget

    class _Proxy(object):
        def __getattr__(self, item):
            if item == "network_params_feed_dict":
                return prop(fake_self)
            return getattr(fake_self, item)

        def __setattr__(self, key, value):
            setattr(fake_self, key, value)

    proxy = _Proxy()
    for name in ("_pad_inputs", "_normalize_data", "_predict", "predict", "adapt", "switch_to_pre_adapt"):
        setattr(fake_self, name, types.MethodType(getattr(MetaMLPDynamicsModel, name), proxy))

    fake_self.adapt(ctx_obs, ctx_act, ctx_next)
    adapted = fake_self._adapted_param_values
    assert len(adapted) == K
    for k in range(K):
        for key in theta.keys():
            out["adapt_theta%d_%s" % (k, key.replace("/", "."))] = adapted[k][key]
    n_per = 6
    rng = np.random.RandomState(51)
    q_obs = prob["norm"]["obs"][0] + prob["norm"]["obs"][1] * rng.normal(size=(K * n_per, prob["obs_dim"]))
    q_act = rng.uniform(prob["low"], prob["high"], size=(K * n_per, prob["act_dim"]))
    out["adapt_meta"] = np.array([K, M, MBS, n_per], np.int64)
    out["adapt_lr"] = np.array(LR)
    out["adapt_query_obs"] = q_obs
    out["adapt_query_act"] = q_act
    post = fake_self.predict(q_obs, q_act)
    assert post.shape == (K * n_per, prob["obs_dim"]), post.shape   # upstream drops the padded tasks
    out["adapt_post_predict"] = post
    fake_self.switch_to_pre_adapt()
    assert fake_self._adapted_param_values is None
    out["adapt_pre_predict"] = fake_self.predict(q_obs, q_act)

    # GrBAL planning: verbatim planner + upstream per-task predict (reading (i): env k uses theta'_k)
    fake_self.adapt(ctx_obs, ctx_act, ctx_next)
    env = FakeEnv(prob, "half_cheetah_env.py")
    ctrl = MPCController("policy", env, fake_self, n_candidates=24, horizon=4)
    np.random.seed(161)
    state = np.random.get_state()
    acts, _ = ctrl.get_actions(prob["obs0"])
    np.random.set_state(state)
    out["grbal_rs_meta"] = np.array([24, 4, 3, 31], np.int64)
    out["grbal_rs_actions"] = ctrl.get_random_action(4 * 24 * 3).reshape((4, 24 * 3, -1))
    out["grbal_rs_step_rewards"] = np.stack(env.step_rewards)
    out["grbal_rs_chosen"] = acts

    # ------------------------------------------------------------------ ReBAL: verbatim RNNMPCController + LSTM stand-in
    import collections
    LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))
    tf.nn.rnn_cell.LSTMStateTuple = LSTMStateTuple
    tf.contrib.rnn.LSTMStateTuple = LSTMStateTuple
    from learning_to_adapt.policies import rnn_mpc_controller as rnn_ctrl_mod
    from learning_to_adapt.dynamics import rnn_dynamics as rnn_dyn_mod
    rnn_ctrl_mod.tf = tf
    RNNDynamicsModel = rnn_dyn_mod.RNNDynamicsModel
    prob = O.make_problem("half_cheetah", hidden_sizes=(32,), n_sets=1, m=2, seed=61)
    HS = 24
    rparams = O.xavier_rnn_params(np.random.RandomState(62), prob["obs_dim"] + prob["act_dim"], HS, prob["obs_dim"], out_scale=0.1)

    rnn_self = types.SimpleNamespace()
    rnn_self.obs_space_dims = prob["obs_dim"]
    rnn_self.action_space_dims = prob["act_dim"]
    rnn_self.normalize_input = True
    rnn_self.normalization = prob["norm"]

    def f_delta_pred(obs3, act3, hidden):
        # graph: concat -> dynamic_rnn (time axis of length 1) -> dense output; float32 feed
        x = np.concatenate([obs3[:, 0, :], act3[:, 0, :]], axis=1).astype(np.float32)
        y, c, h = O.lstm_step(x, np.asarray(hidden[0], np.float32), np.asarray(hidden[1], np.float32), rparams)   # repeat_hidden hands a [c, h] list
        return y[:, None, :], LSTMStateTuple(c, h)

    rnn_self.f_delta_pred = f_delta_pred
    rnn_self._normalize_data = types.MethodType(RNNDynamicsModel._normalize_data, rnn_self)
    rnn_self.predict = types.MethodType(RNNDynamicsModel.predict, rnn_self)
    rnn_self.get_initial_hidden = lambda batch_size: LSTMStateTuple(np.zeros((batch_size, HS), np.float32),
                                                                    np.zeros((batch_size, HS), np.float32))
    env = FakeEnv(prob, "half_cheetah_env.py")
    n_c, hor = 30, 4
    rctrl = rnn_ctrl_mod.RNNMPCController("policy", env, rnn_self, n_candidates=n_c, horizon=hor)
    rctrl.reset(dones=[True, True])
    np.random.seed(171)
    obs_t = np.array(prob["obs0"])
    chosen_seq, hid_c, hid_h = [], [], []
    for step in range(3):                       # three consecutive planning calls: the hidden state carries over
        acts, _ = rctrl.get_actions(obs_t)
        chosen_seq.append(np.array(acts))
        hid_c.append(np.array(rctrl._hidden_state.c))
        hid_h.append(np.array(rctrl._hidden_state.h))
        obs_t = obs_t + 0.05 * np.random.RandomState(step).normal(size=obs_t.shape)
    out["rebal_meta"] = np.array([n_c, hor, 2, 61, HS], np.int64)
    out["rebal_chosen"] = np.stack(chosen_seq)
    out["rebal_hidden_c"] = np.stack(hid_c)
    out["rebal_hidden_h"] = np.stack(hid_h)
    out["rebal_step_rewards"] = np.stack(env.step_rewards)          # [3*H, n*m]
    # get_action on one observation returns ([1, A], {})
    rctrl1 = rnn_ctrl_mod.RNNMPCController("policy", env, rnn_self, n_candidates=10, horizon=2)
    rctrl1.reset(dones=[True])
    np.random.seed(172)
    a1, _ = rctrl1.get_action(prob["obs0"][0])
    out["rebal_get_action_shape"] = np.array(a1.shape, np.int64)
    # recurrent CEM (bug-compatible mask, mean replaced rather than smoothed)
    env = FakeEnv(prob, "half_cheetah_env.py")
    rcem = rnn_ctrl_mod.RNNMPCController("policy", env, rnn_self, n_candidates=40, horizon=3, use_cem=True, num_cem_iters=2,
                                         percent_elites=0.2)
    rcem.reset(dones=[True, True])
    np.random.seed(173)
    acts, _ = rcem.get_actions(np.array(prob["obs0"]))
    out["rebal_cem_meta"] = np.array([40, 3, 2, 2], np.int64)
    out["rebal_cem_chosen"] = np.array(acts)
    out["rebal_cem_last_step_rewards"] = np.stack(env.step_rewards[-3:])

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
