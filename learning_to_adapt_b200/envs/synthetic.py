"""MuJoCo-free stand-ins for the reference env classes: spaces, dt and the closed-form ``reward`` only.

The real MuJoCo step stays on the host and is off the hot path (BASELINE.json north_star); libmujoco131.so is not
available in this image.  These objects carry exactly what the planner reads from an env
(policies/mpc_controller.py:34-39, 67-69, 125): ``action_space``, ``observation_space``, ``dt`` and ``reward``.
``reward`` here is the host-side API method of the env (numpy, float64); the planner never calls it -- the fused
kernel evaluates the same closed form on the device, selected through ``l2a_reward_kind``.
"""
import numpy as np

from learning_to_adapt_b200 import _native as N
from learning_to_adapt_b200.spaces.box import Box

# name -> (obs_dim, act_dim, ctrl limit, dt, reward family)
#   half_cheetah: envs/half_cheetah_env.py:32-37, assets/half_cheetah.xml:40,43,88-93
#   ant         : envs/ant_env.py:31-37, assets/ant.xml:3,71-78
#   arm_7dof    : envs/arm_7dof_env.py:84-89, assets/arm_7dof.xml:4,83-89
ENV_SPECS = {
    "half_cheetah": (20, 6, 1.0, 0.01, N.REWARD_HALF_CHEETAH),
    "ant": (41, 8, 150.0, 0.02, N.REWARD_ANT),
    "arm_7dof": (17, 7, 1.0, 0.02, N.REWARD_ARM),
}

# reference env class name -> reward family, for real reference env objects handed to MPCController
REWARD_KIND_BY_CLASS = {
    "HalfCheetahEnv": N.REWARD_HALF_CHEETAH, "HalfCheetahBlocksEnv": N.REWARD_HALF_CHEETAH,
    "HalfCheetahHFieldEnv": N.REWARD_HALF_CHEETAH, "AntEnv": N.REWARD_ANT, "Arm7DofEnv": N.REWARD_ARM,
}


class SyntheticEnv(object):
    def __init__(self, name="half_cheetah"):
        d, a, lim, dt, kind = ENV_SPECS[name]
        self.name = name
        self.action_space = Box(-lim * np.ones(a), lim * np.ones(a))
        self.observation_space = Box(-np.inf * np.ones(d), np.inf * np.ones(d))
        self.dt = dt
        self.l2a_reward_kind = kind

    def reward(self, obs, action, next_obs):
        assert obs.ndim == 2 and obs.shape == next_obs.shape and obs.shape[0] == action.shape[0]
        if self.l2a_reward_kind == N.REWARD_HALF_CHEETAH:      # envs/half_cheetah_env.py:58-65
            return (next_obs[:, -3] - obs[:, -3]) / self.dt - 0.05 * np.sum(np.square(action), axis=1)
        if self.l2a_reward_kind == N.REWARD_ANT:               # envs/ant_env.py:56-66
            return (next_obs[:, -3] - obs[:, -3]) / self.dt + 0.05
        return -np.linalg.norm(next_obs[:, -3:], axis=1) - 0.005 * np.sum(np.square(action), axis=1)  # arm_7dof_env.py:91-99


class LinearWorldEnv(SyntheticEnv):
    """A steppable stand-in: obs' = obs + 5 dt (obs @ A_task + act @ B), a random stable linear system whose per-task perturbation
    of A plays the role of the crippled leg / changed terrain of the paper's envs.  Gym-style reset() / step() -> (obs, reward,
    done, info) like the reference envs (envs/half_cheetah_env.py:39-56); obs[-3] is the coordinate the reward differentiates.
    It only produces transitions for the sampler: planning never steps it."""

    def __init__(self, name="half_cheetah", seed=0):
        SyntheticEnv.__init__(self, name)
        self.rng = np.random.RandomState(seed)
        D, A = self.observation_space.shape[0], self.action_space.shape[0]
        self.A0 = -0.5 * np.eye(D) + 0.1 * self.rng.normal(size=(D, D))
        self.B = 0.5 * self.rng.normal(size=(A, D))
        self.B[:, -3] += 1.0
        self.reset_task()
        self.obs = np.zeros(D)

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def reset_task(self, value=None):
        self.A = self.A0 + 0.05 * self.rng.normal(size=self.A0.shape)

    def reset(self):
        self.obs = 0.1 * self.rng.normal(size=self.A.shape[0])
        return self.obs.copy()

    def step(self, action):
        action = np.asarray(action, np.float64).reshape(-1)
        nxt = self.obs + self.dt * 5.0 * (self.obs @ self.A + action @ self.B)
        r = float(self.reward(self.obs[None], action[None], nxt[None])[0])
        self.obs = nxt
        return nxt.copy(), r, False, {}


def reward_kind_of(env):
    """Reward family + dt of an env object (ours or one of the reference's classes)."""
    kind = getattr(env, "l2a_reward_kind", None)
    if kind is None:
        kind = REWARD_KIND_BY_CLASS.get(type(env).__name__)
    if kind is None:
        raise NotImplementedError(
            "no fused reward kernel for env class %s: the B200 planner implements the HalfCheetah / Ant / Arm7Dof "
            "closed forms (set env.l2a_reward_kind to pick one); there is no CPU fallback" % type(env).__name__)
    return int(kind), float(env.dt)
