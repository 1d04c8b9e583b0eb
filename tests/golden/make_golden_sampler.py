"""Golden fixture for the sampler-side adaptation windows, produced by the REFERENCE'S OWN
``Sampler.obtain_samples`` (learning_to_adapt/samplers/sampler.py), imported unmodified from /root/reference.

Run in the authoring container only:    python tests/golden/make_golden_sampler.py

Stand-ins around the upstream sampler: tensorflow / gym / pyprind modules are inert stubs (the sampler only constructs a
progress bar), the env is this repo's MuJoCo-free ``LinearWorldEnv`` (the sampler just steps it), and the policy is a
deterministic closed-form stub whose ``dynamics_model`` records what the sampler hands to ``adapt`` at every env step.
The fixture stores those recorded windows, the step at which each was taken, and the returned paths;
tests pin this repo's ``Sampler`` (host 'lists' mode on CPU; device-window mode on the GPU) to them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from make_golden import _Anything, REF  # noqa: E402
from tests.sampler_stubs import RecordingModel, StubPolicy, make_env  # noqa: E402


def main():
    sys.modules["tensorflow"] = _Anything("tensorflow")
    for name in ("gym", "gym.spaces", "pyprind", "mpi4py"):
        sys.modules.setdefault(name, _Anything(name))
    sys.path.insert(0, REF)
    from learning_to_adapt.samplers.sampler import Sampler

    out = {}
    for tag, num_envs, path_len, M, rounds in (("a", 3, 12, 4, 2), ("b", 2, 40, 16, 1)):
        env = make_env()
        model = RecordingModel()
        policy = StubPolicy(env, model)
        sampler = Sampler(env=env, policy=policy, num_rollouts=num_envs, max_path_length=path_len, adapt_batch_size=M)
        for i, e in enumerate(sampler.vec_env.envs):
            e.seed(100 + i)
        sampler.total_samples = rounds * num_envs * path_len        # envs auto-reset and keep going: `rounds` paths per env
        paths = sampler.obtain_samples()
        out["%s_meta" % tag] = np.array([num_envs, path_len, M, rounds], np.int64)
        out["%s_adapt_step" % tag] = np.array(model.steps, np.int64)
        out["%s_adapt_obs" % tag] = np.stack(model.obs)               # [n_adapt, envs, M, D]
        out["%s_adapt_act" % tag] = np.stack(model.act)
        out["%s_adapt_next" % tag] = np.stack(model.nxt)
        out["%s_n_pre_adapt" % tag] = np.array([model.n_switch], np.int64)
        out["%s_path_obs" % tag] = np.stack([p["observations"] for p in paths])
        out["%s_path_act" % tag] = np.stack([p["actions"] for p in paths])
        out["%s_path_rew" % tag] = np.stack([p["rewards"] for p in paths])
        out["%s_path_done" % tag] = np.stack([p["dones"] for p in paths])
        print(tag, "adapt calls", len(model.steps), "paths", len(paths))
    path = os.path.join(HERE, "reference_sampler_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
