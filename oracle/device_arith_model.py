"""TEST INFRASTRUCTURE ONLY -- a numpy MODEL of the arithmetic the fused rollout kernels run (not of the reference): where the
reference carries the planner state in float64 and multiplies in float32 (policies/mpc_controller.py:116-127,
dynamics/mlp_dynamics.py:204-222), the device carries the state, the normalisation and the return in float32 and forms every
dense product on the tensor cores from bf16 halves,
        x @ W  ~=  x_hi @ W_hi + x_hi @ W_lo + x_lo @ W_hi        (x = x_hi + x_lo, W = W_hi + W_lo, fp32 accumulate; the
                                                                    x_lo @ W_lo term, ~2^-16 relative, is dropped)
(learning_to_adapt_b200/csrc/rollout_tc.cuh, rollout_tc2.cuh: "split-bf16").  The model answers on the CPU the question the GPU
parity tests answer by measurement: how far can that choreography drift from the reference-faithful oracle (oracle/mpc_oracle.py)
over H dependent steps, and is north_star's 1e-4 relative bar safe at the BASELINE shapes (Ant's +-150 actions, H = 25)?
It is an error-budget model: accumulation ORDER inside a product differs from the tensor core's, so it predicts magnitudes, not
bits.  Nothing in the product imports this file; tests/test_device_arith_model_cpu.py uses it.
"""
import numpy as np

from oracle import mpc_oracle as O

F32 = np.float32


def to_bf16(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    r = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return (r & np.uint32(0xFFFF0000)).view(F32)


def split(x):
    hi = to_bf16(x)
    return hi, to_bf16(np.asarray(x, F32) - hi)


def dense_split(x, w, b, passes=3):
    """One dense layer the way the kernels form it; passes = 1 keeps only x_hi @ W_hi (a plain bf16 GEMM, for contrast)."""
    xh, xl = split(x)
    wh, wl = split(w)
    y = xh @ wh
    if passes == 3:
        y = y + xh @ wl + xl @ wh
    return (y + b).astype(F32)


def mlp_forward_split(x32, params, passes=3):
    keys = list(params.keys())
    n_layers = len(keys) // 2
    h = np.ascontiguousarray(x32, dtype=F32)
    for l in range(n_layers):
        h = dense_split(h, params[keys[2 * l]], params[keys[2 * l + 1]], passes)
        if l < n_layers - 1:
            h = np.maximum(h, F32(0.0))
    return h


def rollout_returns_model(observations, actions, param_sets, norm, reward_kind, dt, discount=1.0, mode="shared", passes=3):
    """The oracle's rollout_returns (same arguments, same [m, n] result) in the device's choreography: float32 state, reciprocal
    normalisation constants, split-bf16 products, ensemble mean of the denormalised deltas, float32 reward and return."""
    obs0 = np.asarray(observations, np.float64)
    acts = np.asarray(actions, np.float64).astype(F32)            # the candidate tensor is float32 on the device
    h, rows, _ = acts.shape
    m = obs0.shape[0]
    n = rows // m
    mu_o, sd_o = [np.asarray(v, np.float64) for v in norm["obs"]]
    mu_a, sd_a = [np.asarray(v, np.float64) for v in norm["act"]]
    mu_d, sd_d = [np.asarray(v, np.float64) for v in norm["delta"]]
    inv_o = (F32(1.0) / (sd_o + O.EPS).astype(F32)).astype(F32)
    inv_a = (F32(1.0) / (sd_a + O.EPS).astype(F32)).astype(F32)
    mu_o32, mu_a32, mu_d32, sc_d32 = mu_o.astype(F32), mu_a.astype(F32), mu_d.astype(F32), (sd_d + O.EPS).astype(F32)
    s = np.repeat(obs0, n, axis=0).astype(F32)
    ret = np.zeros(rows, F32)
    disc = F32(1.0)
    for t in range(h):
        x = np.concatenate([(s - mu_o32) * inv_o, (acts[t] - mu_a32) * inv_a], axis=1).astype(F32)
        if mode == "shared":
            delta = mlp_forward_split(x, param_sets[0], passes) * sc_d32 + mu_d32
        elif mode == "per_env":
            chunks = np.split(x, len(param_sets), axis=0)
            delta = np.concatenate([mlp_forward_split(c, p, passes) for c, p in zip(chunks, param_sets)], axis=0) * sc_d32 + mu_d32
        elif mode == "ensemble":
            acc = np.zeros_like(s)
            for p in param_sets:
                acc = acc + (mlp_forward_split(x, p, passes) * sc_d32 + mu_d32)
            delta = acc * F32(1.0 / len(param_sets))
        else:
            raise ValueError(mode)
        nxt = (s + delta.astype(F32)).astype(F32)
        asq = np.sum(acts[t] * acts[t], axis=1, dtype=F32)
        if reward_kind == O.REWARD_HALF_CHEETAH:
            r = (nxt[:, -3] - s[:, -3]) / F32(dt) - F32(0.05) * asq
        elif reward_kind == O.REWARD_ANT:
            r = (nxt[:, -3] - s[:, -3]) / F32(dt) + F32(0.05)
        else:
            r = -np.sqrt(np.sum(nxt[:, -3:] * nxt[:, -3:], axis=1, dtype=F32)) - F32(0.005) * asq
        ret = (ret + disc * r.astype(F32)).astype(F32)
        disc = F32(disc * F32(discount))
        s = nxt
    return ret.astype(np.float64).reshape(m, n)
