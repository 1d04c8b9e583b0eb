"""Training of the dynamics models ON the engine's resident parameters: Adam regression for
MLPDynamicsModel.fit (dynamics/mlp_dynamics.py:91-202) and the MAML outer loop for MetaMLPDynamicsModel.fit
(dynamics/meta_mlp_dynamics.py:96-140, 167-274).

The weights never leave the device: `PlanningEngine.param_views` hands out torch views onto the fp32 parameter block the
planning kernels read, forward / backward run in torch autograd on those views, the optimiser updates them in place, and
`refresh_sets` re-tiles them for the tensor-core rollout.  Everything the reference draws from numpy's global stream is drawn
from it here in the same order (train / validation split :119-120 via np.random.shuffle, MAML windows via np.random.randint,
meta_mlp_dynamics.py:353-383); the one thing that cannot be pinned is the order in which TF's `dataset.shuffle` visits the
batches (mlp_dynamics.py:234-236) -- batches are the same consecutive slices, visited in an np.random permutation.
The optimiser is TensorFlow's Adam update, not torch.optim.Adam (they place epsilon differently):
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  theta -= lr_t * m / (sqrt(v) + eps)
with the slot variables living as long as the model (the TF session keeps them across fit() calls).
"""
import time
from collections import OrderedDict

import numpy as np
import torch

ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8          # tf.train.AdamOptimizer defaults (mlp_dynamics.py:34)


def _to_params(np_params, device):
    return [torch.tensor(v, dtype=torch.float32, device=device, requires_grad=True) for v in np_params.values()]


def _forward(x, params):
    h = x
    n_layers = len(params) // 2
    for l in range(n_layers):
        h = h @ params[2 * l] + params[2 * l + 1]
        if l < n_layers - 1:
            h = torch.relu(h)
    return h


def train_test_split(obs, act, delta, test_split_ratio=0.2):
    """mlp_dynamics.py:273-285 -- same draw from numpy's global stream."""
    assert obs.shape[0] == act.shape[0] == delta.shape[0]
    dataset_size = obs.shape[0]
    indices = np.arange(dataset_size)
    np.random.shuffle(indices)
    split_idx = int(dataset_size * (1 - test_split_ratio))
    tr, te = indices[:split_idx], indices[split_idx:]
    return obs[tr], act[tr], delta[tr], obs[te], act[te], delta[te]


def _early_stop_state(valid_loss):
    # mlp_dynamics.py:171-176: start the rolling average above the first validation loss
    if valid_loss < 0:
        return valid_loss / 1.5, valid_loss / 2
    return 1.5 * valid_loss, 2 * valid_loss


class AdamState(object):
    """Slot variables of tf.train.AdamOptimizer for a list of parameter tensors."""

    def __init__(self, params):
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def step(self, params, grads, lr):
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - ADAM_B2 ** self.t) / (1.0 - ADAM_B1 ** self.t)
        with torch.no_grad():
            torch._foreach_mul_(self.m, ADAM_B1)
            torch._foreach_add_(self.m, grads, alpha=1.0 - ADAM_B1)
            torch._foreach_mul_(self.v, ADAM_B2)
            torch._foreach_addcmul_(self.v, grads, grads, value=1.0 - ADAM_B2)
            denom = torch._foreach_sqrt(self.v)
            torch._foreach_add_(denom, ADAM_EPS)
            torch._foreach_addcdiv_(params, self.m, denom, value=-float(lr_t))


def fit_mlp(params, adam, train, test, epochs, batch_size, learning_rate, rolling_average_persitency, verbose=False):
    """params: leaf tensors (views onto the engine's parameters, updated IN PLACE); train / test: dicts of device tensors
    x [n, D+A], y [n, D] (the aggregated, normalised datasets).  Returns dict(epochs, epoch_times, train_loss, valid_loss)."""
    xt, yt, xv, yv = train["x"], train["y"], test["x"], test["y"]
    avg = prev = None
    epoch_times, epoch = [], 0
    batch_losses = []
    for epoch in range(epochs):
        t0 = time.time()
        starts = np.arange(0, xt.shape[0], batch_size)             # dataset.batch(batch_size) ...
        starts = starts[np.random.permutation(len(starts))]        # ... .shuffle(): the batches are shuffled, not the samples
        batch_losses = []
        for s in starts:
            loss = torch.mean((yt[s:s + batch_size] - _forward(xt[s:s + batch_size], params)) ** 2)       # :82
            grads = torch.autograd.grad(loss, params)
            adam.step(params, grads, learning_rate)
            batch_losses.append(loss.detach())
        with torch.no_grad():
            valid_loss = float(torch.mean((yv - _forward(xv, params)) ** 2)) if xv.shape[0] else float(batch_losses[-1])
        if avg is None:
            avg, prev = _early_stop_state(valid_loss)
        avg = rolling_average_persitency * avg + (1.0 - rolling_average_persitency) * valid_loss
        epoch_times.append(time.time() - t0)
        if verbose:
            print("Training DynamicsModel - finished epoch %i -- train loss: %.4f  valid loss: %.4f  valid_loss_mov_avg: %.4f  epoch time: %.2f"
                  % (epoch, float(torch.stack(batch_losses).mean()), valid_loss, avg, epoch_times[-1]))
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return dict(epochs=epoch, epoch_times=epoch_times, train_loss=float(torch.stack(batch_losses).mean()) if batch_losses else float("nan"),
                valid_loss=valid_loss)


def maml_get_batch(data, meta_batch_size, batch_size):
    """meta_mlp_dynamics.py:353-383 -- same two np.random.randint draws; returns index arrays (paths, centres)."""
    num_paths, len_path = data["x"].shape[:2]
    idx_path = np.random.randint(0, num_paths, size=meta_batch_size)
    idx_batch = np.random.randint(batch_size, len_path - batch_size, size=meta_batch_size)
    return idx_path, idx_batch


def maml_losses(params, data, idx_path, idx_batch, batch_size, inner_learning_rate, create_graph=True):
    """(pre_loss, post_loss) of one meta batch: per task, a window of 2*batch_size steps of one path, first half adapts
    (one SGD step on the pre loss, _adapt_sym :409-421), second half evaluates the adapted parameters (:96-140)."""
    pre, post = [], []
    for ip, ib in zip(idx_path, idx_batch):
        xw, yw = data["x"][ip, ib - batch_size:ib + batch_size], data["y"][ip, ib - batch_size:ib + batch_size]
        xp, yp, xq, yq = xw[:batch_size], yw[:batch_size], xw[batch_size:], yw[batch_size:]
        pre_loss = torch.mean((yp - _forward(xp, params)) ** 2)
        grads = torch.autograd.grad(pre_loss, params, create_graph=create_graph)
        adapted = [w - inner_learning_rate * g for w, g in zip(params, grads)]
        post.append(torch.mean((yq - _forward(xq, adapted)) ** 2))
        pre.append(pre_loss)
    return torch.stack(pre).mean(), torch.stack(post).mean()


def fit_maml(params, adam, train, test, epochs, batch_size, meta_batch_size, learning_rate, inner_learning_rate,
             rolling_average_persitency, verbose=False):
    """params: leaf tensors updated in place; train / test: dicts of device tensors x [paths, T, D+A], y [paths, T, D]."""
    n_tr = int(np.prod(train["x"].shape[:2]))
    n_te = int(np.prod(test["x"].shape[:2]))
    steps_train = max(int(n_tr / (meta_batch_size * batch_size * 2)), 1)      # :210-213
    steps_test = max(int(n_te / (meta_batch_size * batch_size * 2)), 1)
    avg = prev = None
    epoch_times, epoch = [], 0
    pre_losses, post_losses = [], []
    valid_loss = float("nan")
    for epoch in range(epochs):
        t0 = time.time()
        pre_losses, post_losses = [], []
        for _ in range(steps_train):
            ip, ib = maml_get_batch(train, meta_batch_size, batch_size)
            pre_loss, post_loss = maml_losses(params, train, ip, ib, batch_size, inner_learning_rate)
            grads = torch.autograd.grad(post_loss, params)                      # train_op minimises the post-update loss (:139)
            adam.step(params, grads, learning_rate)
            pre_losses.append(pre_loss.detach())
            post_losses.append(post_loss.detach())
        valid_losses = []
        with torch.no_grad():
            for _ in range(steps_test):
                ip, ib = maml_get_batch(test, meta_batch_size, batch_size)
                xs = torch.cat([test["x"][p, b - batch_size:b + batch_size] for p, b in zip(ip, ib)])
                ys = torch.cat([test["y"][p, b - batch_size:b + batch_size] for p, b in zip(ip, ib)])
                valid_losses.append(float(torch.mean((ys - _forward(xs, params)) ** 2)))     # self.loss: plain MSE of theta (:242)
        valid_loss = float(np.mean(valid_losses))
        if avg is None:
            avg, prev = _early_stop_state(valid_loss)
        avg = rolling_average_persitency * avg + (1.0 - rolling_average_persitency) * valid_loss
        epoch_times.append(time.time() - t0)
        if verbose:
            print("Training DynamicsModel - finished epoch %i - train loss: %.4f   valid loss: %.4f   valid_loss_mov_avg: %.4f   epoch time: %.2f"
                  % (epoch, float(torch.stack(post_losses).mean()), valid_loss, avg, epoch_times[-1]))
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return dict(epochs=epoch, epoch_times=epoch_times, pre_loss=float(torch.stack(pre_losses).mean()),
                post_loss=float(torch.stack(post_losses).mean()), valid_loss=valid_loss)


def _split(n, ratio, rng):
    idx = np.arange(n)
    rng.shuffle(idx)
    k = int(n * (1 - ratio))
    return idx[:k], idx[k:]


def _lstm_forward(x, c, h, params):
    """x [B, T, in]; TF LSTMCell semantics (gate order i, j, f, o; forget_bias 1)."""
    wk, bk, wo, bo = params
    hs = h.shape[1]
    ys = []
    for t in range(x.shape[1]):
        z = torch.cat([x[:, t], h], dim=1) @ wk + bk
        i, j, f, o = z[:, :hs], z[:, hs:2 * hs], z[:, 2 * hs:3 * hs], z[:, 3 * hs:]
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        h = torch.sigmoid(o) * torch.tanh(c)
        ys.append(h @ wo + bo)
    return torch.stack(ys, dim=1), c, h


def fit_lstm(np_params, obs_n, act_n, delta_n, epochs, batch_size, learning_rate, backprop_steps, valid_split_ratio,
             rolling_average_persitency, device, verbose=False, seed=None):
    """Truncated BPTT (chunks of `backprop_steps`, state carried and detached between chunks) with Adam on
    mean((delta - f(x))^2), validation early stop -- the procedure of rnn_dynamics.py:95-231.  Split and batch order come from
    numpy's global stream unless a seed is given."""
    rng = np.random.RandomState(seed if seed is not None else np.random.randint(2 ** 31 - 1))
    x = torch.tensor(np.concatenate([obs_n, act_n], axis=2).astype(np.float32), device=device)
    y = torch.tensor(delta_n.astype(np.float32), device=device)
    n_paths, T = x.shape[0], x.shape[1]
    tr, te = _split(n_paths, valid_split_ratio, rng)
    if len(te) == 0:
        te = tr
    params = _to_params(np_params, device)
    hs = params[2].shape[0]
    opt = torch.optim.Adam(params, lr=learning_rate)
    avg = prev = None
    for epoch in range(epochs):
        order = rng.permutation(tr)
        for s in range(0, len(order), batch_size):
            idx = torch.as_tensor(order[s:s + batch_size], device=device)
            c = torch.zeros(len(idx), hs, device=device)
            h = torch.zeros(len(idx), hs, device=device)
            for t0 in range(0, T, backprop_steps):
                opt.zero_grad()
                pred, c, h = _lstm_forward(x[idx, t0:t0 + backprop_steps], c, h, params)
                loss = torch.mean((pred - y[idx, t0:t0 + backprop_steps]) ** 2)
                loss.backward()
                opt.step()
                c, h = c.detach(), h.detach()
        with torch.no_grad():
            idx = torch.as_tensor(te, device=device)
            pred, _, _ = _lstm_forward(x[idx], torch.zeros(len(idx), hs, device=device), torch.zeros(len(idx), hs, device=device), params)
            valid_loss = float(torch.mean((pred - y[idx]) ** 2))
        if avg is None:
            avg, prev = _early_stop_state(valid_loss)
        avg = rolling_average_persitency * avg + (1.0 - rolling_average_persitency) * valid_loss
        if verbose:
            print("fit_lstm epoch %d valid %.5f avg %.5f" % (epoch, valid_loss, avg))
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return OrderedDict((k, p.detach().cpu().numpy()) for k, p in zip(np_params.keys(), params))
