"""Host-side training glue (off the planning hot path): Adam regression for MLPDynamicsModel.fit
(dynamics/mlp_dynamics.py:91-202) and the MAML outer loop for MetaMLPDynamicsModel.fit
(dynamics/meta_mlp_dynamics.py:96-140, 167-274), in torch autograd on the engine's device.

These restate the reference's training procedure so the dynamics models stay usable end to end; they are not kernels of
this build (SURVEY.md 8(f) row f2 -- "next").  Inputs are already normalised by the caller, like the TF feed.
"""
from collections import OrderedDict

import numpy as np
import torch


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


def _split(n, ratio, rng):
    idx = np.arange(n)
    rng.shuffle(idx)
    k = int(n * (1 - ratio))
    return idx[:k], idx[k:]


def _early_stop_state(valid_loss):
    # mlp_dynamics.py:171-176: start the rolling average above the first validation loss
    if valid_loss < 0:
        return valid_loss / 1.5, valid_loss / 2
    return 1.5 * valid_loss, 2 * valid_loss


def fit_mlp(np_params, obs_n, act_n, delta_n, epochs, batch_size, learning_rate, valid_split_ratio,
            rolling_average_persitency, device, verbose=False, seed=0):
    rng = np.random.RandomState(seed)
    x = np.concatenate([obs_n, act_n], axis=1).astype(np.float32)
    y = delta_n.astype(np.float32)
    tr, te = _split(x.shape[0], valid_split_ratio, rng)
    xt, yt = torch.tensor(x[tr], device=device), torch.tensor(y[tr], device=device)
    xv, yv = torch.tensor(x[te], device=device), torch.tensor(y[te], device=device)
    params = _to_params(np_params, device)
    opt = torch.optim.Adam(params, lr=learning_rate)
    avg = prev = None
    for epoch in range(epochs):
        # the reference batches then shuffles the batches (mlp_dynamics.py:234-236)
        starts = np.arange(0, xt.shape[0], batch_size)
        rng.shuffle(starts)
        for s in starts:
            opt.zero_grad()
            loss = torch.mean((yt[s:s + batch_size] - _forward(xt[s:s + batch_size], params)) ** 2)
            loss.backward()
            opt.step()
        with torch.no_grad():
            valid_loss = float(torch.mean((yv - _forward(xv, params)) ** 2)) if xv.shape[0] else float(loss)
        if avg is None:
            avg, prev = _early_stop_state(valid_loss)
        avg = rolling_average_persitency * avg + (1.0 - rolling_average_persitency) * valid_loss
        if verbose:
            print("fit_mlp epoch %d valid %.5f avg %.5f" % (epoch, valid_loss, avg))
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return OrderedDict((k, p.detach().cpu().numpy()) for k, p in zip(np_params.keys(), params))


def fit_maml(np_params, obs_n, act_n, delta_n, epochs, batch_size, meta_batch_size, learning_rate, inner_learning_rate,
             valid_split_ratio, rolling_average_persitency, device, verbose=False, seed=0):
    """obs_n/act_n/delta_n: [n_paths, T, dim] normalised.  Each meta-batch element is a 2*batch_size window of one path:
    first half = adaptation data, second half = evaluation data (meta_mlp_dynamics.py:353-383); the outer loss is the mean
    post-update MSE over the meta batch (:96-140), differentiated through the one-step inner update (second order)."""
    rng = np.random.RandomState(seed)
    n_paths, T = obs_n.shape[0], obs_n.shape[1]
    bs = min(batch_size, T // 2)
    assert bs >= 1, "paths are too short for a (pre, post) window"
    x = torch.tensor(np.concatenate([obs_n, act_n], axis=2).astype(np.float32), device=device)
    y = torch.tensor(delta_n.astype(np.float32), device=device)
    tr, te = _split(n_paths, valid_split_ratio, rng)
    if len(te) == 0:
        te = tr
    params = _to_params(np_params, device)
    opt = torch.optim.Adam(params, lr=learning_rate)

    def meta_loss(path_ids):
        losses = []
        for p in path_ids:
            t0 = rng.randint(0, T - 2 * bs + 1)
            xp, yp = x[p, t0:t0 + bs], y[p, t0:t0 + bs]
            xq, yq = x[p, t0 + bs:t0 + 2 * bs], y[p, t0 + bs:t0 + 2 * bs]
            pre = torch.mean((yp - _forward(xp, params)) ** 2)
            grads = torch.autograd.grad(pre, params, create_graph=True)
            adapted = [w - inner_learning_rate * g for w, g in zip(params, grads)]          # _adapt_sym :409-421
            losses.append(torch.mean((yq - _forward(xq, adapted)) ** 2))
        return torch.stack(losses).mean()

    avg = prev = None
    n_batches = max(1, (len(tr) * (T // (2 * bs))) // meta_batch_size)
    for epoch in range(epochs):
        for _ in range(n_batches):
            opt.zero_grad()
            loss = meta_loss(rng.choice(tr, size=meta_batch_size, replace=len(tr) < meta_batch_size))
            loss.backward()
            opt.step()
        valid_loss = float(meta_loss(rng.choice(te, size=meta_batch_size, replace=len(te) < meta_batch_size)).detach())
        if avg is None:
            avg, prev = _early_stop_state(valid_loss)
        avg = rolling_average_persitency * avg + (1.0 - rolling_average_persitency) * valid_loss
        if verbose:
            print("fit_maml epoch %d valid %.5f avg %.5f" % (epoch, valid_loss, avg))
        if prev < avg or epoch == epochs - 1:
            break
        prev = avg
    return OrderedDict((k, p.detach().cpu().numpy()) for k, p in zip(np_params.keys(), params))


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
             rolling_average_persitency, device, verbose=False, seed=0):
    """Truncated BPTT (chunks of `backprop_steps`, state carried and detached between chunks) with Adam on
    mean((delta - f(x))^2), validation early stop -- the procedure of rnn_dynamics.py:95-231."""
    rng = np.random.RandomState(seed)
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
