"""Host-side logic of the candidate shard (N > 1 path) on CPU: world_size-2 gloo, no GPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from learning_to_adapt_b200.parallel import CandidateShard, pack_best, select_best, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 500, 2000, 32768):
        for g in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_select_best_has_numpy_argmax_semantics():
    rng = np.random.RandomState(0)
    G, m, n_loc, A = 4, 6, 5, 3
    returns = rng.normal(size=(m, G * n_loc)).astype(np.float32)
    returns[1, 7] = returns[1, 13] = returns[1].max() + 1.0          # tie across ranks -> lowest global index
    returns[2, 11] = np.nan                                           # NaN wins
    returns[3, 4] = np.nan
    returns[3, 17] = np.nan                                           # first NaN wins
    acts = rng.normal(size=(m, G * n_loc, A)).astype(np.float32)
    packed = []
    for g in range(G):
        sl = slice(g * n_loc, (g + 1) * n_loc)
        loc = returns[:, sl]
        idx = np.array([int(np.argmax(loc[e])) for e in range(m)])
        packed.append(pack_best(torch.tensor(loc[range(m), idx]), torch.tensor(idx + g * n_loc),
                                torch.tensor(acts[range(m), idx + g * n_loc])))
    ret, idx, act = select_best(torch.stack(packed))
    want = np.array([int(np.argmax(returns[e])) for e in range(m)])
    np.testing.assert_array_equal(idx.numpy(), want)
    np.testing.assert_array_equal(act.numpy(), acts[range(m), want])
    np.testing.assert_array_equal(np.isnan(ret.numpy()), np.isnan(returns[range(m), want]))


def test_pack_best_large_indices_exact():
    idx = torch.tensor([0, 65535, 65536, 2 ** 24 + 1, 2 ** 30 + 12345])
    p = pack_best(torch.zeros(5), idx, torch.zeros(5, 2))
    _, got, _ = select_best(p.unsqueeze(0))
    assert got.tolist() == idx.tolist()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = CandidateShard()
        rng = np.random.RandomState(5)                     # same stream on both ranks
        m, n, A = 3, 11, 2
        returns = rng.normal(size=(m, n)).astype(np.float32)
        returns[0, 2] = returns[0, 9] = 10.0                # tie: rank 0's candidate must win
        acts = rng.normal(size=(m, n, A)).astype(np.float32)
        lo, hi = shard_bounds(n, rank, world)
        loc = returns[:, lo:hi]
        idx = np.array([int(np.argmax(loc[e])) for e in range(m)])
        ret, gidx, act = shard.combine(torch.tensor(loc[range(m), idx]), torch.tensor(idx, dtype=torch.int32),
                                       torch.tensor(acts[range(m), lo + idx]), lo)
        want = np.argmax(returns, axis=1)
        ok = (gidx.numpy() == want).all() and np.array_equal(act.numpy(), acts[range(m), want]) and \
            np.array_equal(ret.numpy(), returns[range(m), want])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_combine_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
