"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the candidate-sampling stream of the reference planner, in the decomposition
the device kernels use (learning_to_adapt_b200/csrc/mt19937.cuh).  Nothing in the product imports this file.

The algorithm lives in a third-party dependency of the reference, not under /root/reference:
    numpy (requirements.txt:33 pins numpy==1.15.1) -- the process-wide legacy generator ``np.random`` (RandomState):
    MT19937 (Matsumoto & Nishimura 1998), ``random_sample`` = 53-bit doubles from two 32-bit outputs, ``uniform`` =
    low + (high - low) * random_sample, ``normal`` = the polar Box-Muller method with a cached second value (legacy_gauss).
    NEP 19 froze this stream: the numpy on this box (2.x) produces the same values from the same state as 1.15.1.
Reference call sites: ``get_random_action`` -> np.random.uniform(low, high, (H*N*m, A)) (policies/mpc_controller.py:67-69, 114),
CEM -> np.random.normal(size=(n, m, H*A)) (policies/mpc_controller.py:85).

Pinned by tests/test_mt19937_oracle_cpu.py against numpy itself (values AND the generator state left behind, bit for bit).
The restatement mirrors the kernels' parallel formulation, not numpy's sequential loops:
    raw_blocks     -- the linear recurrence advanced a 624-word block at a time by 227 independent three-element chains
                      (+ element 623 on its own), as mt19937_raw_kernel does
    uniform        -- every element from its own two stream words, as mt19937_uniform_kernel does
    legacy_normal  -- every polar attempt in parallel from its own four stream words, accepted attempts compacted in stream
                      order by a prefix sum, the incoming cached value first, an odd draw leaving f*x1 cached, as
                      mt19937_gauss_kernel / _count / _scan / _scatter do
    state_after    -- (key, pos) after a number of consumed words with numpy's convention pos in [0, 624], as
                      mt19937_state_out_kernel does
"""
import math

import numpy as np

N, M = 624, 397
U32 = np.uint32


def _twist(cur, nxt, far):
    """x[k+624] = x[k+397] ^ ((upper(x[k]) | lower(x[k+1])) >> 1) ^ (odd ? 0x9908b0df : 0)."""
    y = (cur & U32(0x80000000)) | (nxt & U32(0x7FFFFFFF))
    return far ^ (y >> U32(1)) ^ np.where(y & U32(1), U32(0x9908B0DF), U32(0)).astype(U32)


def temper(y):
    y = y.astype(U32)
    y = y ^ (y >> U32(11))
    y = y ^ ((y << U32(7)) & U32(0x9D2C5680))
    y = y ^ ((y << U32(15)) & U32(0xEFC60000))
    return y ^ (y >> U32(18))


def next_block(old):
    """One refill of the 624-word state in the kernel's decomposition: chain k < 227 produces elements k, k+227, k+454 (the
    third only for k < 169); every input is an OLD word except the chain's own previous result; element 623 needs new[0] and
    new[396], which are recomputed from old words (new[396] through new[169])."""
    old = np.asarray(old, U32)
    new = np.empty(N, U32)
    k = np.arange(227)
    n0 = _twist(old[k], old[k + 1], old[k + M])                    # element k      : old[k], old[k+1], old[k+397]
    n1 = _twist(old[k + 227], old[k + 228], n0)                    # element k+227  : old[k+227], old[k+228], new[k]
    kc = np.minimum(k + 454, N - 2)
    n2 = _twist(old[kc], old[kc + 1], n1)                          # element k+454  : old[k+454], old[k+455], new[k+227]
    new[k] = n0
    new[k + 227] = n1
    new[k[:169] + 454] = n2[:169]
    one = lambda i: old[i:i + 1]
    new0 = _twist(one(0), one(1), one(M))
    new169 = _twist(one(169), one(170), one(169 + M))
    new396 = _twist(one(396), one(397), new169)
    new[623] = _twist(one(623), new0, new396)[0]
    return new


def next_block_sequential(old):
    """The textbook loop (mt19937_gen), for cross-checking the decomposition above."""
    x = [int(v) for v in old]
    for i in range(N):
        y = (x[i] & 0x80000000) | (x[(i + 1) % N] & 0x7FFFFFFF)
        x[i] = x[(i + M) % N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
    return np.array(x, U32)


def raw_blocks(key, pos, n_words):
    """Blocks of 624 RAW words: block 0 = the incoming key, block b = the b-th refill; stream word w from here is
    temper(raw[pos + w]).  ceil((pos + n_words) / 624) blocks, block 0 included (at least one)."""
    key = np.asarray(key, U32)
    nblocks = max(1, -(-(int(pos) + int(n_words)) // N))
    out = np.empty((nblocks, N), U32)
    out[0] = key
    for b in range(1, nblocks):
        out[b] = next_block(out[b - 1])
    return out.reshape(-1)


def doubles(raw, pos, count):
    """mt19937_next_double for `count` consecutive draws: a = w0 >> 5, b = w1 >> 6, (a * 2^26 + b) / 2^53."""
    w = temper(raw[pos:pos + 2 * count]).reshape(count, 2)
    a = (w[:, 0] >> U32(5)).astype(np.float64)
    b = (w[:, 1] >> U32(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


def state_after(raw, pos, words):
    """(key, pos) after `words` stream words; numpy's convention: pos = 624 means "block exhausted, refill on the next draw"."""
    if words == 0:
        return raw[:N].copy(), int(pos)
    total = int(pos) + int(words)
    blk, p = divmod(total, N)
    if p == 0:
        blk, p = blk - 1, N
    return raw[blk * N:(blk + 1) * N].copy(), p


def uniform(key, pos, low, high, rows):
    """np.random.uniform(low, high, (rows, A)) from the state (key, pos): values [rows, A] float64, new (key, pos)."""
    low = np.asarray(low, np.float64)
    rng = np.asarray(high, np.float64) - low
    a = low.shape[0]
    total = rows * a
    raw = raw_blocks(key, pos, 2 * total)
    d = doubles(raw, pos, total).reshape(rows, a)
    vals = low + rng * d                                              # separate multiply and add roundings (no fma)
    return vals, state_after(raw, pos, 2 * total)


def legacy_normal(key, pos, has_gauss, cached, count, attempts=None):
    """np.random.normal(size=count) from the state (key, pos, has_gauss, cached): values, new (key, pos, has_gauss, cached).
    `attempts`: the attempt budget generated in parallel (default: enough with a wide margin); the draw consumes only the words
    of the attempts up to the last one it needed."""
    carry = 1 if has_gauss else 0
    z = np.empty(count, np.float64)
    if carry and count > 0:
        z[0] = cached
    need = count - carry
    pairs_needed = (need + 1) // 2 if need > 0 else 0
    if pairs_needed == 0:
        keep = carry if count == 0 else 0
        return z, (np.asarray(key, U32).copy(), int(pos), keep, float(cached) if keep else 0.0)
    if attempts is None:
        attempts = int(pairs_needed * 1.35) + 64                      # acceptance rate pi / 4
    raw = raw_blocks(key, pos, 4 * attempts)
    d = doubles(raw, pos, 2 * attempts).reshape(attempts, 2)
    x1 = 2.0 * d[:, 0] - 1.0
    x2 = 2.0 * d[:, 1] - 1.0
    r2 = x1 * x1 + x2 * x2
    ok = (r2 < 1.0) & (r2 != 0.0)
    rank = np.cumsum(ok) - ok                                         # accepted attempts before attempt i (exclusive scan)
    assert int(ok.sum()) >= pairs_needed, "attempt budget too small"
    take = ok & (rank < pairs_needed)
    idx = np.nonzero(take)[0]
    # C libm log / sqrt, one value at a time (numpy's vectorised log is a different implementation)
    f = np.array([math.sqrt(-2.0 * math.log(r) / r) for r in r2[idx]], np.float64)
    first, second = f * x2[idx], f * x1[idx]                          # the generator returns f*x2, caches f*x1
    o = carry + 2 * np.arange(pairs_needed)
    z[o] = first
    fits = o + 1 < count
    z[o[fits] + 1] = second[fits]
    leftover = need & 1
    consumed = int(idx[-1]) + 1
    k, p = state_after(raw, pos, 4 * consumed)
    return z, (k, p, int(leftover), float(second[-1]) if leftover else 0.0)
