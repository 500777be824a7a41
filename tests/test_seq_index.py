"""CPU: the device path's in-place minibatch index maps (harl_b200/common/seq_index.py) against the oracle's index
functions (oracle/buffers.py -- pinned to the unmodified reference's generators by the ha_train / hatrpo_train goldens,
SURVEY.md Appendix D) and against the buffers' reference-shaped generators, with a recorded permutation replayed.

A buffer whose every element encodes its own (t, n) turns "which rows did the generator gather" into data: the
kernels read row index[r] of the time-major flatten, so gathering the encoded buffer with ``index`` must reproduce the
generator's batch, and the hidden state of sequence j must be the one stored at buffer row index[j].
"""
import numpy as np
import pytest
import torch

from harl_b200.common import seq_index
from oracle import buffers as ob


def _replay(perm):
    orig = torch.randperm
    calls = []

    def fake(n, *a, **k):
        assert n == len(perm)
        calls.append(n)
        return torch.as_tensor(perm)

    return orig, fake, calls


@pytest.mark.parametrize("T,C,nmb,L", [(8, 6, 2, 4), (12, 5, 1, 4), (12, 10, 5, 3), (6, 4, 4, 2)])
def test_chunk_index_map_matches_reference_generator_semantics(T, C, nmb, L):
    rng = np.random.default_rng(T * 100 + C)
    chunks = T * C // L
    perm = rng.permutation(chunks)
    orig, fake, calls = _replay(perm)
    torch.randperm = fake
    try:
        parts = list(seq_index.minibatches(T, C, nmb, "chunk", L, torch.device("cpu")))
    finally:
        torch.randperm = orig
    want = ob.recurrent_chunk_indices(np.arange(chunks) if nmb == 1 else perm, T, C, nmb, L)
    assert len(parts) == nmb == len(want)
    code = (np.arange(T + 1)[:, None] * 1000 + np.arange(C)[None, :]).reshape(-1)  # buffer row -> t*1000 + n
    for (idx, rows, seq_len), (t0, n) in zip(parts, want):
        mb = len(t0)
        assert seq_len == L and rows == L * mb and idx.dtype == torch.int32
        got = code[idx.numpy()].reshape(L, mb)
        for l in range(L):
            np.testing.assert_array_equal(got[l], (t0 + l) * 1000 + n)      # batch row l*mb+j <-> (t0_j + l, n_j)
        np.testing.assert_array_equal(code[idx.numpy()[:mb]], t0 * 1000 + n)  # hidden state of sequence j: row index[j]
        assert idx.numpy().max() < T * C


@pytest.mark.parametrize("T,C,nmb", [(8, 6, 2), (5, 7, 1), (4, 9, 3)])
def test_naive_and_feed_forward_index_maps(T, C, nmb):
    rng = np.random.default_rng(C)
    perm = rng.permutation(C)
    orig, fake, _ = _replay(perm)
    torch.randperm = fake
    try:
        parts = list(seq_index.minibatches(T, C, nmb, "naive", 0, torch.device("cpu")))
    finally:
        torch.randperm = orig
    if nmb == 1:  # one minibatch = every trajectory in env order: the identity map (order of whole sequences only)
        assert parts == [(None, T * C, T)]
    else:
        want = ob.naive_recurrent_indices(perm, T, C, nmb)
        for (idx, rows, seq_len), ids in zip(parts, want):
            k = len(ids)
            assert seq_len == T and rows == T * k
            got = idx.numpy().reshape(T, k)
            for t in range(T):
                np.testing.assert_array_equal(got[t], t * C + ids)          # row t*k+j <-> (t, ids[j]); state: row ids[j]
    permf = rng.permutation(T * C)
    orig, fake, _ = _replay(permf)
    torch.randperm = fake
    try:
        parts = list(seq_index.minibatches(T, C, nmb, "ff", 0, torch.device("cpu")))
    finally:
        torch.randperm = orig
    if nmb == 1:
        assert parts == [(None, T * C, 0)]
    else:
        want = ob.feed_forward_indices(permf, T, C, nmb)
        for (idx, rows, seq_len), w in zip(parts, want):
            assert seq_len == 0 and rows == len(w)
            np.testing.assert_array_equal(idx.numpy(), w)


def test_reference_compatible_generators_agree_with_index_maps():
    """The buffers' reference-shaped generators (kept for API compatibility) and the index maps gather the same rows."""
    from harl_b200.common.buffers.on_policy_actor_buffer import OnPolicyActorBuffer
    from harl_b200.envs.spaces import Box, Discrete
    from tests import util as U

    T, N, L = 8, 6, 4
    cfg = U.base_args(episode_length=T, n_rollout_threads=N, use_recurrent_policy=True, data_chunk_length=L,
                      hidden_sizes=[8])
    b = OnPolicyActorBuffer(cfg, Box(shape=(3,)), Discrete(4), device=torch.device("cpu"))
    enc = (torch.arange(T + 1)[:, None] * 1000 + torch.arange(N)[None, :]).float()
    b.obs.copy_(enc[:, :, None].expand(T + 1, N, 3))
    b.rnn_states.copy_(enc[:, :, None, None].expand(T + 1, N, 1, 8))
    b.masks.copy_(enc[:, :, None])
    b.actions.copy_(enc[:T, :, None])
    perm = np.random.default_rng(3).permutation(T * N // L)
    for use in ("generator", "index"):
        orig, fake, _ = _replay(perm)
        torch.randperm = fake
        try:
            if use == "generator":
                batches = [(s[0][:, 0], s[1][:, 0, 0], s[3][:, 0]) for s in
                           b.recurrent_generator_actor(torch.zeros(T, N, 1), 2, L)]
            else:
                parts = list(seq_index.minibatches(T, N, 2, "chunk", L, torch.device("cpu")))
        finally:
            torch.randperm = orig
    flat = enc.reshape(-1).numpy()
    for (obs, h0, masks), (idx, rows, seq_len) in zip(batches, parts):
        np.testing.assert_array_equal(obs, flat[idx.numpy()])
        np.testing.assert_array_equal(masks, flat[idx.numpy()])
        np.testing.assert_array_equal(h0, flat[idx.numpy()[:rows // seq_len]])


@pytest.mark.parametrize("which", ["actor", "criticEP", "criticFP"])
@pytest.mark.parametrize("kind", ["ff", "naive", "chunk"])
def test_index_maps_reproduce_the_reference_generators(which, kind):
    """tests/golden/generators_index_maps.npz: batches the UNMODIFIED reference generators yielded from buffers whose
    elements encode their own (t, n[, a]) -- i.e. the reference's index maps themselves -- with the permutation it drew.
    Reading the encoded buffer in place through seq_index's index (and taking the hidden state of sequence j from buffer
    row index[j]) must give exactly those batches."""
    from tests import util as U

    g = U.load("generators_index_maps")
    T, N, A, L = (int(g[k]) for k in ("T", "N", "A", "L"))
    perm = g[f"{which}.{kind}.perm"]
    if which == "criticFP":
        C = N * A
        code = (np.arange(T + 1)[:, None, None] * 10000 + np.arange(N)[None, :, None] * 10 + np.arange(A)[None, None, :])
    else:
        C = N
        code = np.arange(T + 1)[:, None] * 1000 + np.arange(N)[None, :]
    flat = code.reshape(-1).astype(np.float32)   # time-major flatten, (n, a) -> n * A + a
    orig, fake, _ = _replay(perm)
    torch.randperm = fake
    try:
        parts = list(seq_index.minibatches(T, C, 2, kind, L, torch.device("cpu")))
    finally:
        torch.randperm = orig
    assert len(parts) == 2
    rows_key = "obs" if which == "actor" else "share_obs"
    for i, (idx, rows, seq_len) in enumerate(parts):
        want = g[f"{which}.{kind}.{i}.{rows_key}"]
        assert rows == len(want)
        np.testing.assert_array_equal(flat[idx.numpy()], want)
        np.testing.assert_array_equal(flat[idx.numpy()], g[f"{which}.{kind}.{i}.masks"])
        if which == "actor":
            np.testing.assert_array_equal(flat[idx.numpy()], g[f"{which}.{kind}.{i}.actions"])   # [T, N] arrays: same row index
            np.testing.assert_array_equal(flat[idx.numpy()], g[f"{which}.{kind}.{i}.adv"])
        else:
            np.testing.assert_array_equal(flat[idx.numpy()], g[f"{which}.{kind}.{i}.value_preds"])
        rnn = g[f"{which}.{kind}.{i}.rnn"]
        if kind == "ff":      # one state per row (unused by feed-forward nets)
            np.testing.assert_array_equal(flat[idx.numpy()], rnn)
        else:                 # one state per sequence: the row the first step of sequence j points at
            assert seq_len == (T if kind == "naive" else L) and len(rnn) == rows // seq_len
            np.testing.assert_array_equal(flat[idx.numpy()[:len(rnn)]], rnn)
