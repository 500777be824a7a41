"""Minibatch index maps of the three reference generators as device int32 tensors.

The kernels read the buffers in place: a minibatch is (index [rows] | None, rows, seq_len).  ``index`` maps batch
row -> buffer row of the time-major flatten (t * C + c, C = N envs, or N * A (env, agent) pairs for the FP critic);
recurrent batches are step-major (row = s * B + j) and their sequence j starts from the hidden state stored at
buffer row index[j] -- exactly what the reference generators gather:

* feed-forward  (on_policy_actor_buffer.py:114-178):  perm(T*C) split into num_mini_batch parts
* naive recurrent (:180-221):  perm(C) split into groups of C // num_mini_batch columns, whole trajectories
* recurrent chunks (:223-326):  chunks of ``data_chunk_length`` steps, perm(T*C/L), state at each chunk start

With a single minibatch the permutation only reorders sums (and the order of whole sequences), so it is skipped.
The permutations are drawn with torch.randperm on the CPU generator, like the reference.
"""
import torch


def minibatches(T, C, num_mini_batch, mode, chunk_len, device):
    """Yield (index int32 tensor or None, rows, seq_len) for one epoch.  mode: 'ff' | 'naive' | 'chunk'."""
    nmb = int(num_mini_batch)
    if mode == "ff":
        rows = T * C
        if nmb == 1:
            yield None, rows, 0
            return
        mb = rows // nmb
        perm = torch.randperm(rows).to(device=device, dtype=torch.int32)
        for i in range(nmb):
            yield perm[i * mb:(i + 1) * mb].contiguous(), mb, 0
        return
    if mode == "naive":
        assert C >= nmb
        if nmb == 1:
            yield None, T * C, T
            return
        k = C // nmb
        perm = torch.randperm(C).to(device)
        steps = torch.arange(T, device=device)
        for i in range(nmb):
            ids = perm[i * k:(i + 1) * k]
            idx = (steps[:, None] * C + ids[None, :]).reshape(-1).to(torch.int32).contiguous()
            yield idx, T * k, T
        return
    L = int(chunk_len)
    assert T % L == 0, "episode_length must be a multiple of data_chunk_length"  # reference :235-238
    per_col = T // L
    chunks = C * per_col
    assert chunks >= 2
    mb = chunks // nmb
    order = torch.arange(chunks, device=device) if nmb == 1 else torch.randperm(chunks).to(device)
    steps = torch.arange(L, device=device)
    for i in range(nmb):
        c = order[i * mb:(i + 1) * mb]
        col, t0 = c // per_col, (c % per_col) * L
        idx = ((t0[None, :] + steps[:, None]) * C + col[None, :]).reshape(-1).to(torch.int32).contiguous()
        yield idx, L * mb, L


def mode_of(use_recurrent_policy, use_naive_recurrent_policy):
    return "chunk" if use_recurrent_policy else ("naive" if use_naive_recurrent_policy else "ff")
