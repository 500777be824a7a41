"""CPU: index-for-index NumPy transcription of the (not yet GPU-run) persistent GRU recurrence kernel
``gru_seq_fwd_kernel`` (harl_b200/csrc/rnn.cu) -- its shared-memory weight regrouping, the lane <-> hidden-unit mapping,
the shuffle broadcast order and the gate arithmetic -- checked against the plain GRU recurrence the per-step kernels
implement.  This pins the kernel's INDEX MATH; CUDA
semantics (shuffles, alignment, occupancy) still need the GPU run (tests/test_gpu_rnn.py, HB_RUN_EXPERIMENTAL=1)."""
import numpy as np
import torch

H = 64


def _fill_smem(whh_t):
    """gp_w[f], f in [0, 64*2*32*4): i = f >> 8, part = (f >> 7) & 1, l = (f >> 2) & 31, c = part*4 + (f & 3)."""
    gp = np.zeros(H * 2 * 32 * 4, np.float32)
    flat = whh_t.reshape(-1)
    for f in range(gp.size):
        i, part, l, c = f >> 8, (f >> 7) & 1, (f >> 2) & 31, ((f >> 7) & 1) * 4 + (f & 3)
        gp[f] = flat[i * 3 * H + (c >> 1) * H + 2 * l + (c & 1)] if c < 6 else 0.0
    return gp


def _emulate(whh_t, bhh, gi, h0, mrow, S, B):
    gp = _fill_smem(whh_t)
    f32 = np.float32
    hs = np.zeros((S * B, H), f32)
    hm_out = np.zeros((S * B, H), f32)
    sig = lambda x: f32(1.0) / (f32(1.0) + np.exp(-x, dtype=f32))
    for j in range(B):
        hx = h0[j, 0::2].copy()   # lane l holds units 2l (x) and 2l + 1 (y)
        hy = h0[j, 1::2].copy()
        for t in range(S):
            row = t * B + j
            m = mrow[row]
            hm0, hm1 = hx * m, hy * m
            hm_out[row, 0::2], hm_out[row, 1::2] = hm0, hm1
            a = np.zeros((32, 6), f32)
            for ii in range(32):
                v0, v1 = hm0[ii], hm1[ii]                      # __shfl_sync(hm0 / hm1, ii)
                for lane in range(32):
                    w0 = ((2 * ii) * 64 + lane) * 4            # float index of w0[0]; w0[32] is +128 floats
                    w1 = ((2 * ii + 1) * 64 + lane) * 4
                    p, q = gp[w0:w0 + 4], gp[w0 + 128:w0 + 132]
                    a[lane] += v0 * np.array([p[0], p[1], p[2], p[3], q[0], q[1]], f32)
                    p1, q1 = gp[w1:w1 + 4], gp[w1 + 128:w1 + 132]
                    a[lane] += v1 * np.array([p1[0], p1[1], p1[2], p1[3], q1[0], q1[1]], f32)
            for lane in range(32):
                e = 2 * lane
                g = gi[row]
                ghr = a[lane, 0:2] + bhh[e:e + 2]
                ghz = a[lane, 2:4] + bhh[H + e:H + e + 2]
                ghn = a[lane, 4:6] + bhh[2 * H + e:2 * H + e + 2]
                r = sig(g[e:e + 2] + ghr)
                z = sig(g[H + e:H + e + 2] + ghz)
                n = np.tanh(g[2 * H + e:2 * H + e + 2] + r * ghn)
                hmv = np.array([hm0[lane], hm1[lane]], f32)
                hn = (f32(1.0) - z) * n + z * hmv
                hx[lane], hy[lane] = hn
                hs[row, e:e + 2] = hn
    return hs, hm_out


def test_persistent_gru_kernel_index_math():
    rng = np.random.default_rng(0)
    S, B = 3, 2
    w_hh = (0.3 * rng.standard_normal((3 * H, H))).astype(np.float32)     # PyTorch layout [3h][h], gates r, z, n
    b_hh = (0.1 * rng.standard_normal(3 * H)).astype(np.float32)
    whh_t = np.ascontiguousarray(w_hh.T)                                  # prepared layout [h][3h]
    gi = rng.standard_normal((S * B, 3 * H)).astype(np.float32)           # = x W_ih^T + b_ih of every step
    h0 = rng.standard_normal((B, H)).astype(np.float32)
    mrow = np.array([1, 1, 0, 1, 1, 1], np.float32)                       # a reset of sequence 0 at step 1
    hs, hm = _emulate(whh_t, b_hh, gi, h0, mrow, S, B)
    # the plain recurrence (what rnn_linear + gru_gate_fwd_kernel compute; PyTorch GRU gate order r, z, n)
    h = torch.from_numpy(h0)
    for t in range(S):
        m = torch.from_numpy(mrow[t * B:(t + 1) * B])[:, None]
        hmask = h * m
        np.testing.assert_allclose(hm[t * B:(t + 1) * B], hmask.numpy(), rtol=0, atol=1e-6)
        gi_t = torch.from_numpy(gi[t * B:(t + 1) * B])
        gh = hmask @ torch.from_numpy(w_hh).T + torch.from_numpy(b_hh)
        r = torch.sigmoid(gi_t[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi_t[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi_t[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * hmask
        np.testing.assert_allclose(hs[t * B:(t + 1) * B], h.numpy(), rtol=0, atol=2e-5)
