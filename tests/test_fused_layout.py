"""CPU transcription of the operand-image addressing of the fused tcgen05 kernels (harl_b200/csrc/fused_update.cu:
img_off, op_kmajor, op_mnmajor, gemm3's k-step advance; umma.cuh: desc).

The kernels rest on one claim: a tile image written ROW by row in the canonical no-swizzle core-matrix layout
``IMG[row/8][feature/8][row%8][8 x fp16]`` can be handed to the tensor core BOTH as a K-major operand (rows = M/N,
features = K: forward and dX GEMMs) and as an MN-major operand (features = M/N, rows = K: weight-gradient GEMMs) by
swapping the leading / stride byte offsets of the shared-memory descriptor -- no transpose, no second copy.  The B200
probe (profiles/probe_umma_layouts.cu) established the hardware's reading of the descriptor fields, restated in
``element_address`` below; this test checks that the offsets the kernel computes address, for every logical element
of every GEMM it issues, exactly the bytes the epilogues wrote.  The constants are parsed out of the CUDA source so
the test fails if the two drift apart.
"""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "harl_b200", "csrc", "fused_update.cu")


def img_off(r, ch, wch):
    """fused_update.cu img_off: byte offset of (row r, 8-feature chunk ch) in an image with wch chunks per row."""
    return ((r >> 3) * wch + ch) * 128 + (r & 7) * 16


def element_address(start, lbo, sbo, major, mn, k):
    """Byte address of logical element (mn, k) of ONE MMA k-step (K = 16 fp16) as the tensor core reads a no-swizzle
    descriptor (start address, LBO, SBO), per the B200 probe:
      K-major  : 8 x 8 core matrices, 16 bytes per mn-row inside; next core matrix along K at +LBO, along MN at +SBO
      MN-major : core matrix = 8 k-rows of 8 mn-elements (16 bytes per k-row); next core matrix along MN at +SBO, along K at
                 +LBO -- the same roles, which is why swapping the two byte offsets transposes the view."""
    if major == "K":
        return start + (mn >> 3) * sbo + (k >> 3) * lbo + (mn & 7) * 16 + (k & 7) * 2
    return start + (mn >> 3) * sbo + (k >> 3) * lbo + (k & 7) * 16 + (mn & 7) * 2


def op_kmajor(base, wch, ch0):
    return dict(start=base + ch0 * 128, lbo=128, sbo=wch * 128, adv=256, major="K")


def op_mnmajor(base, wch, ch0):
    return dict(start=base + ch0 * 128, lbo=wch * 128, sbo=128, adv=2 * wch * 128, major="MN")


def test_source_constants_match_this_transcription():
    s = open(SRC).read()
    assert re.search(r"img_off\(int r, int ch, int wch\) \{ return \(uint32_t\)\(\(\(r >> 3\) \* wch \+ ch\) \* 128 \+ \(r & 7\) \* 16\); \}", s)
    km = re.search(r"Op op_kmajor\(.*?return Op\{(.*?)\};", s, re.S).group(1)
    mn = re.search(r"Op op_mnmajor\(.*?return Op\{(.*?)\};", s, re.S).group(1)
    assert km.replace(" ", "").endswith("128u,(uint32_t)wch*128u,256u")           # lbo, sbo, k-step advance
    assert mn.replace(" ", "").endswith("(uint32_t)wch*128u,128u,2u*(uint32_t)wch*128u")


def _written_image(rows, feats):
    """Address of every (row, feature) as the epilogues write it: 16-byte stores of 8 features at img_off."""
    wch = feats // 8
    addr = np.zeros((rows, feats), np.int64)
    for r in range(rows):
        for f in range(feats):
            addr[r, f] = img_off(r, f >> 3, wch) + (f & 7) * 2
    assert len(np.unique(addr)) == rows * feats and addr.max() == rows * feats * 2 - 2     # a dense, hole-free image
    return addr, wch


def test_kmajor_view_reads_what_the_epilogue_wrote():
    """Forward / dX GEMMs: A = image as [128 rows (M)] x [features (K)], consumed 32-feature chunk by chunk, two k-steps of
    16 per chunk (gemm3: descriptor start + ks * adv)."""
    for feats in (32, 64, 128):
        addr, wch = _written_image(128, feats)
        for c in range(feats // 32):                       # op_kmajor(base, img_bytes, wch, 4 * c), ksteps = 2
            op = op_kmajor(0, wch, 4 * c)
            for ks in range(2):
                for m in range(128):
                    for k in range(16):
                        got = element_address(op["start"] + ks * op["adv"], op["lbo"], op["sbo"], "K", m, k)
                        assert got == addr[m, 32 * c + 16 * ks + k]


def test_mnmajor_view_is_the_transpose_of_the_same_bytes():
    """Weight-gradient GEMMs: the SAME image as [features (M or N)] x [128 rows (K)], eight k-steps of 16 rows
    (gemm3 with ksteps = TILE >> 4)."""
    for feats in (16, 32, 64, 128):
        addr, wch = _written_image(128, feats)
        op = op_mnmajor(0, wch, 0)
        for ks in range(8):
            for f in range(feats):
                for k in range(16):
                    got = element_address(op["start"] + ks * op["adv"], op["lbo"], op["sbo"], "MN", f, k)
                    assert got == addr[16 * ks + k, f]


def test_column_quarters_of_the_pipelined_forward_match_k_chunks():
    """A forward epilogue signals e2m_h when columns [0, 32) (column-half-0 thread) and [64, 96) (half-1 thread) are final;
    the MMA issuer then runs k-chunks 0 and 2 (hd(0), hd(2) / l1(0), l1(2)) and the ring delivers weights in the order
    0, 2, 1, 3 (producer: c = (ci & 1) * 2 + (ci >> 1))."""
    assert [(ci & 1) * 2 + (ci >> 1) for ci in range(4)] == [0, 2, 1, 3]
    nc = 64                                               # columns per thread at H = 128
    first_half = {0: range(0, nc // 2), 1: range(64, 64 + nc // 2)}
    early = sorted(set(c // 32 for h in (0, 1) for c in first_half[h]))
    assert early == [0, 2]
    s = open(SRC).read()
    assert "l1(0, false); l1(2, true);" in s and "hd(0, false); hd(2, true);" in s
    assert "if (c8 == (nc >> 4) - 1) mid();" in s


def test_tmem_and_shared_memory_budgets():
    """The column map of the 512 TMEM columns and the shared-memory carve-up quoted in DESIGN.md section 4a."""
    s = open(SRC).read()
    cols = {k: int(v) for k, v in re.findall(r"constexpr uint32_t (C_\w+) = (\d+);", s)}
    width = dict(C_F=128, C_W1=128, C_W0=64, C_H=16, C_WH=16, C_B1=16, C_B0=16, C_BH=16)
    spans = sorted((cols[k], cols[k] + w) for k, w in width.items())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= 512      # disjoint, inside the allocation
    assert spans[-1][1] == 400                                                               # "400 of 512 TMEM columns"
    TILE, NH, STAGE = 128, 16, 16384
    for k0p, stages in ((32, 3), (64, 2)):
        images = 2 * TILE * k0p * 2 + 2 * 2 * TILE * 128 * 2 + 3 * TILE * NH * 2 + 2 * NH * 128 * 2
        total = images + stages * STAGE + 8 * 1024            # + biases, exchange slots, barriers, alignment (< 8 KB)
        assert total <= 227 * 1024, (k0p, total)
        assert images + (stages + 1) * STAGE + 4 * 1024 > 227 * 1024   # one more ring stage does not fit
