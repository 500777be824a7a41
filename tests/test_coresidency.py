"""Static check of the co-residency rule of DESIGN.md section 4a.

The fused update kernel is persistent (148 CTAs x 320 threads x 168 registers, 222 KB of shared memory, static tile
assignment): anything that can run at the same time on the other stream -- the per-step chain kernels of the actor /
critic update and the exchange kernel that spins for its peers -- must fit on an SM NEXT TO one of its CTAs, or it waits
for the whole persistent launch (and a persistent launch that starts some CTAs late runs up to twice as long).  Register
and shared-memory figures come from the ptxas log of the build (harl_b200/_C/ptxas.log), thread counts from the launch
sites in the sources."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "harl_b200", "_C", "ptxas.log")

REGS_PER_SM, SMEM_PER_SM = 65536, 228 * 1024
FUSED_THREADS, FUSED_SMEM = 320, 222 * 1024 + 1024          # + the 1 KB the driver reserves per CTA

# kernel name fragment -> threads per CTA at its launch site
NEIGHBOURS = {
    "allreduce_oneshot_kernel": 128,      # p2p_comm.cu COMM_THREADS
    "clip_adam_kernel": 256,              # optim.cu
    "fused_slot_reduce_kernel": 256,      # fused_update.cu 32 * SR_GROUPS
    "fused_unfold_kernel": 128,
    "fused_pack_kernel": 256,
    "prepare_kernel": 256,                # layout.cu
    "pack_umma_jobs_kernel": 256,         # tc_gemm.cu
    "masked_moments_kernel": 256,         # gae.cu
    "normalize_kernel": 256,
    "valuenorm_update_kernel": 32,
}


def _entries():
    txt = open(LOG).read()
    out = []
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?Used (\d+) registers.*?(?:, (\d+) bytes smem)?\n", txt, re.S):
        out.append((m.group(1), int(m.group(2)), int(m.group(3) or 0)))
    return out


@pytest.mark.skipif(not os.path.exists(LOG), reason="needs the ptxas log of a build (python -c 'import __graft_entry__ as g; g.build()')")
def test_chain_and_exchange_kernels_fit_beside_a_persistent_update_cta():
    ents = _entries()
    fused = [e for e in ents if "fused_update_kernel" in e[0]]
    assert fused, "fused_update_kernel not in the ptxas log"
    fused_regs = max(e[1] for e in fused)
    assert fused_regs <= 168
    alloc = lambda regs, threads: ((regs + 7) // 8 * 8) * ((threads + 31) // 32 * 32)
    free_regs = REGS_PER_SM - alloc(fused_regs, FUSED_THREADS)
    free_smem = SMEM_PER_SM - FUSED_SMEM
    assert free_regs >= 10000 and free_smem >= 3 * 1024
    seen = set()
    for frag, threads in NEIGHBOURS.items():
        hits = [e for e in ents if frag in e[0]]
        assert hits, frag
        for name, regs, smem in hits:
            assert alloc(regs, threads) <= free_regs, (name, regs, threads, free_regs)
            assert smem + 1024 <= free_smem, (name, smem)
            assert threads + FUSED_THREADS <= 2048
        seen.add(frag)
    assert seen == set(NEIGHBOURS)


def test_launch_sites_use_the_thread_counts_assumed_above():
    src = lambda f: open(os.path.join(ROOT, "harl_b200", "csrc", f)).read()
    assert "constexpr int COMM_THREADS = 128;" in src("p2p_comm.cu")
    assert "clip_adam_kernel<<<adam_ctas, 256, 0, st>>>" in src("optim.cu")
    fu = src("fused_update.cu")
    assert "constexpr int SR_GROUPS = 8" in fu and "32 * fz::SR_GROUPS, 0, st>>>" in fu
    assert "fused_unfold_kernel<<<(cols * 32 + 127) / 128, 128, 0, st>>>" in fu
    assert "fz::fused_pack_kernel<<<cta, 256, 0, st>>>" in fu
    assert "constexpr int THREADS = 320;" in fu
