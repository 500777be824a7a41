"""CPU: the oracle half of tests/test_gpu_zz_wide_heads.py::test_fvp_wide_head_vs_oracle (keeps that test's setup exercised
where no GPU is available)."""
import torch

from tests.test_gpu_zz_wide_heads import _wide_case


def test_wide_case_oracle_side_runs_on_cpu():
    from oracle import trpo as ot

    for recurrent in (False, True):
        cfg, p, _, _, _, batch, vec, _ = _wide_case(recurrent)
        out = ot.fisher_vector_product(p, cfg, "Discrete", batch, vec)
        assert all(torch.isfinite(v).all() for v in out.values())
