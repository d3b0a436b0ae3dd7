"""CPU: the oracle's backward restatements (entropy, softmax, KL_div) against tests/golden/reference_golden_r2.npz
(oracle/make_golden_r2.py: torch autograd through the unmodified reference)."""
import os

import numpy as np
import pytest

from util import assert_close, cases

G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_r2.npz"))


@pytest.mark.parametrize("case", cases(G2, "ent_"))
def test_entropy_bwd_oracle(case, oracle):
    for tag, rtol in (("32", 1e-5), ("64", 1e-11)):
        p, gout = G2[case + "/p" + tag], G2[case + "/gout"].astype(G2[case + "/p" + tag].dtype)
        assert_close(oracle.entropy(p), G2[case + "/ref_map" + tag], rtol=rtol, floor=1.0, what="map" + tag)
        assert_close(oracle.entropy_bwd(p, gout), G2[case + "/ref_gp" + tag], rtol=rtol, what="gp" + tag)
    n = G2[case + "/gout"].size
    gmean = np.full(G2[case + "/gout"].shape, 1.0 / n, dtype=np.float32)
    assert_close(oracle.entropy_bwd(G2[case + "/p32"], gmean), G2[case + "/ref_gp_mean32"], what="gp under mean")


@pytest.mark.parametrize("case", cases(G2, "sm_"))
def test_softmax_bwd_oracle(case, oracle):
    for tag, rtol in (("32", 1e-5), ("64", 1e-11)):
        dt = np.float32 if tag == "32" else np.float64
        p = oracle.softmax(G2[case + "/z"].astype(dt))
        assert_close(p, G2[case + "/ref_p" + tag], rtol=min(rtol, 2e-7), floor=1.0, what="p" + tag)
        assert_close(oracle.softmax_bwd(p, G2[case + "/gp"].astype(dt)), G2[case + "/ref_gz" + tag], rtol=rtol, what="gz" + tag)


@pytest.mark.parametrize("case", cases(G2, "kldiv_"))
def test_kl_div_bwd_oracle(case, oracle):
    for tag, rtol in (("32", 2e-5), ("64", 1e-11)):
        p, q = G2[case + "/p" + tag], G2[case + "/q" + tag]
        gout = G2[case + "/gout"].astype(p.dtype)
        assert_close(oracle.kl_div_fwd(p, q), G2[case + "/ref_map" + tag], rtol=rtol, what="map" + tag)
        gp, gq = oracle.kl_div_bwd(p, q, gout)
        assert_close(gp, G2[case + "/ref_gp" + tag], rtol=rtol, what="gp" + tag)
        assert_close(gq, G2[case + "/ref_gq" + tag], rtol=rtol, what="gq" + tag)
    # element-wise on the fp64 twin: q/p spans many decades, a max-scaled bound alone says nothing about the small ones
    gp, gq = oracle.kl_div_bwd(G2[case + "/p64"], G2[case + "/q64"], G2[case + "/gout"].astype(np.float64))
    np.testing.assert_allclose(gq, G2[case + "/ref_gq64"], rtol=1e-10)
    np.testing.assert_allclose(gp, G2[case + "/ref_gp64"], rtol=1e-9, atol=1e-12)
    n = G2[case + "/gout"].size
    gmean = np.full(G2[case + "/gout"].shape, 0.37 / n, dtype=np.float32)
    gp, gq = oracle.kl_div_bwd(G2[case + "/p32"], G2[case + "/q32"], gmean)
    assert_close(gp, G2[case + "/ref_gp_mean32"], rtol=2e-5, what="gp under 0.37*mean")
    assert_close(gq, G2[case + "/ref_gq_mean32"], rtol=2e-5, what="gq under 0.37*mean")
