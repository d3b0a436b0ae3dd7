"""CPU: pin the oracle (oracle/dct_oracle.c) against outputs of the reference itself.

The fixtures in tests/golden/reference_golden.npz were produced by
oracle/make_golden.py from the unmodified reference (torch CPU).  Integer
results must be bit-exact; floating point within 1e-5 scaled (tests/util.py).
"""
import numpy as np
import pytest

from util import assert_close, cases, lnK


def _names(prefix):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
    return cases(g, prefix)


@pytest.mark.parametrize("case", _names("jsd_"))
def test_jsd_oracle(case, golden, oracle):
    z = golden[case + "/z"]
    K, B = z.shape[0], z.shape[1]
    npx = z.shape[1] * z.shape[3] * z.shape[4]
    w = float(golden[case + "/w"])
    mean, mp, gz = oracle.jsd_logits_fwdbwd(list(z), w)
    assert_close(mp, golden[case + "/ref_map32"], floor=lnK(K), what="map")
    assert_close(mean, golden[case + "/ref_mean32"], floor=lnK(K), what="mean")
    assert_close(np.stack(gz), golden[case + "/ref_gz32"], floor=w / npx, what="grad logits")
    # fp64 twin against fp64 reference: restatement is exact to rounding
    mean64, mp64, gz64 = oracle.jsd_logits_fwdbwd(list(z.astype(np.float64)), w)
    assert_close(mp64, golden[case + "/ref_map64"], rtol=1e-11, floor=lnK(K), what="map64")
    assert_close(np.stack(gz64), golden[case + "/ref_gz64"], rtol=1e-11, floor=w / npx, what="grad64")
    # probs boundary (parity mode): forward from the reference's probs, d/dprobs with upstream gout
    probs = list(golden[case + "/ref_probs32"])
    assert oracle.simplex_violations(probs[0]) == 0
    assert_close(oracle.jsd_fwd(probs), golden[case + "/ref_map32"], floor=lnK(K), what="map from probs")
    gout = golden[case + "/gout"]
    gp = oracle.jsd_bwd(probs, gout)
    assert_close(np.stack(gp), golden[case + "/ref_gp32"], floor=float(np.abs(gout).max()), what="grad probs")
    if case.endswith("identical") and K in (2, 4):
        # identical views -> exactly 0 when (p+..+p)/K is exact; the reference agrees
        assert float(np.abs(mp).max()) == 0.0 and float(np.abs(golden[case + "/ref_map32"]).max()) == 0.0


@pytest.mark.parametrize("case", [c for c in _names("jsd_") if c.endswith("spread")])
def test_jsd_nd_and_entropy_oracle(case, golden, oracle):
    probs = list(golden[case + "/ref_probs32"])
    K = len(probs)
    assert_close(oracle.jsd_fwd(probs), golden[case + "/ref_JSD_map"], floor=lnK(K))
    assert_close(oracle.jsd_fwd(probs).mean(dtype=np.float64), golden[case + "/ref_JSD_reduce"], floor=lnK(K))
    assert_close(oracle.entropy(probs[0]), golden[case + "/ref_entropy0"], floor=1.0)


@pytest.mark.parametrize("case", _names("kl_"))
def test_kl_oracle(case, golden, oracle):
    p, y, gout = golden[case + "/p32"], golden[case + "/y32"], golden[case + "/gout"]
    gs = float(np.abs(gout).max())
    assert_close(oracle.kl_fwd(p, y), golden[case + "/ref_map32"], floor=1.0)
    assert_close(oracle.kl_fwd(p, y).mean(dtype=np.float64), golden[case + "/ref_mean32"], floor=1.0)
    gp, gy = oracle.kl_bwd(p, y, gout)
    assert_close(gp, golden[case + "/ref_gp32"], floor=gs)
    assert_close(gy, golden[case + "/ref_gy32"], floor=gs)
    # trainer composite: softmax -> KL mean -> backward to logits
    n = gout.size
    g1 = oracle.kl_bwd(p, y, np.full(gout.shape, 1.0 / n, np.float32))[0]
    assert_close(oracle.softmax_bwd(p, g1), golden[case + "/ref_gzp_mean32"], floor=1.0 / n)
    m, gpl, gql = oracle.kl_logit(golden[case + "/zy"], golden[case + "/zp"], gout)
    assert_close(m, golden[case + "/ref_logit_map32"], floor=1.0)
    assert_close(gpl, golden[case + "/ref_logit_gpl32"], floor=gs)
    assert_close(gql, golden[case + "/ref_logit_gql32"], floor=gs)
    # KL_Divergence_2D_Logit(p_logit, y_logit) == kl_div_with_logit(q=y_logit, p=p_logit)
    assert_close(m, golden[case + "/ref_logit2d_map32"], floor=1.0)
    assert_close(oracle.kl_div_fwd(p, y), golden[case + "/ref_kldiv_map32"], floor=1.0)
    # fp64 twins
    p64, y64 = golden[case + "/p64"], golden[case + "/y64"]
    assert_close(oracle.kl_fwd(p64, y64), golden[case + "/ref_map64"], rtol=1e-11, floor=1.0)
    assert_close(oracle.softmax(golden[case + "/zp"]), p, floor=1.0)


@pytest.mark.parametrize("case", _names("vat_"))
def test_vat_oracle(case, golden, oracle):
    assert_close(oracle.l2_normalize(golden[case + "/d"]), golden[case + "/ref_l2"])
    adv, noise = oracle.fgsm(golden[case + "/img"], golden[case + "/grad"], 0.05)
    assert np.array_equal(adv, golden[case + "/ref_fgsm_adv"])
    assert np.array_equal(noise, golden[case + "/ref_fgsm_noise"])


@pytest.mark.parametrize("case", [c for c in _names("dice_") if not c.endswith("meter")])
def test_dice_oracle_bit_exact(case, golden, oracle):
    x, gt = golden[case + "/x"], golden[case + "/gt"]
    assert np.array_equal(oracle.dice(x, gt, "2d"), golden[case + "/ref_2d"])
    assert np.array_equal(oracle.dice(x, gt, "3d"), golden[case + "/ref_3d"])
    if case.endswith("_logits"):
        assert bool(golden[case + "/ref_bad_label_raises"])
        bad = gt.copy(); bad[0, 0, 0, 0] = x.shape[1]
        with pytest.raises(AssertionError):
            oracle.dice(x, bad, "2d")


@pytest.mark.parametrize("case", _names("iou_"))
def test_confusion_oracle_bit_exact(case, golden, oracle):
    C = golden[case + "/x0"].shape[1]
    conf = np.zeros((C, C), np.int64)
    for j in range(2):
        conf += oracle.confusion(golden[f"{case}/x{j}"], golden[f"{case}/gt{j}"])
        assert np.array_equal(conf, golden[f"{case}/ref_conf_after{j}"])
    v = oracle.iou_value(conf)
    for k in ("Overall_Acc", "Mean_Acc", "FreqW_Acc", "Validated_Mean_IoU", "Mean_IoU"):
        assert v[k] == golden[f"{case}/ref_{k}"]
    assert np.array_equal(v["Class_IoU"], golden[case + "/ref_Class_IoU"])
    got = oracle.confusion(golden[case + "/pred_map"], golden[case + "/gt1"].squeeze(1), C)
    assert np.array_equal(got, golden[case + "/ref_conf_from_map"])


def test_spec_expf_properties(oracle):
    """The pinned softmax arithmetic: exp(0)==1 exactly, faithful (<=2 ulp) on [-87,0], 0 below."""
    assert oracle.spec_expf(0.0) == 1.0
    assert oracle.spec_expf(-100.0) == 0.0
    rng = np.random.default_rng(0)
    d = -np.abs(rng.standard_normal(20000).astype(np.float32)) * 20
    d = d[d >= -87]
    got = np.array([oracle.spec_expf(float(v)) for v in d], dtype=np.float64)
    ref = np.exp(d.astype(np.float64))
    ulp = np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ulp) <= 2.0
    # margin property used by the CUDA fast path: d <= -2^-17  =>  e <= 1 - 2^-18 (needed: <= 1 - 2^-22)
    for v in (-2.0 ** -17, -2.0 ** -16, -1e-4, -1e-3):
        assert oracle.spec_expf(v) <= 1.0 - 2.0 ** -18
