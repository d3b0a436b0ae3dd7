"""ctypes/numpy front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Loads oracle/libdct_oracle.so (built by oracle/Makefile from dct_oracle.c) and
exposes one numpy-in / numpy-out function per reference function on the hot
path.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline``
/ ``--impl reference`` legs import this module; the product package never does.

Shapes follow the reference: logits/probs ``[B,C,H,W]`` (any trailing spatial
shape is flattened to HW), maps ``[B,H,W]``, labels ``[B,1,H,W]`` / ``[B,H,W]`` int64.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdct_oracle.so")

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("dct_oracle.c", "dct_oracle_body.inc")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libdct_oracle.so"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dcto_jsd_logits_fwdbwd_f32.restype = C.c_double
        _lib.dcto_jsd_logits_fwdbwd_f64.restype = C.c_double
        _lib.dcto_simplex_violations_f32.restype = _i64
        _lib.dcto_simplex_violations_f64.restype = _i64
        _lib.dcto_confusion_from_labels.restype = _i64
        _lib.dcto_ce_f32.restype = _i64
        _lib.dcto_ce_f64.restype = _i64
        _lib.dcto_spec_expf.restype = C.c_float
        _lib.dcto_spec_expf.argtypes = [C.c_float]
        _lib.dcto_num_threads.restype = _int
    return _lib


def num_threads() -> int:
    return int(lib().dcto_num_threads())


def set_num_threads(n: int) -> None:
    lib().dcto_set_num_threads(_int(int(n)))


def _dt(a):
    return np.float64 if np.asarray(a).dtype == np.float64 else np.float32


def _suf(dt):
    return "_f64" if dt == np.float64 else "_f32"


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _bchw(a):
    b, c = a.shape[0], a.shape[1]
    hw = int(np.prod(a.shape[2:], dtype=np.int64)) if a.ndim > 2 else 1
    return b, c, hw


def _ptr(a):
    return _p(a.ctypes.data)


def _ptr_array(arrs):
    return (_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _real(dt, v):
    return C.c_double(v) if dt == np.float64 else C.c_float(v)


def spec_expf(d: float) -> float:
    return float(lib().dcto_spec_expf(C.c_float(d)))


def softmax(z):
    dt = _dt(z); z = _c(z, dt); b, c, hw = _bchw(z)
    out = np.empty_like(z)
    getattr(lib(), "dcto_softmax" + _suf(dt))(_ptr(z), _ptr(out), _i64(b), _int(c), _i64(hw))
    return out


def softmax_bwd(p, gp):
    dt = _dt(p); p = _c(p, dt); gp = _c(gp, dt); b, c, hw = _bchw(p)
    out = np.empty_like(p)
    getattr(lib(), "dcto_softmax_bwd" + _suf(dt))(_ptr(p), _ptr(gp), _ptr(out), _i64(b), _int(c), _i64(hw))
    return out


def simplex_violations(p) -> int:
    dt = _dt(p); p = _c(p, dt); b, c, hw = _bchw(p)
    return int(getattr(lib(), "dcto_simplex_violations" + _suf(dt))(_ptr(p), _i64(b), _int(c), _i64(hw)))


def entropy(p):
    dt = _dt(p); p = _c(p, dt); b, c, hw = _bchw(p)
    out = np.empty((b,) + p.shape[2:], dtype=dt)
    getattr(lib(), "dcto_entropy" + _suf(dt))(_ptr(p), _ptr(out), _i64(b), _int(c), _i64(hw))
    return out


def entropy_bwd(p, gout):
    dt = _dt(p); p = _c(p, dt); gout = _c(gout, dt); b, c, hw = _bchw(p)
    gp = np.empty_like(p)
    getattr(lib(), "dcto_entropy_bwd" + _suf(dt))(_ptr(p), _ptr(gout), _ptr(gp), _i64(b), _int(c), _i64(hw))
    return gp


def jsd_fwd(probs):
    dt = _dt(probs[0]); ps = [_c(p, dt) for p in probs]; b, c, hw = _bchw(ps[0])
    out = np.empty((b,) + ps[0].shape[2:], dtype=dt)
    getattr(lib(), "dcto_jsd_fwd" + _suf(dt))(_ptr_array(ps), _int(len(ps)), _i64(b), _int(c), _i64(hw), _ptr(out))
    return out


def jsd_bwd(probs, gout):
    dt = _dt(probs[0]); ps = [_c(p, dt) for p in probs]; b, c, hw = _bchw(ps[0])
    gout = _c(gout, dt)
    gps = [np.empty_like(p) for p in ps]
    getattr(lib(), "dcto_jsd_bwd" + _suf(dt))(_ptr_array(ps), _int(len(ps)), _i64(b), _int(c), _i64(hw),
                                             _ptr(gout), _ptr_array(gps))
    return gps


def jsd_logits_fwdbwd(logits, weight=1.0, want_map=True, want_grad=True):
    """softmax -> JSD_2D -> .mean() -> *weight -> backward.  Returns (mean, map|None, [grad_k]|None)."""
    dt = _dt(logits[0]); zs = [_c(z, dt) for z in logits]; b, c, hw = _bchw(zs[0])
    mp = np.empty((b,) + zs[0].shape[2:], dtype=dt) if want_map else None
    gz = [np.empty_like(z) for z in zs] if want_grad else None
    mean = getattr(lib(), "dcto_jsd_logits_fwdbwd" + _suf(dt))(
        _ptr_array(zs), _int(len(zs)), _i64(b), _int(c), _i64(hw), _real(dt, weight),
        _ptr(mp) if want_map else _p(0), _ptr_array(gz) if want_grad else _p(0))
    if mean < 0 and (c > 64 or len(zs) > 8):
        raise ValueError("oracle limits: C<=64, K<=8")
    return float(mean), mp, gz


def kl_fwd(p, y, eps=1e-10):
    dt = _dt(p); p = _c(p, dt); y = _c(y, dt); b, c, hw = _bchw(p)
    out = np.empty((b,) + p.shape[2:], dtype=dt)
    getattr(lib(), "dcto_kl_fwd" + _suf(dt))(_ptr(p), _ptr(y), _i64(b), _int(c), _i64(hw), _real(dt, eps), _ptr(out))
    return out


def kl_bwd(p, y, gout, eps=1e-10):
    dt = _dt(p); p = _c(p, dt); y = _c(y, dt); gout = _c(gout, dt); b, c, hw = _bchw(p)
    gp = np.empty_like(p); gy = np.empty_like(p)
    getattr(lib(), "dcto_kl_bwd" + _suf(dt))(_ptr(p), _ptr(y), _i64(b), _int(c), _i64(hw), _real(dt, eps),
                                            _ptr(gout), _ptr(gp), _ptr(gy))
    return gp, gy


def kl_logit(q_logit, p_logit, gout=None):
    """kl_div_with_logit / KL_Divergence_2D_Logit.  Returns (map, grad_p_logit|None, grad_q_logit|None)."""
    dt = _dt(q_logit); ql = _c(q_logit, dt); pl = _c(p_logit, dt); b, c, hw = _bchw(ql)
    if c > 64:
        raise ValueError("oracle limit: C<=64")
    out = np.empty((b,) + ql.shape[2:], dtype=dt)
    if gout is None:
        getattr(lib(), "dcto_kl_logit" + _suf(dt))(_ptr(ql), _ptr(pl), _i64(b), _int(c), _i64(hw), _ptr(out),
                                                  _p(0), _p(0), _p(0))
        return out, None, None
    gout = _c(gout, dt); gpl = np.empty_like(pl); gql = np.empty_like(ql)
    getattr(lib(), "dcto_kl_logit" + _suf(dt))(_ptr(ql), _ptr(pl), _i64(b), _int(c), _i64(hw), _ptr(out),
                                              _ptr(gout), _ptr(gpl), _ptr(gql))
    return out, gpl, gql


def kl_div_fwd(p, q, eps=1e-10):
    dt = _dt(p); p = _c(p, dt); q = _c(q, dt); b, c, hw = _bchw(p)
    out = np.empty((b,) + p.shape[2:], dtype=dt)
    getattr(lib(), "dcto_kl_div_fwd" + _suf(dt))(_ptr(p), _ptr(q), _i64(b), _int(c), _i64(hw), _real(dt, eps), _ptr(out))
    return out


def kl_div_bwd(p, q, gout, eps=1e-10):
    """Gradients of KL_div's map w.r.t. p and q under the upstream map ``gout``."""
    dt = _dt(p); p = _c(p, dt); q = _c(q, dt); gout = _c(gout, dt); b, c, hw = _bchw(p)
    gp = np.empty_like(p); gq = np.empty_like(p)
    getattr(lib(), "dcto_kl_div_bwd" + _suf(dt))(_ptr(p), _ptr(q), _i64(b), _int(c), _i64(hw), _real(dt, eps),
                                                _ptr(gout), _ptr(gp), _ptr(gq))
    return gp, gq


def l2_normalize(d):
    dt = _dt(d); out = np.array(d, dtype=dt, order="C", copy=True)
    b = out.shape[0]; m = out.size // max(b, 1)
    getattr(lib(), "dcto_l2_normalize" + _suf(dt))(_ptr(out), _i64(b), _i64(m))
    return out


def fgsm(img, grad, eps):
    dt = _dt(img); img = _c(img, dt); grad = _c(grad, dt)
    adv = np.empty_like(img); noise = np.empty_like(img)
    getattr(lib(), "dcto_fgsm" + _suf(dt))(_ptr(img), _ptr(grad), _real(dt, eps), _ptr(adv), _ptr(noise), _i64(img.size))
    return adv, noise


def vat_apply(img, d, eps):
    dt = _dt(img); img = _c(img, dt); d = _c(d, dt)
    adv = np.empty_like(img); r = np.empty_like(img)
    getattr(lib(), "dcto_vat_apply" + _suf(dt))(_ptr(img), _ptr(d), _real(dt, eps), _ptr(adv), _ptr(r), _i64(img.size))
    return adv, r


def _labels(gt, b, hw):
    g = np.ascontiguousarray(gt, dtype=np.int64).reshape(b, hw)
    return g


def cross_entropy(z, gt, weight=None, ignore_index=255, reduction="mean", upstream=1.0, gout=None,
                  want_grad=True, n_global=None):
    """CrossEntropyLoss2d(weight, ignore_index)(z, gt) (loss/loss.py:12-25) and its autograd gradient.

    reduction 'mean' | 'sum' | 'none'.  Returns (loss, grad_z, n_bad): loss is a python float for
    mean/sum and a [B,H,W] map for 'none' (then ``gout`` [B,H,W] is the upstream of the map).
    ``n_global`` replaces the 'mean' denominator W (data-parallel global mean with unit weights)."""
    dt = _dt(z); z = _c(z, dt); b, c, hw = _bchw(z); g = _labels(gt, b, hw)
    cw = None if weight is None else _c(np.asarray(weight), dt)
    f = getattr(lib(), "dcto_ce" + _suf(dt))
    sums = np.zeros(2, dtype=np.float64)
    m = np.empty((b,) + z.shape[2:], dtype=dt)
    nul = _p(None)
    bad = int(f(_ptr(z), _ptr(g), nul if cw is None else _ptr(cw), _i64(int(ignore_index)), _i64(b), _int(c), _i64(hw),
                _ptr(m), _ptr(sums), _real(dt, 1.0), nul, nul))
    W = float(sums[1]) if n_global is None else float(n_global)
    if reduction == "mean":
        with np.errstate(invalid="ignore", divide="ignore"):   # every pixel ignored: 0/0 = NaN as in ATen
            loss, gs = float(np.float64(sums[0]) / np.float64(W)), float(np.float64(upstream) / np.float64(W))
    elif reduction == "sum":
        loss, gs = sums[0], upstream
    else:
        loss, gs = m, 1.0
    gz = None
    if want_grad:
        gz = np.empty_like(z)
        go = None if (reduction != "none" or gout is None) else _c(gout, dt)
        f(_ptr(z), _ptr(g), nul if cw is None else _ptr(cw), _i64(int(ignore_index)), _i64(b), _int(c), _i64(hw),
          nul, nul, _real(dt, gs), nul if go is None else _ptr(go), _ptr(gz))
    return loss, gz, bad


def predict(x, mode="dice"):
    x = _c(x, np.float32); b, c, hw = _bchw(x)
    out = np.empty((b,) + x.shape[2:], dtype=np.int64)
    lib().dcto_predict(_ptr(x), _i64(b), _int(c), _i64(hw), _int(0 if mode == "dice" else 1), _ptr(out))
    return out


def dice_counts(x, gt):
    """Returns (counts int64 [B,C,3] = (inter, gt, pred), n_bad_labels)."""
    x = _c(x, np.float32); b, c, hw = _bchw(x); g = _labels(gt, b, hw)
    counts = np.zeros((b, c, 3), dtype=np.int64); bad = _i64(0)
    lib().dcto_dice_counts(_ptr(x), _ptr(g), _i64(b), _int(c), _i64(hw), _ptr(counts), C.byref(bad))
    return counts, int(bad.value)


def dice_from_counts(counts):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    rows, c = counts.shape[0], counts.shape[1]
    out = np.empty((rows, c), dtype=np.float32)
    lib().dcto_dice_from_counts(_ptr(counts), _i64(rows), _int(c), _ptr(out))
    return out


def dice(x, gt, method="2d"):
    """DiceMeter(method).add's appended rows: [B,C] ('2d') or [1,C] ('3d'), float32."""
    counts, bad = dice_counts(x, gt)
    if bad:
        raise AssertionError("labels outside [0,C)")
    if method == "3d":
        counts = counts.sum(0, keepdims=True)
    return dice_from_counts(counts)


def confusion(x, gt, C_=None):
    """IoU.add on scores [B,C,H,W] (or integer maps [B,H,W]) -> int64 [C,C] (rows = gt)."""
    x = np.asarray(x)
    if x.dtype.kind in "iu":
        assert C_ is not None
        p = np.ascontiguousarray(x, dtype=np.int64).reshape(-1)
        g = np.ascontiguousarray(gt, dtype=np.int64).reshape(-1)
        conf = np.zeros((C_, C_), dtype=np.int64)
        bad = lib().dcto_confusion_from_labels(_ptr(p), _ptr(g), _i64(p.size), _int(C_), _ptr(conf))
        if bad:
            raise AssertionError("prediction outside [0,C)")
        return conf
    x = _c(x, np.float32); b, c, hw = _bchw(x); g = _labels(gt, b, hw)
    conf = np.zeros((c, c), dtype=np.int64)
    lib().dcto_confusion_from_scores(_ptr(x), _ptr(g), _i64(b), _int(c), _i64(hw), _ptr(conf))
    return conf


def iou_value(conf):
    """IoU.value() -- generalframework/metrics/iou.py:96-113, numpy float64, NaN-aware."""
    hist = np.asarray(conf)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        valid = hist.sum(axis=1) > 0
        mean_iu = np.nanmean(iu[valid])
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return {"Overall_Acc": acc, "Mean_Acc": acc_cls, "FreqW_Acc": fwavacc,
            "Validated_Mean_IoU": mean_iu, "Mean_IoU": np.nanmean(iu), "Class_IoU": iu.astype(np.float32)}
