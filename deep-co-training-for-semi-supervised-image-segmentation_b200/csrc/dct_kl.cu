// dct_kl.cu -- the adversarial branch's KL family, entropy and softmax as pixelwise ops
// (dct_pixelwise.cuh) plus their C-ABI entry points (include/dct_b200.h).
#include "dct_pixelwise.cuh"

namespace dct {

// softmax statistics of one pixel: on return x[c] = x[c] - max, returns {1/Z, log Z}
template <int CM>
__device__ __forceinline__ void softmax_stats(float (&x)[CM], float (&e)[CM], int C, float& inv, float& lZ) {
    float mx = x[0];
#pragma unroll
    for (int c = 1; c < CM; ++c)
        if (c < C) mx = fmaxf(mx, x[c]);
    float Z = 0.0f;
#pragma unroll
    for (int c = 0; c < CM; ++c)
        if (c < C) {
            float d = x[c] - mx;
            float ev = fexp(d);
            x[c] = d;
            e[c] = ev;
            Z += ev;
        }
    inv = fdiv(1.0f, Z);
    lZ = flog(Z);
}

// KL_Divergence_2D.forward -- generalframework/loss/loss.py:117-134
struct KlProbFwd {
    static constexpr int NIN = 2, NOUT = 0;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = true, USES_UP = false, CHECKS_SIMPLEX = true;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float, float eps, bool& bad) {
        float yy = 0.0f, yp = 0.0f, sp = 0.0f, sy = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = x[0][c], yv = x[1][c];
                sp += pv; sy += yv;
                yy = fmaf(yv, flog(yv + eps), yy);
                yp = fmaf(yv, flog(pv + eps), yp);
            }
        bad |= !simplex_ok(sp) | !simplex_ok(sy);
        return yy - yp;
    }
};

// its derivative: d/dp = -y/(p+eps); d/dy = log(y+eps) + y/(y+eps) - log(p+eps)
template <bool WANT_Y>
struct KlProbBwd {
    static constexpr int NIN = 2, NOUT = 2;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = true;
    static constexpr bool HAS_MAP = false, USES_UP = true, CHECKS_SIMPLEX = false;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float g, float eps, bool&) {
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = x[0][c], yv = x[1][c];
                x[0][c] = -g * fdiv(yv, pv + eps);
                if constexpr (WANT_Y) x[1][c] = g * ((flog(yv + eps) + fdiv(yv, yv + eps)) - flog(pv + eps));
            }
        return 0.0f;
    }
};

// VATGenerator.kl_div_with_logit(q_logit, p_logit) -- generalframework/utils/AEGenerator.py:78-91
// (== KL_Divergence_2D_Logit(p_logit, y_logit=q_logit), loss.py:144-162).  in[0] = q_logit, in[1] = p_logit;
// out[0] = grad q_logit, out[1] = grad p_logit.
template <bool GRAD>
struct KlLogit {
    static constexpr int NIN = 2, NOUT = GRAD ? 2 : 0;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = GRAD;
    static constexpr bool HAS_MAP = true, USES_UP = GRAD, CHECKS_SIMPLEX = false;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float g, float, bool&) {
        float eq[CM], ep[CM];
        float invq, lZq, invp, lZp;
        softmax_stats<CM>(x[0], eq, C, invq, lZq);
        softmax_stats<CM>(x[1], ep, C, invp, lZp);
        float out = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float q = eq[c] * invq;
                float t = (x[0][c] - lZq) - (x[1][c] - lZp);  // log q - log p
                eq[c] = q;
                x[0][c] = t;
                out = fmaf(q, t, out);
            }
        if constexpr (GRAD) {
#pragma unroll
            for (int c = 0; c < CM; ++c)
                if (c < C) {
                    float q = eq[c];
                    x[0][c] = g * q * (x[0][c] - out);
                    x[1][c] = g * (ep[c] * invp - q);
                }
        }
        return out;
    }
};

// KL_Divergence_2D(reduce=True)(softmax(p_logit), y.detach()) + backward to p_logit in one pass
// (generalframework/trainer/cotraining_totalloss.py:391-392).  in[0] = p_logit, in[1] = y_prob; out[0] = grad p_logit.
struct KlFromLogits {
    static constexpr int NIN = 2, NOUT = 1;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = true, USES_UP = true, CHECKS_SIMPLEX = true;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float g, float eps, bool& bad) {
        float e[CM];
        float inv, lZ;
        softmax_stats<CM>(x[0], e, C, inv, lZ);
        float yy = 0.0f, yp = 0.0f, sy = 0.0f, dot = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = e[c] * inv, yv = x[1][c];
                sy += yv;
                float pe = pv + eps;
                yy = fmaf(yv, flog(yv + eps), yy);
                yp = fmaf(yv, flog(pe), yp);
                float gp = -g * fdiv(yv, pe);
                e[c] = pv;
                x[0][c] = gp;
                dot = fmaf(pv, gp, dot);
            }
        bad |= !simplex_ok(sy);
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) x[0][c] = e[c] * (x[0][c] - dot);
        return yy - yp;
    }
};

// KL_div.forward -- loss.py:99-107: sum_c -p*log(q/p + eps)
struct KlDivFwd {
    static constexpr int NIN = 2, NOUT = 0;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = true, USES_UP = false, CHECKS_SIMPLEX = true;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float, float eps, bool& bad) {
        float s = 0.0f, sp = 0.0f, sq = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = x[0][c], qv = x[1][c];
                sp += pv; sq += qv;
                s = fmaf(-pv, logf(qv / pv + eps), s);  // IEEE divide: 0/0 -> NaN exactly like the reference
            }
        bad |= !simplex_ok(sp) | !simplex_ok(sq);
        return s;
    }
};

// Entropy_2D / Entropy forward -- loss.py:53-84
struct EntropyFwd {
    static constexpr int NIN = 1, NOUT = 0;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = true, USES_UP = false, CHECKS_SIMPLEX = true;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[1][CM], int C, float, float, bool& bad) {
        float h = 0.0f, s = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = x[0][c];
                s += pv;
                h = fmaf(pv, flog(pv + kEntEps), h);
            }
        bad |= !simplex_ok(s);
        return -h;
    }
};
struct EntropyBwd {
    static constexpr int NIN = 1, NOUT = 1;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = true;
    static constexpr bool HAS_MAP = false, USES_UP = true, CHECKS_SIMPLEX = false;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[1][CM], int C, float g, float, bool&) {
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                float pv = x[0][c];
                x[0][c] = -g * (flog(pv + kEntEps) + fdiv(pv, pv + kEntEps));
            }
        return 0.0f;
    }
};

// F.softmax(x, 1) as a standalone product op uses the accurate libdevice expf and IEEE division
struct SoftmaxFwd {
    static constexpr int NIN = 1, NOUT = 1;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = false, USES_UP = false, CHECKS_SIMPLEX = false;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[1][CM], int C, float, float, bool&) {
        float mx = x[0][0];
#pragma unroll
        for (int c = 1; c < CM; ++c)
            if (c < C) mx = fmaxf(mx, x[0][c]);
        float Z = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) { x[0][c] = expf(x[0][c] - mx); Z += x[0][c]; }
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) x[0][c] = x[0][c] / Z;
        return 0.0f;
    }
};
struct SoftmaxBwd {  // in[0] = p, in[1] = gp ; out[0] = gx
    static constexpr int NIN = 2, NOUT = 1;
    static constexpr int NDICE = 0;
    static constexpr bool GMAP = false;
    static constexpr bool HAS_MAP = false, USES_UP = false, CHECKS_SIMPLEX = false;
    template <int CM>
    static __device__ __forceinline__ float apply(float (&x)[2][CM], int C, float, float, bool&) {
        float dot = 0.0f;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) dot = fmaf(x[0][c], x[1][c], dot);
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) x[0][c] = x[0][c] * (x[1][c] - dot);
        return 0.0f;
    }
};

static PixArgs make_args(const float* in0, const float* in1, float* out0, float* out1, int C, int64_t HW, float* map,
                         double* sum, Upstream up, float eps, int32_t* flags, void* ws) {
    PixArgs a;
    a.in[0] = in0; a.in[1] = in1; a.out[0] = out0; a.out[1] = out1;
    a.C = C; a.HW = HW; a.map = map; a.sum = sum; a.up = up; a.eps = eps; a.flags = flags;
    a.ws = static_cast<Workspace*>(ws);
    return a;
}

}  // namespace dct

using namespace dct;

extern "C" int dct_kl_fwd_f32(const float* p, const float* y, int C, int64_t B, int64_t HW, float eps, float* map,
                              double* sum, int32_t* flags, void* workspace, void* stream) {
    PixArgs a = make_args(p, y, nullptr, nullptr, C, HW, map, sum, Upstream{nullptr, nullptr, 1.0f}, eps, flags, workspace);
    return pix_launch<KlProbFwd>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_kl_bwd_f32(const float* p, const float* y, int C, int64_t B, int64_t HW, float eps,
                              const float* gmap, const float* gscalar, float gconst, float* grad_p, float* grad_y,
                              void* stream) {
    PixArgs a = make_args(p, y, grad_p, grad_y, C, HW, nullptr, nullptr, Upstream{gmap, gscalar, gconst}, eps, nullptr, nullptr);
    if (grad_y != nullptr) return pix_launch<KlProbBwd<true>>(a, B, static_cast<cudaStream_t>(stream));
    return pix_launch<KlProbBwd<false>>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_kl_logit_f32(const float* q_logit, const float* p_logit, int C, int64_t B, int64_t HW, float* map,
                                double* sum, int has_upstream, const float* gmap, const float* gscalar, float gconst,
                                float* grad_p_logit, float* grad_q_logit, void* workspace, void* stream) {
    PixArgs a = make_args(q_logit, p_logit, grad_q_logit, grad_p_logit, C, HW, map, sum,
                          Upstream{gmap, gscalar, gconst}, 0.0f, nullptr, workspace);
    if (has_upstream) return pix_launch<KlLogit<true>>(a, B, static_cast<cudaStream_t>(stream));
    a.out[0] = a.out[1] = nullptr;
    return pix_launch<KlLogit<false>>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_kl_from_logits_fwdbwd_f32(const float* p_logit, const float* y_prob, int C, int64_t B, int64_t HW,
                                             float eps, float gconst, float* map, double* sum, float* grad_p_logit,
                                             int32_t* flags, void* workspace, void* stream) {
    PixArgs a = make_args(p_logit, y_prob, grad_p_logit, nullptr, C, HW, map, sum, Upstream{nullptr, nullptr, gconst},
                          eps, flags, workspace);
    return pix_launch<KlFromLogits>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_kl_div_fwd_f32(const float* p, const float* q, int C, int64_t B, int64_t HW, float eps, float* map,
                                  double* sum, int32_t* flags, void* workspace, void* stream) {
    PixArgs a = make_args(p, q, nullptr, nullptr, C, HW, map, sum, Upstream{nullptr, nullptr, 1.0f}, eps, flags, workspace);
    return pix_launch<KlDivFwd>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_entropy_fwd_f32(const float* p, int C, int64_t B, int64_t HW, float* map, double* sum,
                                   int32_t* flags, void* workspace, void* stream) {
    PixArgs a = make_args(p, nullptr, nullptr, nullptr, C, HW, map, sum, Upstream{nullptr, nullptr, 1.0f}, 0.0f, flags, workspace);
    return pix_launch<EntropyFwd>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_entropy_bwd_f32(const float* p, int C, int64_t B, int64_t HW, const float* gmap,
                                   const float* gscalar, float gconst, float* grad_p, void* stream) {
    if (grad_p == nullptr) return DCT_ERR_BAD_ARG;
    PixArgs a = make_args(p, nullptr, grad_p, nullptr, C, HW, nullptr, nullptr, Upstream{gmap, gscalar, gconst}, 0.0f, nullptr, nullptr);
    return pix_launch<EntropyBwd>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_softmax_fwd_f32(const float* x, int C, int64_t B, int64_t HW, float* p, void* stream) {
    if (p == nullptr) return DCT_ERR_BAD_ARG;
    PixArgs a = make_args(x, nullptr, p, nullptr, C, HW, nullptr, nullptr, Upstream{nullptr, nullptr, 1.0f}, 0.0f, nullptr, nullptr);
    return pix_launch<SoftmaxFwd>(a, B, static_cast<cudaStream_t>(stream));
}

extern "C" int dct_softmax_bwd_f32(const float* p, const float* gp, int C, int64_t B, int64_t HW, float* gx, void* stream) {
    if (gx == nullptr) return DCT_ERR_BAD_ARG;
    PixArgs a = make_args(p, gp, gx, nullptr, C, HW, nullptr, nullptr, Upstream{nullptr, nullptr, 1.0f}, 0.0f, nullptr, nullptr);
    return pix_launch<SoftmaxBwd>(a, B, static_cast<cudaStream_t>(stream));
}
