// dct_jsd_kernels.cuh -- K-view Jensen-Shannon divergence, forward / backward / fused, sm_100a.
//
// Replaces JSD_2D / JSD / Entropy_2D (generalframework/loss/loss.py:53-84,165-196) and, in logits
// mode, the F.softmax that feeds them (generalframework/models/segmentators.py:46-50).
//
// Data layout: K tensors [B,C,HW] (NCHW): for a fixed (k,c) the HW pixels are contiguous, so a
// warp reading VEC consecutive pixels per lane issues one fully coalesced 128*VEC-byte request
// per (k,c) plane.  All K*C planes of a pixel group are loaded up front (K*C independent
// streaming loads in flight per thread), the per-pixel math runs entirely in registers, and the
// gradients are stored back with the same access pattern.  No shared memory, no tensor cores:
// the kernel is an HBM-bound map/reduce ("JSD kernel").
#pragma once
#include "dct_common.cuh"

namespace dct {

enum JsdMode { kFwd = 0, kBwd = 1, kFwdBwd = 2 };

template <int K>
struct JsdArgs {
    Views<K> v;
    int64_t HW;
    float* map;             // [B,HW] or null
    double* sum;            // or null
    Upstream up;            // kBwd: full upstream; kFwdBwd: gconst only
    int32_t* flags;         // or null
    Workspace* ws;
};

struct JsdArgsRt {          // runtime-(K,C) fallback
    const float* in[DCT_MAX_VIEWS];
    float* grad[DCT_MAX_VIEWS];
    int K, C;
    int64_t HW;
    float* map;
    double* sum;
    Upstream up;
    int32_t* flags;
    Workspace* ws;
};

template <int K>
__device__ __forceinline__ float div_by_K(float s) {
    if constexpr ((K & (K - 1)) == 0) return s * (1.0f / (float)K);  // exact for powers of two
    else return s / (float)K;                                        // the reference divides
}

// One pixel.  In: x[k][c] (probs or logits).  Out: returns the JSD value; if GRAD, x[k][c] is
// overwritten with gK * d JSD / d x[k][c]  (gK = upstream / K).  `bad` is set when a view fails
// the reference's simplex predicate (probs mode only).
template <int K, int C, bool LOGITS, bool GRAD>
__device__ __forceinline__ float jsd_pixel(float (&x)[K][C], float gK, bool& bad) {
    float p[K][C];
    float hsum = 0.0f;  // sum_k sum_c p*log p   (= -sum_k H_k)
    if constexpr (LOGITS) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float mx = x[k][0];
#pragma unroll
            for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[k][c]);
            float Z = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float d = x[k][c] - mx;
                float e = fexp(d);
                x[k][c] = d;
                p[k][c] = e;
                Z += e;
            }
            float inv = fdiv(1.0f, Z);
            float lZ = flog(Z);
            float hk = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float pv = p[k][c] * inv;
                float lp = x[k][c] - lZ;  // log-softmax: differs from log(p+1e-16) by < 1e-16/p, and
                p[k][c] = pv;             // only ever multiplied by p  ->  absolute error < 1e-16
                x[k][c] = lp;
                hk = fmaf(pv, lp, hk);
            }
            hsum += hk;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float s = 0.0f, hk = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float pv = x[k][c];
                s += pv;
                float lp = flog(pv + kEntEps);
                p[k][c] = pv;
                x[k][c] = lp;
                hk = fmaf(pv, lp, hk);
            }
            bad |= !simplex_ok(s);
            hsum += hk;
        }
    }
    float hm = 0.0f;  // sum_c m*log(m+eps)  (= -H(m))
    float am[C];      // per class: log(m+eps) [+ m/(m+eps) in probs mode]
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = p[0][c];
#pragma unroll
        for (int k = 1; k < K; ++k) s += p[k][c];
        float m = div_by_K<K>(s);
        float lm = flog(m + kEntEps);
        hm = fmaf(m, lm, hm);
        if constexpr (GRAD && !LOGITS) lm += fdiv(m, m + kEntEps);
        am[c] = lm;
    }
    const float jsd = div_by_K<K>(hsum) - hm;
    if constexpr (GRAD) {
        if constexpr (LOGITS) {
            // d/dz_kc = gK * p_kc * ((lp_kc - lm_c) - KL(p_k || m));  the +1 terms of
            // d(p log p)/dp cancel inside the softmax backward.
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float kl = 0.0f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float t = x[k][c] - am[c];
                    x[k][c] = t;
                    kl = fmaf(p[k][c], t, kl);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) x[k][c] = gK * p[k][c] * (x[k][c] - kl);
            }
        } else {
            // d/dp_kc = gK * [(log(p+e) + p/(p+e)) - (log(m+e) + m/(m+e))]
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float pv = p[k][c];
                    x[k][c] = gK * ((x[k][c] + fdiv(pv, pv + kEntEps)) - am[c]);
                }
        }
    }
    return jsd;
}

template <int K, int C>
constexpr int jsd_vec() {
    return K * C <= 16 ? 4 : (K * C <= 40 ? 2 : 1);
}

template <int K, int C, int VEC, bool LOGITS, int MODE>
__global__ void __launch_bounds__(256) jsd_kernel(const JsdArgs<K> a) {
    const int64_t HW = a.HW;
    const int64_t gpi = HW / VEC;  // pixel groups per image (host guarantees HW % VEC == 0)
    const int b = blockIdx.y;
    const int64_t img = (int64_t)b * C * HW;
    float gs = 0.0f;
    if constexpr (MODE == kBwd) gs = upstream_scalar(a.up);
    if constexpr (MODE == kFwdBwd) gs = a.up.gconst;
    gs = div_by_K<K>(gs);
    double acc = 0.0;
    bool bad = false;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < gpi; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * VEC;
        FVec<VEC> xin[K][C];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) xin[k][c] = ld_stream<VEC>(a.v.in[k] + img + (int64_t)c * HW + i);
        FVec<VEC> gm;
        if constexpr (MODE == kBwd) {
            if (a.up.gmap != nullptr) gm = ld_stream<VEC>(a.up.gmap + (int64_t)b * HW + i);
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) gm.v[v] = 1.0f;
            }
        }
        FVec<VEC> mapv;
        float part = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float x[K][C];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) x[k][c] = xin[k][c].v[v];
            float gK = gs;
            if constexpr (MODE == kBwd) gK *= gm.v[v];
            float j = jsd_pixel<K, C, LOGITS, MODE != kFwd>(x, gK, bad);
            mapv.v[v] = j;
            part += j;
            if constexpr (MODE != kFwd) {
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c) xin[k][c].v[v] = x[k][c];
            }
        }
        acc += (double)part;
        if constexpr (MODE != kBwd) {
            if (a.map != nullptr) st_stream<VEC>(a.map + (int64_t)b * HW + i, mapv);
        }
        if constexpr (MODE != kFwd) {
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) st_stream<VEC>(a.v.grad[k] + img + (int64_t)c * HW + i, xin[k][c]);
        }
    }
    if constexpr (MODE != kBwd) {
        if constexpr (!LOGITS) {
            if (a.flags != nullptr && __syncthreads_or(bad)) {
                // count per pixel-group is not needed: the contract is "non-zero == violated"
                if (bad) atomicAdd(&a.flags[DCT_FLAG_SIMPLEX], 1);
            }
        }
        grid_sum_to(acc, a.ws, a.sum, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
    }
}

// ---------------------------------------------------------------------------------------------
// Runtime-(K,C) fallback: one pixel per thread, three sweeps over the class axis re-reading the
// inputs (L1/L2 hits).  Correctness path for shapes without a register-tiled instantiation.
// ---------------------------------------------------------------------------------------------
template <bool LOGITS, int MODE>
__global__ void __launch_bounds__(256) jsd_kernel_rt(const JsdArgsRt a) {
    const int K = a.K, C = a.C;
    const int64_t HW = a.HW;
    const int b = blockIdx.y;
    const int64_t img = (int64_t)b * C * HW;
    const float invK = 1.0f / (float)K;
    float gs = 0.0f;
    if constexpr (MODE == kBwd) gs = upstream_scalar(a.up);
    if constexpr (MODE == kFwdBwd) gs = a.up.gconst;
    double acc = 0.0;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        float mx[DCT_MAX_VIEWS], inv[DCT_MAX_VIEWS], lZ[DCT_MAX_VIEWS], kl[DCT_MAX_VIEWS];
        float hsum = 0.0f;
#pragma unroll
        for (int k = 0; k < DCT_MAX_VIEWS; ++k) {
            mx[k] = 0.0f; inv[k] = 1.0f; lZ[k] = 0.0f; kl[k] = 0.0f;
            if (k < K) {
                const float* xb = a.in[k] + img + i;
                if constexpr (LOGITS) {
                    float m = xb[0];
                    for (int c = 1; c < C; ++c) m = fmaxf(m, xb[(int64_t)c * HW]);
                    float Z = 0.0f;
                    for (int c = 0; c < C; ++c) Z += fexp(xb[(int64_t)c * HW] - m);
                    mx[k] = m; inv[k] = fdiv(1.0f, Z); lZ[k] = flog(Z);
                } else {
                    float s = 0.0f;
                    for (int c = 0; c < C; ++c) s += xb[(int64_t)c * HW];
                    bad |= !simplex_ok(s);
                }
            }
        }
        float hm = 0.0f;
        for (int c = 0; c < C; ++c) {
            float s = 0.0f;
            float pv[DCT_MAX_VIEWS], lp[DCT_MAX_VIEWS];
#pragma unroll
            for (int k = 0; k < DCT_MAX_VIEWS; ++k) {
                pv[k] = 0.0f; lp[k] = 0.0f;
                if (k < K) {
                    float xv = a.in[k][img + (int64_t)c * HW + i];
                    if constexpr (LOGITS) { float d = xv - mx[k]; pv[k] = fexp(d) * inv[k]; lp[k] = d - lZ[k]; }
                    else { pv[k] = xv; lp[k] = flog(xv + kEntEps); }
                    s = (k == 0) ? pv[k] : s + pv[k];
                    hsum = fmaf(pv[k], lp[k], hsum);
                }
            }
            float m = ((K & (K - 1)) == 0) ? s * invK : s / (float)K;
            float lm = flog(m + kEntEps);
            hm = fmaf(m, lm, hm);
#pragma unroll
            for (int k = 0; k < DCT_MAX_VIEWS; ++k)
                if (k < K) kl[k] = fmaf(pv[k], lp[k] - lm, kl[k]);
        }
        float hs = ((K & (K - 1)) == 0) ? hsum * invK : hsum / (float)K;
        const float jsd = hs - hm;
        acc += (double)jsd;
        if constexpr (MODE != kBwd) {
            if (a.map != nullptr) a.map[(int64_t)b * HW + i] = jsd;
        }
        if constexpr (MODE != kFwd) {
            float g = gs;
            if constexpr (MODE == kBwd) { if (a.up.gmap != nullptr) g *= a.up.gmap[(int64_t)b * HW + i]; }
            float gK = ((K & (K - 1)) == 0) ? g * invK : g / (float)K;
            for (int c = 0; c < C; ++c) {
                float s = 0.0f;
                float pv[DCT_MAX_VIEWS], lp[DCT_MAX_VIEWS];
#pragma unroll
                for (int k = 0; k < DCT_MAX_VIEWS; ++k) {
                    pv[k] = 0.0f; lp[k] = 0.0f;
                    if (k < K) {
                        float xv = a.in[k][img + (int64_t)c * HW + i];
                        if constexpr (LOGITS) { float d = xv - mx[k]; pv[k] = fexp(d) * inv[k]; lp[k] = d - lZ[k]; }
                        else { pv[k] = xv; lp[k] = flog(xv + kEntEps); }
                        s = (k == 0) ? pv[k] : s + pv[k];
                    }
                }
                float m = ((K & (K - 1)) == 0) ? s * invK : s / (float)K;
                float lm = flog(m + kEntEps);
#pragma unroll
                for (int k = 0; k < DCT_MAX_VIEWS; ++k)
                    if (k < K) {
                        float gv;
                        if constexpr (LOGITS) gv = gK * pv[k] * ((lp[k] - lm) - kl[k]);
                        else gv = gK * ((lp[k] + fdiv(pv[k], pv[k] + kEntEps)) - (lm + fdiv(m, m + kEntEps)));
                        a.grad[k][img + (int64_t)c * HW + i] = gv;
                    }
            }
        }
    }
    if constexpr (MODE != kBwd) {
        if constexpr (!LOGITS) {
            if (a.flags != nullptr && __syncthreads_or(bad)) {
                if (bad) atomicAdd(&a.flags[DCT_FLAG_SIMPLEX], 1);
            }
        }
        grid_sum_to(acc, a.ws, a.sum, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch of one (K,C) instantiation; defined per K in dct_jsd_k*.cu
// ---------------------------------------------------------------------------------------------
struct JsdCall {
    const float* const* views;
    float* const* grads;
    int K, C;
    int64_t B, HW;
    int in_kind, mode;
    float* map;
    double* sum;
    Upstream up;
    int32_t* flags;
    Workspace* ws;
    cudaStream_t stream;
};

// returns DCT_ERR_UNSUPPORTED when (K,C) has no register-tiled instantiation
int jsd_launch_k2(const JsdCall& c);
int jsd_launch_k3(const JsdCall& c);
int jsd_launch_k4(const JsdCall& c);

template <int K, int C>
int jsd_launch_kc(const JsdCall& c) {
    constexpr int VEC = jsd_vec<K, C>();
    JsdArgs<K> a;
    bool al = (c.HW % VEC) == 0 && (c.map == nullptr || aligned(c.map, 4 * VEC)) &&
              (c.up.gmap == nullptr || aligned(c.up.gmap, 4 * VEC));
    for (int k = 0; k < K; ++k) {
        a.v.in[k] = c.views[k];
        a.v.grad[k] = c.grads ? c.grads[k] : nullptr;
        al = al && aligned(a.v.in[k], 4 * VEC) && (a.v.grad[k] == nullptr || aligned(a.v.grad[k], 4 * VEC));
    }
    a.HW = c.HW; a.map = c.map; a.sum = c.sum; a.up = c.up; a.flags = c.flags; a.ws = c.ws;
    const int threads = 256;
    auto go = [&](auto vec_tag) -> int {
        constexpr int V = decltype(vec_tag)::value;
        dim3 grid = image_grid(c.B, c.HW / V, threads);
#define DCT_JSD_GO(LG, MD) jsd_kernel<K, C, V, LG, MD><<<grid, threads, 0, c.stream>>>(a)
        if (c.in_kind == DCT_IN_LOGITS) {
            if (c.mode == kFwd) DCT_JSD_GO(true, kFwd);
            else if (c.mode == kBwd) DCT_JSD_GO(true, kBwd);
            else DCT_JSD_GO(true, kFwdBwd);
        } else {
            if (c.mode == kFwd) DCT_JSD_GO(false, kFwd);
            else if (c.mode == kBwd) DCT_JSD_GO(false, kBwd);
            else DCT_JSD_GO(false, kFwdBwd);
        }
#undef DCT_JSD_GO
        return check_launch();
    };
    if (al) return go(std::integral_constant<int, VEC>{});
    return DCT_ERR_UNSUPPORTED;  // odd HW / misaligned views: the caller falls back to jsd_kernel_rt
}

}  // namespace dct
