// dct_onehot.cu -- class maps, one-hot tensors, functional Dice on one-hot inputs and ensemble voting.
//
// Replaces the tensor helpers the hot path and its callers go through in the reference
// (generalframework/utils/utils.py): pred2class :73-80, probs2class :178-184, class2one_hot :187-198,
// probs2one_hot :201-207, predlogit2one_hot :210-217, one_hot :154-161, intersection :164-168,
// meta_dice / dice_coef / dice_batch :221-235 (the supervised baseline's Dice, trainer/trainer.py:171-175),
// and the evaluation script's Ensembleway._softVoting / _hardVoting (Summary.py:88-120).
// Every predicate of the reference (simplex, sset(...,[0,1]), one_hot) costs a full pass plus a host copy
// for torch.unique; here it is a device-side flag raised by the pass that does the work.
//
// All kernels: grid = (groups, B), a thread handles VEC consecutive pixels of one image, the class planes of
// a [B,C,HW] tensor are walked with coalesced VEC*4-byte streaming accesses.  Integer results are exact.
#include "dct_common.cuh"

namespace dct {

template <int VEC>
struct IVec;
template <>
struct alignas(4) IVec<1> { int v[1]; };
template <>
struct alignas(16) IVec<4> { int v[4]; };

template <int VEC>
__device__ __forceinline__ IVec<VEC> ld_stream_i(const int32_t* p) {
    IVec<VEC> r;
    if constexpr (VEC == 4) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r.v[0]) : "l"(p));
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void st_stream_i(int32_t* p, const IVec<VEC>& r) {
    if constexpr (VEC == 4) {
        asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]) : "memory");
    } else {
        asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(r.v[0]) : "memory");
    }
}

// outputs shared by the class-map producers (each nullable)
struct ClassOut {
    int64_t* cls;      // [B,HW] int64  (pred2class / probs2class)
    uint8_t* cls_u8;   // [B,HW] uint8  (save_images' .astype(np.uint8), utils.py:250)
    int32_t* onehot;   // [B,C,HW] int32 (class2one_hot's dtype, utils.py:194)
    float* onehot_f;   // [B,C,HW] float (Ensembleway._hardVoting returns .float(), Summary.py:120)
};

template <int VEC>
__device__ __forceinline__ void write_class(const ClassOut& o, int C, int64_t HW, int64_t b, int64_t i, const int (&cls)[VEC]) {
    if (o.cls != nullptr) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) o.cls[b * HW + i + v] = cls[v];
    }
    if (o.cls_u8 != nullptr) {
        if constexpr (VEC == 4) {
            const unsigned int w = (unsigned)(cls[0] & 255) | ((unsigned)(cls[1] & 255) << 8) | ((unsigned)(cls[2] & 255) << 16) |
                                   ((unsigned)(cls[3] & 255) << 24);
            *reinterpret_cast<unsigned int*>(o.cls_u8 + b * HW + i) = w;
        } else {
            o.cls_u8[b * HW + i] = (uint8_t)cls[0];
        }
    }
    if (o.onehot != nullptr) {
        for (int c = 0; c < C; ++c) {
            IVec<VEC> r;
#pragma unroll
            for (int v = 0; v < VEC; ++v) r.v[v] = (cls[v] == c) ? 1 : 0;
            st_stream_i<VEC>(o.onehot + (b * C + c) * HW + i, r);
        }
    }
    if (o.onehot_f != nullptr) {
        for (int c = 0; c < C; ++c) {
            FVec<VEC> r;
#pragma unroll
            for (int v = 0; v < VEC; ++v) r.v[v] = (cls[v] == c) ? 1.0f : 0.0f;
            st_stream<VEC>(o.onehot_f + (b * C + c) * HW + i, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scores [B,C,HW] -> class map / one-hot.
//   MODE 0: raw arg-max, torch.max / torch.argmax semantics (first index on ties, NaN maximal)
//   MODE 1: arg-max of softmax(x) under the pinned Dice arithmetic (predlogit2one_hot)
//   MODE 2: MODE 0 + the simplex predicate on the class sum (probs2class / probs2one_hot assert it)
// ---------------------------------------------------------------------------------------------
template <int VEC, int MODE>
__global__ void __launch_bounds__(256) classmap_kernel(const float* x, int C, int64_t HW, ClassOut o, int32_t* flags) {
    const int64_t b = blockIdx.y;
    const float* xb = x + b * C * HW;
    const int64_t gpi = HW / VEC;
    int nbad = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < gpi; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * VEC;
        float m[VEC], m2[VEC], s[VEC];
        int best[VEC];
        bool locked[VEC];
        {
            const FVec<VEC> x0 = ld_stream<VEC>(xb + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                m[v] = x0.v[v]; s[v] = x0.v[v]; best[v] = 0; locked[v] = (m[v] != m[v]);
                m2[v] = -__int_as_float(0x7f800000);
            }
        }
        for (int c = 1; c < C; ++c) {
            const FVec<VEC> xc = ld_stream<VEC>(xb + (int64_t)c * HW + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float val = xc.v[v];
                s[v] += val;
                if constexpr (MODE == 1) {
                    if (val > m[v]) { m2[v] = m[v]; m[v] = val; best[v] = c; }
                    else if (val > m2[v]) m2[v] = val;  // runner-up (a tie with the max included)
                } else {
                    const bool take = !locked[v] && ((val != val) || (val > m[v]));
                    if (take) { m[v] = val; best[v] = c; }
                    locked[v] |= (val != val);
                }
            }
        }
        if constexpr (MODE == 1) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // same fast-path criterion as spec_softmax_argmax (dct_common.cuh): a unique class within 2^-15 of the max
                const float t = m[v] - 3.0517578125e-05f;
                const bool fast = !(m2[v] >= t) & (fabsf(s[v]) <= 3.4028234664e38f);
                if (!fast) best[v] = spec_softmax_argmax_rt(xb + i + v, C, HW);
            }
        }
        if constexpr (MODE == 2) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) nbad += !simplex_ok(s[v]);
        }
        write_class<VEC>(o, C, HW, b, i, best);
    }
    if constexpr (MODE == 2) {
        nbad = __reduce_add_sync(0xffffffffu, nbad);
        if ((threadIdx.x & 31) == 0 && nbad != 0 && flags != nullptr) atomicAdd(&flags[DCT_FLAG_SIMPLEX], nbad);
    }
}

// int64 labels [B,HW] -> int32 one-hot [B,C,HW]; labels outside [0,C) raise the label flag and give an all-zero column
template <int VEC>
__global__ void __launch_bounds__(256) onehot_labels_kernel(const int64_t* labels, int C, int64_t HW, ClassOut o, int32_t* flags) {
    const int64_t b = blockIdx.y;
    const int64_t gpi = HW / VEC;
    int nbad = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < gpi; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * VEC;
        long long lab[VEC];
        ld_labels<VEC>(labels + b * HW + i, lab);
        int cls[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const bool valid = (lab[v] >= 0) & (lab[v] < C);
            nbad += !valid;
            cls[v] = valid ? (int)lab[v] : -1;
        }
        write_class<VEC>(o, C, HW, b, i, cls);
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0 && nbad != 0 && flags != nullptr) atomicAdd(&flags[DCT_FLAG_LABEL], nbad);
}

// ---------------------------------------------------------------------------------------------
// meta_dice on one-hot int32 inputs: counts[b][c] = (|label & pred|, |label|, |pred|) and the one_hot
// predicate of both tensors (values in {0,1}, exactly one 1 per pixel).  `pred` may be null (predicate of
// `label` only; the G column is still counted).  Loop shape is warp-uniform: per class the 32 lanes'
// counts are reduced with REDUX and lane 0 adds them to the warp's private shared-memory slice.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) onehot_dice_kernel(const int32_t* label, const int32_t* pred, int C, int64_t HW,
                                                          unsigned long long* counts, int32_t* flags) {
    extern __shared__ int s_cnt[];  // [warps][C*3]
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int* mine = s_cnt + wid * C * 3;
    for (int j = lane; j < C * 3; j += 32) mine[j] = 0;
    __syncwarp();
    const int32_t* lb = label + b * C * HW;
    const int32_t* pb = pred != nullptr ? pred + b * C * HW : nullptr;
    const int64_t gpi = HW / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int nbad = 0;
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + wid * 32; g0 < gpi; g0 += stride) {
        const int64_t g = g0 + lane;
        const bool active = g < gpi;
        const int64_t i = (active ? g : g0) * VEC;
        int sumL[VEC], sumP[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { sumL[v] = 0; sumP[v] = 0; }
        for (int c = 0; c < C; ++c) {
            IVec<VEC> a = ld_stream_i<VEC>(lb + (int64_t)c * HW + i), p;
            if (pb != nullptr) p = ld_stream_i<VEC>(pb + (int64_t)c * HW + i);
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) p.v[v] = 0;
            }
            int cI = 0, cG = 0, cP = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                nbad += active & (((unsigned)a.v[v] > 1u) | ((unsigned)p.v[v] > 1u));
                sumL[v] += a.v[v]; sumP[v] += p.v[v];
                cI += a.v[v] & p.v[v]; cG += a.v[v]; cP += p.v[v];   // einsum of the int32 tensors themselves
            }
            if (!active) { cI = 0; cG = 0; cP = 0; }
            cI = __reduce_add_sync(0xffffffffu, cI);
            cG = __reduce_add_sync(0xffffffffu, cG);
            cP = __reduce_add_sync(0xffffffffu, cP);
            if (lane == 0) { mine[c * 3] += cI; mine[c * 3 + 1] += cG; mine[c * 3 + 2] += cP; }
        }
        if (active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) nbad += (sumL[v] != 1) | ((pb != nullptr) & (sumP[v] != 1));
        }
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if (lane == 0 && nbad != 0 && flags != nullptr) atomicAdd(&flags[DCT_FLAG_ONEHOT], nbad);
    __syncthreads();
    if (counts != nullptr) {
        for (int j = threadIdx.x; j < C * 3; j += blockDim.x) {
            long long t = 0;
            for (int w = 0; w < nw; ++w) t += s_cnt[w * C * 3 + j];
            if (t) atomicAdd(&counts[b * C * 3 + j], (unsigned long long)t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Ensemble voting over K views (Summary.py:88-120).
//   soft: out[b,c,i] = ((x_0 + x_1) + ... + x_{K-1}) / K   (torch.stack(...).mean(0)), class map = raw arg-max of it
//   hard: per view raw arg-max (pred.max(1)[1]), then np.bincount(votes).argmax() = the most voted class, the
//         SMALLEST class on ties; one-hot of the winner.
// ---------------------------------------------------------------------------------------------
struct VoteViews {
    const float* in[DCT_MAX_VIEWS];
};

template <int VEC>
__global__ void __launch_bounds__(256) vote_soft_kernel(VoteViews vw, int K, int C, int64_t HW, float* mean, ClassOut o) {
    const int64_t b = blockIdx.y;
    const int64_t gpi = HW / VEC;
    const float Kf = (float)K;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < gpi; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * VEC;
        float m[VEC];
        int best[VEC];
        bool locked[VEC];
        for (int c = 0; c < C; ++c) {
            const int64_t off = (b * C + c) * HW + i;
            FVec<VEC> acc = ld_stream<VEC>(vw.in[0] + off);
            for (int k = 1; k < K; ++k) {
                const FVec<VEC> t = ld_stream<VEC>(vw.in[k] + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] = __fadd_rn(acc.v[v], t.v[v]);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc.v[v] = __fdiv_rn(acc.v[v], Kf);
                const float val = acc.v[v];
                if (c == 0) { m[v] = val; best[v] = 0; locked[v] = (val != val); }
                else {
                    const bool take = !locked[v] && ((val != val) || (val > m[v]));
                    if (take) { m[v] = val; best[v] = c; }
                    locked[v] |= (val != val);
                }
            }
            if (mean != nullptr) st_stream<VEC>(mean + off, acc);
        }
        write_class<VEC>(o, C, HW, b, i, best);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) vote_hard_kernel(VoteViews vw, int K, int C, int64_t HW, ClassOut o) {
    const int64_t b = blockIdx.y;
    const int64_t gpi = HW / VEC;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < gpi; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * VEC;
        int votes[DCT_MAX_VIEWS][VEC];
#pragma unroll
        for (int k = 0; k < DCT_MAX_VIEWS; ++k) {
            if (k < K) {
                const float* xb = vw.in[k] + b * C * HW + i;
                float m[VEC];
                bool locked[VEC];
                const FVec<VEC> x0 = ld_stream<VEC>(xb);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { m[v] = x0.v[v]; votes[k][v] = 0; locked[v] = (m[v] != m[v]); }
                for (int c = 1; c < C; ++c) {
                    const FVec<VEC> xc = ld_stream<VEC>(xb + (int64_t)c * HW);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const float val = xc.v[v];
                        const bool take = !locked[v] && ((val != val) || (val > m[v]));
                        if (take) { m[v] = val; votes[k][v] = c; }
                        locked[v] |= (val != val);
                    }
                }
            }
        }
        int win[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            int bc = 0, bn = -1;
#pragma unroll
            for (int k = 0; k < DCT_MAX_VIEWS; ++k) {
                if (k < K) {
                    int n = 0;
#pragma unroll
                    for (int j = 0; j < DCT_MAX_VIEWS; ++j) n += (j < K) & (votes[j][v] == votes[k][v]);
                    const int cand = votes[k][v];
                    if (n > bn || (n == bn && cand < bc)) { bn = n; bc = cand; }
                }
            }
            win[v] = bc;
        }
        write_class<VEC>(o, C, HW, b, i, win);
    }
}

static dim3 pixel_grid(int64_t B, int64_t HW, int vec, int per_thread) {
    int64_t gx = (HW / vec + 256LL * per_thread - 1) / (256LL * per_thread);
    if (gx < 1) gx = 1;
    if (gx > 65535) gx = 65535;
    return dim3((unsigned)gx, (unsigned)B, 1);
}

static bool vec4_ok(int64_t HW, std::initializer_list<const void*> ptrs) {
    if ((HW % 4) != 0) return false;
    for (const void* p : ptrs)
        if (p != nullptr && !aligned(p, 16)) return false;
    return true;
}

static int classout_validate(const ClassOut& o) {
    if (o.cls == nullptr && o.cls_u8 == nullptr && o.onehot == nullptr && o.onehot_f == nullptr) return DCT_ERR_BAD_ARG;
    if (!aligned(o.cls, 8) || !aligned(o.onehot, 4) || !aligned(o.onehot_f, 4)) return DCT_ERR_MISALIGNED;
    return DCT_OK;
}

}  // namespace dct

using namespace dct;

extern "C" int dct_classmap_f32(const float* x, int C, int64_t B, int64_t HW, int mode, int64_t* cls, uint8_t* cls_u8,
                                int32_t* onehot, int32_t* flags, void* stream) {
    if (x == nullptr || C < 1 || B < 1 || HW < 1 || mode < 0 || mode > 2) return DCT_ERR_BAD_ARG;
    if (C > DCT_MAX_CLASSES || B > 65535) return DCT_ERR_UNSUPPORTED;
    ClassOut o{cls, cls_u8, onehot, nullptr};
    int rc = classout_validate(o);
    if (rc != DCT_OK) return rc;
    if (!aligned(x, 4)) return DCT_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool v4 = vec4_ok(HW, {x, cls, onehot}) && aligned(cls_u8, 4);
    const dim3 grid = pixel_grid(B, HW, v4 ? 4 : 1, 2);
#define DCT_LAUNCH_CLASSMAP(V)                                                                     \
    switch (mode) {                                                                                \
        case 0: classmap_kernel<V, 0><<<grid, 256, 0, s>>>(x, C, HW, o, flags); break;             \
        case 1: classmap_kernel<V, 1><<<grid, 256, 0, s>>>(x, C, HW, o, flags); break;             \
        default: classmap_kernel<V, 2><<<grid, 256, 0, s>>>(x, C, HW, o, flags); break;            \
    }
    if (v4) { DCT_LAUNCH_CLASSMAP(4) } else { DCT_LAUNCH_CLASSMAP(1) }
#undef DCT_LAUNCH_CLASSMAP
    return check_launch();
}

extern "C" int dct_onehot_from_labels_i64(const int64_t* labels, int C, int64_t B, int64_t HW, int32_t* onehot,
                                          int32_t* flags, void* stream) {
    if (labels == nullptr || onehot == nullptr || C < 1 || B < 1 || HW < 1) return DCT_ERR_BAD_ARG;
    if (C > DCT_MAX_CLASSES || B > 65535) return DCT_ERR_UNSUPPORTED;
    if (!aligned(labels, 8) || !aligned(onehot, 4)) return DCT_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ClassOut o{nullptr, nullptr, onehot, nullptr};
    if (vec4_ok(HW, {labels, onehot})) onehot_labels_kernel<4><<<pixel_grid(B, HW, 4, 2), 256, 0, s>>>(labels, C, HW, o, flags);
    else onehot_labels_kernel<1><<<pixel_grid(B, HW, 1, 2), 256, 0, s>>>(labels, C, HW, o, flags);
    return check_launch();
}

extern "C" int dct_onehot_dice_counts_i32(const int32_t* label_onehot, const int32_t* pred_onehot, int C, int64_t B,
                                          int64_t HW, int64_t* counts, int32_t* flags, void* stream) {
    if (label_onehot == nullptr || C < 1 || B < 1 || HW < 1) return DCT_ERR_BAD_ARG;
    if (C > DCT_MAX_CLASSES || B > 65535) return DCT_ERR_UNSUPPORTED;
    if (!aligned(label_onehot, 4) || !aligned(pred_onehot, 4) || !aligned(counts, 8)) return DCT_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (counts != nullptr) {
        cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)(B * C * 3), s);
        if (e != cudaSuccess) { g_last_cuda_error = e; return DCT_ERR_CUDA; }
    }
    const size_t smem = sizeof(int) * 8 * C * 3;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(counts);
    if (vec4_ok(HW, {label_onehot, pred_onehot}))
        onehot_dice_kernel<4><<<pixel_grid(B, HW, 4, 4), 256, smem, s>>>(label_onehot, pred_onehot, C, HW, out, flags);
    else
        onehot_dice_kernel<1><<<pixel_grid(B, HW, 1, 4), 256, smem, s>>>(label_onehot, pred_onehot, C, HW, out, flags);
    return check_launch();
}

extern "C" int dct_vote_f32(const float* const* views, int K, int C, int64_t B, int64_t HW, int hard, float* out,
                            int64_t* cls, uint8_t* cls_u8, void* stream) {
    if (views == nullptr || K < 1 || C < 1 || B < 1 || HW < 1) return DCT_ERR_BAD_ARG;
    if (K > DCT_MAX_VIEWS || C > DCT_MAX_CLASSES || B > 65535) return DCT_ERR_UNSUPPORTED;
    if (out == nullptr && cls == nullptr && cls_u8 == nullptr) return DCT_ERR_BAD_ARG;
    VoteViews vw{};
    bool v4 = vec4_ok(HW, {out, cls}) && aligned(cls_u8, 4);
    for (int k = 0; k < K; ++k) {
        if (views[k] == nullptr) return DCT_ERR_BAD_ARG;
        if (!aligned(views[k], 4)) return DCT_ERR_MISALIGNED;
        v4 = v4 && aligned(views[k], 16);
        vw.in[k] = views[k];
    }
    if (!aligned(out, 4) || !aligned(cls, 8)) return DCT_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid = pixel_grid(B, HW, v4 ? 4 : 1, 1);
    if (hard) {
        ClassOut o{cls, cls_u8, nullptr, out};
        if (v4) vote_hard_kernel<4><<<grid, 256, 0, s>>>(vw, K, C, HW, o);
        else vote_hard_kernel<1><<<grid, 256, 0, s>>>(vw, K, C, HW, o);
    } else {
        ClassOut o{cls, cls_u8, nullptr, nullptr};
        if (v4) vote_soft_kernel<4><<<grid, 256, 0, s>>>(vw, K, C, HW, out, o);
        else vote_soft_kernel<1><<<grid, 256, 0, s>>>(vw, K, C, HW, out, o);
    }
    return check_launch();
}
