// dct_tile.cuh -- the tile pipeline: one persistent-CTA skeleton for every per-pixel op of the path.
//
//   HBM --(TMA 1-D bulk copies, mbarrier)--> smem stage [NIN*C][TP] --> registers --> math
//   HBM <--(TMA bulk stores, bulk groups)---- smem stage (results written in place) <----'
//
// A tile is TP = THREADS*PPT consecutive pixels of one image; in NCHW every (tensor, class) plane
// contributes one contiguous TP*4-byte row segment, fetched by ONE cp.async.bulk (SASS UBLKCP).
// STAGES-1 tiles are in flight per CTA at all times, so HBM latency is covered by shared-memory
// depth rather than by resident warps and registers (the register-tiled kernels top out at
// ~68% of the measured HBM peak on ACDC-sized inputs because raising occupancy forces spills).
// Results go back through the same stage with bulk stores; a stage is refilled once its store
// group has finished reading it.  Grid = min(#tiles, 148 * MINB) persistent CTAs, each owning a
// CONTIGUOUS range of tiles (so per-image integer counters rarely need flushing).
//
// Ops plug in through a small static interface (see dct_pixelwise.cuh for the same Ops on the
// register-tiled fallback):
//   NIN, NOUT        number of [B,C,HW] input / output tensors (output n reuses input n's rows)
//   HAS_MAP          produces a per-pixel scalar (map store and/or deterministic grid sum)
//   USES_UP          consumes an upstream gradient (gconst * *gscalar [* gmap[pixel] if GMAP])
//   GMAP             the upstream may carry a per-pixel map (reserves one side row per stage)
//   CHECKS_SIMPLEX   sets the simplex flag through `bad`
//   NDICE            number of leading input tensors whose Dice counts are accumulated (0 = none)
//   apply<CM>(x[NIN][CM], C, g, eps, bad) -> map value; outputs left in x[0..NOUT)
#pragma once
#include "dct_common.cuh"
#include "dct_tma.cuh"

namespace dct {

constexpr int kTileMaxTensors = 8;

struct TileArgs {
    const float* in[kTileMaxTensors];
    float* out[kTileMaxTensors];   // each nullable
    int64_t HW;
    float* map;                    // nullable
    double* sum;                   // nullable
    Upstream up;
    float eps;
    int32_t* flags;                // nullable
    Workspace* ws;
    const int64_t* labels;         // Dice: [B,HW] int64 (NDICE > 0 and non-null => count)
    unsigned long long* counts;    // Dice: [NDICE][B][C][3] (I,G,P), accumulated into
    int64_t count_view_stride;     // B*C*3
    int tiles_per_image;
    int num_tiles;
};

// float rows + optional side rows: int64 labels (Dice ops) and the upstream-gradient map (backward ops)
template <class Op, int CT>
constexpr int tile_row_words() {
    return Op::NIN * CT + (Op::NDICE > 0 ? 2 : 0) + (Op::GMAP ? 1 : 0);
}

template <int WORDS, int PPT, int THREADS, int MINB>
constexpr int tile_stages() {
    // as many stages as fit in this CTA's share of shared memory, between 2 and 8
    constexpr size_t stage = (size_t)WORDS * PPT * THREADS * 4;
    // 228 KB per SM, 1 KB reserved per CTA, ~0.7 KB of static shared memory + barriers in the kernel
    constexpr size_t n = (233472 / MINB - 1024 - 672) / stage;
    return n < 2 ? 2 : (n > 8 ? 8 : (int)n);
}

template <class Op, int CT, int PPT, int THREADS, int STAGES>
struct TileCfg {
    static constexpr int TP = THREADS * PPT;
    static constexpr int ROWS = Op::NIN * CT;                       // float data rows
    static constexpr int WORDS = tile_row_words<Op, CT>();          // 4-byte words per pixel incl. side rows
    static constexpr size_t kStageBytes = (size_t)WORDS * TP * 4;
    static constexpr int kLabelOff = ROWS * TP;                     // in floats; labels are 8-byte, TP*8 bytes
    static constexpr int kGmapOff = (ROWS + (Op::NDICE > 0 ? 2 : 0)) * TP;  // valid when Op::GMAP
    static constexpr size_t kSmemBytes = kStageBytes * STAGES + 8 * STAGES + 128;
};

template <class Op, int CT, int PPT, int THREADS, int STAGES, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) tile_kernel(const TileArgs a) {
    using Cfg = TileCfg<Op, CT, PPT, THREADS, STAGES>;
    constexpr int TP = Cfg::TP, ROWS = Cfg::ROWS, WORDS = Cfg::WORDS, NIN = Op::NIN, NOUT = Op::NOUT, C = CT;
    constexpr bool DICE = Op::NDICE > 0;
    static_assert(!DICE || CT <= 4, "fused Dice counters are packed 8-bit fields: C <= 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kStageBytes * STAGES);
    __shared__ int s_cnt[DICE ? Op::NDICE * CT * 3 : 1];
    const int tid = threadIdx.x;
    const int64_t HW = a.HW;
    const int tpi = a.tiles_per_image;
    const bool do_dice = DICE && a.labels != nullptr;
    const bool has_gmap = Op::GMAP && a.up.gmap != nullptr;

    // contiguous tile range of this CTA
    const int per = a.num_tiles / gridDim.x, extra = a.num_tiles % gridDim.x;
    const int t_begin = blockIdx.x * per + min((int)blockIdx.x, extra);
    const int t_end = t_begin + per + ((int)blockIdx.x < extra ? 1 : 0);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_barrier_init();
    }
    if constexpr (DICE) {
        for (int j = tid; j < Op::NDICE * CT * 3; j += THREADS) s_cnt[j] = 0;
    }
    __syncthreads();

    auto issue_load = [&](int tile, int stage) {  // thread 0 only
        const int b = tile / tpi;
        const int64_t off = (int64_t)(tile - b * tpi) * TP;
        const int64_t rem = HW - off;
        const uint32_t bytes = (uint32_t)((rem < TP ? rem : TP) * 4);
        float* dst = stages + (size_t)stage * WORDS * TP;
        uint32_t total = bytes * ROWS;
        if constexpr (DICE) total += do_dice ? 2u * bytes : 0u;
        if constexpr (Op::GMAP) total += has_gmap ? bytes : 0u;
        tma::mbar_expect_tx(&full[stage], total);
#pragma unroll
        for (int n = 0; n < NIN; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c)
                tma::bulk_load(dst + (n * C + c) * TP, a.in[n] + ((int64_t)b * C + c) * HW + off, bytes, &full[stage]);
        if constexpr (DICE) {
            if (do_dice) tma::bulk_load(dst + Cfg::kLabelOff, a.labels + (int64_t)b * HW + off, 2u * bytes, &full[stage]);
        }
        if constexpr (Op::GMAP) {
            if (has_gmap) tma::bulk_load(dst + Cfg::kGmapOff, a.up.gmap + (int64_t)b * HW + off, bytes, &full[stage]);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (t_begin + s < t_end) issue_load(t_begin + s, s);
    }

    float gs = 1.0f;
    if constexpr (Op::USES_UP) gs = upstream_scalar(a.up);
    double acc = 0.0;
    bool bad = false;
    int nbad_label = 0;
    unsigned int pk[DICE ? Op::NDICE : 1][2];  // packed 8-bit per-class counters: [view][I,P]
    unsigned int pkG = 0u;                     // |gt == c| is the same for every view
    if constexpr (DICE) {
#pragma unroll
        for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = 0u;
    }
    int cur_b = -1, since_flush = 0;

    // flush this thread's packed counters into the CTA's shared counters, then (all threads) to global
    auto flush_counts = [&](int b) {
        if constexpr (DICE) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                int g = (int)((pkG >> (8 * c)) & 0xffu);
                g = __reduce_add_sync(0xffffffffu, g);
#pragma unroll
                for (int n = 0; n < Op::NDICE; ++n) {
                    int vi = (int)((pk[n][0] >> (8 * c)) & 0xffu);
                    int vp = (int)((pk[n][1] >> (8 * c)) & 0xffu);
                    vi = __reduce_add_sync(0xffffffffu, vi);
                    vp = __reduce_add_sync(0xffffffffu, vp);
                    if ((tid & 31) == 0) {
                        if (vi) atomicAdd(&s_cnt[(n * C + c) * 3 + 0], vi);
                        if (g) atomicAdd(&s_cnt[(n * C + c) * 3 + 1], g);
                        if (vp) atomicAdd(&s_cnt[(n * C + c) * 3 + 2], vp);
                    }
                }
            }
            pkG = 0u;
#pragma unroll
            for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = 0u;
            __syncthreads();
            for (int j = tid; j < Op::NDICE * C * 3; j += THREADS) {
                const int v = s_cnt[j];
                if (v) {
                    const int n = j / (C * 3), r = j - n * C * 3;
                    atomicAdd(&a.counts[(int64_t)n * a.count_view_stride + (int64_t)b * C * 3 + r], (unsigned long long)v);
                    s_cnt[j] = 0;
                }
            }
            __syncthreads();
        }
    };

    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        const int b = tile / tpi;
        const int64_t off = (int64_t)(tile - b * tpi) * TP;
        const int64_t rem = HW - off;
        const int len = (int)(rem < TP ? rem : TP);
        float* st = stages + (size_t)stage * WORDS * TP;
        const int p0 = tid * PPT;
        const bool active = p0 < len;
        if constexpr (DICE) {
            if (do_dice && (b != cur_b || since_flush > 255 - PPT)) {  // uniform across the CTA
                if (cur_b >= 0) flush_counts(cur_b);
                cur_b = b;
                since_flush = 0;
            }
        }
        tma::mbar_wait(&full[stage], parity);
        if (active) {
            FVec<PPT> gm;
#pragma unroll
            for (int v = 0; v < PPT; ++v) gm.v[v] = 1.0f;
            if constexpr (Op::GMAP) {
                if (has_gmap) gm = *reinterpret_cast<const FVec<PPT>*>(st + Cfg::kGmapOff + p0);
            }
            uint2 lab[PPT];  // int64 labels as (lo, hi) words
            if constexpr (DICE) {
                if (do_dice) {
#pragma unroll
                    for (int v = 0; v < PPT; ++v) lab[v] = reinterpret_cast<const uint2*>(st + Cfg::kLabelOff)[p0 + v];
                }
            }
            FVec<PPT> xin[NIN][C];
#pragma unroll
            for (int n = 0; n < NIN; ++n)
#pragma unroll
                for (int c = 0; c < C; ++c) xin[n][c] = *reinterpret_cast<const FVec<PPT>*>(st + (n * C + c) * TP + p0);
            FVec<PPT> mapv;
            float part = 0.0f;
#pragma unroll
            for (int v = 0; v < PPT; ++v) {
                float x[NIN][C];
#pragma unroll
                for (int n = 0; n < NIN; ++n)
#pragma unroll
                    for (int c = 0; c < C; ++c) x[n][c] = xin[n][c].v[v];
                if constexpr (DICE) {
                    if (do_dice) {
                        const unsigned int gl = lab[v].x;
                        const bool valid = (lab[v].y == 0u) & (gl < (unsigned int)C);  // 0 <= int64 label < C
                        nbad_label += !valid;
                        const unsigned int gmask = valid ? (1u << (8u * (gl & 3u))) : 0u;  // one-hot byte of the label
                        pkG += gmask;
#pragma unroll
                        for (int n = 0; n < Op::NDICE; ++n) {
                            const unsigned int hot = spec_softmax_argmax_onehot4<C>(x[n]);  // one-hot byte of the prediction
                            pk[n][1] += hot;
                            pk[n][0] += hot & gmask;
                        }
                    }
                }
                const float mv = Op::template apply<C>(x, C, gs * gm.v[v], a.eps, bad);
                mapv.v[v] = mv;
                part += mv;
#pragma unroll
                for (int n = 0; n < NOUT; ++n)
#pragma unroll
                    for (int c = 0; c < C; ++c) xin[n][c].v[v] = x[n][c];
            }
            acc += (double)part;
            if constexpr (Op::HAS_MAP) {
                if (a.map != nullptr) st_stream<PPT>(a.map + (int64_t)b * HW + off + p0, mapv);
            }
#pragma unroll
            for (int n = 0; n < NOUT; ++n)
                if (a.out[n] != nullptr) {
#pragma unroll
                    for (int c = 0; c < C; ++c) *reinterpret_cast<FVec<PPT>*>(st + (n * C + c) * TP + p0) = xin[n][c];
                }
        }
        since_flush += PPT;
        if constexpr (NOUT > 0) tma::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            if constexpr (NOUT > 0) {
                const uint32_t bytes = (uint32_t)len * 4u;
#pragma unroll
                for (int n = 0; n < NOUT; ++n)
                    if (a.out[n] != nullptr) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            tma::bulk_store(a.out[n] + ((int64_t)b * C + c) * HW + off, st + (n * C + c) * TP, bytes);
                    }
                tma::bulk_commit();
            }
            const int next = tile + (STAGES - 1);
            if (next < t_end) {
                // the stage being refilled was drained by the store group committed one iteration ago
                if constexpr (NOUT > 0) tma::bulk_wait_read<1>();
                issue_load(next, (it + STAGES - 1) % STAGES);
            }
        }
    }
    if constexpr (NOUT > 0) {
        if (tid == 0) tma::bulk_wait_all<0>();
    }
    if constexpr (DICE) {
        if (do_dice) {
            if (cur_b >= 0) flush_counts(cur_b);
            nbad_label = __reduce_add_sync(0xffffffffu, nbad_label);
            if ((tid & 31) == 0 && nbad_label != 0 && a.flags != nullptr) atomicAdd(&a.flags[DCT_FLAG_LABEL], nbad_label);
        }
    }
    if constexpr (Op::CHECKS_SIMPLEX) {
        if (a.flags != nullptr && __syncthreads_or(bad)) {
            if (bad) atomicAdd(&a.flags[DCT_FLAG_SIMPLEX], 1);
        }
    }
    if constexpr (Op::HAS_MAP) grid_sum_to(acc, a.ws, a.sum, blockIdx.x, gridDim.x);
}

// Host side: does this problem fit the tile pipeline?  (16-byte aligned rows and segments)
template <class Op>
inline bool tile_eligible(const TileArgs& a, int64_t B) {
    if ((a.HW % 4) != 0 || B * ((a.HW + 255) / 256) > 0x7fffffffLL) return false;
    for (int n = 0; n < Op::NIN; ++n)
        if (!aligned(a.in[n], 16)) return false;
    for (int n = 0; n < Op::NOUT; ++n)
        if (a.out[n] != nullptr && !aligned(a.out[n], 16)) return false;
    if (a.map != nullptr && !aligned(a.map, 16)) return false;
    if (a.up.gmap != nullptr && !aligned(a.up.gmap, 16)) return false;
    if (a.labels != nullptr && !aligned(a.labels, 16)) return false;
    return true;
}

template <class Op, int CT>
int tile_launch_ct(TileArgs a, int64_t B, cudaStream_t stream) {
    constexpr int ROWS = Op::NIN * CT;
    static_assert(ROWS <= 16, "tile pipeline instantiations are for NIN*C <= 16 (larger: register-tiled kernels)");
    // measured on B200 (tools/kbench_tile.cu): two 256-thread CTAs per SM beat one larger CTA for every op
    // (the per-tile CTA barrier of one overlaps the math of the other); math-heavy ops take 2 pixels/thread.
    constexpr int THREADS = 256, MINB = 2;
    constexpr int PPT = ROWS > 8 ? 2 : 4;
    constexpr int STAGES = tile_stages<tile_row_words<Op, CT>(), PPT, THREADS, MINB>();
    using Cfg = TileCfg<Op, CT, PPT, THREADS, STAGES>;
    auto kern = tile_kernel<Op, CT, PPT, THREADS, STAGES, MINB>;
    static bool configured[64] = {};  // per instantiation and device (the attribute is per device function)
    int devid = 0;
    cudaGetDevice(&devid);
    if (devid < 0 || devid >= 64 || !configured[devid]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
        if (e != cudaSuccess) { g_last_cuda_error = e; return DCT_ERR_CUDA; }
        if (devid >= 0 && devid < 64) configured[devid] = true;
    }
    a.tiles_per_image = (int)((a.HW + Cfg::TP - 1) / Cfg::TP);
    a.num_tiles = (int)(a.tiles_per_image * B);
    int grid = kSMs * MINB;
    if (grid > a.num_tiles) grid = a.num_tiles;
    kern<<<grid, THREADS, Cfg::kSmemBytes, stream>>>(a);
    return check_launch();
}

}  // namespace dct
