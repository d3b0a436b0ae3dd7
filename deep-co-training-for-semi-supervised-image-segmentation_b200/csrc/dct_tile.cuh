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
//   USES_UP          consumes an upstream gradient (gconst * *gscalar * gmap[pixel])
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

template <int ROWS, int PPT, int THREADS>
constexpr int tile_stages() {
    // as many stages as fit in ~200 KB, between 2 and 8
    constexpr size_t stage = (size_t)ROWS * PPT * THREADS * 4;
    constexpr size_t n = (200 * 1024) / stage;
    return n < 2 ? 2 : (n > 8 ? 8 : (int)n);
}

template <class Op, int CT, int PPT, int THREADS, int STAGES>
struct TileCfg {
    static constexpr int TP = THREADS * PPT;
    static constexpr int ROWS = Op::NIN * CT;
    static constexpr size_t kStageBytes = (size_t)ROWS * TP * 4;
    static constexpr size_t kSmemBytes = kStageBytes * STAGES + 8 * STAGES + 128;
};

template <class Op, int CT, int PPT, int THREADS, int STAGES>
__global__ void __launch_bounds__(THREADS, 1) tile_kernel(const TileArgs a) {
    using Cfg = TileCfg<Op, CT, PPT, THREADS, STAGES>;
    constexpr int TP = Cfg::TP, ROWS = Cfg::ROWS, NIN = Op::NIN, NOUT = Op::NOUT, C = CT;
    constexpr bool DICE = Op::NDICE > 0;
    static_assert(!DICE || CT <= 4, "fused Dice counters are packed 8-bit fields: C <= 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kStageBytes * STAGES);
    __shared__ int s_cnt[DICE ? Op::NDICE * CT * 3 : 1];
    const int tid = threadIdx.x;
    const int64_t HW = a.HW;
    const int tpi = a.tiles_per_image;

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
        float* dst = stages + (size_t)stage * ROWS * TP;
        tma::mbar_expect_tx(&full[stage], bytes * ROWS);
#pragma unroll
        for (int n = 0; n < NIN; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c)
                tma::bulk_load(dst + (n * C + c) * TP, a.in[n] + ((int64_t)b * C + c) * HW + off, bytes, &full[stage]);
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
    unsigned int pk[DICE ? Op::NDICE : 1][3];  // packed 8-bit per-class counters: [view][I,G,P]
    if constexpr (DICE) {
#pragma unroll
        for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = pk[n][2] = 0u;
    }
    int cur_b = -1, since_flush = 0;
    const bool do_dice = DICE && a.labels != nullptr;

    // flush this thread's packed counters into the CTA's shared counters, then (all threads) to global
    auto flush_counts = [&](int b) {
        if constexpr (DICE) {
#pragma unroll
            for (int n = 0; n < Op::NDICE; ++n)
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        int v = (int)((pk[n][q] >> (8 * c)) & 0xffu);
                        v = __reduce_add_sync(0xffffffffu, v);
                        if ((tid & 31) == 0 && v) atomicAdd(&s_cnt[(n * C + c) * 3 + q], v);
                    }
#pragma unroll
            for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = pk[n][2] = 0u;
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
        float* st = stages + (size_t)stage * ROWS * TP;
        const int p0 = tid * PPT;
        const bool active = p0 < len;
        if constexpr (DICE) {
            if (do_dice && (b != cur_b || since_flush > 255 - PPT)) {  // uniform across the CTA
                if (cur_b >= 0) flush_counts(cur_b);
                cur_b = b;
                since_flush = 0;
            }
        }
        // small per-pixel side inputs come straight from global memory, requested before the wait
        FVec<PPT> gm;
#pragma unroll
        for (int v = 0; v < PPT; ++v) gm.v[v] = 1.0f;
        if constexpr (Op::USES_UP) {
            if (a.up.gmap != nullptr && active) gm = ld_stream<PPT>(a.up.gmap + (int64_t)b * HW + off + p0);
        }
        long long lab[PPT];
        if constexpr (DICE) {
            if (do_dice && active) ld_labels<PPT>(a.labels + (int64_t)b * HW + off + p0, lab);
        }
        tma::mbar_wait(&full[stage], parity);
        if (active) {
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
                        const long long gl = lab[v];
                        const bool valid = (gl >= 0) & (gl < C);
                        nbad_label += !valid;
#pragma unroll
                        for (int n = 0; n < Op::NDICE; ++n) {
                            const int pred = spec_softmax_argmax<C>(x[n]);
                            const unsigned int sh = 8u * (unsigned int)pred;
                            pk[n][2] += 1u << sh;
                            if (valid) {
                                pk[n][1] += 1u << (8u * (unsigned int)gl);
                                pk[n][0] += (unsigned int)(gl == pred) << sh;
                            }
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
    constexpr int PPT = 4, THREADS = 256;
    constexpr int STAGES = tile_stages<ROWS, PPT, THREADS>();
    using Cfg = TileCfg<Op, CT, PPT, THREADS, STAGES>;
    auto kern = tile_kernel<Op, CT, PPT, THREADS, STAGES>;
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
    int grid = kSMs;
    if (grid > a.num_tiles) grid = a.num_tiles;
    kern<<<grid, THREADS, Cfg::kSmemBytes, stream>>>(a);
    return check_launch();
}

}  // namespace dct
