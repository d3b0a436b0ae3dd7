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
constexpr int kTileMaxRows = 80;   // NIN*C rows per stage

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

template <int WORDS, int PPT, int CTHREADS, int MINB>
constexpr int tile_stages() {
    // as many stages as fit in this CTA's share of shared memory, between 2 and 8
    constexpr size_t stage = (size_t)WORDS * PPT * CTHREADS * 4;
    // 228 KB per SM, 1 KB reserved per CTA, < 1 KB of static shared memory + barriers + alignment slack
    constexpr size_t n = (233472 / MINB - 2048) / stage;
    return n < 2 ? 2 : (n > 8 ? 8 : (int)n);
}

template <class Op, int CT, int PPT, int CTHREADS, int STAGES>
struct TileCfg {
    static constexpr int TP = CTHREADS * PPT;                       // pixels per tile
    static constexpr int ROWS = Op::NIN * CT;                       // float data rows
    static constexpr int WORDS = tile_row_words<Op, CT>();          // 4-byte words per pixel incl. side rows
    static constexpr size_t kStageBytes = (size_t)WORDS * TP * 4;
    static constexpr int kLabelOff = ROWS * TP;                     // in floats; labels are 8-byte, TP*8 bytes
    static constexpr int kGmapOff = (ROWS + (Op::NDICE > 0 ? 2 : 0)) * TP;  // valid when Op::GMAP
    static constexpr size_t kSmemBytes = kStageBytes * STAGES + 16 * STAGES + 128;
};

// Warp-specialised persistent kernel: NCW consumer warps + 1 producer warp per CTA.
//   producer (one elected lane): issues every bulk load / store; recycles a stage when its `done`
//       mbarrier (NCW arrivals) has completed and, for ops with outputs, when the bulk store group
//       that drains it has finished reading shared memory;
//   consumers: wait on the stage's `full` mbarrier (transaction bytes), compute in registers, write
//       results back in place, fence to the async proxy, arrive on `done`.  No CTA-wide barrier in
//       the tile loop: a fast warp runs ahead by up to STAGES-1 tiles.
template <class Op, int CT, int PPT, int NCW, int STAGES, int MINB = 1>
__global__ void __launch_bounds__(NCW * 32 + 32, MINB) tile_kernel(const TileArgs a) {
    constexpr int CTHREADS = NCW * 32;
    using Cfg = TileCfg<Op, CT, PPT, CTHREADS, STAGES>;
    constexpr int TP = Cfg::TP, ROWS = Cfg::ROWS, WORDS = Cfg::WORDS, NIN = Op::NIN, NOUT = Op::NOUT, C = CT;
    constexpr bool DICE = Op::NDICE > 0;
    constexpr int LW = (PPT % 2 == 0) ? 2 : 1;   // pixels per math lane group: pairs use packed FP32x2
    constexpr int NG = PPT / LW;
    using T = typename std::conditional<LW == 2, f2, float>::type;
    static_assert(!DICE || CT <= 4, "fused Dice counters are packed 8-bit fields: C <= 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kStageBytes * STAGES);
    uint64_t* done = full + STAGES;
    __shared__ int s_cnt[DICE ? Op::NDICE * CT * 3 : 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool is_producer = tid >= CTHREADS;
    const int64_t HW = a.HW;
    const int tpi = a.tiles_per_image;
    const bool do_dice = DICE && a.labels != nullptr;
    const bool has_gmap = Op::GMAP && a.up.gmap != nullptr;

    // contiguous tile range of this CTA
    const int per = a.num_tiles / gridDim.x, extra = a.num_tiles % gridDim.x;
    const int t_begin = blockIdx.x * per + min((int)blockIdx.x, extra);
    const int n_tiles = per + ((int)blockIdx.x < extra ? 1 : 0);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(&full[s], 1);
            tma::mbar_init(&done[s], NCW);
        }
        tma::fence_barrier_init();
    }
    if constexpr (DICE) {
        for (int j = tid; j < Op::NDICE * CT * 3; j += blockDim.x) s_cnt[j] = 0;
    }
    __syncthreads();
    pdl_wait();               // the previous grid has completed and its writes are visible (no-op without PDL)

    double acc = 0.0;
    bool bad = false;

    if (is_producer) {
        if (lane == 0) {
            auto issue_load = [&](int i) {  // i-th tile of this CTA into stage i % STAGES
                const int tile = t_begin + i, stage = i % STAGES;
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
            for (int i = 0; i < STAGES && i < n_tiles; ++i) issue_load(i);
            for (int i = 0; i < n_tiles; ++i) {
                const int stage = i % STAGES;
                tma::mbar_wait(&done[stage], (uint32_t)(i / STAGES) & 1u);  // every consumer warp is through with tile i
                if constexpr (NOUT > 0) {
                    const int tile = t_begin + i;
                    const int b = tile / tpi;
                    const int64_t off = (int64_t)(tile - b * tpi) * TP;
                    const int64_t rem = HW - off;
                    const uint32_t bytes = (uint32_t)((rem < TP ? rem : TP) * 4);
                    const float* st = stages + (size_t)stage * WORDS * TP;
#pragma unroll
                    for (int n = 0; n < NOUT; ++n)
                        if (a.out[n] != nullptr) {
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                tma::bulk_store(a.out[n] + ((int64_t)b * C + c) * HW + off, st + (n * C + c) * TP, bytes);
                        }
                    tma::bulk_commit();
                    // tile i-1's store group has drained its stage once at most one group is still reading
                    if (i >= 1 && i - 1 + STAGES < n_tiles) {
                        tma::bulk_wait_read<1>();
                        issue_load(i - 1 + STAGES);
                    }
                } else {
                    if (i + STAGES < n_tiles) issue_load(i + STAGES);
                }
            }
            if constexpr (NOUT > 0) tma::bulk_wait_all<0>();
        }
        __syncwarp();
    } else {
        float gs = 1.0f;
        if constexpr (Op::USES_UP) gs = upstream_scalar(a.up);
        int nbad_label = 0;
        unsigned int pk[DICE ? Op::NDICE : 1][2];  // packed 8-bit per-class counters: [view][I,P]
        unsigned int pkG = 0u;                     // |gt == c| is the same for every view
        if constexpr (DICE) {
#pragma unroll
            for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = 0u;
        }
        int cur_b = -1, since_flush = 0;
        auto consumer_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory"); };

        // flush this thread's packed counters into the CTA's shared counters, then (all consumers) to global
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
                        if (lane == 0) {
                            if (vi) atomicAdd(&s_cnt[(n * C + c) * 3 + 0], vi);
                            if (g) atomicAdd(&s_cnt[(n * C + c) * 3 + 1], g);
                            if (vp) atomicAdd(&s_cnt[(n * C + c) * 3 + 2], vp);
                        }
                    }
                }
                pkG = 0u;
#pragma unroll
                for (int n = 0; n < Op::NDICE; ++n) pk[n][0] = pk[n][1] = 0u;
                consumer_sync();
                for (int j = tid; j < Op::NDICE * C * 3; j += CTHREADS) {
                    const int v = s_cnt[j];
                    if (v) {
                        const int n = j / (C * 3), r = j - n * C * 3;
                        atomicAdd(&a.counts[(int64_t)n * a.count_view_stride + (int64_t)b * C * 3 + r], (unsigned long long)v);
                        s_cnt[j] = 0;
                    }
                }
                consumer_sync();
            }
        };

        for (int i = 0; i < n_tiles; ++i) {
            const int tile = t_begin + i;
            const int stage = i % STAGES;
            const int b = tile / tpi;
            const int64_t off = (int64_t)(tile - b * tpi) * TP;
            const int64_t rem = HW - off;
            const int len = (int)(rem < TP ? rem : TP);
            float* st = stages + (size_t)stage * WORDS * TP;
            const int p0 = tid * PPT;
            const bool active = p0 < len;
            if constexpr (DICE) {
                if (do_dice && (b != cur_b || since_flush > 255 - PPT)) {  // uniform across the consumers
                    if (cur_b >= 0) flush_counts(cur_b);
                    cur_b = b;
                    since_flush = 0;
                }
            }
            tma::mbar_wait(&full[stage], (uint32_t)(i / STAGES) & 1u);
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
                        if constexpr (PPT % 2 == 0) {
#pragma unroll
                            for (int v = 0; v < PPT / 2; ++v) {  // two labels per 128-bit shared load
                                const uint4 q = reinterpret_cast<const uint4*>(st + Cfg::kLabelOff)[(p0 >> 1) + v];
                                lab[2 * v] = make_uint2(q.x, q.y);
                                lab[2 * v + 1] = make_uint2(q.z, q.w);
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < PPT; ++v) lab[v] = reinterpret_cast<const uint2*>(st + Cfg::kLabelOff)[p0 + v];
                        }
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
                for (int gi = 0; gi < NG; ++gi) {
                    T x[NIN][C];
#pragma unroll
                    for (int n = 0; n < NIN; ++n)
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if constexpr (LW == 2) x[n][c] = mk2(xin[n][c].v[2 * gi], xin[n][c].v[2 * gi + 1]);
                            else x[n][c] = xin[n][c].v[gi];
                        }
                    if constexpr (DICE) {
                        if (do_dice) {
#pragma unroll
                            for (int j = 0; j < LW; ++j) {
                                const uint2 lb = lab[gi * LW + j];
                                const bool valid = (lb.y == 0u) & (lb.x < (unsigned int)C);  // 0 <= int64 label < C
                                nbad_label += !valid;
                                const unsigned int gmask = valid ? (1u << (8u * (lb.x & 3u))) : 0u;  // one-hot byte of the label
                                pkG += gmask;
#pragma unroll
                                for (int n = 0; n < Op::NDICE; ++n) {
                                    float xs[C];
#pragma unroll
                                    for (int c = 0; c < C; ++c) xs[c] = vget(x[n][c], j);
                                    const unsigned int hot = spec_softmax_argmax_onehot4<C>(xs);  // one-hot byte of the prediction
                                    pk[n][1] += hot;
                                    pk[n][0] += hot & gmask;
                                }
                            }
                        }
                    }
                    T gv;
                    if constexpr (LW == 2) gv = mk2(gs * gm.v[2 * gi], gs * gm.v[2 * gi + 1]);
                    else gv = gs * gm.v[gi];
                    const T mv = Op::template apply<C, T>(x, C, gv, a.eps, bad);
                    if constexpr (LW == 2) { mapv.v[2 * gi] = vget(mv, 0); mapv.v[2 * gi + 1] = vget(mv, 1); }
                    else mapv.v[gi] = vget(mv, 0);
                    part += vhsum(mv);
#pragma unroll
                    for (int n = 0; n < NOUT; ++n)
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if constexpr (LW == 2) { xin[n][c].v[2 * gi] = vget(x[n][c], 0); xin[n][c].v[2 * gi + 1] = vget(x[n][c], 1); }
                            else xin[n][c].v[gi] = vget(x[n][c], 0);
                        }
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
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&done[stage]);
        }
        if constexpr (DICE) {
            if (do_dice) {
                if (cur_b >= 0) flush_counts(cur_b);
                nbad_label = __reduce_add_sync(0xffffffffu, nbad_label);
                if (lane == 0 && nbad_label != 0 && a.flags != nullptr) atomicAdd(&a.flags[DCT_FLAG_LABEL], nbad_label);
            }
        }
    }
    // this CTA's tiles are done: the next kernel in the stream may be scheduled onto the SMs that drain first and run
    // its prologue; its pdl_wait() still holds it until this whole grid (epilogue below included) has completed
    pdl_launch_dependents();
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
    static_assert(ROWS <= kTileMaxRows, "tile pipeline instantiations are for NIN*C <= 80");
    // Launch shapes measured on B200 (tools/kbench_tile.cu, profiles/):
    //   rows <= 8        4 pixels/thread, 8 consumer warps, 2 CTAs/SM
    //   rows <= 16       2 pixels/thread (math-heavy: ACDC K*C = 12..16), 8 warps, 2 CTAs/SM
    //   rows <= 40       Cityscapes C = 19 with 2 tensors: 2 pixels/thread keeps the packed FP32x2 math; one CTA per SM
    //                    of 4 warps (8 for read-only ops) so that a [rows][TP] stage leaves room for >= 2..5 stages
    //   rows <= 80       one pixel/thread (a pixel pair would need > 255 registers), 4 warps
    constexpr int PPT = ROWS <= 8 ? 4 : (ROWS <= 40 ? 2 : 1);
    constexpr int NCW = ROWS <= 16 ? 8 : (ROWS <= 40 ? (Op::NOUT == 0 ? 8 : 4) : 4);
    constexpr int MINB = ROWS <= 16 ? 2 : (ROWS <= 40 ? 1 : (ROWS <= 60 ? 2 : 1));
    constexpr int STAGES = tile_stages<tile_row_words<Op, CT>(), PPT, NCW * 32, MINB>();
    using Cfg = TileCfg<Op, CT, PPT, NCW * 32, STAGES>;
    auto kern = tile_kernel<Op, CT, PPT, NCW, STAGES, MINB>;
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
    cudaError_t e = launch_pdl(kern, dim3(grid), dim3(NCW * 32 + 32), Cfg::kSmemBytes, stream, a);
    if (e != cudaSuccess) { g_last_cuda_error = e; return DCT_ERR_CUDA; }
    return check_launch();
}

}  // namespace dct
