// dct_jsd_tma.cuh -- tile-pipelined JSD kernel: the TMA engine streams class-major tiles through
// shared memory while the warps compute.
//
// One persistent CTA per SM walks over pixel tiles (TP consecutive pixels of one image).  For each
// tile the K*C planes' TP-pixel row segments (contiguous in NCHW) are fetched by 1-D bulk copies
// (cp.async.bulk -> SASS UBLKCP) into a [K*C][TP] shared-memory stage, completion signalled on an
// mbarrier; STAGES-1 tiles are always in flight, so HBM latency is hidden by shared-memory depth
// instead of by resident warps/registers.  Threads read their pixels from the stage (conflict-free,
// row stride TP), do the per-pixel math in registers, write the gradients back IN PLACE into the
// stage, and one thread drains the stage to global memory with bulk stores (async proxy), after which
// the stage is refilled.  Loads and stores never pass through the register file of a stalled warp.
#pragma once
#include "dct_jsd_kernels.cuh"
#include "dct_tma.cuh"

namespace dct {

template <int K, int C, int PPT, int THREADS, int STAGES>
struct JsdTmaCfg {
    static constexpr int TP = THREADS * PPT;   // pixels per tile
    static constexpr int ROWS = K * C;
    static constexpr size_t kStageBytes = (size_t)ROWS * TP * 4;
    static constexpr size_t kSmemBytes = kStageBytes * STAGES + 8 * STAGES + 128;
};

template <int K, int C, int PPT, int THREADS, int STAGES, bool LOGITS, int MODE, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) jsd_tma_kernel(const JsdArgs<K> a, const int tiles_per_image, const int num_tiles) {
    using Cfg = JsdTmaCfg<K, C, PPT, THREADS, STAGES>;
    constexpr int TP = Cfg::TP, ROWS = Cfg::ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kStageBytes * STAGES);
    const int tid = threadIdx.x;
    const int64_t HW = a.HW;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();

    auto issue_load = [&](int tile, int stage) {  // called by thread 0 only
        const int b = tile / tiles_per_image;
        const int64_t off = (int64_t)(tile - b * tiles_per_image) * TP;
        const int64_t rem = HW - off;
        const uint32_t bytes = (uint32_t)((rem < TP ? rem : TP) * 4);
        float* dst = stages + (size_t)stage * ROWS * TP;
        tma::mbar_expect_tx(&full[stage], bytes * ROWS);
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c)
                tma::bulk_load(dst + (k * C + c) * TP, a.v.in[k] + ((int64_t)b * C + c) * HW + off, bytes, &full[stage]);
    };

    const int first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            const int t = first + s * stride;
            if (t < num_tiles) issue_load(t, s);
        }
    }

    float gs = a.up.gconst;
    gs = div_by_K<K>(gs);
    double acc = 0.0;
    bool bad = false;
    int it = 0;
    for (int tile = first; tile < num_tiles; tile += stride, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        const int b = tile / tiles_per_image;
        const int64_t off = (int64_t)(tile - b * tiles_per_image) * TP;
        const int64_t rem = HW - off;
        const int len = (int)(rem < TP ? rem : TP);
        float* st = stages + (size_t)stage * ROWS * TP;
        tma::mbar_wait(&full[stage], parity);
        const int p0 = tid * PPT;
        if (p0 < len) {
            FVec<PPT> xin[K][C];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) xin[k][c] = *reinterpret_cast<const FVec<PPT>*>(st + (k * C + c) * TP + p0);
            FVec<PPT> mapv;
            float part = 0.0f;
#pragma unroll
            for (int v = 0; v < PPT; ++v) {
                float x[K][C];
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c) x[k][c] = xin[k][c].v[v];
                float j = jsd_pixel<K, C, LOGITS, MODE != kFwd>(x, gs, bad);
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
            if (a.map != nullptr) st_stream<PPT>(a.map + (int64_t)b * HW + off + p0, mapv);
            if constexpr (MODE != kFwd) {
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c) *reinterpret_cast<FVec<PPT>*>(st + (k * C + c) * TP + p0) = xin[k][c];
            }
        }
        if constexpr (MODE != kFwd) tma::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            if constexpr (MODE != kFwd) {
                const uint32_t bytes = (uint32_t)len * 4u;
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        tma::bulk_store(a.v.grad[k] + ((int64_t)b * C + c) * HW + off, st + (k * C + c) * TP, bytes);
                tma::bulk_commit();
            }
            const int next = tile + (STAGES - 1) * stride;
            if (next < num_tiles) {
                // the stage being refilled was drained by the store group committed one iteration ago
                if constexpr (MODE != kFwd) tma::bulk_wait_read<1>();
                issue_load(next, (it + STAGES - 1) % STAGES);
            }
        }
    }
    if constexpr (MODE != kFwd) {
        if (tid == 0) tma::bulk_wait_all<0>();
    }
    if constexpr (!LOGITS) {
        if (a.flags != nullptr && __syncthreads_or(bad)) {
            if (bad) atomicAdd(&a.flags[DCT_FLAG_SIMPLEX], 1);
        }
    }
    grid_sum_to(acc, a.ws, a.sum, blockIdx.x, gridDim.x);
}

}  // namespace dct
