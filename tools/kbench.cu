// kbench.cu -- developer micro-benchmark for kernel variants (not part of the product library).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I<csrc> tools/kbench.cu -o tools/kbench
// Times each JSD fwd+bwd variant with CUDA events over rotating buffer sets (inputs never in L2).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dct_jsd_tma.cuh"

using namespace dct;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int K>
struct Sets {
    std::vector<JsdArgs<K>> a;
};

__global__ void fill(float* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        p[i] = ((h & 0xffff) / 65535.0f - 0.5f) * 12.0f;
    }
}

template <int K, int C, int VEC, int MINB>
void run(const char* tag, int64_t B, int64_t HW, int threads, int reps) {
    const int R = 4;
    size_t n = (size_t)B * C * HW;
    std::vector<JsdArgs<K>> sets(R);
    Workspace* ws; CK(cudaMalloc(&ws, sizeof(Workspace))); CK(cudaMemset(ws, 0, sizeof(Workspace)));
    double* sum; CK(cudaMalloc(&sum, 8));
    for (int r = 0; r < R; ++r) {
        for (int k = 0; k < K; ++k) {
            float *in, *g;
            CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&g, n * 4));
            fill<<<1024, 256>>>(in, n, 17u * r + k);
            sets[r].v.in[k] = in; sets[r].v.grad[k] = g;
        }
        sets[r].HW = HW; sets[r].map = nullptr; sets[r].sum = sum; sets[r].up = Upstream{nullptr, nullptr, 1e-6f};
        sets[r].flags = nullptr; sets[r].ws = ws;
    }
    dim3 grid = image_grid(B, HW / VEC, threads);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) jsd_kernel<K, C, VEC, true, kFwdBwd, MINB><<<grid, threads>>>(sets[i % R]);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) jsd_kernel<K, C, VEC, true, kFwdBwd, MINB><<<grid, threads>>>(sets[i % R]);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double us = ms * 1e3 / reps;
    double bytes = (double)B * HW * (2.0 * K * C * 4);
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, jsd_kernel<K, C, VEC, true, kFwdBwd, MINB>));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jsd_kernel<K, C, VEC, true, kFwdBwd, MINB>, threads, 0));
    printf("%-10s K=%d C=%2d VEC=%d MINB=%d thr=%3d regs=%3d occ=%2d CTAs/SM grid=%5ux%u  %8.2f us  %7.1f GB/s  %6.2f Gpix/s\n", tag, K, C, VEC,
           MINB, threads, fa.numRegs, occ, grid.x, grid.y, us, bytes / us / 1e3, B * HW / us / 1e3);
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < K; ++k) { cudaFree((void*)sets[r].v.in[k]); cudaFree(sets[r].v.grad[k]); }
    cudaFree(ws); cudaFree(sum);
}

template <int K, int C, int PPT, int THREADS, int STAGES, int MINB>
void run_tma(const char* tag, int64_t B, int64_t HW, int reps) {
    using Cfg = JsdTmaCfg<K, C, PPT, THREADS, STAGES>;
    auto kern = jsd_tma_kernel<K, C, PPT, THREADS, STAGES, true, kFwdBwd, MINB>;
    const int R = 4;
    size_t n = (size_t)B * C * HW;
    std::vector<JsdArgs<K>> sets(R);
    Workspace* ws; CK(cudaMalloc(&ws, sizeof(Workspace))); CK(cudaMemset(ws, 0, sizeof(Workspace)));
    double* sum; CK(cudaMalloc(&sum, 8));
    for (int r = 0; r < R; ++r) {
        for (int k = 0; k < K; ++k) {
            float *in, *g;
            CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&g, n * 4));
            fill<<<1024, 256>>>(in, n, 17u * r + k);
            sets[r].v.in[k] = in; sets[r].v.grad[k] = g;
        }
        sets[r].HW = HW; sets[r].map = nullptr; sets[r].sum = sum; sets[r].up = Upstream{nullptr, nullptr, 1e-6f};
        sets[r].flags = nullptr; sets[r].ws = ws;
    }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    int tpi = (int)((HW + Cfg::TP - 1) / Cfg::TP);
    int num_tiles = (int)(tpi * B);
    int grid = 148 * MINB; if (grid > num_tiles) grid = num_tiles;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) kern<<<grid, THREADS, Cfg::kSmemBytes>>>(sets[i % R], tpi, num_tiles);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kern<<<grid, THREADS, Cfg::kSmemBytes>>>(sets[i % R], tpi, num_tiles);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double us = ms * 1e3 / reps;
    double bytes = (double)B * HW * (2.0 * K * C * 4);
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    // correctness vs the register kernel on set 0
    std::vector<float> g1(n), g2(n);
    kern<<<grid, THREADS, Cfg::kSmemBytes>>>(sets[0], tpi, num_tiles);
    CK(cudaMemcpy(g1.data(), sets[0].v.grad[K - 1], n * 4, cudaMemcpyDeviceToHost));
    double s1; CK(cudaMemcpy(&s1, sum, 8, cudaMemcpyDeviceToHost));
    jsd_kernel<K, C, jsd_vec<K, C>(), true, kFwdBwd, 1><<<image_grid(B, HW / jsd_vec<K, C>(), 256), 256>>>(sets[0]);
    CK(cudaMemcpy(g2.data(), sets[0].v.grad[K - 1], n * 4, cudaMemcpyDeviceToHost));
    double s2; CK(cudaMemcpy(&s2, sum, 8, cudaMemcpyDeviceToHost));
    double md = 0; for (size_t i = 0; i < n; ++i) { double d = fabs((double)g1[i] - g2[i]); if (d > md) md = d; }
    printf("%-6s TMA K=%d C=%2d PPT=%d thr=%4d stages=%d minb=%d regs=%3d smem=%3zuKB grid=%4d tiles=%5d  %8.2f us  %7.1f GB/s  %6.2f Gpix/s  maxdiff=%.2e sum %.6f/%.6f\n",
           tag, K, C, PPT, THREADS, STAGES, MINB, fa.numRegs, Cfg::kSmemBytes / 1024, grid, num_tiles, us, bytes / us / 1e3, B * HW / us / 1e3, md, s1, s2);
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < K; ++k) { cudaFree((void*)sets[r].v.in[k]); cudaFree(sets[r].v.grad[k]); }
    cudaFree(ws); cudaFree(sum);
}

int main(int argc, char** argv) {
    int reps = argc > 1 ? atoi(argv[1]) : 40;
    run_tma<3, 4, 1, 1024, 4, 1>("c2", 32, 65536, reps);
    run_tma<3, 4, 2, 512, 4, 1>("c2", 32, 65536, reps);
    run_tma<3, 4, 4, 256, 4, 1>("c2", 32, 65536, reps);
    run_tma<3, 4, 1, 512, 4, 2>("c2", 32, 65536, reps);
    run_tma<3, 4, 2, 256, 4, 2>("c2", 32, 65536, reps);
    run_tma<3, 4, 4, 256, 2, 2>("c2", 32, 65536, reps);
    run_tma<3, 4, 2, 512, 2, 2>("c2", 32, 65536, reps);
    run_tma<3, 4, 1, 512, 3, 3>("c2", 32, 65536, reps);
    run_tma<3, 4, 1, 256, 4, 4>("c2", 32, 65536, reps);
    run_tma<3, 4, 2, 256, 3, 3>("c2", 32, 65536, reps);
    run_tma<3, 4, 4, 512, 2, 1>("c2", 32, 65536, reps);
    run_tma<2, 19, 1, 256, 2, 2>("c4/4", 4, 524288, reps);
    run_tma<2, 19, 1, 512, 2, 1>("c4/4", 4, 524288, reps);
    run_tma<2, 19, 2, 256, 2, 1>("c4/4", 4, 524288, reps);
    run_tma<2, 19, 1, 256, 4, 1>("c4/4", 4, 524288, reps);
    run_tma<2, 2, 4, 256, 4, 2>("c3", 8, 262144, reps);
    run_tma<2, 2, 2, 512, 4, 2>("c3", 8, 262144, reps);
    run_tma<2, 4, 2, 512, 4, 2>("c1x8", 32, 65536, reps);
    // c2: K=3 C=4 B=32 256x256
    run<3, 4, 4, 1>("c2", 32, 65536, 256, reps);
    run<3, 4, 4, 2>("c2", 32, 65536, 256, reps);
    run<3, 4, 4, 3>("c2", 32, 65536, 256, reps);
    run<3, 4, 4, 4>("c2", 32, 65536, 256, reps);
    run<3, 4, 4, 4>("c2", 32, 65536, 128, reps);
    run<3, 4, 2, 2>("c2", 32, 65536, 256, reps);
    run<3, 4, 2, 4>("c2", 32, 65536, 256, reps);
    run<3, 4, 2, 6>("c2", 32, 65536, 256, reps);
    run<3, 4, 2, 6>("c2", 32, 65536, 128, reps);
    run<3, 4, 1, 4>("c2", 32, 65536, 256, reps);
    run<3, 4, 1, 8>("c2", 32, 65536, 256, reps);
    run<3, 4, 1, 8>("c2", 32, 65536, 128, reps);
    // c1/c3-like
    run<2, 4, 4, 2>("c1x8", 32, 65536, 256, reps);
    run<2, 4, 4, 4>("c1x8", 32, 65536, 256, reps);
    run<2, 4, 2, 6>("c1x8", 32, 65536, 256, reps);
    run<2, 2, 4, 4>("c3", 8, 262144, 256, reps);
    run<2, 2, 4, 6>("c3", 8, 262144, 256, reps);
    run<2, 2, 2, 8>("c3", 8, 262144, 256, reps);
    // c4: K=2 C=19 B=16 512x1024 (reduced B=4 to bound memory: 4 sets x 2 x 2 x 160 MB)
    run<2, 19, 2, 1>("c4/4", 4, 524288, 256, reps);
    run<2, 19, 2, 2>("c4/4", 4, 524288, 256, reps);
    run<2, 19, 1, 2>("c4/4", 4, 524288, 256, reps);
    run<2, 19, 1, 3>("c4/4", 4, 524288, 256, reps);
    run<2, 19, 1, 4>("c4/4", 4, 524288, 256, reps);
    run<2, 19, 1, 4>("c4/4", 4, 524288, 128, reps);
    return 0;
}
