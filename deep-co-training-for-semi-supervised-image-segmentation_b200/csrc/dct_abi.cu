// dct_abi.cu -- library-level entry points of include/dct_b200.h
#include "dct_common.cuh"

#include <cstdlib>
#include <cstring>

namespace dct {
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("DCT_B200_PDL");
        return !(e != nullptr && std::strcmp(e, "0") == 0);
    }();
    return on;
}

static unsigned long long* g_trace_buf = nullptr;
static int g_trace_max_ctas = 0, g_trace_max_launches = 0, g_trace_launch = 0;
unsigned long long* trace_next(int grid) {
    if (g_trace_buf == nullptr || grid > g_trace_max_ctas || g_trace_launch >= g_trace_max_launches) return nullptr;
    return g_trace_buf + (size_t)(g_trace_launch++) * g_trace_max_ctas * kTraceSlots;
}
}  // namespace dct

using namespace dct;

extern "C" int dct_dev_trace_begin(void* buf, int max_ctas, int max_launches) {
    if (buf == nullptr || max_ctas < 1 || max_launches < 1) return DCT_ERR_BAD_ARG;
    g_trace_buf = static_cast<unsigned long long*>(buf);
    g_trace_max_ctas = max_ctas; g_trace_max_launches = max_launches; g_trace_launch = 0;
    return DCT_OK;
}
extern "C" int dct_dev_trace_end(void) {
    const int n = g_trace_launch;
    g_trace_buf = nullptr; g_trace_launch = 0;
    return n;
}

extern "C" int dct_abi_version(void) { return DCT_ABI_VERSION; }

extern "C" const char* dct_error_string(int code) {
    switch (code) {
        case DCT_OK: return "ok";
        case DCT_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size, or inconsistent options)";
        case DCT_ERR_UNSUPPORTED: return "unsupported shape (K > 8, C > 64 or B > 65535)";
        case DCT_ERR_MISALIGNED: return "misaligned pointer";
        case DCT_ERR_CUDA: return "CUDA launch failed (see dct_last_cuda_error)";
        case DCT_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        default: return "unknown error code";
    }
}

extern "C" const char* dct_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda_error); }

extern "C" int dct_device_check(int ordinal) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || ordinal < 0 || ordinal >= n) { g_last_cuda_error = e; return DCT_ERR_NO_DEVICE; }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ordinal);
    if (e != cudaSuccess || major != 10) { g_last_cuda_error = e; return DCT_ERR_NO_DEVICE; }
    return DCT_OK;
}

extern "C" size_t dct_workspace_bytes(void) { return sizeof(Workspace); }

// ---- fused exchange plumbing (include/dct_b200.h "Fused cross-rank exchange") ----
extern "C" size_t dct_peer_pub_bytes(void) { return sizeof(PeerPub); }

static_assert(sizeof(cudaIpcMemHandle_t) == DCT_IPC_HANDLE_BYTES, "CUDA IPC handles are 64 bytes");

extern "C" int dct_mailbox_create(size_t bytes, void** dev_ptr, void* ipc_handle) {
    if (bytes == 0 || dev_ptr == nullptr || ipc_handle == nullptr) return DCT_ERR_BAD_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        g_last_cuda_error = e;
        if (p != nullptr) cudaFree(p);
        (void)cudaGetLastError();
        return DCT_ERR_CUDA;
    }
    std::memcpy(ipc_handle, &h, sizeof(h));
    *dev_ptr = p;
    return DCT_OK;
}

extern "C" int dct_mailbox_open(const void* ipc_handle, void** dev_ptr) {
    if (ipc_handle == nullptr || dev_ptr == nullptr) return DCT_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { g_last_cuda_error = e; (void)cudaGetLastError(); return DCT_ERR_CUDA; }
    *dev_ptr = p;
    return DCT_OK;
}

extern "C" int dct_mailbox_close(void* dev_ptr, int owned) {
    if (dev_ptr == nullptr) return DCT_ERR_BAD_ARG;
    cudaError_t e = owned ? cudaFree(dev_ptr) : cudaIpcCloseMemHandle(dev_ptr);
    if (e != cudaSuccess) { g_last_cuda_error = e; (void)cudaGetLastError(); return DCT_ERR_CUDA; }
    return DCT_OK;
}

namespace dct {
// One thread per (value j, peer p): what cannot depend on the previous launch -- the mailbox pointers -- is fetched while
// that launch is still draining; after griddepcontrol.wait one parallel round of loads (counter + sums) and the stores remain.
__global__ void __launch_bounds__(DCT_PUB_MAX_VALUES * DCT_MAX_PEERS) exchange_publish_kernel(const PeerPub pub) {
    const int t = threadIdx.x, j = t / DCT_MAX_PEERS, p = t % DCT_MAX_PEERS;
    const bool on = j < pub.n && p < pub.world;
    unsigned long long* mb = on ? pub.mailbox_table[p] : nullptr;   // immutable after PeerExchange.__init__
    pdl_wait();   // the previous launch of the stream has completed and its sums are visible
    // the counter is read AFTER the wait: the previous launch may itself be a publication (two in a row, or the deferred
    // plus the final one), and its `*pub.seq = q` is only guaranteed visible once that grid has completed
    const unsigned long long q = __ldcg(pub.seq) + 1ull;
    if (on) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(__ldcg(pub.src + j));
        const unsigned long long tag = (q & 0xffffffffull) << 32;
        const size_t row = ((size_t)(q % (unsigned long long)pub.nslots) * pub.world + pub.rank) * DCT_PUB_ROW_WORDS;
        volatile unsigned long long* vd = mb + row + 2 * j;   // st.volatile: relaxed, system scope; self-validating words
        vd[0] = tag | (bits & 0xffffffffull);
        vd[1] = tag | (bits >> 32);
    }
    __syncthreads();
    if (t == 0) *pub.seq = q;
}
int check_pub(const dct_peer_pub* d) {
    if (d == nullptr || d->src == nullptr || d->seq == nullptr) return DCT_ERR_BAD_ARG;
    if (d->n < 1 || d->n > DCT_PUB_MAX_VALUES || d->world < 1 || d->world > DCT_MAX_PEERS || d->rank < 0 ||
        d->rank >= d->world || d->nslots < 1)
        return DCT_ERR_BAD_ARG;
    if (d->mailbox_table == nullptr) return DCT_ERR_BAD_ARG;
    if (!aligned(d->src, 8) || !aligned(d->seq, 8) || !aligned(d->mailbox_table, 8)) return DCT_ERR_MISALIGNED;
    return DCT_OK;
}
}  // namespace dct

extern "C" int dct_exchange_publish(const dct_peer_pub* desc, void* stream) {
    int rc = check_pub(desc);
    if (rc != DCT_OK) return rc;
    PeerPub pub;
    std::memcpy(&pub, desc, sizeof(pub));
    cudaError_t e = launch_pdl(exchange_publish_kernel, dim3(1), dim3(DCT_PUB_MAX_VALUES * DCT_MAX_PEERS), 0,
                               static_cast<cudaStream_t>(stream), pub);
    if (e != cudaSuccess) { g_last_cuda_error = e; return DCT_ERR_CUDA; }
    return check_launch();
}
