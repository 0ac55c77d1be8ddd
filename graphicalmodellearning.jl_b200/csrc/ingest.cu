// Ingest of the reference's own input type: `learn` receives the histogram as a K x (N+1) matrix of Real -- Int64 from
// `sample` (src/sampling.jl:54), Float64 from `readdlm` (test/runtests.jl:71) -- column-major, column 1 the counts,
// column 1+i spin i (src/GraphicalModelLearning.jl:76-81).  At the headline size that matrix is 80 GB while its content is
// 10 GB of +-1 bytes, so the narrowing runs on the HOST, next to the data: a pool of host threads converts column blocks
// into pinned int8 staging buffers (validating that every entry is exactly -1 or +1) and streams them to their final
// rows of the device histogram; only 1 byte per spin crosses the host link.  The counts column goes over as float64.
// What used to be `Int8.(samples[:, 2:end])` in the Julia shim (a single-threaded 80 GB -> 10 GB copy plus a second pass
// over the result) is now part of the measured call.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace gml {
namespace {

constexpr int64_t INGEST_CHUNK = 1 << 21;     // elements per work unit (2 Mi spins -> 2 MiB of int8)

template <class T> inline bool narrow_chunk(const T* __restrict__ src, int8_t* __restrict__ dst, int64_t n) {
    unsigned bad = 0;
    for (int64_t k = 0; k < n; ++k) {
        const T v = src[k];
        bad |= (unsigned)!(v == (T)1 || v == (T)-1);
        dst[k] = (int8_t)v;
    }
    return bad == 0;
}

inline bool narrow_any(const void* col, int dtype, int64_t k0, int8_t* dst, int64_t n) {
    switch (dtype) {
        case GML_B200_DTYPE_F64: return narrow_chunk(static_cast<const double*>(col) + k0, dst, n);
        case GML_B200_DTYPE_I64: return narrow_chunk(static_cast<const int64_t*>(col) + k0, dst, n);
        case GML_B200_DTYPE_F32: return narrow_chunk(static_cast<const float*>(col) + k0, dst, n);
        case GML_B200_DTYPE_I32: return narrow_chunk(static_cast<const int32_t*>(col) + k0, dst, n);
        default: return narrow_chunk(static_cast<const int8_t*>(col) + k0, dst, n);
    }
}

inline size_t dtype_size(int dtype) {
    switch (dtype) {
        case GML_B200_DTYPE_F64: case GML_B200_DTYPE_I64: return 8;
        case GML_B200_DTYPE_F32: case GML_B200_DTYPE_I32: return 4;
        default: return 1;
    }
}

inline double count_at(const void* col, int dtype, int64_t k) {
    switch (dtype) {
        case GML_B200_DTYPE_F64: return static_cast<const double*>(col)[k];
        case GML_B200_DTYPE_I64: return (double)static_cast<const int64_t*>(col)[k];
        case GML_B200_DTYPE_F32: return (double)static_cast<const float*>(col)[k];
        case GML_B200_DTYPE_I32: return (double)static_cast<const int32_t*>(col)[k];
        default: return (double)static_cast<const int8_t*>(col)[k];
    }
}

}  // namespace

// samples: column-major K_total x (N+1) matrix with leading dimension ld (elements); rows [k0, k0+K) are ingested.
// d_base: int8 [>= N rows x Kp] device rows (pitch Kp), d_counts: double [K] on the device.  n_threads <= 0: all host cores.
void ingest_matrix(const void* samples, int dtype, int64_t ld, int64_t k0, int64_t K, int32_t N, int8_t* d_base, int64_t Kp,
                   double* d_counts, int n_threads, double* out_host_ms) {
    GML_REQUIRE(dtype >= GML_B200_DTYPE_F64 && dtype <= GML_B200_DTYPE_I8, "unknown element type of the samples matrix");
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t esz = dtype_size(dtype);
    const char* base = static_cast<const char*>(samples);
    const int64_t chunks_per_col = ceil_div(K, INGEST_CHUNK);
    const int64_t n_units = chunks_per_col * N;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(T, 64), n_units));
    int dev = 0;
    GML_CUDA(cudaGetDevice(&dev));

    // counts column: float64 through one pinned buffer
    double* h_counts = nullptr;
    GML_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_counts), sizeof(double) * (size_t)K));
    const char* col0 = base + (size_t)k0 * esz;
    for (int64_t k = 0; k < K; ++k) h_counts[k] = count_at(col0, dtype, k);
    cudaStream_t st0;
    GML_CUDA(cudaStreamCreateWithFlags(&st0, cudaStreamNonBlocking));
    GML_CUDA(cudaMemcpyAsync(d_counts, h_counts, sizeof(double) * (size_t)K, cudaMemcpyHostToDevice, st0));

    std::atomic<int64_t> next{0};
    std::atomic<int> bad{0}, cuda_err{0};
    auto worker = [&]() {
        if (cudaSetDevice(dev) != cudaSuccess) { cuda_err = 1; return; }
        int8_t* buf[2] = {nullptr, nullptr};
        cudaEvent_t ev[2];
        cudaStream_t st;
        bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int b = 0; b < 2 && ok; ++b)
            ok = cudaMallocHost(reinterpret_cast<void**>(&buf[b]), (size_t)INGEST_CHUNK) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) == cudaSuccess;
        int b = 0;
        bool used[2] = {false, false};
        while (ok) {
            const int64_t u = next.fetch_add(1);
            if (u >= n_units || bad.load(std::memory_order_relaxed)) break;
            const int64_t i = u / chunks_per_col, c = u % chunks_per_col;
            const int64_t a = c * INGEST_CHUNK, n = std::min<int64_t>(INGEST_CHUNK, K - a);
            if (used[b] && cudaEventSynchronize(ev[b]) != cudaSuccess) { ok = false; break; }
            const char* col = base + ((size_t)(1 + i) * (size_t)ld) * esz;
            if (!narrow_any(col, dtype, k0 + a, buf[b], n)) { bad = 1; break; }
            ok = cudaMemcpyAsync(d_base + (size_t)i * (size_t)Kp + a, buf[b], (size_t)n, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                 cudaEventRecord(ev[b], st) == cudaSuccess;
            used[b] = true;
            b ^= 1;
        }
        if (!ok) cuda_err = 1;
        cudaStreamSynchronize(st);
        for (int q = 0; q < 2; ++q) { if (buf[q]) { cudaFreeHost(buf[q]); cudaEventDestroy(ev[q]); } }
        cudaStreamDestroy(st);
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    cudaStreamSynchronize(st0);
    cudaStreamDestroy(st0);
    cudaFreeHost(h_counts);
    if (out_host_ms) *out_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    GML_REQUIRE(bad.load() == 0, "histogram spins must be exactly -1 or +1");
    if (cuda_err.load()) {
        set_error(std::string("matrix ingest: CUDA error: ") + cudaGetErrorString(cudaGetLastError()));
        throw CudaError{GML_B200_ECUDA};
    }
}

}  // namespace gml
