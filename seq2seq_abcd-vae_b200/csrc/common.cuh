// Shared device/host helpers for the ABCD-VAE sm_100a kernels.
// Everything here is internal to the shared library; the public surface is include/abcd_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define ABCD_H 256          // RNN hidden size the recurrent kernels are specialised for
#define ABCD_M 256          // MLP hidden size the decoder kernels are specialised for

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local last-error string; nothing throws across the C ABI)
// ---------------------------------------------------------------------------------------------
void abcd_set_error(const char* fmt, ...);
#define ABCD_OK 0
#define ABCD_ERR_ARG (-1)
#define ABCD_ERR_CUDA (-2)
#define ABCD_ERR_UNSUPPORTED (-3)
#define ABCD_ERR_WORKSPACE (-4)

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            abcd_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ABCD_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define ARG_CHECK(cond, ...)                 \
    do {                                     \
        if (!(cond)) {                       \
            abcd_set_error(__VA_ARGS__);     \
            return ABCD_ERR_ARG;             \
        }                                    \
    } while (0)

#define TRY(expr)                \
    do {                         \
        int _r = (expr);         \
        if (_r != ABCD_OK) return _r; \
    } while (0)

// launch accounting + optional per-kernel-group CUDA-event timing (api.cu)
void abcd_note_launch(int n);
#define KERNEL_CHECK()                    \
    do {                                  \
        abcd_note_launch(1);              \
        CUDA_TRY(cudaGetLastError());     \
    } while (0)

enum ProfId {
    PROF_ENC_FWD = 0, PROF_ENC_BWD, PROF_DEC_FWD, PROF_DEC_BWD, PROF_GEMM, PROF_ROWWISE, PROF_OPTIM, PROF_STFT,
    PROF_NOISE, PROF_COUNT
};
void abcd_prof_begin(int id, cudaStream_t s);
void abcd_prof_end(int id, cudaStream_t s);
struct ProfScope {
    int id; cudaStream_t s;
    ProfScope(int id_, cudaStream_t s_) : id(id_), s(s_) { abcd_prof_begin(id, s); }
    ~ProfScope() { abcd_prof_end(id, s); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace blob (the library never allocates device memory).
struct Arena {
    char* base;
    size_t cap;
    size_t off;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// same, for code where the threads of a warp reach the wait at different points (no .aligned)
__device__ __forceinline__ void cluster_wait_acquire_divergent() {
    asm volatile("barrier.cluster.wait.acquire;\n" ::: "memory");
}
// Full cluster barrier with release/acquire semantics at cluster scope: global-memory writes made
// by any thread of any CTA of the cluster before the barrier are visible to every thread after it.
__device__ __forceinline__ void cluster_sync_all() {
    cluster_arrive_release();
    cluster_wait_acquire();
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
    return r;
}

// 16-byte async copy global -> shared, bypassing L1 (.cg): the sources are rows another CTA of the
// cluster wrote before the last cluster barrier, so they must come from L2.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// L2-coherent scalar/vector loads for data produced by other CTAs inside the same launch.
__device__ __forceinline__ float ld_cg(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_cg2(const float* p) {
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red /*>=33 floats*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ double block_sum_d(double v, double* red /*>=33 doubles*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
#endif  // __CUDACC__
