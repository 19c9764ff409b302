// Probe for the TS form of tcgen05.mma (A operand in tensor memory): D[128,N] = A[128,K] * B[K,N], kind::tf32,
// single pass (raw fp32 operands: the tensor core ignores the low 13 mantissa bits).  One CTA, 128 threads.
//   A -> tensor memory with tcgen05.st.32x32b (lane = row m, one 32-bit column per k)
//   B -> shared memory, MN-major SWIZZLE_128B_BASE32B tiles (the layout every other kernel here uses)
// Used by tools/ts_probe.py to establish the layout facts the next step of the recurrent kernels relies on
// (weights resident in tensor memory as the A operand, activations streamed as B).
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int PB_CHUNK = 4096;   // [32 k][32 n]

__global__ void __launch_bounds__(128, 1) ts_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D, int K, int N, int variant) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* b_tile = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t d_col = 0, a_col = 256;
    // A rows -> tensor memory: this thread's lane is row m = 32*warp + lane, 16 columns (k values) per store
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t r[16];
        // variant bit 2: lanes 32q + i (i < 16) hold row 16q + i (the 16-lanes-per-quadrant pattern); other lanes zero
        const int arow = (variant & 4) ? ((lane < 16) ? warp * 16 + lane : -1) : m;
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((arow >= 0 && k0 + j < K) ? A[(long)arow * K + k0 + j] : 0.f);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_col + (uint32_t)k0;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
            ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
              "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    // B -> MN-major tile (32-k slabs stacked: slab s at s * (N/32 chunks) * 4 KB)
    const int nch = (N + 31) / 32;
    for (int idx = tid; idx < K * (nch * 8); idx += 128) {
        const int k = idx / (nch * 8), n4 = idx % (nch * 8);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n4 * 4 < N) v = *reinterpret_cast<const float4*>(B + (long)k * N + n4 * 4);
        const int slab = k >> 5, kl = k & 31;
        *reinterpret_cast<float4*>(b_tile + (size_t)slab * nch * PB_CHUNK + umma::tile_off<PB_CHUNK>(kl, n4)) = v;
    }
    // variant bit 3: SS form -- A as MN-major tiles [k][m] in shared memory behind the B tiles (Mi rows)
    char* a_tile = b_tile + (size_t)((K + 31) / 32) * nch * PB_CHUNK;
    {
        const int Mi = (variant & 2) ? 64 : 128;
        for (int idx = tid; idx < K * Mi; idx += 128) {
            const int mm = idx % Mi, k = idx / Mi;
            *reinterpret_cast<float*>(a_tile + (size_t)(k >> 5) * (Mi / 32) * PB_CHUNK + umma::tile_off<PB_CHUNK>(k & 31, mm >> 2) + (mm & 3) * 4) = A[(long)mm * K + k];
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        umma::fence_after_sync();
        // instruction descriptor: c F32, a/b TF32, a_major per variant bit 0 (0: K-major, 1: MN-major), b MN-major
        const int Mi = (variant & 2) ? 64 : 128;   // variant bit 1: M = 64 (A rows 0..63 in lanes 0..63)
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(Mi >> 4) << 24);
        if (variant & 1) idesc |= (1u << 15);
        for (int kg = 0; kg < K / 8; ++kg) {
            const int slab = kg >> 2, kq = kg & 3;
            const uint64_t db = umma::make_desc(umma::smem_u32(b_tile + (size_t)slab * nch * PB_CHUNK) + kq * 1024, PB_CHUNK, 512);
            const uint32_t ta = tmem + a_col + (uint32_t)(kg * 8);
            const uint32_t acc = kg > 0 ? 1u : 0u;
            if (variant & 8) {
                const uint64_t da = umma::make_desc(umma::smem_u32(a_tile + (size_t)slab * (Mi / 32) * PB_CHUNK) + kq * 1024, PB_CHUNK, 512);
                umma::mma_tf32(tmem + d_col, da, db, idesc | (1u << 15), acc);
            } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                ::"r"(tmem + d_col), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + d_col + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c0 + j < N) D[(long)m * N + c0 + j] = __uint_as_float(r[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

}  // namespace

int umma_ts_probe(const float* A, const float* B, float* D, int K, int N, int variant, cudaStream_t stream) {
    ARG_CHECK(K >= 8 && K <= 256 && K % 8 == 0 && N >= 16 && N <= 256 && N % 16 == 0, "ts probe: K %d / N %d unsupported", K, N);
    const size_t smem = (size_t)((K + 31) / 32) * (((N + 31) / 32) + 4) * PB_CHUNK + 1024;
    CUDA_TRY(cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ts_probe_kernel<<<1, 128, smem, stream>>>(A, B, D, K, N, variant);
    KERNEL_CHECK();
    return ABCD_OK;
}
