// tcgen05 (3xTF32) GEMM for the tall "activation x weight" products outside the time loops:
//     C[M,N] = act(alpha * A[M,K] * op(B) + bias[n])      A row-major (k contiguous), M large (all packed frames)
//     op(B) = B^T with B [N,K] row-major  (NT: y = x W^T -- hoisted input projection, offset-predictor hidden layer;
//                                          reference: torch.nn.LSTM weight_ih / MLP Linear, model.py:53,327-329)
//          or B with B [K,N] row-major    (NN: dx = dy W  -- data gradients)
// Same tensor-core machinery as umma_gemm.cu (MN-major SWIZZLE_128B_BASE32B operand tiles, hi/lo split, fp32
// accumulation in tensor memory).  Operands whose rows run along k are turned into MN-major tiles by a 4x4 register
// transpose during the split.  One CTA = one 128-row x (<=256)-column output tile, the whole K range (K is small
// here: 68..1024), 32-row K slabs, epilogue straight from tensor memory to C (bias / scale / tanh fused).
#include "sgemm.cuh"
#include "umma.cuh"

namespace {

constexpr int RM_THREADS = 256;
constexpr int RM_BM = 128;
constexpr int RM_BN = 256;
constexpr int RM_SLAB = 32;
constexpr int RM_CHUNK = 4096;                     // [32 k][32 mn] fp32
constexpr int RM_PITCH = RM_SLAB + 4;              // raw k-contiguous rows: 36 floats
constexpr int RM_A_TILE = 4 * RM_CHUNK, RM_B_TILE = 8 * RM_CHUNK;
constexpr int RM_RAW_A = RM_BM * RM_PITCH * 4;     // bytes
constexpr int RM_RAW_B = RM_BN * RM_PITCH * 4;     // bytes (k-major B; the n-major form needs 32*256*4, less)
constexpr int RM_STAGE = RM_RAW_A + RM_RAW_B;
constexpr int RM_STAGES = 2;
constexpr int RM_EP_PITCH = 36;                     // epilogue bounce tile: [32 rows][36] floats per warp (8 x 4608 B in the staging area)
constexpr size_t RM_SMEM = 2 * RM_A_TILE + 2 * RM_B_TILE + RM_STAGES * RM_STAGE + 1024;

struct RmArgs {
    const float* A; long lda;
    const float* B; long ldb;
    float* C; long ldc;
    const float* bias;
    float alpha;
    int act;
    int M, N, K;
};

using umma::smem_u32;

// raw piece p (4 floats) of row r lives at piece position p ^ ((r >> 3) & 3): reading 4 consecutive rows x one piece
// for 32 row quads then touches every bank group equally (same scheme as rnn_ts.cuh)
__device__ __forceinline__ int raw_off(int r, int p) { return r * RM_PITCH + ((p ^ ((r >> 3) & 3)) << 2); }

// rows 4q..4q+3, k-quad kq of a k-contiguous raw block -> four MN-major 16-byte pieces
__device__ __forceinline__ void split_transposed(const float* raw, int q, int kq, bool k_valid, char* hi, char* lo) {
    float4 r0, r1, r2, r3;
    if (k_valid) {
        r0 = *reinterpret_cast<const float4*>(raw + raw_off(4 * q + 0, kq));
        r1 = *reinterpret_cast<const float4*>(raw + raw_off(4 * q + 1, kq));
        r2 = *reinterpret_cast<const float4*>(raw + raw_off(4 * q + 2, kq));
        r3 = *reinterpret_cast<const float4*>(raw + raw_off(4 * q + 3, kq));
    } else {
        r0 = r1 = r2 = r3 = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int k = kq * 4;
    umma::split_store(hi, lo, umma::tile_off<RM_CHUNK>(k + 0, q), make_float4(r0.x, r1.x, r2.x, r3.x));
    umma::split_store(hi, lo, umma::tile_off<RM_CHUNK>(k + 1, q), make_float4(r0.y, r1.y, r2.y, r3.y));
    umma::split_store(hi, lo, umma::tile_off<RM_CHUNK>(k + 2, q), make_float4(r0.z, r1.z, r2.z, r3.z));
    umma::split_store(hi, lo, umma::tile_off<RM_CHUNK>(k + 3, q), make_float4(r0.w, r1.w, r2.w, r3.w));
}

template <bool B_KM>
__global__ void __launch_bounds__(RM_THREADS, 1) umma_rm_kernel(RmArgs g) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    char* a_hi = base;
    char* a_lo = a_hi + RM_A_TILE;
    char* b_hi = a_lo + RM_A_TILE;
    char* b_lo = b_hi + RM_B_TILE;
    char* stage = b_lo + RM_B_TILE;
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * RM_BM, n0 = blockIdx.y * RM_BN;
    const int nloc = min(RM_BN, g.N - n0);
    const int n_pad = (nloc + 15) & ~15;
    const int nslab = (g.K + RM_SLAB - 1) / RM_SLAB;

    if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = umma::make_idesc(RM_BM, n_pad);

    auto issue_slab = [&](int s, int slot) {
        if (s < nslab) {
            const int k0 = s * RM_SLAB;
            float* ra = reinterpret_cast<float*>(stage + (size_t)slot * RM_STAGE);
            float* rb = reinterpret_cast<float*>(stage + (size_t)slot * RM_STAGE + RM_RAW_A);
            for (int idx = tid; idx < RM_BM * 8; idx += RM_THREADS) {
                const int r = idx >> 3, p = idx & 7;
                if (m0 + r < g.M && k0 + p * 4 < g.K) cp_async16(ra + raw_off(r, p), g.A + (long)(m0 + r) * g.lda + k0 + p * 4);
            }
            if (B_KM) {
                for (int idx = tid; idx < RM_BN * 8; idx += RM_THREADS) {
                    const int r = idx >> 3, p = idx & 7;
                    if (r < nloc && k0 + p * 4 < g.K) cp_async16(rb + raw_off(r, p), g.B + (long)(n0 + r) * g.ldb + k0 + p * 4);
                }
            } else {
                for (int idx = tid; idx < RM_SLAB * (RM_BN / 4); idx += RM_THREADS) {
                    const int k = idx >> 6, n4 = idx & 63;
                    if (k0 + k < g.K && n4 * 4 < nloc) cp_async16(rb + k * RM_BN + n4 * 4, g.B + (long)(k0 + k) * g.ldb + n0 + n4 * 4);
                }
            }
        }
        cp_async_commit();
    };
    for (int s = 0; s < RM_STAGES; ++s) issue_slab(s, s);

    for (int s = 0; s < nslab; ++s) {
        const int slot = s % RM_STAGES;
        const int k0 = s * RM_SLAB;
        cp_async_wait<RM_STAGES - 1>();
        __syncthreads();
        if (s > 0) umma::mbar_wait(&mma_bar, (uint32_t)((s - 1) & 1));   // operand tiles are free again
        const float* ra = reinterpret_cast<const float*>(stage + (size_t)slot * RM_STAGE);
        const float* rb = reinterpret_cast<const float*>(stage + (size_t)slot * RM_STAGE + RM_RAW_A);
        {   // A: 32 row quads x 8 k quads = one item per thread
            const int q = tid & 31, kq = tid >> 5;
            split_transposed(ra, q, kq, k0 + kq * 4 < g.K, a_hi, a_lo);
        }
        if (B_KM) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * RM_THREADS;
                const int q = idx & 63, kq = idx >> 6;
                if (q * 4 < n_pad) split_transposed(rb, q, kq, k0 + kq * 4 < g.K, b_hi, b_lo);
            }
        } else {
            for (int idx = tid; idx < RM_SLAB * (RM_BN / 4); idx += RM_THREADS) {
                const int k = idx >> 6, n4 = idx & 63;
                if (n4 * 4 >= n_pad) continue;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + k < g.K && n4 * 4 < nloc) v = *reinterpret_cast<const float4*>(rb + k * RM_BN + n4 * 4);
                umma::split_store(b_hi, b_lo, umma::tile_off<RM_CHUNK>(k, n4), v);
            }
        }
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int kg = 0; kg < RM_SLAB / 8; ++kg) {
                const uint64_t dah = umma::make_desc(ah + kg * 1024, RM_CHUNK, 512);
                const uint64_t dal = umma::make_desc(al + kg * 1024, RM_CHUNK, 512);
                const uint64_t dbh = umma::make_desc(bh + kg * 1024, RM_CHUNK, 512);
                const uint64_t dbl = umma::make_desc(bl + kg * 1024, RM_CHUNK, 512);
                umma::mma_tf32(tmem_d, dah, dbh, idesc, (s > 0 || kg > 0) ? 1u : 0u);
                umma::mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                umma::mma_tf32(tmem_d, dal, dbh, idesc, 1u);
            }
            umma::commit(&mma_bar);
        }
        issue_slab(s + RM_STAGES, slot);
    }
    // ---- epilogue: tensor memory -> registers -> per-warp transpose in shared memory -> C ----
    // A thread reads ROW lane of the accumulator; storing that directly would scatter 16-byte pieces over 32 rows per
    // instruction.  Each warp bounces its 32 x 32 block through the (now free) staging area and writes 128-byte runs.
    cp_async_wait<0>();
    if (nslab > 0) umma::mbar_wait(&mma_bar, (uint32_t)((nslab - 1) & 1));
    umma::fence_after_sync();
    const int lane_q = warp & 3;
    float* tile = reinterpret_cast<float*>(stage) + warp * (32 * RM_EP_PITCH);
    const int rq = lane >> 3, cq = lane & 7;          // read-back: rows rq, rq + 4, ...; column quad cq
#pragma unroll 1
    for (int c0 = (warp >> 2) * 32; c0 < n_pad; c0 += 64) {
        uint32_t ra[16], rb[16];
        if (nslab > 0) {
            const uint32_t taddr = tmem_d + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)c0;
            umma::tmem_ld16(taddr, ra);
            umma::tmem_ld16(taddr + 16, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ra[j] = rb[j] = 0u;
        }
        __syncwarp();                                 // the previous block has been read back
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<float4*>(tile + lane * RM_EP_PITCH + 4 * q) =
                make_float4(__uint_as_float(ra[4 * q]), __uint_as_float(ra[4 * q + 1]), __uint_as_float(ra[4 * q + 2]), __uint_as_float(ra[4 * q + 3]));
            *reinterpret_cast<float4*>(tile + lane * RM_EP_PITCH + 16 + 4 * q) =
                make_float4(__uint_as_float(rb[4 * q]), __uint_as_float(rb[4 * q + 1]), __uint_as_float(rb[4 * q + 2]), __uint_as_float(rb[4 * q + 3]));
        }
        __syncwarp();
        const int c = c0 + 4 * cq;
        if (c < nloc) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias) b = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = rq + 4 * i;
                const int m = m0 + lane_q * 32 + row;
                if (m < g.M) {
                    float4 v = *reinterpret_cast<const float4*>(tile + row * RM_EP_PITCH + 4 * cq);
                    v.x = v.x * g.alpha + b.x; v.y = v.y * g.alpha + b.y; v.z = v.z * g.alpha + b.z; v.w = v.w * g.alpha + b.w;
                    if (g.act == 1) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
                    *reinterpret_cast<float4*>(g.C + (long)m * g.ldc + n0 + c) = v;
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<256>(tmem_d);
}

}  // namespace

bool umma_rm_eligible(const GemmArgs& g) {
    if (!g.a_kmajor || g.splits != 1 || g.accumulate || g.bias_on_m) return false;
    if (g.b_kgather || g.b_nsrc.enabled || g.m_dst.enabled) return false;
    if (g.M < 2048 || g.N < 64 || g.K < 32) return false;                 // tall activation matrices only
    if (g.K % 4 != 0 || g.N % 4 != 0) return false;
    auto aligned = [](const void* p, long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; };
    if (!aligned(g.A, g.lda) || !aligned(g.B, g.ldb) || !aligned(g.C, g.ldc)) return false;
    if (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15) != 0) return false;
    return true;
}

int umma_rm_launch(const GemmArgs& g, cudaStream_t stream) {
    RmArgs a;
    a.A = g.A; a.lda = g.lda; a.B = g.B; a.ldb = g.ldb; a.C = g.C; a.ldc = g.ldc;
    a.bias = g.bias; a.alpha = g.alpha; a.act = g.act; a.M = g.M; a.N = g.N; a.K = g.K;
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(umma_rm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(umma_rm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
        configured = true;
    }
    dim3 grid((g.M + RM_BM - 1) / RM_BM, (g.N + RM_BN - 1) / RM_BN);
    if (g.b_kmajor) umma_rm_kernel<true><<<grid, RM_THREADS, RM_SMEM, stream>>>(a);
    else umma_rm_kernel<false><<<grid, RM_THREADS, RM_SMEM, stream>>>(a);
    KERNEL_CHECK();
    return ABCD_OK;
}
