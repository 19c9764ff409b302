// tcgen05 (5th-gen tensor core) GEMM for the weight-gradient reductions:
//     D[M,N] = sum_k A[k,m] * B[g(k),n]          A: [K, lda] fp32 (m contiguous), B: [*, ldb] fp32 (n contiguous)
// i.e. dW = dG^T [x|h]_prev over all packed frames (reference: autograd of torch.nn.LSTM/LSTMCell/Linear weights,
// ABCD-VAE/modules/model.py:53,295,327-329).  Both operands are "MN-major" for the tensor core (the reduction index
// is the row index in memory), which tcgen05.mma kind::tf32 reads directly from shared memory.
//
// fp32 fidelity: 3xTF32 split.  Every operand value x is split into hi = x with the low 13 mantissa bits cleared
// (exactly representable in TF32) and lo = x - hi (exact in fp32, |lo| <= 2^-10 |x|); the accumulator in tensor
// memory receives hi*hi + hi*lo + lo*hi in fp32.  The dropped lo*lo term is <= 2^-20 relative per product.  The
// tensor core's fp32 accumulation truncates, so the error grows with the rows accumulated per CTA (~1.5e-8 per
// row, relative to the largest output); split-K keeps that to a few hundred rows.
//
// Structure per CTA (256 threads, one 128-row M tile, all N <= 256 columns, one K split), 32-row K slabs:
//   all threads   cp.async (16-byte pieces, optional row gather on B) the raw fp32 rows of slab s+2 into a 2-slot
//                 staging ring, wait for slab s and for the previous slab's MMAs (mbarrier armed by tcgen05.commit), split the
//                 staged fp32 rows into hi and lo tiles in the canonical MN-major layout for 32-bit operands
//                 (cute::UMMA::Layout_MN_SW128_32B_Atom: atoms of 4 k-rows x 128 B, 32-byte chunks XOR row index)
//   fence.proxy.async + CTA barrier; the freed staging slot is refilled two slabs ahead
//   one thread    12 x tcgen05.mma.cta_group::1.kind::tf32 (4 k-groups x 3 split products), tcgen05.commit
//   epilogue      tcgen05.ld 32x32b (warp w owns TMEM lanes 32(w%4)..+31, warps 0-3 / 4-7 take alternate column
//                 blocks) -> fp32 partials [split][M][N]; the common split-K reduce kernel applies alpha / gate
//                 map / accumulate.
// Operands whose B rows are not 16-byte aligned (the encoder's user-supplied [sumT, 65] input) take the same path
// with 4-byte cp.async pieces.
#include "sgemm.cuh"
#include "umma.cuh"

namespace {

constexpr int UM_THREADS = 256;
constexpr int UM_BM = 128;
constexpr int UM_SLAB = 32;            // K rows per slab = 4 MMA k-groups of 8
constexpr int UM_CHUNK_BYTES = 4096;   // one [32 k][32 mn] fp32 chunk
constexpr int UM_MAX_N = 256;
constexpr int UM_STAGES = 2;

struct UmmaArgs {
    const float* A; long lda;
    const float* B; long ldb; const int* b_kgather;
    int M, N, K;
    int n_pad;         // N rounded up to 16
    int k_per_split;   // multiple of UM_SLAB
    int bulk_b;        // B rows are 16-byte aligned and padded to 4 columns: bulk copies; else plain loads
    float* partials;   // [splits][M][N]
};

using umma::smem_u32; using umma::make_desc; using umma::make_idesc; using umma::mma_tf32; using umma::mbar_init;
using umma::mbar_wait; using umma::split_store;
__device__ __forceinline__ void umma_commit(uint64_t* bar) { umma::commit(bar); }
__device__ __forceinline__ uint32_t tile_off(int k, int mn4) { return umma::tile_off<UM_CHUNK_BYTES>(k, mn4); }

template <int NCH /* 32-column chunks of B: 1..8 */>
__global__ void __launch_bounds__(UM_THREADS, 1) umma_tn_kernel(UmmaArgs g) {
    extern __shared__ __align__(1024) char smem_raw[];
    constexpr int NB = NCH * 32;                               // staged B columns per row
    constexpr int A_TILE = 4 * UM_CHUNK_BYTES, B_TILE = NCH * UM_CHUNK_BYTES;
    constexpr int STAGE_BYTES = UM_SLAB * (UM_BM + NB) * 4;    // raw fp32 rows: [32][128] of A then [32][NB] of B
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    char* a_hi = base;
    char* a_lo = a_hi + A_TILE;
    char* b_hi = a_lo + A_TILE;
    char* b_lo = b_hi + B_TILE;
    char* stage = b_lo + B_TILE;                               // UM_STAGES x STAGE_BYTES
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ int brow_sm[4][UM_SLAB];   // source row of B per reduction index (gather resolved one slab ahead), -1: none

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * UM_BM;
    const int kbeg = blockIdx.z * g.k_per_split;
    const int kend = min(g.K, kbeg + g.k_per_split);
    const int nslab = kend > kbeg ? (kend - kbeg + UM_SLAB - 1) / UM_SLAB : 0;
    constexpr int TMEM_COLS = NCH <= 1 ? 32 : (NCH <= 2 ? 64 : (NCH <= 4 ? 128 : 256));

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sh)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = tmem_base_sh;
    const uint32_t idesc = make_idesc(UM_BM, g.n_pad);

    // Fill staging slot `slot` with slab s: every thread issues its share of 16-byte (4-byte for unaligned B rows)
    // asynchronous copies L2 -> shared memory and commits them as one cp.async group (empty groups keep the
    // group count uniform).  Missing rows (past the split, or gathered index -1) are skipped; the split pass
    // zero-fills them.
    auto issue_slab = [&](int s, int slot) {
        if (s < nslab) {
            const int k0s = kbeg + s * UM_SLAB;
            char* st = stage + (size_t)slot * STAGE_BYTES;
            for (int idx = tid; idx < UM_SLAB * (UM_BM / 4); idx += UM_THREADS) {
                const int k = idx >> 5, mn4 = idx & 31;
                if (k0s + k < kend && m0 + mn4 * 4 < g.M)   // a ragged last M tile: pieces past M are zero-filled by the split pass
                    cp_async16(st + k * (UM_BM * 4) + mn4 * 16, g.A + (long)(k0s + k) * g.lda + m0 + mn4 * 4);
            }
            char* stb = st + UM_SLAB * UM_BM * 4;
            if (g.bulk_b) {
                const int n4 = (g.N + 3) >> 2;   // 16-byte pieces per row (rows are padded to 4 columns)
                for (int idx = tid; idx < UM_SLAB * (NB / 4); idx += UM_THREADS) {
                    const int k = idx / (NB / 4), mn4 = idx % (NB / 4);
                    if (mn4 < n4) {
                        const long brow = brow_sm[s & 3][k];
                        if (brow >= 0) cp_async16(stb + k * (NB * 4) + mn4 * 16, g.B + brow * g.ldb + mn4 * 4);
                    }
                }
            } else {
                for (int idx = tid; idx < UM_SLAB * NB; idx += UM_THREADS) {
                    const int k = idx / NB, n = idx % NB;
                    if (n < g.N) {
                        const long brow = brow_sm[s & 3][k];
                        if (brow >= 0) {
                            unsigned sa = smem_u32(stb + k * (NB * 4) + n * 4);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(g.B + brow * g.ldb + n) : "memory");
                        }
                    }
                }
            }
        }
        cp_async_commit();
    };

    // gather indices are resolved a slab before the copies that need them, so no copy waits on an index load
    auto load_rows = [&](int s) {
        if (tid < UM_SLAB) {
            const int gk = kbeg + s * UM_SLAB + tid;
            int r = -1;
            if (s < nslab && gk < kend) r = g.b_kgather ? __ldg(g.b_kgather + gk) : gk;
            brow_sm[s & 3][tid] = r;
        }
    };
    load_rows(0); load_rows(1); load_rows(2);
    __syncthreads();
    for (int s = 0; s < UM_STAGES; ++s) issue_slab(s, s);

    for (int s = 0; s < nslab; ++s) {
        const int slot = s % UM_STAGES;
        const int k0 = kbeg + s * UM_SLAB;
        cp_async_wait<UM_STAGES - 1>();                                     // this thread's copies of slab s landed
        __syncthreads();                                                    // ... and everyone else's
        load_rows(s + 3);
        if (s > 0) mbar_wait(&mma_bar, (uint32_t)((s - 1) & 1));            // tiles are free again
        const char* st = stage + (size_t)slot * STAGE_BYTES;
        // ---- split staged fp32 rows into swizzled hi / lo tiles ----
        for (int idx = tid; idx < UM_SLAB * (UM_BM / 4); idx += UM_THREADS) {
            const int k = idx >> 5, mn4 = idx & 31;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + k < kend && m0 + mn4 * 4 < g.M) v = *reinterpret_cast<const float4*>(st + k * (UM_BM * 4) + mn4 * 16);
            split_store(a_hi, a_lo, tile_off(k, mn4), v);
        }
        for (int idx = tid; idx < UM_SLAB * (NB / 4); idx += UM_THREADS) {
            const int k = idx / (NB / 4), mn4 = idx % (NB / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mn4 * 4 < g.N) {
                const bool present = brow_sm[s & 3][k] >= 0;
                if (present) {
                    v = *reinterpret_cast<const float4*>(st + UM_SLAB * UM_BM * 4 + k * (NB * 4) + mn4 * 16);
                    const int n = mn4 * 4;   // columns past N hold stale staging bytes: clear them
                    if (n + 1 >= g.N) v.y = 0.f;
                    if (n + 2 >= g.N) v.z = 0.f;
                    if (n + 3 >= g.N) v.w = 0.f;
                }
            }
            split_store(b_hi, b_lo, tile_off(k, mn4), v);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy accesses ordered before async-proxy ones
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int kg = 0; kg < UM_SLAB / 8; ++kg) {
                // one MMA consumes K = 8 rows = two 4-row atoms (SBO = 512 B) of every 32-column chunk (LBO = 4 KB)
                const uint64_t dah = make_desc(ah + kg * 1024, UM_CHUNK_BYTES, 512);
                const uint64_t dal = make_desc(al + kg * 1024, UM_CHUNK_BYTES, 512);
                const uint64_t dbh = make_desc(bh + kg * 1024, UM_CHUNK_BYTES, 512);
                const uint64_t dbl = make_desc(bl + kg * 1024, UM_CHUNK_BYTES, 512);
                mma_tf32(tmem_d, dah, dbh, idesc, (s > 0 || kg > 0) ? 1u : 0u);
                mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                mma_tf32(tmem_d, dal, dbh, idesc, 1u);
            }
            umma_commit(&mma_bar);
        }
        // every thread is past its reads of staging[slot]: refill it two slabs ahead
        issue_slab(s + UM_STAGES, slot);
    }
    // ---- epilogue: TMEM -> registers -> per-warp transpose in shared memory -> partials ----
    // A thread reads ROW lane of the accumulator; storing that directly scatters every store instruction over 32 rows.
    // Each warp bounces its 32 x 32 block through the (now free) staging area and writes contiguous 128-byte runs.
    cp_async_wait<0>();
    if (nslab > 0) mbar_wait(&mma_bar, (uint32_t)((nslab - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int lane_q = warp & 3;                       // TMEM lane quarter this warp may access
    float* tile = reinterpret_cast<float*>(stage) + warp * (32 * 33);
    float* out = g.partials + ((long)blockIdx.z * g.M + m0 + lane_q * 32) * g.N;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 32; c0 < g.n_pad; c0 += 64) {
        uint32_t ra[16], rb[16];
        if (nslab > 0) {
            const uint32_t taddr = tmem_d + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)c0;
            umma::tmem_ld16(taddr, ra);
            umma::tmem_ld16(taddr + 16, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ra[j] = rb[j] = 0u;
        }
        __syncwarp();                                  // the previous block has been read back
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            tile[lane * 33 + j] = __uint_as_float(ra[j]);
            tile[lane * 33 + 16 + j] = __uint_as_float(rb[j]);
        }
        __syncwarp();
        if (c0 + lane < g.N) {
            const int rows = min(32, g.M - (m0 + lane_q * 32));
            for (int r = 0; r < rows; ++r) out[(long)r * g.N + c0 + lane] = tile[r * 33 + lane];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

template <int NCH>
size_t umma_smem_bytes() {
    return (size_t)(2 * 4 + 2 * NCH) * UM_CHUNK_BYTES + (size_t)UM_STAGES * UM_SLAB * (UM_BM + NCH * 32) * 4 + 1024;
}

template <int NCH>
int launch_nch(const UmmaArgs& a, int splits, cudaStream_t stream) {
    const size_t smem = umma_smem_bytes<NCH>();
    static bool configured = false;   // per instantiation; setting the attribute is idempotent
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(umma_tn_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((a.M + UM_BM - 1) / UM_BM, 1, splits);
    umma_tn_kernel<NCH><<<grid, UM_THREADS, smem, stream>>>(a);
    KERNEL_CHECK();
    return ABCD_OK;
}

}  // namespace

// M: whole 128-row tiles, or a ragged last tile (rows past M are computed on zero-filled operands and dropped) when the
// padding stays below ~50%
bool umma_tn_shape_ok(int M, int N, int K) {
    if (M % UM_BM != 0 && (M < 64 || M % UM_BM < 64)) return false;
    if (N < 1 || N > UM_MAX_N) return false;
    return K >= 256;                                            // below that not worth a tensor-memory round trip
}

bool umma_tn_eligible(const GemmArgs& g) {
    if (g.a_kmajor || g.b_kmajor) return false;                 // TN form only
    if (!umma_tn_shape_ok(g.M, g.N, g.K)) return false;
    if (g.lda % 4 != 0 || (reinterpret_cast<uintptr_t>(g.A) & 15) != 0) return false;
    if (((g.M + 3) & ~3) > g.lda) return false;                 // the last 16-byte piece of a ragged M tile stays inside the row
    if (!g.partials) return false;
    return true;
}

// Split count that gives every SM one CTA; the caller's partials buffer must hold splits*M*N floats.
int umma_tn_pick_splits(int M, int K) {
    int mtiles = (M + UM_BM - 1) / UM_BM;
    int s = (148 + mtiles - 1) / mtiles;
    if (s * mtiles > 148) s = 148 / mtiles;                      // one wave
    int max_s = (K + 4 * UM_SLAB - 1) / (4 * UM_SLAB);          // at least 4 slabs per split
    if (s > max_s) s = max_s;
    return s < 1 ? 1 : s;
}

int umma_tn_launch(const GemmArgs& g, int splits, cudaStream_t stream) {
    UmmaArgs a;
    a.A = g.A; a.lda = g.lda; a.B = g.B; a.ldb = g.ldb; a.b_kgather = g.b_kgather;
    a.M = g.M; a.N = g.N; a.K = g.K;
    a.n_pad = (g.N + 15) & ~15;
    int per = (g.K + splits - 1) / splits;
    a.k_per_split = (per + UM_SLAB - 1) / UM_SLAB * UM_SLAB;
    // bulk copies move whole 16-byte pieces: rows must be aligned and padded up to a multiple of 4 columns
    a.bulk_b = (g.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (((g.N + 3) & ~3) <= g.ldb);
    a.partials = g.partials;
    const int nch = (a.n_pad + 31) / 32;
    switch (nch) {
        case 1: return launch_nch<1>(a, splits, stream);
        case 2: return launch_nch<2>(a, splits, stream);
        case 3: return launch_nch<3>(a, splits, stream);
        case 4: return launch_nch<4>(a, splits, stream);
        case 5: return launch_nch<5>(a, splits, stream);
        case 6: return launch_nch<6>(a, splits, stream);
        case 7: return launch_nch<7>(a, splits, stream);
        default: return launch_nch<8>(a, splits, stream);
    }
}
