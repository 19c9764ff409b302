// Tensor-core (tcgen05, 3xTF32) version of rnn::phase_gemm for the persistent recurrent kernels.
//
// Same contract as phase_gemm(): on return the thread (cg, rg) holds, in acc[i][0..3], the sums for rows
// rg + i*RG and columns cg*4..cg*4+3 of  A[rows, K0+K1] * Wsm[K0+K1, LDW],  so the fused cell updates are untouched.
//
// Why: the FFMA loop is bound by the shared-memory pipe (an LDS.128 costs 4 pipe cycles per warp; a 10x4 thread
// tile needs 14 of them per 160 FFMAs).  The tensor core reads its operands from shared memory itself, so per
// operand element the SM spends one LDS + two STS (the hi/lo split) per STEP instead of one LDS per 10 FFMAs.
//
// Pipeline per phase, 16-row K slabs (two MMA k-groups):
//   cp.async     raw fp32 A rows of slab s+3 -> 3-slot ring (L2 -> smem, 16-byte pieces)
//   all threads  split slab s of A (raw ring) and of W (resident fp32 slice) into hi / lo operand tiles, MN-major
//                SWIZZLE_128B_BASE32B layout (umma.cuh), stage s&1; a 4x4 register transpose turns the row-major A
//                rows into the MN-major 16-byte pieces (4 consecutive rows at one k)
//   one thread   6 x tcgen05.mma kind::tf32 (2 k-groups x {hi*hi, hi*lo, lo*hi}) into the TMEM accumulator
//                [128 lanes = tile rows][NCOLS columns], tcgen05.commit -> the stage's mbarrier
//   epilogue     tcgen05.ld (8 warps: lane quarter = warp&3, column half = warp>>2) -> shared-memory bounce
//                [row][NCOLS+4] -> each thread reads its acc tile.
// Rows >= nrows of the A tiles are never written: they hold stale (finite or not) data and only feed accumulator
// rows nobody reads.  M is always 128 (the tensor core's time per MMA does not drop below M=128).
#pragma once
#include "rnn_common.cuh"
#include "umma.cuh"

namespace rnn {

constexpr int TC_SLAB = 16;                    // K rows per slab
constexpr int TC_CHUNK = TC_SLAB * 128;        // bytes of one [16 k][32 mn] operand chunk
constexpr int TC_A_CHUNKS = (ROWS + 31) / 32;  // 96 >= ROWS tile rows are addressable without aliasing live data
constexpr int TC_RAW_PITCH = TC_SLAB + 4;      // floats; 80-byte rows: conflict-free LDS.128 down a column of rows
constexpr int TC_RAW_RING = 3;

template <int NCOLS>
struct TcShape {
    static_assert(NCOLS % 16 == 0 && NCOLS >= 16 && NCOLS <= 256, "MMA N must be a multiple of 16 (M = 128)");
    static constexpr int B_CHUNKS = (NCOLS + 31) / 32;
    static constexpr int A_BYTES = TC_A_CHUNKS * TC_CHUNK;
    static constexpr int B_BYTES = B_CHUNKS * TC_CHUNK;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);   // [A_hi][A_lo][B_hi][B_lo]; the M=128 A descriptors run one chunk into the next tile
    static constexpr int TILE_BYTES = 2 * STAGE_BYTES;
    static constexpr int BOUNCE_PITCH = NCOLS + 4;
    static_assert(ROWS * BOUNCE_PITCH * 4 <= TILE_BYTES, "accumulator bounce must fit in the operand tiles");
    static constexpr int TMEM_COLS = NCOLS <= 32 ? 32 : (NCOLS <= 64 ? 64 : (NCOLS <= 128 ? 128 : 256));
};
constexpr size_t tc_raw_bytes() { return (size_t)TC_RAW_RING * ROWS * TC_RAW_PITCH * sizeof(float); }

struct TcCtx {
    char* tiles;          // 1024-byte aligned; sized for the largest phase of the kernel
    float* raw;           // TC_RAW_RING x ROWS x TC_RAW_PITCH
    uint64_t* bars;       // [0], [1]: operand stage consumed; [2]: phase done
    uint32_t tmem;        // accumulator base (lane 0, first column)
    uint32_t uses[3];     // commits issued so far on each barrier (uniform over the CTA)
};

// One warp allocates TMEM_COLS accumulator columns, one thread initialises the barriers.  Ends with a CTA barrier.
template <int TMEM_COLS>
__device__ __forceinline__ void tc_init(TcCtx& c, char* tiles, float* raw, uint64_t* bars, uint32_t* tmem_slot, int tid) {
    c.tiles = tiles; c.raw = raw; c.bars = bars;
    c.uses[0] = c.uses[1] = c.uses[2] = 0;
    if ((tid >> 5) == 0) umma::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = *tmem_slot;
}
template <int TMEM_COLS>
__device__ __forceinline__ void tc_fini(TcCtx& c, int tid) {
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<TMEM_COLS>(c.tmem);
}

// Raw staging: every thread owns up to RAW_ROUNDS (row, 16-byte piece) slots of a slab; the source row pointers
// are slab-invariant and resolved once per phase.
template <int NT>
struct TcRawSrc {
    static constexpr int ROUNDS = (ROWS * (TC_SLAB / 4) + NT - 1) / NT;
    const float* p0[ROUNDS];
    const float* p1[ROUNDS];
    int dst[ROUNDS];      // float offset inside a ring slot, -1: no slot
    int piece[ROUNDS];
};
template <int NT, class RP0, class RP1>
__device__ __forceinline__ void tc_raw_setup(TcRawSrc<NT>& rs, int nrows, RP0 rp0, RP1 rp1, int tid) {
#pragma unroll
    for (int j = 0; j < TcRawSrc<NT>::ROUNDS; ++j) {
        const int idx = tid + j * NT;
        const int r = idx >> 2, pc = idx & 3;
        rs.dst[j] = -1; rs.p0[j] = nullptr; rs.p1[j] = nullptr; rs.piece[j] = pc;
        if (r < nrows) {
            rs.dst[j] = r * TC_RAW_PITCH + ((pc ^ ((r >> 3) & 3)) << 2);   // piece swizzle: conflict-free row-quad reads
            rs.p0[j] = rp0(r); rs.p1[j] = rp1(r);
        }
    }
}
template <int NT>
__device__ __forceinline__ void tc_stage_raw(float* slot, const TcRawSrc<NT>& rs, int k_begin, int K, int K0) {
#pragma unroll
    for (int j = 0; j < TcRawSrc<NT>::ROUNDS; ++j) {
        if (rs.dst[j] < 0) continue;
        const int k = k_begin + rs.piece[j] * 4;
        const float* src = nullptr;
        if (k < K) {
            if (k < K0) { if (rs.p0[j]) src = rs.p0[j] + k; }
            else { if (rs.p1[j]) src = rs.p1[j] + (k - K0); }
        }
        float* d = slot + rs.dst[j];
        if (src) cp_async16(d, src);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// NT threads; NCOLS = MMA N (LDW rounded up to 16); LDW = pitch (floats) and column count of the resident slice.
// Roles: warp 0 issues the MMAs of slab s-1 while warps 1..7 split slab s (threads 32..111: one A item = 4 rows x
// 4 k each; threads 112..255: the W rows), so the issue latency hides behind the split.  All per-thread
// addressing is slab-invariant and computed once per phase.
template <int NT, int NCOLS, int LDW, class RP0, class RP1>
__device__ __forceinline__ void phase_gemm_tc(float (&acc)[RT][4], TcCtx& c, int K0, RP0 rp0, int K1, RP1 rp1,
                                              int nrows, const float* Wsm, const ThreadMap& tm, int tid,
                                              unsigned long long* dbg = nullptr) {
    using S = TcShape<NCOLS>;
    static_assert(NT == 256, "role and epilogue mapping assume 8 warps");
    static_assert(LDW % 4 == 0 && LDW <= NCOLS, "resident slice pitch");
    constexpr int NQ_MAX = ROWS / 4;                       // row quads of a full tile
    constexpr int A_THREADS = NQ_MAX * (TC_SLAB / 4);      // 80
    constexpr int W_THREADS = NT - 32 - A_THREADS;         // 144
    constexpr int W4 = NCOLS / 4;
    constexpr int NW = TC_SLAB * W4;                       // W pieces per slab
    constexpr int W_ROUNDS = (NW + W_THREADS - 1) / W_THREADS;
    const int K = K0 + K1;
    const int nsl = (K + TC_SLAB - 1) / TC_SLAB;
    const int nq = (nrows + 3) >> 2;                       // active row quads
    const int warp = tid >> 5;
    const int st = tid - 32;                               // index among the split threads

    // ---- slab-invariant addressing ----
    int a_raw = 0; uint32_t a_off[4] = {0, 0, 0, 0}; bool a_on = false;
    int w_src[W_ROUNDS]; uint32_t w_off[W_ROUNDS]; int w_kl[W_ROUNDS];
    if (st >= 0 && st < A_THREADS) {
        const int q = st % NQ_MAX, kq = st / NQ_MAX;
        a_on = q < nq;
        a_raw = (q * 4) * TC_RAW_PITCH + ((kq ^ ((q >> 1) & 3)) << 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) a_off[j] = umma::tile_off<TC_CHUNK>(kq * 4 + j, q);
    }
#pragma unroll
    for (int j = 0; j < W_ROUNDS; ++j) {
        const int idx = (st - A_THREADS) + j * W_THREADS;
        w_kl[j] = -1; w_src[j] = 0; w_off[j] = 0;
        if (st >= A_THREADS && idx < NW) {
            const int kl = idx / W4, mn4 = idx % W4;
            w_kl[j] = kl;
            w_src[j] = (mn4 * 4 < LDW) ? kl * LDW + mn4 * 4 : -1;
            w_off[j] = umma::tile_off<TC_CHUNK>(kl, mn4);
        }
    }
    const uint32_t idesc = umma::make_idesc(128, NCOLS);
    const uint64_t d_ah = umma::make_desc(umma::smem_u32(c.tiles), TC_CHUNK, 512);
    const uint64_t d_al = d_ah + (uint64_t)(S::A_BYTES >> 4);
    const uint64_t d_bh = d_al + (uint64_t)(S::A_BYTES >> 4);
    const uint64_t d_bl = d_bh + (uint64_t)(S::B_BYTES >> 4);

    long long t0 = dbg ? clock64() : 0;
    TcRawSrc<NT> rs;
    tc_raw_setup<NT>(rs, nrows, rp0, rp1, tid);
    for (int s = 0; s < TC_RAW_RING; ++s) {
        if (s < nsl) tc_stage_raw<NT>(c.raw + s * (ROWS * TC_RAW_PITCH), rs, s * TC_SLAB, K, K0);
        cp_async_commit();
    }
    cp_async_wait<TC_RAW_RING - 1>();
    __syncthreads();                                       // slab 0 visible to all
    if (dbg) { long long t1 = clock64(); dbg[0] += t1 - t0; t0 = t1; }

    for (int sl = 0; sl <= nsl; ++sl) {
        if (warp == 0) {
            // ---- issue the MMAs of slab sl-1 (its tiles were completed at the previous barrier) ----
            if (sl > 0 && tid == 0) {
                umma::fence_after_sync();
                const int ps = sl - 1, stg = ps & 1;
                const uint64_t so = (uint64_t)((stg * S::STAGE_BYTES) >> 4);
#pragma unroll
                for (int kg = 0; kg < TC_SLAB / 8; ++kg) {
                    const uint64_t o = so + (uint64_t)((kg * 1024) >> 4);
                    umma::mma_tf32(c.tmem, d_ah + o, d_bh + o, idesc, (ps > 0 || kg > 0) ? 1u : 0u);
                    umma::mma_tf32(c.tmem, d_ah + o, d_bl + o, idesc, 1u);
                    umma::mma_tf32(c.tmem, d_al + o, d_bh + o, idesc, 1u);
                }
                umma::commit(&c.bars[stg]);
                if (ps == nsl - 1) umma::commit(&c.bars[2]);
                if (dbg) { long long t1 = clock64(); dbg[1] += t1 - t0; t0 = t1; }
            }
            if (sl < nsl) cp_async_wait<TC_RAW_RING - 2>();
            if (dbg) { long long t1 = clock64(); dbg[2] += t1 - t0; t0 = t1; }
        } else if (sl < nsl) {
            const int stg = sl & 1;
            char* a_hi = c.tiles + stg * S::STAGE_BYTES;
            char* a_lo = a_hi + S::A_BYTES;
            char* b_hi = a_lo + S::A_BYTES;
            char* b_lo = b_hi + S::B_BYTES;
            const uint32_t used = c.uses[stg];
            if (used > 0) umma::mbar_wait(&c.bars[stg], (used - 1) & 1);   // MMAs of slab sl-2 consumed this stage
            if (dbg) { long long t1 = clock64(); dbg[1] += t1 - t0; t0 = t1; }
            if (a_on) {
                // A: rows 4q..4q+3, k-quad kq -> four MN-major pieces (one per k)
                const float* rb = c.raw + (sl % TC_RAW_RING) * (ROWS * TC_RAW_PITCH) + a_raw;
                const float4 r0 = *reinterpret_cast<const float4*>(rb);
                const float4 r1 = *reinterpret_cast<const float4*>(rb + TC_RAW_PITCH);
                const float4 r2 = *reinterpret_cast<const float4*>(rb + 2 * TC_RAW_PITCH);
                const float4 r3 = *reinterpret_cast<const float4*>(rb + 3 * TC_RAW_PITCH);
                umma::split_store(a_hi, a_lo, a_off[0], make_float4(r0.x, r1.x, r2.x, r3.x));
                umma::split_store(a_hi, a_lo, a_off[1], make_float4(r0.y, r1.y, r2.y, r3.y));
                umma::split_store(a_hi, a_lo, a_off[2], make_float4(r0.z, r1.z, r2.z, r3.z));
                umma::split_store(a_hi, a_lo, a_off[3], make_float4(r0.w, r1.w, r2.w, r3.w));
            }
            {
                const float* wb = Wsm + sl * (TC_SLAB * LDW);
                float4 v[W_ROUNDS];
#pragma unroll
                for (int j = 0; j < W_ROUNDS; ++j) {
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (w_kl[j] >= 0 && w_src[j] >= 0 && sl * TC_SLAB + w_kl[j] < K) v[j] = *reinterpret_cast<const float4*>(wb + w_src[j]);
                }
#pragma unroll
                for (int j = 0; j < W_ROUNDS; ++j)
                    if (w_kl[j] >= 0) umma::split_store(b_hi, b_lo, w_off[j], v[j]);
            }
            if (dbg) { long long t1 = clock64(); dbg[2] += t1 - t0; t0 = t1; }
            umma::fence_proxy_async();
            cp_async_wait<TC_RAW_RING - 2>();              // this thread's pieces of slab sl+1 have landed
        }
        if (sl < nsl) c.uses[sl & 1] += 1;                 // the commit warp 0 issues in the next iteration
        umma::fence_before_sync();
        __syncthreads();                                   // tiles complete; raw slab sl+1 visible; raw slab sl free
        if (dbg) { long long t1 = clock64(); dbg[3] += t1 - t0; t0 = t1; }
        if (sl < nsl) {
            if (sl + TC_RAW_RING < nsl)
                tc_stage_raw<NT>(c.raw + (sl % TC_RAW_RING) * (ROWS * TC_RAW_PITCH), rs, (sl + TC_RAW_RING) * TC_SLAB, K, K0);
            cp_async_commit();
        }
        if (dbg) { long long t1 = clock64(); dbg[4] += t1 - t0; t0 = t1; }
    }
    cp_async_wait<0>();
    c.uses[2] += 1;
    umma::mbar_wait(&c.bars[2], (c.uses[2] - 1) & 1);      // every MMA of the phase has completed
    umma::fence_after_sync();

    // accumulator -> bounce (reuses the operand tiles)
    float* bounce = reinterpret_cast<float*>(c.tiles);
    {
        const int warp = tid >> 5, lane = tid & 31;
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane;
        if (q * 32 < nrows) {
#pragma unroll 1
            for (int c0 = half * (NCOLS / 2); c0 < (half + 1) * (NCOLS / 2); c0 += 16) {
                uint32_t r[16];
                umma::tmem_ld16(c.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                if (row < nrows) {
                    float* o = bounce + row * S::BOUNCE_PITCH + c0;
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<float4*>(o + 4 * v) = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                                          __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tm.active && tm.ks == 0 && tm.cg * 4 < LDW) {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = tm.rg + i * RG;
            if (r < nrows) {
                const float4 v = *reinterpret_cast<const float4*>(bounce + r * S::BOUNCE_PITCH + tm.cg * 4);
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
    }
    __syncthreads();   // bounce (= operand tiles) free for the next phase
}

}  // namespace rnn
