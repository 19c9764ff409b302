// Building blocks of the persistent, cluster-partitioned recurrent kernels (enc.cu, dec.cu).
//
// Execution model.  One thread-block cluster owns one "row tile" of up to ROWS=80 segments (sequence b belongs
// to tile b % n_tiles, so every tile sees the same length distribution and row activity inside a tile is a
// shrinking prefix, exactly like PackedSequence.batch_sizes; 512 segments -> 7 tiles, which is the number of
// 16-CTA clusters a B200 keeps resident at once).  Each CTA owns a slice of the 256 hidden units and keeps its
// slice of every weight matrix resident in shared memory for the whole launch.  A time step is a short sequence
// of phases; in each phase a CTA computes
//     out[rows, my columns] = A[rows, K] * Wslice[K, my columns]
// where the A rows are activations written to global memory (L2) by all CTAs of the cluster in the previous
// phase.  Those activation buffers are the tensors the backward pass needs anyway, so the exchange costs no
// extra traffic.  Phases are separated by the hardware cluster barrier (release/acquire), not a grid sync.
//
// phase_gemm(): A rows are streamed L2 -> shared memory with cp.async.cg in double-buffered K-chunks (one CTA
// barrier per chunk); each thread accumulates a 5-row x 4-column register tile from register-double-buffered
// shared-memory fragments (for gate GEMMs the 4 columns are the 4 gates of one hidden unit, so the cell update is
// thread-local).  Small-N phases split K across thread groups and reduce through shared memory.
#pragma once
#include "common.cuh"

namespace rnn {

constexpr int CL = 16;             // CTAs per cluster (decoder; the encoder kernels use 8)
constexpr int U = ABCD_H / CL;     // hidden units per CTA
constexpr int RT = 10;             // rows per thread (more FFMAs per LDS.128: an LDS.128 costs 4 smem-pipe cycles per warp)
constexpr int NRG = 8;             // row groups
constexpr int RG = NRG;            // thread of row group rg owns the interleaved rows rg + i*RG (i < RT): retiring
                                   // segments thin every row group evenly, so all warps keep equal work
constexpr int ROWS = RT * NRG;     // rows (segments) per tile
constexpr int NTHREADS = 256;

// Row tiles of a batch of B segments.  At least ceil(B / ROWS) so that a tile fits the per-thread register tile; for small
// batches (a rank's shard of the publication batch under strong scaling: 64 segments) MORE, smaller tiles -- up to the 7 clusters
// of 16 CTAs (or 2 x 7 clusters of 8) the GPU keeps co-resident -- because the per-step cost of a tile shrinks with its rows while
// an idle SM buys nothing: 64 segments run as 7 tiles of 9-10 rows instead of one tile of 64.
constexpr int MAX_CORESIDENT_TILES = 7;
constexpr int MIN_TILE_ROWS = 8;
__host__ __device__ inline int num_tiles(int B) {
    const int need = (B + ROWS - 1) / ROWS;
    int spread = (B + MIN_TILE_ROWS - 1) / MIN_TILE_ROWS;
    if (spread > MAX_CORESIDENT_TILES) spread = MAX_CORESIDENT_TILES;
    return need > spread ? need : spread;
}

struct ThreadMap {
    int cg, rg, ks;   // column group (4 columns), row group, K-split index
    bool active;
};
// Lane assignment.  A warp covers 8 column groups x 4 row groups when the phase has a multiple of 8 column groups:
// its weight-fragment load is then 128 contiguous bytes (one shared-memory wavefront, where 16 groups x 2 rows
// would need two) and its activation-fragment load is 4 rows x 16 B (one wavefront; the row pitch is 4 banks).
// Other column-group counts use the plain (cg fastest) order, which already mixes several row groups per warp.
__device__ __forceinline__ ThreadMap make_map(int tid, int ncg, int KS) {
    ThreadMap m;
    if ((ncg & 7) == 0) {
        const int lane = tid & 31, w = tid >> 5;
        const int ncg_hi = ncg >> 3;           // warps along the column groups
        m.cg = (w % ncg_hi) * 8 + (lane & 7);
        m.rg = ((w / ncg_hi) % (NRG / 4)) * 4 + (lane >> 3);
        m.ks = w / (ncg_hi * (NRG / 4));
    } else {
        m.cg = tid % ncg;
        m.rg = (tid / ncg) % NRG;
        m.ks = tid / (ncg * NRG);
    }
    m.active = m.ks < KS;
    return m;
}
__host__ __device__ inline int pick_ks(int ncg, int nthreads = NTHREADS) {
    int ks = nthreads / (ncg * NRG);
    return ks < 1 ? 1 : ks;
}

struct Stage {
    float* buf0;
    float* buf1;
};
// Two staging buffers of ROWS rows x (kc + 4) floats (the 4-float pad makes the row pitch 4 banks mod 32).
__host__ __device__ inline Stage make_stage(float* base, int kc) {
    Stage s;
    s.buf0 = base; s.buf1 = base + ROWS * (kc + 4);
    return s;
}
__host__ __device__ inline size_t stage_floats(int kc) { return (size_t)2 * ROWS * (kc + 4); }
// K-split factor of a phase whose cross-split reduction scratch ((KS-1) x ROWS x ldw floats) must fit in the staging buffers
__host__ __device__ inline int pick_ks_fit(int ncg, int ldw, int kc, int nthreads = NTHREADS) {
    int ks = pick_ks(ncg, nthreads);
    int cap = 1 + (int)(stage_floats(kc) / (size_t)(ROWS * ldw));
    return ks < cap ? ks : cap;
}

__device__ __forceinline__ void zero_acc(float (&acc)[RT][4]) {
#pragma unroll
    for (int i = 0; i < RT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
}

// A row r of the tile is the concatenation [rp0(r)[0,K0) | rp1(r)[0,K1)] (a null pointer means zeros).
// K0 and K1 are multiples of 4 and all row pointers are 16-byte aligned.  KC = staged K-chunk (32/64/128/256).
template <int NT, int KC, class RP0, class RP1>
__device__ __forceinline__ void stage_chunk(float* dst, int k_begin, int k_len, int nrows, int K0,
                                            RP0 rp0, RP1 rp1, int tid) {
    constexpr int LG = KC == 256 ? 6 : KC == 128 ? 5 : (KC == 64 ? 4 : 3);   // log2 of 16-byte pieces per staged row
    constexpr int LDA = KC + 4;
    const int total = nrows << LG;
    for (int idx = tid; idx < total; idx += NT) {
        const int r = idx >> LG, p4 = (idx & ((1 << LG) - 1)) * 4;
        if (p4 >= k_len) continue;
        const int k = k_begin + p4;
        const float* src;
        if (k < K0) { src = rp0(r); if (src) src += k; }
        else { src = rp1(r); if (src) src += (k - K0); }
        float* d = dst + r * LDA + p4;
        if (src) cp_async16(d, src);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Inner product over n k-quads for the first NS row slots of this thread, branch-free.  All shared-memory
// strides are compile-time (LDA = staged row pitch, LDW = weight-slice row pitch), so every fragment load is
// [pointer + immediate]; the fragments of quad q+1 are fetched while quad q is multiplied (register double
// buffering), so the LDS latency hides behind 16*NS FFMAs.
template <int NS>
struct Frag {
    float4 w[4];
    float4 a[NS];
};
template <int NS, int LDA, int LDW>
__device__ __forceinline__ void load_frag(Frag<NS>& f, const float* a_rg, const float* w) {
    f.w[0] = *reinterpret_cast<const float4*>(w);
    f.w[1] = *reinterpret_cast<const float4*>(w + LDW);
    f.w[2] = *reinterpret_cast<const float4*>(w + 2 * LDW);
    f.w[3] = *reinterpret_cast<const float4*>(w + 3 * LDW);
#pragma unroll
    for (int i = 0; i < NS; ++i) f.a[i] = *reinterpret_cast<const float4*>(a_rg + i * RG * LDA);
}
template <int NS>
__device__ __forceinline__ void fma_frag(float (&acc)[RT][4], const Frag<NS>& f) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float4 a = f.a[i];
        acc[i][0] = fmaf(a.x, f.w[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, f.w[0].y, acc[i][1]);
        acc[i][2] = fmaf(a.x, f.w[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, f.w[0].w, acc[i][3]);
        acc[i][0] = fmaf(a.y, f.w[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, f.w[1].y, acc[i][1]);
        acc[i][2] = fmaf(a.y, f.w[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, f.w[1].w, acc[i][3]);
        acc[i][0] = fmaf(a.z, f.w[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, f.w[2].y, acc[i][1]);
        acc[i][2] = fmaf(a.z, f.w[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, f.w[2].w, acc[i][3]);
        acc[i][0] = fmaf(a.w, f.w[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, f.w[3].y, acc[i][1]);
        acc[i][2] = fmaf(a.w, f.w[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, f.w[3].w, acc[i][3]);
    }
}
template <int NS, int LDA, int LDW>
__device__ __forceinline__ void mma_range(float (&acc)[RT][4], const float* a_rg /* &As[rg][k4_begin*4] */,
                                          const float* w /* &Wsm[k_begin + k4_begin*4][cg*4] */, int n) {
    if (n <= 0) return;
    Frag<NS> f0, f1;
    load_frag<NS, LDA, LDW>(f0, a_rg, w);
    int it = 0;
    for (; it + 2 <= n; it += 2) {
        load_frag<NS, LDA, LDW>(f1, a_rg + 4, w + 4 * LDW);
        fma_frag<NS>(acc, f0);
        a_rg += 8; w += 8 * LDW;
        if (it + 2 < n) load_frag<NS, LDA, LDW>(f0, a_rg, w);   // only when another quad follows
        fma_frag<NS>(acc, f1);
    }
    if (it < n) fma_frag<NS>(acc, f0);
}

// One K-chunk.  K-splits own contiguous k ranges of the chunk.
template <int KC, int LDW>
__device__ __forceinline__ void mma_chunk(float (&acc)[RT][4], const float* As, int k_len,
                                          const float* Wk /* &Wsm[k_begin][cg*4] */, const ThreadMap& tm,
                                          int KS, int nslots) {
    constexpr int LDA = KC + 4;
    const int nk4 = k_len >> 2;
    const int per = (nk4 + KS - 1) / KS;
    const int b = tm.ks * per;
    const int n = min(per, nk4 - b);
    const float* a_rg = As + tm.rg * LDA + b * 4;
    const float* w = Wk + (b * 4) * LDW;
    switch (nslots) {
        case 1: mma_range<1, LDA, LDW>(acc, a_rg, w, n); break;
        case 2: mma_range<2, LDA, LDW>(acc, a_rg, w, n); break;
        case 3: mma_range<3, LDA, LDW>(acc, a_rg, w, n); break;
        case 4: mma_range<4, LDA, LDW>(acc, a_rg, w, n); break;
        case 5: mma_range<5, LDA, LDW>(acc, a_rg, w, n); break;
        case 6: mma_range<6, LDA, LDW>(acc, a_rg, w, n); break;
        case 7: mma_range<7, LDA, LDW>(acc, a_rg, w, n); break;
        case 8: mma_range<8, LDA, LDW>(acc, a_rg, w, n); break;
        case 9: mma_range<9, LDA, LDW>(acc, a_rg, w, n); break;
        default: mma_range<RT, LDA, LDW>(acc, a_rg, w, n); break;
    }
}

// acc += A[rows, K0+K1] * Wsm[K0+K1, LDW] for this thread's tile.  On return the threads with ks == 0
// hold the full sums for rows rg + i*RG (i < RT) and columns cg*4 .. cg*4+3.  Ends with a CTA barrier, so the
// staging buffers may be reused immediately.  `wait_chunk` >= 1: the caller has issued a cluster arrive; the matching
// wait is executed just before chunk `wait_chunk` is staged.  `dbg` (optional) accumulates clock cycles of the calling thread:
// [0] staging issue, [1] wait + barrier, [2] multiply, [3] chunks, [4] tail barrier.
template <int NT, int KC, int LDW, class RP0, class RP1>
__device__ __forceinline__ void phase_gemm(float (&acc)[RT][4], const Stage& st, int K0, RP0 rp0, int K1,
                                           RP1 rp1, int nrows, const float* Wsm, const ThreadMap& tm,
                                           int KS, int tid, unsigned long long* dbg = nullptr, int wait_chunk = -1) {
    const int K = K0 + K1;
    const int nslots = (nrows + RG - 1) / RG;   // active 16-row slots (uniform over the CTA)
    const int nch = (K + KC - 1) / KC;
    const float* Wcol = Wsm + tm.cg * 4;
    long long d0 = dbg ? clock64() : 0;
    stage_chunk<NT, KC>(st.buf0, 0, min(KC, K), nrows, K0, rp0, rp1, tid);
    cp_async_commit();
    if (dbg) { long long d1 = clock64(); dbg[0] += d1 - d0; }
    for (int c = 0; c < nch; ++c) {
        long long e0 = dbg ? clock64() : 0;
        cp_async_wait<0>();
        __syncthreads();   // chunk c visible to all; everyone is done with chunk c-1's buffer
        long long e1 = dbg ? clock64() : 0;
        if (c + 1 < nch) {
            // split-phase cluster barrier: rows of chunks >= wait_chunk are produced by other CTAs in the phase that ended
            // with a cluster arrive; everything before them overlaps the barrier latency
            if (c + 1 == wait_chunk) cluster_wait_acquire();
            const int kb = (c + 1) * KC;
            stage_chunk<NT, KC>((c & 1) ? st.buf0 : st.buf1, kb, min(KC, K - kb), nrows, K0, rp0, rp1, tid);
            cp_async_commit();
        }
        long long e2 = dbg ? clock64() : 0;
        if (tm.active) {
            const int kb = c * KC;
            mma_chunk<KC, LDW>(acc, (c & 1) ? st.buf1 : st.buf0, min(KC, K - kb), Wcol + kb * LDW, tm, KS, nslots);
        }
        if (dbg) { long long e3 = clock64(); dbg[1] += e1 - e0; dbg[0] += e2 - e1; dbg[2] += e3 - e2; dbg[3] += 1; }
    }
    if (wait_chunk >= nch) cluster_wait_acquire();   // (never with the callers' shapes; keeps arrive/wait paired)
    long long f0 = dbg ? clock64() : 0;
    __syncthreads();
    if (dbg) dbg[4] += clock64() - f0;
    if (KS > 1) {
        // cross-split reduction through the (now idle) staging buffers
        float* scratch = st.buf0;
        if (tm.active && tm.ks > 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
                if (i < nslots)
                    *reinterpret_cast<float4*>(scratch + ((tm.ks - 1) * ROWS + tm.rg + i * RG) * LDW + tm.cg * 4) =
                        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (tm.active && tm.ks == 0) {
            for (int s = 1; s < KS; ++s) {
#pragma unroll
                for (int i = 0; i < RT; ++i)
                    if (i < nslots) {
                        float4 v = *reinterpret_cast<const float4*>(scratch + ((s - 1) * ROWS + tm.rg + i * RG) * LDW + tm.cg * 4);
                        acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
                    }
            }
        }
        __syncthreads();
    }
}

// out[rows, 4 ncg] = A[rows, 4*K4] * W[4*K4, LDW] on operands that are already in shared memory; work items are
// (column group, row group) pairs with the usual 10-row x 4-column register tile (rows rg + 8 i) and the
// register-double-buffered inner loop of rnn_common.cuh.  Results go to `out` (global, row pitch ldo).
template <int K4, int LDA, int LDW>
__device__ __forceinline__ void local_gemm_store(const float* A, const float* W, int ncg, int na, float* out, long ldo, int tid) {
    const int nslots = (na + RG - 1) / RG;
    for (int item = tid; item < ncg * NRG; item += NTHREADS) {
        const int cg = item % ncg, rg = item / ncg;
        if (rg >= na) continue;
        const float* a = A + rg * LDA;
        const float* w = W + cg * 4;
        float acc[RT][4];
        zero_acc(acc);
        switch (nslots) {
            case 1: mma_range<1, LDA, LDW>(acc, a, w, K4); break;
            case 2: mma_range<2, LDA, LDW>(acc, a, w, K4); break;
            case 3: mma_range<3, LDA, LDW>(acc, a, w, K4); break;
            case 4: mma_range<4, LDA, LDW>(acc, a, w, K4); break;
            case 5: mma_range<5, LDA, LDW>(acc, a, w, K4); break;
            case 6: mma_range<6, LDA, LDW>(acc, a, w, K4); break;
            case 7: mma_range<7, LDA, LDW>(acc, a, w, K4); break;
            case 8: mma_range<8, LDA, LDW>(acc, a, w, K4); break;
            case 9: mma_range<9, LDA, LDW>(acc, a, w, K4); break;
            default: mma_range<RT, LDA, LDW>(acc, a, w, K4); break;
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = rg + i * RG;
            if (r < na) *reinterpret_cast<float4*>(out + (long)r * ldo + cg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

// Partial-sum reductions over the CL = 16 CTAs of a cluster read 16 values per output from L2.  To keep every load of a
// step in flight at once (one L2 round trip instead of one per item), FOUR lanes share an output: lane `sub` (= tid & 3)
// loads and adds sources 4 sub .. 4 sub + 3 in order, then quad_sum() combines the four sub-sums with two butterfly
// exchanges -- ((s0..s3) + (s4..s7)) + ((s8..s11) + (s12..s15)), the same bits in all four lanes: a fixed order, so results
// stay reproducible run to run.  Must be executed by all 32 lanes of a warp.
__device__ __forceinline__ float quad_sum(float a) {
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    return a;
}
__device__ __forceinline__ float4 quad_sum4(float4 a) {
    a.x = quad_sum(a.x); a.y = quad_sum(a.y); a.z = quad_sum(a.z); a.w = quad_sum(a.w);
    return a;
}
__device__ __forceinline__ float pick4(const float4& a, int i) { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }

struct NullRow {
    __device__ __forceinline__ const float* operator()(int) const { return nullptr; }
};

// number of active rows of tile `tile` when `bs` sequences are active (sequence b = tile + n_tiles*r)
__device__ __forceinline__ int active_rows(int bs, int tile, int n_tiles) {
    int n = bs - tile;
    return n <= 0 ? 0 : (n + n_tiles - 1) / n_tiles;
}

}  // namespace rnn
