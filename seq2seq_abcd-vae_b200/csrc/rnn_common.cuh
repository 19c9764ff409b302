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
constexpr int RT = 5;              // rows per thread
constexpr int RG = 16;             // row groups; thread rg owns rows rg + i*RG
constexpr int ROWS = RT * RG;      // rows (segments) per tile
constexpr int NTHREADS = 256;

struct ThreadMap {
    int cg, rg, ks;   // column group (4 columns), row group, K-split index
    bool active;
};
// Threads of a warp share as few distinct A rows (2) and as many contiguous W columns (16 groups) as possible.
__device__ __forceinline__ ThreadMap make_map(int tid, int ncg, int KS) {
    ThreadMap m;
    if (ncg <= 16) {
        m.cg = tid % ncg;
        m.rg = (tid / ncg) % RG;
        m.ks = tid / (ncg * RG);
    } else {   // ncg == 32 (8-CTA encoder clusters, 512 threads): column groups split in two halves of 16
        m.cg = (tid & 15) + 16 * (tid / (16 * RG));
        m.rg = (tid >> 4) % RG;
        m.ks = 0;
    }
    m.active = m.ks < KS;
    return m;
}
__host__ __device__ inline int pick_ks(int ncg, int nthreads = NTHREADS) {
    int ks = nthreads / (ncg * RG);
    return ks < 1 ? 1 : ks;
}

struct Stage {
    float* buf0;
    float* buf1;
    int lda;     // kc + 4 floats
    int kc;      // K-chunk: 32, 64 or 128
    int lg_ppr;  // log2(kc / 4): 16-byte pieces per staged row
};
__host__ __device__ inline Stage make_stage(float* base, int kc) {
    Stage s;
    s.kc = kc; s.lda = kc + 4;
    s.lg_ppr = kc == 256 ? 6 : (kc == 128 ? 5 : (kc == 64 ? 4 : 3));
    s.buf0 = base; s.buf1 = base + ROWS * (kc + 4);
    return s;
}
// As rows retire, the same two buffers hold fewer rows of longer K-chunks (fewer chunk boundaries per phase):
// the chunk doubles while `rows` staged rows of it still fit in one buffer sized for ROWS rows of st.kc.
__device__ __forceinline__ Stage grow_stage(const Stage& st, int rows, int K) {
    Stage s = st;
    const int cap = ROWS * (st.kc + 4);
    while (s.kc < 256 && s.kc < K && rows * (2 * s.kc + 4) <= cap) {
        s.kc *= 2; s.lg_ppr += 1;
    }
    s.lda = s.kc + 4;
    return s;
}
__host__ __device__ inline size_t stage_floats(int kc) { return (size_t)2 * ROWS * (kc + 4); }

__device__ __forceinline__ void zero_acc(float (&acc)[RT][4]) {
#pragma unroll
    for (int i = 0; i < RT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
}

// A row r of the tile is the concatenation [rp0(r)[0,K0) | rp1(r)[0,K1)] (a null pointer means zeros).
// K0 and K1 are multiples of 4 and all row pointers are 16-byte aligned.
template <int NT, class RP0, class RP1>
__device__ __forceinline__ void stage_chunk(float* dst, const Stage& st, int k_begin, int k_len, int nrows, int K0,
                                            RP0 rp0, RP1 rp1, int tid) {
    const int ppr_mask = (1 << st.lg_ppr) - 1;
    const int total = nrows << st.lg_ppr;
    for (int idx = tid; idx < total; idx += NT) {
        const int r = idx >> st.lg_ppr, p4 = (idx & ppr_mask) * 4;
        if (p4 >= k_len) continue;
        const int k = k_begin + p4;
        const float* src;
        if (k < K0) { src = rp0(r); if (src) src += k; }
        else { src = rp1(r); if (src) src += (k - K0); }
        float* d = dst + r * st.lda + p4;
        if (src) cp_async16(d, src);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Inner product over k4 in [k4_begin, k4_end) for the first NS row slots of this thread, branch-free: the
// shared-memory fragments of step k4+1 are fetched while step k4 is multiplied (register double buffering), so
// the LDS latency hides behind 16*NS FFMAs.
template <int NS>
struct Frag {
    float4 w[4];
    float4 a[NS];
};
template <int NS>
__device__ __forceinline__ void load_frag(Frag<NS>& f, const float* a_rg, int lda, const float* w, int ldw) {
    f.w[0] = *reinterpret_cast<const float4*>(w);
    f.w[1] = *reinterpret_cast<const float4*>(w + ldw);
    f.w[2] = *reinterpret_cast<const float4*>(w + 2 * ldw);
    f.w[3] = *reinterpret_cast<const float4*>(w + 3 * ldw);
#pragma unroll
    for (int i = 0; i < NS; ++i) f.a[i] = *reinterpret_cast<const float4*>(a_rg + i * RG * lda);
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
template <int NS>
__device__ __forceinline__ void mma_range(float (&acc)[RT][4], const float* a_rg /* &As[rg][k4_begin*4] */, int lda,
                                          const float* w /* &Wsm[k_begin + k4_begin*4][cg*4] */, int ldw, int n) {
    if (n <= 0) return;
    const int wstep = 4 * ldw;
    Frag<NS> f0, f1;
    load_frag<NS>(f0, a_rg, lda, w, ldw);
    int it = 0;
    for (; it + 2 <= n; it += 2) {
        load_frag<NS>(f1, a_rg + 4, lda, w + wstep, ldw);
        fma_frag<NS>(acc, f0);
        a_rg += 8; w += 2 * wstep;
        // the prefetch past the end of the range stays inside the staging row / weight slice padding rules:
        // it is only issued when another step follows
        if (it + 2 < n) load_frag<NS>(f0, a_rg, lda, w, ldw);
        fma_frag<NS>(acc, f1);
    }
    if (it < n) fma_frag<NS>(acc, f0);
}

// One K-chunk.  K-splits own contiguous k ranges of the chunk.
__device__ __forceinline__ void mma_chunk(float (&acc)[RT][4], const float* As, int lda, int k_len,
                                          const float* Wk /* &Wsm[k_begin][cg*4] */, int ldw, const ThreadMap& tm,
                                          int KS, int nslots) {
    const int nk4 = k_len >> 2;
    const int per = (nk4 + KS - 1) / KS;
    const int b = tm.ks * per;
    const int n = min(per, nk4 - b);
    const float* a_rg = As + tm.rg * lda + b * 4;
    const float* w = Wk + (long)(b * 4) * ldw;
    switch (nslots) {
        case 1: mma_range<1>(acc, a_rg, lda, w, ldw, n); break;
        case 2: mma_range<2>(acc, a_rg, lda, w, ldw, n); break;
        case 3: mma_range<3>(acc, a_rg, lda, w, ldw, n); break;
        case 4: mma_range<4>(acc, a_rg, lda, w, ldw, n); break;
        default: mma_range<RT>(acc, a_rg, lda, w, ldw, n); break;
    }
}

// acc += A[rows, K0+K1] * Wsm[K0+K1, ncols] for this thread's tile.  On return the threads with ks == 0
// hold the full sums for rows rg + i*RG (i < RT) and columns cg*4 .. cg*4+3.  Ends with a CTA barrier, so the
// staging buffers may be reused immediately.
template <int NT, class RP0, class RP1>
__device__ __forceinline__ void phase_gemm(float (&acc)[RT][4], const Stage& st_full, int K0, RP0 rp0, int K1,
                                           RP1 rp1, int nrows, const float* Wsm, int ldw, const ThreadMap& tm,
                                           int KS, int tid) {
    const int K = K0 + K1;
    const int nslots = (nrows + RG - 1) / RG;
    const Stage st = grow_stage(st_full, nslots * RG, K);   // the multiply reads whole 16-row slots
    const int nch = (K + st.kc - 1) / st.kc;
    const float* Wcol = Wsm + tm.cg * 4;
    stage_chunk<NT>(st.buf0, st, 0, min(st.kc, K), nrows, K0, rp0, rp1, tid);
    cp_async_commit();
    for (int c = 0; c < nch; ++c) {
        cp_async_wait<0>();
        __syncthreads();   // chunk c visible to all; everyone is done with chunk c-1's buffer
        if (c + 1 < nch) {
            const int kb = (c + 1) * st.kc;
            stage_chunk<NT>((c & 1) ? st.buf0 : st.buf1, st, kb, min(st.kc, K - kb), nrows, K0, rp0, rp1, tid);
            cp_async_commit();
        }
        if (tm.active) {
            const int kb = c * st.kc;
            mma_chunk(acc, (c & 1) ? st.buf1 : st.buf0, st.lda, min(st.kc, K - kb), Wcol + (long)kb * ldw, ldw, tm, KS, nslots);
        }
    }
    __syncthreads();
    if (KS > 1) {
        // cross-split reduction through the (now idle) staging buffers
        float* scratch = st.buf0;
        const int ncols = ldw;   // scratch row pitch: one row of the weight slice
        if (tm.active && tm.ks > 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
                if (i < nslots)
                    *reinterpret_cast<float4*>(scratch + ((tm.ks - 1) * ROWS + tm.rg + i * RG) * ncols + tm.cg * 4) =
                        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (tm.active && tm.ks == 0) {
            for (int s = 1; s < KS; ++s) {
#pragma unroll
                for (int i = 0; i < RT; ++i)
                    if (i < nslots) {
                        float4 v = *reinterpret_cast<const float4*>(scratch + ((s - 1) * ROWS + tm.rg + i * RG) * ncols + tm.cg * 4);
                        acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
                    }
            }
        }
        __syncthreads();
    }
}

struct NullRow {
    __device__ __forceinline__ const float* operator()(int) const { return nullptr; }
};

// number of active rows of tile `tile` when `bs` sequences are active (sequence b = tile + n_tiles*r)
__device__ __forceinline__ int active_rows(int bs, int tile, int n_tiles) {
    int n = bs - tile;
    return n <= 0 ? 0 : (n + n_tiles - 1) / n_tiles;
}

}  // namespace rnn
