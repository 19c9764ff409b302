// Building blocks of the persistent, cluster-partitioned recurrent kernels (enc.cu, dec.cu).
//
// Execution model.  One thread-block cluster of CL=16 CTAs owns one "row tile" of up to 64 segments
// (sequence b belongs to tile b % n_tiles, so every tile sees the same length distribution and row
// activity inside a tile is a shrinking prefix, exactly like PackedSequence.batch_sizes).  Each CTA owns
// 16 of the 256 hidden units and keeps its slice of every weight matrix resident in shared memory for the
// whole launch.  A time step is a short sequence of phases; in each phase a CTA computes
//     out[rows, my columns] = A[rows, K] * Wslice[K, my columns]
// where the A rows are activations written to global memory (L2) by all CTAs of the cluster in the previous
// phase.  Those activation buffers are the tensors the backward pass needs anyway, so the exchange costs no
// extra traffic.  Phases are separated by the hardware cluster barrier (release/acquire), not a grid sync.
//
// phase_gemm(): A rows are streamed L2 -> shared memory with cp.async.cg in double-buffered K-chunks; each
// thread accumulates a 4-row x 4-column register tile (for gate GEMMs the 4 columns are the 4 gates of one
// hidden unit, so the cell update is thread-local).  Small-N phases split K across thread groups and reduce
// through shared memory.
#pragma once
#include "common.cuh"

namespace rnn {

constexpr int CL = 16;             // CTAs per cluster
constexpr int U = ABCD_H / CL;     // hidden units per CTA
constexpr int ROWS = 64;           // rows (segments) per tile
constexpr int RT = 4;              // rows per thread
constexpr int RG = ROWS / RT;      // row groups; thread rg owns rows rg + i*RG
constexpr int NTHREADS = 256;

struct ThreadMap {
    int cg, rg, ks;   // column group (4 columns), row group, K-split index
    bool active;
};
__device__ __forceinline__ ThreadMap make_map(int tid, int ncg, int KS) {
    ThreadMap m;
    m.cg = tid % ncg;
    m.rg = (tid / ncg) % RG;
    m.ks = tid / (ncg * RG);
    m.active = m.ks < KS;
    return m;
}
__host__ __device__ inline int pick_ks(int ncg) {
    int ks = NTHREADS / (ncg * RG);
    return ks < 1 ? 1 : ks;
}

struct Stage {
    float* buf0;
    float* buf1;
    int lda;   // kc + 4 floats
    int kc;    // K-chunk (multiple of 4)
};

__device__ __forceinline__ void zero_acc(float (&acc)[RT][4]) {
#pragma unroll
    for (int i = 0; i < RT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
}

// A row r of the tile is the concatenation [rp0(r)[0,K0) | rp1(r)[0,K1)] (a null pointer means zeros).
// K0 and K1 are multiples of 4 and all row pointers are 16-byte aligned.
template <class RP0, class RP1>
__device__ __forceinline__ void stage_chunk(float* dst, int lda, int k_begin, int k_len, int nrows, int K0,
                                            RP0 rp0, RP1 rp1, int tid) {
    const int ppr = k_len >> 2;   // 16-byte pieces per row
    const int total = nrows * ppr;
    for (int idx = tid; idx < total; idx += NTHREADS) {
        int r = idx / ppr, p = idx - r * ppr;
        int k = k_begin + p * 4;
        const float* src;
        if (k < K0) { src = rp0(r); if (src) src += k; }
        else { src = rp1(r); if (src) src += (k - K0); }
        float* d = dst + r * lda + p * 4;
        if (src) cp_async16(d, src);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ void mma_chunk(float (&acc)[RT][4], const float* As, int lda, int k_len,
                                          const float* Wk /* &Wsm[k_begin][cg*4] */, int ldw, const ThreadMap& tm,
                                          int KS, int nslots) {
    const int nk4 = k_len >> 2;
    for (int k4 = tm.ks; k4 < nk4; k4 += KS) {
        const float* w = Wk + (k4 * 4) * ldw;
        const float4 w0 = *reinterpret_cast<const float4*>(w);
        const float4 w1 = *reinterpret_cast<const float4*>(w + ldw);
        const float4 w2 = *reinterpret_cast<const float4*>(w + 2 * ldw);
        const float4 w3 = *reinterpret_cast<const float4*>(w + 3 * ldw);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            if (i < nslots) {
                const float4 a = *reinterpret_cast<const float4*>(As + (tm.rg + i * RG) * lda + k4 * 4);
                acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
                acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
                acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
                acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
                acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
                acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
                acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
                acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
            }
        }
    }
}

// acc += A[rows, K0+K1] * Wsm[K0+K1, ncols] for this thread's tile.  On return the threads with ks == 0
// hold the full sums for rows rg + i*RG (i < RT) and columns cg*4 .. cg*4+3.  Ends with a CTA barrier, so the
// staging buffers may be reused immediately.
template <class RP0, class RP1>
__device__ __forceinline__ void phase_gemm(float (&acc)[RT][4], const Stage& st, int K0, RP0 rp0, int K1,
                                           RP1 rp1, int nrows, const float* Wsm, int ldw, const ThreadMap& tm,
                                           int KS, int tid) {
    const int K = K0 + K1;
    const int nch = (K + st.kc - 1) / st.kc;
    const int nslots = (nrows + RG - 1) / RG;
    const float* Wcol = Wsm + tm.cg * 4;
    stage_chunk(st.buf0, st.lda, 0, min(st.kc, K), nrows, K0, rp0, rp1, tid);
    cp_async_commit();
    for (int c = 0; c < nch; ++c) {
        float* cur = (c & 1) ? st.buf1 : st.buf0;
        if (c + 1 < nch) {
            float* nxt = (c & 1) ? st.buf0 : st.buf1;
            int kb = (c + 1) * st.kc;
            stage_chunk(nxt, st.lda, kb, min(st.kc, K - kb), nrows, K0, rp0, rp1, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (tm.active) {
            int kb = c * st.kc;
            mma_chunk(acc, cur, st.lda, min(st.kc, K - kb), Wcol + (long)kb * ldw, ldw, tm, KS, nslots);
        }
        __syncthreads();
    }
    if (KS > 1) {
        // cross-split reduction through the (now idle) staging buffer 0
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
