// Generic fp32 GEMM used for every non-recurrent contraction on the path (hoisted input projection,
// MLPs outside the time loop, weight gradients over all packed frames).  See sgemm.cu.
#pragma once
#include "common.cuh"

// Gate-slot map between the PyTorch "standard" gate-major row order (row = g*H + u) and the
// unit-major interleaved order the recurrent kernels use (col = u*4 + slot).
struct GateMap {
    int enabled;   // 0: identity
    int H;         // units
    int gate_of_slot[4];  // slot -> standard gate index, or -1 (slot unused)
};
static inline GateMap gate_map_none() {
    GateMap g;
    g.enabled = 0; g.H = 0;
    g.gate_of_slot[0] = g.gate_of_slot[1] = g.gate_of_slot[2] = g.gate_of_slot[3] = -1;
    return g;
}
static inline GateMap gate_map(int H, int s0, int s1, int s2, int s3) {
    GateMap g;
    g.enabled = 1; g.H = H;
    g.gate_of_slot[0] = s0; g.gate_of_slot[1] = s1; g.gate_of_slot[2] = s2; g.gate_of_slot[3] = s3;
    return g;
}

struct GemmArgs {
    int M, N, K;
    // A(m,k): a_kmajor ? A[m*lda + k] : A[k*lda + m]
    const float* A; long lda; int a_kmajor;
    // B(k,n): b_kmajor ? B[n*ldb + k] : B[k*ldb + n]
    const float* B; long ldb; int b_kmajor;
    const int* b_kgather;   // optional (n-major B only): source row of B for reduction index k, -1 => zero row
    GateMap b_nsrc;         // optional (k-major B only): physical n (interleaved) -> source row of B / bias index
    GateMap m_dst;          // optional: physical m (interleaved) -> destination row of C / bias index
    float* C; long ldc;     // C[m*ldc + n]
    float alpha;
    const float* bias; int bias_on_m;   // bias over n (default) or over m
    int act;                // 0 none, 1 tanh
    int accumulate;         // C += ...
    int splits;             // split-K factor (>=1); >1 needs `partials` of splits*M*N floats
    float* partials;
};
static inline GemmArgs gemm_args_zero() {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.alpha = 1.f; g.splits = 1;
    g.b_nsrc = gate_map_none(); g.m_dst = gate_map_none();
    return g;
}

// Enqueue C = alpha * A*B (+bias)(tanh) on `stream`. Returns ABCD_OK or an error code.
int sgemm_launch(const GemmArgs& g, cudaStream_t stream);
// Heuristic split-K factor for an (M,N,K) problem and the number of floats of partial workspace.
int sgemm_pick_splits(int M, int N, int K);

// out[c] (+)= alpha * sum_r A[r*lda + c]  (bias gradients), with optional gate map on the column index.
// A caller-owned scratch buffer (see ColsumScratch) enables a row-split two-pass reduction for tall inputs.
int colsum_launch(const float* A, long lda, int rows, int cols, float alpha, float* out, int accumulate,
                  GateMap dst, cudaStream_t stream);
// RAII: inside its scope the split-K reduce / epilogue kernels of sgemm_launch and the column sums of colsum_launch issued by
// this thread are DEFERRED and run as at most three batched launches at flush() (or at the end of the scope): the
// weight-gradient sections issue ~10 GEMMs and ~10 column sums back to back, whose ~30 tiny epilogue launches become three.
// `pool` (caller's workspace) holds every deferred GEMM's split-K partials and the column sums' scratch until the flush; when
// it runs out the batch is flushed early.  GemmArgs::partials / ColsumScratch are ignored inside the scope.
struct EpilogueBatch {
    EpilogueBatch(float* pool, size_t pool_floats, cudaStream_t stream);
    int flush();
    ~EpilogueBatch();
};
// RAII: lends `floats` floats of the caller's workspace to colsum_launch calls made by this thread in this scope.
struct ColsumScratch {
    ColsumScratch(float* p, size_t floats);
    ~ColsumScratch();
};
