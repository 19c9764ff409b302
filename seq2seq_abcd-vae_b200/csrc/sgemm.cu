// fp32 FFMA GEMM for sm_100a: C[M,N] = alpha * A[M,K] * B[K,N] (+bias)(tanh), register-tiled with
// double-buffered shared-memory staging.  Three operand layouts cover the whole path:
//   NT  (A k-major, B k-major)  y = x W^T           hoisted input projection, MLP forward
//   NN  (A k-major, B n-major)  dx = dy W           data gradients outside the time loop
//   TN  (A m-major, B n-major)  dW = dy^T x         weight gradients, reduction over packed frames
// The TN form takes an optional row-gather on the reduction index (frame -> previous frame of the same
// sequence) and split-K with a deterministic second-pass reduce.  Gate maps let the recurrent kernels keep
// their unit-major interleaved gate order while parameters/gradients stay in PyTorch's gate-major order
// (reference layouts: torch.nn.LSTM/GRU weight_ih_l0 [4H|3H, F], ABCD-VAE/modules/model.py:53).
#include <stdlib.h>
#include "sgemm.cuh"

namespace {

constexpr int BK = 16;

__device__ __forceinline__ int gate_src_row(const GateMap& gm, int phys) {
    // physical interleaved index u*4+slot -> standard row gate*H+u (or -1)
    if (!gm.enabled) return phys;
    int u = phys >> 2, s = phys & 3;
    int g = gm.gate_of_slot[s];
    return g < 0 ? -1 : g * gm.H + u;
}

constexpr int CLK_SIZE = 8;   // cluster split-K factor of the small-problem path

__device__ __forceinline__ float4 ld_dsmem4(const float* local, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(local), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(r) : "memory");
    return v;
}

// CLK: the CTAs of a (1,1,CLK_SIZE) cluster split K; their partial tiles meet in distributed shared memory, every CTA
// reduces BM/CLK_SIZE rows of the tile in a fixed order and applies the epilogue (small problems that cannot fill the GPU
// with output tiles alone: the 512-row sampler / MLP GEMMs; deterministic, no partials workspace).
template <int BM, int BN, int TM, int TN, bool A_KM, bool B_KM, bool CLK = false>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(GemmArgs g, int a_vec, int b_vec) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];
    __shared__ __align__(16) float Cs[CLK ? BM : 1][CLK ? BN + 4 : 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K range (multiples of BK)
    const int ksplit = gridDim.z;
    const int ktiles_total = (g.K + BK - 1) / BK;
    const int ktiles_per = (ktiles_total + ksplit - 1) / ksplit;
    const int kt_begin = blockIdx.z * ktiles_per;
    const int kt_end = min(ktiles_total, kt_begin + ktiles_per);

    constexpr int A_LD = BM * BK / 4 / NT;   // float4 loads per thread per tile
    constexpr int B_LD = BN * BK / 4 / NT;
    static_assert(A_LD >= 1 && B_LD >= 1, "tile too small for thread count");
    float4 ra[A_LD], rb[B_LD];

    auto load_a = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KM) {
                int kq = idx % (BK / 4), m = idx / (BK / 4);
                int gm = m0 + m, gk = k0 + kq * 4;
                if (gm < g.M) {
                    const float* src = g.A + (long)gm * g.lda + gk;
                    if (a_vec && gk + 3 < g.K) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (gk + 0 < g.K) v.x = src[0];
                        if (gk + 1 < g.K) v.y = src[1];
                        if (gk + 2 < g.K) v.z = src[2];
                        if (gk + 3 < g.K) v.w = src[3];
                    }
                }
            } else {
                int mq = idx % (BM / 4), k = idx / (BM / 4);
                int gm = m0 + mq * 4, gk = k0 + k;
                if (gk < g.K) {
                    const float* src = g.A + (long)gk * g.lda + gm;
                    if (a_vec && gm + 3 < g.M) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (gm + 0 < g.M) v.x = src[0];
                        if (gm + 1 < g.M) v.y = src[1];
                        if (gm + 2 < g.M) v.z = src[2];
                        if (gm + 3 < g.M) v.w = src[3];
                    }
                }
            }
            ra[i] = v;
        }
    };
    auto load_b = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KM) {
                int kq = idx % (BK / 4), n = idx / (BK / 4);
                int gn = n0 + n, gk = k0 + kq * 4;
                if (gn < g.N) {
                    int row = gate_src_row(g.b_nsrc, gn);
                    if (row >= 0) {
                        const float* src = g.B + (long)row * g.ldb + gk;
                        if (b_vec && gk + 3 < g.K) v = *reinterpret_cast<const float4*>(src);
                        else {
                            if (gk + 0 < g.K) v.x = src[0];
                            if (gk + 1 < g.K) v.y = src[1];
                            if (gk + 2 < g.K) v.z = src[2];
                            if (gk + 3 < g.K) v.w = src[3];
                        }
                    }
                }
            } else {
                int nq = idx % (BN / 4), k = idx / (BN / 4);
                int gn = n0 + nq * 4, gk = k0 + k;
                if (gk < g.K) {
                    long row = g.b_kgather ? (long)g.b_kgather[gk] : (long)gk;
                    if (row >= 0) {
                        const float* src = g.B + row * g.ldb + gn;
                        if (b_vec && gn + 3 < g.N) v = *reinterpret_cast<const float4*>(src);
                        else {
                            if (gn + 0 < g.N) v.x = src[0];
                            if (gn + 1 < g.N) v.y = src[1];
                            if (gn + 2 < g.N) v.z = src[2];
                            if (gn + 3 < g.N) v.w = src[3];
                        }
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_ab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int idx = tid + i * NT;
            if (A_KM) {
                int kq = idx % (BK / 4), m = idx / (BK / 4);
                As[buf][kq * 4 + 0][m] = ra[i].x;
                As[buf][kq * 4 + 1][m] = ra[i].y;
                As[buf][kq * 4 + 2][m] = ra[i].z;
                As[buf][kq * 4 + 3][m] = ra[i].w;
            } else {
                int mq = idx % (BM / 4), k = idx / (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][k][mq * 4]) = ra[i];
            }
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int idx = tid + i * NT;
            if (B_KM) {
                int kq = idx % (BK / 4), n = idx / (BK / 4);
                Bs[buf][kq * 4 + 0][n] = rb[i].x;
                Bs[buf][kq * 4 + 1][n] = rb[i].y;
                Bs[buf][kq * 4 + 2][n] = rb[i].z;
                Bs[buf][kq * 4 + 3][n] = rb[i].w;
            } else {
                int nq = idx % (BN / 4), k = idx / (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][k][nq * 4]) = rb[i];
            }
        }
    };

    // thread -> output micro-tile: TM rows as TM/4 float4 groups strided by BM/(TM/4), same for cols
    constexpr int TXN = BN / TN, TYN = BM / TM;
    const int tx = tid % TXN, ty = tid / TXN;
    constexpr int MG = TM / 4, NG = TN / 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kt_begin < kt_end) {
        load_a(kt_begin);
        load_b(kt_begin);
        store_ab(0);
        __syncthreads();
        for (int kt = kt_begin; kt < kt_end; ++kt) {
            const int buf = (kt - kt_begin) & 1;
            const bool has_next = kt + 1 < kt_end;
            if (has_next) { load_a(kt + 1); load_b(kt + 1); }
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int gq = 0; gq < MG; ++gq) {
                    float4 v = *reinterpret_cast<const float4*>(&As[buf][k][gq * (TYN * 4) + ty * 4]);
                    a[gq * 4 + 0] = v.x; a[gq * 4 + 1] = v.y; a[gq * 4 + 2] = v.z; a[gq * 4 + 3] = v.w;
                }
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) {
                    float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][gq * (TXN * 4) + tx * 4]);
                    b[gq * 4 + 0] = v.x; b[gq * 4 + 1] = v.y; b[gq * 4 + 2] = v.z; b[gq * 4 + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (has_next) {
                store_ab(buf ^ 1);
                __syncthreads();
            }
        }
    }

    if (CLK) {
        // partial tile -> own shared memory; reduce a slice of the tile across the cluster; epilogue
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int gq = 0; gq < NG; ++gq)
                *reinterpret_cast<float4*>(&Cs[(i / 4) * (TYN * 4) + ty * 4 + (i % 4)][gq * (TXN * 4) + tx * 4]) =
                    make_float4(acc[i][gq * 4], acc[i][gq * 4 + 1], acc[i][gq * 4 + 2], acc[i][gq * 4 + 3]);
        cluster_sync_all();
        const unsigned rank = cluster_ctarank();
        constexpr int RPC = BM / CLK_SIZE;
        for (int idx = tid; idx < RPC * (BN / 4); idx += NT) {
            const int r = rank * RPC + idx / (BN / 4), c4 = idx % (BN / 4);
            float4 sum = ld_dsmem4(&Cs[r][c4 * 4], 0);
#pragma unroll
            for (unsigned pr = 1; pr < CLK_SIZE; ++pr) {
                const float4 v = ld_dsmem4(&Cs[r][c4 * 4], pr);
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            const int m = m0 + r, nb = n0 + c4 * 4;
            if (m >= g.M || nb >= g.N) continue;
            const int mrow = gate_src_row(g.m_dst, m);
            if (mrow < 0) continue;
            float v[4] = {sum.x, sum.y, sum.z, sum.w};
            float* dst = g.C + (long)mrow * g.ldc + nb;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (nb + q >= g.N) continue;
                float t = g.alpha * v[q];
                if (g.bias) {
                    int bi = g.bias_on_m ? mrow : gate_src_row(g.b_nsrc, nb + q);
                    t += g.bias[bi < 0 ? 0 : bi];
                }
                if (g.act == 1) t = tanhf(t);
                dst[q] = g.accumulate ? dst[q] + t : t;
            }
        }
        cluster_sync_all();   // peers may still be reading this CTA's tile
        return;
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (TYN * 4) + ty * 4 + (i % 4);
        if (m >= g.M) continue;
        const int mrow = ksplit > 1 ? m : gate_src_row(g.m_dst, m);
        if (mrow < 0) continue;
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
            const int nb = n0 + gq * (TXN * 4) + tx * 4;   // 4 consecutive output columns
            if (nb >= g.N) continue;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = acc[i][gq * 4 + q];
            float* dst;
            if (ksplit > 1) {
                dst = g.partials + ((long)blockIdx.z * g.M + m) * g.N + nb;
            } else {
                dst = g.C + (long)mrow * g.ldc + nb;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float t = g.alpha * v[q];
                    if (g.bias && nb + q < g.N) {
                        int bi = g.bias_on_m ? mrow : gate_src_row(g.b_nsrc, nb + q);
                        t += g.bias[bi < 0 ? 0 : bi];
                    }
                    if (g.act == 1) t = tanhf(t);
                    v[q] = t;
                }
            }
            const bool vec = (nb + 3 < g.N) && (((uintptr_t)dst & 15) == 0);
            if (vec) {
                float4 o = make_float4(v[0], v[1], v[2], v[3]);
                if (ksplit == 1 && g.accumulate) {
                    float4 old = *reinterpret_cast<const float4*>(dst);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(dst) = o;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (nb + q < g.N) dst[q] = (ksplit == 1 && g.accumulate) ? dst[q] + v[q] : v[q];
            }
        }
    }
}

__device__ __forceinline__ void splitk_reduce_body(const GemmArgs& g, long first, long stride) {
    long total = (long)g.M * g.N;
    for (long idx = first; idx < total; idx += stride) {
        int m = (int)(idx / g.N), n = (int)(idx % g.N);
        float s = 0.f;
        for (int z = 0; z < g.splits; ++z) s += g.partials[(long)z * total + idx];
        int mrow = gate_src_row(g.m_dst, m);
        if (mrow < 0) continue;
        float v = g.alpha * s;
        if (g.bias) {
            int nb = gate_src_row(g.b_nsrc, n);
            v += g.bias_on_m ? g.bias[mrow] : g.bias[nb < 0 ? 0 : nb];
        }
        if (g.act == 1) v = tanhf(v);
        float* dst = g.C + (long)mrow * g.ldc + n;
        if (g.accumulate) v += *dst;
        *dst = v;
    }
}
__global__ void splitk_reduce_kernel(GemmArgs g) {
    splitk_reduce_body(g, blockIdx.x * (long)blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x);
}
// deferred epilogues of several GEMMs in one launch (EpilogueBatch): blocks [bstart[j], bstart[j+1]) work on job j
constexpr int EB_MAX = 16;
struct ReduceJobs {
    int n;
    int bstart[EB_MAX + 1];
    GemmArgs g[EB_MAX];
};
__global__ void splitk_reduce_batch_kernel(const __grid_constant__ ReduceJobs J) {
    int j = 0;
    while (j + 1 < J.n && (int)blockIdx.x >= J.bstart[j + 1]) ++j;
    const int nb = J.bstart[j + 1] - J.bstart[j], lb = blockIdx.x - J.bstart[j];
    splitk_reduce_body(J.g[j], lb * (long)blockDim.x + threadIdx.x, (long)nb * blockDim.x);
}

// Column sums (bias gradients).  blockIdx.x: 32-column strip, blockIdx.y: row split; 8 warps stride the rows of
// the split, deterministic tree reduce in shared memory.  With gridDim.y > 1 the per-split sums go to
// scratch[y][cols] and colsum_finish_kernel adds them in a fixed order.
__device__ __forceinline__ void colsum_body(const float* __restrict__ A, long lda, int rows, int cols, float alpha,
                                            float* __restrict__ out, int accumulate, const GateMap& dst, float* __restrict__ scratch,
                                            int strip, int split, int nsplit) {
    __shared__ float red[8][33];
    const int c = strip * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    const int per = (rows + nsplit - 1) / nsplit;
    const int r0 = split * per, r1 = min(rows, r0 + per);
    float s = 0.f;
    if (c < cols) {
        // four independent partial sums keep four loads in flight (the loop is latency-bound); fixed order
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int r = r0 + w;
        for (; r + 24 < r1; r += 32) {
            s0 += A[(long)r * lda + c];
            s1 += A[(long)(r + 8) * lda + c];
            s2 += A[(long)(r + 16) * lda + c];
            s3 += A[(long)(r + 24) * lda + c];
        }
        for (; r < r1; r += 8) s0 += A[(long)r * lda + c];
        s = (s0 + s1) + (s2 + s3);
    }
    red[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        if (nsplit > 1) {
            scratch[(long)split * cols + c] = t;
        } else {
            int d = gate_src_row(dst, c);
            if (d >= 0) {
                t *= alpha;
                out[d] = accumulate ? out[d] + t : t;
            }
        }
    }
}
__global__ void colsum_kernel(const float* __restrict__ A, long lda, int rows, int cols, float alpha,
                              float* __restrict__ out, int accumulate, GateMap dst, float* __restrict__ scratch) {
    colsum_body(A, lda, rows, cols, alpha, out, accumulate, dst, scratch, blockIdx.x, blockIdx.y, gridDim.y);
}
// strip: 32-column strip; warp w adds splits w, w+8, ... (four loads in flight), warp 0 adds the 8 warps: fixed order
__device__ __forceinline__ void colsum_finish_body(const float* __restrict__ scratch, int nsplit, int cols, float alpha,
                                                   float* __restrict__ out, int accumulate, const GateMap& dst, int strip) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = strip * 32 + lane;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    if (c < cols) {
        int y = w;
        for (; y + 24 < nsplit; y += 32) {
            t0 += scratch[(long)y * cols + c];
            t1 += scratch[(long)(y + 8) * cols + c];
            t2 += scratch[(long)(y + 16) * cols + c];
            t3 += scratch[(long)(y + 24) * cols + c];
        }
        for (; y < nsplit; y += 8) t0 += scratch[(long)y * cols + c];
    }
    red[w][lane] = (t0 + t1) + (t2 + t3);
    __syncthreads();
    if (w == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][lane];
        int d = gate_src_row(dst, c);
        if (d >= 0) {
            t *= alpha;
            out[d] = accumulate ? out[d] + t : t;
        }
    }
}
__global__ void colsum_finish_kernel(const float* __restrict__ scratch, int nsplit, int cols, float alpha,
                                     float* __restrict__ out, int accumulate, GateMap dst) {
    colsum_finish_body(scratch, nsplit, cols, alpha, out, accumulate, dst, blockIdx.x);
}
// deferred column sums of several inputs in one launch per pass (EpilogueBatch)
struct ColsumJob {
    const float* A; long lda; int rows, cols; float alpha; float* out; int accumulate; GateMap dst; float* scratch; int nsplit, strips;
};
struct ColsumJobs {
    int n;
    int bstart[EB_MAX + 1];
    ColsumJob j[EB_MAX];
};
__global__ void colsum_batch_kernel(const __grid_constant__ ColsumJobs J) {
    int j = 0;
    while (j + 1 < J.n && (int)blockIdx.x >= J.bstart[j + 1]) ++j;
    const ColsumJob& q = J.j[j];
    const int lb = blockIdx.x - J.bstart[j];
    colsum_body(q.A, q.lda, q.rows, q.cols, q.alpha, q.out, q.accumulate, q.dst, q.scratch, lb % q.strips, lb / q.strips, q.nsplit);
}
__global__ void colsum_finish_batch_kernel(const __grid_constant__ ColsumJobs J) {   // bstart: strips of the jobs with nsplit > 1
    int j = 0;
    while (j + 1 < J.n && (int)blockIdx.x >= J.bstart[j + 1]) ++j;
    const ColsumJob& q = J.j[j];
    colsum_finish_body(q.scratch, q.nsplit, q.cols, q.alpha, q.out, q.accumulate, q.dst, blockIdx.x - J.bstart[j]);
}

template <int BM, int BN, int TM, int TN>
int launch_cfg(const GemmArgs& g, int a_vec, int b_vec, cudaStream_t stream) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.splits);
    dim3 block((BM / TM) * (BN / TN));
    if (g.a_kmajor && g.b_kmajor) sgemm_kernel<BM, BN, TM, TN, true, true><<<grid, block, 0, stream>>>(g, a_vec, b_vec);
    else if (g.a_kmajor && !g.b_kmajor) sgemm_kernel<BM, BN, TM, TN, true, false><<<grid, block, 0, stream>>>(g, a_vec, b_vec);
    else if (!g.a_kmajor && !g.b_kmajor) sgemm_kernel<BM, BN, TM, TN, false, false><<<grid, block, 0, stream>>>(g, a_vec, b_vec);
    else {
        abcd_set_error("sgemm: (A m-major, B k-major) layout is not on the path");
        return ABCD_ERR_UNSUPPORTED;
    }
    KERNEL_CHECK();
    return ABCD_OK;
}

template <bool A_KM, bool B_KM>
int launch_clk(const GemmArgs& g, int a_vec, int b_vec, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((g.N + 63) / 64, (g.M + 63) / 64, CLK_SIZE);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = CLK_SIZE;
    cfg.attrs = attr; cfg.numAttrs = 1;
    abcd_note_launch(1);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, sgemm_kernel<64, 64, 4, 4, A_KM, B_KM, true>, g, a_vec, b_vec));
    return ABCD_OK;
}

}  // namespace

// 128x128 tiles only when they alone fill the GPU; otherwise 64x64 tiles (4x the CTAs, shorter K loops per CTA)
static bool use_small_tiles(int M, int N, int splits) {
    if (M <= 64 || N <= 64) return true;
    long big = (long)((M + 127) / 128) * ((N + 127) / 128);
    return big * splits < 74;
}

bool umma_tn_eligible(const GemmArgs& g);
bool umma_tn_shape_ok(int M, int N, int K);
int umma_tn_pick_splits(int M, int K);
int umma_tn_launch(const GemmArgs& g, int splits, cudaStream_t stream);
bool umma_rm_eligible(const GemmArgs& g);
int umma_rm_launch(const GemmArgs& g, cudaStream_t stream);

// ABCD_UMMA=0 forces the FFMA path for every GEMM (A/B comparisons, debugging)
static bool umma_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ABCD_UMMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// Split-K factor for the reduction-over-frames (TN) GEMMs; callers size their partials buffer with it.
int sgemm_pick_splits(int M, int N, int K) {
    if (umma_enabled() && umma_tn_shape_ok(M, N, K)) return umma_tn_pick_splits(M, K);
    long tiles = (long)((M + 127) / 128) * ((N + 127) / 128);
    if (M <= 64 || N <= 64) tiles = (long)((M + 63) / 64) * ((N + 63) / 64);
    int ktiles = (K + BK - 1) / BK;
    int s = 1;
    // fill ~2 waves of 148 SMs, keep at least 8 k-tiles per split
    while (tiles * s < 296 && ktiles / (s * 2) >= 8 && s < 64) s *= 2;
    return s;
}

// ---- EpilogueBatch: deferred split-K epilogues and column sums of the calling thread ----
namespace {
struct EpilogueBatchState {
    float* pool; size_t cap, used;
    cudaStream_t stream;
    ReduceJobs rj;
    ColsumJobs cj;
};
thread_local EpilogueBatchState* t_eb = nullptr;
int reduce_blocks(const GemmArgs& g) {
    long total = (long)g.M * g.N;
    int blocks = (int)((total + 255) / 256);
    return blocks > 148 * 8 ? 148 * 8 : (blocks < 1 ? 1 : blocks);
}
int eb_flush(EpilogueBatchState* e) {
    if (e->rj.n > 0) {
        int total = 0;
        for (int j = 0; j < e->rj.n; ++j) { e->rj.bstart[j] = total; total += reduce_blocks(e->rj.g[j]); }
        e->rj.bstart[e->rj.n] = total;
        ProfScope ps(PROF_GEMM, e->stream);
        splitk_reduce_batch_kernel<<<total, 256, 0, e->stream>>>(e->rj);
        KERNEL_CHECK();
        e->rj.n = 0;
    }
    if (e->cj.n > 0) {
        ProfScope ps(PROF_GEMM, e->stream);
        int total = 0;
        for (int j = 0; j < e->cj.n; ++j) { e->cj.bstart[j] = total; total += e->cj.j[j].strips * e->cj.j[j].nsplit; }
        e->cj.bstart[e->cj.n] = total;
        colsum_batch_kernel<<<total, 256, 0, e->stream>>>(e->cj);
        KERNEL_CHECK();
        ColsumJobs fin;
        fin.n = 0;
        total = 0;
        for (int j = 0; j < e->cj.n; ++j)
            if (e->cj.j[j].nsplit > 1) { fin.bstart[fin.n] = total; fin.j[fin.n++] = e->cj.j[j]; total += e->cj.j[j].strips; }
        if (fin.n > 0) {
            fin.bstart[fin.n] = total;
            colsum_finish_batch_kernel<<<total, 256, 0, e->stream>>>(fin);
            KERNEL_CHECK();
        }
        e->cj.n = 0;
    }
    e->used = 0;
    return ABCD_OK;
}
// `floats` floats of the pool (flushing the batch first when it is full); nullptr when the request can never fit
float* eb_take(EpilogueBatchState* e, size_t floats, int* err) {
    floats = (floats + 3) & ~(size_t)3;
    *err = ABCD_OK;
    if (floats > e->cap) return nullptr;
    if (e->used + floats > e->cap) { *err = eb_flush(e); if (*err != ABCD_OK) return nullptr; }
    float* p = e->pool + e->used;
    e->used += floats;
    return p;
}
}  // namespace
EpilogueBatch::EpilogueBatch(float* pool, size_t pool_floats, cudaStream_t stream) {
    EpilogueBatchState* e = new EpilogueBatchState();
    e->pool = pool; e->cap = pool ? pool_floats : 0; e->used = 0; e->stream = stream;
    e->rj.n = 0; e->cj.n = 0;
    t_eb = e;
}
int EpilogueBatch::flush() { return t_eb ? eb_flush(t_eb) : ABCD_OK; }
EpilogueBatch::~EpilogueBatch() {
    if (t_eb) { eb_flush(t_eb); delete t_eb; t_eb = nullptr; }
}
// split-K epilogue of `g`: deferred into the active batch, or launched now
static int reduce_or_defer(const GemmArgs& g, cudaStream_t stream) {
    if (t_eb && t_eb->stream == stream) {
        if (t_eb->rj.n == EB_MAX) TRY(eb_flush(t_eb));   // (never with a job's partials still pending: they were taken before)
        t_eb->rj.g[t_eb->rj.n++] = g;
        return ABCD_OK;
    }
    splitk_reduce_kernel<<<reduce_blocks(g), 256, 0, stream>>>(g);
    KERNEL_CHECK();
    return ABCD_OK;
}

int sgemm_launch(const GemmArgs& gin, cudaStream_t stream) {
    GemmArgs g = gin;
    if (g.M <= 0 || g.N <= 0) return ABCD_OK;
    ARG_CHECK(g.K >= 0 && g.A && g.B && g.C, "sgemm: null operand or negative K");
    if (g.splits < 1) g.splits = 1;
    ARG_CHECK(g.splits == 1 || g.partials, "sgemm: split-K needs a partials workspace");
    ARG_CHECK(!(g.b_kgather && g.b_kmajor), "sgemm: k-gather needs n-major B");
    ARG_CHECK(!(g.b_nsrc.enabled && !g.b_kmajor), "sgemm: n source map needs k-major B");
    ProfScope ps(PROF_GEMM, stream);
    if (umma_enabled() && umma_rm_eligible(g)) return umma_rm_launch(g, stream);   // tall row-major products on tcgen05
    const bool tn = umma_enabled() && umma_tn_eligible(g);
    if (t_eb && t_eb->stream == stream && (tn || g.splits > 1)) {
        // deferred epilogue: the partials live in the batch's pool until its flush
        if (t_eb->rj.n == EB_MAX) TRY(eb_flush(t_eb));
        int err;
        float* part = eb_take(t_eb, (size_t)g.splits * g.M * g.N, &err);
        if (err != ABCD_OK) return err;
        if (part) g.partials = part;
        else TRY(eb_flush(t_eb));   // larger than the pool: the caller's own buffer, nothing else may be pending in it
    }
    if (tn) {
        // tensor-core path (tcgen05 3xTF32): partial tiles per split, then the common reduce/epilogue kernel
        TRY(umma_tn_launch(g, g.splits, stream));
        return reduce_or_defer(g, stream);
    }
    const int a_vec = ((uintptr_t)g.A % 16 == 0) && (g.lda % 4 == 0);
    const int b_vec = ((uintptr_t)g.B % 16 == 0) && (g.ldb % 4 == 0);
    int r;
    // small problems (few output tiles, long K): split K over an 8-CTA cluster, reduce through distributed shared memory
    const long tiles64 = (long)((g.M + 63) / 64) * ((g.N + 63) / 64);
    if (g.splits == 1 && !g.b_kgather && tiles64 <= 64 && g.K >= 256 && !(!g.a_kmajor && g.b_kmajor)) {
        if (g.a_kmajor && g.b_kmajor) return launch_clk<true, true>(g, a_vec, b_vec, stream);
        if (g.a_kmajor && !g.b_kmajor) return launch_clk<true, false>(g, a_vec, b_vec, stream);
        return launch_clk<false, false>(g, a_vec, b_vec, stream);
    }
    if (use_small_tiles(g.M, g.N, g.splits)) r = launch_cfg<64, 64, 4, 4>(g, a_vec, b_vec, stream);
    else r = launch_cfg<128, 128, 8, 8>(g, a_vec, b_vec, stream);
    if (r != ABCD_OK) return r;
    if (g.splits > 1) return reduce_or_defer(g, stream);
    return ABCD_OK;
}

static thread_local float* t_colsum_scratch = nullptr;
static thread_local size_t t_colsum_floats = 0;
ColsumScratch::ColsumScratch(float* p, size_t floats) { t_colsum_scratch = p; t_colsum_floats = floats; }
ColsumScratch::~ColsumScratch() { t_colsum_scratch = nullptr; t_colsum_floats = 0; }

int colsum_launch(const float* A, long lda, int rows, int cols, float alpha, float* out, int accumulate,
                  GateMap dst, cudaStream_t stream) {
    float* scratch = t_colsum_scratch;
    size_t scratch_floats = t_colsum_floats;
    if (cols <= 0) return ABCD_OK;
    const int strips = (cols + 31) / 32;
    if (t_eb && t_eb->stream == stream) {
        // deferred: one batched launch per pass at the flush; scratch from the batch's pool
        if (t_eb->cj.n == EB_MAX) TRY(eb_flush(t_eb));
        int nsplit = (148 * 4 + strips - 1) / strips;
        if (nsplit > rows / 64) nsplit = rows / 64;
        if (nsplit < 1) nsplit = 1;
        float* sc = nullptr;
        if (nsplit > 1) {
            int err;
            sc = eb_take(t_eb, (size_t)nsplit * cols, &err);
            if (err != ABCD_OK) return err;
            if (!sc) nsplit = 1;
        }
        ColsumJob& q = t_eb->cj.j[t_eb->cj.n++];
        q.A = A; q.lda = lda; q.rows = rows; q.cols = cols; q.alpha = alpha; q.out = out; q.accumulate = accumulate; q.dst = dst;
        q.scratch = sc; q.nsplit = nsplit; q.strips = strips;
        return ABCD_OK;
    }
    ProfScope ps(PROF_GEMM, stream);
    // enough row splits to cover the GPU (~4 CTAs/SM), at least 64 rows per split
    int nsplit = 1;
    if (scratch) {
        nsplit = (148 * 4 + strips - 1) / strips;
        if (nsplit > rows / 64) nsplit = rows / 64;
        if ((size_t)nsplit * cols > scratch_floats) nsplit = (int)(scratch_floats / cols);
        if (nsplit < 1) nsplit = 1;
    }
    colsum_kernel<<<dim3(strips, nsplit), 256, 0, stream>>>(A, lda, rows, cols, alpha, out, accumulate, dst, scratch);
    KERNEL_CHECK();
    if (nsplit > 1) {
        colsum_finish_kernel<<<strips, 256, 0, stream>>>(scratch, nsplit, cols, alpha, out, accumulate, dst);
        KERNEL_CHECK();
    }
    return ABCD_OK;
}
