// The ABCD sampler as ONE forward and ONE backward kernel (BASELINE.json north star, item 3; reference:
// ABCD-VAE/modules/model.py:581-639).  Input is the output xc [B, f] of the to_code_like MLP (two GEMMs, sampler.cu):
//
//   forward   logits = xc codebook / sqrt(f)                                  (model.py:588-589)
//             y      = softmax(logits)  |  softmax((logits + Gumbel) / tau)   (model.py:601-604; noise from the caller or
//                                                                              drawn in the kernel from the Philox stream)
//             feats  = y codebook^T                                           (model.py:605)
//             KL     = Dirichlet terms * B/N + sum_b sum_k q (log q - E_k)    (model.py:615-639)
//   backward  d logits = (1/tau) y (dY - <y, dY>) + g_KL q ((log q - E) - <q, log q - E>),   dY = d feats codebook
//             d xc     = d logits codebook^T / sqrt(f),   d phi (posterior_shape_logits) in closed form,
//             d codebook = xc^T d logits / sqrt(f) + d feats^T y   (two reductions over the batch, left to the GEMM path)
//
// Layout: a CTA keeps the whole codebook [f, K] in shared memory with an odd row pitch (K | 1 floats), so that both walks are
// bank-conflict free: along k for a fixed j (logits, dY: lane = category) and along j for a fixed k (feats, d xc:
// lane = feature).  One warp per segment, 8 segments per CTA pass; the per-row reductions are warp shuffles.  The batch
// reductions (KL, column sums Q_k of q) are per-CTA partials added in a FIXED order by whichever CTA finishes last
// (ticket counter), so results are reproducible.  Special functions of the Dirichlet terms in fp64, as in sampler.cu.
#include "sampler.cuh"
#include "sgemm.cuh"
#include "philox.cuh"

namespace {

constexpr int SF_THREADS = 256;
constexpr int SF_WARPS = SF_THREADS / 32;
constexpr int SF_MAX_KI = 8;     // categories per lane: K <= 256
constexpr int SF_MAX_FI = 16;    // features per lane:   f <= 512
constexpr int SF_MAX_CTAS = 148;

__device__ double sf_digamma(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
__device__ double sf_trigamma(double x) {
    double r = 0.0;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    double f = 1.0 / (x * x);
    return r + 1.0 / x + 0.5 * f + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30))));
}
__device__ __forceinline__ double block_max_d(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double m = red[0];
    for (int i = 1; i < SF_WARPS; ++i) m = fmax(m, red[i]);
    __syncthreads();
    return m;
}

// Dirichlet terms (model.py:615-626), computed redundantly by every CTA (K special-function evaluations):
//   s = softmax(phi), alpha = s N + c, E_k = psi(alpha_k) - psi(sum alpha);  returns q_pi - p_pi (valid in every thread)
// E_sm / s_sm: [K] shared-memory outputs.  Same fp32 rounding points for s and alpha as the reference.
__device__ double dirichlet_block(int K, const float* __restrict__ phi, float prior_c, float n_data, float* E_sm, float* s_sm, double* red) {
    const int tid = threadIdx.x;
    double mx = -1e300;
    for (int k = tid; k < K; k += SF_THREADS) mx = fmax(mx, (double)phi[k]);
    mx = block_max_d(mx, red);
    double z = 0.0;
    for (int k = tid; k < K; k += SF_THREADS) z += exp((double)phi[k] - mx);
    z = block_sum_d(z, red);
    double asum = 0.0;
    for (int k = tid; k < K; k += SF_THREADS) {
        const float sk = (float)(exp((double)phi[k] - mx) / z);
        s_sm[k] = sk;
        asum += (double)(sk * n_data + prior_c);
    }
    asum = block_sum_d(asum, red);
    const double psiA = sf_digamma(asum);
    double t_lg = 0.0, t_am1E = 0.0, t_E = 0.0;
    for (int k = tid; k < K; k += SF_THREADS) {
        const double ak = (double)(s_sm[k] * n_data + prior_c);
        const double e = sf_digamma(ak) - psiA;
        E_sm[k] = (float)e;
        t_lg += lgamma(ak);
        t_am1E += (ak - 1.0) * e;
        t_E += e;
    }
    t_lg = block_sum_d(t_lg, red);
    t_am1E = block_sum_d(t_am1E, red);
    t_E = block_sum_d(t_E, red);
    const double c = (double)prior_c;
    const double q_pi = lgamma(asum) - t_lg + t_am1E;
    const double p_pi = lgamma(c * K) - lgamma(c) * K + (c - 1.0) * t_E;
    return q_pi - p_pi;
}

struct SfArgs {
    int B, K, f, mode;
    int rpc;               // segments per CTA pass (<= 8 warps): small batches spread over more SMs, each streaming the codebook for fewer rows
    float inv_tau, alpha, prior_c, n_data;
    const float* codebook; const float* phi;
    const float* xc; const float* gumbel;
    unsigned long long seed, offset; const unsigned long long* d_base;
    float* logits; float* y; float* feats; float* kl_out;
    float* partial;        // [gridDim.x] KL partials
    unsigned int* ticket;  // zero on entry; reset by the last CTA
    // backward
    const float* d_feats; const float* g_kl;
    float* dl_scaled;      // [B, K] d logits / sqrt(f): the left factor of the codebook-gradient GEMM
    float* d_xc;           // [B, f]
    float* Qpart;          // [gridDim.x][K]
    float* d_phi;          // accumulated
};

// codebook [f][K] (global, row-major) -> shared memory with row pitch P; batches of independent 16-byte loads in flight
__device__ __forceinline__ void load_codebook(float* cb, const float* __restrict__ g, int f, int K) {
    const int P = K | 1;   // odd row pitch: a column walk (fixed k, consecutive j) touches 32 different banks
    if ((K & 3) == 0) {
        const int K4 = K >> 2, n4 = f * K4;
        const float4* g4 = reinterpret_cast<const float4*>(g);
        constexpr int UNR = 8;
        for (int base = threadIdx.x; base < n4; base += SF_THREADS * UNR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = base + u * SF_THREADS;
                v[u] = idx < n4 ? __ldg(g4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = base + u * SF_THREADS;
                if (idx < n4) {
                    const int j = idx / K4, k = (idx - j * K4) << 2;
                    float* d = cb + j * P + k;
                    d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                }
            }
        }
    } else {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int j = w; j < f; j += SF_WARPS)
            for (int k = lane; k < K; k += 32) cb[j * P + k] = __ldg(g + (long)j * K + k);
    }
}

// out[k] = sum_j v[j] cb[j][k] for this lane's categories k = lane + 32 i  (KI = ceil(K / 32) <= SF_MAX_KI)
template <int KI>
__device__ __forceinline__ void dot_over_features_t(const float* cb, const float* v, int f, int K, int lane, float (&out)[SF_MAX_KI]) {
    const int P = K | 1;
    float acc[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) acc[i] = 0.f;
    const bool full = K == 32 * KI;
    int j = 0;
    for (; j + 4 <= f; j += 4) {   // four rows of the codebook in flight per iteration
        float a[4], c[4][KI];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = v[j + u];
            const float* row = cb + (j + u) * P + lane;
#pragma unroll
            for (int i = 0; i < KI; ++i) c[u][i] = (full || lane + 32 * i < K) ? row[32 * i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < KI; ++i) acc[i] = fmaf(a[u], c[u][i], acc[i]);
    }
    for (; j < f; ++j) {
        const float a = v[j];
        const float* row = cb + j * P + lane;
#pragma unroll
        for (int i = 0; i < KI; ++i)
            if (full || lane + 32 * i < K) acc[i] = fmaf(a, row[32 * i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < SF_MAX_KI; ++i) out[i] = i < KI ? acc[i < KI ? i : 0] : 0.f;
}
__device__ __forceinline__ void dot_over_features(const float* cb, const float* v, int f, int K, int lane, float (&out)[SF_MAX_KI]) {
    const int ki = (K + 31) >> 5;
    if (ki <= 1) dot_over_features_t<1>(cb, v, f, K, lane, out);
    else if (ki <= 2) dot_over_features_t<2>(cb, v, f, K, lane, out);
    else if (ki <= 4) dot_over_features_t<4>(cb, v, f, K, lane, out);
    else dot_over_features_t<8>(cb, v, f, K, lane, out);
}
// out[j] = sum_k v[k] cb[j][k] for this lane's features j = lane + 32 m  (FI = ceil(f / 32) <= SF_MAX_FI)
template <int FI>
__device__ __forceinline__ void dot_over_categories_t(const float* cb, const float* v, int f, int K, int lane, float (&out)[SF_MAX_FI]) {
    const int P = K | 1;
    float acc[FI];
#pragma unroll
    for (int m = 0; m < FI; ++m) acc[m] = 0.f;
    const bool full = f == 32 * FI;
    const float* col = cb + lane * P;
    int k = 0;
    for (; k + 2 <= K; k += 2) {
        float a[2], c[2][FI];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u] = v[k + u];
#pragma unroll
            for (int m = 0; m < FI; ++m) c[u][m] = (full || lane + 32 * m < f) ? col[32 * m * P + k + u] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int m = 0; m < FI; ++m) acc[m] = fmaf(a[u], c[u][m], acc[m]);
    }
    for (; k < K; ++k) {
        const float a = v[k];
#pragma unroll
        for (int m = 0; m < FI; ++m)
            if (full || lane + 32 * m < f) acc[m] = fmaf(a, col[32 * m * P + k], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < SF_MAX_FI; ++m) out[m] = m < FI ? acc[m < FI ? m : 0] : 0.f;
}
__device__ __forceinline__ void dot_over_categories(const float* cb, const float* v, int f, int K, int lane, float (&out)[SF_MAX_FI]) {
    const int fi = (f + 31) >> 5;
    if (fi <= 2) dot_over_categories_t<2>(cb, v, f, K, lane, out);
    else if (fi <= 4) dot_over_categories_t<4>(cb, v, f, K, lane, out);
    else if (fi <= 8) dot_over_categories_t<8>(cb, v, f, K, lane, out);
    else dot_over_categories_t<16>(cb, v, f, K, lane, out);
}

// dynamic shared memory: cb [f][K+1] | E [K] | s [K] | rowf [8][f] | rowk [8][K] | qw [8][K] (backward only)
__global__ void __launch_bounds__(SF_THREADS, 1) sampler_fused_fwd_kernel(SfArgs a) {
    extern __shared__ __align__(16) float sf_smem[];
    __shared__ double red[40];
    __shared__ float redf[40];
    __shared__ unsigned int last_sh;
    const int K = a.K, f = a.f, P = K | 1;
    float* cb = sf_smem;
    float* E = cb + f * P;
    float* s_sm = E + K;
    float* rowf = s_sm + K;
    float* rowk = rowf + SF_WARPS * f;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    load_codebook(cb, a.codebook, f, K);
    const double prior = dirichlet_block(K, a.phi, a.prior_c, a.n_data, E, s_sm, red);
    __syncthreads();
    float klacc = 0.f;
    float* xr = rowf + w * f;
    float* yr = rowk + w * K;
    unsigned long long base = a.offset;
    if (a.d_base) base += __ldg(a.d_base);
    for (int b = blockIdx.x * a.rpc + w; w < a.rpc && b < a.B; b += gridDim.x * a.rpc) {
        for (int j = lane; j < f; j += 32) xr[j] = a.xc[(long)b * f + j];
        __syncwarp();
        float l[SF_MAX_KI];
        dot_over_features(cb, xr, f, K, lane, l);
        float v[SF_MAX_KI];
        float mx = -INFINITY, mxl = -INFINITY;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i) {
            const int k = lane + 32 * i;
            if (k < K) {
                l[i] *= a.alpha;
                a.logits[(long)b * K + k] = l[i];
                float g = 0.f;
                if (a.mode) {
                    if (a.gumbel) g = a.gumbel[(long)b * K + k];
                    else {
                        const unsigned long long e = (unsigned long long)b * K + k;
                        uint32_t r[4];
                        philox4(a.seed, base + (e >> 2), r);
                        const uint32_t c = (e & 3) == 0 ? r[0] : (e & 3) == 1 ? r[1] : (e & 3) == 2 ? r[2] : r[3];
                        g = -logf(-logf(u01(c)));
                    }
                }
                v[i] = a.mode ? (l[i] + g) * a.inv_tau : l[i];
                mx = fmaxf(mx, v[i]);
                mxl = fmaxf(mxl, l[i]);
            }
        }
        mx = warp_max(mx);
        mxl = warp_max(mxl);
        float z = 0.f, zl = 0.f;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i)
            if (lane + 32 * i < K) { z += expf(v[i] - mx); zl += expf(l[i] - mxl); }
        z = warp_sum(z);
        zl = warp_sum(zl);
        const float iz = 1.f / z, lzl = logf(zl);
        float srow = 0.f;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i) {
            const int k = lane + 32 * i;
            if (k < K) {
                const float yv = expf(v[i] - mx) * iz;
                a.y[(long)b * K + k] = yv;
                yr[k] = yv;
                const float lq = l[i] - mxl - lzl;
                srow += expf(lq) * (lq - E[k]);
            }
        }
        srow = warp_sum(srow);
        if (lane == 0) klacc += srow;
        __syncwarp();
        float o[SF_MAX_FI];
        dot_over_categories(cb, yr, f, K, lane, o);
#pragma unroll
        for (int m = 0; m < SF_MAX_FI; ++m)
            if (lane + 32 * m < f) a.feats[(long)b * f + lane + 32 * m] = o[m];
        __syncwarp();
    }
    // KL: per-CTA partial, then the last CTA adds all partials in a fixed order
    const float tot = block_sum(klacc, redf);
    if (tid == 0) {
        a.partial[blockIdx.x] = tot;
        __threadfence();
        last_sh = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (last_sh) {
        __threadfence();
        double sacc = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += SF_THREADS) sacc += (double)ld_cg(a.partial + i);
        sacc = block_sum_d(sacc, red);
        if (tid == 0) {
            *a.kl_out = (float)(prior * (double)((float)a.B / a.n_data) + sacc);
            *a.ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(SF_THREADS, 1) sampler_fused_bwd_kernel(SfArgs a) {
    extern __shared__ __align__(16) float sf_smem[];
    __shared__ double red[40];
    __shared__ unsigned int last_sh;
    const int K = a.K, f = a.f, P = K | 1;
    float* cb = sf_smem;
    float* E = cb + f * P;
    float* s_sm = E + K;
    float* rowf = s_sm + K;
    float* rowk = rowf + SF_WARPS * f;
    float* qw = rowk + SF_WARPS * K;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    load_codebook(cb, a.codebook, f, K);
    (void)dirichlet_block(K, a.phi, a.prior_c, a.n_data, E, s_sm, red);
    for (int k = tid; k < SF_WARPS * K; k += SF_THREADS) qw[k] = 0.f;
    __syncthreads();
    const float g = *a.g_kl;
    float* dr = rowf + w * f;
    float* lr = rowk + w * K;
    float* qrow = qw + w * K;
    for (int b = blockIdx.x * a.rpc + w; w < a.rpc && b < a.B; b += gridDim.x * a.rpc) {
        for (int j = lane; j < f; j += 32) dr[j] = a.d_feats[(long)b * f + j];
        __syncwarp();
        float dY[SF_MAX_KI];
        dot_over_features(cb, dr, f, K, lane, dY);
        float l[SF_MAX_KI], yv[SF_MAX_KI];
        float sdot = 0.f, mxl = -INFINITY;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i) {
            const int k = lane + 32 * i;
            l[i] = 0.f; yv[i] = 0.f;
            if (k < K) {
                l[i] = a.logits[(long)b * K + k];
                yv[i] = a.y[(long)b * K + k];
                sdot += yv[i] * dY[i];
                mxl = fmaxf(mxl, l[i]);
            }
        }
        sdot = warp_sum(sdot);
        mxl = warp_max(mxl);
        float zl = 0.f;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i)
            if (lane + 32 * i < K) zl += expf(l[i] - mxl);
        zl = warp_sum(zl);
        const float lzl = logf(zl);
        float srow = 0.f;
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i) {
            const int k = lane + 32 * i;
            if (k < K) { const float lq = l[i] - mxl - lzl; srow += expf(lq) * (lq - E[k]); }
        }
        srow = warp_sum(srow);
#pragma unroll
        for (int i = 0; i < SF_MAX_KI; ++i) {
            const int k = lane + 32 * i;
            if (k < K) {
                const float lq = l[i] - mxl - lzl, q = expf(lq);
                const float dl = a.inv_tau * yv[i] * (dY[i] - sdot) + g * q * ((lq - E[k]) - srow);
                const float dls = dl * a.alpha;
                a.dl_scaled[(long)b * K + k] = dls;
                lr[k] = dls;
                qrow[k] += q;
            }
        }
        __syncwarp();
        float o[SF_MAX_FI];
        dot_over_categories(cb, lr, f, K, lane, o);
#pragma unroll
        for (int m = 0; m < SF_MAX_FI; ++m)
            if (lane + 32 * m < f) a.d_xc[(long)b * f + lane + 32 * m] = o[m];
        __syncwarp();
    }
    __syncthreads();
    // column sums Q_k = sum_b q_bk: warps in order, then CTAs in order (last CTA), then the closed-form d phi
    for (int k = tid; k < K; k += SF_THREADS) {
        float t = 0.f;
        for (int i = 0; i < SF_WARPS; ++i) t += qw[i * K + k];
        a.Qpart[(size_t)blockIdx.x * K + k] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last_sh = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (last_sh) {
        __threadfence();
        float* Q = qw;   // reuse: [K]
        for (int k = tid; k < K; k += SF_THREADS) {
            float t = 0.f;
            for (int i = 0; i < (int)gridDim.x; ++i) t += ld_cg(a.Qpart + (size_t)i * K + k);
            Q[k] = t;
        }
        __syncthreads();
        double dot = 0.0;
        for (int k = tid; k < K; k += SF_THREADS) {
            const double ak = (double)(s_sm[k] * a.n_data + a.prior_c);
            const double h = sf_trigamma(ak) * ((double)a.B * s_sm[k] - (double)Q[k]);
            dot += (double)s_sm[k] * h;
        }
        dot = block_sum_d(dot, red);
        for (int k = tid; k < K; k += SF_THREADS) {
            const double ak = (double)(s_sm[k] * a.n_data + a.prior_c);
            const double h = sf_trigamma(ak) * ((double)a.B * s_sm[k] - (double)Q[k]);
            a.d_phi[k] += (float)((double)g * (double)a.n_data * s_sm[k] * (h - dot));
        }
        if (tid == 0) *a.ticket = 0u;
    }
}

size_t sf_smem_bytes(int f, int K, bool bwd) {
    return ((size_t)f * (K | 1) + 2 * (size_t)K + (size_t)SF_WARPS * f + (size_t)SF_WARPS * K * (bwd ? 2 : 1)) * sizeof(float);
}
int sf_rows_per_cta(int B) { return B <= 2 * SF_MAX_CTAS ? 2 : (B <= 4 * SF_MAX_CTAS ? 4 : SF_WARPS); }
int sf_grid(int B) {
    const int r = sf_rows_per_cta(B);
    int g = (B + r - 1) / r;
    return g > SF_MAX_CTAS ? SF_MAX_CTAS : (g < 1 ? 1 : g);
}

}  // namespace

int sampler_fused_supported(int f, int K) {
    return K >= 1 && K <= 32 * SF_MAX_KI && f >= 1 && f <= 32 * SF_MAX_FI && sf_smem_bytes(f, K, true) <= 220 * 1024;
}
// ws: [ticket (4 floats) | KL partials [148] | Qpart [148][K]]
size_t sampler_fused_ws_floats(int K) { return 4 + SF_MAX_CTAS + (size_t)SF_MAX_CTAS * K; }

static SfArgs sf_args(const SamplerParams& P, int B, float* ws) {
    SfArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.K = P.K; a.f = P.f; a.rpc = sf_rows_per_cta(B);
    a.alpha = 1.f / sqrtf((float)P.f); a.prior_c = P.prior_c;
    a.codebook = P.codebook; a.phi = P.shape_logits;
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.partial = ws + 4;
    a.Qpart = ws + 4 + SF_MAX_CTAS;
    return a;
}

int sampler_fused_forward(const SamplerParams& P, int B, const float* xc, int mode, float tau, const float* gumbel,
                          unsigned long long seed, unsigned long long offset, const unsigned long long* d_base, float n_data,
                          float* logits, float* y, float* feats, float* kl_out, float* ws, cudaStream_t stream) {
    ARG_CHECK(sampler_fused_supported(P.f, P.K), "fused sampler: feature_dim %d x num_categories %d does not fit (use the unfused calls)", P.f, P.K);
    ARG_CHECK(B > 0 && xc && logits && y && feats && kl_out && ws && P.codebook && P.shape_logits, "fused sampler: null argument");
    static bool prepared = false;
    if (!prepared) {
        CUDA_TRY(cudaFuncSetAttribute(sampler_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(sampler_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        prepared = true;
    }
    SfArgs a = sf_args(P, B, ws);
    a.mode = mode; a.inv_tau = mode ? 1.f / tau : 1.f; a.n_data = n_data;
    a.xc = xc; a.gumbel = gumbel; a.seed = seed; a.offset = offset; a.d_base = d_base;
    a.logits = logits; a.y = y; a.feats = feats; a.kl_out = kl_out;
    ProfScope ps(PROF_ROWWISE, stream);
    CUDA_TRY(cudaMemsetAsync(a.ticket, 0, 16, stream));
    sampler_fused_fwd_kernel<<<sf_grid(B), SF_THREADS, sf_smem_bytes(P.f, P.K, false), stream>>>(a);
    KERNEL_CHECK();
    return ABCD_OK;
}

// d_xc [B,f] is written; the codebook and posterior_shape_logits gradients are ACCUMULATED into G.  scratch: B*K floats.
int sampler_fused_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* xc, const float* logits,
                           const float* y, int mode, float tau, float n_data, const float* d_feats, const float* g_kl,
                           float* d_xc, float* scratch, float* ws, cudaStream_t stream) {
    ARG_CHECK(sampler_fused_supported(P.f, P.K), "fused sampler: feature_dim %d x num_categories %d does not fit (use the unfused calls)", P.f, P.K);
    ARG_CHECK(B > 0 && xc && logits && y && d_feats && g_kl && d_xc && scratch && ws && G.codebook && G.shape_logits, "fused sampler bwd: null argument");
    SfArgs a = sf_args(P, B, ws);
    a.mode = mode; a.inv_tau = mode ? 1.f / tau : 1.f; a.n_data = n_data;
    a.logits = const_cast<float*>(logits); a.y = const_cast<float*>(y);
    a.d_feats = d_feats; a.g_kl = g_kl; a.dl_scaled = scratch; a.d_xc = d_xc; a.d_phi = G.shape_logits;
    {
        ProfScope ps(PROF_ROWWISE, stream);
        CUDA_TRY(cudaMemsetAsync(a.ticket, 0, 16, stream));
        sampler_fused_bwd_kernel<<<sf_grid(B), SF_THREADS, sf_smem_bytes(P.f, P.K, true), stream>>>(a);
        KERNEL_CHECK();
    }
    GemmArgs g = gemm_args_zero();   // d codebook += xc^T (d logits / sqrt(f))
    g.M = P.f; g.N = P.K; g.K = B;
    g.A = xc; g.lda = P.f; g.a_kmajor = 0;
    g.B = scratch; g.ldb = P.K; g.b_kmajor = 0;
    g.C = G.codebook; g.ldc = P.K; g.accumulate = 1;
    TRY(sgemm_launch(g, stream));
    g = gemm_args_zero();            // d codebook += d feats^T y
    g.M = P.f; g.N = P.K; g.K = B;
    g.A = d_feats; g.lda = P.f; g.a_kmajor = 0;
    g.B = y; g.ldb = P.K; g.b_kmajor = 0;
    g.C = G.codebook; g.ldc = P.K; g.accumulate = 1;
    TRY(sgemm_launch(g, stream));
    return ABCD_OK;
}
