// ABCD sampler (ABCD-VAE/modules/model.py:538-673): to_code_like MLP -> scaled dot-product logits against the
// K-entry codebook -> softmax / Gumbel-softmax -> codebook-weighted feature, plus the Dirichlet-prior KL and
// the closed-form backward of all three.  Also the isotropic-Gaussian latent of the plain/ package
// (plain/modules/model.py:538-567) and generic MLP forward/backward used by both.
//
// Contractions go through sgemm.cu; the row-wise softmax/Gumbel/KL work is one warp per segment.
#include "sgemm.cuh"
#include "sampler.cuh"

namespace {

// ---- special functions in double (K values per launch; accuracy matters more than speed) ----------
__device__ double digamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
__device__ double trigamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    double f = 1.0 / (x * x);
    return r + 1.0 / x + 0.5 * f + f / x * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30))));
}

// Dirichlet terms of kl_divergence (model.py:615-626). One block.
//   s = softmax(phi), alpha = s*N + c, A = sum alpha, E_k = psi(alpha_k) - psi(A)
//   prior = lgamma(A) - sum lgamma(alpha) + sum (alpha-1)E - lgamma(cK) + K lgamma(c) - (c-1) sum E
__global__ void dirichlet_kernel(int K, const float* __restrict__ phi, float prior_c, float n_data,
                                 float* __restrict__ E, float* __restrict__ s_out, double* __restrict__ prior_out) {
    __shared__ double red[40];
    const int tid = threadIdx.x;
    double mx = -1e300;
    for (int k = tid; k < K; k += blockDim.x) mx = fmax(mx, (double)phi[k]);
    // block max via sum trick is wrong; do an explicit shared reduce
    __shared__ double smx[256];
    smx[tid] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) { if (tid < o) smx[tid] = fmax(smx[tid], smx[tid + o]); __syncthreads(); }
    mx = smx[0];
    double z = 0.0;
    for (int k = tid; k < K; k += blockDim.x) z += exp((double)phi[k] - mx);
    z = block_sum_d(z, red);
    double asum = 0.0;
    for (int k = tid; k < K; k += blockDim.x) {
        // follow the reference's fp32 rounding points for s and alpha, then evaluate special functions in double
        float sk = (float)(exp((double)phi[k] - mx) / z);
        float ak = sk * n_data + prior_c;
        s_out[k] = sk;
        asum += (double)ak;
    }
    asum = block_sum_d(asum, red);
    const double psiA = digamma_d(asum);
    double t_lg = 0.0, t_am1E = 0.0, t_E = 0.0;
    for (int k = tid; k < K; k += blockDim.x) {
        double ak = (double)(s_out[k] * n_data + prior_c);
        double e = digamma_d(ak) - psiA;
        E[k] = (float)e;
        t_lg += lgamma(ak);
        t_am1E += (ak - 1.0) * e;
        t_E += e;
    }
    t_lg = block_sum_d(t_lg, red);
    t_am1E = block_sum_d(t_am1E, red);
    t_E = block_sum_d(t_E, red);
    if (tid == 0) {
        double c = (double)prior_c;
        double q_pi = lgamma(asum) - t_lg + t_am1E;
        double p_pi = lgamma(c * K) - lgamma(c) * K + (c - 1.0) * t_E;
        *prior_out = q_pi - p_pi;
    }
}

// One warp per row: q = softmax(logits), row term sum_k q (log q - E_k).
__global__ void kl_rows_kernel(int B, int K, const float* __restrict__ logits, const float* __restrict__ E,
                               float* __restrict__ partial) {
    __shared__ float red[40];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float acc = 0.f;
    for (int b = blockIdx.x * nw + w; b < B; b += gridDim.x * nw) {
        const float* l = logits + (long)b * K;
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, l[k]);
        mx = warp_max(mx);
        float z = 0.f;
        for (int k = lane; k < K; k += 32) z += expf(l[k] - mx);
        z = warp_sum(z);
        float lz = logf(z);
        float s = 0.f;
        for (int k = lane; k < K; k += 32) {
            float lq = l[k] - mx - lz;
            s += expf(lq) * (lq - E[k]);
        }
        s = warp_sum(s);
        if (lane == 0) acc += s;
    }
    float tot = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void kl_final_kernel(int nparts, const float* __restrict__ partial, const double* __restrict__ prior,
                                float batch_over_n, float* __restrict__ kl_out) {
    __shared__ double red[40];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += (double)partial[i];
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) *kl_out = (float)(*prior * (double)batch_over_n + s);
}

// dKL/dlogits_bj = g * q_bj [ (log q_bj - E_j) - sum_k q_bk (log q_bk - E_k) ];  also the column sums Q_k = sum_b q_bk,
// reproducibly: every warp adds its rows (in order) into its own row of shared memory, the block adds the warps in order
// and writes Qpart[block][K]; kl_shape_bwd_kernel adds the blocks in order.  Dynamic shared memory: (warps) x K floats.
__global__ void kl_rows_bwd_kernel(int B, int K, const float* __restrict__ logits, const float* __restrict__ E,
                                   const float* __restrict__ g_kl, float* __restrict__ dlogits, float* __restrict__ Qpart) {
    extern __shared__ float q_warp[];   // [nw][K]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float g = *g_kl;
    float* qw = q_warp + (size_t)w * K;
    for (int k = lane; k < K; k += 32) qw[k] = 0.f;
    for (int b = blockIdx.x * nw + w; b < B; b += gridDim.x * nw) {
        const float* l = logits + (long)b * K;
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, l[k]);
        mx = warp_max(mx);
        float z = 0.f;
        for (int k = lane; k < K; k += 32) z += expf(l[k] - mx);
        z = warp_sum(z);
        float lz = logf(z);
        float s = 0.f;
        for (int k = lane; k < K; k += 32) {
            float lq = l[k] - mx - lz;
            s += expf(lq) * (lq - E[k]);
        }
        s = warp_sum(s);
        for (int k = lane; k < K; k += 32) {
            float lq = l[k] - mx - lz;
            float q = expf(lq);
            dlogits[(long)b * K + k] = g * q * ((lq - E[k]) - s);
            qw[k] += q;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float tot = 0.f;
        for (int i = 0; i < nw; ++i) tot += q_warp[(size_t)i * K + k];
        Qpart[(size_t)blockIdx.x * K + k] = tot;
    }
}

// dKL/dphi_j = g * N s_j (h_j - sum_k s_k h_k),  h_k = psi'(alpha_k) (B s_k - Q_k)   [SURVEY.md 8(a) a10]
__global__ void kl_shape_bwd_kernel(int K, int B, const float* __restrict__ s, const float* __restrict__ Qpart, int nparts,
                                    float* __restrict__ Q, float prior_c, float n_data, const float* __restrict__ g_kl,
                                    float* __restrict__ d_phi) {
    __shared__ double red[40];
    const int tid = threadIdx.x;
    for (int k = tid; k < K; k += blockDim.x) {   // block partials in a fixed order
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;   // four loads in flight; the order is still fixed
        int i = 0;
        for (; i + 3 < nparts; i += 4) {
            t0 += Qpart[(size_t)i * K + k];
            t1 += Qpart[(size_t)(i + 1) * K + k];
            t2 += Qpart[(size_t)(i + 2) * K + k];
            t3 += Qpart[(size_t)(i + 3) * K + k];
        }
        for (; i < nparts; ++i) t0 += Qpart[(size_t)i * K + k];
        Q[k] = (t0 + t1) + (t2 + t3);
    }
    __syncthreads();
    double dot = 0.0;
    for (int k = tid; k < K; k += blockDim.x) {
        double ak = (double)(s[k] * n_data + prior_c);
        double h = trigamma_d(ak) * ((double)B * s[k] - (double)Q[k]);
        dot += (double)s[k] * h;
    }
    dot = block_sum_d(dot, red);
    const double g = (double)*g_kl;
    for (int k = tid; k < K; k += blockDim.x) {
        double ak = (double)(s[k] * n_data + prior_c);
        double h = trigamma_d(ak) * ((double)B * s[k] - (double)Q[k]);
        d_phi[k] += (float)(g * (double)n_data * s[k] * (h - dot));
    }
}

// y = softmax(logits) (mode 0) or softmax((logits + G)/tau) (mode 1, G = Gumbel noise); model.py:601-604
__global__ void sample_rows_kernel(int B, int K, const float* __restrict__ logits, int mode, float inv_tau,
                                   const float* __restrict__ noise, float* __restrict__ y) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x * nw + w; b < B; b += gridDim.x * nw) {
        const float* l = logits + (long)b * K;
        const float* gn = noise ? noise + (long)b * K : nullptr;
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) {
            float v = mode ? (l[k] + gn[k]) * inv_tau : l[k];
            mx = fmaxf(mx, v);
        }
        mx = warp_max(mx);
        float z = 0.f;
        for (int k = lane; k < K; k += 32) {
            float v = mode ? (l[k] + gn[k]) * inv_tau : l[k];
            z += expf(v - mx);
        }
        z = warp_sum(z);
        float iz = 1.f / z;
        for (int k = lane; k < K; k += 32) {
            float v = mode ? (l[k] + gn[k]) * inv_tau : l[k];
            y[(long)b * K + k] = expf(v - mx) * iz;
        }
    }
}

// dlogits = inv_tau * y (dy - sum_k y_k dy_k)
__global__ void sample_rows_bwd_kernel(int B, int K, const float* __restrict__ y, const float* __restrict__ dy,
                                       float inv_tau, float* __restrict__ dlogits) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x * nw + w; b < B; b += gridDim.x * nw) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += y[(long)b * K + k] * dy[(long)b * K + k];
        s = warp_sum(s);
        for (int k = lane; k < K; k += 32)
            dlogits[(long)b * K + k] = inv_tau * y[(long)b * K + k] * (dy[(long)b * K + k] - s);
    }
}

// probs = softmax(logits), category = first index of the row maximum (encode.py:32)
__global__ void softmax_argmax_kernel(int B, int K, const float* __restrict__ logits, float* __restrict__ probs,
                                      int* __restrict__ argmax) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x * nw + w; b < B; b += gridDim.x * nw) {
        const float* l = logits + (long)b * K;
        float mx = -INFINITY; int mi = 0x7fffffff;
        for (int k = lane; k < K; k += 32) { float v = l[k]; if (v > mx) { mx = v; mi = k; } }
        for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(0xffffffffu, mx, o);
            int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
        }
        float z = 0.f;
        for (int k = lane; k < K; k += 32) z += expf(l[k] - mx);
        z = warp_sum(z);
        float iz = 1.f / z;
        if (probs) for (int k = lane; k < K; k += 32) probs[(long)b * K + k] = expf(l[k] - mx) * iz;
        if (argmax && lane == 0) argmax[b] = mi;
    }
}

// Per-batch perplexities logged by Learner.train (learning.py:171-178). One block.
//   out[0] = exp(mean_b H(q_b)), out[1] = exp(H(normalised mean_b q_b)), out[2] = exp(H(softmax(phi)))
__global__ void stats_kernel(int B, int K, const float* __restrict__ logits, const float* __restrict__ phi,
                             float* __restrict__ colmean /*[K] scratch*/, float* __restrict__ out) {
    __shared__ float red[40];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < K; k += blockDim.x) colmean[k] = 0.f;
    __syncthreads();
    float ent = 0.f;
    for (int b = w; b < B; b += nw) {
        const float* l = logits + (long)b * K;
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, l[k]);
        mx = warp_max(mx);
        float z = 0.f;
        for (int k = lane; k < K; k += 32) z += expf(l[k] - mx);
        z = warp_sum(z);
        float lz = logf(z), e = 0.f;
        for (int k = lane; k < K; k += 32) {
            float lq = l[k] - mx - lz, q = expf(lq);
            e -= q * lq;
            atomicAdd(&colmean[k], q);   // log-line statistic only: the last bits may differ between runs
        }
        e = warp_sum(e);
        if (lane == 0) ent += e;
    }
    float tot_ent = block_sum(ent, red);
    __syncthreads();
    float tot = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) tot += colmean[k];
    tot = block_sum(tot, red);
    float e2 = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { float pm = colmean[k] / tot; if (pm > 0.f) e2 -= pm * logf(pm); }
    e2 = block_sum(e2, red);
    // shape perplexity
    float mx = -INFINITY;
    for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, phi[k]);
    __shared__ float smx[256];
    smx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) { if (threadIdx.x < o) smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]); __syncthreads(); }
    mx = smx[0];
    float z = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) z += expf(phi[k] - mx);
    z = block_sum(z, red);
    float lz = logf(z), e3 = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { float lq = phi[k] - mx - lz; e3 -= expf(lq) * lq; }
    e3 = block_sum(e3, red);
    if (threadIdx.x == 0) { out[0] = expf(tot_ent / B); out[1] = expf(e2); out[2] = expf(e3); }
}

__global__ void tanh_bwd_kernel(long n, float* __restrict__ d, const float* __restrict__ hid) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float h = hid[i];
        d[i] *= (1.f - h * h);
    }
}

// Gaussian latent (plain/): z = mean + exp(lv/2) eps ; KL = -1/2 sum (1 + lv - mean^2 - exp(lv))
__global__ void gauss_fwd_kernel(long n, const float* __restrict__ mean, const float* __restrict__ lv,
                                 const float* __restrict__ eps, float* __restrict__ z, float* __restrict__ partial) {
    __shared__ float red[40];
    float kl = 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float m = mean[i], l = lv[i];
        if (z) z[i] = m + expf(0.5f * l) * eps[i];
        kl += -0.5f * (1.f + l - m * m - expf(l));
    }
    float tot = block_sum(kl, red);
    if (threadIdx.x == 0 && partial) partial[blockIdx.x] = tot;
}
__global__ void gauss_bwd_kernel(long n, const float* __restrict__ mean, const float* __restrict__ lv,
                                 const float* __restrict__ eps, const float* __restrict__ dz,
                                 const float* __restrict__ g_kl, float* __restrict__ dmean, float* __restrict__ dlv) {
    const float g = g_kl ? *g_kl : 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float m = mean[i], l = lv[i];
        float d = dz ? dz[i] : 0.f;
        dmean[i] = d + g * m;
        dlv[i] = d * 0.5f * expf(0.5f * l) * eps[i] + g * 0.5f * (expf(l) - 1.f);
    }
}
__global__ void sum_partials_kernel(int n, const float* __restrict__ partial, float* __restrict__ out) {
    __shared__ double red[40];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) *out = (float)s;
}

constexpr int ROW_BLOCKS = 148;

}  // namespace

// ------------------------------------------------------------------------------------------------
// MLP: Linear -> Tanh -> Linear (model.py:316-334)
// ------------------------------------------------------------------------------------------------
int mlp_forward(const MlpParams& P, int rows, const float* x, float* hid, float* y, cudaStream_t stream) {
    GemmArgs g = gemm_args_zero();
    g.M = rows; g.N = P.hidden; g.K = P.in;
    g.A = x; g.lda = P.in; g.a_kmajor = 1;
    g.B = P.w1; g.ldb = P.in; g.b_kmajor = 1;
    g.C = hid; g.ldc = P.hidden; g.bias = P.b1; g.act = 1;
    TRY(sgemm_launch(g, stream));
    g = gemm_args_zero();
    g.M = rows; g.N = P.out; g.K = P.hidden;
    g.A = hid; g.lda = P.hidden; g.a_kmajor = 1;
    g.B = P.w2; g.ldb = P.hidden; g.b_kmajor = 1;
    g.C = y; g.ldc = P.out; g.bias = P.b2;
    TRY(sgemm_launch(g, stream));
    return ABCD_OK;
}

// dy [rows,out] -> parameter grads (accumulated) and dx [rows,in] (overwritten, may be null).
// scratch: rows*hidden floats.
int mlp_backward(const MlpParams& P, const MlpGrads& G, int rows, const float* x, const float* hid, const float* dy,
                 float* dx, float* scratch, cudaStream_t stream) {
    GemmArgs g = gemm_args_zero();   // dW2 += dy^T hid
    g.M = P.out; g.N = P.hidden; g.K = rows;
    g.A = dy; g.lda = P.out; g.a_kmajor = 0;
    g.B = hid; g.ldb = P.hidden; g.b_kmajor = 0;
    g.C = G.w2; g.ldc = P.hidden; g.accumulate = 1;
    TRY(sgemm_launch(g, stream));
    TRY(colsum_launch(dy, P.out, rows, P.out, 1.f, G.b2, 1, gate_map_none(), stream));
    g = gemm_args_zero();            // dhid = dy W2
    g.M = rows; g.N = P.hidden; g.K = P.out;
    g.A = dy; g.lda = P.out; g.a_kmajor = 1;
    g.B = P.w2; g.ldb = P.hidden; g.b_kmajor = 0;
    g.C = scratch; g.ldc = P.hidden;
    TRY(sgemm_launch(g, stream));
    tanh_bwd_kernel<<<ROW_BLOCKS * 2, 256, 0, stream>>>((long)rows * P.hidden, scratch, hid);
    KERNEL_CHECK();
    g = gemm_args_zero();            // dW1 += dpre^T x
    g.M = P.hidden; g.N = P.in; g.K = rows;
    g.A = scratch; g.lda = P.hidden; g.a_kmajor = 0;
    g.B = x; g.ldb = P.in; g.b_kmajor = 0;
    g.C = G.w1; g.ldc = P.in; g.accumulate = 1;
    TRY(sgemm_launch(g, stream));
    TRY(colsum_launch(scratch, P.hidden, rows, P.hidden, 1.f, G.b1, 1, gate_map_none(), stream));
    if (dx) {
        g = gemm_args_zero();        // dx = dpre W1
        g.M = rows; g.N = P.in; g.K = P.hidden;
        g.A = scratch; g.lda = P.hidden; g.a_kmajor = 1;
        g.B = P.w1; g.ldb = P.in; g.b_kmajor = 0;
        g.C = dx; g.ldc = P.in;
        TRY(sgemm_launch(g, stream));
    }
    return ABCD_OK;
}

// ------------------------------------------------------------------------------------------------
// ABCDSampler.forward: logits = to_code_like(x) @ codebook / sqrt(f)   (model.py:588-589)
// ------------------------------------------------------------------------------------------------
int sampler_logits_forward(const SamplerParams& P, int B, const float* x, float* hid, float* xc, float* logits,
                           cudaStream_t stream) {
    TRY(mlp_forward(P.mlp, B, x, hid, xc, stream));
    GemmArgs g = gemm_args_zero();
    g.M = B; g.N = P.K; g.K = P.f;
    g.A = xc; g.lda = P.f; g.a_kmajor = 1;
    g.B = P.codebook; g.ldb = P.K; g.b_kmajor = 0;
    g.C = logits; g.ldc = P.K; g.alpha = 1.f / sqrtf((float)P.f);
    TRY(sgemm_launch(g, stream));
    return ABCD_OK;
}

// scratch: B*max(hidden, f) + B*f floats
int sampler_logits_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* x, const float* hid,
                            const float* xc, const float* dlogits, float* dx, float* scratch, cudaStream_t stream) {
    const float alpha = 1.f / sqrtf((float)P.f);
    float* dxc = scratch;                       // [B,f]
    float* mlp_scratch = scratch + (size_t)B * P.f;
    GemmArgs g = gemm_args_zero();   // d codebook += alpha * xc^T dlogits
    g.M = P.f; g.N = P.K; g.K = B;
    g.A = xc; g.lda = P.f; g.a_kmajor = 0;
    g.B = dlogits; g.ldb = P.K; g.b_kmajor = 0;
    g.C = G.codebook; g.ldc = P.K; g.accumulate = 1; g.alpha = alpha;
    TRY(sgemm_launch(g, stream));
    g = gemm_args_zero();            // dxc = alpha * dlogits codebook^T
    g.M = B; g.N = P.f; g.K = P.K;
    g.A = dlogits; g.lda = P.K; g.a_kmajor = 1;
    g.B = P.codebook; g.ldb = P.K; g.b_kmajor = 1;
    g.C = dxc; g.ldc = P.f; g.alpha = alpha;
    TRY(sgemm_launch(g, stream));
    TRY(mlp_backward(P.mlp, G.mlp, B, x, hid, dxc, dx, mlp_scratch, stream));
    return ABCD_OK;
}

// ABCDSampler.sample (model.py:592-606)
int sampler_sample_forward(const SamplerParams& P, int B, const float* logits, int mode, float tau, const float* noise,
                           float* y, float* features, cudaStream_t stream) {
    ARG_CHECK(mode == 0 || noise, "sampler: Gumbel mode needs a noise buffer");
    sample_rows_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(B, P.K, logits, mode, mode ? 1.f / tau : 1.f, noise, y);
    KERNEL_CHECK();
    GemmArgs g = gemm_args_zero();   // features = y codebook^T
    g.M = B; g.N = P.f; g.K = P.K;
    g.A = y; g.lda = P.K; g.a_kmajor = 1;
    g.B = P.codebook; g.ldb = P.K; g.b_kmajor = 1;
    g.C = features; g.ldc = P.f;
    TRY(sgemm_launch(g, stream));
    return ABCD_OK;
}

// scratch: B*K floats
int sampler_sample_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* y, int mode, float tau,
                            const float* dfeat, float* dlogits, float* scratch, cudaStream_t stream) {
    float* dy = scratch;
    GemmArgs g = gemm_args_zero();   // d codebook += dfeat^T y
    g.M = P.f; g.N = P.K; g.K = B;
    g.A = dfeat; g.lda = P.f; g.a_kmajor = 0;
    g.B = y; g.ldb = P.K; g.b_kmajor = 0;
    g.C = G.codebook; g.ldc = P.K; g.accumulate = 1;
    TRY(sgemm_launch(g, stream));
    g = gemm_args_zero();            // dy = dfeat codebook
    g.M = B; g.N = P.K; g.K = P.f;
    g.A = dfeat; g.lda = P.f; g.a_kmajor = 1;
    g.B = P.codebook; g.ldb = P.K; g.b_kmajor = 0;
    g.C = dy; g.ldc = P.K;
    TRY(sgemm_launch(g, stream));
    sample_rows_bwd_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(B, P.K, y, dy, mode ? 1.f / tau : 1.f, dlogits);
    KERNEL_CHECK();
    return ABCD_OK;
}

// ABCDSampler.kl_divergence (model.py:608-639).  ws: 1 double, E[K], s[K], Q[K], partial[ROW_BLOCKS], Qpart[ROW_BLOCKS][K].
size_t sampler_kl_ws_floats(int K) { return (size_t)3 * K + ROW_BLOCKS + 8 + (size_t)ROW_BLOCKS * K; }

static void kl_ws(float* ws, int K, float** E, float** s, float** Q, float** partial, double** prior) {
    *prior = reinterpret_cast<double*>(ws);   // first 8 bytes (ws is 256B aligned by the caller)
    *E = ws + 4; *s = *E + K; *Q = *s + K; *partial = *Q + K;
}

int sampler_kl_forward(const SamplerParams& P, int B, const float* logits, float n_data, float* kl_out, float* ws,
                       cudaStream_t stream) {
    ARG_CHECK(P.K <= 4096, "sampler: too many categories");
    float *E, *s, *Q, *partial; double* prior;
    kl_ws(ws, P.K, &E, &s, &Q, &partial, &prior);
    dirichlet_kernel<<<1, 256, 0, stream>>>(P.K, P.shape_logits, P.prior_c, n_data, E, s, prior);
    KERNEL_CHECK();
    kl_rows_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(B, P.K, logits, E, partial);
    KERNEL_CHECK();
    kl_final_kernel<<<1, 256, 0, stream>>>(ROW_BLOCKS, partial, prior, (float)B / n_data, kl_out);
    KERNEL_CHECK();
    return ABCD_OK;
}

int sampler_kl_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* logits, float n_data,
                        const float* g_kl, float* dlogits, float* ws, cudaStream_t stream) {
    float *E, *s, *Q, *partial; double* prior;
    kl_ws(ws, P.K, &E, &s, &Q, &partial, &prior);
    ARG_CHECK(P.K <= 4096, "sampler: too many categories");
    float* Qpart = partial + ROW_BLOCKS;
    int nw = 12288 / P.K;                         // warps per block: (warps) x K floats of shared memory <= 48 KB
    nw = nw > 8 ? 8 : (nw < 1 ? 1 : nw);
    kl_rows_bwd_kernel<<<ROW_BLOCKS, 32 * nw, (size_t)nw * P.K * sizeof(float), stream>>>(B, P.K, logits, E, g_kl, dlogits, Qpart);
    KERNEL_CHECK();
    kl_shape_bwd_kernel<<<1, 256, 0, stream>>>(P.K, B, s, Qpart, ROW_BLOCKS, Q, P.prior_c, n_data, g_kl, G.shape_logits);
    KERNEL_CHECK();
    return ABCD_OK;
}

int sampler_stats(int B, int K, const float* logits, const float* shape_logits, float* scratch_K, float* out3,
                  cudaStream_t stream) {
    stats_kernel<<<1, 256, 0, stream>>>(B, K, logits, shape_logits, scratch_K, out3);
    KERNEL_CHECK();
    return ABCD_OK;
}

int softmax_argmax(int B, int K, const float* logits, float* probs, int* argmax, cudaStream_t stream) {
    softmax_argmax_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(B, K, logits, probs, argmax);
    KERNEL_CHECK();
    return ABCD_OK;
}

int gauss_latent_forward(long n, const float* mean, const float* lv, const float* eps, float* z, float* kl_out,
                         float* partial /*ROW_BLOCKS floats*/, cudaStream_t stream) {
    gauss_fwd_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(n, mean, lv, eps, z, partial);
    KERNEL_CHECK();
    if (kl_out) {
        sum_partials_kernel<<<1, 256, 0, stream>>>(ROW_BLOCKS, partial, kl_out);
        KERNEL_CHECK();
    }
    return ABCD_OK;
}
int gauss_latent_backward(long n, const float* mean, const float* lv, const float* eps, const float* dz,
                          const float* g_kl, float* dmean, float* dlv, cudaStream_t stream) {
    gauss_bwd_kernel<<<ROW_BLOCKS, 256, 0, stream>>>(n, mean, lv, eps, dz, g_kl, dmean, dlv);
    KERNEL_CHECK();
    return ABCD_OK;
}
