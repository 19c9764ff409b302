// Fused global-norm gradient clipping + SGD step over one flat parameter/gradient arena, and Philox noise.
//   torch.nn.utils.clip_grad_norm_(params, max_norm)  (ABCD-VAE/learning.py:161)
//        total = ||g||_2 over ALL parameters; g *= min(1, max_norm / (total + 1e-6))
//   torch.optim.SGD(lr, momentum).step()               (ABCD-VAE/learning.py:163,256)
//        buf = momentum*buf + g ; p -= lr*buf   (buf starts at zero, dampening 0, no nesterov/weight decay)
// Two launches: (1) per-CTA partial sums of squares, (2) every CTA re-reduces the few hundred partials in a
// fixed order (deterministic, no atomics), derives the clip coefficient and applies the step.  Algorithmic
// traffic 16 B/parameter (4 read for the norm, 8 read + 4 written for the step), +8 with momentum.
#include "common.cuh"
#include "stft.cuh"
#include "philox.cuh"

namespace {

constexpr int OPT_BLOCKS = 592;   // 4 CTAs per SM
constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
    __shared__ double red[40];
    double s = 0.0;
    const long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float a = 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = (double)a;
    if (blockIdx.x == 0)
        for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += (double)g[i] * (double)g[i];
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = (float)s;
}

__global__ void __launch_bounds__(OPT_THREADS) clip_sgd_kernel(float* __restrict__ p, float* __restrict__ g,
                                                               float* __restrict__ mbuf, long n, float lr, float momentum,
                                                               float max_norm, int zero_grads,
                                                               const float* __restrict__ partial, int n_partial,
                                                               float* __restrict__ norm_out) {
    __shared__ double red[40];
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) s += (double)partial[i];
    s = block_sum_d(s, red);
    const float total = (float)sqrt(s);
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = total;
    float coef = max_norm / (total + 1e-6f);
    coef = coef < 1.f ? coef : 1.f;
    if (max_norm <= 0.f) coef = 1.f;
    const long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(mbuf);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 gv = g4[i], pv = p4[i];
        gv.x *= coef; gv.y *= coef; gv.z *= coef; gv.w *= coef;
        float4 st = gv;
        if (mbuf) {
            float4 mv = m4[i];
            mv.x = momentum * mv.x + gv.x; mv.y = momentum * mv.y + gv.y;
            mv.z = momentum * mv.z + gv.z; mv.w = momentum * mv.w + gv.w;
            m4[i] = mv; st = mv;
        }
        pv.x -= lr * st.x; pv.y -= lr * st.y; pv.z -= lr * st.z; pv.w -= lr * st.w;
        p4[i] = pv;
        g4[i] = zero_grads ? make_float4(0.f, 0.f, 0.f, 0.f) : gv;
    }
    if (blockIdx.x == 0)
        for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float gv = g[i] * coef, st = gv;
            if (mbuf) { float mv = momentum * mbuf[i] + gv; mbuf[i] = mv; st = mv; }
            p[i] -= lr * st;
            g[i] = zero_grads ? 0.f : gv;
        }
}

// kind 0: standard normal (Box-Muller); kind 1: standard Gumbel  -log(-log U)  (= -log of an Exp(1) draw,
// torch.nn.functional.gumbel_softmax's noise, model.py:604)
__global__ void philox_kernel(float* __restrict__ out, long n, uint64_t seed, uint64_t offset, int kind,
                              const unsigned long long* __restrict__ d_base) {
    // d_base (optional): a device-resident stream position added to the launch-time offset, so that a launch recorded in a
    // CUDA graph draws fresh numbers on every replay (the graph advances the counter with counter_add_kernel)
    if (d_base) offset += (uint64_t)__ldg(d_base);
    const long n4 = (n + 3) >> 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4(seed, offset + (uint64_t)i, r);
        float v[4];
        if (kind == 0) {
            float u0 = u01(r[0]), u1 = u01(r[1]), u2 = u01(r[2]), u3 = u01(r[3]);
            float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u2));
            float s0, c0, s1, c1;
            sincospif(2.f * u1, &s0, &c0);
            sincospif(2.f * u3, &s1, &c1);
            v[0] = m0 * c0; v[1] = m0 * s0; v[2] = m1 * c1; v[3] = m1 * s1;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = -logf(-logf(u01(r[j])));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i * 4 + j < n) out[i * 4 + j] = v[j];
    }
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long delta) { *ctr += delta; }

}  // namespace

int counter_add(unsigned long long* d_ctr, unsigned long long delta, cudaStream_t stream) {
    ARG_CHECK(d_ctr, "counter_add: null counter");
    counter_add_kernel<<<1, 1, 0, stream>>>(d_ctr, delta);
    KERNEL_CHECK();
    return ABCD_OK;
}

int optim_partial_count() { return OPT_BLOCKS; }

int gradnorm_clip_sgd(float* params, float* grads, float* momentum_buf, long n, float lr, float momentum,
                      float max_norm, int zero_grads, float* norm_out, float* partial, int n_partial,
                      cudaStream_t stream) {
    ARG_CHECK(params && grads && partial && n_partial >= OPT_BLOCKS, "gradnorm_clip_sgd: bad arguments");
    ARG_CHECK(((uintptr_t)params % 16 == 0) && ((uintptr_t)grads % 16 == 0) && (!momentum_buf || (uintptr_t)momentum_buf % 16 == 0),
              "gradnorm_clip_sgd: arenas must be 16-byte aligned");
    if (n <= 0) return ABCD_OK;
    ProfScope ps(PROF_OPTIM, stream);
    sumsq_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(grads, n, partial);
    KERNEL_CHECK();
    clip_sgd_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(params, grads, momentum_buf, n, lr, momentum, max_norm,
                                                          zero_grads, partial, OPT_BLOCKS, norm_out);
    KERNEL_CHECK();
    return ABCD_OK;
}

int philox_fill(float* out, long n, unsigned long long seed, unsigned long long offset, int kind, cudaStream_t stream,
                const unsigned long long* d_base) {
    if (n <= 0) return ABCD_OK;
    ProfScope ps(PROF_NOISE, stream);
    long n4 = (n + 3) >> 2;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    philox_kernel<<<blocks, 256, 0, stream>>>(out, n, (uint64_t)seed, (uint64_t)offset, kind, d_base);
    KERNEL_CHECK();
    return ABCD_OK;
}
