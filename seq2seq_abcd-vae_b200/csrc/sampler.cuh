#pragma once
#include "common.cuh"

struct MlpParams { int in, hidden, out; const float* w1; const float* b1; const float* w2; const float* b2; };
struct MlpGrads { float* w1; float* b1; float* w2; float* b2; };

struct SamplerParams {
    MlpParams mlp;            // to_code_like: in -> hidden -> f
    int f, K;
    const float* codebook;    // [f, K]
    const float* shape_logits;   // [K]
    float prior_c;
};
struct SamplerGrads { MlpGrads mlp; float* codebook; float* shape_logits; };

int mlp_forward(const MlpParams& P, int rows, const float* x, float* hid, float* y, cudaStream_t stream);
int mlp_backward(const MlpParams& P, const MlpGrads& G, int rows, const float* x, const float* hid, const float* dy,
                 float* dx, float* scratch, cudaStream_t stream);
int sampler_logits_forward(const SamplerParams& P, int B, const float* x, float* hid, float* xc, float* logits, cudaStream_t stream);
int sampler_logits_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* x, const float* hid,
                            const float* xc, const float* dlogits, float* dx, float* scratch, cudaStream_t stream);
int sampler_sample_forward(const SamplerParams& P, int B, const float* logits, int mode, float tau, const float* noise,
                           float* y, float* features, cudaStream_t stream);
int sampler_sample_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* y, int mode, float tau,
                            const float* dfeat, float* dlogits, float* scratch, cudaStream_t stream);
size_t sampler_kl_ws_floats(int K);
int sampler_kl_forward(const SamplerParams& P, int B, const float* logits, float n_data, float* kl_out, float* ws, cudaStream_t stream);
int sampler_kl_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* logits, float n_data,
                        const float* g_kl, float* dlogits, float* ws, cudaStream_t stream);
int sampler_stats(int B, int K, const float* logits, const float* shape_logits, float* scratch_K, float* out3, cudaStream_t stream);
int softmax_argmax(int B, int K, const float* logits, float* probs, int* argmax, cudaStream_t stream);
int gauss_latent_forward(long n, const float* mean, const float* lv, const float* eps, float* z, float* kl_out, float* partial, cudaStream_t stream);
int gauss_latent_backward(long n, const float* mean, const float* lv, const float* eps, const float* dz, const float* g_kl,
                          float* dmean, float* dlv, cudaStream_t stream);

// fused ABCD sampler (sampler_fused.cu)
int sampler_fused_supported(int f, int K);
size_t sampler_fused_ws_floats(int K);
int sampler_fused_forward(const SamplerParams& P, int B, const float* xc, int mode, float tau, const float* gumbel,
                          unsigned long long seed, unsigned long long offset, const unsigned long long* d_base, float n_data,
                          float* logits, float* y, float* feats, float* kl_out, float* ws, cudaStream_t stream);
int sampler_fused_backward(const SamplerParams& P, const SamplerGrads& G, int B, const float* xc, const float* logits,
                           const float* y, int mode, float tau, float n_data, const float* d_feats, const float* g_kl,
                           float* d_xc, float* scratch, float* ws, cudaStream_t stream);
