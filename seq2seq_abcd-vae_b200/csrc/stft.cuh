#pragma once
#include "common.cuh"
int stft_logamp(const void* wav, int is_i16, long n_samples, const long long* onset, const int* nsamp, int B,
                const int* offsets, int T_max, int sumT, int n_fft, int hop, int center, const float* window,
                float eps, float normalizer, float* out, long ld_out, cudaStream_t stream);
int gradnorm_clip_sgd(float* params, float* grads, float* momentum_buf, long n, float lr, float momentum,
                      float max_norm, int zero_grads, float* norm_out, float* partial, int n_partial,
                      cudaStream_t stream);
int optim_partial_count();
int philox_fill(float* out, long n, unsigned long long seed, unsigned long long offset, int kind, cudaStream_t stream,
                const unsigned long long* d_base = nullptr);
int counter_add(unsigned long long* d_ctr, unsigned long long delta, cudaStream_t stream);
