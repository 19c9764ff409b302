// Log-amplitude STFT front end written straight into packed (time-major) order.
// Mirrors data_utils.STFT.__call__ (ABCD-VAE/modules/data_utils.py:131-139) followed by the
// log_and_normalize transform (ABCD-VAE/learning.py:466):
//   x.stft(n_fft, hop, window, center=True [reflect pad n_fft/2], onesided) -> |.| -> [T,F] -> log(. + eps)/N
// with T = 1 + n/hop frames per segment and F = n_fft/2 + 1 bins.  Samples are used at raw int16 scale
// (data_utils.py:94, no /32768).
//
// A CTA of 256 threads takes 16 packed frames at a time: the windowed frames are staged in shared memory and every
// thread evaluates one DFT bin for 4 frames at once (the frame samples are warp-wide broadcasts, the twiddles come from a
// [n_fft][64] cos/sin matrix in shared memory whose rows are conflict-free across the 64 bins; frame lengths above 128
// fall back to indexing the n_fft-entry table).  The Nyquist bin is an alternating sum.  Direct O(n_fft * F) evaluation:
// at 8 ms frames that is 16 k FMA per frame, the same order as a radix-2 FFT plus its shuffles at this size.
#include "common.cuh"
#include "stft.cuh"

namespace {

struct StftK {
    const void* wav; int is_i16; long n_samples;
    const long long* onset; const int* nsamp;   // [B] in batch (sorted) order
    const int* offsets; int T_max; int sumT; int B;
    int n_fft, hop, center, F;
    const float* window;
    float eps, normalizer;
    float* out; long ld_out;
};

constexpr int ST_FR = 16;        // frames per CTA iteration
constexpr int ST_THREADS = 256;
constexpr int ST_KB = 64;        // bins per pass
constexpr int ST_OFFS = 2048;    // packed time-step offsets cached in shared memory

template <bool MATRIX>
__global__ void __launch_bounds__(ST_THREADS) stft_logamp_kernel(StftK p) {
    extern __shared__ __align__(16) float sm[];
    const int N = p.n_fft, NP = N + 4;
    float* cs = sm;                       // cos table [N]
    float* sn = cs + N;                   // sin table [N]
    float* fr = sn + N;                   // windowed frames [ST_FR][NP]
    float* Wc = fr + ST_FR * NP;          // MATRIX: cos [N][ST_KB], then sin [N][ST_KB]
    float* Ws = Wc + N * ST_KB;
    int* offs = reinterpret_cast<int*>(MATRIX ? Ws + N * ST_KB : Wc);   // packed offsets [min(T_max, ST_OFFS)] for the frame lookup
    const int n_offs = min(p.T_max, ST_OFFS);
    __shared__ int f_t[ST_FR], f_n[ST_FR];
    __shared__ long f_on[ST_FR];
    const int tid = threadIdx.x;
    for (int j = tid; j < n_offs; j += ST_THREADS) offs[j] = p.offsets[j];
    for (int j = tid; j < N; j += ST_THREADS) {
        float s, c;
        sincospif(2.0f * (float)j / (float)N, &s, &c);
        cs[j] = c; sn[j] = s;
    }
    __syncthreads();
    if (MATRIX) {
        for (int idx = tid; idx < N * ST_KB; idx += ST_THREADS) {
            const int j = idx / ST_KB, k = idx % ST_KB;
            const int w = (j * k) % N;
            Wc[idx] = cs[w]; Ws[idx] = sn[w];
        }
    }
    const short* w16 = reinterpret_cast<const short*>(p.wav);
    const float* w32 = reinterpret_cast<const float*>(p.wav);
    const int kl = tid & (ST_KB - 1), fg = tid / ST_KB;          // bin within a pass, frame group (4 frames each)
    const int nyq = (N % 2 == 0) ? N / 2 : -1;                     // bin with purely real alternating twiddle
    for (int f0 = blockIdx.x * ST_FR; f0 < p.sumT; f0 += gridDim.x * ST_FR) {
        __syncthreads();
        if (tid < ST_FR) {
            const int fi = f0 + tid;
            int t = 0, n = 0; long on = 0;
            if (fi < p.sumT) {
                // packed index -> (t, b): largest t with offsets[t] <= fi
                int lo = 0, hi = p.T_max;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    const int o = mid < n_offs ? offs[mid] : p.offsets[mid];
                    if (o <= fi) lo = mid; else hi = mid;
                }
                t = lo;
                const int b = fi - (lo < n_offs ? offs[lo] : p.offsets[lo]);
                on = p.onset[b]; n = p.nsamp[b];
            }
            f_t[tid] = t; f_n[tid] = n; f_on[tid] = on;
        }
        __syncthreads();
        for (int idx = tid; idx < ST_FR * N; idx += ST_THREADS) {
            const int f = idx / N, j = idx % N;
            const int n = f_n[f];
            int i = f_t[f] * p.hop + j - (p.center ? N / 2 : 0);
            if (i < 0) i = -i;
            if (i >= n) i = 2 * (n - 1) - i;
            float v = 0.f;
            if (i >= 0 && i < n) {
                const long gi = f_on[f] + i;
                if (gi >= 0 && gi < p.n_samples) v = p.is_i16 ? (float)w16[gi] : w32[gi];
            }
            fr[f * NP + j] = v * p.window[j];
        }
        __syncthreads();
        const float* fb = fr + fg * 4 * NP;
        for (int k0 = 0; k0 < p.F; k0 += ST_KB) {
            const int k = k0 + kl;
            if (k >= p.F) continue;
            float re[4] = {0.f, 0.f, 0.f, 0.f}, im[4] = {0.f, 0.f, 0.f, 0.f};
            if (k == nyq) {
                for (int j = 0; j < N; ++j) {
                    const float sg = (j & 1) ? -1.f : 1.f;
#pragma unroll
                    for (int f = 0; f < 4; ++f) re[f] = fmaf(sg, fb[f * NP + j], re[f]);
                }
            } else if (MATRIX && k0 == 0) {
                for (int j = 0; j < N; j += 4) {
                    float4 v[4];
#pragma unroll
                    for (int f = 0; f < 4; ++f) v[f] = *reinterpret_cast<const float4*>(fb + f * NP + j);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float c = Wc[(j + jj) * ST_KB + kl], s = Ws[(j + jj) * ST_KB + kl];
#pragma unroll
                        for (int f = 0; f < 4; ++f) {
                            const float x = jj == 0 ? v[f].x : jj == 1 ? v[f].y : jj == 2 ? v[f].z : v[f].w;
                            re[f] = fmaf(x, c, re[f]);
                            im[f] = fmaf(-x, s, im[f]);
                        }
                    }
                }
            } else {
                int w = 0;
                for (int j = 0; j < N; ++j) {
                    const float c = cs[w], s = sn[w];
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const float x = fb[f * NP + j];
                        re[f] = fmaf(x, c, re[f]);
                        im[f] = fmaf(-x, s, im[f]);
                    }
                    w += k; if (w >= N) w -= N;
                }
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int fi = f0 + fg * 4 + f;
                if (fi >= p.sumT) continue;
                const float mag = sqrtf(re[f] * re[f] + im[f] * im[f]);
                // normalizer == 0 selects the plain amplitude (data_utils.STFT without the log transform)
                p.out[(long)fi * p.ld_out + k] = p.normalizer != 0.f ? logf(mag + p.eps) / p.normalizer : mag;
            }
        }
    }
}

}  // namespace

int stft_logamp(const void* wav, int is_i16, long n_samples, const long long* onset, const int* nsamp, int B,
                const int* offsets, int T_max, int sumT, int n_fft, int hop, int center, const float* window,
                float eps, float normalizer, float* out, long ld_out, cudaStream_t stream) {
    ARG_CHECK(n_fft >= 2 && n_fft <= 2048 && hop >= 1, "stft: n_fft %d / hop %d unsupported", n_fft, hop);
    ARG_CHECK(wav && onset && nsamp && offsets && window && out, "stft: null argument");
    if (sumT <= 0) return ABCD_OK;
    ProfScope ps(PROF_STFT, stream);
    StftK k;
    k.wav = wav; k.is_i16 = is_i16; k.n_samples = n_samples; k.onset = onset; k.nsamp = nsamp;
    k.offsets = offsets; k.T_max = T_max; k.sumT = sumT; k.B = B;
    k.n_fft = n_fft; k.hop = hop; k.center = center; k.F = n_fft / 2 + 1;
    k.window = window; k.eps = eps; k.normalizer = normalizer; k.out = out; k.ld_out = ld_out;
    const int groups = (sumT + ST_FR - 1) / ST_FR;
    const bool matrix = n_fft <= 128 && n_fft % 4 == 0;
    size_t smem = (size_t)(2 * n_fft + ST_FR * (n_fft + 4) + (matrix ? 2 * n_fft * ST_KB : 0) + (T_max < ST_OFFS ? T_max : ST_OFFS)) * sizeof(float);
    if (matrix) {
        static bool configured = false;
        if (!configured) {
            CUDA_TRY(cudaFuncSetAttribute(stft_logamp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = true;
        }
        const int blocks = groups < 148 * 2 ? groups : 148 * 2;
        stft_logamp_kernel<true><<<blocks, ST_THREADS, smem, stream>>>(k);
    } else {
        static bool configured = false;
        if (!configured) {
            CUDA_TRY(cudaFuncSetAttribute(stft_logamp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured = true;
        }
        ARG_CHECK(smem <= 200 * 1024, "stft: n_fft %d needs too much shared memory", n_fft);
        const int blocks = groups < 148 * 4 ? groups : 148 * 4;
        stft_logamp_kernel<false><<<blocks, ST_THREADS, smem, stream>>>(k);
    }
    KERNEL_CHECK();
    return ABCD_OK;
}
