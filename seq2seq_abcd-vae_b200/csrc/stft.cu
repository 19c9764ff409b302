// Log-amplitude STFT front end written straight into packed (time-major) order.
// Mirrors data_utils.STFT.__call__ (ABCD-VAE/modules/data_utils.py:131-139) followed by the
// log_and_normalize transform (ABCD-VAE/learning.py:466):
//   x.stft(n_fft, hop, window, center=True [reflect pad n_fft/2], onesided) -> |.| -> [T,F] -> log(. + eps)/N
// with T = 1 + n/hop frames per segment and F = n_fft/2 + 1 bins.  Samples are used at raw int16 scale
// (data_utils.py:94, no /32768).
//
// One CTA per packed frame at a time: stage the windowed frame in shared memory, then each thread evaluates
// DFT bins against a shared twiddle table.  (Direct O(n_fft * F) evaluation: at 8 ms frames this is ~16 k FMA
// per frame and the kernel is far off the critical path of a training step; a radix-2 warp FFT is the planned
// replacement for the bulk-encode case.)
#include "common.cuh"
#include "stft.cuh"

namespace {

struct StftK {
    const void* wav; int is_i16; long n_samples;
    const int* onset; const int* nsamp;   // [B] in batch (sorted) order
    const int* offsets; int T_max; int sumT; int B;
    int n_fft, hop, center, F;
    const float* window;
    float eps, normalizer;
    float* out; long ld_out;
};

__global__ void __launch_bounds__(128) stft_logamp_kernel(StftK p) {
    extern __shared__ float sm[];
    float* cs = sm;                 // cos table [n_fft]
    float* sn = cs + p.n_fft;       // sin table [n_fft]
    float* fr = sn + p.n_fft;       // windowed frame [n_fft]
    const int N = p.n_fft;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float s, c;
        sincospif(2.0f * (float)j / (float)N, &s, &c);
        cs[j] = c; sn[j] = s;
    }
    const short* w16 = reinterpret_cast<const short*>(p.wav);
    const float* w32 = reinterpret_cast<const float*>(p.wav);
    for (int fi = blockIdx.x; fi < p.sumT; fi += gridDim.x) {
        // packed index -> (t, b): largest t with offsets[t] <= fi
        int lo = 0, hi = p.T_max;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (p.offsets[mid] <= fi) lo = mid; else hi = mid; }
        const int t = lo, b = fi - p.offsets[lo];
        const long on = p.onset[b];
        const int n = p.nsamp[b];
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            int i = t * p.hop + j - (p.center ? N / 2 : 0);
            if (i < 0) i = -i;
            if (i >= n) i = 2 * (n - 1) - i;
            float v = 0.f;
            if (i >= 0 && i < n) {
                long gi = on + i;
                if (gi >= 0 && gi < p.n_samples) v = p.is_i16 ? (float)w16[gi] : w32[gi];
            }
            fr[j] = v * p.window[j];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < p.F; k += blockDim.x) {
            float re = 0.f, im = 0.f;
            int idx = 0;
            for (int j = 0; j < N; ++j) {
                float v = fr[j];
                re = fmaf(v, cs[idx], re);
                im = fmaf(-v, sn[idx], im);
                idx += k; if (idx >= N) idx -= N;
            }
            float mag = sqrtf(re * re + im * im);
            // normalizer == 0 selects the plain amplitude (data_utils.STFT without the log transform)
            p.out[(long)fi * p.ld_out + k] = p.normalizer != 0.f ? logf(mag + p.eps) / p.normalizer : mag;
        }
    }
}

}  // namespace

int stft_logamp(const void* wav, int is_i16, long n_samples, const int* onset, const int* nsamp, int B,
                const int* offsets, int T_max, int sumT, int n_fft, int hop, int center, const float* window,
                float eps, float normalizer, float* out, long ld_out, cudaStream_t stream) {
    ARG_CHECK(n_fft >= 2 && n_fft <= 4096 && hop >= 1, "stft: n_fft %d / hop %d unsupported", n_fft, hop);
    ARG_CHECK(wav && onset && nsamp && offsets && window && out, "stft: null argument");
    if (sumT <= 0) return ABCD_OK;
    ProfScope ps(PROF_STFT, stream);
    StftK k;
    k.wav = wav; k.is_i16 = is_i16; k.n_samples = n_samples; k.onset = onset; k.nsamp = nsamp;
    k.offsets = offsets; k.T_max = T_max; k.sumT = sumT; k.B = B;
    k.n_fft = n_fft; k.hop = hop; k.center = center; k.F = n_fft / 2 + 1;
    k.window = window; k.eps = eps; k.normalizer = normalizer; k.out = out; k.ld_out = ld_out;
    int blocks = sumT < 148 * 16 ? sumT : 148 * 16;
    size_t smem = (size_t)3 * n_fft * sizeof(float);
    stft_logamp_kernel<<<blocks, 128, smem, stream>>>(k);
    KERNEL_CHECK();
    return ABCD_OK;
}
