// extern "C" surface of libabcd_b200.so -- thin argument translation onto the internal kernels.
#include <stdarg.h>
#include "../../include/abcd_b200.h"
#include "common.cuh"
#include "enc.cuh"
#include "sampler.cuh"
#include "sgemm.cuh"
#include "stft.cuh"

int enc_max_clusters(int which, int* out);
int enc_read_timing(unsigned long long* out5);
int rnn_set_tc(int on);
int dec_read_timing(unsigned long long* out16);
int measure_cluster_exchange(int cluster_size, int iters, int mode, float* d_out2, float* d_scratch, cudaStream_t stream);
int dec_max_clusters(int which, int F, int* out);

static thread_local char g_err[1024] = "";
void abcd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch accounting and per-kernel-group event timing ---------------------------------------------
#include <atomic>
#include <mutex>
#include <vector>
static std::atomic<long long> g_launches{0};
void abcd_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfSpan { cudaEvent_t a, b; };
static std::vector<ProfSpan> g_spans[PROF_COUNT];
static std::vector<cudaEvent_t> g_open[PROF_COUNT];
static const char* const g_prof_names[PROF_COUNT] = {
    "enc_fwd_kernel", "enc_bwd_kernel", "dec_fwd_kernel", "dec_bwd_kernel", "sgemm+colsum", "rowwise", "clip_sgd",
    "stft_logamp", "philox"};

void abcd_prof_begin(int id, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_open[id].push_back(e);
}
void abcd_prof_end(int id, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_open[id].empty()) return;
    cudaEvent_t a = g_open[id].back();
    g_open[id].pop_back();
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
    cudaEventRecord(b, s);
    g_spans[id].push_back(ProfSpan{a, b});
}

static PackedMeta to_meta(const abcd_packed_t* pk) {
    PackedMeta m;
    m.B = pk->B; m.T_max = pk->T_max; m.sumT = pk->sumT;
    m.batch_sizes = pk->batch_sizes; m.offsets = pk->offsets; m.lengths = pk->lengths;
    return m;
}
static MlpParams to_mlp(const abcd_mlp_params_t& p) { return MlpParams{p.in, p.hidden, p.out, p.w1, p.b1, p.w2, p.b2}; }
static MlpGrads to_mlpg(const abcd_mlp_grads_t& g) { return MlpGrads{g.w1, g.b1, g.w2, g.b2}; }
static SamplerParams to_sp(const abcd_sampler_params_t* p) {
    SamplerParams s;
    s.mlp = to_mlp(p->mlp); s.f = p->f; s.K = p->K; s.codebook = p->codebook; s.shape_logits = p->shape_logits; s.prior_c = p->prior_c;
    return s;
}
static SamplerGrads to_sg(const abcd_sampler_grads_t* g) {
    SamplerGrads s;
    s.mlp = to_mlpg(g->mlp); s.codebook = g->codebook; s.shape_logits = g->shape_logits;
    return s;
}
static EncParams to_ep(const abcd_enc_params_t* p) {
    EncParams e;
    e.cell = p->cell; e.F = p->F; e.H = p->H; e.bidir = p->bidir;
    for (int d = 0; d < 2; ++d) { e.w_ih[d] = p->w_ih[d]; e.w_hh[d] = p->w_hh[d]; e.b_ih[d] = p->b_ih[d]; e.b_hh[d] = p->b_hh[d]; }
    return e;
}
static DecParams to_dp(const abcd_dec_params_t* p) {
    DecParams d;
    d.cell = p->cell; d.F = p->F; d.H = p->H; d.M = p->M; d.fin = p->fin;
    d.w_f2h = p->w_f2h; d.b_f2h = p->b_f2h; d.w_ih = p->w_ih; d.w_hh = p->w_hh; d.b_ih = p->b_ih; d.b_hh = p->b_hh;
    for (int i = 0; i < 2; ++i) { d.em_w1[i] = p->em_w1[i]; d.em_b1[i] = p->em_b1[i]; d.em_w2[i] = p->em_w2[i]; d.em_b2[i] = p->em_b2[i]; }
    d.off_w1 = p->off_w1; d.off_b1 = p->off_b1; d.off_w2 = p->off_w2; d.off_b2 = p->off_b2;
    return d;
}
static DecGrads to_dg(const abcd_dec_grads_t* p) {
    DecGrads d;
    d.w_f2h = p->w_f2h; d.b_f2h = p->b_f2h; d.w_ih = p->w_ih; d.w_hh = p->w_hh; d.b_ih = p->b_ih; d.b_hh = p->b_hh;
    for (int i = 0; i < 2; ++i) { d.em_w1[i] = p->em_w1[i]; d.em_b1[i] = p->em_b1[i]; d.em_w2[i] = p->em_w2[i]; d.em_b2[i] = p->em_b2[i]; }
    d.off_w1 = p->off_w1; d.off_b1 = p->off_b1; d.off_w2 = p->off_w2; d.off_b2 = p->off_b2;
    return d;
}
#define S(stream) ((cudaStream_t)(stream))

extern "C" {

const char* abcd_last_error(void) { return g_err; }
const char* abcd_version(void) { return "abcd_b200 0.2 (sm_100a: cluster-persistent fp32 recurrences, tcgen05 3xTF32 GEMMs)"; }

// which: 0 enc_fwd, 1 enc_bwd, 2 dec_fwd, 3 dec_bwd -> number of co-resident clusters (or negative error)
int abcd_max_active_clusters(int which, int F) {
    int n = 0;
    int rc = which < 2 ? enc_max_clusters(which, &n) : dec_max_clusters(which - 2, F, &n);
    return rc == ABCD_OK ? n : rc;
}
long long abcd_launch_count(void) { return g_launches.load(); }
// debug (ABCD_TIMING=1): accumulated clock cycles of enc_fwd CTA 0: setup, gemm, epilogue, cluster barrier, steps
int abcd_debug_enc_timing(unsigned long long* out10) { return enc_read_timing(out10); }
int abcd_set_rnn_tensorcore(int on) { return rnn_set_tc(on); }
int abcd_debug_dec_timing(unsigned long long* out16) { return dec_read_timing(out16); }
int abcd_measure_cluster_exchange(int cluster_size, int iters, int mode, float* out2, float* scratch, void* stream) {
    return measure_cluster_exchange(cluster_size, iters, mode, out2, scratch, (cudaStream_t)stream);
}
int abcd_profile_groups(void) { return PROF_COUNT; }
const char* abcd_profile_name(int id) { return (id >= 0 && id < PROF_COUNT) ? g_prof_names[id] : ""; }
void abcd_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
// Synchronises the recorded events, returns total milliseconds and span count of a group, and clears it.
int abcd_profile_read(int id, float* total_ms, int* count) {
    if (id < 0 || id >= PROF_COUNT) return ABCD_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    float tot = 0.f;
    int n = 0;
    for (auto& sp : g_spans[id]) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { tot += ms; ++n; }
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    g_spans[id].clear();
    if (total_ms) *total_ms = tot;
    if (count) *count = n;
    return ABCD_OK;
}

int abcd_init(int device) {
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        abcd_set_error("abcd_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return ABCD_ERR_UNSUPPORTED;
    }
    TRY(enc_init());
    TRY(dec_init());
    return ABCD_OK;
}

int abcd_stft_logamp(const void* wav, int wav_is_int16, int64_t n_samples, const int64_t* seg_onset,
                     const int32_t* seg_nsamp, const abcd_packed_t* pk, int n_fft, int hop, int center,
                     const float* window, float eps, float normalizer, float* out, int64_t ld_out, void* stream) {
    ARG_CHECK(pk, "stft: null packed descriptor");
    return stft_logamp(wav, wav_is_int16, (long)n_samples, (const long long*)seg_onset, seg_nsamp, pk->B, pk->offsets, pk->T_max, pk->sumT,
                       n_fft, hop, center, window, eps, normalizer, out, (long)ld_out, S(stream));
}

size_t abcd_enc_workspace_bytes(int cell, int F, int B, int sumT, int bidir, int save) {
    return enc_workspace_bytes(cell, F, B, sumT, bidir, save);
}
int abcd_enc_fwd(const abcd_enc_params_t* p, const abcd_packed_t* pk, const float* x, int ldx, float* last_hidden,
                 int save, void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && pk && x && last_hidden && ws, "enc_fwd: null argument");
    return enc_forward(to_ep(p), to_meta(pk), x, ldx, last_hidden, save, ws, ws_bytes, S(stream));
}
int abcd_enc_bwd(const abcd_enc_params_t* p, const abcd_enc_grads_t* g, const abcd_packed_t* pk, const float* x, int ldx,
                 const float* d_last_hidden, void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && g && pk && x && d_last_hidden && ws, "enc_bwd: null argument");
    EncGrads G;
    for (int d = 0; d < 2; ++d) { G.w_ih[d] = g->w_ih[d]; G.w_hh[d] = g->w_hh[d]; G.b_ih[d] = g->b_ih[d]; G.b_hh[d] = g->b_hh[d]; }
    return enc_backward(to_ep(p), G, to_meta(pk), x, ldx, d_last_hidden, ws, ws_bytes, S(stream));
}

int abcd_sampler_logits_fwd(const abcd_sampler_params_t* p, int B, const float* x, float* hid, float* xc, float* logits, void* stream) {
    return sampler_logits_forward(to_sp(p), B, x, hid, xc, logits, S(stream));
}
int abcd_sampler_logits_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* x,
                            const float* hid, const float* xc, const float* d_logits, float* d_x, float* scratch, void* stream) {
    return sampler_logits_backward(to_sp(p), to_sg(g), B, x, hid, xc, d_logits, d_x, scratch, S(stream));
}
int abcd_sampler_sample_fwd(const abcd_sampler_params_t* p, int B, const float* logits, int mode, float tau,
                            const float* noise, float* y, float* features, void* stream) {
    return sampler_sample_forward(to_sp(p), B, logits, mode, tau, noise, y, features, S(stream));
}
int abcd_sampler_sample_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* y, int mode,
                            float tau, const float* d_features, float* d_logits, float* scratch, void* stream) {
    return sampler_sample_backward(to_sp(p), to_sg(g), B, y, mode, tau, d_features, d_logits, scratch, S(stream));
}
size_t abcd_sampler_kl_ws_floats(int K) { return sampler_kl_ws_floats(K); }
int abcd_sampler_kl_fwd(const abcd_sampler_params_t* p, int B, const float* logits, float n_data, float* kl_out, float* ws, void* stream) {
    return sampler_kl_forward(to_sp(p), B, logits, n_data, kl_out, ws, S(stream));
}
int abcd_sampler_kl_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* logits,
                        float n_data, const float* g_kl, float* d_logits, float* ws, void* stream) {
    return sampler_kl_backward(to_sp(p), to_sg(g), B, logits, n_data, g_kl, d_logits, ws, S(stream));
}
int abcd_sampler_fused_supported(int f, int K) { return sampler_fused_supported(f, K); }
size_t abcd_sampler_fused_ws_floats(int K) { return sampler_fused_ws_floats(K); }
int abcd_sampler_fused_fwd(const abcd_sampler_params_t* p, int B, const float* xc, int mode, float tau, const float* gumbel_noise,
                           uint64_t seed, uint64_t offset, const uint64_t* d_base, float entire_data_size,
                           float* logits, float* y, float* features, float* kl_out, float* ws, void* stream) {
    ARG_CHECK(p, "sampler_fused_fwd: null argument");
    return sampler_fused_forward(to_sp(p), B, xc, mode, tau, gumbel_noise, seed, offset, (const unsigned long long*)d_base,
                                 entire_data_size, logits, y, features, kl_out, ws, S(stream));
}
int abcd_sampler_fused_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* xc,
                           const float* logits, const float* y, int mode, float tau, float entire_data_size,
                           const float* d_features, const float* g_kl, float* d_xc, float* scratch, float* ws, void* stream) {
    ARG_CHECK(p && g, "sampler_fused_bwd: null argument");
    return sampler_fused_backward(to_sp(p), to_sg(g), B, xc, logits, y, mode, tau, entire_data_size, d_features, g_kl, d_xc,
                                  scratch, ws, S(stream));
}
int abcd_sampler_stats(int B, int K, const float* logits, const float* shape_logits, float* scratch_K, float* out3, void* stream) {
    return sampler_stats(B, K, logits, shape_logits, scratch_K, out3, S(stream));
}
int abcd_softmax_argmax(int B, int K, const float* logits, float* probs, int32_t* argmax, void* stream) {
    return softmax_argmax(B, K, logits, probs, argmax, S(stream));
}
int abcd_mlp_fwd(const abcd_mlp_params_t* p, int rows, const float* x, float* hid, float* y, void* stream) {
    return mlp_forward(to_mlp(*p), rows, x, hid, y, S(stream));
}
int abcd_mlp_bwd(const abcd_mlp_params_t* p, const abcd_mlp_grads_t* g, int rows, const float* x, const float* hid,
                 const float* d_y, float* d_x, float* scratch, void* stream) {
    return mlp_backward(to_mlp(*p), to_mlpg(*g), rows, x, hid, d_y, d_x, scratch, S(stream));
}
int abcd_gauss_latent_fwd(int64_t n, const float* mean, const float* logvar, const float* eps, float* z, float* kl_out,
                          float* partial, void* stream) {
    return gauss_latent_forward((long)n, mean, logvar, eps, z, kl_out, partial, S(stream));
}
int abcd_gauss_latent_bwd(int64_t n, const float* mean, const float* logvar, const float* eps, const float* d_z,
                          const float* g_kl, float* d_mean, float* d_logvar, void* stream) {
    return gauss_latent_backward((long)n, mean, logvar, eps, d_z, g_kl, d_mean, d_logvar, S(stream));
}

size_t abcd_dec_workspace_bytes(int cell, int F, int fin, int B, int sumT) { return dec_workspace_bytes(cell, F, fin, B, sumT); }
int abcd_dec_fwd(const abcd_dec_params_t* p, const abcd_packed_t* pk, const float* feat_in, const float* eps, int ld_eps,
                 const float* gt, int ld_gt, const float* is_offset, float* losses, float* out_x, float* out_mean,
                 float* out_logvar, float* out_offset_logit, void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && pk && feat_in && ws, "dec_fwd: null argument");
    return dec_forward(to_dp(p), to_meta(pk), feat_in, eps, ld_eps, gt, ld_gt, is_offset, losses, out_x, out_mean,
                       out_logvar, out_offset_logit, ws, ws_bytes, S(stream));
}
int abcd_dec_bwd(const abcd_dec_params_t* p, const abcd_dec_grads_t* g, const abcd_packed_t* pk, const float* feat_in,
                 const float* eps, int ld_eps, const float* gt, int ld_gt, const float* is_offset, const float* g_losses,
                 float* d_feat_in, void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && g && pk && feat_in && ws, "dec_bwd: null argument");
    return dec_backward(to_dp(p), to_dg(g), to_meta(pk), feat_in, eps, ld_eps, gt, ld_gt, is_offset, g_losses, d_feat_in,
                        ws, ws_bytes, S(stream));
}

int abcd_dec_bwd_data(const abcd_dec_params_t* p, const abcd_packed_t* pk, const float* feat_in,
                      const float* eps, int ld_eps, const float* gt, int ld_gt, const float* is_offset, const float* g_losses,
                      float* d_feat_in, void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && pk && feat_in && ws, "dec_bwd_data: null argument");
    DecGrads none;
    memset(&none, 0, sizeof(none));
    return dec_backward(to_dp(p), none, to_meta(pk), feat_in, eps, ld_eps, gt, ld_gt, is_offset, g_losses, d_feat_in,
                        ws, ws_bytes, S(stream), 1);
}
int abcd_dec_bwd_weights(const abcd_dec_params_t* p, const abcd_dec_grads_t* g, const abcd_packed_t* pk, const float* feat_in,
                         void* ws, size_t ws_bytes, void* stream) {
    ARG_CHECK(p && g && pk && feat_in && ws, "dec_bwd_weights: null argument");
    return dec_backward(to_dp(p), to_dg(g), to_meta(pk), feat_in, nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr,
                        ws, ws_bytes, S(stream), 2);
}

int abcd_optim_partial_count(void) { return optim_partial_count(); }
int abcd_gradnorm_clip_sgd(float* params, float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                           float max_norm, int zero_grads, float* norm_out, float* partial, void* stream) {
    return gradnorm_clip_sgd(params, grads, momentum_buf, (long)n, lr, momentum, max_norm, zero_grads, norm_out, partial,
                             optim_partial_count(), S(stream));
}
int abcd_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    return philox_fill(out, (long)n, seed, offset, 0, S(stream));
}
int abcd_philox_gumbel(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    return philox_fill(out, (long)n, seed, offset, 1, S(stream));
}
int abcd_philox_fill(float* out, int64_t n, int kind, uint64_t seed, uint64_t offset, const uint64_t* d_base, void* stream) {
    ARG_CHECK(kind == 0 || kind == 1, "philox_fill: kind must be 0 (normal) or 1 (Gumbel)");
    return philox_fill(out, (long)n, seed, offset, kind, S(stream), (const unsigned long long*)d_base);
}
int abcd_counter_add(uint64_t* d_counter, uint64_t delta, void* stream) {
    return counter_add((unsigned long long*)d_counter, delta, S(stream));
}

int abcd_sgemm(int M, int N, int K, const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
               float* C, int64_t ldc, float alpha, const float* bias, int act, int accumulate, int splits,
               float* partials, void* stream) {
    GemmArgs g = gemm_args_zero();
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = (long)lda; g.a_kmajor = a_kmajor;
    g.B = B; g.ldb = (long)ldb; g.b_kmajor = b_kmajor; g.C = C; g.ldc = (long)ldc;
    g.alpha = alpha; g.bias = bias; g.act = act; g.accumulate = accumulate;
    g.splits = splits > 0 ? splits : sgemm_pick_splits(M, N, K);
    if (!partials) g.splits = 1;
    g.partials = partials;
    return sgemm_launch(g, S(stream));
}

}  // extern "C"
