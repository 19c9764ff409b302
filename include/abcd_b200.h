/* abcd_b200.h -- C ABI of the B200-native ABCD-VAE hot path (libabcd_b200.so).
 *
 * The reference (tkc-morita/seq2seq_abcd-vae) has no FFI: its operator interface for this path is the Python
 * class surface of ABCD-VAE/modules/model.py and modules/data_utils.py, which dispatches to stock PyTorch
 * ops.  Each entry point below replaces the PyTorch dispatch of one of those methods; the reference-side
 * binding is a ctypes stub inside the same-named nn.Module.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name starts with h_; buffers (including workspaces and the
 *    tensors saved for backward) are owned by the caller; the library never allocates device memory;
 *  - functions only enqueue work on `stream` (a cudaStream_t passed as void*); they never synchronise;
 *  - return 0 on success, a negative code otherwise; abcd_last_error() returns a thread-local message;
 *  - nothing throws across the ABI; re-entrant across streams/threads after abcd_init();
 *  - fp32 throughout: the recurrent contractions (weights resident in tensor memory) and the GEMMs outside the time loops
 *    run on the tensor cores as 3xTF32 (hi/lo split, fp32 accumulation; <= 3e-6 + 1.5e-8*K relative to the largest
 *    output); the decoder forward and the FFMA fallbacks are IEEE fp32 FFMA; special functions of the Dirichlet KL in fp64;
 *  - recurrent kernels are specialised for rnn_hidden_size = mlp_hidden_size = 256 and output_size <= 129;
 *    anything else is rejected with ABCD_ERR_UNSUPPORTED/ABCD_ERR_ARG -- there is no fallback path.
 *
 * Packed layout (torch.nn.utils.rnn.PackedSequence, ABCD-VAE/modules/data_utils.py:179): frames are time-major,
 * frame (t, b) of the b-th longest segment lives at row offsets[t] + b, b < batch_sizes[t].
 */
#ifndef ABCD_B200_H
#define ABCD_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ABCD_OK 0
#define ABCD_ERR_ARG (-1)
#define ABCD_ERR_CUDA (-2)
#define ABCD_ERR_UNSUPPORTED (-3)
#define ABCD_ERR_WORKSPACE (-4)

#define ABCD_CELL_LSTM 0
#define ABCD_CELL_GRU 1

typedef struct {
    int32_t B, T_max, sumT;
    const int32_t* batch_sizes; /* [T_max]   PackedSequence.batch_sizes */
    const int32_t* offsets;     /* [T_max+1] exclusive prefix sum of batch_sizes */
    const int32_t* lengths;     /* [B]       frames per segment, descending */
} abcd_packed_t;

/* RNN_Variational_Encoder (model.py:40-79): state_dict keys rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l0[_reverse] */
typedef struct {
    int32_t cell, F, H, bidir;
    const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
} abcd_enc_params_t;
typedef struct {
    float* w_ih[2]; float* w_hh[2]; float* b_ih[2]; float* b_hh[2];
} abcd_enc_grads_t;

/* MLP (model.py:316-334): whole_network.{0,2}.{weight,bias} */
typedef struct { int32_t in, hidden, out; const float* w1; const float* b1; const float* w2; const float* b2; } abcd_mlp_params_t;
typedef struct { float* w1; float* b1; float* w2; float* b2; } abcd_mlp_grads_t;

/* ABCDSampler (model.py:538-673) */
typedef struct {
    abcd_mlp_params_t mlp;     /* to_code_like */
    int32_t f, K;
    const float* codebook;     /* [f, K] */
    const float* shape_logits; /* posterior_shape_logits [K] */
    float prior_c;             /* prior_concentration */
} abcd_sampler_params_t;
typedef struct { abcd_mlp_grads_t mlp; float* codebook; float* shape_logits; } abcd_sampler_grads_t;

/* RNN_Variational_Decoder (model.py:84-285) */
typedef struct {
    int32_t cell, F, H, M, fin;
    const float* w_f2h; const float* b_f2h;                                     /* feature2hidden */
    const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh; /* rnn_cell.cell */
    const float* em_w1[2]; const float* em_b1[2]; const float* em_w2[2]; const float* em_b2[2]; /* emission_sampler.to_parameters.mlps.{0,1} */
    const float* off_w1; const float* off_b1; const float* off_w2; const float* off_b2;         /* offset_predictor */
} abcd_dec_params_t;
typedef struct {
    float* w_f2h; float* b_f2h;
    float* w_ih; float* w_hh; float* b_ih; float* b_hh;
    float* em_w1[2]; float* em_b1[2]; float* em_w2[2]; float* em_b2[2];
    float* off_w1; float* off_b1; float* off_w2; float* off_b2;
} abcd_dec_grads_t;

/* ---- library ------------------------------------------------------------------------------------- */
int abcd_init(int device);                 /* requires compute capability 10.x; sets kernel attributes */
const char* abcd_last_error(void);
const char* abcd_version(void);

/* co-resident 16-CTA clusters of a persistent kernel (0 enc_fwd, 1 enc_bwd, 2 dec_fwd, 3 dec_bwd) for output size F */
int abcd_max_active_clusters(int which, int F);
/* debug (env ABCD_TIMING=1): per-phase clock cycles of CTA 0 of the last enc_fwd launch: out10[0..4] = setup, gemm,
 * epilogue, barrier, steps; out10[5..9] = breakdown of the gemm phase */
int abcd_debug_enc_timing(unsigned long long* out10);
/* same for the last decoder launch (env ABCD_TIMING=1: forward kernel, 3: backward kernel; env ABCD_TIMING_CTA=k: the CTA that
 * records, default 0): out16[0..10] = cycles per phase in program order, out16[15] = steps.  The phase list depends on the
 * kernel that ran -- tools/diag.py names the slots of each (two-barrier forward kernel: x wait, x product, cell + publish,
 * barrier A, P2, P2 epilogue + P3a, accumulator read-back, barrier B, P3b, x push, -, with out16[10..13] = MMA issue time of the
 * issuing thread, arrive A, arrive B, P3b loads + sums; backward: wait, B3, B3 epilogue, B2', barrier, -, gate gradients, B1'
 * read-back, barrier, dx reduce + push, B1' relayout, with out16[11] = B1' MMA issue -> done) */
int abcd_debug_dec_timing(unsigned long long* out16);
/* measurement (bench.py "latency_roofline"): nanoseconds per round of the publish -> cluster barrier -> consume exchange that
 * every phase of the persistent recurrent kernels performs, on ONE cluster of `cluster_size` CTAs x 256 threads.
 * mode 0: barrier only; 1: one token through L2 per round; 2: 16-way fan-in of a 128-byte row per CTA.
 * out2[0] = ns per round (device float), scratch >= 16*32 floats. */
int abcd_measure_cluster_exchange(int cluster_size, int iters, int mode, float* out2, float* scratch, void* stream);
/* Recurrent contractions of the encoder forward / backward and the decoder backward (model.py:60-66, 165-196): on the tensor
 * cores (tcgen05 kind::tf32, weights resident in tensor memory, 3xTF32 split, fp32-grade; default) or on the FFMA path.  on = 1/0; -1 only queries.  Returns
 * the previous setting.  Initial value: env ABCD_RNN_TC (default 1).  This is a PROCESS-WIDE debug switch (PyTorch runs backward
 * passes on its own threads, so it cannot be thread-local): the only mutable global of the library; every compute entry point is
 * re-entrant across streams and threads as long as the switch is not flipped concurrently. */
int abcd_set_rnn_tensorcore(int on);
/* kernel launches enqueued by this library since load (what bench.py reports as gpu_launches) */
long long abcd_launch_count(void);
/* optional CUDA-event timing per kernel group on the launching stream (off by default; used by bench.py for the
 * roofline numbers).  abcd_profile_read synchronises the group's events, returns their summed duration and clears. */
int abcd_profile_groups(void);
const char* abcd_profile_name(int id);
void abcd_profile_enable(int on);
int abcd_profile_read(int id, float* total_ms, int* count);

/* ---- front end: data_utils.STFT.__call__ + log_and_normalize (data_utils.py:131-139, learning.py:466) */
int abcd_stft_logamp(const void* wav, int wav_is_int16, int64_t n_samples,
                     const int64_t* seg_onset, const int32_t* seg_nsamp, /* [B] first sample (int64: corpora past 2^31 samples) / sample count, batch order */
                     const abcd_packed_t* pk, int n_fft, int hop, int center, const float* window /*[n_fft]*/,
                     float eps, float normalizer, float* out /*[sumT, ld_out]*/, int64_t ld_out, void* stream);

/* ---- encoder: RNN_Variational_Encoder.forward (model.py:60-66) ------------------------------------- */
size_t abcd_enc_workspace_bytes(int cell, int F, int B, int sumT, int bidir, int save_for_backward);
int abcd_enc_fwd(const abcd_enc_params_t* p, const abcd_packed_t* pk, const float* x /*[sumT, ldx]*/, int ldx,
                 float* last_hidden /*[B, hidden_size_total]*/, int save_for_backward,
                 void* workspace, size_t workspace_bytes, void* stream);
int abcd_enc_bwd(const abcd_enc_params_t* p, const abcd_enc_grads_t* g /*accumulated into*/, const abcd_packed_t* pk,
                 const float* x, int ldx, const float* d_last_hidden, void* workspace, size_t workspace_bytes, void* stream);

/* ---- ABCD sampler --------------------------------------------------------------------------------- */
/* ABCDSampler.forward (model.py:581-590). hid [B,hidden], xc [B,f] are saved for backward. */
int abcd_sampler_logits_fwd(const abcd_sampler_params_t* p, int B, const float* x, float* hid, float* xc, float* logits, void* stream);
int abcd_sampler_logits_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* x,
                            const float* hid, const float* xc, const float* d_logits, float* d_x /*[B,in]*/,
                            float* scratch /* B*(f + max(hidden,f)) floats */, void* stream);
/* ABCDSampler.sample (model.py:592-606). mode 0: softmax (no_sample); 1: Gumbel-softmax with noise [B,K] */
int abcd_sampler_sample_fwd(const abcd_sampler_params_t* p, int B, const float* logits, int mode, float tau,
                            const float* gumbel_noise, float* y /*[B,K] saved*/, float* features /*[B,f]*/, void* stream);
int abcd_sampler_sample_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* y, int mode,
                            float tau, const float* d_features, float* d_logits, float* scratch /* B*K floats */, void* stream);
/* ABCDSampler.kl_divergence (model.py:608-639). kl_ws (abcd_sampler_kl_ws_floats(K) floats, 256B aligned) is saved. */
size_t abcd_sampler_kl_ws_floats(int K);
int abcd_sampler_kl_fwd(const abcd_sampler_params_t* p, int B, const float* logits, float entire_data_size,
                        float* kl_out /*scalar*/, float* kl_ws, void* stream);
int abcd_sampler_kl_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* logits,
                        float entire_data_size, const float* g_kl /*scalar*/, float* d_logits, float* kl_ws, void* stream);
/* The three calls above (after the to_code_like MLP) as ONE forward and ONE backward kernel -- BASELINE.json north star item 3:
 * logits = xc codebook / sqrt(f); y = softmax(logits) (mode 0) or softmax((logits + G) / tau) (mode 1); features = y codebook^T;
 * KL as abcd_sampler_kl_fwd.  G: gumbel_noise [B,K], or NULL to draw it in the kernel from the Philox stream
 * (element e of the [B,K] tensor = component e & 3 of block offset + *d_base + (e >> 2): the same draws abcd_philox_fill makes).
 * The codebook lives in shared memory: needs abcd_sampler_fused_supported(f, K) (f <= 512, K <= 256, f*K*4 <= ~200 KB).
 * ws: abcd_sampler_fused_ws_floats(K) floats (256-byte aligned), no state carried between calls.
 * bwd: d_features [B,f] and the scalar g_kl are the upstream gradients; d_xc [B,f] is written, the codebook and
 * posterior_shape_logits gradients are accumulated into g; scratch: B*K floats. */
int abcd_sampler_fused_supported(int f, int K);
size_t abcd_sampler_fused_ws_floats(int K);
int abcd_sampler_fused_fwd(const abcd_sampler_params_t* p, int B, const float* xc /*[B,f]*/, int mode, float tau, const float* gumbel_noise,
                           uint64_t seed, uint64_t offset, const uint64_t* d_base, float entire_data_size,
                           float* logits /*[B,K]*/, float* y /*[B,K] saved*/, float* features /*[B,f]*/, float* kl_out /*scalar*/,
                           float* ws, void* stream);
int abcd_sampler_fused_bwd(const abcd_sampler_params_t* p, const abcd_sampler_grads_t* g, int B, const float* xc,
                           const float* logits, const float* y, int mode, float tau, float entire_data_size,
                           const float* d_features, const float* g_kl, float* d_xc, float* scratch, float* ws, void* stream);
/* per-batch perplexities of Learner.train (learning.py:171-178): out3 = cluster, batch-mean, Dirichlet-shape */
int abcd_sampler_stats(int B, int K, const float* logits, const float* shape_logits, float* scratch_K, float* out3, void* stream);
/* Encoder.encode (encode.py:28-32): probs = softmax(logits) and the row arg-max category */
int abcd_softmax_argmax(int B, int K, const float* logits, float* probs, int32_t* argmax, void* stream);

/* generic MLP (used by plain/ Sampler: plain/modules/model.py:538-567) */
int abcd_mlp_fwd(const abcd_mlp_params_t* p, int rows, const float* x, float* hid, float* y, void* stream);
int abcd_mlp_bwd(const abcd_mlp_params_t* p, const abcd_mlp_grads_t* g, int rows, const float* x, const float* hid,
                 const float* d_y, float* d_x, float* scratch /* rows*hidden floats */, void* stream);
/* isotropic Gaussian latent: sample_from_isotropic_gaussian / kl_isotropic_to_standard_gaussian (model.py:17-28) */
int abcd_gauss_latent_fwd(int64_t n, const float* mean, const float* logvar, const float* eps, float* z, float* kl_out,
                          float* partial /*148 floats*/, void* stream);
int abcd_gauss_latent_bwd(int64_t n, const float* mean, const float* logvar, const float* eps, const float* d_z,
                          const float* g_kl, float* d_mean, float* d_logvar, void* stream);

/* ---- decoder: RNN_Variational_Decoder._forward_unidirectional (model.py:165-196) ------------------- */
size_t abcd_dec_workspace_bytes(int cell, int F, int fin, int B, int sumT);
int abcd_dec_fwd(const abcd_dec_params_t* p, const abcd_packed_t* pk, const float* feat_in /*[B,fin]*/,
                 const float* eps /*[sumT, ld_eps] N(0,1) draws in packed order*/, int ld_eps,
                 const float* gt /*[sumT, ld_gt] or NULL*/, int ld_gt, const float* is_offset /*[sumT] or NULL*/,
                 float* losses /*[2]: emission NLL, offset BCE*/,
                 float* out_x, float* out_mean, float* out_logvar /*[sumT,F] or NULL*/, float* out_offset_logit /*[sumT] or NULL*/,
                 void* workspace, size_t workspace_bytes, void* stream);
int abcd_dec_bwd(const abcd_dec_params_t* p, const abcd_dec_grads_t* g, const abcd_packed_t* pk, const float* feat_in,
                 const float* eps, int ld_eps, const float* gt, int ld_gt, const float* is_offset,
                 const float* g_losses /*[2] upstream gradients of the two losses*/, float* d_feat_in /*[B,fin] or NULL*/,
                 void* workspace, size_t workspace_bytes, void* stream);

/* abcd_dec_bwd in two halves, so that the caller can put the weight-gradient reductions on a second stream where they
 * overlap the rest of the backward pass (sampler backward, encoder BPTT): _data = loss-head gradients + BPTT + d_feat_in
 * (the critical path), _weights = every parameter gradient, reading what _data left in the workspace.  The caller orders
 * _weights after _data (event) and joins the second stream before using the gradients. */
int abcd_dec_bwd_data(const abcd_dec_params_t* p, const abcd_packed_t* pk, const float* feat_in,
                      const float* eps, int ld_eps, const float* gt, int ld_gt, const float* is_offset, const float* g_losses,
                      float* d_feat_in, void* workspace, size_t workspace_bytes, void* stream);
int abcd_dec_bwd_weights(const abcd_dec_params_t* p, const abcd_dec_grads_t* g, const abcd_packed_t* pk, const float* feat_in,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- clip_grad_norm_ + SGD.step over a flat arena (learning.py:161-163) ---------------------------- */
int abcd_optim_partial_count(void);
int abcd_gradnorm_clip_sgd(float* params, float* grads, float* momentum_buf /*or NULL*/, int64_t n, float lr,
                           float momentum, float max_norm, int zero_grads, float* norm_out /*scalar or NULL*/,
                           float* partial /*abcd_optim_partial_count() floats*/, void* stream);

/* ---- noise --------------------------------------------------------------------------------------- */
int abcd_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);
int abcd_philox_gumbel(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);
/* the same generators with a DEVICE-resident stream position: draw i uses Philox counter offset + *d_base + i/4 (d_base may be
 * NULL).  A training step recorded in a CUDA graph bakes `offset` (its position inside the step) and ends with
 * abcd_counter_add(d_base, draws of the step), so that every replay continues the stream (model.py:19, 604: fresh noise per step). */
int abcd_philox_fill(float* out, int64_t n, int kind /*0 normal, 1 Gumbel*/, uint64_t seed, uint64_t offset, const uint64_t* d_base, void* stream);
int abcd_counter_add(uint64_t* d_counter, uint64_t delta, void* stream);

/* ---- Encoder.encode_dataset (encode.py:38-55): host-side long-format CSV rows of one batch ------------- */
/* Category-major like DataFrame.melt: for k < K, for b < B: "<h_data_ix[b]>,<k>,<h_vals[b*ld+k]><suffix b>\n", the value
 * printed as pandas.to_csv prints a float32 (shortest round-trip digits; scientific outside [1e-4, 1e6)).  h_suffix /
 * h_suffix_off ([B+1] byte offsets; both may be NULL) carry the pre-formatted annotation cells of each segment including
 * their leading comma.  ALL pointers are HOST pointers.  Returns the number of bytes written to h_out, or a negative
 * error code; h_out must hold B*K*abcd_csv_row_bound(longest suffix) bytes. */
int64_t abcd_csv_row_bound(int64_t max_suffix_bytes);
int64_t abcd_csv_format_long(const float* h_vals, int32_t B, int32_t K, int64_t ld, const int64_t* h_data_ix,
                             const char* h_suffix, const int64_t* h_suffix_off, char* h_out, int64_t out_cap);

/* ---- generic fp32 GEMM (exposed for tests): C[M,N] = alpha * A * B (+bias)(tanh) -------------------- */
int abcd_sgemm(int M, int N, int K, const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
               float* C, int64_t ldc, float alpha, const float* bias, int act, int accumulate, int splits,
               float* partials, void* stream);

#ifdef __cplusplus
}
#endif
#endif
