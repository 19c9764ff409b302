// Decoder recurrence with self-feedback: RNN_Variational_Decoder._forward_unidirectional
// (ABCD-VAE/modules/model.py:165-196) and its backward.  Step t feeds the Gaussian sample drawn from step
// t-1's emission parameters back into the cell (model.py:177-180), so nothing can be hoisted out of the
// time loop: the per-step chain is  cell([x_t | h_{t-1}]) -> 2 emission MLPs -> reparameterised sample.
//
// One persistent cluster kernel per direction of time (forward / BPTT).  Per step, three phases separated by
// hardware cluster barriers; every weight slice stays resident in shared memory (see rnn_common.cuh):
//   fwd  P1 gates+cell (K = H+Fp)      P2 MLP hidden layers (K = H)        P3 mean/log-var/sample/NLL (K-split over the hidden units)
//   bwd  B3 d(hidden pre-act) (K=2Fp)  B2' dh + gate gradients             B1' [dh_prev | dx]   (both K-split over the CTA's own slice)
// The offset predictor (model.py:191-195) does not feed the recurrence and runs as plain GEMMs over all
// frames, as do all weight gradients.
#include <stdlib.h>
#include "rnn_common.cuh"
#include "rnn_ts.cuh"
#include "sgemm.cuh"
#include "enc.cuh"

using namespace rnn;

namespace {

constexpr int DEC_N1P = 392;   // backward: pitch of the resident [W_hh | W_ih] rows (H + pad4(F) <= 388 for F <= 129)
constexpr size_t SMEM_MAX = 232448;   // 227 KB of dynamic shared memory per CTA on sm_100
constexpr float LOG_2PI = 1.8378770664093453f;
constexpr int DEC_MAX_F = 129;

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int xcols_per_cta(int F) { return (F + CL - 1) / CL; }

struct DecK {
    int cell, F, Fp, n_tiles, B, T_max;
    int kc;   // staged K-chunk of this launch (largest of 128/64/32 that fits beside the resident weights)
    int ld_eps, ld_gt;
    const int* batch_sizes; const int* offsets; const int* lengths;
    // parameters
    const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
    const float* em_w1[2]; const float* em_b1[2]; const float* em_w2[2]; const float* em_b2[2];
    // activations (see DecWs)
    float* Hall; float* Call; float* Xall; float* Gs; float* Hid; float* Mean; float* Lv;
    const float* eps; const float* gt;
    int timing;   // debug: per-phase clock cycles of CTA `timing_cta` -> g_dec_timing
    int timing_cta;
    float* nll_partial;   // [gridDim.x]
    float* Pbuf;          // [n_tiles][CL][ROWS][2Fp] partial emission parameters exchanged inside a cluster
    char* himg; char* ximg;   // tensor-core forward: per tile, 2 parities of operand-tile images of h (K = H) and x (K = Fp)
    // backward only
    float* dML; float* dHidPre; float* dG; const float* dHoff; float* dHC0;
    float* dbias;   // [n_tiles][4H]: per-tile column sums of dG (recurrent bias gradients), interleaved layout
    float* dbias1;  // [n_tiles][2M]: per-tile column sums of dHidPre (bias gradients of the two MLP input layers)
    float* Pb2;     // [n_tiles][CL][ROWS][H]     partial dh from the MLP input layers, exchanged inside a cluster
    float* Pb1;     // [n_tiles][CL][ROWS][H+Fp]  partial [dh_prev | dx] from the cell
};

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
__device__ unsigned long long g_dec_timing[16];

// TC = true: the gate contraction P1 runs on tcgen05 with [W_hh | W_ih]^T of this CTA's 64 gate columns resident in tensor
// memory (TS form, M = 64; rnn_ts.cuh); shared memory is then
// [activation tiles + raw staging ring (also the P2 staging buffers) | W_lo tiles | W1 | W2s | reduction scratch].
template <int CELL, int KC, bool TC = false>
__global__ void __launch_bounds__(NTHREADS, 1) dec_fwd_kernel(DecK p) {
    extern __shared__ __align__(1024) float smem_fwd[];
    float* const smem = smem_fwd;
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int XC = xcols_per_cta(F);
    constexpr int HSP = 2 * U + 4;            // pitch of the CTA's own hidden-layer slice kept in shared memory
    const int ts_nsl = (H + Fp + TS_SLAB - 1) / TS_SLAB;               // K slabs of the gate contraction
    // TC: [activation tiles (aliased by the P2 staging) | W_lo tiles | raw ring | W1 | W2s]; the operand tiles must stay
    // 1024-byte aligned (the swizzle is a function of the absolute shared-memory address)
    char* ts_wlo = reinterpret_cast<char*>(smem) + (size_t)TSI_STAGES * TS_STAGE_BYTES;
    const int ts_nslx = (Fp + TS_SLAB - 1) / TS_SLAB;                  // slabs of the x image
    float* Wg = smem;                         // FFMA: [H+Fp][U*4]: recurrent rows first, then the input rows
    float* W1 = TC ? reinterpret_cast<float*>(ts_wlo + ts_wlo_bytes_m<U * 4>(H + Fp)) : Wg + (Fp + H) * (U * 4);   // [H][2U]
    float* W2s = W1 + H * (2 * U);            // [U][2Fp]: rows of the output layers that multiply this CTA's hidden units
    const Stage st = make_stage(TC ? smem : W2s + U * 2 * Fp, KC);
    float* red = TC ? W2s + U * 2 * Fp : W2s + U * 2 * Fp + stage_floats(KC);   // 40 floats
    float* Hs = st.buf0;                      // [ROWS][HSP] (the staging buffers are idle between phases)
    __shared__ uint64_t ts_bars[TS_STAGES + 1];
    __shared__ uint64_t tsi_bars[2 * TSI_STAGES + 1];
    __shared__ uint32_t ts_tmem_slot;
    TsCtx ts;
    TsImgCtx tsi;
    char* himg[2] = {nullptr, nullptr};
    char* ximg[2] = {nullptr, nullptr};

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;
    const int xc0 = rank * XC;
    const int xc1 = min(F, xc0 + XC);

    // ---- resident weights ----
    if (TC) {
        static_assert(!TC || (size_t)2 * ROWS * (KC + 4) * sizeof(float) <= (size_t)TSI_STAGES * TS_STAGE_BYTES, "P2 staging must fit in the activation tiles");
        if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
        auto wval = [&](int col, int k) -> float {
            const int u = col >> 2, s = col & 3;
            if (k < H) {
                const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
                return gate >= 0 ? __ldg(p.w_hh + (long)(gate * H + u0 + u) * H + k) : 0.f;
            }
            const int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            return (gate >= 0 && k - H < F) ? __ldg(p.w_ih + (long)(gate * H + u0 + u) * F + (k - H)) : 0.f;
        };
        ts_init_m<U * 4>(ts, reinterpret_cast<char*>(smem), ts_wlo, nullptr, ts_bars, &ts_tmem_slot, ts_nsl, wval, tid);
        ts_img_init_bars(tsi, reinterpret_cast<char*>(smem), ts_wlo, tsi_bars, ts.tmem, tid);
        himg[0] = p.himg + (size_t)(tile * 2) * ts_image_bytes(ABCD_H); himg[1] = himg[0] + ts_image_bytes(ABCD_H);
        ximg[0] = p.ximg + (size_t)(tile * 2) * ts_nslx * TS_STAGE_BYTES; ximg[1] = ximg[0] + (size_t)ts_nslx * TS_STAGE_BYTES;
        // clear the images once (padding k >= F of the x image must be zero; stale rows must at least be finite)
        {
            float4* z = reinterpret_cast<float4*>(himg[0]);
            const size_t n16 = 2 * ts_image_bytes(ABCD_H) / 16;
            for (size_t i = (size_t)rank * NTHREADS + tid; i < n16; i += (size_t)CL * NTHREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* zx = reinterpret_cast<float4*>(ximg[0]);
            const size_t nx16 = (size_t)2 * ts_nslx * TS_STAGE_BYTES / 16;
            for (size_t i = (size_t)rank * NTHREADS + tid; i < nx16; i += (size_t)CL * NTHREADS) zx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ts_image_publish();
        cluster_sync_all();
    }
    for (int idx = tid; !TC && idx < (Fp + H) * (U * 4); idx += NTHREADS) {
        int col = idx % (U * 4), k = idx / (U * 4);
        int u = col >> 2, s = col & 3;
        float v = 0.f;
        if (k >= H) {
            int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            if (k - H < F && gate >= 0) v = p.w_ih[(long)(gate * H + u0 + u) * F + (k - H)];
        } else {
            int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            if (gate >= 0) v = p.w_hh[(long)(gate * H + u0 + u) * H + k];
        }
        Wg[idx] = v;
    }
    for (int idx = tid; idx < H * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);
        int which = col / U, u = col % U;
        W1[idx] = p.em_w1[which][(long)(u0 + u) * H + k];
    }
    for (int idx = tid; idx < U * 2 * Fp; idx += NTHREADS) {
        int c = idx % (2 * Fp), kk = idx / (2 * Fp);
        int which = c >= Fp, cc = which ? c - Fp : c;     // mean columns [0,Fp), log-variance columns [Fp,2Fp)
        W2s[idx] = cc < F ? p.em_w2[which][(long)cc * M + u0 + kk] : 0.f;
    }

    // ---- per-thread constants for the three phases ----
    const int KS1 = pick_ks_fit(U, U * 4, KC);
    const ThreadMap tm1 = make_map(tid, U, KS1);
    const int unit = u0 + tm1.cg;
    float bias1[4];
    if (CELL == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) bias1[g] = p.b_ih[g * H + unit] + p.b_hh[g * H + unit];
    } else {
        bias1[0] = p.b_ih[unit] + p.b_hh[unit];
        bias1[1] = p.b_ih[H + unit] + p.b_hh[H + unit];
        bias1[2] = p.b_ih[2 * H + unit];
        bias1[3] = p.b_hh[2 * H + unit];
    }
    const int KS2 = pick_ks_fit(2 * U / 4, 2 * U, KC);
    const ThreadMap tm2 = make_map(tid, 2 * U / 4, KS2);
    float bias2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int col = tm2.cg * 4 + j;
        bias2[j] = p.em_b1[col / U][u0 + col % U];
    }
    const int ncg3 = 2 * Fp / 4;                                      // column groups of the local emission GEMM
    float* Pmine = p.Pbuf + (long)blockIdx.x * ROWS * (2 * Fp);       // this CTA's partial emission parameters
    const float* Ptile = p.Pbuf + (long)tile * CL * ROWS * (2 * Fp);  // partials of all CTAs of the cluster

    // recurrent state of (row, unit) pairs owned by this thread
    float creg[RT], hreg[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        int b = tile + p.n_tiles * (tm1.rg + i * RG);
        creg[i] = 0.f; hreg[i] = 0.f;
        if (b < p.B) {
            hreg[i] = p.Hall[(long)b * H + unit];
            if (CELL == 0) creg[i] = p.Call[(long)b * H + unit];
        }
    }
    const int Tt = p.lengths[tile];
    float nll = 0.f;
    const int n_tiles = p.n_tiles, B = p.B;
    const float* Hall = p.Hall; const float* Xall = p.Xall;
    if (TC) {
        // images of the initial state (read by step 0 with parity (0 - 1) & 1 = 1)
        if (tm1.ks == 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const int r = tm1.rg + i * RG;
                if (tile + p.n_tiles * r < p.B) ts_image_store(himg[1], unit, r, hreg[i]);
            }
        }
        for (int o = tid; o < ROWS * XC; o += NTHREADS) {
            const int r = o / XC, c = xc0 + o % XC;
            const long b = tile + (long)n_tiles * r;
            if (c < xc1 && b < B) ts_image_store(ximg[1], c, r, Xall[b * Fp + c]);
        }
        ts_image_publish();
        cluster_sync_all();
    }
    __syncthreads();
    const bool timing = p.timing == 1 && blockIdx.x == 0 && tid == (TC ? 64 : 0);
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }

    for (int t = 0; t < Tt; ++t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const long off_t = p.offsets[t];
        const long prev_base = t == 0 ? 0 : (long)B + p.offsets[t - 1];

        // ---- P1: gates + cell ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rpx = [=](int r) -> const float* { return Xall + (prev_base + tile + (long)n_tiles * r) * Fp; };
            auto rph = [=](int r) -> const float* { return Hall + (prev_base + tile + (long)n_tiles * r) * ABCD_H; };
            DEC_TICK(0)
            // h_{t-1} has been visible since the first barrier of the previous step; x_t only after the barrier whose
            // arrive ended that step: the h part of the contraction overlaps that barrier
            if (TC) phase_gemm_ts_img2<U * 4>(acc, tsi, himg[(t - 1) & 1], H / TS_SLAB, ximg[(t - 1) & 1], ts_nslx, na, tm1, tid, t > 0);
            else phase_gemm<NTHREADS, KC, U * 4>(acc, st, H, rph, Fp, rpx, na, Wg, tm1, KS1, tid, timing ? tacc + 10 : nullptr, t > 0 ? H / KC : -1);
            DEC_TICK(1)
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                int r = tm1.rg + i * RG;
                if (r >= na || tm1.ks != 0) continue;
                long frame = off_t + tile + (long)n_tiles * r;
                float h;
                if (CELL == 0) {
                    float ig = sigmoid_acc(acc[i][0] + bias1[0]);
                    float fg = sigmoid_acc(acc[i][1] + bias1[1]);
                    float gg = tanhf(acc[i][2] + bias1[2]);
                    float og = sigmoid_acc(acc[i][3] + bias1[3]);
                    float c = fg * creg[i] + ig * gg;
                    h = og * tanhf(c);
                    creg[i] = c;
                    *reinterpret_cast<float4*>(p.Gs + frame * (4 * H) + unit * 4) = make_float4(ig, fg, gg, og);
                    p.Call[((long)B + frame) * H + unit] = c;
                } else {
                    float rg_ = sigmoid_acc(acc[i][0] + bias1[0]);
                    float zg = sigmoid_acc(acc[i][1] + bias1[1]);
                    float hn = acc[i][3] + bias1[3];
                    float ng = tanhf(acc[i][2] + bias1[2] + rg_ * hn);
                    h = (1.f - zg) * ng + zg * hreg[i];
                    *reinterpret_cast<float4*>(p.Gs + frame * (4 * H) + unit * 4) = make_float4(rg_, zg, ng, hn);
                }
                hreg[i] = h;
                p.Hall[((long)B + frame) * H + unit] = h;
                if (TC) ts_image_store(himg[t & 1], unit, r, h);
            }
        }
        if (TC) ts_image_publish();
        DEC_TICK(2)
        cluster_sync_all();
        DEC_TICK(3)

        // ---- P2: hidden layers of the two emission MLPs ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rph = [=](int r) -> const float* { return Hall + ((long)B + off_t + tile + (long)n_tiles * r) * ABCD_H; };
            phase_gemm<NTHREADS, KC, 2 * U>(acc, st, H, rph, 0, NullRow(), na, W1, tm2, KS2, tid);
            DEC_TICK(4)
            if (tm2.active && tm2.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm2.rg + i * RG;
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm2.cg * 4;
                    int which = col / U, uu = col % U;
                    float4 v = make_float4(tanhf(acc[i][0] + bias2[0]), tanhf(acc[i][1] + bias2[1]),
                                           tanhf(acc[i][2] + bias2[2]), tanhf(acc[i][3] + bias2[3]));
                    *reinterpret_cast<float4*>(p.Hid + frame * (2 * M) + which * M + u0 + uu) = v;
                    *reinterpret_cast<float4*>(Hs + r * HSP + col) = v;
                }
            }
            __syncthreads();
            // ---- P3a: this CTA's share of the emission parameters.  The output layers contract over the hidden
            // units, which are partitioned over the cluster: every CTA multiplies its own 2U hidden columns (still
            // in shared memory) by the matching rows of the output weights and publishes a partial [rows][2Fp];
            // the owners of the x columns add the CL partials after the barrier (fixed order). ----
            for (int item = tid; item < ncg3 * 16; item += NTHREADS) {
                const int cg = item % ncg3, r0 = item / ncg3;   // rows r0 + 16 i
                if (r0 >= na) continue;
                const float* a = Hs + r0 * HSP + ((cg * 4 < Fp) ? 0 : U);
                const float* w = W2s + cg * 4;
                float o[5][4];
#pragma unroll
                for (int i = 0; i < 5; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
                for (int kq = 0; kq < U / 4; ++kq) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + (kq * 4 + 0) * (2 * Fp));
                    const float4 w1 = *reinterpret_cast<const float4*>(w + (kq * 4 + 1) * (2 * Fp));
                    const float4 w2 = *reinterpret_cast<const float4*>(w + (kq * 4 + 2) * (2 * Fp));
                    const float4 w3 = *reinterpret_cast<const float4*>(w + (kq * 4 + 3) * (2 * Fp));
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 av = *reinterpret_cast<const float4*>(a + i * 16 * HSP + kq * 4);
                        o[i][0] = fmaf(av.x, w0.x, o[i][0]); o[i][1] = fmaf(av.x, w0.y, o[i][1]); o[i][2] = fmaf(av.x, w0.z, o[i][2]); o[i][3] = fmaf(av.x, w0.w, o[i][3]);
                        o[i][0] = fmaf(av.y, w1.x, o[i][0]); o[i][1] = fmaf(av.y, w1.y, o[i][1]); o[i][2] = fmaf(av.y, w1.z, o[i][2]); o[i][3] = fmaf(av.y, w1.w, o[i][3]);
                        o[i][0] = fmaf(av.z, w2.x, o[i][0]); o[i][1] = fmaf(av.z, w2.y, o[i][1]); o[i][2] = fmaf(av.z, w2.z, o[i][2]); o[i][3] = fmaf(av.z, w2.w, o[i][3]);
                        o[i][0] = fmaf(av.w, w3.x, o[i][0]); o[i][1] = fmaf(av.w, w3.y, o[i][1]); o[i][2] = fmaf(av.w, w3.z, o[i][2]); o[i][3] = fmaf(av.w, w3.w, o[i][3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int r = r0 + 16 * i;
                    if (r < na) *reinterpret_cast<float4*>(Pmine + (long)r * (2 * Fp) + cg * 4) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
                }
            }
        }
        DEC_TICK(5)
        cluster_sync_all();
        DEC_TICK(6)

        // ---- P3b: sum the partials of this CTA's x columns, reparameterised sample, NLL ----
        for (int o = tid; o < na * XC; o += NTHREADS) {
            const int r = o / XC, c = xc0 + o % XC;
            if (c >= xc1) continue;
            const long frame = off_t + tile + (long)n_tiles * r;
            const float e = __ldg(p.eps + frame * p.ld_eps + c);
            const float g = p.gt ? __ldg(p.gt + frame * p.ld_gt + c) : 0.f;
            float pm[CL], pv[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) {
                const float* q = Ptile + ((long)src * ROWS + r) * (2 * Fp) + c;
                pm[src] = ld_cg(q);
                pv[src] = ld_cg(q + Fp);
            }
            float mean = __ldg(p.em_b2[0] + c), lv = __ldg(p.em_b2[1] + c);
#pragma unroll
            for (int src = 0; src < CL; ++src) { mean += pm[src]; lv += pv[src]; }
            const float x = mean + expf(0.5f * lv) * e;
            p.Mean[frame * Fp + c] = mean;
            p.Lv[frame * Fp + c] = lv;
            p.Xall[((long)B + frame) * Fp + c] = x;
            if (TC) ts_image_store(ximg[t & 1], c, r, x);
            if (p.gt) {
                const float d = g - mean;
                nll += 0.5f * (LOG_2PI + lv + d * expf(-lv) * d);
            }
        }
        DEC_TICK(7)
        if (TC) ts_image_publish();
        DEC_TICK(8)
        cluster_arrive_release();   // matched by the wait inside the next step's P1 (or the one after the loop)
        DEC_TICK(9)
    }
#undef DEC_TICK
    if (Tt > 0) cluster_wait_acquire();
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) g_dec_timing[i] = tacc[i]; }
    float tot = block_sum(nll, red);
    if (tid == 0) p.nll_partial[blockIdx.x] = tot;
    if (TC) ts_fini(ts, tid);
}

// -------------------------------------------------------------------------------------------------
// forward, pair variant (default for output_size <= 128): the gate contraction P1 runs on tcgen05 with full-rate M = 128
// instructions -- the two CTAs of a pair share the 128 gate columns of 32 units and split K = [h | x] between them, hi AND lo
// parts of the weights resident in tensor memory (rnn_ts.cuh, "pair variant").  P2 / P3 are those of dec_fwd_kernel.
// Shared memory: [activation ring (also the P2 staging buffers, the hidden-layer slice and the accumulator bounce) | W1 | W2s | scratch]
// -------------------------------------------------------------------------------------------------
constexpr int DFP_KC = 64;   // staged K-chunk of P2 (its two staging buffers alias the activation ring)
static_assert((size_t)2 * ROWS * (DFP_KC + 4) * sizeof(float) <= (size_t)TSP_STAGES * TS_STAGE_BYTES, "P2 staging must fit in the activation ring");

template <int CELL>
__global__ void __launch_bounds__(NTHREADS, 1) dec_fwd_pair_kernel(DecK p) {
    extern __shared__ __align__(1024) float smem_pair[];
    float* const smem = smem_pair;
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int XC = xcols_per_cta(F);
    constexpr int KC = DFP_KC;
    constexpr int HSP = 2 * U + 4;            // pitch of the CTA's own hidden-layer slice kept in shared memory
    const int ts_nslx = (Fp + TS_SLAB - 1) / TS_SLAB;                  // slabs of the x image
    const int nsl_total = H / TS_SLAB + ts_nslx;
    float* W1 = reinterpret_cast<float*>(reinterpret_cast<char*>(smem) + (size_t)TSP_STAGES * TS_STAGE_BYTES);   // [H][2U]
    float* W2s = W1 + H * (2 * U);            // [U][2Fp]: rows of the output layers that multiply this CTA's hidden units
    float* red = W2s + U * 2 * Fp;            // 40 floats
    const Stage st = make_stage(smem, KC);
    float* Hs = st.buf0;                      // [ROWS][HSP] (the staging buffers are idle between phases)
    __shared__ uint64_t tsp_bars[TSP_STAGES + 1];
    __shared__ uint32_t tsp_tmem_slot;
    TsPairCtx tsp;

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;                  // the 16 units this CTA finalises (cell update, MLP slices)
    const int pair_u0 = (rank >> 1) * 2 * U;  // the 32 units whose gate columns the pair contracts
    const int xc0 = rank * XC;
    const int xc1 = min(F, xc0 + XC);

    // ---- resident weights ----
    if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
    {
        auto wval = [&](int col, int k) -> float {
            const int u = pair_u0 + (col >> 2), s = col & 3;
            if (k < H) {
                const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
                return gate >= 0 ? __ldg(p.w_hh + (long)(gate * H + u) * H + k) : 0.f;
            }
            const int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            return (gate >= 0 && k - H < F) ? __ldg(p.w_ih + (long)(gate * H + u) * F + (k - H)) : 0.f;
        };
        ts_pair_init(tsp, reinterpret_cast<char*>(smem), tsp_bars, &tsp_tmem_slot, rank & 1, nsl_total, wval, tid);
    }
    char* himg[2]; char* ximg[2];
    himg[0] = p.himg + (size_t)(tile * 2) * ts_image_bytes(ABCD_H); himg[1] = himg[0] + ts_image_bytes(ABCD_H);
    ximg[0] = p.ximg + (size_t)(tile * 2) * ts_nslx * TS_STAGE_BYTES; ximg[1] = ximg[0] + (size_t)ts_nslx * TS_STAGE_BYTES;
    // clear the images once (padding k >= F of the x image must be zero; stale rows must at least be finite)
    {
        float4* z = reinterpret_cast<float4*>(himg[0]);
        const size_t n16 = 2 * ts_image_bytes(ABCD_H) / 16;
        for (size_t i = (size_t)rank * NTHREADS + tid; i < n16; i += (size_t)CL * NTHREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* zx = reinterpret_cast<float4*>(ximg[0]);
        const size_t nx16 = (size_t)2 * ts_nslx * TS_STAGE_BYTES / 16;
        for (size_t i = (size_t)rank * NTHREADS + tid; i < nx16; i += (size_t)CL * NTHREADS) zx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ts_image_publish();
    cluster_sync_all();
    for (int idx = tid; idx < H * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);
        int which = col / U, u = col % U;
        W1[idx] = p.em_w1[which][(long)(u0 + u) * H + k];
    }
    for (int idx = tid; idx < U * 2 * Fp; idx += NTHREADS) {
        int c = idx % (2 * Fp), kk = idx / (2 * Fp);
        int which = c >= Fp, cc = which ? c - Fp : c;     // mean columns [0,Fp), log-variance columns [Fp,2Fp)
        W2s[idx] = cc < F ? p.em_w2[which][(long)cc * M + u0 + kk] : 0.f;
    }

    // ---- per-thread constants for the three phases ----
    const ThreadMap tm1 = make_map(tid, U, 1);      // cell update: warps 0..3, (unit, 8 row groups)
    const int unit = u0 + tm1.cg;
    float bias1[4];
    if (CELL == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) bias1[g] = p.b_ih[g * H + unit] + p.b_hh[g * H + unit];
    } else {
        bias1[0] = p.b_ih[unit] + p.b_hh[unit];
        bias1[1] = p.b_ih[H + unit] + p.b_hh[H + unit];
        bias1[2] = p.b_ih[2 * H + unit];
        bias1[3] = p.b_hh[2 * H + unit];
    }
    const int KS2 = pick_ks_fit(2 * U / 4, 2 * U, KC);
    const ThreadMap tm2 = make_map(tid, 2 * U / 4, KS2);
    float bias2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int col = tm2.cg * 4 + j;
        bias2[j] = p.em_b1[col / U][u0 + col % U];
    }
    const int ncg3 = 2 * Fp / 4;                                      // column groups of the local emission GEMM
    float* Pmine = p.Pbuf + (long)blockIdx.x * ROWS * (2 * Fp);       // this CTA's partial emission parameters
    const float* Ptile = p.Pbuf + (long)tile * CL * ROWS * (2 * Fp);  // partials of all CTAs of the cluster
    // pair exchange of the gate partial sums (the backward pass's partial buffer is idle during the forward pass)
    float* pair_send = p.Pb2 + (size_t)(blockIdx.x ^ 1) * ROWS * H;
    const float* pair_recv = p.Pb2 + (size_t)blockIdx.x * ROWS * H;

    // recurrent state of (row, unit) pairs owned by this thread
    float creg[RT], hreg[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        int b = tile + p.n_tiles * (tm1.rg + i * RG);
        creg[i] = 0.f; hreg[i] = 0.f;
        if (b < p.B && tm1.active) {
            hreg[i] = p.Hall[(long)b * H + unit];
            if (CELL == 0) creg[i] = p.Call[(long)b * H + unit];
        }
    }
    const int Tt = p.lengths[tile];
    float nll = 0.f;
    const int n_tiles = p.n_tiles, B = p.B;
    const float* Hall = p.Hall; const float* Xall = p.Xall;
    // images of the initial state (read by step 0 with parity (0 - 1) & 1 = 1)
    if (tm1.active) {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = tm1.rg + i * RG;
            if (tile + p.n_tiles * r < p.B) ts_image_store(himg[1], unit, r, hreg[i]);
        }
    }
    for (int o = tid; o < ROWS * XC; o += NTHREADS) {
        const int r = o / XC, c = xc0 + o % XC;
        const long b = tile + (long)n_tiles * r;
        if (c < xc1 && b < B) ts_image_store(ximg[1], c, r, Xall[b * Fp + c]);
    }
    ts_image_publish();
    cluster_sync_all();
    __syncthreads();
    const bool timing = p.timing == 1 && blockIdx.x == 0 && tid == 64;
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }

    for (int t = 0; t < Tt; ++t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const long off_t = p.offsets[t];

        // ---- P1: gates + cell ----
        {
            float acc[RT][4];
            zero_acc(acc);
            DEC_TICK(0)
            // h_{t-1} has been visible since the first barrier of the previous step; x_t only after the barrier whose
            // arrive ended that step: the h slabs of the contraction overlap that barrier
            phase_gemm_ts_pair(acc, tsp, himg[(t - 1) & 1], ximg[(t - 1) & 1], pair_send, pair_recv, na, tm1, tid, t > 0);
            DEC_TICK(1)
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                int r = tm1.rg + i * RG;
                if (r >= na || !tm1.active) continue;
                long frame = off_t + tile + (long)n_tiles * r;
                float h;
                if (CELL == 0) {
                    float ig = sigmoid_acc(acc[i][0] + bias1[0]);
                    float fg = sigmoid_acc(acc[i][1] + bias1[1]);
                    float gg = tanhf(acc[i][2] + bias1[2]);
                    float og = sigmoid_acc(acc[i][3] + bias1[3]);
                    float c = fg * creg[i] + ig * gg;
                    h = og * tanhf(c);
                    creg[i] = c;
                    *reinterpret_cast<float4*>(p.Gs + frame * (4 * H) + unit * 4) = make_float4(ig, fg, gg, og);
                    p.Call[((long)B + frame) * H + unit] = c;
                } else {
                    float rg_ = sigmoid_acc(acc[i][0] + bias1[0]);
                    float zg = sigmoid_acc(acc[i][1] + bias1[1]);
                    float hn = acc[i][3] + bias1[3];
                    float ng = tanhf(acc[i][2] + bias1[2] + rg_ * hn);
                    h = (1.f - zg) * ng + zg * hreg[i];
                    *reinterpret_cast<float4*>(p.Gs + frame * (4 * H) + unit * 4) = make_float4(rg_, zg, ng, hn);
                }
                hreg[i] = h;
                p.Hall[((long)B + frame) * H + unit] = h;
                ts_image_store(himg[t & 1], unit, r, h);
            }
        }
        ts_image_publish();
        DEC_TICK(2)
        cluster_sync_all();
        DEC_TICK(3)

        // ---- P2: hidden layers of the two emission MLPs ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rph = [=](int r) -> const float* { return Hall + ((long)B + off_t + tile + (long)n_tiles * r) * ABCD_H; };
            phase_gemm<NTHREADS, KC, 2 * U>(acc, st, H, rph, 0, NullRow(), na, W1, tm2, KS2, tid);
            DEC_TICK(4)
            if (tm2.active && tm2.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm2.rg + i * RG;
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm2.cg * 4;
                    int which = col / U, uu = col % U;
                    float4 v = make_float4(tanhf(acc[i][0] + bias2[0]), tanhf(acc[i][1] + bias2[1]),
                                           tanhf(acc[i][2] + bias2[2]), tanhf(acc[i][3] + bias2[3]));
                    *reinterpret_cast<float4*>(p.Hid + frame * (2 * M) + which * M + u0 + uu) = v;
                    *reinterpret_cast<float4*>(Hs + r * HSP + col) = v;
                }
            }
            __syncthreads();
            // ---- P3a: this CTA's share of the emission parameters (K-split over the hidden units, see dec_fwd_kernel) ----
            for (int item = tid; item < ncg3 * 16; item += NTHREADS) {
                const int cg = item % ncg3, r0 = item / ncg3;   // rows r0 + 16 i
                if (r0 >= na) continue;
                const float* a = Hs + r0 * HSP + ((cg * 4 < Fp) ? 0 : U);
                const float* w = W2s + cg * 4;
                float o[5][4];
#pragma unroll
                for (int i = 0; i < 5; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
                for (int kq = 0; kq < U / 4; ++kq) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + (kq * 4 + 0) * (2 * Fp));
                    const float4 w1 = *reinterpret_cast<const float4*>(w + (kq * 4 + 1) * (2 * Fp));
                    const float4 w2 = *reinterpret_cast<const float4*>(w + (kq * 4 + 2) * (2 * Fp));
                    const float4 w3 = *reinterpret_cast<const float4*>(w + (kq * 4 + 3) * (2 * Fp));
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 av = *reinterpret_cast<const float4*>(a + i * 16 * HSP + kq * 4);
                        o[i][0] = fmaf(av.x, w0.x, o[i][0]); o[i][1] = fmaf(av.x, w0.y, o[i][1]); o[i][2] = fmaf(av.x, w0.z, o[i][2]); o[i][3] = fmaf(av.x, w0.w, o[i][3]);
                        o[i][0] = fmaf(av.y, w1.x, o[i][0]); o[i][1] = fmaf(av.y, w1.y, o[i][1]); o[i][2] = fmaf(av.y, w1.z, o[i][2]); o[i][3] = fmaf(av.y, w1.w, o[i][3]);
                        o[i][0] = fmaf(av.z, w2.x, o[i][0]); o[i][1] = fmaf(av.z, w2.y, o[i][1]); o[i][2] = fmaf(av.z, w2.z, o[i][2]); o[i][3] = fmaf(av.z, w2.w, o[i][3]);
                        o[i][0] = fmaf(av.w, w3.x, o[i][0]); o[i][1] = fmaf(av.w, w3.y, o[i][1]); o[i][2] = fmaf(av.w, w3.z, o[i][2]); o[i][3] = fmaf(av.w, w3.w, o[i][3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int r = r0 + 16 * i;
                    if (r < na) *reinterpret_cast<float4*>(Pmine + (long)r * (2 * Fp) + cg * 4) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
                }
            }
        }
        DEC_TICK(5)
        cluster_sync_all();
        DEC_TICK(6)

        // ---- P3b: sum the partials of this CTA's x columns, reparameterised sample, NLL ----
        for (int o = tid; o < na * XC; o += NTHREADS) {
            const int r = o / XC, c = xc0 + o % XC;
            if (c >= xc1) continue;
            const long frame = off_t + tile + (long)n_tiles * r;
            const float e = __ldg(p.eps + frame * p.ld_eps + c);
            const float g = p.gt ? __ldg(p.gt + frame * p.ld_gt + c) : 0.f;
            float pm[CL], pv[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) {
                const float* q = Ptile + ((long)src * ROWS + r) * (2 * Fp) + c;
                pm[src] = ld_cg(q);
                pv[src] = ld_cg(q + Fp);
            }
            float mean = __ldg(p.em_b2[0] + c), lv = __ldg(p.em_b2[1] + c);
#pragma unroll
            for (int src = 0; src < CL; ++src) { mean += pm[src]; lv += pv[src]; }
            const float x = mean + expf(0.5f * lv) * e;
            p.Mean[frame * Fp + c] = mean;
            p.Lv[frame * Fp + c] = lv;
            p.Xall[((long)B + frame) * Fp + c] = x;
            ts_image_store(ximg[t & 1], c, r, x);
            if (p.gt) {
                const float d = g - mean;
                nll += 0.5f * (LOG_2PI + lv + d * expf(-lv) * d);
            }
        }
        DEC_TICK(7)
        ts_image_publish();
        DEC_TICK(8)
        cluster_arrive_release();   // matched by the wait inside the next step's P1 (or the one after the loop)
        DEC_TICK(9)
    }
#undef DEC_TICK
    if (Tt > 0) cluster_wait_acquire();
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) g_dec_timing[i] = tacc[i]; }
    float tot = block_sum(nll, red);
    if (tid == 0) p.nll_partial[blockIdx.x] = tot;
    ts_pair_fini(tsp, tid);
}

// -------------------------------------------------------------------------------------------------
// forward, pair variant ONE STEP AHEAD (default for output_size <= 88; rnn_ts.cuh "one step ahead"): the recurrent product
// h_t W_hh of step t + 1 runs on tcgen05 underneath the emission MLPs of step t; the critical path of P1 is the input product
// x_t W_ih (FFMA, K = Fp) plus the cell update, both on all 256 threads (16 units x 16 row groups, 5 rows per thread):
//   P1   gates = gpre (h_{t-1} W_hh, ready) + x_{t-1} W_ih;  cell;  h_t published (rows + operand-tile image)   -- cluster barrier A --
//        [thread 224: bulk copies of the h_t image]                  (saved gates / cell state are stored under the barrier)
//   P2   hidden layers of the two emission MLPs (staged FFMA, K = H)
//        [thread 224: MMAs of h_t W_hh while warps 0..6 do P3a]      P3a partial emission parameters (K-split)
//        accumulator read-back: own half -> gpre, partner's half -> L2                                            -- cluster barrier B --
//   P3b  emission parameters of this CTA's x columns, sample x_t, NLL;  x_t is PUSHED into the staging buffer of every CTA of
//        the cluster through distributed shared memory with st.async: every 4-byte store also signals the consumer's mbarrier,
//        so the third exchange of a step needs no fence, no cluster barrier and no L2 round trip               -- mbarrier (x bytes) --
//        [off the critical path: Mean / Lv / Xall stores for the backward pass, partner's half of gpre added]
// Shared memory: [h ring 4 x 24 KB | W1 | W2s | Wx | gpre | staging (P2: 2 x KC = 64;  x rows [ROWS][Fp + 4]) | scratch]
// -------------------------------------------------------------------------------------------------
constexpr int DFA_KC = 64;     // staged K-chunk of P2
constexpr int DFA_MAX_FP = 88; // x rows [ROWS][Fp + 4] must fit in the P2 staging area, weights in shared memory
constexpr int DFA_STAGES = ABCD_H / TS_SLAB / 2;   // ring stages = h slabs of one CTA of a pair
constexpr int DFA_MMA_TID = 224;                   // lane 0 of warp 7: copy + MMA thread (warp 7 has no share of the P2 epilogue and of P3a)
constexpr int DFA_RT = 5;                          // P1: rows per thread (rows rgx + 16 j)
__host__ __device__ constexpr size_t dfa_stage_floats() { return (size_t)2 * ROWS * (DFA_KC + 4); }
constexpr int DFA_P2_ONE_ROWS = 40;                // P2 stages all of K = H at once when at most this many rows are active
static_assert((size_t)DFA_P2_ONE_ROWS * (ABCD_H + 4) <= dfa_stage_floats(), "one-chunk P2 staging must fit in the staging area");
static_assert(DFA_STAGES <= TSP_STAGES && DFA_STAGES <= TSP_MAX_SLABS, "ring / tensor-memory geometry of the pair variant");
static_assert((size_t)ROWS * (DFA_MAX_FP + 4) <= dfa_stage_floats(), "x rows must fit in the staging area");
static_assert(2 * ROWS * ((DFA_MAX_FP / 4 + CL - 1) / CL) <= 2 * NTHREADS, "P3b: two lanes per (row, column group) item");
static_assert(NTHREADS == 16 * U && ROWS == 16 * DFA_RT, "P1 thread map: 16 units x 16 row groups");

// One 4-byte store into the shared memory of CTA `cta` of the cluster that also signals 4 bytes on that CTA's mbarrier (same
// offsets as in this CTA): data and notification travel together, so the consumer needs neither a fence nor a cluster barrier.
__device__ __forceinline__ void st_async_cluster_f32(uint32_t local_smem_addr, uint32_t local_bar_addr, unsigned cta, float v) {
    uint32_t remote, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local_smem_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rbar) : "r"(local_bar_addr), "r"(cta));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n"
                 ::"r"(remote), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}

// acc[j][0..3] += x[row rg + 16 j][0, 4 nk4) * W[.., 4 columns of this thread] (row pitch LDW)   for j < NJ (warp-uniform)
__device__ __forceinline__ void st_async_cluster_v4(uint32_t local_smem_addr, uint32_t local_bar_addr, unsigned cta, float a, float b, float c, float d) {
    uint32_t remote, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local_smem_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rbar) : "r"(local_bar_addr), "r"(cta));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                 ::"r"(remote), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(rbar) : "memory");
}

template <int NJ, int LDW>
__device__ __forceinline__ void rows_gemm(float (&acc)[DFA_RT][4], const float* xr, int ldx, const float* wq, int nk4) {
#pragma unroll 2
    for (int kq = 0; kq < nk4; ++kq) {
        const float4 w0 = *reinterpret_cast<const float4*>(wq + (kq * 4 + 0) * LDW);
        const float4 w1 = *reinterpret_cast<const float4*>(wq + (kq * 4 + 1) * LDW);
        const float4 w2 = *reinterpret_cast<const float4*>(wq + (kq * 4 + 2) * LDW);
        const float4 w3 = *reinterpret_cast<const float4*>(wq + (kq * 4 + 3) * LDW);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(xr + j * 16 * ldx + kq * 4);
            acc[j][0] = fmaf(a.x, w0.x, acc[j][0]); acc[j][1] = fmaf(a.x, w0.y, acc[j][1]); acc[j][2] = fmaf(a.x, w0.z, acc[j][2]); acc[j][3] = fmaf(a.x, w0.w, acc[j][3]);
            acc[j][0] = fmaf(a.y, w1.x, acc[j][0]); acc[j][1] = fmaf(a.y, w1.y, acc[j][1]); acc[j][2] = fmaf(a.y, w1.z, acc[j][2]); acc[j][3] = fmaf(a.y, w1.w, acc[j][3]);
            acc[j][0] = fmaf(a.z, w2.x, acc[j][0]); acc[j][1] = fmaf(a.z, w2.y, acc[j][1]); acc[j][2] = fmaf(a.z, w2.z, acc[j][2]); acc[j][3] = fmaf(a.z, w2.w, acc[j][3]);
            acc[j][0] = fmaf(a.w, w3.x, acc[j][0]); acc[j][1] = fmaf(a.w, w3.y, acc[j][1]); acc[j][2] = fmaf(a.w, w3.z, acc[j][2]); acc[j][3] = fmaf(a.w, w3.w, acc[j][3]);
        }
    }
}

template <int CELL>
__global__ void __launch_bounds__(NTHREADS, 1) dec_fwd_ahead_kernel(DecK p) {
    extern __shared__ __align__(1024) float smem_ahead[];
    float* const smem = smem_ahead;
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int LDX = Fp + 4;                   // pitch of the x rows
    constexpr int KC = DFA_KC;
    constexpr int HSP = 2 * U + 4;            // pitch of the CTA's own hidden-layer slice kept in shared memory
    float* W1 = reinterpret_cast<float*>(reinterpret_cast<char*>(smem) + (size_t)DFA_STAGES * TS_STAGE_BYTES);   // [H][2U]
    float* W2s = W1 + H * (2 * U);            // [U][2Fp]: rows of the output layers that multiply this CTA's hidden units
    float* Wx = W2s + U * 2 * Fp;             // [Fp][4U]: input rows of the gate weights of this CTA's units
    float* gpre = Wx + Fp * (4 * U);          // [ROWS][TSP_GP]: h_{t-1} W_hh of this CTA's gate columns
    float* stage_base = gpre + ROWS * TSP_GP;
    float* red = stage_base + dfa_stage_floats();   // 64 floats
    const Stage st = make_stage(stage_base, KC);
    const Stage st1 = make_stage(stage_base, ABCD_H);   // one-chunk staging of P2 (only buf0, at most DFA_P2_ONE_ROWS rows)
    float* Hs = st.buf0;                      // [ROWS][HSP] (the staging buffers are idle between phases)
    float* xs = stage_base;                   // [ROWS][LDX]: x_{t-1}, written by the column owners of the whole cluster during their P3b
    __shared__ uint64_t tsp_bars[TSP_STAGES + 1];
    __shared__ uint64_t xbar;                 // bytes of x_t pushed into xs by the column owners (one phase per step)
    __shared__ uint32_t tsp_tmem_slot;
    TsPairCtx tsp;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) umma::mbar_init(&xbar, 1);  // (made visible cluster-wide by the fence + barriers of ts_pair_init / the prologue)
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;                  // the 16 units this CTA finalises (cell update, MLP slices)
    const int pair_u0 = (rank >> 1) * 2 * U;  // the 32 units whose gate columns the pair contracts
    const int ng_mine = (Fp / 4 - rank + CL - 1) / CL;   // x column groups (4 columns) owned by this CTA: rank, rank + 16, ...

    // ---- resident weights ----
    if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
    {
        auto wval = [&](int col, int k) -> float {   // recurrent weights only: k < H
            const int u = pair_u0 + (col >> 2), s = col & 3;
            const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            return gate >= 0 ? __ldg(p.w_hh + (long)(gate * H + u) * H + k) : 0.f;
        };
        ts_pair_init(tsp, reinterpret_cast<char*>(smem), tsp_bars, &tsp_tmem_slot, rank & 1, H / TS_SLAB, wval, tid);
    }
    char* himg[2];
    himg[0] = p.himg + (size_t)(tile * 2) * ts_image_bytes(ABCD_H); himg[1] = himg[0] + ts_image_bytes(ABCD_H);
    {   // clear the images once (rows that never become active are still fetched in 32-row chunks: keep them finite)
        float4* z = reinterpret_cast<float4*>(himg[0]);
        const size_t n16 = 2 * ts_image_bytes(ABCD_H) / 16;
        for (size_t i = (size_t)rank * NTHREADS + tid; i < n16; i += (size_t)CL * NTHREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ts_image_publish();
    cluster_sync_all();
    for (int idx = tid; idx < H * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);
        int which = col / U, u = col % U;
        W1[idx] = p.em_w1[which][(long)(u0 + u) * H + k];
    }
    for (int idx = tid; idx < U * 2 * Fp; idx += NTHREADS) {
        int c = idx % (2 * Fp), kk = idx / (2 * Fp);
        int which = c >= Fp, cc = which ? c - Fp : c;     // mean columns [0,Fp), log-variance columns [Fp,2Fp)
        W2s[idx] = cc < F ? p.em_w2[which][(long)cc * M + u0 + kk] : 0.f;
    }
    for (int idx = tid; idx < Fp * 4 * U; idx += NTHREADS) {
        const int col = idx % (4 * U), k = idx / (4 * U);
        const int u = u0 + (col >> 2), s = col & 3;
        const int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
        Wx[idx] = (gate >= 0 && k < F) ? p.w_ih[(long)(gate * H + u) * F + k] : 0.f;
    }

    // ---- per-thread constants for the three phases ----
    const int cgx = tid & (U - 1), rgx = tid >> 4;   // P1: unit, row group (rows rgx + 16 j, j < 5)
    const int unit = u0 + cgx;
    float bias1[4];
    if (CELL == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) bias1[g] = p.b_ih[g * H + unit] + p.b_hh[g * H + unit];
    } else {
        bias1[0] = p.b_ih[unit] + p.b_hh[unit];
        bias1[1] = p.b_ih[H + unit] + p.b_hh[H + unit];
        bias1[2] = p.b_ih[2 * H + unit];
        bias1[3] = p.b_hh[2 * H + unit];
    }
    const int KS2 = pick_ks_fit(2 * U / 4, 2 * U, KC);
    const ThreadMap tm2 = make_map(tid, 2 * U / 4, KS2);
    float bias2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int col = tm2.cg * 4 + j;
        bias2[j] = p.em_b1[col / U][u0 + col % U];
    }
    const int ncg3 = 2 * Fp / 4;                                      // column groups of the local emission GEMM
    const int NG = Fp / 4;                                            // x column groups
    // partial emission parameters, laid out for writer and reader alike: [source CTA][x column group][mean | log-variance][row][4]
    float* Pmine = p.Pbuf + (long)blockIdx.x * ROWS * (2 * Fp);
    const float* Ptile = p.Pbuf + (long)tile * CL * ROWS * (2 * Fp);
    // pair exchange of the gate partial sums (the backward pass's partial buffer is idle during the forward pass)
    float* pair_send = p.Pb2 + (size_t)(blockIdx.x ^ 1) * ROWS * H;
    const float* pair_recv = p.Pb2 + (size_t)blockIdx.x * ROWS * H;
    const uint32_t xs_addr = umma::smem_u32(xs), xbar_addr = umma::smem_u32(&xbar);
    // P3b item order: the four warps that do NOT own accumulator lanes of the pair product come first, so that at typical row
    // counts the owners start adding the partner's half of gpre (an L2 round trip) while the others sum the emission partials
    const bool recv_warp = (((warp & 3) >> 1) == (rank & 1));
    const int p3b_tid = ((recv_warp ? 4 : 0) + (((warp >> 2) << 1) | (warp & 1))) * 32 + (tid & 31);

    // recurrent state of (row, unit) pairs owned by this thread
    float creg[DFA_RT], hreg[DFA_RT];
#pragma unroll
    for (int j = 0; j < DFA_RT; ++j) {
        int b = tile + p.n_tiles * (rgx + 16 * j);
        creg[j] = 0.f; hreg[j] = 0.f;
        if (b < p.B) {
            hreg[j] = p.Hall[(long)b * H + unit];
            if (CELL == 0) creg[j] = p.Call[(long)b * H + unit];
        }
    }
    const int Tt = p.lengths[tile];
    float nll = 0.f;
    const int n_tiles = p.n_tiles, B = p.B;
    const float* Hall = p.Hall;
    // image of the initial state and its recurrent product (gpre of step 0)
#pragma unroll
    for (int j = 0; j < DFA_RT; ++j) {
        const int r = rgx + 16 * j;
        if (tile + p.n_tiles * r < p.B) ts_image_store(himg[1], unit, r, hreg[j]);
    }
    ts_image_publish();
    cluster_sync_all();
    if (Tt > 0) {
        const int na0 = active_rows(p.batch_sizes[0], tile, n_tiles);
        if (tid == DFA_MMA_TID) { ts_pair_fetch(tsp, himg[1], na0); ts_pair_issue(tsp, na0); }
        __syncwarp();
        ts_pair_drain_send(tsp, gpre, pair_send, na0, tid);
        cluster_sync_all();
        ts_pair_drain_recv(tsp, gpre, pair_recv, na0, tid);
    }
    __syncthreads();
    const bool timing = p.timing == 1 && blockIdx.x == p.timing_cta && tid == 64;
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
    unsigned long long issue_cycles = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }

    for (int t = 0; t < Tt; ++t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const int na_next = t + 1 < Tt ? active_rows(p.batch_sizes[t + 1], tile, n_tiles) : 0;
        const long off_t = p.offsets[t];

        // ---- P1: gates + cell ----
        {
            float acc[DFA_RT][4];
#pragma unroll
            for (int j = 0; j < DFA_RT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
            if (t > 0) {
                umma::mbar_wait(&xbar, (uint32_t)(t - 1) & 1u);   // C: x_{t-1} of every column owner has landed in xs
                __syncthreads();          // gpre was completed by other threads of this CTA (partner's half) after barrier B
                DEC_TICK(0)
                // rows of this warp: 2 row groups; the trip count is warp-uniform (rows >= na are computed on stale data and dropped).
                // (Tried: one gate column per lane with broadcast activations, and 10-row tiles on four warps adding into gpre in
                // place -- 4.7 k and 3.9 k cycles against 3.8 k for this form.)
                const int njw = min(DFA_RT, (na - 2 * warp + 15) >> 4);
                const float* xr = xs + rgx * LDX;
                const float* wq = Wx + cgx * 4;
                switch (njw) {
                    case 1: rows_gemm<1, 4 * U>(acc, xr, LDX, wq, Fp >> 2); break;
                    case 2: rows_gemm<2, 4 * U>(acc, xr, LDX, wq, Fp >> 2); break;
                    case 3: rows_gemm<3, 4 * U>(acc, xr, LDX, wq, Fp >> 2); break;
                    case 4: rows_gemm<4, 4 * U>(acc, xr, LDX, wq, Fp >> 2); break;
                    case 5: rows_gemm<5, 4 * U>(acc, xr, LDX, wq, Fp >> 2); break;
                    default: break;
                }
            }
            DEC_TICK(1)
            float4 gsv[DFA_RT];
            float csv[DFA_RT];
#pragma unroll
            for (int j = 0; j < DFA_RT; ++j) {
                const int r = rgx + 16 * j;
                if (r >= na) continue;
                const long frame = off_t + tile + (long)n_tiles * r;
                const float4 hp = *reinterpret_cast<const float4*>(gpre + r * TSP_GP + cgx * 4);
                const float a0 = acc[j][0] + hp.x, a1 = acc[j][1] + hp.y, a2 = acc[j][2] + hp.z, a3 = acc[j][3] + hp.w;
                float h;
                if (CELL == 0) {
                    float ig = sigmoid_acc(a0 + bias1[0]);
                    float fg = sigmoid_acc(a1 + bias1[1]);
                    float gg = tanhf(a2 + bias1[2]);
                    float og = sigmoid_acc(a3 + bias1[3]);
                    float c = fg * creg[j] + ig * gg;
                    h = og * tanhf(c);
                    creg[j] = c;
                    gsv[j] = make_float4(ig, fg, gg, og);
                    csv[j] = c;
                } else {
                    float rg_ = sigmoid_acc(a0 + bias1[0]);
                    float zg = sigmoid_acc(a1 + bias1[1]);
                    float hn = a3 + bias1[3];
                    float ng = tanhf(a2 + bias1[2] + rg_ * hn);
                    h = (1.f - zg) * ng + zg * hreg[j];
                    gsv[j] = make_float4(rg_, zg, ng, hn);
                }
                hreg[j] = h;
                p.Hall[((long)B + frame) * H + unit] = h;
                // (a transposed copy in shared memory + 16-byte image stores, 4x fewer sectors, measured slower: 1.23 vs 1.17 ms)
                if (na_next > 0) ts_image_store(himg[t & 1], unit, r, h);
            }
            ts_image_publish();
            DEC_TICK(2)
            cluster_arrive_release();   // A: h_t on its way to the other CTAs
            DEC_TICK(11)
            // saved for the backward pass only: stored under the barrier
#pragma unroll
            for (int j = 0; j < DFA_RT; ++j) {
                const int r = rgx + 16 * j;
                if (r >= na) continue;
                const long frame = off_t + tile + (long)n_tiles * r;
                *reinterpret_cast<float4*>(p.Gs + frame * (4 * H) + unit * 4) = gsv[j];
                if (CELL == 0) p.Call[((long)B + frame) * H + unit] = csv[j];
            }
            cluster_wait_acquire();
        }
        DEC_TICK(3)
        if (tid == DFA_MMA_TID && na_next > 0) ts_pair_fetch(tsp, himg[t & 1], na_next);

        // ---- P2: hidden layers of the two emission MLPs ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rph = [=](int r) -> const float* { return Hall + ((long)B + off_t + tile + (long)n_tiles * r) * ABCD_H; };
            // few active rows (most steps): all of h_t[rows, H] fits in the staging area -> one L2 round trip, one CTA barrier
            if (na <= DFA_P2_ONE_ROWS) phase_gemm<NTHREADS, ABCD_H, 2 * U>(acc, st1, H, rph, 0, NullRow(), na, W1, tm2, KS2, tid);
            else phase_gemm<NTHREADS, KC, 2 * U>(acc, st, H, rph, 0, NullRow(), na, W1, tm2, KS2, tid);
            DEC_TICK(4)
            if (warp == 7) {
                // (warp 7 holds a K-split share of P2 only: no epilogue rows, no P3a items)
                if (tid == DFA_MMA_TID && na_next > 0) {
                    const long long i0 = p.timing == 1 ? clock64() : 0;
                    ts_pair_issue(tsp, na_next);
                    if (p.timing == 1) issue_cycles += clock64() - i0;
                }
                __syncwarp();
            } else {
                if (tm2.active && tm2.ks == 0) {
#pragma unroll
                    for (int i = 0; i < RT; ++i) {
                        int r = tm2.rg + i * RG;
                        if (r >= na) continue;
                        long frame = off_t + tile + (long)n_tiles * r;
                        int col = tm2.cg * 4;
                        int which = col / U, uu = col % U;
                        float4 v = make_float4(tanhf(acc[i][0] + bias2[0]), tanhf(acc[i][1] + bias2[1]),
                                               tanhf(acc[i][2] + bias2[2]), tanhf(acc[i][3] + bias2[3]));
                        *reinterpret_cast<float4*>(p.Hid + frame * (2 * M) + which * M + u0 + uu) = v;
                        *reinterpret_cast<float4*>(Hs + r * HSP + col) = v;
                    }
                }
                asm volatile("bar.sync 1, 224;\n" ::: "memory");   // warps 0..6: the hidden slice is complete
                // ---- P3a: this CTA's share of the emission parameters (K-split over the hidden units): out[rows][2Fp] =
                // Hs[rows][U of the head] * W2s[U][2Fp], lane = output column (warps 0..4 cover the 2Fp <= 160 columns; weights in
                // registers, activations as warp-wide broadcasts) ----
                const int colp = warp * 32 + (tid & 31);
                if (colp < 2 * Fp) {
                    const int which = colp >= Fp, cc = colp - which * Fp;
                    float wreg[U];
#pragma unroll
                    for (int kk = 0; kk < U; ++kk) wreg[kk] = W2s[kk * (2 * Fp) + colp];
                    float* dst = Pmine + (((long)(cc >> 2) * 2 + which) * ROWS) * 4 + (cc & 3);
                    const float* hrow = Hs + which * U;
#pragma unroll 2
                    for (int r = 0; r < na; ++r) {
                        float o = 0.f;
#pragma unroll
                        for (int kq = 0; kq < U / 4; ++kq) {
                            const float4 av = *reinterpret_cast<const float4*>(hrow + r * HSP + kq * 4);
                            o = fmaf(av.x, wreg[kq * 4 + 0], o); o = fmaf(av.y, wreg[kq * 4 + 1], o);
                            o = fmaf(av.z, wreg[kq * 4 + 2], o); o = fmaf(av.w, wreg[kq * 4 + 3], o);
                        }
                        dst[(long)r * 4] = o;
                    }
                }
            }
        }
        DEC_TICK(5)
        // recurrent product of the NEXT step: own gate columns -> gpre, the partner's -> its receive buffer
        if (na_next > 0) ts_pair_drain_send(tsp, gpre, pair_send, na_next, tid);
        DEC_TICK(6)
        cluster_arrive_release();   // B: partial emission parameters and partner partials visible; every CTA is done with its staging area
        DEC_TICK(12)
        cluster_wait_acquire();
        DEC_TICK(7)
        // x_t arrives as 4-byte pushes from the owners of its columns: na_next rows x Fp columns
        if (tid == 0 && na_next > 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(xbar_addr), "r"((uint32_t)(na_next * Fp * 4)) : "memory");
        }

        // ---- P3b: emission parameters of this CTA's x column groups (group g of 4 columns belongs to rank g % 16), reparameterised
        // sample, NLL; x_t pushed to every CTA of the cluster, 16 bytes per store.  Two lanes per (row, group) item: the even lane
        // sums the 16 partial means, the odd lane the 16 partial log-variances (one L2 round trip, fixed order), then they swap ----
        {
            const int n_items = na * ng_mine;
            for (int vt = p3b_tid; (vt & ~31) < 2 * n_items; vt += NTHREADS) {   // (warp-uniform: the lanes exchange their sums with shuffles)
                const int item = vt >> 1, which = vt & 1;
                const bool valid = item < n_items;
                const int r = valid ? item / ng_mine : 0, c0 = 4 * (rank + CL * (valid ? item % ng_mine : 0));
                const long frame = off_t + tile + (long)n_tiles * r;
                // noise / ground truth of this lane's two columns come from HBM: issued first, their latency hides behind the partial sums
                const int cj = c0 + 2 * which;
                float ev[2], gv[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    ev[j] = (valid && cj + j < F) ? __ldg(p.eps + frame * p.ld_eps + cj + j) : 0.f;
                    gv[j] = (valid && p.gt && cj + j < F) ? __ldg(p.gt + frame * p.ld_gt + cj + j) : 0.f;
                }
                float4 pp[CL];
                const float* src0 = Ptile + (((long)(c0 >> 2) * 2 + which) * ROWS + r) * 4;
#pragma unroll
                for (int src = 0; src < CL; ++src) pp[src] = valid ? ld_cg4(src0 + (long)src * ROWS * (2 * Fp)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 acc4;
                acc4.x = c0 + 0 < F ? __ldg(p.em_b2[which] + c0 + 0) : 0.f; acc4.y = c0 + 1 < F ? __ldg(p.em_b2[which] + c0 + 1) : 0.f;
                acc4.z = c0 + 2 < F ? __ldg(p.em_b2[which] + c0 + 2) : 0.f; acc4.w = c0 + 3 < F ? __ldg(p.em_b2[which] + c0 + 3) : 0.f;
#pragma unroll
                for (int src = 0; src < CL; ++src) { acc4.x += pp[src].x; acc4.y += pp[src].y; acc4.z += pp[src].z; acc4.w += pp[src].w; }
                if (c0 + 0 >= F) acc4.x = 0.f;   // padding columns carry zeros
                if (c0 + 1 >= F) acc4.y = 0.f;
                if (c0 + 2 >= F) acc4.z = 0.f;
                if (c0 + 3 >= F) acc4.w = 0.f;
                float4 oth;
                oth.x = __shfl_xor_sync(0xffffffffu, acc4.x, 1); oth.y = __shfl_xor_sync(0xffffffffu, acc4.y, 1);
                oth.z = __shfl_xor_sync(0xffffffffu, acc4.z, 1); oth.w = __shfl_xor_sync(0xffffffffu, acc4.w, 1);
                const float4 mean4 = which ? oth : acc4, lv4 = which ? acc4 : oth;
                // this lane finalises columns cj, cj + 1; the sample quad is completed by one more swap
                const float m0 = which ? mean4.z : mean4.x, m1 = which ? mean4.w : mean4.y;
                const float l0 = which ? lv4.z : lv4.x, l1 = which ? lv4.w : lv4.y;
                float x0 = 0.f, x1 = 0.f;
                if (cj < F) {
                    x0 = m0 + expf(0.5f * l0) * ev[0];
                    if (p.gt && valid) { const float d = gv[0] - m0; nll += 0.5f * (LOG_2PI + l0 + d * expf(-l0) * d); }
                }
                if (cj + 1 < F) {
                    x1 = m1 + expf(0.5f * l1) * ev[1];
                    if (p.gt && valid) { const float d = gv[1] - m1; nll += 0.5f * (LOG_2PI + l1 + d * expf(-l1) * d); }
                }
                const float y0 = __shfl_xor_sync(0xffffffffu, x0, 1), y1 = __shfl_xor_sync(0xffffffffu, x1, 1);
                const float4 x4 = which ? make_float4(y0, y1, x0, x1) : make_float4(x0, x1, y0, y1);
                DEC_TICK(13)
                if (valid && r < na_next) {
                    const uint32_t a = xs_addr + (uint32_t)(r * LDX + c0) * 4u;
#pragma unroll
                    for (unsigned q = 0; q < (unsigned)CL / 2; ++q) st_async_cluster_v4(a, xbar_addr, 8u * which + q, x4.x, x4.y, x4.z, x4.w);
                }
                if (valid) {   // saved for the backward pass
                    *reinterpret_cast<float4*>((which ? p.Lv : p.Mean) + frame * Fp + c0) = which ? lv4 : mean4;
                    if (which == 0) *reinterpret_cast<float4*>(p.Xall + ((long)B + frame) * Fp + c0) = x4;
                }
            }
        }
        DEC_TICK(8)
        // off the critical path (the next step waits for x_t on its mbarrier; WAR hazards on the exchange buffers are covered by
        // barriers A and B of the next step): partner's half of the next gate pre-activations
        if (na_next > 0) ts_pair_drain_recv(tsp, gpre, pair_recv, na_next, tid);
        DEC_TICK(9)
    }
#undef DEC_TICK
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) if (i != 10) g_dec_timing[i] = tacc[i]; }
    if (p.timing == 1 && blockIdx.x == p.timing_cta && tid == DFA_MMA_TID) g_dec_timing[10] = issue_cycles;
    float tot = block_sum(nll, red);
    if (tid == 0) p.nll_partial[blockIdx.x] = tot;
    ts_pair_fini(tsp, tid);
}

// -------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------
// Backward.  Every contraction over a dimension that is partitioned over the cluster (the hidden units of the MLPs,
// the gate columns of the cell) is computed as a partial product on the CTA's OWN slice -- which it has just produced
// and still holds in shared memory -- followed by an in-cluster reduction of the partials through L2:
//   B3  dHidPre[:, my 2U]   = (dML[:, 2Fp] W2^T) * (1 - Hid^2)        staged GEMM (K = 2Fp), slice kept in smem
//   B2' partial dh[:, H]    = dHidPre[:, my 2U] W1[my 2U, H]          local GEMM -> Pb2        -- cluster barrier --
//       dh[:, my U]         = sum of the CL partials + recurrent carry + offset-head gradient; gate gradients dG
//   B1' partial [dh_prev|dx] = dG[:, my 4U] [W_hh | W_ih][my 4U, H+Fp] local GEMM -> Pb1        -- cluster barrier --
//       dh carry[:, my U], dx[:, my x columns] = sums of the partials; dx is folded into dML of the previous frame
//                                                                                              -- cluster barrier --
template <int CELL, int KC>
__global__ void __launch_bounds__(NTHREADS, 1) dec_bwd_kernel(DecK p) {
    extern __shared__ __align__(16) float smem[];
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int XC = xcols_per_cta(F);
    const int N1 = H + Fp;                    // columns of the B1' partial: [dh_prev (H) | dx (Fp)]
    constexpr int DSP = 2 * U + 4;            // pitch of the dHidPre slice in shared memory
    constexpr int GSP = 4 * U + 4;            // pitch of the dG slice in shared memory
    float* WB3 = smem;                        // [2Fp][2U]
    float* W1s = WB3 + 2 * Fp * (2 * U);      // [2U][H]: rows of the MLP input layers of this CTA's hidden units
    float* Wb1 = W1s + 2 * U * H;             // [4U][DEC_N1P]: rows of [W_hh | W_ih] of this CTA's gate columns
    float* dh_sm = Wb1 + 4 * U * DEC_N1P;     // [ROWS][U]
    float* dc_sm = dh_sm + ROWS * U;          // [ROWS][U]
    const Stage st = make_stage(dc_sm + ROWS * U, KC);
    float* Ds = st.buf0;                      // [ROWS][DSP]  (staging is idle outside B3)
    float* Gsm = st.buf0;                     // [ROWS][GSP]  (Ds is dead by then; may run into buf1)

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;
    const int xc0 = rank * XC;
    const int xc1 = min(F, xc0 + XC);

    for (int idx = tid; idx < 2 * Fp * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);
        int which = col / U, u = col % U;
        int kk = which == 0 ? k : k - Fp;
        float v = 0.f;
        if (kk >= 0 && kk < F) v = p.em_w2[which][(long)kk * M + u0 + u];
        WB3[idx] = v;
    }
    for (int idx = tid; idx < 2 * U * H; idx += NTHREADS) {
        int j = idx % H, k = idx / H;
        W1s[idx] = p.em_w1[k / U][(long)(u0 + k % U) * H + j];
    }
    for (int idx = tid; idx < 4 * U * DEC_N1P; idx += NTHREADS) {
        int j = idx % DEC_N1P, kphys = idx / DEC_N1P;
        int uu = u0 + (kphys >> 2), s = kphys & 3;
        float v = 0.f;
        if (j < H) {
            int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            if (gate >= 0) v = p.w_hh[(long)(gate * H + uu) * H + j];
        } else {
            int c = j - H;
            int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            if (gate >= 0 && c < F) v = p.w_ih[(long)(gate * H + uu) * F + c];
        }
        Wb1[idx] = v;
    }
    for (int idx = tid; idx < ROWS * U; idx += NTHREADS) { dh_sm[idx] = 0.f; dc_sm[idx] = 0.f; }

    const int KS3 = pick_ks_fit(2 * U / 4, 2 * U, KC);
    const ThreadMap tm3 = make_map(tid, 2 * U / 4, KS3);
    constexpr int PRG = NTHREADS / U;   // pointwise mapping: 16 row groups x 16 units, rows prg + i*PRG
    constexpr int PR = ROWS / PRG;      // rows per pointwise thread
    const int pu = tid % U, prg = tid / U;
    const int unit = u0 + pu;

    const int Tt = p.lengths[tile];
    const int n_tiles = p.n_tiles, B = p.B;
    const float* dML = p.dML;
    float* Pb2 = p.Pb2 + (long)blockIdx.x * ROWS * H;
    const float* Pb2t = p.Pb2 + (long)tile * CL * ROWS * H;
    float* Pb1 = p.Pb1 + (long)blockIdx.x * ROWS * N1;
    const float* Pb1t = p.Pb1 + (long)tile * CL * ROWS * N1;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of sum(dG) over rows and steps
    float4 hsum = make_float4(0.f, 0.f, 0.f, 0.f);   // ... and of sum(dHidPre) (bias gradients of the MLP input layers)
    __syncthreads();
    const bool timing = p.timing == 3 && blockIdx.x == 0 && tid == 0;
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }

    for (int t = Tt - 1; t >= 0; --t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const long off_t = p.offsets[t];
        const long prev_base = t == 0 ? 0 : (long)B + p.offsets[t - 1];

        // saved forward activations for the gate-gradient step: fetched now, consumed after the first barrier
        float4 g4v[PR];
        float cv[PR], pv[PR], hov[PR];
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            g4v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[i] = pv[i] = hov[i] = 0.f;
            if (r >= na) continue;
            long b = tile + (long)n_tiles * r;
            long frame = off_t + b;
            long prow = prev_base + b;
            g4v[i] = __ldg(reinterpret_cast<const float4*>(p.Gs + frame * (4 * H) + unit * 4));
            hov[i] = __ldg(p.dHoff + frame * H + unit);
            if (CELL == 0) {
                cv[i] = __ldg(p.Call + ((long)B + frame) * H + unit);
                pv[i] = __ldg(p.Call + prow * H + unit);
            } else {
                pv[i] = __ldg(p.Hall + prow * H + unit);
            }
        }

        // ---- B3: gradient w.r.t. this CTA's MLP hidden pre-activations ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rp = [=](int r) -> const float* { return dML + (off_t + tile + (long)n_tiles * r) * (2 * Fp); };
            // tanh outputs of this thread's tile: fetched before the GEMM so the latency hides behind it
            float4 hdv[RT];
            if (tm3.active && tm3.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm3.rg + i * RG;
                    hdv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm3.cg * 4;
                    hdv[i] = __ldg(reinterpret_cast<const float4*>(p.Hid + frame * (2 * M) + (col / U) * M + u0 + col % U));
                }
            }
            // the saved-activation loads above do not depend on other CTAs: they overlap the barrier that ended the
            // previous step (whose dML updates feed the contraction below)
            if (t < Tt - 1) cluster_wait_acquire();
            DEC_TICK(0)
            phase_gemm<NTHREADS, KC, 2 * U>(acc, st, 2 * Fp, rp, 0, NullRow(), na, WB3, tm3, KS3, tid);
            DEC_TICK(1)
            if (tm3.active && tm3.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm3.rg + i * RG;
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm3.cg * 4;
                    int which = col / U, uu = col % U;
                    long o = frame * (2 * M) + which * M + u0 + uu;
                    float4 hd = hdv[i];
                    float4 v = make_float4(acc[i][0] * (1.f - hd.x * hd.x), acc[i][1] * (1.f - hd.y * hd.y),
                                           acc[i][2] * (1.f - hd.z * hd.z), acc[i][3] * (1.f - hd.w * hd.w));
                    *reinterpret_cast<float4*>(p.dHidPre + o) = v;
                    *reinterpret_cast<float4*>(Ds + r * DSP + col) = v;
                    hsum.x += v.x; hsum.y += v.y; hsum.z += v.z; hsum.w += v.w;
                }
            }
            __syncthreads();
        }
        // ---- B2': partial dh over all H units from this CTA's hidden slice ----
        DEC_TICK(2)
        local_gemm_store<2 * U / 4, DSP, ABCD_H>(Ds, W1s, H / 4, na, Pb2, H, tid);
        DEC_TICK(3)
        cluster_sync_all();
        DEC_TICK(4)

        // ---- dh of this CTA's units = sum of the partials (+ carry), then the gate gradients ----
        for (int idx = tid; idx < na * (U / 4); idx += NTHREADS) {
            const int r = idx / (U / 4), q = idx % (U / 4);
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb2t + ((long)src * ROWS + r) * H + u0 + q * 4);
            float4 a = *reinterpret_cast<float4*>(dh_sm + r * U + q * 4);
#pragma unroll
            for (int src = 0; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            *reinterpret_cast<float4*>(dh_sm + r * U + q * 4) = a;
        }
        __syncthreads();
        DEC_TICK(5)
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            if (r >= na) continue;
            long frame = off_t + tile + (long)n_tiles * r;
            float dh = dh_sm[r * U + pu] + hov[i];
            float carry = dc_sm[r * U + pu];
            const float4 g4 = g4v[i];
            float4 dgv;
            if (CELL == 0) {
                float c = cv[i];
                float c_prev = pv[i];
                float tc = tanhf(c);
                float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
                float dc = carry + dh * og * (1.f - tc * tc);
                dgv.x = dc * gg * ig * (1.f - ig);
                dgv.y = dc * c_prev * fg * (1.f - fg);
                dgv.z = dc * ig * (1.f - gg * gg);
                dgv.w = dh * tc * og * (1.f - og);
                dc_sm[r * U + pu] = dc * fg;
            } else {
                dh += carry;
                float h_prev = pv[i];
                float rg_ = g4.x, zg = g4.y, ng = g4.z, hn = g4.w;
                float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
                dgv.x = dn_pre * hn * rg_ * (1.f - rg_);
                dgv.y = dh * (h_prev - ng) * zg * (1.f - zg);
                dgv.z = dn_pre;
                dgv.w = dn_pre * rg_;
                dc_sm[r * U + pu] = dh * zg;
            }
            *reinterpret_cast<float4*>(p.dG + frame * (4 * H) + unit * 4) = dgv;
            *reinterpret_cast<float4*>(Gsm + r * GSP + pu * 4) = dgv;
            bsum.x += dgv.x; bsum.y += dgv.y; bsum.z += dgv.z; bsum.w += dgv.w;
        }
        __syncthreads();

        // ---- B1': partial [dh_{t-1} | dx_t] from this CTA's gate columns ----
        DEC_TICK(6)
        local_gemm_store<U, GSP, DEC_N1P>(Gsm, Wb1, N1 / 4, na, Pb1, N1, tid);
        DEC_TICK(7)
        cluster_sync_all();
        DEC_TICK(8)

        // ---- recurrent carry for this CTA's units, dx for its x columns (folded into dML of the previous frame) ----
        // x columns first (longest dependency chain: partials + saved activations + read-modify-write), one column per
        // thread; then the dh quads.  All loads of an item are issued before its first use.
        if (t > 0) {
            const long pbase = p.offsets[t - 1] + tile;
            for (int idx = tid; idx < na * XC; idx += NTHREADS) {
                const int r = idx / XC, c = xc0 + idx % XC;
                if (c >= xc1) continue;
                const long pf = pbase + (long)n_tiles * r;   // frame that produced x_t
                float v[CL];
#pragma unroll
                for (int src = 0; src < CL; ++src) v[src] = ld_cg(Pb1t + ((long)src * ROWS + r) * N1 + H + c);
                const float lv = __ldg(p.Lv + pf * Fp + c);
                const float e = __ldg(p.eps + pf * p.ld_eps + c);
                float* dm = p.dML + pf * (2 * Fp) + c;
                const float m0 = dm[0], m1 = dm[Fp];
                float dx = v[0];
#pragma unroll
                for (int src = 1; src < CL; ++src) dx += v[src];
                dm[0] = m0 + dx;
                dm[Fp] = m1 + dx * 0.5f * expf(0.5f * lv) * e;
            }
        }
        for (int idx = tid; idx < na * (U / 4); idx += NTHREADS) {
            const int r = idx / (U / 4), q = idx % (U / 4);
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb1t + ((long)src * ROWS + r) * N1 + u0 + q * 4);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            *reinterpret_cast<float4*>(dh_sm + r * U + q * 4) = a;
        }
        DEC_TICK(9)
        cluster_arrive_release();   // matched by the wait after the next step's prefetch (or the one after the loop)
        DEC_TICK(10)
    }
#undef DEC_TICK
    if (Tt > 0) cluster_wait_acquire();
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) g_dec_timing[i] = tacc[i]; }
    // gradient w.r.t. the initial state produced by feature2hidden (h0 = even, c0 = odd columns; model.py:100)
    for (int i = 0; i < PR; ++i) {
        int r = prg + i * PRG;
        long b = tile + (long)n_tiles * r;
        if (b >= B || r >= ROWS) continue;
        if (CELL == 0) {
            p.dHC0[b * (2 * H) + 2 * unit] = dh_sm[r * U + pu];
            p.dHC0[b * (2 * H) + 2 * unit + 1] = dc_sm[r * U + pu];
        } else {
            p.dHC0[b * H + unit] = dh_sm[r * U + pu] + dc_sm[r * U + pu];
        }
    }
    // recurrent bias gradients: fixed-order reduction over the row groups, one row of partial sums per tile
    {
        __syncthreads();
        float4* red = reinterpret_cast<float4*>(st.buf0);
        red[prg * U + pu] = bsum;
        __syncthreads();
        if (tid < U) {
            float4 a = red[tid];
            for (int g = 1; g < PRG; ++g) {
                const float4 b = red[g * U + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(p.dbias + (long)tile * (4 * H) + (u0 + tid) * 4) = a;
        }
        // MLP input-layer bias gradients: threads of the B3 epilogue hold sums for (4 columns, their row group)
        __syncthreads();
        if (tm3.active && tm3.ks == 0) red[tm3.rg * (2 * U / 4) + tm3.cg] = hsum;
        __syncthreads();
        if (tid < 2 * U / 4) {
            float4 a = red[tid];
            for (int g = 1; g < NRG; ++g) {
                const float4 b = red[g * (2 * U / 4) + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const int col = tid * 4;   // column of this CTA's slice: [which][u0 + col % U]
            *reinterpret_cast<float4*>(p.dbias1 + (long)tile * (2 * M) + (col / U) * M + u0 + col % U) = a;
        }
    }
}

// geometry of the tensor-core backward kernel
constexpr int DBT_M1 = 384;                                          // padded columns of the B1' result (3 blocks of 128)
constexpr int DBT_DCOLS = 80;                                        // accumulator columns per block (N <= ROWS)
constexpr int DBT_A1_COL = 256, DBT_A2_COL = 256 + 3 * 4 * U;        // tensor-memory columns of the resident weights
constexpr int DBT_WLO1_BYTES = (4 * U / TS_SLAB) * (DBT_M1 / 32) * TS_CHUNK;   // 96 KB
constexpr int DBT_WLO2_BYTES = (ABCD_H / 32) * TS_CHUNK;             // 32 KB (2U = 32 = one slab)
static_assert(3 * DBT_DCOLS <= 256 && DBT_A2_COL + 2 * 2 * U <= 512, "tensor-memory budget");
static_assert(2 * U == TS_SLAB && 4 * U == 2 * TS_SLAB, "slab geometry of the K-split products");

// Tensor-core version (F <= 128): the two K-split products run transposed on tcgen05 with the weight rows resident in
// tensor memory as the A operand (TS form, rnn_ts.cuh):
//   B2'  D^T[H units, rows]        = W1^T[H, 2U]            * dHidPre^T[2U, rows]      (2 blocks of 128 lanes, K = 32)
//   B1'  D^T[H + Fp columns, rows] = [W_hh | W_ih]^T[.., 4U] * dG^T[4U, rows]           (3 blocks, K = 64 in two halves)
// Tensor memory: accumulators 3 x 80 columns, W_hi of B1' 3 x 64 columns, W_hi of B2' 2 x 32 columns (496 of 512).  The low
// parts of the weights live in shared memory as MN-major tiles; the CTA's own dHidPre / dG slices are re-laid from shared
// memory into hi/lo B-operand tiles (4x4 register transposes).  B3 stays a staged FFMA GEMM.
template <int CELL, int KC>
__global__ void __launch_bounds__(NTHREADS, 1) dec_bwd_ts_kernel(DecK p) {
    // (a distinct symbol: the alignment of an extern shared array is a property of the symbol, and `smem` is declared
    // 16-byte aligned by the FFMA kernels of this file)
    extern __shared__ __align__(1024) float smem_ts[];
    float* const smem = smem_ts;
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int XC = xcols_per_cta(F);
    const int N1 = H + Fp;                    // columns of the B1' partial: [dh_prev (H) | dx (Fp)]
    constexpr int DSP = 2 * U + 4;            // pitch of the dHidPre slice in shared memory
    constexpr int GSP = 4 * U + 4;            // pitch of the dG slice in shared memory
    char* b_hi = reinterpret_cast<char*>(smem);                     // one 32-k slab of [k][rows] operand tiles
    char* b_lo = b_hi + TS_B_BYTES;
    char* wlo1 = b_lo + TS_B_BYTES;                                 // B1': [2 slabs][12 chunks (384 columns)][4 KB]
    char* wlo2 = wlo1 + DBT_WLO1_BYTES;                             // B2': [1 slab][8 chunks (256 units)][4 KB]
    float* WB3 = reinterpret_cast<float*>(wlo2 + DBT_WLO2_BYTES);   // [2Fp][2U]
    float* dh_sm = WB3 + 2 * Fp * (2 * U);    // [ROWS][U]
    float* dc_sm = dh_sm + ROWS * U;          // [ROWS][U]
    const Stage st = make_stage(dc_sm + ROWS * U, KC);
    float* Ds = st.buf0;                      // [ROWS][DSP]  (staging is idle outside B3), piece-swizzled
    float* Gsm = st.buf0;                     // [ROWS][GSP]  (Ds is dead by then; may run into buf1), piece-swizzled
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;
    const int xc0 = rank * XC;
    const int xc1 = min(F, xc0 + XC);

    for (int idx = tid; idx < 2 * Fp * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);
        int which = col / U, u = col % U;
        int kk = which == 0 ? k : k - Fp;
        float v = 0.f;
        if (kk >= 0 && kk < F) v = p.em_w2[which][(long)kk * M + u0 + u];
        WB3[idx] = v;
    }
    const int warp = tid >> 5, lane = tid & 31;
    if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    uint32_t bar_uses = 0;
    {
        // element (gate row kphys of this CTA, column j of [W_hh | W_ih]); columns >= H + F are zero
        auto w1 = [&](int kphys, int j) -> float {
            const int uu = u0 + (kphys >> 2), s = kphys & 3;
            if (j < H) {
                const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
                return gate >= 0 ? __ldg(p.w_hh + (long)(gate * H + uu) * H + j) : 0.f;
            }
            const int c = j - H;
            const int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            return (gate >= 0 && c < F) ? __ldg(p.w_ih + (long)(gate * H + uu) * F + c) : 0.f;
        };
        auto w2 = [&](int k, int j) -> float { return __ldg(p.em_w1[k / U] + (long)(u0 + k % U) * H + j); };
        auto tmem_store16 = [&](uint32_t taddr, const uint32_t (&r)[16]) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        };
        const int lq = warp & 3;
        // hi parts into tensor memory: lane = output column / unit within a 128-wide block
        for (int b = warp >> 2; b < 3; b += 2) {
            const int j = b * 128 + lq * 32 + lane;
            for (int k0 = 0; k0 < 4 * U; k0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(j < N1 ? w1(k0 + q, j) : 0.f);
                tmem_store16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(DBT_A1_COL + b * 4 * U + k0), r);
            }
        }
        {
            const int b = warp >> 2;
            const int j = b * 128 + lq * 32 + lane;
            for (int k0 = 0; k0 < 2 * U; k0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(w2(k0 + q, j));
                tmem_store16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(DBT_A2_COL + b * 2 * U + k0), r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        // lo parts: MN-major tiles [k][column]
        for (int idx = tid; idx < 4 * U * (DBT_M1 / 4); idx += NTHREADS) {
            const int j4 = idx % (DBT_M1 / 4), k = idx / (DBT_M1 / 4);
            float4 v;
            v.x = 4 * j4 + 0 < N1 ? w1(k, 4 * j4 + 0) : 0.f; v.y = 4 * j4 + 1 < N1 ? w1(k, 4 * j4 + 1) : 0.f;
            v.z = 4 * j4 + 2 < N1 ? w1(k, 4 * j4 + 2) : 0.f; v.w = 4 * j4 + 3 < N1 ? w1(k, 4 * j4 + 3) : 0.f;
            *reinterpret_cast<float4*>(wlo1 + (size_t)(k / TS_SLAB) * (DBT_M1 / 32) * TS_CHUNK + umma::tile_off<TS_CHUNK>(k % TS_SLAB, j4)) = umma::lo_of(v);
        }
        for (int idx = tid; idx < 2 * U * (H / 4); idx += NTHREADS) {
            const int j4 = idx % (H / 4), k = idx / (H / 4);
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.em_w1[k / U] + (long)(u0 + k % U) * H + j4 * 4));
            *reinterpret_cast<float4*>(wlo2 + umma::tile_off<TS_CHUNK>(k, j4)) = umma::lo_of(v);
        }
        umma::fence_proxy_async();
    }
    for (int idx = tid; idx < ROWS * U; idx += NTHREADS) { dh_sm[idx] = 0.f; dc_sm[idx] = 0.f; }

    const int KS3 = pick_ks_fit(2 * U / 4, 2 * U, KC);
    const ThreadMap tm3 = make_map(tid, 2 * U / 4, KS3);
    constexpr int PRG = NTHREADS / U;   // pointwise mapping: 16 row groups x 16 units, rows prg + i*PRG
    constexpr int PR = ROWS / PRG;      // rows per pointwise thread
    const int pu = tid % U, prg = tid / U;
    const int unit = u0 + pu;

    const int Tt = p.lengths[tile];
    const int n_tiles = p.n_tiles, B = p.B;
    const float* dML = p.dML;
    float* Pb2 = p.Pb2 + (long)blockIdx.x * ROWS * H;
    const float* Pb2t = p.Pb2 + (long)tile * CL * ROWS * H;
    float* Pb1 = p.Pb1 + (long)blockIdx.x * ROWS * N1;
    const float* Pb1t = p.Pb1 + (long)tile * CL * ROWS * N1;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of sum(dG) over rows and steps
    float4 hsum = make_float4(0.f, 0.f, 0.f, 0.f);   // ... and of sum(dHidPre) (bias gradients of the MLP input layers)
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    const bool timing = p.timing == 3 && blockIdx.x == 0 && tid == 32;
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }
    for (int t = Tt - 1; t >= 0; --t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const long off_t = p.offsets[t];
        const long prev_base = t == 0 ? 0 : (long)B + p.offsets[t - 1];

        // saved forward activations for the gate-gradient step: fetched now, consumed after the first barrier
        float4 g4v[PR];
        float cv[PR], pv[PR], hov[PR];
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            g4v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[i] = pv[i] = hov[i] = 0.f;
            if (r >= na) continue;
            long b = tile + (long)n_tiles * r;
            long frame = off_t + b;
            long prow = prev_base + b;
            g4v[i] = __ldg(reinterpret_cast<const float4*>(p.Gs + frame * (4 * H) + unit * 4));
            hov[i] = __ldg(p.dHoff + frame * H + unit);
            if (CELL == 0) {
                cv[i] = __ldg(p.Call + ((long)B + frame) * H + unit);
                pv[i] = __ldg(p.Call + prow * H + unit);
            } else {
                pv[i] = __ldg(p.Hall + prow * H + unit);
            }
        }

        // ---- B3: gradient w.r.t. this CTA's MLP hidden pre-activations ----
        {
            float acc[RT][4];
            zero_acc(acc);
            auto rp = [=](int r) -> const float* { return dML + (off_t + tile + (long)n_tiles * r) * (2 * Fp); };
            // tanh outputs of this thread's tile: fetched before the GEMM so the latency hides behind it
            float4 hdv[RT];
            if (tm3.active && tm3.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm3.rg + i * RG;
                    hdv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm3.cg * 4;
                    hdv[i] = __ldg(reinterpret_cast<const float4*>(p.Hid + frame * (2 * M) + (col / U) * M + u0 + col % U));
                }
            }
            // the saved-activation loads above do not depend on other CTAs: they overlap the barrier that ended the
            // previous step (whose dML updates feed the contraction below)
            if (t < Tt - 1) cluster_wait_acquire();
            DEC_TICK(0)
            phase_gemm<NTHREADS, KC, 2 * U>(acc, st, 2 * Fp, rp, 0, NullRow(), na, WB3, tm3, KS3, tid);
            DEC_TICK(1)
            if (tm3.active && tm3.ks == 0) {
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    int r = tm3.rg + i * RG;
                    if (r >= na) continue;
                    long frame = off_t + tile + (long)n_tiles * r;
                    int col = tm3.cg * 4;
                    int which = col / U, uu = col % U;
                    long o = frame * (2 * M) + which * M + u0 + uu;
                    float4 hd = hdv[i];
                    float4 v = make_float4(acc[i][0] * (1.f - hd.x * hd.x), acc[i][1] * (1.f - hd.y * hd.y),
                                           acc[i][2] * (1.f - hd.z * hd.z), acc[i][3] * (1.f - hd.w * hd.w));
                    *reinterpret_cast<float4*>(p.dHidPre + o) = v;
                    *reinterpret_cast<float4*>(Ds + r * DSP + ((((col >> 2) ^ ((r >> 3) & 3))) << 2)) = v;
                    hsum.x += v.x; hsum.y += v.y; hsum.z += v.z; hsum.w += v.w;
                }
            }
            __syncthreads();
        }
        // ---- B2': partial dh over all H units from this CTA's hidden slice ----
        const int nq = (na + 3) >> 2;
        const int n_mma = (na + 15) & ~15;
        const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_ss = idesc_ts | (1u << 15);
        // re-lay `pieces` 16-byte pieces (4 k each) starting at piece `pc0` of the rows of `src` (pitch SP, piece-swizzled)
        // into the [k][row] operand tiles
        auto relayout = [&](const float* src, int SP, int pc0) {
            for (int idx = tid; idx < (ROWS / 4) * (TS_SLAB / 4); idx += NTHREADS) {
                const int q = idx % (ROWS / 4), kq = idx / (ROWS / 4);
                if (q >= nq) continue;
                const float* rb = src + (4 * q) * SP + (((pc0 + kq) ^ ((q >> 1) & 3)) << 2);
                const float4 r0 = *reinterpret_cast<const float4*>(rb);
                const float4 r1 = *reinterpret_cast<const float4*>(rb + SP);
                const float4 r2 = *reinterpret_cast<const float4*>(rb + 2 * SP);
                const float4 r3 = *reinterpret_cast<const float4*>(rb + 3 * SP);
                umma::split_store(b_hi, b_lo, umma::tile_off<TS_CHUNK>(kq * 4 + 0, q), make_float4(r0.x, r1.x, r2.x, r3.x));
                umma::split_store(b_hi, b_lo, umma::tile_off<TS_CHUNK>(kq * 4 + 1, q), make_float4(r0.y, r1.y, r2.y, r3.y));
                umma::split_store(b_hi, b_lo, umma::tile_off<TS_CHUNK>(kq * 4 + 2, q), make_float4(r0.z, r1.z, r2.z, r3.z));
                umma::split_store(b_hi, b_lo, umma::tile_off<TS_CHUNK>(kq * 4 + 3, q), make_float4(r0.w, r1.w, r2.w, r3.w));
            }
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();
        };
        // accumulator blocks D^T[column][row] -> out[column][ROWS] (global partials stay transposed: every thread writes 64
        // contiguous bytes per tcgen05.ld; the owners reduce them after the barrier)
        auto push_blocks = [&](int nblk, int ncols, float* out) {
            const int lq = warp & 3;
            for (int b = warp >> 2; b < nblk; b += 2) {
                const int j = b * 128 + lq * 32 + lane;
#pragma unroll 1
                for (int r0 = 0; r0 < na; r0 += 16) {
                    uint32_t r[16];
                    umma::tmem_ld16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * DBT_DCOLS + r0), r);
                    if (j < ncols) {
                        float4* o = reinterpret_cast<float4*>(out + (long)j * ROWS + r0);
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            o[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                    }
                }
            }
            umma::fence_before_sync();
        };
        DEC_TICK(2)
        // ---- B2' on the tensor cores: partial dh over all H units from this CTA's hidden slice (K = 2U = 32) ----
        relayout(Ds, DSP, 0);
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo), wl = umma::smem_u32(wlo2);
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                    const uint64_t dbh = umma::make_desc(bh + kg * 1024, TS_CHUNK, 512);
                    const uint64_t dbl = umma::make_desc(bl + kg * 1024, TS_CHUNK, 512);
                    const uint64_t dwl = umma::make_desc(wl + b * 4 * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
                    const uint32_t ta = tmem + (uint32_t)(DBT_A2_COL + b * 2 * U + kg * 8);
                    const uint32_t td = tmem + (uint32_t)(b * DBT_DCOLS);
                    mma_tf32_ts(td, ta, dbh, idesc_ts, kg > 0 ? 1u : 0u);
                    mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                    umma::mma_tf32(td, dwl, dbh, idesc_ss, 1u);
                }
            }
            umma::commit(&bar);
        }
        bar_uses += 1;
        umma::mbar_wait(&bar, (bar_uses - 1) & 1);
        umma::fence_after_sync();
        push_blocks(2, H, Pb2);
        DEC_TICK(3)
        cluster_sync_all();
        DEC_TICK(4)

        // ---- dh of this CTA's units = sum of the partials (+ carry), then the gate gradients ----
        for (int idx = tid; idx < ((na + 3) >> 2) * U; idx += NTHREADS) {
            const int rq = idx % ((na + 3) >> 2), j = idx / ((na + 3) >> 2);   // 4 rows of one unit: contiguous in the transposed partials
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb2t + ((long)src * H + u0 + j) * ROWS + 4 * rq);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * rq + i < na) dh_sm[(4 * rq + i) * U + j] += av[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            if (r >= na) continue;
            long frame = off_t + tile + (long)n_tiles * r;
            float dh = dh_sm[r * U + pu] + hov[i];
            float carry = dc_sm[r * U + pu];
            const float4 g4 = g4v[i];
            float4 dgv;
            if (CELL == 0) {
                float c = cv[i];
                float c_prev = pv[i];
                float tc = tanhf(c);
                float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
                float dc = carry + dh * og * (1.f - tc * tc);
                dgv.x = dc * gg * ig * (1.f - ig);
                dgv.y = dc * c_prev * fg * (1.f - fg);
                dgv.z = dc * ig * (1.f - gg * gg);
                dgv.w = dh * tc * og * (1.f - og);
                dc_sm[r * U + pu] = dc * fg;
            } else {
                dh += carry;
                float h_prev = pv[i];
                float rg_ = g4.x, zg = g4.y, ng = g4.z, hn = g4.w;
                float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
                dgv.x = dn_pre * hn * rg_ * (1.f - rg_);
                dgv.y = dh * (h_prev - ng) * zg * (1.f - zg);
                dgv.z = dn_pre;
                dgv.w = dn_pre * rg_;
                dc_sm[r * U + pu] = dh * zg;
            }
            *reinterpret_cast<float4*>(p.dG + frame * (4 * H) + unit * 4) = dgv;
            *reinterpret_cast<float4*>(Gsm + r * GSP + ((pu ^ ((r >> 3) & 3)) << 2)) = dgv;
            bsum.x += dgv.x; bsum.y += dgv.y; bsum.z += dgv.z; bsum.w += dgv.w;
        }
        __syncthreads();

        // ---- B1': partial [dh_{t-1} | dx_t] from this CTA's gate columns ----
        DEC_TICK(6)
        // ---- B1' on the tensor cores: partial [dh_{t-1} | dx_t] from this CTA's gate columns (K = 4U = 64, two halves) ----
        for (int half = 0; half < 2; ++half) {
            if (half == 1) umma::mbar_wait(&bar, (bar_uses - 1) & 1);   // the MMAs of the first half have read the tiles
            relayout(Gsm, GSP, half * (TS_SLAB / 4));
            if (tid == 0) {
                umma::fence_after_sync();
                const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo), wl = umma::smem_u32(wlo1);
#pragma unroll 1
                for (int b = 0; b < 3; ++b) {
#pragma unroll
                    for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                        const uint64_t dbh = umma::make_desc(bh + kg * 1024, TS_CHUNK, 512);
                        const uint64_t dbl = umma::make_desc(bl + kg * 1024, TS_CHUNK, 512);
                        const uint64_t dwl = umma::make_desc(wl + half * ((DBT_M1 / 32) * TS_CHUNK) + b * 4 * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
                        const uint32_t ta = tmem + (uint32_t)(DBT_A1_COL + b * 4 * U + half * TS_SLAB + kg * 8);
                        const uint32_t td = tmem + (uint32_t)(b * DBT_DCOLS);
                        mma_tf32_ts(td, ta, dbh, idesc_ts, (half > 0 || kg > 0) ? 1u : 0u);
                        mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                        umma::mma_tf32(td, dwl, dbh, idesc_ss, 1u);
                    }
                }
                umma::commit(&bar);
            }
            bar_uses += 1;
        }
        umma::mbar_wait(&bar, (bar_uses - 1) & 1);
        umma::fence_after_sync();
        push_blocks(3, N1, Pb1);
        DEC_TICK(7)
        cluster_sync_all();
        DEC_TICK(8)

        // ---- recurrent carry for this CTA's units, dx for its x columns (folded into dML of the previous frame) ----
        // x columns first (longest dependency chain: partials + saved activations + read-modify-write), one column per
        // thread; then the dh quads.  All loads of an item are issued before its first use.
        if (t > 0) {
            const long pbase = p.offsets[t - 1] + tile;
            for (int idx = tid; idx < na * XC; idx += NTHREADS) {
                const int r = idx % na, c = xc0 + idx / na;   // rows fastest: contiguous in the transposed partials
                if (c >= xc1) continue;
                const long pf = pbase + (long)n_tiles * r;   // frame that produced x_t
                float v[CL];
#pragma unroll
                for (int src = 0; src < CL; ++src) v[src] = ld_cg(Pb1t + ((long)src * N1 + H + c) * ROWS + r);
                const float lv = __ldg(p.Lv + pf * Fp + c);
                const float e = __ldg(p.eps + pf * p.ld_eps + c);
                float* dm = p.dML + pf * (2 * Fp) + c;
                const float m0 = dm[0], m1 = dm[Fp];
                float dx = v[0];
#pragma unroll
                for (int src = 1; src < CL; ++src) dx += v[src];
                dm[0] = m0 + dx;
                dm[Fp] = m1 + dx * 0.5f * expf(0.5f * lv) * e;
            }
        }
        for (int idx = tid; idx < ((na + 3) >> 2) * U; idx += NTHREADS) {
            const int rq = idx % ((na + 3) >> 2), j = idx / ((na + 3) >> 2);
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb1t + ((long)src * N1 + u0 + j) * ROWS + 4 * rq);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * rq + i < na) dh_sm[(4 * rq + i) * U + j] = av[i];
        }
        DEC_TICK(9)
        cluster_arrive_release();   // matched by the wait after the next step's prefetch (or the one after the loop)
        DEC_TICK(10)
    }
#undef DEC_TICK
    if (Tt > 0) cluster_wait_acquire();
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) g_dec_timing[i] = tacc[i]; }
    // gradient w.r.t. the initial state produced by feature2hidden (h0 = even, c0 = odd columns; model.py:100)
    for (int i = 0; i < PR; ++i) {
        int r = prg + i * PRG;
        long b = tile + (long)n_tiles * r;
        if (b >= B || r >= ROWS) continue;
        if (CELL == 0) {
            p.dHC0[b * (2 * H) + 2 * unit] = dh_sm[r * U + pu];
            p.dHC0[b * (2 * H) + 2 * unit + 1] = dc_sm[r * U + pu];
        } else {
            p.dHC0[b * H + unit] = dh_sm[r * U + pu] + dc_sm[r * U + pu];
        }
    }
    // recurrent bias gradients: fixed-order reduction over the row groups, one row of partial sums per tile
    {
        __syncthreads();
        float4* red = reinterpret_cast<float4*>(st.buf0);
        red[prg * U + pu] = bsum;
        __syncthreads();
        if (tid < U) {
            float4 a = red[tid];
            for (int g = 1; g < PRG; ++g) {
                const float4 b = red[g * U + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(p.dbias + (long)tile * (4 * H) + (u0 + tid) * 4) = a;
        }
        // MLP input-layer bias gradients: threads of the B3 epilogue hold sums for (4 columns, their row group)
        __syncthreads();
        if (tm3.active && tm3.ks == 0) red[tm3.rg * (2 * U / 4) + tm3.cg] = hsum;
        __syncthreads();
        if (tid < 2 * U / 4) {
            float4 a = red[tid];
            for (int g = 1; g < NRG; ++g) {
                const float4 b = red[g * (2 * U / 4) + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const int col = tid * 4;   // column of this CTA's slice: [which][u0 + col % U]
            *reinterpret_cast<float4*>(p.dbias1 + (long)tile * (2 * M) + (col / U) * M + u0 + col % U) = a;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// Push variant of the tensor-core backward (default for pad4(F) <= 68): the third exchange of a step -- dx_t folded into dML of the
// previous frame, which every CTA then needs in full for B3 -- no longer goes through L2 and a cluster barrier.  The owner of
// an x column updates dML in global memory (the weight-gradient GEMMs read it later) and PUSHES the updated pair
// (d mean, d log-variance) into the B3 input buffer of every CTA of the cluster with st.async (8 bytes + 8 bytes signalled on
// the consumer's mbarrier).  B3 then is a GEMM on shared-memory operands (no staging, no chunk barriers), on all 256 threads
// (8 column groups x 16 row groups x 2 K halves).  The reduction of the recurrent carry dh (not needed before the gate
// gradients) moves off the critical path, under the MMAs of B2'.  Two cluster barriers per step instead of three.
template <int CELL>
__global__ void __launch_bounds__(NTHREADS, 1) dec_bwd_push_kernel(DecK p) {
    constexpr int KC = 64;
    // (a distinct symbol: the alignment of an extern shared array is a property of the symbol, and `smem` is declared
    // 16-byte aligned by the FFMA kernels of this file)
    extern __shared__ __align__(1024) float smem_push[];
    float* const smem = smem_push;
    const int H = ABCD_H, M = ABCD_M;
    const int F = p.F, Fp = p.Fp;
    const int N1 = H + Fp;                    // columns of the B1' partial: [dh_prev (H) | dx (Fp)]
    constexpr int DSP = 2 * U + 4;            // pitch of the dHidPre slice in shared memory
    constexpr int GSP = 4 * U + 4;            // pitch of the dG slice in shared memory
    char* b_hi = reinterpret_cast<char*>(smem);                     // one 32-k slab of [k][rows] operand tiles
    char* b_lo = b_hi + TS_B_BYTES;
    char* wlo1 = b_lo + TS_B_BYTES;                                 // B1': [2 slabs][12 chunks (384 columns)][4 KB]
    char* wlo2 = wlo1 + DBT_WLO1_BYTES;                             // B2': [1 slab][8 chunks (256 units)][4 KB]
    float* WB3 = reinterpret_cast<float*>(wlo2 + DBT_WLO2_BYTES);   // [2Fp][2U]
    float* dh_sm = WB3 + 2 * Fp * (2 * U);    // [ROWS][U]
    float* dc_sm = dh_sm + ROWS * U;          // [ROWS][U]
    const Stage st = make_stage(dc_sm + ROWS * U, KC);
    const int LDA = 2 * Fp;                   // pitch of the B3 input rows
    float* As = st.buf0;                      // [ROWS][2Fp]: dML rows of the frame B3 works on, column 2c = d mean_c, 2c + 1 = d log-variance_c;
                                              // written by the owners of the x columns of the whole cluster (st.async) after the second barrier
    float* Ds = st.buf0;                      // [ROWS][DSP]  (As is dead once B3 has multiplied), piece-swizzled
    float* Gsm = st.buf0;                     // [ROWS][GSP]  (Ds is dead by then; may run into buf1), piece-swizzled
    __shared__ uint64_t bar;
    __shared__ uint64_t dbar;                 // bytes of the dML rows pushed into As (one phase per step)
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x / CL;
    const int u0 = rank * U;
    const int ng_mine = (Fp / 4 - rank + CL - 1) / CL;   // x column groups (4 columns) owned by this CTA: rank, rank + 16, ...

    for (int idx = tid; idx < 2 * Fp * 2 * U; idx += NTHREADS) {
        int col = idx % (2 * U), k = idx / (2 * U);   // row k of the interleaved input: x column k / 2, mean (0) or log-variance (1) head
        int which = col / U, u = col % U;
        float v = 0.f;
        if ((k & 1) == which && (k >> 1) < F) v = p.em_w2[which][(long)(k >> 1) * M + u0 + u];
        WB3[idx] = v;
    }
    const int warp = tid >> 5, lane = tid & 31;
    if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_init(&dbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    uint32_t bar_uses = 0;
    {
        // element (gate row kphys of this CTA, column j of [W_hh | W_ih]); columns >= H + F are zero
        auto w1 = [&](int kphys, int j) -> float {
            const int uu = u0 + (kphys >> 2), s = kphys & 3;
            if (j < H) {
                const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
                return gate >= 0 ? __ldg(p.w_hh + (long)(gate * H + uu) * H + j) : 0.f;
            }
            const int c = j - H;
            const int gate = CELL == 0 ? s : (s == 3 ? -1 : s);
            return (gate >= 0 && c < F) ? __ldg(p.w_ih + (long)(gate * H + uu) * F + c) : 0.f;
        };
        auto w2 = [&](int k, int j) -> float { return __ldg(p.em_w1[k / U] + (long)(u0 + k % U) * H + j); };
        auto tmem_store16 = [&](uint32_t taddr, const uint32_t (&r)[16]) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        };
        const int lq = warp & 3;
        // hi parts into tensor memory: lane = output column / unit within a 128-wide block
        for (int b = warp >> 2; b < 3; b += 2) {
            const int j = b * 128 + lq * 32 + lane;
            for (int k0 = 0; k0 < 4 * U; k0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(j < N1 ? w1(k0 + q, j) : 0.f);
                tmem_store16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(DBT_A1_COL + b * 4 * U + k0), r);
            }
        }
        {
            const int b = warp >> 2;
            const int j = b * 128 + lq * 32 + lane;
            for (int k0 = 0; k0 < 2 * U; k0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(w2(k0 + q, j));
                tmem_store16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(DBT_A2_COL + b * 2 * U + k0), r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        // lo parts: MN-major tiles [k][column]
        for (int idx = tid; idx < 4 * U * (DBT_M1 / 4); idx += NTHREADS) {
            const int j4 = idx % (DBT_M1 / 4), k = idx / (DBT_M1 / 4);
            float4 v;
            v.x = 4 * j4 + 0 < N1 ? w1(k, 4 * j4 + 0) : 0.f; v.y = 4 * j4 + 1 < N1 ? w1(k, 4 * j4 + 1) : 0.f;
            v.z = 4 * j4 + 2 < N1 ? w1(k, 4 * j4 + 2) : 0.f; v.w = 4 * j4 + 3 < N1 ? w1(k, 4 * j4 + 3) : 0.f;
            *reinterpret_cast<float4*>(wlo1 + (size_t)(k / TS_SLAB) * (DBT_M1 / 32) * TS_CHUNK + umma::tile_off<TS_CHUNK>(k % TS_SLAB, j4)) = umma::lo_of(v);
        }
        for (int idx = tid; idx < 2 * U * (H / 4); idx += NTHREADS) {
            const int j4 = idx % (H / 4), k = idx / (H / 4);
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.em_w1[k / U] + (long)(u0 + k % U) * H + j4 * 4));
            *reinterpret_cast<float4*>(wlo2 + umma::tile_off<TS_CHUNK>(k, j4)) = umma::lo_of(v);
        }
        umma::fence_proxy_async();
    }
    for (int idx = tid; idx < ROWS * U; idx += NTHREADS) { dh_sm[idx] = 0.f; dc_sm[idx] = 0.f; }

    // B3 thread map: 8 column groups (4 columns) x 16 row groups (rows rg + 16 j, j < 5) x 2 halves of K = 2Fp
    const int b3_cg = tid & 7, b3_rg = (tid >> 3) & 15, b3_ks = tid >> 7;
    const int b3_nq0 = (Fp / 2 + 1) / 2, b3_nq = b3_ks == 0 ? b3_nq0 : Fp / 2 - b3_nq0;   // k-quads of this thread's half
    float* b3_scratch = reinterpret_cast<float*>(b_hi);   // [128][20]: partial tiles of the second K half (the operand tiles are idle during B3)
    const uint32_t as_addr = umma::smem_u32(As), dbar_addr = umma::smem_u32(&dbar);
    constexpr int PRG = NTHREADS / U;   // pointwise mapping: 16 row groups x 16 units, rows prg + i*PRG
    constexpr int PR = ROWS / PRG;      // rows per pointwise thread
    const int pu = tid % U, prg = tid / U;
    const int unit = u0 + pu;

    const int Tt = p.lengths[tile];
    const int n_tiles = p.n_tiles, B = p.B;
    const float* dML = p.dML;
    float* Pb2 = p.Pb2 + (long)blockIdx.x * ROWS * H;
    const float* Pb2t = p.Pb2 + (long)tile * CL * ROWS * H;
    float* Pb1 = p.Pb1 + (long)blockIdx.x * ROWS * N1;
    const float* Pb1t = p.Pb1 + (long)tile * CL * ROWS * N1;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of sum(dG) over rows and steps
    float4 hsum = make_float4(0.f, 0.f, 0.f, 0.f);   // ... and of sum(dHidPre) (bias gradients of the MLP input layers)
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    // dML rows (row r of frame t - 1, this CTA's x columns) -> every CTA's As; for rows that were active at step t (r < na_t) the
    // gradient dx_t that came back through the cell is folded in first.  Pb1t: partials of step t ([16][ROWS][H + Fp]).
    auto push_dml = [&](int t_prev, int na_t) {
        const int na_p = active_rows(p.batch_sizes[t_prev], tile, n_tiles);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(dbar_addr), "r"((uint32_t)(na_p * 2 * Fp * 4)) : "memory");
        const long pbase = p.offsets[t_prev] + tile;
        // items: (row, pair of x columns): 16 bytes of the interleaved B3 input row
        for (int idx = tid; idx < na_p * ng_mine * 2; idx += NTHREADS) {
            const int hp = idx % (ng_mine * 2), r = idx / (ng_mine * 2);   // column pairs fastest: adjacent in the row-major partials
            const int c0 = 4 * (rank + CL * (hp >> 1)) + 2 * (hp & 1);   // columns c0, c0 + 1
            const long pf = pbase + (long)n_tiles * r;                   // frame that produced x_t
            float* dm = p.dML + pf * (2 * Fp) + c0;
            float2 m0 = *reinterpret_cast<const float2*>(dm), m1 = *reinterpret_cast<const float2*>(dm + Fp);
            if (r < na_t) {
                // saved activations and noise come from HBM: issued first, their latency hides behind the partial sums
                const float2 lv = __ldg(reinterpret_cast<const float2*>(p.Lv + pf * Fp + c0));
                float e[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) e[j] = c0 + j < F ? __ldg(p.eps + pf * p.ld_eps + c0 + j) : 0.f;
                float2 v[CL];
#pragma unroll
                for (int src = 0; src < CL; ++src) v[src] = ld_cg2(Pb1t + ((long)src * ROWS + r) * N1 + H + c0);
                float dx[2] = {v[0].x, v[0].y};
#pragma unroll
                for (int src = 1; src < CL; ++src) { dx[0] += v[src].x; dx[1] += v[src].y; }
                m0.x += dx[0]; m0.y += dx[1];
                m1.x += dx[0] * 0.5f * expf(0.5f * lv.x) * e[0]; m1.y += dx[1] * 0.5f * expf(0.5f * lv.y) * e[1];
            }
            // padding columns carry zeros (their rows of the output weights are zero, but stale bits could be NaN)
            if (c0 + 0 >= F) { m0.x = 0.f; m1.x = 0.f; }
            if (c0 + 1 >= F) { m0.y = 0.f; m1.y = 0.f; }
            if (r < na_t) {
                *reinterpret_cast<float2*>(dm) = m0;
                *reinterpret_cast<float2*>(dm + Fp) = m1;
            }
            const uint32_t a = as_addr + (uint32_t)(r * LDA + 2 * c0) * 4u;
#pragma unroll
            for (unsigned q = 0; q < (unsigned)CL; ++q) st_async_cluster_v4(a, dbar_addr, q, m0.x, m1.x, m0.y, m1.y);
        }
    };
    // recurrent carry dh of this CTA's units = sum of the B1' partials of the step that has just finished (rows < na_t)
    auto reduce_dh = [&](int na_t) {
        for (int idx = tid; idx < na_t * (U / 4); idx += NTHREADS) {
            const int jq = idx % (U / 4), r = idx / (U / 4);
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb1t + ((long)src * ROWS + r) * N1 + u0 + 4 * jq);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            *reinterpret_cast<float4*>(dh_sm + r * U + 4 * jq) = a;
        }
    };
    cluster_sync_all();   // every CTA's mbarriers are initialised before the first remote store
    if (Tt > 0) push_dml(Tt - 1, 0);
    int na_last = 0;      // active rows of the step whose B1' partials have not been reduced into dh_sm yet

    const bool timing = p.timing == 3 && blockIdx.x == p.timing_cta && tid == 32;
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define DEC_TICK(slot) if (timing) { long long now_ = clock64(); tacc[slot] += now_ - tick_; tick_ = now_; }
    for (int t = Tt - 1; t >= 0; --t) {
        long long tick_ = timing ? clock64() : 0;
        const int na = active_rows(p.batch_sizes[t], tile, n_tiles);
        const long off_t = p.offsets[t];
        const long prev_base = t == 0 ? 0 : (long)B + p.offsets[t - 1];

        // saved forward activations for the gate-gradient step: fetched now, consumed after the first barrier
        float4 g4v[PR];
        float cv[PR], pv[PR], hov[PR];
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            g4v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[i] = pv[i] = hov[i] = 0.f;
            if (r >= na) continue;
            long b = tile + (long)n_tiles * r;
            long frame = off_t + b;
            long prow = prev_base + b;
            g4v[i] = __ldg(reinterpret_cast<const float4*>(p.Gs + frame * (4 * H) + unit * 4));
            hov[i] = __ldg(p.dHoff + frame * H + unit);
            if (CELL == 0) {
                cv[i] = __ldg(p.Call + ((long)B + frame) * H + unit);
                pv[i] = __ldg(p.Call + prow * H + unit);
            } else {
                pv[i] = __ldg(p.Hall + prow * H + unit);
            }
        }

        // ---- B3: gradient w.r.t. this CTA's MLP hidden pre-activations (operands in shared memory) ----
        {
            float acc[DFA_RT][4];
#pragma unroll
            for (int j = 0; j < DFA_RT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
            // tanh outputs of this thread's tile: fetched before the GEMM so the latency hides behind it
            float4 hdv[DFA_RT];
            if (b3_ks == 0) {
#pragma unroll
                for (int j = 0; j < DFA_RT; ++j) {
                    const int r = b3_rg + 16 * j;
                    hdv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r >= na) continue;
                    const long frame = off_t + tile + (long)n_tiles * r;
                    const int col = b3_cg * 4;
                    hdv[j] = __ldg(reinterpret_cast<const float4*>(p.Hid + frame * (2 * M) + (col / U) * M + u0 + col % U));
                }
            }
            umma::mbar_wait(&dbar, (uint32_t)(Tt - 1 - t) & 1u);   // the dML rows of frame t have landed in As
            DEC_TICK(0)
            {
                const int njw = min(DFA_RT, (na - 4 * (warp & 3) + 15) >> 4);   // warp-uniform trip count (its 4 row groups)
                const float* ar = As + b3_rg * LDA + b3_ks * b3_nq0 * 4;
                const float* wq = WB3 + (b3_ks * b3_nq0 * 4) * (2 * U) + b3_cg * 4;
                switch (njw) {
                    case 1: rows_gemm<1, 2 * U>(acc, ar, LDA, wq, b3_nq); break;
                    case 2: rows_gemm<2, 2 * U>(acc, ar, LDA, wq, b3_nq); break;
                    case 3: rows_gemm<3, 2 * U>(acc, ar, LDA, wq, b3_nq); break;
                    case 4: rows_gemm<4, 2 * U>(acc, ar, LDA, wq, b3_nq); break;
                    case 5: rows_gemm<5, 2 * U>(acc, ar, LDA, wq, b3_nq); break;
                    default: break;
                }
            }
            if (b3_ks == 1) {
#pragma unroll
                for (int j = 0; j < DFA_RT; ++j)
                    *reinterpret_cast<float4*>(b3_scratch + ((tid - 128) * DFA_RT + j) * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            }
            __syncthreads();   // partial tiles visible; every thread is done reading As (Ds aliases it)
            DEC_TICK(1)
            if (b3_ks == 0) {
#pragma unroll
                for (int j = 0; j < DFA_RT; ++j) {
                    const int r = b3_rg + 16 * j;
                    if (r >= na) continue;
                    const float4 o2 = *reinterpret_cast<const float4*>(b3_scratch + (tid * DFA_RT + j) * 4);
                    const long frame = off_t + tile + (long)n_tiles * r;
                    const int col = b3_cg * 4;
                    const int which = col / U, uu = col % U;
                    const long o = frame * (2 * M) + which * M + u0 + uu;
                    const float4 hd = hdv[j];
                    const float4 v = make_float4((acc[j][0] + o2.x) * (1.f - hd.x * hd.x), (acc[j][1] + o2.y) * (1.f - hd.y * hd.y),
                                                 (acc[j][2] + o2.z) * (1.f - hd.z * hd.z), (acc[j][3] + o2.w) * (1.f - hd.w * hd.w));
                    *reinterpret_cast<float4*>(p.dHidPre + o) = v;
                    *reinterpret_cast<float4*>(Ds + r * DSP + ((((col >> 2) ^ ((r >> 3) & 3))) << 2)) = v;
                    hsum.x += v.x; hsum.y += v.y; hsum.z += v.z; hsum.w += v.w;
                }
            }
            __syncthreads();
        }
        // ---- B2': partial dh over all H units from this CTA's hidden slice ----
        const int nq = (na + 3) >> 2;
        const int n_mma = (na + 15) & ~15;
        const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_ss = idesc_ts | (1u << 15);
        // re-lay `pieces` 16-byte pieces (4 k each) starting at piece `pc0` of the rows of `src` (pitch SP, piece-swizzled)
        // into the [k][row] operand tiles
        // `toff`: byte offset of the destination tile inside b_hi / b_lo (a second 32-row tile when at most 32 rows are active)
        auto relayout_nosync = [&](const float* src, int SP, int pc0, uint32_t toff) {
            for (int idx = tid; idx < (ROWS / 4) * (TS_SLAB / 4); idx += NTHREADS) {
                const int q = idx % (ROWS / 4), kq = idx / (ROWS / 4);
                if (q >= nq) continue;
                const float* rb = src + (4 * q) * SP + (((pc0 + kq) ^ ((q >> 1) & 3)) << 2);
                const float4 r0 = *reinterpret_cast<const float4*>(rb);
                const float4 r1 = *reinterpret_cast<const float4*>(rb + SP);
                const float4 r2 = *reinterpret_cast<const float4*>(rb + 2 * SP);
                const float4 r3 = *reinterpret_cast<const float4*>(rb + 3 * SP);
                umma::split_store(b_hi, b_lo, toff + umma::tile_off<TS_CHUNK>(kq * 4 + 0, q), make_float4(r0.x, r1.x, r2.x, r3.x));
                umma::split_store(b_hi, b_lo, toff + umma::tile_off<TS_CHUNK>(kq * 4 + 1, q), make_float4(r0.y, r1.y, r2.y, r3.y));
                umma::split_store(b_hi, b_lo, toff + umma::tile_off<TS_CHUNK>(kq * 4 + 2, q), make_float4(r0.z, r1.z, r2.z, r3.z));
                umma::split_store(b_hi, b_lo, toff + umma::tile_off<TS_CHUNK>(kq * 4 + 3, q), make_float4(r0.w, r1.w, r2.w, r3.w));
            }
        };
        auto relayout = [&](const float* src, int SP, int pc0) {
            relayout_nosync(src, SP, pc0, 0u);
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();
        };
        // accumulator blocks D^T[column][row] -> out[row][ncols] (global partials, row-major: the 32 lanes of a warp hold 32
        // consecutive columns, so every store instruction writes 128 contiguous bytes; the owners reduce them after the barrier)
        auto push_blocks = [&](int nblk, int ncols, float* out) {
            const int lq = warp & 3;
            for (int b = warp >> 2; b < nblk; b += 2) {
                const int j = b * 128 + lq * 32 + lane;
#pragma unroll 1
                for (int r0 = 0; r0 < na; r0 += 32) {
                    const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * DBT_DCOLS + r0);
                    float* o = out + (long)r0 * ncols + j;
                    if (r0 + 16 < na) {   // 32 rows with one load (r0 + 32 <= 80 accumulator columns for r0 = 0, 32)
                        uint32_t r[32];
                        umma::tmem_ld32(taddr, r);
                        if (j < ncols) {
#pragma unroll
                            for (int v = 0; v < 32; ++v)
                                if (r0 + v < na) o[(long)v * ncols] = __uint_as_float(r[v]);
                        }
                    } else {
                        uint32_t r[16];
                        umma::tmem_ld16(taddr, r);
                        if (j < ncols) {
#pragma unroll
                            for (int v = 0; v < 16; ++v)
                                if (r0 + v < na) o[(long)v * ncols] = __uint_as_float(r[v]);
                        }
                    }
                }
            }
            umma::fence_before_sync();
        };
        DEC_TICK(2)
        // ---- B2' on the tensor cores: partial dh over all H units from this CTA's hidden slice (K = 2U = 32) ----
        relayout(Ds, DSP, 0);
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo), wl = umma::smem_u32(wlo2);
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                    const uint64_t dbh = umma::make_desc(bh + kg * 1024, TS_CHUNK, 512);
                    const uint64_t dbl = umma::make_desc(bl + kg * 1024, TS_CHUNK, 512);
                    const uint64_t dwl = umma::make_desc(wl + b * 4 * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
                    const uint32_t ta = tmem + (uint32_t)(DBT_A2_COL + b * 2 * U + kg * 8);
                    const uint32_t td = tmem + (uint32_t)(b * DBT_DCOLS);
                    mma_tf32_ts(td, ta, dbh, idesc_ts, kg > 0 ? 1u : 0u);
                    mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                    umma::mma_tf32(td, dwl, dbh, idesc_ss, 1u);
                }
            }
            umma::commit(&bar);
        }
        bar_uses += 1;
        // under the MMAs: the recurrent carry left by the previous iteration's B1' (its partials stay valid until this step's B1')
        if (na_last > 0) reduce_dh(na_last);
        umma::mbar_wait(&bar, (bar_uses - 1) & 1);
        umma::fence_after_sync();
        push_blocks(2, H, Pb2);
        DEC_TICK(3)
        cluster_sync_all();
        DEC_TICK(4)

        // ---- dh of this CTA's units = sum of the partials (+ carry), then the gate gradients ----
        for (int idx = tid; idx < na * (U / 4); idx += NTHREADS) {
            const int jq = idx % (U / 4), r = idx / (U / 4);   // 4 units of one row: contiguous in the row-major partials
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pb2t + ((long)src * ROWS + r) * H + u0 + 4 * jq);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            float4* d = reinterpret_cast<float4*>(dh_sm + r * U + 4 * jq);
            const float4 o = *d;
            *d = make_float4(o.x + a.x, o.y + a.y, o.z + a.z, o.w + a.w);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            int r = prg + i * PRG;
            if (r >= na) continue;
            long frame = off_t + tile + (long)n_tiles * r;
            float dh = dh_sm[r * U + pu] + hov[i];
            float carry = dc_sm[r * U + pu];
            const float4 g4 = g4v[i];
            float4 dgv;
            if (CELL == 0) {
                float c = cv[i];
                float c_prev = pv[i];
                float tc = tanhf(c);
                float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
                float dc = carry + dh * og * (1.f - tc * tc);
                dgv.x = dc * gg * ig * (1.f - ig);
                dgv.y = dc * c_prev * fg * (1.f - fg);
                dgv.z = dc * ig * (1.f - gg * gg);
                dgv.w = dh * tc * og * (1.f - og);
                dc_sm[r * U + pu] = dc * fg;
            } else {
                dh += carry;
                float h_prev = pv[i];
                float rg_ = g4.x, zg = g4.y, ng = g4.z, hn = g4.w;
                float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
                dgv.x = dn_pre * hn * rg_ * (1.f - rg_);
                dgv.y = dh * (h_prev - ng) * zg * (1.f - zg);
                dgv.z = dn_pre;
                dgv.w = dn_pre * rg_;
                dc_sm[r * U + pu] = dh * zg;
            }
            *reinterpret_cast<float4*>(p.dG + frame * (4 * H) + unit * 4) = dgv;
            *reinterpret_cast<float4*>(Gsm + r * GSP + ((pu ^ ((r >> 3) & 3)) << 2)) = dgv;
            bsum.x += dgv.x; bsum.y += dgv.y; bsum.z += dgv.z; bsum.w += dgv.w;
        }
        __syncthreads();

        // ---- B1': partial [dh_{t-1} | dx_t] from this CTA's gate columns ----
        DEC_TICK(6)
        // ---- B1' on the tensor cores: partial [dh_{t-1} | dx_t] from this CTA's gate columns (K = 4U = 64, two halves) ----
        if (n_mma <= 32) {
            // at most 32 active rows: a K half needs one 4 KB chunk of the operand tiles, so both halves are laid out side by side
            // and all 72 MMAs are issued in one go (one relayout round, one completion wait)
            relayout_nosync(Gsm, GSP, 0, 0u);
            relayout_nosync(Gsm, GSP, TS_SLAB / 4, (uint32_t)TS_CHUNK);
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();
            DEC_TICK(10)
            if (tid == 0) {
                umma::fence_after_sync();
                const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo), wl = umma::smem_u32(wlo1);
                // consecutive instructions go to different accumulator blocks (an instruction that accumulates into the block
                // of its predecessor waits for it)
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                        const uint64_t dbh = umma::make_desc(bh + half * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
                        const uint64_t dbl = umma::make_desc(bl + half * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                            for (int b = 0; b < 3; ++b) {
                                const uint32_t ta = tmem + (uint32_t)(DBT_A1_COL + b * 4 * U + half * TS_SLAB + kg * 8);
                                const uint32_t td = tmem + (uint32_t)(b * DBT_DCOLS);
                                if (prod == 0) mma_tf32_ts(td, ta, dbh, idesc_ts, (half > 0 || kg > 0) ? 1u : 0u);
                                else if (prod == 1) mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                                else umma::mma_tf32(td, umma::make_desc(wl + half * ((DBT_M1 / 32) * TS_CHUNK) + b * 4 * TS_CHUNK + kg * 1024, TS_CHUNK, 512), dbh, idesc_ss, 1u);
                            }
                        }
                    }
                }
                umma::commit(&bar);
            }
            bar_uses += 1;
        } else
        for (int half = 0; half < 2; ++half) {
            if (half == 1) umma::mbar_wait(&bar, (bar_uses - 1) & 1);   // the MMAs of the first half have read the tiles
            relayout(Gsm, GSP, half * (TS_SLAB / 4));
            if (tid == 0) {
                umma::fence_after_sync();
                const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo), wl = umma::smem_u32(wlo1);
#pragma unroll 1
                for (int b = 0; b < 3; ++b) {
#pragma unroll
                    for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                        const uint64_t dbh = umma::make_desc(bh + kg * 1024, TS_CHUNK, 512);
                        const uint64_t dbl = umma::make_desc(bl + kg * 1024, TS_CHUNK, 512);
                        const uint64_t dwl = umma::make_desc(wl + half * ((DBT_M1 / 32) * TS_CHUNK) + b * 4 * TS_CHUNK + kg * 1024, TS_CHUNK, 512);
                        const uint32_t ta = tmem + (uint32_t)(DBT_A1_COL + b * 4 * U + half * TS_SLAB + kg * 8);
                        const uint32_t td = tmem + (uint32_t)(b * DBT_DCOLS);
                        mma_tf32_ts(td, ta, dbh, idesc_ts, (half > 0 || kg > 0) ? 1u : 0u);
                        mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                        umma::mma_tf32(td, dwl, dbh, idesc_ss, 1u);
                    }
                }
                umma::commit(&bar);
            }
            bar_uses += 1;
        }
        umma::mbar_wait(&bar, (bar_uses - 1) & 1);
        umma::fence_after_sync();
        DEC_TICK(11)
        push_blocks(3, N1, Pb1);
        DEC_TICK(7)
        cluster_sync_all();
        DEC_TICK(8)

        // ---- dx for this CTA's x columns, folded into dML of the previous frame and pushed to every CTA's B3 input ----
        if (t > 0) push_dml(t - 1, na);
        na_last = na;
        DEC_TICK(9)
    }
#undef DEC_TICK
    if (na_last > 0) reduce_dh(na_last);   // carry out of step 0: gradient of the initial state
    __syncthreads();
    if (timing) { tacc[15] = Tt; for (int i = 0; i < 16; ++i) g_dec_timing[i] = tacc[i]; }
    // gradient w.r.t. the initial state produced by feature2hidden (h0 = even, c0 = odd columns; model.py:100)
    for (int i = 0; i < PR; ++i) {
        int r = prg + i * PRG;
        long b = tile + (long)n_tiles * r;
        if (b >= B || r >= ROWS) continue;
        if (CELL == 0) {
            p.dHC0[b * (2 * H) + 2 * unit] = dh_sm[r * U + pu];
            p.dHC0[b * (2 * H) + 2 * unit + 1] = dc_sm[r * U + pu];
        } else {
            p.dHC0[b * H + unit] = dh_sm[r * U + pu] + dc_sm[r * U + pu];
        }
    }
    // recurrent bias gradients: fixed-order reduction over the row groups, one row of partial sums per tile
    {
        __syncthreads();
        float4* red = reinterpret_cast<float4*>(st.buf0);
        red[prg * U + pu] = bsum;
        __syncthreads();
        if (tid < U) {
            float4 a = red[tid];
            for (int g = 1; g < PRG; ++g) {
                const float4 b = red[g * U + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(p.dbias + (long)tile * (4 * H) + (u0 + tid) * 4) = a;
        }
        // MLP input-layer bias gradients: threads of the B3 epilogue hold sums for (4 columns, their row group)
        __syncthreads();
        if (b3_ks == 0) red[b3_rg * (2 * U / 4) + b3_cg] = hsum;
        __syncthreads();
        if (tid < 2 * U / 4) {
            float4 a = red[tid];
            for (int g = 1; g < 16; ++g) {
                const float4 b = red[g * (2 * U / 4) + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const int col = tid * 4;   // column of this CTA's slice: [which][u0 + col % U]
            *reinterpret_cast<float4*>(p.dbias1 + (long)tile * (2 * M) + (col / U) * M + u0 + col % U) = a;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

// -------------------------------------------------------------------------------------------------
// helper kernels outside the time loop
// -------------------------------------------------------------------------------------------------
// hc0 [B][2H] interleaved (h even, c odd) -> Hall[0..B), Call[0..B)   (GRU: hc0 [B][H] -> Hall)
__global__ void split_hc0_kernel(int cell, int B, const float* __restrict__ hc0, float* __restrict__ Hall, float* __restrict__ Call) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * ABCD_H) return;
    if (cell == 0) {
        Hall[idx] = hc0[2 * (long)idx];
        Call[idx] = hc0[2 * (long)idx + 1];
    } else {
        Hall[idx] = hc0[idx];
    }
}

__global__ void final_sum_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    __shared__ double red[40];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) *out = (float)s;
}

// offset logit = hid . w2 + b2 ; BCE-with-logits (sum) against is_offset  (model.py:122,191-195)
__global__ void offset_head_kernel(int sumT, const float* __restrict__ hid, const float* __restrict__ w2,
                                   const float* __restrict__ b2, const float* __restrict__ y,
                                   float* __restrict__ logit_out, float* __restrict__ partial) {
    __shared__ float red[40];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float loss = 0.f;
    for (long n = (long)blockIdx.x * nw + w; n < sumT; n += (long)gridDim.x * nw) {
        float s = 0.f;
        for (int j = lane; j < ABCD_M; j += 32) s += hid[n * ABCD_M + j] * w2[j];
        s = warp_sum(s) + b2[0];
        if (lane == 0) {
            logit_out[n] = s;
            if (y) loss += fmaxf(s, 0.f) - s * y[n] + log1pf(expf(-fabsf(s)));
        }
    }
    float tot = block_sum(loss, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// d logit = (sigmoid(logit) - y) * g_off ;  dHidOffPre[n][j] = dlogit * w2[j] * (1 - hid^2)
__global__ void offset_head_bwd_kernel(int sumT, const float* __restrict__ hid, const float* __restrict__ w2,
                                       const float* __restrict__ logit, const float* __restrict__ y,
                                       const float* __restrict__ g_losses, float* __restrict__ dlogit,
                                       float* __restrict__ dpre) {
    const float g = g_losses[1];
    long total = (long)sumT * ABCD_M;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long n = idx / ABCD_M; int j = (int)(idx % ABCD_M);
        float dl = (sigmoid_acc(logit[n]) - y[n]) * g;
        float hd = hid[idx];
        dpre[idx] = dl * w2[j] * (1.f - hd * hd);
        if (j == 0) dlogit[n] = dl;
    }
}

// dML[n] = [ d nll/d mean | d nll/d logvar ] * g_em   (padding columns zero)   (model.py:30-37)
__global__ void nll_grad_kernel(int sumT, int F, int Fp, const float* __restrict__ mean, const float* __restrict__ lv,
                                const float* __restrict__ gt, int ld_gt, const float* __restrict__ g_losses,
                                float* __restrict__ dML) {
    const float g = g_losses[0];
    long total = (long)sumT * Fp;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long n = idx / Fp; int c = (int)(idx % Fp);
        float dm = 0.f, dl = 0.f;
        if (c < F) {
            float d = gt[n * ld_gt + c] - mean[idx];
            float e = expf(-lv[idx]);
            dm = -d * e * g;
            dl = 0.5f * (1.f - d * d * e) * g;
        }
        dML[n * (2 * Fp) + c] = dm;
        dML[n * (2 * Fp) + Fp + c] = dl;
    }
}

__global__ void dec_prev_index_kernel(int T_max, int B, const int* __restrict__ batch_sizes, const int* __restrict__ offsets,
                                      int* __restrict__ prev) {
    for (int t = blockIdx.x; t < T_max; t += gridDim.x) {
        int bs = batch_sizes[t], off = offsets[t];
        for (int b = threadIdx.x; b < bs; b += blockDim.x) prev[off + b] = t > 0 ? B + offsets[t - 1] + b : b;
    }
}

__global__ void unpad_copy_kernel(long rows, int F, int ld_src, const float* __restrict__ src, float* __restrict__ dst) {
    long total = rows * F;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long n = idx / F; int c = (int)(idx % F);
        dst[idx] = src[n * ld_src + c];
    }
}

size_t dec_fwd_fixed_floats(int F) {   // resident weights + reduction scratch
    int Fp = pad4(F);
    return (size_t)((Fp + ABCD_H) * U * 4 + ABCD_H * 2 * U + U * 2 * Fp + 64);
}
size_t dec_bwd_fixed_floats(int F) {   // resident weights + dh/dc carries
    int Fp = pad4(F);
    return (size_t)(2 * Fp * 2 * U + 2 * U * ABCD_H + 4 * U * DEC_N1P + 2 * ROWS * U);
}
int pick_kc(size_t fixed_floats, size_t limit = SMEM_MAX) {
    for (int kc : {128, 64, 32})
        if ((fixed_floats + stage_floats(kc)) * sizeof(float) <= limit) return kc;
    return 0;
}
int dec_fwd_kc(int F) { return pick_kc(dec_fwd_fixed_floats(F), SMEM_MAX - 1024); }   // the forward kernel also has static shared memory
int dec_bwd_kc(int F) { return pick_kc(dec_bwd_fixed_floats(F)); }
size_t dec_fwd_smem(int F) { return (dec_fwd_fixed_floats(F) + stage_floats(dec_fwd_kc(F))) * sizeof(float); }
size_t dec_bwd_smem(int F) { return (dec_bwd_fixed_floats(F) + stage_floats(dec_bwd_kc(F))) * sizeof(float); }
// tensor-core forward: [tiles + raw ring | W_lo | W1 | W2s | scratch]
size_t dec_fwd_ts_smem(int F) {
    const int Fp = pad4(F);
    return (size_t)TSI_STAGES * TS_STAGE_BYTES + ts_wlo_bytes_m<U * 4>(ABCD_H + Fp) + (size_t)(ABCD_H * 2 * U + U * 2 * Fp + 64) * sizeof(float);
}
// pair variant (default): [activation ring | W1 | W2s | scratch]
size_t dec_fwd_pair_smem(int F) {
    const int Fp = pad4(F);
    return (size_t)TSP_STAGES * TS_STAGE_BYTES + (size_t)(ABCD_H * 2 * U + U * 2 * Fp + 64) * sizeof(float);
}
// one-step-ahead pair variant (default): [h ring | W1 | W2s | Wx | gpre | staging | scratch]
size_t dec_fwd_ahead_smem(int F) {
    const int Fp = pad4(F);
    return (size_t)DFA_STAGES * TS_STAGE_BYTES + ((size_t)(ABCD_H * 2 * U + U * 2 * Fp + Fp * 4 * U + ROWS * TSP_GP + 64) + dfa_stage_floats()) * sizeof(float);
}
bool dec_fwd_use_ahead(int F) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("ABCD_DEC_FWD"); off = (e && strcmp(e, "ahead")) ? 1 : 0; }   // any other explicit choice disables it
    return !off && rnn_use_tc() && pad4(F) <= DFA_MAX_FP && dec_fwd_ahead_smem(F) <= SMEM_MAX - 1024;
}
// ABCD_DEC_FWD=ahead (default) | pair | ffma | tc64.  The pair variant pays one more exchange per step than the FFMA kernel and wins once a
// tile has enough rows for the tensor-core contraction to matter (measured: 1.41 vs 1.52 ms at 73 rows per tile, 1.20 vs 1.18 at 37, 1.03 vs 0.94 ms at
// 10); below DEC_PAIR_MIN_ROWS rows per tile (small shards under strong scaling) the FFMA kernel runs.
constexpr int DEC_PAIR_MIN_ROWS = 48;
int dec_fwd_mode(int F, int B = 1 << 30) {
    static int opt = -1;
    if (opt < 0) {
        const char* e = getenv("ABCD_DEC_FWD");
        opt = (e && !strcmp(e, "ffma")) ? 0 : (e && !strcmp(e, "tc64")) ? 1 : 2;
        const char* e2 = getenv("ABCD_DEC_FWD_TC");
        if (e2 && e2[0] == '1') opt = 1;
    }
    const int nsl = ABCD_H / TS_SLAB + (pad4(F) + TS_SLAB - 1) / TS_SLAB;
    static int force = -1;
    if (force < 0) { const char* e = getenv("ABCD_DEC_FWD"); force = (e && !strcmp(e, "pair")) ? 1 : 0; }
    const bool rows_ok = force || (B + num_tiles(B) - 1) / num_tiles(B) >= DEC_PAIR_MIN_ROWS;
    if (opt == 2 && rows_ok && rnn_use_tc() && (nsl + 1) / 2 <= TSP_MAX_SLABS && dec_fwd_pair_smem(F) <= SMEM_MAX - 1024) return 2;
    return opt == 2 ? 0 : opt;
}
// The M = 64 tensor-core gate contraction of the forward kernel is parity-green but not yet faster than the FFMA phase
// (1.68 vs 1.52 ms: eleven serial split -> barrier -> issue rounds per step): opt-in with ABCD_DEC_FWD_TC=1.
bool dec_fwd_use_ts(int F) {
    static int opt = -1;
    if (opt < 0) { const char* e = getenv("ABCD_DEC_FWD_TC"); opt = (e && e[0] == '1') ? 1 : 0; }
    const int nsl = (ABCD_H + pad4(F) + TS_SLAB - 1) / TS_SLAB;
    (void)opt;
    return dec_fwd_mode(F) == 1 && rnn_use_tc() && TS_D_COLS_M + nsl * TS_SLAB <= 512 && dec_fwd_ts_smem(F) <= SMEM_MAX - 1024;
}
// tensor-core backward: [operand tiles | W_lo tiles | WB3 | dh/dc carries | B3 staging (also the dHidPre / dG slices)]
size_t dec_bwd_ts_fixed_bytes(int F) {
    const int Fp = pad4(F);
    return (size_t)2 * TS_B_BYTES + DBT_WLO1_BYTES + DBT_WLO2_BYTES + (size_t)(2 * Fp * 2 * U + 2 * ROWS * U) * sizeof(float);
}
int dec_bwd_ts_kc(int F) {
    for (int kc : {64, 32})
        if (dec_bwd_ts_fixed_bytes(F) + stage_floats(kc) * sizeof(float) <= SMEM_MAX - 1024) return kc;
    return 0;
}
bool dec_bwd_use_ts(int F) { return rnn_use_tc() && ABCD_H + pad4(F) <= DBT_M1 && dec_bwd_ts_kc(F) != 0; }
size_t dec_bwd_ts_smem(int F) { return dec_bwd_ts_fixed_bytes(F) + stage_floats(dec_bwd_ts_kc(F)) * sizeof(float); }
// push variant (default): same shared-memory layout with the 64-wide staging area, which must hold the B3 input rows [ROWS][2Fp];
// ABCD_DEC_BWD=ts selects the barrier-exchange kernel
bool dec_bwd_use_push(int F) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("ABCD_DEC_BWD"); off = (e && strcmp(e, "push")) ? 1 : 0; }
    return !off && dec_bwd_use_ts(F) && dec_bwd_ts_kc(F) == 64 && (size_t)ROWS * 2 * pad4(F) <= stage_floats(64);
}

template <class K>
int launch_cluster_dec(K kernel, int n_clusters, size_t smem, cudaStream_t stream, const DecK& params) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(n_clusters * CL);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    abcd_note_launch(1);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
    return ABCD_OK;
}

struct DecWs {
    float *hc0, *Hall, *Call, *Xall, *Gs, *Hid, *Mean, *Lv, *HidOff, *OffLogit, *nll_partial, *bce_partial, *Pbuf;
    char *himg, *ximg;
    float *dML, *dHidPre, *dG, *dHoff, *dHidOffPre, *dLogit, *dHC0, *dbias, *dbias1, *Pb2, *Pb1, *partials;
    int* prev;
    size_t partial_floats;
    size_t bytes;
};
constexpr int BCE_BLOCKS = 296;

DecWs dec_carve(void* base, size_t cap, int cell, int F, int fin, int B, int sumT) {
    Arena a(base, cap);
    DecWs w;
    memset(&w, 0, sizeof(w));
    const int H = ABCD_H, M = ABCD_M, Fp = pad4(F);
    const int hc = cell == 0 ? 2 * H : H;
    const int n_tiles = num_tiles(B);
    w.hc0 = a.take<float>((size_t)B * hc);
    w.Hall = a.take<float>((size_t)(B + sumT) * H);
    w.Call = a.take<float>((size_t)(B + sumT) * H);
    w.Xall = a.take<float>((size_t)(B + sumT) * Fp);
    w.Gs = a.take<float>((size_t)sumT * 4 * H);
    w.Hid = a.take<float>((size_t)sumT * 2 * M);
    w.Mean = a.take<float>((size_t)sumT * Fp);
    w.Lv = a.take<float>((size_t)sumT * Fp);
    w.HidOff = a.take<float>((size_t)sumT * M);
    w.OffLogit = a.take<float>((size_t)sumT);
    w.nll_partial = a.take<float>((size_t)n_tiles * CL);
    w.bce_partial = a.take<float>(BCE_BLOCKS);
    w.Pbuf = a.take<float>((size_t)n_tiles * CL * ROWS * 2 * Fp);
    w.himg = reinterpret_cast<char*>(a.take<float>((size_t)n_tiles * 2 * ts_image_bytes(H) / sizeof(float)));
    w.ximg = reinterpret_cast<char*>(a.take<float>((size_t)n_tiles * 2 * ((Fp + TS_SLAB - 1) / TS_SLAB) * TS_STAGE_BYTES / sizeof(float)));
    w.dML = a.take<float>((size_t)sumT * 2 * Fp);
    w.dHidPre = a.take<float>((size_t)sumT * 2 * M);
    w.dG = a.take<float>((size_t)sumT * 4 * H);
    w.dHoff = a.take<float>((size_t)sumT * H);
    w.dHidOffPre = a.take<float>((size_t)sumT * M);
    w.dLogit = a.take<float>((size_t)sumT);
    w.dHC0 = a.take<float>((size_t)B * hc);
    w.dbias = a.take<float>((size_t)n_tiles * 4 * H);
    w.dbias1 = a.take<float>((size_t)n_tiles * 2 * M);
    w.Pb2 = a.take<float>((size_t)n_tiles * CL * ROWS * H);
    w.Pb1 = a.take<float>((size_t)n_tiles * CL * ROWS * (H + Fp));
    w.prev = a.take<int>((size_t)sumT);
    // split-K partials of ALL weight-gradient GEMMs (their epilogues are deferred and batched, see EpilogueBatch) plus the
    // scratch of the batched column sums
    size_t mx = 0;
    {
        int dims[][3] = {{4 * H, H, sumT}, {4 * H, F, sumT}, {M, H, sumT}, {M, H, sumT}, {M, H, sumT}, {F, M, sumT}, {F, M, sumT}, {1, M, sumT}, {hc, fin, B}};
        for (auto& d : dims) mx += (((size_t)sgemm_pick_splits(d[0], d[1], d[2]) * d[0] * d[1]) + 3) & ~(size_t)3;
        mx += (size_t)12 * 600 * 64;   // column sums: <= ~600 row splits x 32-column strips per job
    }
    w.partial_floats = mx;
    w.partials = a.take<float>(mx);
    w.bytes = a.off;
    return w;
}

DecK make_k(const DecParams& P, const PackedMeta& pk, const DecWs& w) {
    DecK k;
    memset(&k, 0, sizeof(k));
    k.cell = P.cell; k.F = P.F; k.Fp = pad4(P.F);
    k.n_tiles = num_tiles(pk.B); k.B = pk.B; k.T_max = pk.T_max;
    k.batch_sizes = pk.batch_sizes; k.offsets = pk.offsets; k.lengths = pk.lengths;
    k.w_ih = P.w_ih; k.w_hh = P.w_hh; k.b_ih = P.b_ih; k.b_hh = P.b_hh;
    for (int i = 0; i < 2; ++i) { k.em_w1[i] = P.em_w1[i]; k.em_b1[i] = P.em_b1[i]; k.em_w2[i] = P.em_w2[i]; k.em_b2[i] = P.em_b2[i]; }
    k.Hall = w.Hall; k.Call = w.Call; k.Xall = w.Xall; k.Gs = w.Gs; k.Hid = w.Hid; k.Mean = w.Mean; k.Lv = w.Lv;
    k.nll_partial = w.nll_partial; k.Pbuf = w.Pbuf; k.himg = w.himg; k.ximg = w.ximg;
    k.timing = getenv("ABCD_TIMING") ? atoi(getenv("ABCD_TIMING")) : 0;   // 1: forward kernel, 3: backward kernel
    k.timing_cta = getenv("ABCD_TIMING_CTA") ? atoi(getenv("ABCD_TIMING_CTA")) : 0;
    k.dML = w.dML; k.dHidPre = w.dHidPre; k.dG = w.dG; k.dHoff = w.dHoff; k.dHC0 = w.dHC0; k.dbias = w.dbias; k.dbias1 = w.dbias1; k.Pb2 = w.Pb2; k.Pb1 = w.Pb1;
    return k;
}

int check_dec(const DecParams& P) {
    ARG_CHECK(P.H == ABCD_H, "decoder: rnn_hidden_size %d unsupported (kernels are specialised for %d)", P.H, ABCD_H);
    ARG_CHECK(P.M == ABCD_M, "decoder: mlp_hidden_size %d unsupported (kernels are specialised for %d)", P.M, ABCD_M);
    ARG_CHECK(P.F >= 1 && P.F <= DEC_MAX_F, "decoder: output_size %d unsupported (1..%d)", P.F, DEC_MAX_F);
    ARG_CHECK(P.cell == 0 || P.cell == 1, "decoder: rnn_type must be LSTM or GRU");
    return ABCD_OK;
}

}  // namespace

int rnn_max_active_clusters(const void* fn, size_t smem, int cluster, int threads, int* out);
int dec_read_timing(unsigned long long* out16) {
    CUDA_TRY(cudaMemcpyFromSymbol(out16, g_dec_timing, 16 * sizeof(unsigned long long)));
    return ABCD_OK;
}
int dec_max_clusters(int which, int F, int* out) {
    if (which == 1 && dec_bwd_use_push(F))
        return rnn_max_active_clusters((const void*)dec_bwd_push_kernel<0>, dec_bwd_ts_smem(F), CL, NTHREADS, out);
    if (which == 1 && dec_bwd_use_ts(F))   // the tensor-memory backward
        return rnn_max_active_clusters(dec_bwd_ts_kc(F) == 64 ? (const void*)dec_bwd_ts_kernel<0, 64> : (const void*)dec_bwd_ts_kernel<0, 32>,
                                       dec_bwd_ts_smem(F), CL, NTHREADS, out);
    if (which == 0 && dec_fwd_use_ahead(F))
        return rnn_max_active_clusters((const void*)dec_fwd_ahead_kernel<0>, dec_fwd_ahead_smem(F), CL, NTHREADS, out);
    if (which == 0 && dec_fwd_mode(F) == 2)
        return rnn_max_active_clusters((const void*)dec_fwd_pair_kernel<0>, dec_fwd_pair_smem(F), CL, NTHREADS, out);
    const int kc = which == 0 ? dec_fwd_kc(F) : dec_bwd_kc(F);
    const void* fn = which == 0 ? (kc == 128 ? (const void*)dec_fwd_kernel<0, 128> : kc == 64 ? (const void*)dec_fwd_kernel<0, 64> : (const void*)dec_fwd_kernel<0, 32>)
                                : (kc == 128 ? (const void*)dec_bwd_kernel<0, 128> : kc == 64 ? (const void*)dec_bwd_kernel<0, 64> : (const void*)dec_bwd_kernel<0, 32>);
    return rnn_max_active_clusters(fn, which == 0 ? dec_fwd_smem(F) : dec_bwd_smem(F), CL, NTHREADS, out);
}

int dec_init() {
    TRY(rnn_prepare_kernel((const void*)dec_bwd_push_kernel<0>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_push_kernel<1>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_ts_kernel<0, 64>, SMEM_MAX - 1024));   // the kernel also has static shared memory
    TRY(rnn_prepare_kernel((const void*)dec_bwd_ts_kernel<1, 64>, SMEM_MAX - 1024));   // the kernel also has static shared memory
    TRY(rnn_prepare_kernel((const void*)dec_bwd_ts_kernel<0, 32>, SMEM_MAX - 1024));   // the kernel also has static shared memory
    TRY(rnn_prepare_kernel((const void*)dec_bwd_ts_kernel<1, 32>, SMEM_MAX - 1024));   // the kernel also has static shared memory
    TRY(rnn_prepare_kernel((const void*)dec_fwd_ahead_kernel<0>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_ahead_kernel<1>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_pair_kernel<0>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_pair_kernel<1>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<0, 64, true>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<1, 64, true>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<0, 128>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<1, 128>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<0, 128>, SMEM_MAX));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<1, 128>, SMEM_MAX));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<0, 64>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<1, 64>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<0, 32>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_fwd_kernel<1, 32>, SMEM_MAX - 1024));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<0, 64>, SMEM_MAX));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<1, 64>, SMEM_MAX));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<0, 32>, SMEM_MAX));
    TRY(rnn_prepare_kernel((const void*)dec_bwd_kernel<1, 32>, SMEM_MAX));
    return ABCD_OK;
}

size_t dec_workspace_bytes(int cell, int F, int fin, int B, int sumT) {
    return dec_carve(nullptr, 0, cell, F, fin, B, sumT).bytes;
}

int dec_forward(const DecParams& P, const PackedMeta& pk, const float* feat_in, const float* eps, int ld_eps,
                const float* gt, int ld_gt, const float* is_offset, float* losses, float* out_x, float* out_mean,
                float* out_logvar, float* out_offset_logit, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TRY(check_dec(P));
    ARG_CHECK(pk.B > 0 && pk.T_max > 0 && pk.sumT > 0, "decoder: empty packed batch");
    ARG_CHECK(eps != nullptr, "decoder: emission noise buffer is required");
    const int H = ABCD_H, M = ABCD_M, Fp = pad4(P.F);
    const int hc = P.cell == 0 ? 2 * H : H;
    DecWs w = dec_carve(ws, ws_bytes, P.cell, P.F, P.fin, pk.B, pk.sumT);
    if (w.bytes > ws_bytes) { abcd_set_error("decoder: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return ABCD_ERR_WORKSPACE; }
    const int n_tiles = num_tiles(pk.B);

    // feature2hidden (model.py:169-170): hc0 = feat_in W^T + b
    GemmArgs g = gemm_args_zero();
    g.M = pk.B; g.N = hc; g.K = P.fin;
    g.A = feat_in; g.lda = P.fin; g.a_kmajor = 1;
    g.B = P.w_f2h; g.ldb = P.fin; g.b_kmajor = 1;
    g.C = w.hc0; g.ldc = hc; g.bias = P.b_f2h;
    TRY(sgemm_launch(g, stream));
    split_hc0_kernel<<<(pk.B * H + 255) / 256, 256, 0, stream>>>(P.cell, pk.B, w.hc0, w.Hall, w.Call);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemsetAsync(w.Xall, 0, (size_t)(pk.B + pk.sumT) * Fp * sizeof(float), stream));

    DecK k = make_k(P, pk, w);
    k.eps = eps; k.ld_eps = ld_eps; k.gt = gt; k.ld_gt = ld_gt;
    k.kc = dec_fwd_kc(P.F);
    ARG_CHECK(k.kc > 0, "decoder: weights for output_size %d do not fit in shared memory", P.F);
    {
        ProfScope ps(PROF_DEC_FWD, stream);
        const size_t sm = dec_fwd_smem(P.F);
        if (dec_fwd_use_ahead(P.F)) {
            k.kc = DFA_KC;
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_ahead_kernel<0>, n_tiles, dec_fwd_ahead_smem(P.F), stream, k));
            else TRY(launch_cluster_dec(dec_fwd_ahead_kernel<1>, n_tiles, dec_fwd_ahead_smem(P.F), stream, k));
        } else if (dec_fwd_mode(P.F, pk.B) == 2) {
            k.kc = DFP_KC;
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_pair_kernel<0>, n_tiles, dec_fwd_pair_smem(P.F), stream, k));
            else TRY(launch_cluster_dec(dec_fwd_pair_kernel<1>, n_tiles, dec_fwd_pair_smem(P.F), stream, k));
        } else if (dec_fwd_use_ts(P.F)) {
            k.kc = 64;
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_kernel<0, 64, true>, n_tiles, dec_fwd_ts_smem(P.F), stream, k));
            else TRY(launch_cluster_dec(dec_fwd_kernel<1, 64, true>, n_tiles, dec_fwd_ts_smem(P.F), stream, k));
        } else if (k.kc == 128) {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_kernel<0, 128>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_fwd_kernel<1, 128>, n_tiles, sm, stream, k));
        } else if (k.kc == 64) {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_kernel<0, 64>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_fwd_kernel<1, 64>, n_tiles, sm, stream, k));
        } else {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_fwd_kernel<0, 32>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_fwd_kernel<1, 32>, n_tiles, sm, stream, k));
        }
    }
    if (gt && losses) {
        final_sum_kernel<<<1, 256, 0, stream>>>(w.nll_partial, n_tiles * CL, losses + 0);
        KERNEL_CHECK();
    }
    // offset predictor over all frames (model.py:191)
    g = gemm_args_zero();
    g.M = pk.sumT; g.N = M; g.K = H;
    g.A = w.Hall + (size_t)pk.B * H; g.lda = H; g.a_kmajor = 1;
    g.B = P.off_w1; g.ldb = H; g.b_kmajor = 1;
    g.C = w.HidOff; g.ldc = M; g.bias = P.off_b1; g.act = 1;
    TRY(sgemm_launch(g, stream));
    offset_head_kernel<<<BCE_BLOCKS, 256, 0, stream>>>(pk.sumT, w.HidOff, P.off_w2, P.off_b2, is_offset, w.OffLogit, w.bce_partial);
    KERNEL_CHECK();
    if (is_offset && losses) {
        final_sum_kernel<<<1, 256, 0, stream>>>(w.bce_partial, BCE_BLOCKS, losses + 1);
        KERNEL_CHECK();
    }
    const int cp_blocks = 592;
    if (out_x) unpad_copy_kernel<<<cp_blocks, 256, 0, stream>>>(pk.sumT, P.F, Fp, w.Xall + (size_t)pk.B * Fp, out_x);
    if (out_mean) unpad_copy_kernel<<<cp_blocks, 256, 0, stream>>>(pk.sumT, P.F, Fp, w.Mean, out_mean);
    if (out_logvar) unpad_copy_kernel<<<cp_blocks, 256, 0, stream>>>(pk.sumT, P.F, Fp, w.Lv, out_logvar);
    if (out_offset_logit) CUDA_TRY(cudaMemcpyAsync(out_offset_logit, w.OffLogit, (size_t)pk.sumT * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    KERNEL_CHECK();
    return ABCD_OK;
}

// parts: bit 0 = data path (loss-head gradients, BPTT, gradient of the features), bit 1 = weight gradients (reductions over all
// packed frames of the buffers the data path left in the workspace; may run on another stream once the data path is done)
int dec_backward(const DecParams& P, const DecGrads& G, const PackedMeta& pk, const float* feat_in, const float* eps,
                 int ld_eps, const float* gt, int ld_gt, const float* is_offset, const float* g_losses,
                 float* d_feat_in, void* ws, size_t ws_bytes, cudaStream_t stream, int parts) {
    TRY(check_dec(P));
    const bool do_data = parts & 1, do_wgrad = parts & 2;
    ARG_CHECK(!do_data || (gt && is_offset && g_losses && eps), "decoder bwd: ground truth, offsets, noise and loss gradients are required");
    const int H = ABCD_H, M = ABCD_M, F = P.F, Fp = pad4(P.F);
    const int hc = P.cell == 0 ? 2 * H : H;
    DecWs w = dec_carve(ws, ws_bytes, P.cell, P.F, P.fin, pk.B, pk.sumT);
    if (w.bytes > ws_bytes) { abcd_set_error("decoder bwd: workspace too small"); return ABCD_ERR_WORKSPACE; }
    ColsumScratch colsum_scratch(w.partials, w.partial_floats);
    const int n_tiles = num_tiles(pk.B);
    const int sumT = pk.sumT;
    const float* Hfr = w.Hall + (size_t)pk.B * H;   // h of frame n

    GemmArgs g = gemm_args_zero();
    if (do_data) {
    // ---- loss-head gradients over all frames ----
    nll_grad_kernel<<<592, 256, 0, stream>>>(sumT, F, Fp, w.Mean, w.Lv, gt, ld_gt, g_losses, w.dML);
    KERNEL_CHECK();
    offset_head_bwd_kernel<<<592, 256, 0, stream>>>(sumT, w.HidOff, P.off_w2, w.OffLogit, is_offset, g_losses, w.dLogit, w.dHidOffPre);
    KERNEL_CHECK();
    // dHoff = dHidOffPre * off_w1
    g.M = sumT; g.N = H; g.K = M;
    g.A = w.dHidOffPre; g.lda = M; g.a_kmajor = 1;
    g.B = P.off_w1; g.ldb = H; g.b_kmajor = 0;
    g.C = w.dHoff; g.ldc = H;
    TRY(sgemm_launch(g, stream));

    // ---- BPTT ----
    DecK k = make_k(P, pk, w);
    k.eps = eps; k.ld_eps = ld_eps; k.gt = gt; k.ld_gt = ld_gt;
    k.kc = dec_bwd_kc(P.F);
    ARG_CHECK(k.kc > 0, "decoder bwd: weights for output_size %d do not fit in shared memory", P.F);
    {
        ProfScope ps(PROF_DEC_BWD, stream);
        const size_t sm = dec_bwd_smem(P.F);
        if (dec_bwd_use_push(P.F)) {
            k.kc = 64;
            if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_push_kernel<0>, n_tiles, dec_bwd_ts_smem(P.F), stream, k));
            else TRY(launch_cluster_dec(dec_bwd_push_kernel<1>, n_tiles, dec_bwd_ts_smem(P.F), stream, k));
        } else if (dec_bwd_use_ts(P.F)) {
            k.kc = dec_bwd_ts_kc(P.F);
            const size_t smt = dec_bwd_ts_smem(P.F);
            if (k.kc == 64) {
                if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_ts_kernel<0, 64>, n_tiles, smt, stream, k));
                else TRY(launch_cluster_dec(dec_bwd_ts_kernel<1, 64>, n_tiles, smt, stream, k));
            } else {
                if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_ts_kernel<0, 32>, n_tiles, smt, stream, k));
                else TRY(launch_cluster_dec(dec_bwd_ts_kernel<1, 32>, n_tiles, smt, stream, k));
            }
        } else if (k.kc == 128) {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_kernel<0, 128>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_bwd_kernel<1, 128>, n_tiles, sm, stream, k));
        } else if (k.kc == 64) {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_kernel<0, 64>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_bwd_kernel<1, 64>, n_tiles, sm, stream, k));
        } else {
            if (P.cell == 0) TRY(launch_cluster_dec(dec_bwd_kernel<0, 32>, n_tiles, sm, stream, k));
            else TRY(launch_cluster_dec(dec_bwd_kernel<1, 32>, n_tiles, sm, stream, k));
        }
    }

    if (d_feat_in) {   // gradient of the decoder input features (feature2hidden data gradient): on the critical path
        g = gemm_args_zero();
        g.M = pk.B; g.N = P.fin; g.K = hc;
        g.A = w.dHC0; g.lda = hc; g.a_kmajor = 1;
        g.B = P.w_f2h; g.ldb = P.fin; g.b_kmajor = 0;
        g.C = d_feat_in; g.ldc = P.fin;
        TRY(sgemm_launch(g, stream));
    }
    }   // do_data
    if (!do_wgrad) return ABCD_OK;

    // ---- weight gradients: reductions over all packed frames (epilogues and column sums deferred into three batched launches) ----
    EpilogueBatch epilogues(w.partials, w.partial_floats, stream);
    dec_prev_index_kernel<<<min(pk.T_max, 592), 128, 0, stream>>>(pk.T_max, pk.B, pk.batch_sizes, pk.offsets, w.prev);
    KERNEL_CHECK();
    auto tn = [&](int Mo, int No, const float* A, long lda, const float* Bm, long ldb, const int* gather, GateMap md,
                  float* C, long ldc, int Kred) -> int {
        GemmArgs q = gemm_args_zero();
        q.M = Mo; q.N = No; q.K = Kred;
        q.A = A; q.lda = lda; q.a_kmajor = 0;
        q.B = Bm; q.ldb = ldb; q.b_kmajor = 0; q.b_kgather = gather;
        q.m_dst = md; q.C = C; q.ldc = ldc; q.accumulate = 1;
        q.splits = sgemm_pick_splits(Mo, No, Kred); q.partials = w.partials;
        if ((size_t)q.splits * Mo * No > w.partial_floats) q.splits = 1;
        return sgemm_launch(q, stream);
    };
    const GateMap map_ih = P.cell == 0 ? gate_map(H, 0, 1, 2, 3) : gate_map(H, 0, 1, 2, -1);
    const GateMap map_hh = P.cell == 0 ? gate_map(H, 0, 1, 2, 3) : gate_map(H, 0, 1, -1, 2);
    TRY(tn(4 * H, F, w.dG, 4 * H, w.Xall, Fp, w.prev, map_ih, G.w_ih, F, sumT));
    TRY(tn(4 * H, H, w.dG, 4 * H, w.Hall, H, w.prev, map_hh, G.w_hh, H, sumT));
    // the BPTT kernel already reduced dG over the frames of every tile
    TRY(colsum_launch(w.dbias, 4 * H, num_tiles(pk.B), 4 * H, 1.f, G.b_ih, 1, map_ih, stream));
    TRY(colsum_launch(w.dbias, 4 * H, num_tiles(pk.B), 4 * H, 1.f, G.b_hh, 1, map_hh, stream));
    for (int i = 0; i < 2; ++i) {
        TRY(tn(M, H, w.dHidPre + i * M, 2 * M, Hfr, H, nullptr, gate_map_none(), G.em_w1[i], H, sumT));
        // the BPTT kernel already reduced dHidPre over the frames of every tile
        TRY(colsum_launch(w.dbias1 + i * M, 2 * M, num_tiles(pk.B), M, 1.f, G.em_b1[i], 1, gate_map_none(), stream));
        TRY(tn(F, M, w.dML + i * Fp, 2 * Fp, w.Hid + i * M, 2 * M, nullptr, gate_map_none(), G.em_w2[i], M, sumT));
        TRY(colsum_launch(w.dML + i * Fp, 2 * Fp, sumT, F, 1.f, G.em_b2[i], 1, gate_map_none(), stream));
    }
    TRY(tn(M, H, w.dHidOffPre, M, Hfr, H, nullptr, gate_map_none(), G.off_w1, H, sumT));
    TRY(colsum_launch(w.dHidOffPre, M, sumT, M, 1.f, G.off_b1, 1, gate_map_none(), stream));
    TRY(tn(1, M, w.dLogit, 1, w.HidOff, M, nullptr, gate_map_none(), G.off_w2, M, sumT));
    TRY(colsum_launch(w.dLogit, 1, sumT, 1, 1.f, G.off_b2, 1, gate_map_none(), stream));
    // feature2hidden
    TRY(tn(hc, P.fin, w.dHC0, hc, feat_in, P.fin, nullptr, gate_map_none(), G.w_f2h, P.fin, pk.B));
    TRY(colsum_launch(w.dHC0, hc, pk.B, hc, 1.f, G.b_f2h, 1, gate_map_none(), stream));
    return epilogues.flush();
}
