// Encoder recurrence: bidirectional single-layer LSTM/GRU over packed variable-length segments, returning
// only the final states, i.e. RNN_Variational_Encoder.forward (ABCD-VAE/modules/model.py:60-66) and its
// backward.  Gate order follows torch.nn.LSTM (i,f,g,o) / torch.nn.GRU (r,z,n) (model.py:53).
//
// Forward = hoisted input projection (tensor-core GEMM over all packed frames, gates stored unit-major interleaved)
//         + one persistent cluster kernel per launch that runs both directions concurrently.
// Backward = persistent BPTT kernel (K-split local GEMM + in-cluster reduction) producing pre-activation gate
//         gradients for every frame and the bias-gradient sums + weight gradients as tensor-core TN GEMMs over all frames.
#include <stdlib.h>
#include "rnn_common.cuh"
#include "rnn_ts.cuh"
#include "sgemm.cuh"
#include "enc.cuh"

using namespace rnn;

namespace {

// The encoder runs 8-CTA clusters (32 hidden units and 512 threads per CTA): 2 directions x 7 row tiles = 14
// clusters are resident together, where 16-CTA clusters would need two waves.
constexpr int ENC_CL = 8;
constexpr int ENC_U = ABCD_H / ENC_CL;
constexpr int ENC_NT = NRG * ENC_U;
constexpr int ENC_KC_FWD = 128;

// -------------------------------------------------------------------------------------------------
// small helper kernels
// -------------------------------------------------------------------------------------------------
// Input-projection operands for the tensor-core GEMM: weight_ih re-homed as [4H physical columns (u*4+slot)][Fpad]
// (16-byte aligned rows, zero padded) and the combined bias in the same interleaved order.
__global__ void prep_ih_kernel(int cell, int F, int Fpad, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                               const float* __restrict__ b_hh, float* __restrict__ w_out, float* __restrict__ b_out) {
    const int H = ABCD_H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * H * Fpad; idx += gridDim.x * blockDim.x) {
        const int phys = idx / Fpad, k = idx % Fpad;
        const int u = phys >> 2, s = phys & 3;
        const int gate = cell == 0 ? s : (s == 3 ? -1 : s);
        w_out[idx] = (gate >= 0 && k < F) ? w_ih[(long)(gate * H + u) * F + k] : 0.f;
        if (k == 0) {
            float b = 0.f;
            if (gate >= 0) {
                b = b_ih[gate * H + u];
                if (cell == 0 || gate < 2) b += b_hh[gate * H + u];
            }
            b_out[phys] = b;
        }
    }
}

__global__ void pad_rows_kernel(long rows, int F, int Fpad, const float* __restrict__ src, int ld_src, float* __restrict__ dst) {
    long total = rows * Fpad;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long n = idx / Fpad; int c = (int)(idx % Fpad);
        dst[idx] = c < F ? src[n * ld_src + c] : 0.f;
    }
}

// prev_fwd[n]: packed index of the same sequence at t-1 (or -1); prev_bwd[n]: at t+1 (or -1)
__global__ void prev_index_kernel(int T_max, const int* __restrict__ batch_sizes, const int* __restrict__ offsets,
                                  int* __restrict__ prev_fwd, int* __restrict__ prev_bwd) {
    for (int t = blockIdx.x; t < T_max; t += gridDim.x) {
        int bs = batch_sizes[t], off = offsets[t];
        int bs_next = (t + 1 < T_max) ? batch_sizes[t + 1] : 0;
        for (int b = threadIdx.x; b < bs; b += blockDim.x) {
            prev_fwd[off + b] = t > 0 ? offsets[t - 1] + b : -1;
            prev_bwd[off + b] = b < bs_next ? offsets[t + 1] + b : -1;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
__device__ unsigned long long g_enc_timing[10];

struct EncFwdK {
    int n_tiles, B, ndirs, save, out_ld;
    int timing;   // debug: accumulate per-phase clock cycles of CTA 0 into g_enc_timing
    const int* batch_sizes; const int* offsets; const int* lengths;
    const float* w_hh[2];
    const float* b_hh[2];
    const float* xproj[2];   // [sumT][4H] interleaved
    float* hs[2];            // save: [sumT][H]; else ping-pong [2][B][H]
    float* gs[2];            // save: [sumT][4H] interleaved activations
    float* cs[2];            // save (LSTM): [sumT][H]
    float* out;              // [B][out_ld]
    char* himg[2];           // tensor-core path: [clusters of the direction][2 parities] operand-tile images of h (rnn_ts.cuh)
};

// TC = true: the per-step gate contraction runs on the tensor cores with the recurrent weights resident in tensor
// memory (rnn_ts.cuh); shared memory is then [activation tiles][W_lo tiles][raw staging ring] instead of
// [W_hh slice][staging buffers].
template <int CELL, bool TC>
__global__ void __launch_bounds__(ENC_NT, 1) enc_fwd_kernel(EncFwdK p) {
    constexpr int CL = ENC_CL, U = ENC_U, NTHREADS = ENC_NT;   // 8-CTA clusters, 32 units per CTA, 256 threads
    extern __shared__ __align__(1024) float smem[];
    float* Wsm = smem;                                         // FFMA path: [H][U*4]
    const Stage st = make_stage(Wsm + ABCD_H * (U * 4), ENC_KC_FWD);
    __shared__ uint64_t ts_bars[TS_STAGES + 1];
    __shared__ uint64_t tsi_bars[2 * TSI_STAGES + 1];
    __shared__ uint32_t ts_tmem_slot;
    TsCtx ts;
    TsImgCtx tsi;
    char* img[2] = {nullptr, nullptr};       // this cluster's two h images (written in step s, read in step s+1)

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL;
    const int tile = cid % p.n_tiles, dir = cid / p.n_tiles;
    const int u0 = rank * U;
    const int H = ABCD_H;

    if (TC) {
        static_assert(ENC_U * 4 == TS_M, "one CTA = 128 gate columns = the tensor-memory lanes");
        if ((umma::smem_u32(smem) & 1023u) != 0) __trap();      // operand tiles need 1024-byte alignment
        char* base = reinterpret_cast<char*>(smem);
        const float* W = p.w_hh[dir];
        auto wrow = [=](int col) -> const float* {
            const int u = col >> 2, s = col & 3;
            const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            return gate >= 0 ? W + (long)(gate * ABCD_H + u0 + u) * ABCD_H : nullptr;
        };
        char* wlo = base + (size_t)TSI_STAGES * TS_STAGE_BYTES;
        ts_init<ABCD_H>(ts, base, wlo, nullptr, ts_bars, &ts_tmem_slot, wrow, tid);
        ts_img_init_bars(tsi, base, wlo, tsi_bars, ts.tmem, tid);
        // rows that join later (reverse direction) must read h_prev = 0: clear both images once
        img[0] = p.himg[dir] + (size_t)(tile * 2) * ts_image_bytes(ABCD_H);
        img[1] = img[0] + ts_image_bytes(ABCD_H);
        {
            float4* z = reinterpret_cast<float4*>(img[0]);
            const size_t n16 = 2 * ts_image_bytes(ABCD_H) / 16;
            for (size_t i = (size_t)rank * NTHREADS + tid; i < n16; i += (size_t)CL * NTHREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ts_image_publish();
        umma::fence_before_sync();
        cluster_sync_all();
        umma::fence_after_sync();
    } else {
        // resident recurrent weights: Wsm[k][u*4+slot] = W_hh[gate(slot)*H + u0+u][k]
        const float* W = p.w_hh[dir];
        for (int idx = tid; idx < U * 4 * (H / 4); idx += NTHREADS) {
            int kq = idx % (H / 4), col = idx / (H / 4);
            int u = col >> 2, s = col & 3;
            int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gate >= 0) v = *reinterpret_cast<const float4*>(W + (long)(gate * H + u0 + u) * H + kq * 4);
            Wsm[(kq * 4 + 0) * (U * 4) + col] = v.x;
            Wsm[(kq * 4 + 1) * (U * 4) + col] = v.y;
            Wsm[(kq * 4 + 2) * (U * 4) + col] = v.z;
            Wsm[(kq * 4 + 3) * (U * 4) + col] = v.w;
        }
    }
    const ThreadMap tm = make_map(tid, U, 1);   // cg == unit, ks == 0 for all 256 threads
    const int unit = u0 + tm.cg;
    const float bhn = CELL == 1 ? p.b_hh[dir][2 * H + unit] : 0.f;

    int len[RT];
    float creg[RT], hreg[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        int b = tile + p.n_tiles * (tm.rg + i * RG);
        len[i] = b < p.B ? p.lengths[b] : 0;
        creg[i] = 0.f; hreg[i] = 0.f;
    }
    const int Tt = p.lengths[tile];
    const float* xp_base = p.xproj[dir];
    float* hs = p.hs[dir];
    __syncthreads();

    unsigned long long tacc[5] = {0, 0, 0, 0, 0};
    unsigned long long pg[5] = {0, 0, 0, 0, 0};
    const bool timing = p.timing && blockIdx.x == 0 && tid == ((TC && p.timing == 1) ? 32 : 0);
    for (int s = 0; s < Tt; ++s) {
        long long c0 = timing ? clock64() : 0;
        const int t = dir == 0 ? s : Tt - 1 - s;
        const int na = active_rows(p.batch_sizes[t], tile, p.n_tiles);
        const int off_t = p.offsets[t];
        // rows that already carry state from the previous step of this direction
        int na_prev, off_prev;
        if (dir == 0) { na_prev = s > 0 ? na : 0; off_prev = s > 0 ? p.offsets[t - 1] : 0; }
        else { na_prev = s > 0 ? active_rows(p.batch_sizes[t + 1], tile, p.n_tiles) : 0; off_prev = s > 0 ? p.offsets[t + 1] : 0; }

        // projected inputs for this thread's rows (issued early; consumed after the GEMM)
        float4 xp[RT];
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = tm.rg + i * RG;
            xp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < na) {
                long frame = off_t + tile + p.n_tiles * r;
                xp[i] = __ldg(reinterpret_cast<const float4*>(xp_base + frame * (4 * H) + unit * 4));
            }
        }
        // the loads above are independent of other CTAs: they overlap the barrier that ended the previous step
        if (s > 0) cluster_wait_acquire();
        float acc[RT][4];
        zero_acc(acc);
        long long c1 = timing ? clock64() : 0;
        if (na_prev > 0) {
            const int n_tiles = p.n_tiles, B = p.B, save = p.save;
            auto rp = [=](int r) -> const float* {
                if (r >= na_prev) return nullptr;
                long b = tile + n_tiles * r;
                long row = save ? (long)off_prev + b : (long)((s - 1) & 1) * B + b;
                return hs + row * ABCD_H;
            };
            if (TC) phase_gemm_ts_img<ABCD_H>(acc, tsi, img[(s - 1) & 1], na, tm, tid);
            else phase_gemm<ENC_NT, ENC_KC_FWD, ENC_U * 4>(acc, st, H, rp, 0, NullRow(), na, Wsm, tm, 1, tid, timing ? pg : nullptr);
        }
        long long c2 = timing ? clock64() : 0;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = tm.rg + i * RG;
            if (r >= na) continue;
            long b = tile + p.n_tiles * r;
            long frame = off_t + b;
            float h;
            if (CELL == 0) {
                float ig = sigmoid_acc(acc[i][0] + xp[i].x);
                float fg = sigmoid_acc(acc[i][1] + xp[i].y);
                float gg = tanhf(acc[i][2] + xp[i].z);
                float og = sigmoid_acc(acc[i][3] + xp[i].w);
                float c = fg * creg[i] + ig * gg;
                h = og * tanhf(c);
                creg[i] = c;
                if (p.save) {
                    *reinterpret_cast<float4*>(p.gs[dir] + frame * (4 * H) + unit * 4) = make_float4(ig, fg, gg, og);
                    p.cs[dir][frame * H + unit] = c;
                }
            } else {
                float rg_ = sigmoid_acc(acc[i][0] + xp[i].x);
                float zg = sigmoid_acc(acc[i][1] + xp[i].y);
                float hn = acc[i][3] + bhn;
                float ng = tanhf(xp[i].z + rg_ * hn);
                h = (1.f - zg) * ng + zg * hreg[i];
                if (p.save)
                    *reinterpret_cast<float4*>(p.gs[dir] + frame * (4 * H) + unit * 4) = make_float4(rg_, zg, ng, hn);
            }
            hreg[i] = h;
            long row = p.save ? frame : (long)(s & 1) * p.B + b;
            hs[row * H + unit] = h;
            if (TC) ts_image_store(img[s & 1], unit, r, h);
            const bool last = dir == 0 ? (t == len[i] - 1) : (t == 0);
            if (last) {
                float* o = p.out + b * p.out_ld;
                if (CELL == 0) { o[dir * 2 * H + unit] = h; o[dir * 2 * H + H + unit] = creg[i]; }
                else o[dir * H + unit] = h;
            }
        }
        long long c3 = timing ? clock64() : 0;
        if (TC) ts_image_publish();
        cluster_arrive_release();   // matched by the wait after the next step's input loads (or the one after the loop)
        if (timing) {
            long long c4 = clock64();
            tacc[0] += c1 - c0; tacc[1] += c2 - c1; tacc[2] += c3 - c2; tacc[3] += c4 - c3; tacc[4] += 1;
        }
    }
    if (Tt > 0) cluster_wait_acquire();
    if (timing) for (int i = 0; i < 5; ++i) { g_enc_timing[i] = tacc[i]; g_enc_timing[5 + i] = pg[i]; }
    if (TC) ts_fini(ts, tid);
}

// -------------------------------------------------------------------------------------------------
// backward (BPTT): produces dG[dir][frame][u*4+slot] (pre-activation gate gradients)
// -------------------------------------------------------------------------------------------------
struct EncBwdK {
    int n_tiles, B, ndirs, out_ld;
    const int* batch_sizes; const int* offsets; const int* lengths;
    const float* w_hh[2];
    const float* hs[2]; const float* gs[2]; const float* cs[2];
    const float* d_out;   // [B][out_ld]
    float* dg[2];         // [sumT][4H]
    float* dbias[2];      // [n_tiles][4H]: per-tile column sums of dg (bias gradients), interleaved layout
    float* pbuf;          // [2 parities][gridDim.x][ROWS][H]: partial dh exchanged inside a cluster
};

// The gate gradients dG_t of a step are produced unit by unit (thread-local cell backward), so the contraction
// dh_{t-1} = dG_t W_hh runs over a dimension that is partitioned over the cluster: every CTA multiplies ITS dG slice
// (still in shared memory) by the matching rows of W_hh into a partial dh over all H units, publishes it through L2,
// and after the cluster barrier the owner of each unit adds the CL partials in a fixed order.  The partial buffers
// alternate between steps, so one cluster barrier per step is enough.
template <int CELL>
__global__ void __launch_bounds__(ENC_NT, 1) enc_bwd_kernel(EncBwdK p) {
    constexpr int CL = ENC_CL, U = ENC_U, NTHREADS = ENC_NT;
    constexpr int GSP = 4 * U + 4;           // pitch of the dG slice in shared memory
    extern __shared__ __align__(16) float smem[];
    const int H = ABCD_H;
    float* Wr = smem;                        // [4U][H]: Wr[u*4+slot][j] = W_hh[gate(slot)*H + u0+u][j]
    float* dh_sm = Wr + 4 * U * H;           // [ROWS][U] recurrent dh carried to the next (earlier) step
    float* dc_sm = dh_sm + ROWS * U;         // [ROWS][U] LSTM: dc carry; GRU: direct dh*z carry
    float* Gsm = dc_sm + ROWS * U;           // [ROWS][GSP] this step's dG slice

    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL;
    const int tile = cid % p.n_tiles, dir = cid / p.n_tiles;
    const int u0 = rank * U;
    {
        const float* W = p.w_hh[dir];
        for (int idx = tid; idx < 4 * U * (H / 4); idx += NTHREADS) {
            int jq = idx % (H / 4), kphys = idx / (H / 4);
            int uu = u0 + (kphys >> 2), s = kphys & 3;
            int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gate >= 0) v = *reinterpret_cast<const float4*>(W + (long)(gate * H + uu) * H + jq * 4);
            *reinterpret_cast<float4*>(Wr + kphys * H + jq * 4) = v;
        }
        for (int idx = tid; idx < ROWS * U; idx += NTHREADS) { dh_sm[idx] = 0.f; dc_sm[idx] = 0.f; }
    }
    // pointwise mapping: thread -> (unit pu, rows prg + i*RG)
    const int pu = tid % U, prg = tid / U;
    const int unit = u0 + pu;
    int len[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        int b = tile + p.n_tiles * (prg + i * RG);
        len[i] = b < p.B ? p.lengths[b] : 0;
    }
    const int Tt = p.lengths[tile];
    const float* gs = p.gs[dir];
    const float* cs = p.cs[dir];
    const float* hs = p.hs[dir];
    float* dg = p.dg[dir];
    const long pb_stride = (long)gridDim.x * ROWS * H;             // one parity buffer
    float* Pmine = p.pbuf + (long)blockIdx.x * ROWS * H;
    const float* Pclu = p.pbuf + (long)cid * CL * ROWS * H;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of the bias gradient (sum of dG over rows and steps)
    __syncthreads();

    // Saved activations of a step are fetched one step ahead (issued before the previous step's local GEMM), so their
    // L2 latency hides behind that GEMM instead of sitting in front of the cell backward.
    float4 g4v[RT];
    float cv[RT], pv[RT], doh[RT], doc[RT];
    auto step_geometry = [&](int s, int& t, int& na, int& off_t, int& na_src, int& off_src) {
        // walk the forward recurrence of this direction backwards
        t = dir == 0 ? Tt - 1 - s : s;
        na = active_rows(p.batch_sizes[t], tile, p.n_tiles);
        off_t = p.offsets[t];
        // "previous" frame in the forward recurrence of this direction (source of h_prev / c_prev)
        if (dir == 0) { na_src = t > 0 ? na : 0; off_src = t > 0 ? p.offsets[t - 1] : 0; }
        else { na_src = (t + 1 < Tt) ? active_rows(p.batch_sizes[t + 1], tile, p.n_tiles) : 0; off_src = (t + 1 < Tt) ? p.offsets[t + 1] : 0; }
    };
    auto prefetch = [&](int s) {
        int t, na, off_t, na_src, off_src;
        step_geometry(s, t, na, off_t, na_src, off_src);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = prg + i * RG;
            g4v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[i] = pv[i] = doh[i] = doc[i] = 0.f;
            if (r >= na) continue;
            long b = tile + p.n_tiles * r;
            long frame = off_t + b;
            g4v[i] = __ldg(reinterpret_cast<const float4*>(gs + frame * (4 * H) + unit * 4));
            const bool is_final = dir == 0 ? (t == len[i] - 1) : (t == 0);
            const float* dout = p.d_out + b * p.out_ld;
            if (CELL == 0) {
                cv[i] = __ldg(cs + frame * H + unit);
                if (r < na_src) pv[i] = __ldg(cs + ((long)off_src + b) * H + unit);
                if (is_final) { doh[i] = __ldg(dout + dir * 2 * H + unit); doc[i] = __ldg(dout + dir * 2 * H + H + unit); }
            } else {
                if (r < na_src) pv[i] = __ldg(hs + ((long)off_src + b) * H + unit);
                if (is_final) doh[i] = __ldg(dout + dir * H + unit);
            }
        }
    };
    if (Tt > 0) prefetch(0);

    for (int s = 0; s < Tt; ++s) {
        int t, na, off_t, na_src, off_src;
        step_geometry(s, t, na, off_t, na_src, off_src);
        (void)off_src;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = prg + i * RG;
            if (r >= na) continue;
            long b = tile + p.n_tiles * r;
            long frame = off_t + b;
            float dh = dh_sm[r * U + pu];
            float carry = dc_sm[r * U + pu];
            const float4 g4 = g4v[i];
            float4 dgv;
            if (CELL == 0) {
                float dc = carry;
                dh += doh[i]; dc += doc[i];
                float c = cv[i];
                float c_prev = pv[i];
                float tc = tanhf(c);
                float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
                dc += dh * og * (1.f - tc * tc);
                dgv.x = dc * gg * ig * (1.f - ig);
                dgv.y = dc * c_prev * fg * (1.f - fg);
                dgv.z = dc * ig * (1.f - gg * gg);
                dgv.w = dh * tc * og * (1.f - og);
                dc_sm[r * U + pu] = dc * fg;
            } else {
                dh += carry + doh[i];
                float h_prev = pv[i];
                float rg_ = g4.x, zg = g4.y, ng = g4.z, hn = g4.w;
                float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
                dgv.x = dn_pre * hn * rg_ * (1.f - rg_);
                dgv.y = dh * (h_prev - ng) * zg * (1.f - zg);
                dgv.z = dn_pre;
                dgv.w = dn_pre * rg_;
                dc_sm[r * U + pu] = dh * zg;
            }
            *reinterpret_cast<float4*>(dg + frame * (4 * H) + unit * 4) = dgv;
            *reinterpret_cast<float4*>(Gsm + r * GSP + pu * 4) = dgv;
            bsum.x += dgv.x; bsum.y += dgv.y; bsum.z += dgv.z; bsum.w += dgv.w;
        }
        __syncthreads();
        if (s + 1 < Tt) prefetch(s + 1);
        // partial dh_prev[rows, H] = dG[rows, my 4U] * W_hh[my 4U, H]; only rows that have a previous frame
        const long par = (long)(s & 1) * pb_stride;
        if (na_src > 0) local_gemm_store<U, GSP, ABCD_H>(Gsm, Wr, H / 4, na_src, Pmine + par, H, tid);
        cluster_sync_all();
        for (int idx = tid; idx < na_src * (U / 4); idx += NTHREADS) {
            const int r = idx / (U / 4), q = idx % (U / 4);
            float4 v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg4(Pclu + par + ((long)src * ROWS + r) * H + u0 + q * 4);
            float4 a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) { a.x += v[src].x; a.y += v[src].y; a.z += v[src].z; a.w += v[src].w; }
            *reinterpret_cast<float4*>(dh_sm + r * U + q * 4) = a;
        }
        __syncthreads();
    }
    // bias gradients: fixed-order reduction over the row groups, one row of partial sums per tile
    {
        float4* red = reinterpret_cast<float4*>(Gsm);
        red[prg * U + pu] = bsum;
        __syncthreads();
        if (tid < U) {
            float4 a = red[tid];
            for (int g = 1; g < NTHREADS / U; ++g) {
                const float4 b = red[g * U + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(p.dbias[dir] + (long)tile * (4 * H) + (u0 + tid) * 4) = a;
        }
    }
}

// Tensor-core version of the BPTT kernel: the K-split product  partial dh[rows, H] = dG[rows, my 4U] * W_hh[my 4U, H]  is
// computed transposed,  D^T[H units, rows] = W^T[H, 4U] * dG^T[4U, rows],  with W^T (two blocks of 128 units x 128 gate
// rows) resident in tensor memory as the A operand of tcgen05.mma (TS form, rnn_ts.cuh), its low part W - trunc_tf32(W)
// resident in shared memory, and the gate gradients -- produced by this CTA's own cell backward -- re-laid into hi/lo
// B-operand tiles straight from shared memory (no global round trip).  The recurrent carries dh / dc live in registers.
constexpr int EB_GSP = 4 * ENC_U + 4;                      // pitch of the dG slice in shared memory
constexpr int EB_HALF_K = 2 * TS_SLAB;                     // gate rows per sub-phase (the operand tiles hold half of K)
constexpr int EB_TILE_BYTES = 2 * TS_B_BYTES;              // two 32-k slabs of [k][rows] tiles (hi or lo)
constexpr int EB_DCOLS = 96;                               // accumulator columns per 128-unit block
size_t enc_bwd_ts_smem() { return (size_t)2 * EB_TILE_BYTES + ts_wlo_bytes(4 * ENC_U) * 2 + (size_t)ROWS * EB_GSP * sizeof(float); }

template <int CELL>
__global__ void __launch_bounds__(ENC_NT, 1) enc_bwd_ts_kernel(EncBwdK p) {
    constexpr int CL = ENC_CL, U = ENC_U, NTHREADS = ENC_NT;
    constexpr int KG = 4 * U;                              // 128 gate rows of this CTA = reduction length
    extern __shared__ __align__(1024) float smem[];
    const int H = ABCD_H;
    char* b_hi = reinterpret_cast<char*>(smem);            // [2 slabs][3 chunks][4 KB]
    char* b_lo = b_hi + EB_TILE_BYTES;
    char* wlo = b_lo + EB_TILE_BYTES;                      // [KG/32 slabs][8 chunks (256 units)][4 KB]
    float* Gsm = reinterpret_cast<float*>(wlo + (size_t)(KG / TS_SLAB) * (H / 32) * TS_CHUNK);   // [ROWS][EB_GSP], piece-swizzled
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL;
    const int tile = cid % p.n_tiles, dir = cid / p.n_tiles;
    const int u0 = rank * U;
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
        // gate row k (= uu*4 + slot) of this CTA -> row of W_hh (or -1)
        const float* W = p.w_hh[dir];
        auto wrow = [=](int k) -> const float* {
            const int uu = u0 + (k >> 2), s = k & 3;
            const int gate = CELL == 0 ? s : (s == 0 ? 0 : s == 1 ? 1 : s == 3 ? 2 : -1);
            return gate >= 0 ? W + (long)(gate * ABCD_H + uu) * ABCD_H : nullptr;
        };
        // hi: W^T into tensor memory, block b = warp >> 2, lane = unit within the block
        {
            const int b = warp >> 2, lq = warp & 3;
            const int j = b * 128 + lq * 32 + lane;
            for (int k0 = 0; k0 < KG; k0 += 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float* w = wrow(k0 + q);
                    r[q] = __float_as_uint(w ? __ldg(w + j) : 0.f);
                }
                const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(TS_A_COL + b * KG + k0);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                    ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        // lo: MN-major tiles [k][unit]
        for (int idx = tid; idx < KG * (H / 4); idx += NTHREADS) {
            const int j4 = idx % (H / 4), k = idx / (H / 4);
            const float* w = wrow(k);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w) v = umma::lo_of(__ldg(reinterpret_cast<const float4*>(w + j4 * 4)));
            *reinterpret_cast<float4*>(wlo + (size_t)(k / TS_SLAB) * (H / 32) * TS_CHUNK + umma::tile_off<TS_CHUNK>(k % TS_SLAB, j4)) = v;
        }
        umma::fence_proxy_async();
    }
    // pointwise mapping: thread -> (unit pu, rows prg + i*RG); a warp holds the 32 units of one row group
    const int pu = tid % U, prg = tid / U;
    const int unit = u0 + pu;
    int len[RT];
    float dh_reg[RT], dc_reg[RT];   // recurrent carries of this thread's (row, unit) pairs
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        int b = tile + p.n_tiles * (prg + i * RG);
        len[i] = b < p.B ? p.lengths[b] : 0;
        dh_reg[i] = 0.f; dc_reg[i] = 0.f;
    }
    const int Tt = p.lengths[tile];
    const float* gs = p.gs[dir];
    const float* cs = p.cs[dir];
    const float* hs = p.hs[dir];
    float* dg = p.dg[dir];
    const long pb_stride = (long)gridDim.x * ROWS * H;             // one parity buffer
    float* Pmine = p.pbuf + (long)blockIdx.x * ROWS * H;
    const float* Pclu = p.pbuf + (long)cid * CL * ROWS * H;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    float4 g4v[RT];
    float cv[RT], pv[RT], doh[RT], doc[RT];
    auto step_geometry = [&](int s, int& t, int& na, int& off_t, int& na_src, int& off_src) {
        t = dir == 0 ? Tt - 1 - s : s;
        na = active_rows(p.batch_sizes[t], tile, p.n_tiles);
        off_t = p.offsets[t];
        if (dir == 0) { na_src = t > 0 ? na : 0; off_src = t > 0 ? p.offsets[t - 1] : 0; }
        else { na_src = (t + 1 < Tt) ? active_rows(p.batch_sizes[t + 1], tile, p.n_tiles) : 0; off_src = (t + 1 < Tt) ? p.offsets[t + 1] : 0; }
    };
    auto prefetch = [&](int s) {
        int t, na, off_t, na_src, off_src;
        step_geometry(s, t, na, off_t, na_src, off_src);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = prg + i * RG;
            g4v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cv[i] = pv[i] = doh[i] = doc[i] = 0.f;
            if (r >= na) continue;
            long b = tile + p.n_tiles * r;
            long frame = off_t + b;
            g4v[i] = __ldg(reinterpret_cast<const float4*>(gs + frame * (4 * H) + unit * 4));
            const bool is_final = dir == 0 ? (t == len[i] - 1) : (t == 0);
            const float* dout = p.d_out + b * p.out_ld;
            if (CELL == 0) {
                cv[i] = __ldg(cs + frame * H + unit);
                if (r < na_src) pv[i] = __ldg(cs + ((long)off_src + b) * H + unit);
                if (is_final) { doh[i] = __ldg(dout + dir * 2 * H + unit); doc[i] = __ldg(dout + dir * 2 * H + H + unit); }
            } else {
                if (r < na_src) pv[i] = __ldg(hs + ((long)off_src + b) * H + unit);
                if (is_final) doh[i] = __ldg(dout + dir * H + unit);
            }
        }
    };
    if (Tt > 0) prefetch(0);

    for (int s = 0; s < Tt; ++s) {
        int t, na, off_t, na_src, off_src;
        step_geometry(s, t, na, off_t, na_src, off_src);
        (void)off_src;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            int r = prg + i * RG;
            if (r >= na) continue;
            long b = tile + p.n_tiles * r;
            long frame = off_t + b;
            float dh = dh_reg[i];
            float carry = dc_reg[i];
            const float4 g4 = g4v[i];
            float4 dgv;
            if (CELL == 0) {
                float dc = carry;
                dh += doh[i]; dc += doc[i];
                float c = cv[i];
                float c_prev = pv[i];
                float tc = tanhf(c);
                float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
                dc += dh * og * (1.f - tc * tc);
                dgv.x = dc * gg * ig * (1.f - ig);
                dgv.y = dc * c_prev * fg * (1.f - fg);
                dgv.z = dc * ig * (1.f - gg * gg);
                dgv.w = dh * tc * og * (1.f - og);
                dc_reg[i] = dc * fg;
            } else {
                dh += carry + doh[i];
                float h_prev = pv[i];
                float rg_ = g4.x, zg = g4.y, ng = g4.z, hn = g4.w;
                float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
                dgv.x = dn_pre * hn * rg_ * (1.f - rg_);
                dgv.y = dh * (h_prev - ng) * zg * (1.f - zg);
                dgv.z = dn_pre;
                dgv.w = dn_pre * rg_;
                dc_reg[i] = dh * zg;
            }
            dh_reg[i] = 0.f;   // consumed; rows without a previous frame restart from zero
            *reinterpret_cast<float4*>(dg + frame * (4 * H) + unit * 4) = dgv;
            *reinterpret_cast<float4*>(Gsm + r * EB_GSP + ((pu ^ ((r >> 3) & 3)) << 2)) = dgv;
            bsum.x += dgv.x; bsum.y += dgv.y; bsum.z += dgv.z; bsum.w += dgv.w;
        }
        __syncthreads();
        if (s + 1 < Tt) prefetch(s + 1);
        const long par = (long)(s & 1) * pb_stride;
        if (na_src > 0) {
            const int nq = (na_src + 3) >> 2;
            const int n_mma = (na_src + 15) & ~15;
            const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_ss = idesc_ts | (1u << 15);
            for (int half = 0; half < 2; ++half) {
                if (half == 1) umma::mbar_wait(&bar, (bar_uses - 1) & 1);   // the MMAs of the first half have read the tiles
                // re-lay gate rows [64 half, 64 half + 64) of the dG slice into [k][row] operand tiles (4x4 transposes)
                for (int idx = tid; idx < (ROWS / 4) * (EB_HALF_K / 4); idx += NTHREADS) {
                    const int q = idx % (ROWS / 4), kq = idx / (ROWS / 4);   // row quad, k quad within the half
                    if (q >= nq) continue;
                    const int pc = half * (EB_HALF_K / 4) + kq;            // 16-byte piece (= unit) within a Gsm row
                    const int sw = (q >> 1) & 3;
                    const float* rb = Gsm + (4 * q) * EB_GSP + ((pc ^ sw) << 2);
                    const float4 r0 = *reinterpret_cast<const float4*>(rb);
                    const float4 r1 = *reinterpret_cast<const float4*>(rb + EB_GSP);
                    const float4 r2 = *reinterpret_cast<const float4*>(rb + 2 * EB_GSP);
                    const float4 r3 = *reinterpret_cast<const float4*>(rb + 3 * EB_GSP);
                    const uint32_t base = (uint32_t)((kq >> 3) * TS_B_BYTES);
                    const int kl = (kq & 7) * 4;
                    umma::split_store(b_hi, b_lo, base + umma::tile_off<TS_CHUNK>(kl + 0, q), make_float4(r0.x, r1.x, r2.x, r3.x));
                    umma::split_store(b_hi, b_lo, base + umma::tile_off<TS_CHUNK>(kl + 1, q), make_float4(r0.y, r1.y, r2.y, r3.y));
                    umma::split_store(b_hi, b_lo, base + umma::tile_off<TS_CHUNK>(kl + 2, q), make_float4(r0.z, r1.z, r2.z, r3.z));
                    umma::split_store(b_hi, b_lo, base + umma::tile_off<TS_CHUNK>(kl + 3, q), make_float4(r0.w, r1.w, r2.w, r3.w));
                }
                umma::fence_proxy_async();
                umma::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    umma::fence_after_sync();
                    const uint64_t d_bh0 = umma::make_desc(umma::smem_u32(b_hi), TS_CHUNK, 512);
                    const uint64_t d_bl0 = umma::make_desc(umma::smem_u32(b_lo), TS_CHUNK, 512);
                    const uint64_t d_wl0 = umma::make_desc(umma::smem_u32(wlo), TS_CHUNK, 512);
#pragma unroll 1
                    for (int b = 0; b < 2; ++b) {
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
                            for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                                const uint64_t dbh = d_bh0 + (uint64_t)((sl * TS_B_BYTES + kg * 1024) >> 4);
                                const uint64_t dbl = d_bl0 + (uint64_t)((sl * TS_B_BYTES + kg * 1024) >> 4);
                                const uint64_t dwl = d_wl0 + (uint64_t)(((half * 2 + sl) * ((H / 32) * TS_CHUNK) + b * 4 * TS_CHUNK + kg * 1024) >> 4);
                                const uint32_t ta = tmem + (uint32_t)(TS_A_COL + b * KG + half * EB_HALF_K + sl * TS_SLAB + kg * 8);
                                const uint32_t td = tmem + (uint32_t)(b * EB_DCOLS);
                                mma_tf32_ts(td, ta, dbh, idesc_ts, (half > 0 || sl > 0 || kg > 0) ? 1u : 0u);
                                mma_tf32_ts(td, ta, dbl, idesc_ts, 1u);
                                umma::mma_tf32(td, dwl, dbh, idesc_ss, 1u);
                            }
                        }
                    }
                    umma::commit(&bar);
                }
                bar_uses += 1;
            }
            umma::mbar_wait(&bar, (bar_uses - 1) & 1);
            umma::fence_after_sync();
            // accumulators D^T[unit][row] -> partial dh[row][unit] (global, reduced by the owners after the barrier)
            {
                const int b = warp >> 2, lq = warp & 3;
                const int j = b * 128 + lq * 32 + lane;
#pragma unroll 1
                for (int r0 = 0; r0 < na_src; r0 += 16) {
                    uint32_t r[16];
                    umma::tmem_ld16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * EB_DCOLS + r0), r);
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj)
                        if (r0 + jj < na_src) Pmine[par + (long)(r0 + jj) * H + j] = __uint_as_float(r[jj]);
                }
            }
            umma::fence_before_sync();
        }
        cluster_sync_all();
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = prg + i * RG;
            if (r >= na_src) continue;
            float v[CL];
#pragma unroll
            for (int src = 0; src < CL; ++src) v[src] = ld_cg(Pclu + par + ((long)src * ROWS + r) * H + unit);
            float a = v[0];
#pragma unroll
            for (int src = 1; src < CL; ++src) a += v[src];
            dh_reg[i] = a;
        }
    }
    // bias gradients: fixed-order reduction over the row groups, one row of partial sums per tile
    {
        __syncthreads();
        float4* red = reinterpret_cast<float4*>(Gsm);
        red[prg * U + pu] = bsum;
        __syncthreads();
        if (tid < U) {
            float4 a = red[tid];
            for (int g = 1; g < NTHREADS / U; ++g) {
                const float4 b = red[g * U + tid];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(p.dbias[dir] + (long)tile * (4 * H) + (u0 + tid) * 4) = a;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

size_t enc_fwd_smem(bool tc) {
    if (tc) return (size_t)TSI_STAGES * TS_STAGE_BYTES + ts_wlo_bytes(ABCD_H);
    return (size_t)(ABCD_H * ENC_U * 4 + stage_floats(ENC_KC_FWD)) * sizeof(float);
}
// tensor-core recurrence (default); env ABCD_RNN_TC=0 or abcd_set_rnn_tensorcore(0) selects the FFMA path
}  // namespace
// process-wide debug switch (parity tests run the FFMA kernels through it); autograd runs the backward pass on another thread,
// so it cannot be thread-local.  Not part of the re-entrant compute surface: set it before issuing work from several threads.
static int g_rnn_tc = -1;
bool rnn_use_tc() {
    if (g_rnn_tc < 0) { const char* e = getenv("ABCD_RNN_TC"); g_rnn_tc = (e && e[0] == '0') ? 0 : 1; }
    return g_rnn_tc == 1;
}
namespace {
size_t enc_bwd_smem() { return (size_t)(4 * ENC_U * ABCD_H + 2 * ROWS * ENC_U + ROWS * (4 * ENC_U + 4)) * sizeof(float); }

template <class K, class P>
int launch_cluster(K kernel, int n_clusters, size_t smem, cudaStream_t stream, const P& params) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(n_clusters * ENC_CL);
    cfg.blockDim = dim3(ENC_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ENC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    abcd_note_launch(1);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
    return ABCD_OK;
}

}  // namespace

int rnn_prepare_kernel(const void* fn, size_t smem) {
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    return ABCD_OK;
}

// How many clusters of `fn` can be resident at once (diagnostics; DESIGN.md "cluster residency").
int rnn_max_active_clusters(const void* fn, size_t smem, int cluster, int threads, int* out) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cluster * 64);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUDA_TRY(cudaOccupancyMaxActiveClusters(out, fn, &cfg));
    return ABCD_OK;
}
int enc_max_clusters(int which, int* out) {
    const bool tc = rnn_use_tc();
    // the kernels the launchers pick with the current tensor-core setting
    const void* fn = which == 0 ? (tc ? (const void*)enc_fwd_kernel<0, true> : (const void*)enc_fwd_kernel<0, false>)
                                : (tc ? (const void*)enc_bwd_ts_kernel<0> : (const void*)enc_bwd_kernel<0>);
    const size_t smem = which == 0 ? enc_fwd_smem(tc) : (tc ? enc_bwd_ts_smem() : enc_bwd_smem());
    return rnn_max_active_clusters(fn, smem, ENC_CL, ENC_NT, out);
}

int rnn_set_tc(int on) {
    const int prev = rnn_use_tc() ? 1 : 0;
    if (on == 0 || on == 1) g_rnn_tc = on;
    return prev;
}

int enc_read_timing(unsigned long long* out5) {
    CUDA_TRY(cudaMemcpyFromSymbol(out5, g_enc_timing, 10 * sizeof(unsigned long long)));
    return ABCD_OK;
}

int enc_init() {
    TRY(rnn_prepare_kernel((const void*)enc_fwd_kernel<0, false>, enc_fwd_smem(false)));
    TRY(rnn_prepare_kernel((const void*)enc_fwd_kernel<1, false>, enc_fwd_smem(false)));
    TRY(rnn_prepare_kernel((const void*)enc_fwd_kernel<0, true>, enc_fwd_smem(true)));
    TRY(rnn_prepare_kernel((const void*)enc_fwd_kernel<1, true>, enc_fwd_smem(true)));
    TRY(rnn_prepare_kernel((const void*)enc_bwd_ts_kernel<0>, enc_bwd_ts_smem()));
    TRY(rnn_prepare_kernel((const void*)enc_bwd_ts_kernel<1>, enc_bwd_ts_smem()));
    TRY(rnn_prepare_kernel((const void*)enc_bwd_kernel<0>, enc_bwd_smem()));
    TRY(rnn_prepare_kernel((const void*)enc_bwd_kernel<1>, enc_bwd_smem()));
    return ABCD_OK;
}

// ---- workspace layout --------------------------------------------------------------------------
struct EncWs {
    float* bias_comb[2];
    char* himg[2];      // operand-tile images of h for the tensor-core forward kernel
    float* wih_pad[2];  // [4H][pad4(F)] interleaved, zero-padded copy of weight_ih
    float* xproj[2];
    float* hs[2];
    float* gs[2];
    float* cs[2];
    float* dg[2];
    float* dbias[2];
    float* pbuf;
    int* prev_fwd;
    int* prev_bwd;
    float* xpad;      // [sumT][pad4(F)] copy of the input with 16-byte aligned rows (bulk-copy friendly)
    float* partials;
    size_t partial_floats;
    size_t bytes;
};

static EncWs enc_carve(void* base, size_t cap, int cell, int F, int B, int sumT, int ndirs, int save) {
    Arena a(base, cap);
    EncWs w;
    memset(&w, 0, sizeof(w));
    const int H = ABCD_H;
    for (int d = 0; d < ndirs; ++d) {
        w.bias_comb[d] = a.take<float>(4 * H);
        w.wih_pad[d] = a.take<float>((size_t)4 * H * ((F + 3) & ~3));
        w.himg[d] = reinterpret_cast<char*>(a.take<float>((size_t)num_tiles(B) * 2 * ts_image_bytes(H) / sizeof(float)));
        w.xproj[d] = a.take<float>((size_t)sumT * 4 * H);
        w.hs[d] = a.take<float>(save ? (size_t)sumT * H : (size_t)2 * B * H);
        if (save) {
            w.gs[d] = a.take<float>((size_t)sumT * 4 * H);
            if (cell == 0) w.cs[d] = a.take<float>((size_t)sumT * H);
        }
    }
    w.xpad = a.take<float>((size_t)sumT * ((F + 3) & ~3));
    if (save) {
        // dG aliases xproj (dead after the forward recurrence) -- both [sumT][4H] per direction
        for (int d = 0; d < ndirs; ++d) w.dg[d] = w.xproj[d];
        for (int d = 0; d < ndirs; ++d) w.dbias[d] = a.take<float>((size_t)num_tiles(B) * 4 * H);
        w.pbuf = a.take<float>((size_t)2 * ndirs * num_tiles(B) * ENC_CL * ROWS * H);
        w.prev_fwd = a.take<int>(sumT);
        w.prev_bwd = a.take<int>(sumT);
        int s1 = sgemm_pick_splits(4 * H, H, sumT), s2 = sgemm_pick_splits(4 * H, F, sumT);
        // split-K partials of all four weight-gradient GEMMs (deferred, batched epilogues) + scratch of the column sums
        w.partial_floats = (size_t)ndirs * 4 * H * ((size_t)s1 * H + (size_t)s2 * F) + 64 + (size_t)4 * 600 * 64;
        w.partials = a.take<float>(w.partial_floats);
    }
    w.bytes = a.off;
    return w;
}

size_t enc_workspace_bytes(int cell, int F, int B, int sumT, int bidir, int save) {
    return enc_carve(nullptr, 0, cell, F, B, sumT, bidir ? 2 : 1, save).bytes;
}

int enc_forward(const EncParams& P, const PackedMeta& pk, const float* x, int ldx, float* last_hidden,
                int save, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ARG_CHECK(P.H == ABCD_H, "encoder: rnn_hidden_size %d unsupported (kernels are specialised for %d)", P.H, ABCD_H);
    ARG_CHECK(P.cell == 0 || P.cell == 1, "encoder: rnn_type must be LSTM or GRU");
    ARG_CHECK(pk.B > 0 && pk.T_max > 0 && pk.sumT > 0, "encoder: empty packed batch");
    const int ndirs = P.bidir ? 2 : 1;
    const int H = ABCD_H;
    EncWs w = enc_carve(ws, ws_bytes, P.cell, P.F, pk.B, pk.sumT, ndirs, save);
    if (w.bytes > ws_bytes) { abcd_set_error("encoder: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return ABCD_ERR_WORKSPACE; }
    const int n_tiles = num_tiles(pk.B);
    const int ngates = P.cell == 0 ? 4 : 3;

    EncFwdK k;
    memset(&k, 0, sizeof(k));
    k.n_tiles = n_tiles; k.B = pk.B; k.ndirs = ndirs; k.save = save;
    k.timing = getenv("ABCD_TIMING") ? atoi(getenv("ABCD_TIMING")) : 0;
    k.out_ld = ndirs * H * (P.cell == 0 ? 2 : 1);
    k.batch_sizes = pk.batch_sizes; k.offsets = pk.offsets; k.lengths = pk.lengths;
    k.out = last_hidden;
    // the input rows are re-homed with a 16-byte aligned pitch so the tensor-core GEMMs can stream them
    const int Fpad = (P.F + 3) & ~3;
    const bool use_pad = (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0);
    if (use_pad) {
        pad_rows_kernel<<<592, 256, 0, stream>>>(pk.sumT, P.F, Fpad, x, ldx, w.xpad);
        KERNEL_CHECK();
    }
    for (int d = 0; d < ndirs; ++d) {
        prep_ih_kernel<<<148, 256, 0, stream>>>(P.cell, P.F, Fpad, P.w_ih[d], P.b_ih[d], P.b_hh[d], w.wih_pad[d], w.bias_comb[d]);
        KERNEL_CHECK();
        // xproj[frame][u*4+slot] = x W_ih^T + bias (gate-interleaved by construction of the operands)
        GemmArgs g = gemm_args_zero();
        g.M = pk.sumT; g.N = 4 * H; g.K = (use_pad || ldx >= Fpad) ? Fpad : P.F;
        g.A = use_pad ? w.xpad : x; g.lda = use_pad ? Fpad : ldx; g.a_kmajor = 1;
        g.B = w.wih_pad[d]; g.ldb = Fpad; g.b_kmajor = 1;
        g.C = w.xproj[d]; g.ldc = 4 * H;
        g.bias = w.bias_comb[d];
        TRY(sgemm_launch(g, stream));
        k.w_hh[d] = P.w_hh[d]; k.b_hh[d] = P.b_hh[d];
        k.xproj[d] = w.xproj[d]; k.hs[d] = w.hs[d]; k.gs[d] = w.gs[d]; k.cs[d] = w.cs[d]; k.himg[d] = w.himg[d];
    }
    (void)ngates;
    {
        ProfScope ps(PROF_ENC_FWD, stream);
        const bool tc = rnn_use_tc();
        if (P.cell == 0) {
            if (tc) TRY(launch_cluster(enc_fwd_kernel<0, true>, n_tiles * ndirs, enc_fwd_smem(true), stream, k));
            else TRY(launch_cluster(enc_fwd_kernel<0, false>, n_tiles * ndirs, enc_fwd_smem(false), stream, k));
        } else {
            if (tc) TRY(launch_cluster(enc_fwd_kernel<1, true>, n_tiles * ndirs, enc_fwd_smem(true), stream, k));
            else TRY(launch_cluster(enc_fwd_kernel<1, false>, n_tiles * ndirs, enc_fwd_smem(false), stream, k));
        }
    }
    return ABCD_OK;
}

int enc_backward(const EncParams& P, const EncGrads& G, const PackedMeta& pk, const float* x, int ldx,
                 const float* d_last_hidden, void* ws, size_t ws_bytes, cudaStream_t stream) {
    ARG_CHECK(P.H == ABCD_H, "encoder: rnn_hidden_size %d unsupported", P.H);
    const int ndirs = P.bidir ? 2 : 1;
    const int H = ABCD_H;
    EncWs w = enc_carve(ws, ws_bytes, P.cell, P.F, pk.B, pk.sumT, ndirs, 1);
    if (w.bytes > ws_bytes) { abcd_set_error("encoder bwd: workspace too small"); return ABCD_ERR_WORKSPACE; }
    ColsumScratch colsum_scratch(w.partials, w.partial_floats);
    const int n_tiles = num_tiles(pk.B);

    EncBwdK k;
    memset(&k, 0, sizeof(k));
    k.n_tiles = n_tiles; k.B = pk.B; k.ndirs = ndirs;
    k.out_ld = ndirs * H * (P.cell == 0 ? 2 : 1);
    k.batch_sizes = pk.batch_sizes; k.offsets = pk.offsets; k.lengths = pk.lengths;
    k.d_out = d_last_hidden;
    k.pbuf = w.pbuf;
    for (int d = 0; d < ndirs; ++d) {
        k.w_hh[d] = P.w_hh[d]; k.hs[d] = w.hs[d]; k.gs[d] = w.gs[d]; k.cs[d] = w.cs[d]; k.dg[d] = w.dg[d]; k.dbias[d] = w.dbias[d];
    }
    {
        ProfScope ps(PROF_ENC_BWD, stream);
        if (rnn_use_tc()) {
            if (P.cell == 0) TRY(launch_cluster(enc_bwd_ts_kernel<0>, n_tiles * ndirs, enc_bwd_ts_smem(), stream, k));
            else TRY(launch_cluster(enc_bwd_ts_kernel<1>, n_tiles * ndirs, enc_bwd_ts_smem(), stream, k));
        } else {
            if (P.cell == 0) TRY(launch_cluster(enc_bwd_kernel<0>, n_tiles * ndirs, enc_bwd_smem(), stream, k));
            else TRY(launch_cluster(enc_bwd_kernel<1>, n_tiles * ndirs, enc_bwd_smem(), stream, k));
        }
    }

    prev_index_kernel<<<min(pk.T_max, 592), 128, 0, stream>>>(pk.T_max, pk.batch_sizes, pk.offsets, w.prev_fwd, w.prev_bwd);
    KERNEL_CHECK();
    EpilogueBatch epilogues(w.partials, w.partial_floats, stream);   // split-K epilogues + bias column sums: three launches at the end
    for (int d = 0; d < ndirs; ++d) {
        const GateMap map_ih = P.cell == 0 ? gate_map(H, 0, 1, 2, 3) : gate_map(H, 0, 1, 2, -1);
        const GateMap map_hh = P.cell == 0 ? gate_map(H, 0, 1, 2, 3) : gate_map(H, 0, 1, -1, 2);
        // dW_hh[gate rows][H] += dG^T * H_prev
        GemmArgs g = gemm_args_zero();
        g.M = 4 * H; g.N = H; g.K = pk.sumT;
        g.A = w.dg[d]; g.lda = 4 * H; g.a_kmajor = 0;
        g.B = w.hs[d]; g.ldb = H; g.b_kmajor = 0; g.b_kgather = d == 0 ? w.prev_fwd : w.prev_bwd;
        g.m_dst = map_hh;
        g.C = G.w_hh[d]; g.ldc = H; g.accumulate = 1;
        g.splits = sgemm_pick_splits(g.M, g.N, g.K); g.partials = w.partials;
        TRY(sgemm_launch(g, stream));
        // dW_ih += dG^T * X
        // the input rows are re-homed with a 16-byte aligned pitch so the tensor-core path can bulk-copy them
        const int Fpad = (P.F + 3) & ~3;
        // (the padded copy of x was made by the forward pass and lives in the same workspace)
        const bool use_pad = (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0);
        g.N = P.F; g.B = use_pad ? w.xpad : x; g.ldb = use_pad ? Fpad : ldx; g.b_kgather = nullptr;
        g.m_dst = map_ih; g.C = G.w_ih[d]; g.ldc = P.F;
        g.splits = sgemm_pick_splits(g.M, g.N, g.K);
        TRY(sgemm_launch(g, stream));
        // bias gradients: the BPTT kernel already reduced dG over the frames of every tile
        TRY(colsum_launch(w.dbias[d], 4 * H, n_tiles, 4 * H, 1.f, G.b_ih[d], 1, map_ih, stream));
        TRY(colsum_launch(w.dbias[d], 4 * H, n_tiles, 4 * H, 1.f, G.b_hh[d], 1, map_hh, stream));
    }
    return epilogues.flush();
}
