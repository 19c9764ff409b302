// Tensor-core recurrent contractions with the weights RESIDENT IN TENSOR MEMORY (tcgen05.mma, TS form).
//
//     D^T[weight columns (lanes), rows] = W^T[columns, K] * act^T[K, rows]
//
// The CTA's weight slice is the A operand: its TF32 "hi" part lives in tensor memory for the whole launch (lane = output
// column, one 32-bit column per k; written once with tcgen05.st -- the tensor core ignores the low 13 mantissa bits, so the
// raw fp32 weights are stored), its "lo" part (W - trunc_tf32(W)) lives in shared memory as MN-major operand tiles.  Per
// step only the ACTIVATIONS are split into hi / lo B-operand tiles ([k][row], SWIZZLE_128B_BASE32B) and every 8-wide k-group
// issues three MMAs into the same accumulator:
//     W_hi (tensor memory) x a_hi,   W_hi (tensor memory) x a_lo,   W_lo (shared memory) x a_hi
// i.e. the 3xTF32 split at fp32-grade accuracy with no per-step work on the weights.
//
// This file holds the pieces shared by the kernels: weight set-up (ts_init, ts_init_m), the TS-form MMA wrapper, and the
// image-fed phases used by the forward kernels -- the producers of the activations write them as hi/lo operand-tile images
// to global memory, the consumer fetches whole tile stages with bulk copies (phase_gemm_ts_img, phase_gemm_ts_img2) and
// bounces the accumulator through shared memory (transposed back to [row][column]) into the per-thread tile phase_gemm()
// produces, so the fused cell update is shared with the FFMA path.  The backward kernels (enc.cu, dec.cu) re-lay their own
// dG / dHidPre slices from shared memory and issue the MMAs inline.
//
// Layout facts established with tools/ts_probe.py: the tensor-memory A operand is addressed [lane = m][column = k], the
// instruction descriptor must say A is K-major (MN-major is an illegal instruction), N may be any multiple of 16 <= 256;
// with M = 64 (TS and SS form alike) row m of A and of the accumulator lives in lane 32 (m / 16) + m % 16; M = 64
// instructions run at about half the rate of M = 128 ones.
#pragma once
#include "rnn_common.cuh"
#include "umma.cuh"

namespace rnn {

constexpr int TS_M = 128;                                  // gate columns per CTA = tensor-memory lanes
constexpr int TS_SLAB = 32;                                // k rows per slab = 4 MMA k-groups
constexpr int TS_CHUNK = TS_SLAB * 128;                    // bytes of one [32 k][32 mn] chunk
constexpr int TS_B_CHUNKS = (ROWS + 31) / 32;              // 96 >= ROWS rows addressable
constexpr int TS_B_BYTES = TS_B_CHUNKS * TS_CHUNK;         // one hi (or lo) activation tile
constexpr int TS_STAGE_BYTES = 2 * TS_B_BYTES;             // [h_hi | h_lo]
constexpr int TS_STAGES = 2;
constexpr int TS_BOUNCE_PITCH = TS_M + 4;
constexpr int TS_D_COL = 0, TS_A_COL = 256;                // tensor-memory columns: accumulator, weights
constexpr int TS_D_COLS_M = 96;                            // two-source variant: accumulator in columns [0, 96), weights from 96
static_assert(ROWS % 16 == 0, "MMA N granularity");
static_assert(ROWS * TS_BOUNCE_PITCH * 4 <= TS_STAGES * TS_STAGE_BYTES, "accumulator bounce must fit in the activation tiles");

constexpr size_t ts_wlo_bytes(int K) { return (size_t)(K / TS_SLAB) * (TS_M / 32) * TS_CHUNK; }
constexpr size_t ts_tiles_bytes() { return (size_t)TS_STAGES * TS_STAGE_BYTES; }

struct TsCtx {
    char* tiles;          // TS_STAGES activation stages (1024-byte aligned)
    char* wlo;            // resident W_lo tiles: [K/32 slabs][4 chunks][4 KB]
    float* raw;           // unused by the image-fed phases (kept for staged experiments)
    uint64_t* bars;       // [0..TS_STAGES): stage consumed; [TS_STAGES]: phase done
    uint32_t tmem;
    uint32_t uses[TS_STAGES + 1];
};

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// Allocates tensor memory, loads the weight slice (wrow(c) = pointer to the K weights of gate column c, or nullptr for an
// all-zero column) into tensor memory (hi) and shared memory (lo tiles).  Ends with a CTA barrier.
template <int K, class WRow>
__device__ __forceinline__ void ts_init(TsCtx& c, char* tiles, char* wlo, float* raw, uint64_t* bars, uint32_t* tmem_slot,
                                        WRow wrow, int tid) {
    c.tiles = tiles; c.wlo = wlo; c.raw = raw; c.bars = bars;
#pragma unroll
    for (int i = 0; i <= TS_STAGES; ++i) c.uses[i] = 0;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i <= TS_STAGES; ++i) umma::mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = *tmem_slot;
    if (warp < 4) {
        // hi part: raw fp32 rows of W^T into tensor memory, lane = gate column
        const float* w = wrow(warp * 32 + lane);
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t r[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w) v = __ldg(reinterpret_cast<const float4*>(w + k0 + 4 * q));
                r[4 * q] = __float_as_uint(v.x); r[4 * q + 1] = __float_as_uint(v.y);
                r[4 * q + 2] = __float_as_uint(v.z); r[4 * q + 3] = __float_as_uint(v.w);
            }
            const uint32_t taddr = c.tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(TS_A_COL + k0);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    // lo part: MN-major tiles [k][gate column]
    for (int idx = tid; idx < TS_M * (K / 4); idx += NTHREADS) {
        const int col = idx % TS_M, kq = idx / TS_M;
        const float* w = wrow(col);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w) v = umma::lo_of(__ldg(reinterpret_cast<const float4*>(w + kq * 4)));
        const float lo[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kq * 4 + j;
            char* slab = wlo + (size_t)(k / TS_SLAB) * (TS_M / 32) * TS_CHUNK;
            *reinterpret_cast<float*>(slab + umma::tile_off<TS_CHUNK>(k % TS_SLAB, col >> 2) + (col & 3) * 4) = lo[j];
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
}
__device__ __forceinline__ void ts_fini(TsCtx& c, int tid) {
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(c.tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// Image variant: the PRODUCERS of the activations write them (hi = raw fp32, lo = x - trunc_tf32(x)) straight in the
// B-operand tile layout to a global "image" ([K/32 slabs][hi 12 KB | lo 12 KB], one image per cluster and step parity);
// the consumer fetches whole 24 KB stages with two cp.async.bulk copies that complete on an mbarrier -- no staging ring, no
// per-step split pass, no generic->async proxy fence in the loop.  One thread produces (copies), one thread issues MMAs.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TSI_STAGES = 4;
constexpr size_t ts_image_bytes(int K) { return (size_t)(K / TS_SLAB) * TS_STAGE_BYTES; }   // one image (hi + lo of all slabs)

// byte offset of activation element (k, row) inside an image, hi part (the lo part is TS_B_BYTES further)
__device__ __forceinline__ size_t ts_image_off(int k, int row) {
    return (size_t)(k / TS_SLAB) * TS_STAGE_BYTES + umma::tile_off<TS_CHUNK>(k % TS_SLAB, row >> 2) + (size_t)(row & 3) * 4;
}
__device__ __forceinline__ void ts_image_store(char* image, int k, int row, float v) {
    char* p = image + ts_image_off(k, row);
    *reinterpret_cast<float*>(p) = v;
    *reinterpret_cast<float*>(p + TS_B_BYTES) = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}
// make this thread's image stores visible to the bulk-copy engine (async proxy) of other CTAs; call before the cluster arrive
__device__ __forceinline__ void ts_image_publish() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

struct TsImgCtx {
    char* tiles;          // TSI_STAGES stages of [h_hi | h_lo] (1024-byte aligned)
    char* wlo;
    uint64_t* full;       // [TSI_STAGES] bytes landed
    uint64_t* empty;      // [TSI_STAGES] MMAs consumed the stage
    uint64_t* done;       // phase done
    uint32_t tmem;
    uint32_t n_full[TSI_STAGES], n_empty[TSI_STAGES], n_done;   // completed phases (uniform bookkeeping)
};
__device__ __forceinline__ void ts_img_init_bars(TsImgCtx& c, char* tiles, char* wlo, uint64_t* bars, uint32_t tmem, int tid) {
    c.tiles = tiles; c.wlo = wlo; c.full = bars; c.empty = bars + TSI_STAGES; c.done = bars + 2 * TSI_STAGES; c.tmem = tmem;
#pragma unroll
    for (int i = 0; i < TSI_STAGES; ++i) { c.n_full[i] = 0; c.n_empty[i] = 0; }
    c.n_done = 0;
    if (tid == 0) {
        for (int i = 0; i < 2 * TSI_STAGES + 1; ++i) umma::mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
}

// acc += act[rows, K] * W[K, my 128 gate columns]; the activations come from `image` (see above).
template <int K>
__device__ __forceinline__ void phase_gemm_ts_img(float (&acc)[RT][4], TsImgCtx& c, const char* image, int nrows,
                                                  const ThreadMap& tm, int tid) {
    constexpr int NSL = K / TS_SLAB;
    const int warp = tid >> 5;
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(TS_M >> 4) << 24);
    const uint32_t idesc_ss = idesc_ts | (1u << 15);
    const uint32_t cbytes = (uint32_t)((n_mma + 31) / 32) * TS_CHUNK;   // bytes of the row chunks the MMAs read (per hi / lo tile)
    if (tid == 32) {
        // ---- copy thread: two bulk copies per slab into the ring ----
        for (int s = 0; s < NSL; ++s) {
            const int stg = s % TSI_STAGES;
            if (s >= TSI_STAGES) umma::mbar_wait(&c.empty[stg], (c.n_empty[stg] + (s / TSI_STAGES) - 1) & 1);
            const uint32_t bar = umma::smem_u32(&c.full[stg]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * cbytes) : "memory");
            const uint32_t dst = umma::smem_u32(c.tiles + (size_t)stg * TS_STAGE_BYTES);
            const char* src = image + (size_t)s * TS_STAGE_BYTES;
            // only the 32-row chunks that hold active rows: hi part, then lo part
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst), "l"(src), "r"(cbytes), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst + (uint32_t)TS_B_BYTES), "l"(src + TS_B_BYTES), "r"(cbytes), "r"(bar) : "memory");
        }
    } else if (tid == 0) {
        // ---- MMA thread ----
        const uint64_t d_bh0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
        const uint64_t d_wl0 = umma::make_desc(umma::smem_u32(c.wlo), TS_CHUNK, 512);
        for (int s = 0; s < NSL; ++s) {
            const int stg = s % TSI_STAGES;
            umma::mbar_wait(&c.full[stg], (c.n_full[stg] + s / TSI_STAGES) & 1);
            umma::fence_after_sync();
            const uint64_t sbh = d_bh0 + (uint64_t)((stg * TS_STAGE_BYTES) >> 4);
            const uint64_t swl = d_wl0 + (uint64_t)((s * (TS_M / 32) * TS_CHUNK) >> 4);
#pragma unroll
            for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                const uint64_t dbh = sbh + (uint64_t)((kg * 1024) >> 4);
                const uint64_t dbl = dbh + (uint64_t)(TS_B_BYTES >> 4);
                const uint64_t dwl = swl + (uint64_t)((kg * 1024) >> 4);
                const uint32_t ta = c.tmem + (uint32_t)(TS_A_COL + s * TS_SLAB + kg * 8);
                mma_tf32_ts(c.tmem + TS_D_COL, ta, dbh, idesc_ts, (s > 0 || kg > 0) ? 1u : 0u);
                mma_tf32_ts(c.tmem + TS_D_COL, ta, dbl, idesc_ts, 1u);
                umma::mma_tf32(c.tmem + TS_D_COL, dwl, dbh, idesc_ss, 1u);
            }
            umma::commit(&c.empty[stg]);
            if (s == NSL - 1) umma::commit(c.done);
        }
    }
    // uniform bookkeeping of the barrier phases this call went through
#pragma unroll
    for (int i = 0; i < TSI_STAGES; ++i) {
        const uint32_t uses = (NSL - i + TSI_STAGES - 1) / TSI_STAGES;   // slabs that used stage i
        c.n_full[i] += uses; c.n_empty[i] += uses;
    }
    umma::mbar_wait(c.done, c.n_done & 1);
    c.n_done += 1;
    umma::fence_after_sync();

    float* bounce = reinterpret_cast<float*>(c.tiles);
    {
        const int lane = tid & 31;
        const int lq = warp & 3, half = warp >> 2;
        const int col = lq * 32 + lane;
        constexpr int HALF_ROWS = ((ROWS / 2 + 15) / 16) * 16;
#pragma unroll 1
        for (int r0 = half * HALF_ROWS; r0 < min(ROWS, (half + 1) * HALF_ROWS); r0 += 16) {
            if (r0 >= nrows) break;
            uint32_t r[16];
            umma::tmem_ld16(c.tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(TS_D_COL + r0), r);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (r0 + j < nrows) bounce[(r0 + j) * TS_BOUNCE_PITCH + col] = __uint_as_float(r[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tm.active && tm.ks == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = tm.rg + i * RG;
            if (r < nrows) {
                const float4 v = *reinterpret_cast<const float4*>(bounce + r * TS_BOUNCE_PITCH + tm.cg * 4);
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
    }
    // the bounce aliases the ring: the next phase's bulk copies (async proxy) must not overtake these generic reads
    umma::fence_proxy_async();
    __syncthreads();
}

// Two-image variant with M weight columns (decoder forward): slabs [0, nsl0) come from `image0` (h_{t-1}), slabs
// [nsl0, nsl0 + nsl1) from `image1` (x_t).  `wait_cluster`: a cluster arrive is pending and image1 is only valid after the
// matching wait; every thread executes that wait, the copy and MMA threads after their first TSI_STAGES slabs so that the
// image0 part of the contraction overlaps the barrier.  Requires nsl0 >= TSI_STAGES.
template <int M>
__device__ __forceinline__ void phase_gemm_ts_img2(float (&acc)[RT][4], TsImgCtx& c, const char* image0, int nsl0,
                                                   const char* image1, int nsl1, int nrows, const ThreadMap& tm, int tid,
                                                   bool wait_cluster) {
    constexpr int BOUNCE_PITCH = M + 4;
    const int nsl = nsl0 + nsl1;
    const int warp = tid >> 5;
    // the ring may have been used through the generic proxy since the last call (staging buffers of other phases alias it)
    umma::fence_proxy_async();
    __syncthreads();
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t idesc_ss = idesc_ts | (1u << 15);
    const uint32_t cbytes = (uint32_t)((n_mma + 31) / 32) * TS_CHUNK;
    if (tid == 32) {
        for (int s = 0; s < nsl; ++s) {
            if (s == TSI_STAGES && wait_cluster) cluster_wait_acquire_divergent();
            const int stg = s % TSI_STAGES;
            if (s >= TSI_STAGES) umma::mbar_wait(&c.empty[stg], (c.n_empty[stg] + (s / TSI_STAGES) - 1) & 1);
            const uint32_t bar = umma::smem_u32(&c.full[stg]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * cbytes) : "memory");
            const uint32_t dst = umma::smem_u32(c.tiles + (size_t)stg * TS_STAGE_BYTES);
            const char* src = s < nsl0 ? image0 + (size_t)s * TS_STAGE_BYTES : image1 + (size_t)(s - nsl0) * TS_STAGE_BYTES;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst), "l"(src), "r"(cbytes), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst + (uint32_t)TS_B_BYTES), "l"(src + TS_B_BYTES), "r"(cbytes), "r"(bar) : "memory");
        }
    } else if (tid == 0) {
        const uint64_t d_bh0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
        const uint64_t d_wl0 = umma::make_desc(umma::smem_u32(c.wlo), TS_CHUNK, 512);
        for (int s = 0; s < nsl; ++s) {
            if (s == TSI_STAGES && wait_cluster) cluster_wait_acquire_divergent();
            const int stg = s % TSI_STAGES;
            umma::mbar_wait(&c.full[stg], (c.n_full[stg] + s / TSI_STAGES) & 1);
            umma::fence_after_sync();
            const uint64_t sbh = d_bh0 + (uint64_t)((stg * TS_STAGE_BYTES) >> 4);
            const uint64_t swl = d_wl0 + (uint64_t)((s * (M / 32) * TS_CHUNK) >> 4);
#pragma unroll
            for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                const uint64_t dbh = sbh + (uint64_t)((kg * 1024) >> 4);
                const uint64_t dbl = dbh + (uint64_t)(TS_B_BYTES >> 4);
                const uint64_t dwl = swl + (uint64_t)((kg * 1024) >> 4);
                const uint32_t ta = c.tmem + (uint32_t)(TS_D_COLS_M + s * TS_SLAB + kg * 8);
                mma_tf32_ts(c.tmem, ta, dbh, idesc_ts, (s > 0 || kg > 0) ? 1u : 0u);
                mma_tf32_ts(c.tmem, ta, dbl, idesc_ts, 1u);
                umma::mma_tf32(c.tmem, dwl, dbh, idesc_ss, 1u);
            }
            umma::commit(&c.empty[stg]);
            if (s == nsl - 1) umma::commit(c.done);
        }
    } else if (wait_cluster) {
        cluster_wait_acquire_divergent();
    }
#pragma unroll
    for (int i = 0; i < TSI_STAGES; ++i) {
        const uint32_t uses = (nsl - i + TSI_STAGES - 1) / TSI_STAGES;
        c.n_full[i] += uses; c.n_empty[i] += uses;
    }
    umma::mbar_wait(c.done, c.n_done & 1);
    c.n_done += 1;
    umma::fence_after_sync();

    float* bounce = reinterpret_cast<float*>(c.tiles);
    {
        const int lane = tid & 31;
        const int lq = warp & 3, half = warp >> 2;
        const int col = M == 128 ? lq * 32 + lane : (lane < 16 ? lq * 16 + lane : -1);   // M = 64: 16 accumulator rows per lane quadrant
        constexpr int HALF_ROWS = ((ROWS / 2 + 15) / 16) * 16;
#pragma unroll 1
        for (int r0 = half * HALF_ROWS; r0 < min(ROWS, (half + 1) * HALF_ROWS); r0 += 16) {
            if (r0 >= nrows) break;
            uint32_t r[16];
            umma::tmem_ld16(c.tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)r0, r);
            if (col >= 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (r0 + j < nrows) bounce[(r0 + j) * BOUNCE_PITCH + col] = __uint_as_float(r[j]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tm.active && tm.ks == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = tm.rg + i * RG;
            if (r < nrows) {
                const float4 v = *reinterpret_cast<const float4*>(bounce + r * BOUNCE_PITCH + tm.cg * 4);
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
    }
    umma::fence_proxy_async();
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Weight set-up for M weight columns per CTA (M = 64 or 128) and an arbitrary number of K slabs: used by the decoder forward
// gate contraction, [h_{t-1} | x_t] * [W_hh | W_ih]^T (phase_gemm_ts_img2 above).
// ---------------------------------------------------------------------------------------------------------------
template <int M>
constexpr size_t ts_wlo_bytes_m(int Ktot) { return (size_t)((Ktot + TS_SLAB - 1) / TS_SLAB) * (M / 32) * TS_CHUNK; }

// wval(col, k): weight of output column col (< M) at reduction index k (< nsl * TS_SLAB; zero beyond the real K).
template <int M, class WVal>
__device__ __forceinline__ void ts_init_m(TsCtx& c, char* tiles, char* wlo, float* raw, uint64_t* bars, uint32_t* tmem_slot,
                                          int nsl, WVal wval, int tid) {
    static_assert(M == 64 || M == 128, "tensor-memory block height");
    c.tiles = tiles; c.wlo = wlo; c.raw = raw; c.bars = bars;
#pragma unroll
    for (int i = 0; i <= TS_STAGES; ++i) c.uses[i] = 0;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i <= TS_STAGES; ++i) umma::mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = *tmem_slot;
    const int Kp = nsl * TS_SLAB;
    if (warp < 4) {
        // M = 128: row (output column) m lives in lane m.  M = 64 (tools/ts_probe.py): row m lives in lane 32 (m / 16) + m % 16,
        // for the A operand and for the accumulator alike; lanes 16..31 of every quadrant are unused (zero-filled here).
        const int col = M == 128 ? warp * 32 + lane : (lane < 16 ? warp * 16 + lane : -1);
        for (int k0 = 0; k0 < Kp; k0 += 16) {
            uint32_t r[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(col >= 0 ? wval(col, k0 + q) : 0.f);
            const uint32_t taddr = c.tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(TS_D_COLS_M + k0);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    for (int idx = tid; idx < M * Kp; idx += NTHREADS) {
        const int col = idx % M, k = idx / M;
        const float w = wval(col, k);
        const float lo = w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
        char* slab = wlo + (size_t)(k / TS_SLAB) * (M / 32) * TS_CHUNK;
        *reinterpret_cast<float*>(slab + umma::tile_off<TS_CHUNK>(k % TS_SLAB, col >> 2) + (col & 3) * 4) = lo;
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
}

// ---------------------------------------------------------------------------------------------------------------
// Pair variant (decoder forward, default): full-rate M = 128 instructions inside a 16-CTA cluster.
//
// A 16-way column split leaves 64 gate columns per CTA, and M = 64 instructions run at half rate.  Instead, the two CTAs
// of a PAIR (cluster ranks 2c, 2c+1) share the 128 gate columns of 32 hidden units and split the REDUCTION dimension
// [h_{t-1} | x_t] (K = H + Fp, in 32-row slabs) between them: rank 2c takes the first ceil(NS/2) slabs (all of them rows of
// h), rank 2c+1 the rest (the last h slabs and every x slab).  Each CTA keeps BOTH the hi and the lo part of its
// [128 columns x K/2] weight block in tensor memory (2 x 192 of the 512 columns), so every one of the three products of the
// 3xTF32 split is a TS-form instruction and no weight byte is read from shared memory in the time loop:
//     D^T[128, rows] (+)= W_hi x a_hi,  W_hi x a_lo,  W_lo x a_hi          (accumulator: columns [0, 96))
// The partial sums of the two CTAs are exchanged through L2: each CTA finalises the 16 units (64 accumulator lanes) it
// owns -- lanes [64 r', 64 r' + 64) for pair rank r' -- so it publishes the OTHER 64 lanes to its partner's receive
// buffer, a cluster barrier follows, and the owner adds the partner's partial to its own lanes while bouncing them
// (transposed to [row][gate column]) through shared memory into the per-thread tile of the cell update.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TSP_STAGES = 6;                              // one ring stage per slab of a CTA (<= 6 slabs): no reuse inside a phase
constexpr int TSP_MAX_SLABS = 6;
constexpr int TSP_A_HI = TS_D_COLS_M;                      // tensor-memory columns: accumulator [0, 96), W_hi, W_lo
constexpr int TSP_A_LO = TSP_A_HI + TSP_MAX_SLABS * TS_SLAB;
constexpr int TSP_BOUNCE_PITCH = 64 + 4;
static_assert(TSP_A_LO + TSP_MAX_SLABS * TS_SLAB <= 512, "tensor-memory budget of the pair variant");
static_assert(ROWS * TSP_BOUNCE_PITCH * 4 <= TSP_STAGES * TS_STAGE_BYTES, "accumulator bounce must fit in the activation ring");

struct TsPairCtx {
    char* tiles;          // TSP_STAGES stages of [a_hi | a_lo] (1024-byte aligned)
    uint64_t* full;       // [TSP_STAGES] bytes landed
    uint64_t* done;       // MMAs of the phase complete
    uint32_t tmem;
    uint32_t n_phase;     // phases completed (uniform): parity of every barrier
    int kh, nsl, slab0;   // pair rank, number of slabs of this CTA, first global slab
};

// Tensor memory allocation + resident weights (hi and lo) of this CTA's K half.  wval(col, k): weight of gate column col
// (< 128, of the pair) at reduction index k of [h | x] (zero beyond the real K).  Ends with a CTA barrier.
template <class WVal>
__device__ __forceinline__ void ts_pair_init(TsPairCtx& c, char* tiles, uint64_t* bars, uint32_t* tmem_slot, int kh, int nsl_total,
                                             WVal wval, int tid) {
    c.tiles = tiles; c.full = bars; c.done = bars + TSP_STAGES; c.n_phase = 0;
    const int ns0 = (nsl_total + 1) / 2;
    c.kh = kh; c.slab0 = kh ? ns0 : 0; c.nsl = kh ? nsl_total - ns0 : ns0;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i <= TSP_STAGES; ++i) umma::mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = *tmem_slot;
    if (warp < 4) {
        const int col = warp * 32 + lane;      // M = 128: gate column m lives in lane m
        const int kbase = c.slab0 * TS_SLAB;
        for (int k0 = 0; k0 < c.nsl * TS_SLAB; k0 += 16) {
            uint32_t rh[16], rl[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float w = wval(col, kbase + k0 + q);
                rh[q] = __float_as_uint(w);
                rl[q] = __float_as_uint(w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u));
            }
            const uint32_t lane_base = c.tmem + ((uint32_t)(warp * 32) << 16);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(lane_base + (uint32_t)(TSP_A_HI + k0)), "r"(rh[0]), "r"(rh[1]), "r"(rh[2]), "r"(rh[3]), "r"(rh[4]), "r"(rh[5]), "r"(rh[6]), "r"(rh[7]),
                  "r"(rh[8]), "r"(rh[9]), "r"(rh[10]), "r"(rh[11]), "r"(rh[12]), "r"(rh[13]), "r"(rh[14]), "r"(rh[15]) : "memory");
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(lane_base + (uint32_t)(TSP_A_LO + k0)), "r"(rl[0]), "r"(rl[1]), "r"(rl[2]), "r"(rl[3]), "r"(rl[4]), "r"(rl[5]), "r"(rl[6]), "r"(rl[7]),
                  "r"(rl[8]), "r"(rl[9]), "r"(rl[10]), "r"(rl[11]), "r"(rl[12]), "r"(rl[13]), "r"(rl[14]), "r"(rl[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
}
__device__ __forceinline__ void ts_pair_fini(TsPairCtx& c, int tid) {
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(c.tmem);
}

// acc += ([h | x][rows, K] * W[K, pair's 128 gate columns]) restricted to the 16 units (64 columns) this CTA owns, as the
// per-thread tile of make_map(tid, 16, .) (threads with tm.ks == 0).  himage / ximage: operand-tile images of h_{t-1} (8 slabs)
// and x_t; send / recv: receive buffer of the partner / of this CTA ([64 lanes][ROWS] floats in global memory).
// wait_cluster: a cluster arrive is pending; the x image is only valid after the matching wait, which every thread executes
// here (the copy and MMA threads after the h slabs are in flight, so that the h part of the contraction overlaps it).
__device__ __forceinline__ void phase_gemm_ts_pair(float (&acc)[RT][4], TsPairCtx& c, const char* himage, const char* ximage,
                                                   float* send, const float* recv, int nrows, const ThreadMap& tm, int tid,
                                                   bool wait_cluster) {
    constexpr int NSL_H = ABCD_H / TS_SLAB;
    const int warp = tid >> 5, lane = tid & 31;
    // the ring may have been used through the generic proxy since the last call (P2 staging and the bounce alias it)
    umma::fence_proxy_async();
    __syncthreads();
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t cbytes = (uint32_t)((n_mma + 31) / 32) * TS_CHUNK;
    const uint32_t par = c.n_phase & 1;
    // slabs of h come first in this CTA's list; the wait (if any) goes between the last h slab and the first x slab
    const int n_h = max(0, min(c.nsl, NSL_H - c.slab0));
    if (tid == 32) {
        for (int s = 0; s < c.nsl; ++s) {
            if (s == n_h && wait_cluster) cluster_wait_acquire_divergent();
            const int gs = c.slab0 + s;
            const uint32_t bar = umma::smem_u32(&c.full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * cbytes) : "memory");
            const uint32_t dst = umma::smem_u32(c.tiles + (size_t)s * TS_STAGE_BYTES);
            const char* src = gs < NSL_H ? himage + (size_t)gs * TS_STAGE_BYTES : ximage + (size_t)(gs - NSL_H) * TS_STAGE_BYTES;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst), "l"(src), "r"(cbytes), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst + (uint32_t)TS_B_BYTES), "l"(src + TS_B_BYTES), "r"(cbytes), "r"(bar) : "memory");
        }
        if (n_h == c.nsl && wait_cluster) cluster_wait_acquire_divergent();
    } else if (tid == 0) {
        const uint64_t d_b0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
        for (int s = 0; s < c.nsl; ++s) {
            if (s == n_h && wait_cluster) cluster_wait_acquire_divergent();
            umma::mbar_wait(&c.full[s], par);
            umma::fence_after_sync();
            const uint64_t sb = d_b0 + (uint64_t)((s * TS_STAGE_BYTES) >> 4);
#pragma unroll
            for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
                const uint64_t dbh = sb + (uint64_t)((kg * 1024) >> 4);
                const uint64_t dbl = dbh + (uint64_t)(TS_B_BYTES >> 4);
                const uint32_t ta_hi = c.tmem + (uint32_t)(TSP_A_HI + s * TS_SLAB + kg * 8);
                const uint32_t ta_lo = c.tmem + (uint32_t)(TSP_A_LO + s * TS_SLAB + kg * 8);
                mma_tf32_ts(c.tmem, ta_hi, dbh, idesc, (s > 0 || kg > 0) ? 1u : 0u);
                mma_tf32_ts(c.tmem, ta_hi, dbl, idesc, 1u);
                mma_tf32_ts(c.tmem, ta_lo, dbh, idesc, 1u);
            }
        }
        umma::commit(c.done);
        if (n_h == c.nsl && wait_cluster) cluster_wait_acquire_divergent();
    } else if (wait_cluster) {
        cluster_wait_acquire_divergent();
    }
    __syncwarp();
    umma::mbar_wait(c.done, par);
    c.n_phase += 1;
    umma::fence_after_sync();

    // accumulator lanes: quadrants 2 kh, 2 kh + 1 are this CTA's own units, the other two go to the partner.
    // Warps w and w + 4 share lane quadrant w & 3 and split the 16-row blocks: warp < 4 takes blocks 0..2, warp >= 4 blocks 3..4.
    const int lq = warp & 3;
    const bool own = (lq >> 1) == c.kh;
    const int col64 = (lq & 1) * 32 + lane;                // lane within the 64-lane half
    const int rb0 = warp < 4 ? 0 : 3, rb1 = warp < 4 ? 3 : 5;
    uint32_t keep[3][16];
#pragma unroll
    for (int bi = 0; bi < 3; ++bi) {
        const int rb = rb0 + bi;
        if (rb < rb1 && rb * 16 < n_mma) {
            umma::tmem_ld16(c.tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(rb * 16), keep[bi]);
            if (!own) {
                float4* dst = reinterpret_cast<float4*>(send + (size_t)col64 * ROWS + rb * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    dst[q] = make_float4(__uint_as_float(keep[bi][4 * q]), __uint_as_float(keep[bi][4 * q + 1]),
                                         __uint_as_float(keep[bi][4 * q + 2]), __uint_as_float(keep[bi][4 * q + 3]));
            }
        }
    }
    umma::fence_before_sync();
    cluster_sync_all();           // partner partials published / visible
    float* bounce = reinterpret_cast<float*>(c.tiles);
    if (own) {
#pragma unroll
        for (int bi = 0; bi < 3; ++bi) {
            const int rb = rb0 + bi;
            if (rb < rb1 && rb * 16 < n_mma) {
                const float* src = recv + (size_t)col64 * ROWS + rb * 16;
                float4 pv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) pv[q] = ld_cg4(src + 4 * q);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float a[4] = {pv[q].x, pv[q].y, pv[q].z, pv[q].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = rb * 16 + 4 * q + j;
                        if (r < nrows) bounce[r * TSP_BOUNCE_PITCH + col64] = __uint_as_float(keep[bi][4 * q + j]) + a[j];
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tm.active && tm.ks == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int r = tm.rg + i * RG;
            if (r < nrows) {
                const float4 v = *reinterpret_cast<const float4*>(bounce + r * TSP_BOUNCE_PITCH + tm.cg * 4);
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
    }
    // the bounce aliases the ring: later bulk copies (async proxy) must not overtake these generic accesses
    umma::fence_proxy_async();
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Pair variant, ONE STEP AHEAD (decoder forward, default for output_size <= 88).  The gate pre-activations of step t + 1 are
// h_t W_hh + x_t W_ih, and h_t is final a third of a step before x_t is: the recurrent product h_t W_hh (K = H, split over the
// pair as above: 4 slabs per CTA, hi and lo weights in tensor memory) is therefore issued right after the cluster barrier that
// publishes h_t and runs on the tensor pipe WHILE the CUDA cores compute the emission MLPs of step t; its partial-sum exchange
// between the two CTAs rides on the cluster barrier that step already has between the hidden and the output layers.  What is
// left on the critical path of step t + 1 is the small input product x_t W_ih (K = Fp, FFMA, one staged chunk).  The four
// pieces below are called at the points of the step where their inputs are ready:
//   ts_pair_fetch       (one thread)  bulk copies of this CTA's h slabs from the operand-tile image into the ring
//   ts_pair_issue       (one thread)  waits for the slabs, issues the 3xTF32 MMAs, commits
//   ts_pair_drain_send  (all threads) accumulator read-back: own 64 gate columns -> gpre [row][64 (+4)] in shared memory,
//                                     the partner's 64 columns -> its receive buffer in global memory;  a cluster barrier follows
//   ts_pair_drain_recv  (all threads) adds the partner's partial to gpre
// The ring is touched by the async proxy only (bulk copies in, MMAs out), so no proxy fences are needed in the loop.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ts_pair_fetch(TsPairCtx& c, const char* himage, int nrows) {
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t cbytes = (uint32_t)((n_mma + 31) / 32) * TS_CHUNK;
    for (int s = 0; s < c.nsl; ++s) {
        const uint32_t bar = umma::smem_u32(&c.full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * cbytes) : "memory");
        const uint32_t dst = umma::smem_u32(c.tiles + (size_t)s * TS_STAGE_BYTES);
        const char* src = himage + (size_t)(c.slab0 + s) * TS_STAGE_BYTES;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(dst), "l"(src), "r"(cbytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(dst + (uint32_t)TS_B_BYTES), "l"(src + TS_B_BYTES), "r"(cbytes), "r"(bar) : "memory");
    }
}
__device__ __forceinline__ void ts_pair_issue(TsPairCtx& c, int nrows) {
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t par = c.n_phase & 1;
    const uint64_t d_b0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
    for (int s = 0; s < c.nsl; ++s) {
        umma::mbar_wait(&c.full[s], par);
        umma::fence_after_sync();
        const uint64_t sb = d_b0 + (uint64_t)((s * TS_STAGE_BYTES) >> 4);
#pragma unroll
        for (int kg = 0; kg < TS_SLAB / 8; ++kg) {
            const uint64_t dbh = sb + (uint64_t)((kg * 1024) >> 4);
            const uint64_t dbl = dbh + (uint64_t)(TS_B_BYTES >> 4);
            const uint32_t ta_hi = c.tmem + (uint32_t)(TSP_A_HI + s * TS_SLAB + kg * 8);
            const uint32_t ta_lo = c.tmem + (uint32_t)(TSP_A_LO + s * TS_SLAB + kg * 8);
            mma_tf32_ts(c.tmem, ta_hi, dbh, idesc, (s > 0 || kg > 0) ? 1u : 0u);
            mma_tf32_ts(c.tmem, ta_hi, dbl, idesc, 1u);
            mma_tf32_ts(c.tmem, ta_lo, dbh, idesc, 1u);
        }
    }
    umma::commit(c.done);
}
constexpr int TSP_GP = 64 + 4;   // pitch of gpre
// send / recv: [ROWS][64] floats in global memory (row-major: the 32 lanes of a warp hold 32 consecutive gate columns, so every
// store and load instruction covers 128 contiguous bytes)
__device__ __forceinline__ void ts_pair_drain_send(TsPairCtx& c, float* gpre, float* send, int nrows, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int n_mma = (nrows + 15) & ~15;
    umma::mbar_wait(c.done, c.n_phase & 1);
    c.n_phase += 1;
    umma::fence_after_sync();
    // accumulator lanes: quadrants 2 kh, 2 kh + 1 are this CTA's own units, the other two go to the partner.
    // Warps w and w + 4 share lane quadrant w & 3 and split the 32-row blocks: warp < 4 takes rows [0, 32) and [64, 80), warp >= 4 [32, 64).
    const int lq = warp & 3;
    const bool own = (lq >> 1) == c.kh;
    const int col64 = (lq & 1) * 32 + lane;
    float* dst = own ? gpre + col64 : send + col64;
    const int pitch = own ? TSP_GP : 64;
#pragma unroll 1
    for (int blk = (warp < 4 ? 0 : 1); blk < 3; blk += 2) {
        const int r0 = blk * 32;
        if (r0 >= n_mma) break;
        const uint32_t taddr = c.tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)r0;
        if (r0 + 16 < nrows && r0 + 32 <= TS_D_COLS_M) {
            uint32_t v[32];
            umma::tmem_ld32(taddr, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (r0 + j < nrows) dst[(r0 + j) * pitch] = __uint_as_float(v[j]);
        } else {
            uint32_t v[16];
            umma::tmem_ld16(taddr, v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (r0 + j < nrows) dst[(r0 + j) * pitch] = __uint_as_float(v[j]);
        }
    }
    umma::fence_before_sync();
}
__device__ __forceinline__ void ts_pair_drain_recv(const TsPairCtx& c, float* gpre, const float* recv, int nrows, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int lq = warp & 3;
    if ((lq >> 1) != c.kh) return;
    const int col64 = (lq & 1) * 32 + lane;
    // the four receiving warps: columns by (lq & 1, lane), rows interleaved between warp < 4 and warp >= 4
#pragma unroll 1
    for (int r0 = (warp < 4 ? 0 : 8); r0 < nrows; r0 += 16) {
        float pv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[j] = r0 + j < nrows ? ld_cg(recv + (size_t)(r0 + j) * 64 + col64) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (r0 + j < nrows) gpre[(r0 + j) * TSP_GP + col64] += pv[j];
    }
}

}  // namespace rnn
