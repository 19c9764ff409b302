// Tensor-core gate contraction with the recurrent weights RESIDENT IN TENSOR MEMORY (tcgen05.mma, TS form).
//
//     D^T[128 gate columns, rows] = W^T[128, K] * h^T[K, rows]            (K = 256 hidden units of the previous step)
//
// The CTA's weight slice is the A operand: its TF32 "hi" part lives in tensor memory for the whole launch (lane = gate
// column, one 32-bit column per k: 256 of the 512 columns; written once with tcgen05.st -- the tensor core ignores the low
// 13 mantissa bits, so the raw fp32 weights are stored), its "lo" part (W - trunc_tf32(W)) lives in shared memory as
// MN-major operand tiles (128 KB).  Per step only the ACTIVATIONS are split: the staged rows of h are turned into hi / lo
// B-operand tiles (4x4 register transpose into the MN-major SWIZZLE_128B_BASE32B layout, rows along N) and every 8-wide
// k-group issues three MMAs into the same accumulator:
//     W_hi (tensor memory) x h_hi,   W_hi (tensor memory) x h_lo,   W_lo (shared memory) x h_hi
// i.e. the 3xTF32 split at fp32-grade accuracy with no per-step work on the weights (the SS-form predecessor of this file
// re-split the weights every step and lost to the FFMA path for that reason).  The accumulator is bounced through shared
// memory (transposed back to [row][gate column]) into the same per-thread tile phase_gemm() produces, so the fused cell
// update is shared with the FFMA path.
//
// Layout facts established with tools/ts_probe.py: the tensor-memory A operand is addressed [lane = m][column = k], the
// instruction descriptor must say A is K-major (MN-major is an illegal instruction), N may be any multiple of 16 <= 256.
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
constexpr int TS_RAW_PITCH = TS_SLAB + 4;                  // floats
constexpr int TS_RAW_RING = 3;
constexpr int TS_BOUNCE_PITCH = TS_M + 4;
constexpr int TS_D_COL = 0, TS_A_COL = 256;                // tensor-memory columns: accumulator, weights
constexpr int TS_D_COLS_M = 96;                            // two-source variant: accumulator in columns [0, 96), weights from 96
constexpr int TS_WORKERS = NTHREADS - 32;                  // warp 0 only issues MMAs
static_assert(ROWS % 16 == 0, "MMA N granularity");
static_assert(ROWS * TS_BOUNCE_PITCH * 4 <= TS_STAGES * TS_STAGE_BYTES, "accumulator bounce must fit in the activation tiles");

constexpr size_t ts_wlo_bytes(int K) { return (size_t)(K / TS_SLAB) * (TS_M / 32) * TS_CHUNK; }
constexpr size_t ts_tiles_bytes() { return (size_t)TS_STAGES * TS_STAGE_BYTES; }
constexpr size_t ts_raw_bytes() { return (size_t)TS_RAW_RING * ROWS * TS_RAW_PITCH * sizeof(float); }

struct TsCtx {
    char* tiles;          // TS_STAGES activation stages (1024-byte aligned)
    char* wlo;            // resident W_lo tiles: [K/32 slabs][4 chunks][4 KB]
    float* raw;           // TS_RAW_RING x ROWS x TS_RAW_PITCH staged fp32 rows
    uint64_t* bars;       // [0..TS_STAGES): stage consumed; [TS_STAGES]: phase done
    uint32_t tmem;
    uint32_t uses[TS_STAGES + 1];
};

__device__ __forceinline__ int ts_raw_off(int r, int p) { return r * TS_RAW_PITCH + ((p ^ ((r >> 3) & 3)) << 2); }

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

// acc += h[rows, K] * W[K, my 128 gate columns] for this thread's tile (same contract as phase_gemm: the thread (cg, rg)
// receives rows rg + i*RG, columns cg*4..cg*4+3).  rp(r) = pointer to the K activations of tile row r (or nullptr).
// `wait_cluster`: a cluster arrive is pending; the matching wait is executed before the first activation row is fetched.
template <int K, class RP>
__device__ __forceinline__ void phase_gemm_ts(float (&acc)[RT][4], TsCtx& c, RP rp, int nrows, const ThreadMap& tm, int tid) {
    static_assert(NTHREADS == 256 && K % TS_SLAB == 0, "role mapping");
    constexpr int NSL = K / TS_SLAB;
    constexpr int RAW_ROUNDS = (ROWS * (TS_SLAB / 4) + TS_WORKERS - 1) / TS_WORKERS;   // (row, piece) copies per worker per slab
    const int warp = tid >> 5, wt = tid - 32;
    const int nq = (nrows + 3) >> 2;
    const int n_mma = (nrows + 15) & ~15;                 // MMA N: active rows rounded up to 16
    // instruction descriptors: TS form (A in tensor memory, K-major), SS form for the W_lo product (A MN-major in smem)
    const uint32_t idesc_common = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(TS_M >> 4) << 24);
    const uint32_t idesc_ts = idesc_common, idesc_ss = idesc_common | (1u << 15);

    // ---- slab-invariant work assignment ----
    const float* src[RAW_ROUNDS]; int dst[RAW_ROUNDS]; int piece[RAW_ROUNDS];
#pragma unroll
    for (int j = 0; j < RAW_ROUNDS; ++j) {
        const int idx = wt + j * TS_WORKERS;
        const int r = idx >> 3, pc = idx & 7;
        dst[j] = -1; src[j] = nullptr; piece[j] = pc;
        if (wt >= 0 && r < nrows && idx < ROWS * 8) { dst[j] = ts_raw_off(r, pc); src[j] = rp(r); }
    }
    auto stage_raw = [&](int s) {
        if (s < NSL) {
            float* slot = c.raw + (s % TS_RAW_RING) * (ROWS * TS_RAW_PITCH);
#pragma unroll
            for (int j = 0; j < RAW_ROUNDS; ++j) {
                if (dst[j] < 0) continue;
                if (src[j]) cp_async16(slot + dst[j], src[j] + s * TS_SLAB + piece[j] * 4);
                else *reinterpret_cast<float4*>(slot + dst[j]) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };
    // split item of this worker: row quad q, k quad kq of a slab (20 x 8 = 160 items)
    const int q = wt % (ROWS / 4), kq = wt / (ROWS / 4);
    const bool split_on = wt >= 0 && kq < TS_SLAB / 4 && q < nq;
    uint32_t b_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b_off[j] = umma::tile_off<TS_CHUNK>((kq & 7) * 4 + j, q);
    const int raw0 = ts_raw_off(4 * q, kq & 7);

    // descriptors of stage 0 / slab 0; the others differ by constant address offsets (>> 4) in the low 14 bits
    const uint64_t d_bh0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
    const uint64_t d_wl0 = umma::make_desc(umma::smem_u32(c.wlo), TS_CHUNK, 512);

    for (int s = 0; s < TS_RAW_RING; ++s) stage_raw(s);
    cp_async_wait<TS_RAW_RING - 1>();
    __syncthreads();                                       // raw slab 0 visible to all

    for (int s = 0; s < NSL; ++s) {
        const int stg = s & 1;
        char* b_hi = c.tiles + stg * TS_STAGE_BYTES;
        char* b_lo = b_hi + TS_B_BYTES;
        if (wt >= 0) {
            if (c.uses[stg] > 0) umma::mbar_wait(&c.bars[stg], (c.uses[stg] - 1) & 1);   // MMAs of slab s-2 consumed this stage
            if (split_on) {
                const float* rb = c.raw + (s % TS_RAW_RING) * (ROWS * TS_RAW_PITCH);
                const float4 r0 = *reinterpret_cast<const float4*>(rb + raw0);
                const float4 r1 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 1, kq & 7));
                const float4 r2 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 2, kq & 7));
                const float4 r3 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 3, kq & 7));
                umma::split_store(b_hi, b_lo, b_off[0], make_float4(r0.x, r1.x, r2.x, r3.x));
                umma::split_store(b_hi, b_lo, b_off[1], make_float4(r0.y, r1.y, r2.y, r3.y));
                umma::split_store(b_hi, b_lo, b_off[2], make_float4(r0.z, r1.z, r2.z, r3.z));
                umma::split_store(b_hi, b_lo, b_off[3], make_float4(r0.w, r1.w, r2.w, r3.w));
            }
            umma::fence_proxy_async();
            cp_async_wait<TS_RAW_RING - 2>();              // this thread's pieces of raw slab s+1 have landed
        }
        umma::fence_before_sync();
        __syncthreads();                                   // tiles of slab s complete; raw slab s+1 visible; raw slab s free
        c.uses[stg] += 1;
        if (tid == 0) {
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
            umma::commit(&c.bars[stg]);
            if (s == NSL - 1) umma::commit(&c.bars[TS_STAGES]);
        }
        if (wt >= 0) stage_raw(s + TS_RAW_RING);           // refills raw[s % ring], read by the split above
    }
    cp_async_wait<0>();
    c.uses[TS_STAGES] += 1;
    umma::mbar_wait(&c.bars[TS_STAGES], (c.uses[TS_STAGES] - 1) & 1);   // every MMA of the phase has completed
    umma::fence_after_sync();

    // accumulator D^T[gate column][row] -> bounce[row][gate column] (reuses the activation tiles)
    float* bounce = reinterpret_cast<float*>(c.tiles);
    {
        const int lane = tid & 31;
        const int lq = warp & 3, half = warp >> 2;
        const int col = lq * 32 + lane;
        constexpr int HALF_ROWS = ((ROWS / 2 + 15) / 16) * 16;   // 48: rows [0,48) / [48,80)
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
    __syncthreads();   // bounce (= activation tiles) free for the next phase
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
// Two-source variant with M weight columns per CTA (M = 64 or 128): rows are the concatenation [rp0(r)[0,K0) | rp1(r)[0,K1)]
// (K0 a multiple of TS_SLAB, K1 a multiple of 4; the tail of the last slab is zero-filled).  Used by the decoder forward
// gate contraction, [h_{t-1} | x_t] * [W_hh | W_ih]^T.  `wait_slab` >= 1: a cluster arrive is pending and the matching wait
// is executed just before the first slab >= wait_slab is staged (the rows of rp1 are produced by other CTAs).
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

template <int M, class RP0, class RP1>
__device__ __forceinline__ void phase_gemm_ts2(float (&acc)[RT][4], TsCtx& c, int K0, RP0 rp0, int K1, RP1 rp1, int nrows,
                                               const ThreadMap& tm, int tid, int wait_slab = -1) {
    static_assert(NTHREADS == 256, "role mapping");
    constexpr int RAW_ROUNDS = (ROWS * (TS_SLAB / 4) + TS_WORKERS - 1) / TS_WORKERS;
    constexpr int BOUNCE_PITCH = M + 4;
    const int nsl = (K0 + K1 + TS_SLAB - 1) / TS_SLAB;
    const int warp = tid >> 5, wt = tid - 32;
    const int nq = (nrows + 3) >> 2;
    const int n_mma = (nrows + 15) & ~15;
    const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t idesc_ss = idesc_ts | (1u << 15);

    const float* src0[RAW_ROUNDS]; const float* src1[RAW_ROUNDS]; int dst[RAW_ROUNDS]; int piece[RAW_ROUNDS];
#pragma unroll
    for (int j = 0; j < RAW_ROUNDS; ++j) {
        const int idx = wt + j * TS_WORKERS;
        const int r = idx >> 3, pc = idx & 7;
        dst[j] = -1; src0[j] = nullptr; src1[j] = nullptr; piece[j] = pc;
        if (wt >= 0 && r < nrows && idx < ROWS * 8) { dst[j] = ts_raw_off(r, pc); src0[j] = rp0(r); src1[j] = rp1(r); }
    }
    auto stage_raw = [&](int s) {
        if (s < nsl) {
            float* slot = c.raw + (s % TS_RAW_RING) * (ROWS * TS_RAW_PITCH);
#pragma unroll
            for (int j = 0; j < RAW_ROUNDS; ++j) {
                if (dst[j] < 0) continue;
                const int k = s * TS_SLAB + piece[j] * 4;
                const float* sp = nullptr;
                if (k < K0) { if (src0[j]) sp = src0[j] + k; }
                else if (k - K0 < K1) { if (src1[j]) sp = src1[j] + (k - K0); }
                if (sp) cp_async16(slot + dst[j], sp);
                else *reinterpret_cast<float4*>(slot + dst[j]) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };
    const int q = wt % (ROWS / 4), kq = wt / (ROWS / 4);
    const bool split_on = wt >= 0 && kq < TS_SLAB / 4 && q < nq;
    uint32_t b_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b_off[j] = umma::tile_off<TS_CHUNK>((kq & 7) * 4 + j, q);

    const uint64_t d_bh0 = umma::make_desc(umma::smem_u32(c.tiles), TS_CHUNK, 512);
    const uint64_t d_wl0 = umma::make_desc(umma::smem_u32(c.wlo), TS_CHUNK, 512);
    bool waited = wait_slab < 0;
    for (int s = 0; s < TS_RAW_RING; ++s) {
        if (!waited && s >= wait_slab) { cluster_wait_acquire(); waited = true; }
        stage_raw(s);
    }
    cp_async_wait<TS_RAW_RING - 1>();
    __syncthreads();

    for (int s = 0; s < nsl; ++s) {
        const int stg = s & 1;
        char* b_hi = c.tiles + stg * TS_STAGE_BYTES;
        char* b_lo = b_hi + TS_B_BYTES;
        if (wt >= 0) {
            if (c.uses[stg] > 0) umma::mbar_wait(&c.bars[stg], (c.uses[stg] - 1) & 1);
            if (split_on) {
                const float* rb = c.raw + (s % TS_RAW_RING) * (ROWS * TS_RAW_PITCH);
                const float4 r0 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 0, kq & 7));
                const float4 r1 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 1, kq & 7));
                const float4 r2 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 2, kq & 7));
                const float4 r3 = *reinterpret_cast<const float4*>(rb + ts_raw_off(4 * q + 3, kq & 7));
                umma::split_store(b_hi, b_lo, b_off[0], make_float4(r0.x, r1.x, r2.x, r3.x));
                umma::split_store(b_hi, b_lo, b_off[1], make_float4(r0.y, r1.y, r2.y, r3.y));
                umma::split_store(b_hi, b_lo, b_off[2], make_float4(r0.z, r1.z, r2.z, r3.z));
                umma::split_store(b_hi, b_lo, b_off[3], make_float4(r0.w, r1.w, r2.w, r3.w));
            }
            umma::fence_proxy_async();
            cp_async_wait<TS_RAW_RING - 2>();
        }
        umma::fence_before_sync();
        __syncthreads();
        c.uses[stg] += 1;
        if (tid == 0) {
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
#ifndef TS2_SKIP2
                mma_tf32_ts(c.tmem, ta, dbl, idesc_ts, 1u);
#endif
#ifndef TS2_SKIP3
                umma::mma_tf32(c.tmem, dwl, dbh, idesc_ss, 1u);
#endif
            }
            umma::commit(&c.bars[stg]);
            if (s == nsl - 1) umma::commit(&c.bars[TS_STAGES]);
        }
        if (s + TS_RAW_RING < nsl && !waited && s + TS_RAW_RING >= wait_slab) { cluster_wait_acquire(); waited = true; }
        if (wt >= 0) stage_raw(s + TS_RAW_RING);
        else cp_async_commit();
    }
    if (!waited) cluster_wait_acquire();
    cp_async_wait<0>();
    c.uses[TS_STAGES] += 1;
    umma::mbar_wait(&c.bars[TS_STAGES], (c.uses[TS_STAGES] - 1) & 1);
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
    __syncthreads();
}

}  // namespace rnn
