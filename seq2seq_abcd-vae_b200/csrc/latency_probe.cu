// Measured floor of the per-timestep exchange every persistent recurrent kernel performs (bench.py "latency_roofline"):
// one cluster of CL CTAs x 256 threads runs `iters` dependent rounds of
//     mode 0:  barrier.cluster.arrive.release + barrier.cluster.wait.acquire
//     mode 1:  thread 0 of every CTA stores a token to global memory, cluster barrier, reads the token its neighbour wrote
//              (ld.global.cg) and derives the next token from it -- the publish -> barrier -> consume chain of one phase
//     mode 2:  mode 1 with a 32-float row per CTA written by one warp and 16 rows read back per thread (the fan-in of a partial-sum
//              reduction over a 16-CTA cluster)
// and reports nanoseconds per round from %globaltimer of CTA 0.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256, 1) cluster_exchange_kernel(float* slots /*[CL*32]*/, float* out_ns, int iters, int mode, int cl) {
    const int tid = threadIdx.x;
    const unsigned rank = cluster_ctarank();
    float token = 1.0f + rank;
    if (tid < 32) slots[rank * 32 + tid] = token;
    cluster_sync_all();
    const unsigned long long t0 = globaltimer_ns();
    for (int it = 0; it < iters; ++it) {
        if (mode == 1) {
            if (tid == 0) slots[rank * 32] = token;
        } else if (mode == 2) {
            if (tid < 32) slots[rank * 32 + tid] = token;
        }
        cluster_arrive_release();
        cluster_wait_acquire();
        if (mode == 1) {
            token = ld_cg(slots + ((rank + 1) % cl) * 32) * 0.5f + 1.0f;
        } else if (mode == 2) {
            float s = 0.f;
            for (int src = 0; src < cl; ++src) s += ld_cg(slots + src * 32 + (tid & 31));
            token = s * 0.03125f + 1.0f;
        }
        if (mode != 0) {
            // the next round overwrites the slots: everybody must have read them first (same as the kernels' second barrier)
            cluster_arrive_release();
            cluster_wait_acquire();
        }
    }
    const unsigned long long t1 = globaltimer_ns();
    if (blockIdx.x == 0 && tid == 0) {
        out_ns[0] = (float)((double)(t1 - t0) / (double)iters);
        out_ns[1] = token;
    }
}

}  // namespace

int measure_cluster_exchange(int cluster_size, int iters, int mode, float* d_out2, float* d_scratch, cudaStream_t stream) {
    ARG_CHECK(cluster_size >= 1 && cluster_size <= 16 && iters > 0 && mode >= 0 && mode <= 2 && d_out2 && d_scratch, "cluster exchange probe: bad argument");
    static bool prepared = false;
    if (!prepared) {
        CUDA_TRY(cudaFuncSetAttribute(cluster_exchange_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        prepared = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    abcd_note_launch(1);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, cluster_exchange_kernel, d_scratch, d_out2, iters, mode, cluster_size));
    return ABCD_OK;
}
