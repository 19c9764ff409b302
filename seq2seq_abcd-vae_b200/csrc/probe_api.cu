// extern "C" surface of libabcd_b200_probe.so (tools/ts_probe.py): layout probes that are not part of the product library.
#include <stdarg.h>
#include "common.cuh"

int umma_ts_probe(const float* A, const float* B, float* D, int K, int N, int variant, cudaStream_t stream);

static thread_local char g_err[512] = "";
void abcd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void abcd_note_launch(int) {}
void abcd_prof_begin(int, cudaStream_t) {}
void abcd_prof_end(int, cudaStream_t) {}

extern "C" const char* abcd_probe_last_error(void) { return g_err; }
/* D[128,N] = A[128,K] B[K,N] with the A operand of tcgen05.mma read from tensor memory (TS form), single-pass TF32;
 * variant bit 0 selects the a_major bit of the instruction descriptor */
extern "C" int abcd_debug_ts_probe(const float* A, const float* B, float* D, int K, int N, int variant, void* stream) {
    return umma_ts_probe(A, B, D, K, N, variant, (cudaStream_t)stream);
}
