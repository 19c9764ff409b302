// Host side of Encoder.encode_dataset (ABCD-VAE/encode.py:38-55): the reference builds, per batch, a pandas frame of the
// [B, K] result, melts it into long format (category-major: all segments of category 0, then category 1, ...), merges the
// annotation columns and appends to a CSV.  At 1 M segments x 128 categories that is 128 M text rows, so the row formatter is
// native: one pass over the batch writes "<data_ix>,<k>,<value><annotation suffix>\n" straight into a caller-owned buffer.
//
// Values are float32 and are printed exactly the way pandas.to_csv prints a float32 column (numpy scalar str):
// the SHORTEST digit string that round-trips the float32, positional for 1e-4 <= |x| < 1e6 (at least one digit after the
// point), otherwise scientific with an exponent of at least two digits ("1e-05", "3.4e-12"); nan / inf / -inf as numpy
// spells them.  No CUDA in this file: it compiles with the host compiler nvcc drives.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../../include/abcd_b200.h"
#include "common.cuh"

namespace {

// decimal digits (no point, no sign) of the shortest round-trip representation of |v| and its decimal exponent:
// |v| = 0.d1 d2 ... dn x 10^decpt  ->  returns n
inline int shortest_digits(float v, char* digits, int* decpt) {
    char tmp[48];
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific);   // d[.ddd]e[+-]xx, shortest
    int n = 0;
    const char* p = tmp;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[n++] = *p;
    int e = 0;
    bool neg = false;
    ++p;
    if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
    for (; p < r.ptr; ++p) e = e * 10 + (*p - '0');
    *decpt = (neg ? -e : e) + 1;
    while (n > 1 && digits[n - 1] == '0') --n;   // "1.0e+00"-style padding never appears in shortest mode; be safe
    return n;
}

inline char* put_int(char* o, long long v) {
    auto r = std::to_chars(o, o + 24, v);
    return r.ptr;
}

inline char* put_f32(char* o, float v) {
    if (std::isnan(v)) { memcpy(o, "nan", 3); return o + 3; }
    if (std::isinf(v)) { if (v < 0) *o++ = '-'; memcpy(o, "inf", 3); return o + 3; }
    if (std::signbit(v)) { *o++ = '-'; v = -v; }
    if (v == 0.f) { memcpy(o, "0.0", 3); return o + 3; }
    char d[16];
    int decpt;
    const int n = shortest_digits(v, d, &decpt);
    const double a = (double)v;
    if (a < 1.e6 && a >= 1.e-4) {                  // numpy float32 scalar str: positional (TrimMode_LeaveOneZero) in [1e-4, 1e6)
        if (decpt <= 0) {
            *o++ = '0'; *o++ = '.';
            for (int i = 0; i < -decpt; ++i) *o++ = '0';
            memcpy(o, d, n); o += n;
        } else if (decpt >= n) {
            memcpy(o, d, n); o += n;
            for (int i = n; i < decpt; ++i) *o++ = '0';
            *o++ = '.'; *o++ = '0';
        } else {
            memcpy(o, d, decpt); o += decpt;
            *o++ = '.';
            memcpy(o, d + decpt, n - decpt); o += n - decpt;
        }
        return o;
    }
    // scientific, TrimMode_DptZeros ("1e-05", "1.2345678e-05"), exponent with at least two digits
    *o++ = d[0];
    if (n > 1) { *o++ = '.'; memcpy(o, d + 1, n - 1); o += n - 1; }
    int e = decpt - 1;
    *o++ = 'e';
    if (e < 0) { *o++ = '-'; e = -e; } else *o++ = '+';
    if (e < 10) *o++ = '0';
    return put_int(o, e);
}

}  // namespace

extern "C" int64_t abcd_csv_row_bound(int64_t max_suffix_bytes) {
    // 20 (int64) + 1 + 11 (int32) + 1 + 16 (float32: sign, 9 digits, point, e-xx) + suffix + newline
    return 20 + 1 + 11 + 1 + 16 + max_suffix_bytes + 1;
}

extern "C" int64_t abcd_csv_format_long(const float* h_vals, int32_t B, int32_t K, int64_t ld, const int64_t* h_data_ix,
                                        const char* h_suffix, const int64_t* h_suffix_off, char* h_out, int64_t out_cap) {
    if (!h_vals || !h_data_ix || !h_out || B < 0 || K < 0 || ld < K) {
        abcd_set_error("csv_format_long: bad argument");
        return ABCD_ERR_ARG;
    }
    int64_t max_suffix = 0;
    if (h_suffix && h_suffix_off)
        for (int b = 0; b < B; ++b) max_suffix = std::max<int64_t>(max_suffix, h_suffix_off[b + 1] - h_suffix_off[b]);
    if (out_cap < (int64_t)B * K * abcd_csv_row_bound(max_suffix)) {
        abcd_set_error("csv_format_long: output buffer too small (%lld bytes for %d x %d rows)", (long long)out_cap, B, K);
        return ABCD_ERR_WORKSPACE;
    }
    char* o = h_out;
    for (int k = 0; k < K; ++k) {
        char kbuf[16];
        char* ke = put_int(kbuf, k);
        *ke++ = ',';
        const int klen = (int)(ke - kbuf);
        for (int b = 0; b < B; ++b) {
            o = put_int(o, h_data_ix[b]);
            *o++ = ',';
            memcpy(o, kbuf, klen); o += klen;
            o = put_f32(o, h_vals[(int64_t)b * ld + k]);
            if (h_suffix && h_suffix_off) {
                const int64_t n = h_suffix_off[b + 1] - h_suffix_off[b];
                memcpy(o, h_suffix + h_suffix_off[b], n); o += n;
            }
            *o++ = '\n';
        }
    }
    return o - h_out;
}
