// _Py_HashDouble restated for device and host (the host build of tests/host_emul/ checks it against Python's own hash()).
#pragma once
#include <cstdint>
#include <math.h>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace pct {

// _Py_HashDouble (Python/pyhash.c) for finite doubles: value mod (2^61 - 1) with sign, -1 -> -2.
// Integer restatement (round 2; the frexp loop below was 25 % of the continuous candidates kernel): a normal double is M * 2^E with the 53-bit
// integer M = mantissa | 2^52 and E = exponent - 1075; modulo the Mersenne prime P = 2^61 - 1 a multiplication by 2^k is a rotation by k within
// 61 bits, and M < P, so the reduced value is rot61(M, E mod 61) — exactly what the loop accumulates 28 bits at a time.  Zero -> 0; subnormals,
// infinities and NaN (never coordinates) take the loop / CPython's special cases.  tests/test_host_emul_stability.py checks it against Python's hash().
__host__ __device__ __forceinline__ uint64_t hash_double_loop(double v);
__host__ __device__ __forceinline__ uint64_t hash_double(double v) {
    union { double d; uint64_t u; } cv;
    cv.d = v;
    const uint64_t bits = cv.u;
    const int ex = (int)((bits >> 52) & 0x7FF);
    if (ex == 0 || ex == 0x7FF) return hash_double_loop(v);  // zero / subnormal / inf / nan
    const uint64_t MOD = (1ull << 61) - 1;
    const uint64_t M = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    int r = (ex - 1075) % 61;
    if (r < 0) r += 61;
    const uint64_t x = r ? (((M << r) & MOD) | (M >> (61 - r))) : M;
    int64_t h = (bits >> 63) ? -(int64_t)x : (int64_t)x;
    if (h == -1) h = -2;
    return (uint64_t)h;
}
__host__ __device__ __forceinline__ uint64_t hash_double_loop(double v) {
    const uint64_t MOD = (1ull << 61) - 1;
    if (v == 0.0) return 0;
    int e;
    double m = frexp(v, &e);
    int sign = 1;
    if (m < 0) { sign = -1; m = -m; }
    uint64_t x = 0;
    while (m != 0.0) {
        x = ((x << 28) & MOD) | (x >> (61 - 28));
        m *= 268435456.0;
        e -= 28;
        uint64_t y = (uint64_t)m;
        m -= (double)y;
        x += y;
        if (x >= MOD) x -= MOD;
    }
    e = e >= 0 ? e % 61 : 61 - 1 - ((-1 - e) % 61);
    x = ((x << e) & MOD) | (x >> (61 - e));
    int64_t r = (int64_t)x * sign;
    if (r == -1) r = -2;
    return (uint64_t)r;
}

}  // namespace pct
