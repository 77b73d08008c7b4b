// siphash.cuh -- SipHash-1-3 with a zero key == CPython >= 3.11 hash(str) under PYTHONHASHSEED=0
// (the k-mer order of the reference, main:85,94).
#pragma once
#include <cstdint>

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

#define SIPROUND                                                                                   \
    do {                                                                                           \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);                              \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                                                   \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                                                   \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);                              \
    } while (0)

// byte-pointer form (any address space); used off the hot path
__device__ __forceinline__ int64_t py_siphash13_bytes(const uint8_t *p, int len)
{
    if (len <= 0) return 0;
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    int n = len;
    while (n >= 8) {
        uint64_t m = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) m |= (uint64_t)p[q] << (8 * q);
        v3 ^= m;
        SIPROUND;
        v0 ^= m;
        p += 8; n -= 8;
    }
    uint64_t b = (uint64_t)len << 56;
    for (int q = 0; q < n; ++q) b |= (uint64_t)p[q] << (8 * q);
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    int64_t x = (int64_t)((v0 ^ v1) ^ (v2 ^ v3));
    return x == -1 ? -2 : x;
}
