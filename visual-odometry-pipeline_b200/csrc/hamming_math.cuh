// 256-bit Hamming distance as a carry-save adder tree over prefix-XOR descriptors (used by match_u8.cu; also compiled
// for the host by tests/host_math_shim.cpp, which checks it against a plain popcount on the CPU-only build box).
//
// Both descriptors are kept in a prefix-XOR form: words 2, 5 and 6 hold w0^w1^w2, w3^w4^w5 and w0^...^w6.  With
// x[w] = a[w] ^ b[w] on such words, x[2], x[5] and x[6] ARE the sum outputs of the first three full adders of the tree
// (sa = x0^x1^x2, sb = x3^x4^x5, sc = sa^sb^x6 in terms of the plain XOR words), and every carry follows from two adder
// inputs and the sum with one LOP3.  8 XOR + 3 carries + one (xor3, maj3) adder over the carries = 13 logic operations
// and 4 POPC per distance (a plain tree: 16), same result bit for bit:
//   distance = popc(sc) + popc(x7) + 2 popc(t) + 4 popc(f),   t = ca^cb^cc,  f = maj(ca, cb, cc).
#pragma once
#include <stdint.h>

#ifndef VO_HD
#define VO_HD __host__ __device__ __forceinline__
#endif

namespace vo {

VO_HD uint32_t hm_xor3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a ^ b ^ c;
#endif
}
VO_HD uint32_t hm_maj3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return (a & b) | (a & c) | (b & c);
#endif
}
// Carry of a full adder whose third input is known only through the sum: maj(x, y, s ^ x ^ y) = (x & y) | ((x ^ y) & ~s).
VO_HD uint32_t hm_carry_from_sum(uint32_t x, uint32_t y, uint32_t s) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(d) : "r"(x), "r"(y), "r"(s));
    return d;
#else
    return (x & y) | ((x ^ y) & ~s);
#endif
}

VO_HD void hamming_prefix_form(uint32_t (&w)[8]) {
    w[2] ^= w[0] ^ w[1];
    w[5] ^= w[3] ^ w[4];
    w[6] ^= w[2] ^ w[5];
}

// a, b in prefix-XOR form -> the four words whose weighted popcounts make up the distance.
struct HammingPlanes {
    uint32_t ones_a, ones_b, twos, fours;
};
VO_HD HammingPlanes hamming_planes(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t x[8];  // x[2], x[5], x[6] are the sums sa, sb, sc
#pragma unroll
    for (int w = 0; w < 8; ++w) x[w] = a[w] ^ b[w];
    const uint32_t ca = hm_carry_from_sum(x[0], x[1], x[2]);
    const uint32_t cb = hm_carry_from_sum(x[3], x[4], x[5]);
    const uint32_t cc = hm_carry_from_sum(x[2], x[5], x[6]);
    HammingPlanes p;
    p.ones_a = x[6];
    p.ones_b = x[7];
    p.twos = hm_xor3(ca, cb, cc);
    p.fours = hm_maj3(ca, cb, cc);
    return p;
}

}  // namespace vo
