// k1_helpers.cuh -- tile geometry and device helpers shared by the extraction kernels (k1_extract.cu for
// assemblies, k1q_reads.cu for reads): TMA bulk copy + mbarrier, 2-bit packing, invalid-base / quality masks,
// 128-bit sliding-OR used for window validity.
#pragma once
#include "common.cuh"

namespace ska {

constexpr int K1_T = 4096;                       // window-start positions per CTA
constexpr int K1_THREADS = 256;
constexpr int K1_LEAD = 32;                      // staged bytes before the tile
constexpr int K1_TAIL = 96;                      // staged bytes after the tile (>= k for k <= 63, 32-aligned)
constexpr int K1_BYTES = K1_LEAD + K1_T + K1_TAIL;
constexpr int K1_G32 = K1_BYTES / 32;            // groups of 32 bases in the staged region
// host-packed batches (2-bit codes + invalid mask): staged positions [t0 - K1P_LEAD, t0 + K1_T + K1P_TAIL); both byte
// streams (4 and 8 positions per byte) then start and end on 16-byte boundaries, as the bulk copy needs
constexpr int K1P_LEAD = 128;
constexpr int K1P_TAIL = 128;
constexpr int K1P_BYTES = K1P_LEAD + K1_T + K1P_TAIL;      // positions staged
static_assert(K1P_LEAD == FRONT_PAD && (K1P_BYTES / 8) % 16 == 0 && K1P_LEAD >= K1_LEAD && K1P_TAIL >= K1_TAIL, "packed tile geometry");


struct B128 { uint64_t lo, hi; };
__device__ __forceinline__ B128 shr128(B128 v, int n) {   // 0 <= n < 64
    B128 r;
    r.lo = n ? (v.lo >> n) | (v.hi << (64 - n)) : v.lo;
    r.hi = v.hi >> n;
    return r;
}
__device__ __forceinline__ B128 or128(B128 a, B128 b) { return {a.lo | b.lo, a.hi | b.hi}; }
// bit n of result = OR of v[n .. n+len-1]
__device__ __forceinline__ B128 slide_or(B128 v, int len) {
    int covered = 1;
    while (covered * 2 <= len) { v = or128(v, shr128(v, covered)); covered *= 2; }
    if (covered < len) v = or128(v, shr128(v, len - covered));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    } while (!done);
}

// 4 ASCII bases in a u32 -> 8 bits of 2-bit codes, base i at bits 2i   ((b >> 1) & 3, bit_encoding.rs:34-36)
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    uint32_t c = (w >> 1) & 0x03030303u;
    uint32_t x = c | (c >> 6);
    return (x | (x >> 12)) & 0xFFu;
}
// 4 ASCII bases -> 4 bits, bit i set iff base i is invalid ((b & 0xF) == 14, bit_encoding.rs:52-54)
__device__ __forceinline__ uint32_t bad4(uint32_t w) {
    uint32_t t = (w & 0x0F0F0F0Fu) ^ 0x0E0E0E0Eu;            // zero byte <=> invalid
    uint32_t z = ~(t + 0x7F7F7F7Fu) & 0x80808080u;           // bit 7 of byte set <=> byte was zero
    uint32_t m = z >> 7;
    m |= m >> 7;
    return (m | (m >> 14)) & 0xFu;
}
// 4 quality bytes -> 4 bits, bit i set iff NOT (qual - 33 > min_qual)   (split_kmer.rs:66-71, u8 wrap).  0xFF marks a base of a
// record that has no quality string (a FASTA file inside a read sample): `None => true` there
__device__ __forceinline__ uint32_t badq4(uint32_t q, uint32_t min_qual) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t v = ((q >> (8 * i)) - 33u) & 0xFFu;
        r |= ((v > min_qual || v == 222u) ? 0u : 1u) << i;              // byte 0xFF = "record without quality": always valid
    }
    return r;
}
// 16 bases LE (base i at bits 2i) -> BE (base i at bits 30-2i)
__device__ __forceinline__ uint32_t le_to_be(uint32_t x) {
    uint32_t y = __brev(x);
    return ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
}


}  // namespace ska
