// Per-thread arithmetic of the BCR merge pass (bcr.cu: k_bcr_merge), host + device so that tests/emu can run exactly this code on
// the CPU.  A thread owns kWords x 4 consecutive output positions of a tile.  Inputs per word: the insert flags of its four
// positions (bytes 0/1), the inserted symbols (bytes, valid where flagged) and the staged old symbols (consecutive bytes: the old
// symbols of a tile keep their order).
#pragma once
#include <cstdint>
#if defined(__CUDACC__)
#define BCR_HD __host__ __device__ __forceinline__
#else
#define BCR_HD inline
#endif

namespace fmg {

BCR_HD uint32_t bcr_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
// PRMT in its default mode: byte i of the result = byte (selector nibble i) of the 8 bytes b:a
BCR_HD uint32_t bcr_byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = (uint64_t)b << 32 | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}
// selector that spreads consecutive bytes 0,1,2.. over the non-insert positions of a word; f4 = its four flags as bits
BCR_HD uint32_t bcr_spread_selector(uint32_t f4) {
    uint32_t sel = 0, nxt = 0;
    for (int t = 0; t < 4; ++t) if (!((f4 >> t) & 1)) sel |= (nxt++) << (4 * t);
    return sel;
}
BCR_HD uint32_t bcr_flag_bits(uint32_t flags) { return (flags * 0x01020408u) >> 24; }       // bytes 0/1 -> four bits
BCR_HD uint32_t bcr_flag_count(uint32_t flags) { return (flags * 0x01010101u) >> 24; }      // bytes 0/1 (or small sums) -> their sum

// One output word.  `staged` = the old symbols as 32-bit words, sb = staged byte of the word's first old symbol, spread = the
// 16-entry selector table.  Old symbols fill the positions that are not inserts in order; inserts come from `syms`.
BCR_HD uint32_t bcr_merge_word(const uint32_t *staged, int sb, uint32_t flags, uint32_t syms, const uint32_t *spread) {
    const uint32_t *src = staged + (sb >> 2);
    const uint32_t sel = spread[bcr_flag_bits(flags)] + (uint32_t)(sb & 3) * 0x1111u;
    const uint32_t ins = flags * 0xffu;
    return (bcr_byte_perm(src[0], src[1], sel) & ~ins) | (syms & ins);
}

// Bit planes of a thread's symbols (symbol 4w+b at bit 8b+w of each plane, kWords <= 8) -> one mask per base A,C,G,T (1..4)
template <int kWords>
BCR_HD void bcr_base_masks(const uint32_t *o, uint32_t eq[4]) {
    const uint32_t kLsb = 0x01010101u;
    uint32_t p0 = 0, p1 = 0, p2 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 0; w < kWords; ++w) {
        p0 |= (o[w] & kLsb) << w;
        p1 |= w >= 1 ? (o[w] & (kLsb << 1)) << (w - 1) : (o[w] >> 1) & kLsb;
        p2 |= w >= 2 ? (o[w] & (kLsb << 2)) << (w - 2) : (o[w] >> (2 - w)) & (kLsb << w);
    }
    eq[0] = p0 & ~p1 & ~p2; eq[1] = ~p0 & p1 & ~p2; eq[2] = p0 & p1 & ~p2; eq[3] = ~p0 & ~p1 & p2;
}
// packed A,C,G,T counts (4 x 16 bit) of the thread
BCR_HD uint64_t bcr_pack_counts(const uint32_t eq[4]) {
    return (uint64_t)(bcr_popc(eq[0]) | (bcr_popc(eq[1]) << 16)) | ((uint64_t)(bcr_popc(eq[2]) | (bcr_popc(eq[3]) << 16)) << 32);
}
// positions of the thread before byte (sh / 8) of word w, in the bit layout of the masks
BCR_HD uint32_t bcr_earlier_mask(int w, int sh) {
    const uint32_t kLsb = 0x01010101u;
    return (kLsb * ((1u << w) - 1u)) | ((kLsb << w) & ((1u << sh) - 1u));
}
// rank of the insert at byte (sh / 8) of word w among equal symbols of the tile: counts before the thread + equal symbols at
// earlier positions of the thread; 0 for a symbol that is not a base
BCR_HD uint32_t bcr_insert_rank(uint32_t word, int w, int sh, const uint32_t eq[4], uint64_t cnt_before) {
    const uint32_t c = (word >> sh) & 0xffu;
    if (c < 1 || c > 4) return 0;
    const uint32_t e = c == 1 ? eq[0] : c == 2 ? eq[1] : c == 3 ? eq[2] : eq[3];
    return (uint32_t)((cnt_before >> (16 * (c - 1))) & 0xffff) + bcr_popc(e & bcr_earlier_mask(w, sh));
}

}  // namespace fmg
