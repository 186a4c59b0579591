// rng.cuh — NumPy-exact SeedSequence -> PCG64DXSM -> Lemire dice on the device.
//
// Replaces, bit for bit, what the reference obtains from NumPy at
//   src/farkle/utils/random.py:156,188,225  (SeedSequence / PCG64DXSM / Generator)
//   src/farkle/game/engine.py:101           (Generator.integers(1, 7, size=n))
//   src/farkle/simulation/run_tournament.py:318 (Generator.permutation)
// NumPy's algorithms (numpy/random/bit_generator.pyx, src/pcg64/pcg64.h,
// src/distributions/distributions.c) are restated here from their published form.
#pragma once
#include <cstdint>

namespace fb {

// ---- SeedSequence ---------------------------------------------------------
constexpr uint32_t SS_INIT_A = 0x43b0d7e5u;
constexpr uint32_t SS_MULT_A = 0x931e8875u;
constexpr uint32_t SS_INIT_B = 0x8b51f9ddu;
constexpr uint32_t SS_MULT_B = 0x58f38dedu;
constexpr uint32_t SS_MIX_L = 0xca01f9ddu;
constexpr uint32_t SS_MIX_R = 0x4973f715u;

__host__ __device__ __forceinline__ uint32_t ss_hashmix(uint32_t v, uint32_t& hc) {
    v ^= hc;
    hc *= SS_MULT_A;
    v *= hc;
    v ^= v >> 16;
    return v;
}
__host__ __device__ __forceinline__ uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = SS_MIX_L * x - SS_MIX_R * y;
    r ^= r >> 16;
    return r;
}

// The 18-word coordinate entropy of src/farkle/utils/random.py:80-124.
struct Coord {
    uint32_t purpose;
    uint64_t root_seed, k, shuffle_index, pair_id, order, game_index, seat_index, replicate_index;
};

// Entropy pool of SeedSequence(coordinate_entropy(...)).  The loops are fully
// unrolled so the data-independent hash-constant chain folds at compile time.
__host__ __device__ __forceinline__ void ss_pool_coord(const Coord& c, uint32_t pool[4]) {
    uint32_t e[18];
    e[0] = 2u;  // RNG_SCHEME_VERSION, src/farkle/utils/random.py:13
    e[1] = c.purpose;
    const uint64_t v[8] = {c.root_seed, c.k,          c.shuffle_index, c.pair_id,
                           c.order,     c.game_index, c.seat_index,    c.replicate_index};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        e[2 + 2 * i] = (uint32_t)v[i];
        e[3 + 2 * i] = (uint32_t)(v[i] >> 32);
    }
    uint32_t hc = SS_INIT_A;
#pragma unroll
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(e[i], hc);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
#pragma unroll
    for (int s = 4; s < 18; s++)
#pragma unroll
        for (int d = 0; d < 4; d++) pool[d] = ss_mix(pool[d], ss_hashmix(e[s], hc));
}

// The same pool in two stages.  Words 0..11 (scheme, purpose, root_seed, k, shuffle_index, pair_id,
// order) are shared by every seat of a shuffle / H2H block, so the seed kernels hash them once
// per shuffle and every (game, seat) thread only absorbs the last six words.  The hash-constant
// chain is data independent: after n hashmix calls it is INIT_A * MULT_A^n.
__host__ __device__ constexpr uint32_t ss_hash_const_after(int calls) {
    uint32_t h = SS_INIT_A;
    for (int i = 0; i < calls; i++) h *= SS_MULT_A;
    return h;
}
constexpr int SS_PREFIX_WORDS = 12;
constexpr int SS_PREFIX_CALLS = 4 + 12 + 4 * (SS_PREFIX_WORDS - 4);  // 48
__host__ __device__ __forceinline__ void ss_pool_prefix(const Coord& c, uint32_t pool[4]) {
    const uint32_t e[SS_PREFIX_WORDS] = {2u, c.purpose, (uint32_t)c.root_seed, (uint32_t)(c.root_seed >> 32),
                                         (uint32_t)c.k, (uint32_t)(c.k >> 32), (uint32_t)c.shuffle_index,
                                         (uint32_t)(c.shuffle_index >> 32), (uint32_t)c.pair_id,
                                         (uint32_t)(c.pair_id >> 32), (uint32_t)c.order, (uint32_t)(c.order >> 32)};
    uint32_t hc = SS_INIT_A;
#pragma unroll
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(e[i], hc);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
#pragma unroll
    for (int s = 4; s < SS_PREFIX_WORDS; s++)
#pragma unroll
        for (int d = 0; d < 4; d++) pool[d] = ss_mix(pool[d], ss_hashmix(e[s], hc));
}
__host__ __device__ __forceinline__ void ss_pool_suffix(uint32_t pool[4], uint64_t game_index, uint64_t seat_index,
                                                        uint64_t replicate_index) {
    const uint32_t e[6] = {(uint32_t)game_index, (uint32_t)(game_index >> 32), (uint32_t)seat_index,
                           (uint32_t)(seat_index >> 32), (uint32_t)replicate_index,
                           (uint32_t)(replicate_index >> 32)};
    uint32_t hc = ss_hash_const_after(SS_PREFIX_CALLS);
#pragma unroll
    for (int s = 0; s < 6; s++)
#pragma unroll
        for (int d = 0; d < 4; d++) pool[d] = ss_mix(pool[d], ss_hashmix(e[s], hc));
}

// Generic pool for arbitrary entropy (fb_seedseq_generate).
__host__ __device__ inline void ss_pool_generic(const uint32_t* e, int n, uint32_t pool[4]) {
    uint32_t hc = SS_INIT_A;
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < n ? e[i] : 0u, hc);
    for (int s = 0; s < 4; s++)
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
    for (int s = 4; s < n; s++)
        for (int d = 0; d < 4; d++) pool[d] = ss_mix(pool[d], ss_hashmix(e[s], hc));
}

template <int N>
__host__ __device__ __forceinline__ void ss_generate(const uint32_t pool[4], uint32_t out[N]) {
    uint32_t hc = SS_INIT_B;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t v = pool[i & 3] ^ hc;
        hc *= SS_MULT_B;
        v *= hc;
        v ^= v >> 16;
        out[i] = v;
    }
}

// ---- PCG64DXSM --------------------------------------------------------------
constexpr uint64_t PCG_CHEAP_MULT = 0xda942042e4dd58b5ULL;
constexpr uint64_t PCG_DEF_MULT_HI = 2549297995355413924ULL;
constexpr uint64_t PCG_DEF_MULT_LO = 4865540595714422341ULL;

struct Pcg {
    uint64_t hi, lo;    // 128-bit LCG state
    uint64_t ihi, ilo;  // 128-bit increment (odd)
};

#ifdef __CUDA_ARCH__
#define FB_UMULHI(a, b) __umul64hi((a), (b))
#else
#define FB_UMULHI(a, b) ((uint64_t)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#endif

// state = state * MULT128 + inc with the full 128-bit default multiplier
// (seeding only; pcg_setseq_128_srandom_r).
__host__ __device__ __forceinline__ void pcg_step_full(Pcg& g) {
    uint64_t lo = g.lo * PCG_DEF_MULT_LO;
    uint64_t hi = FB_UMULHI(g.lo, PCG_DEF_MULT_LO) + g.lo * PCG_DEF_MULT_HI + g.hi * PCG_DEF_MULT_LO;
    uint64_t nlo = lo + g.ilo;
    g.hi = hi + g.ihi + (nlo < lo ? 1u : 0u);
    g.lo = nlo;
}

// PCG64DXSM(SeedSequence) seeding from generate_state(4, uint64) = 8 uint32 words.
__host__ __device__ __forceinline__ void pcg_seed_words(Pcg& g, const uint32_t w[8]) {
    const uint64_t w0 = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    const uint64_t w1 = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
    const uint64_t w2 = (uint64_t)w[4] | ((uint64_t)w[5] << 32);
    const uint64_t w3 = (uint64_t)w[6] | ((uint64_t)w[7] << 32);
    // inc = (initseq << 1) | 1, initseq = w2:w3
    g.ihi = (w2 << 1) | (w3 >> 63);
    g.ilo = (w3 << 1) | 1u;
    g.hi = 0;
    g.lo = 0;
    pcg_step_full(g);
    // state += initstate (w0:w1)
    uint64_t nlo = g.lo + w1;
    g.hi = g.hi + w0 + (nlo < g.lo ? 1u : 0u);
    g.lo = nlo;
    pcg_step_full(g);
}

__host__ __device__ __forceinline__ void pcg_seed_pool(Pcg& g, const uint32_t pool[4]) {
    uint32_t w[8];
    ss_generate<8>(pool, w);
    pcg_seed_words(g, w);
}

__host__ __device__ __forceinline__ void pcg_seed_coord(Pcg& g, const Coord& c) {
    uint32_t pool[4], w[8];
    ss_pool_coord(c, pool);
    ss_generate<8>(pool, w);
    pcg_seed_words(g, w);
}

// DXSM output of the current state (pcg_cm_random_r computes it from the
// pre-step state).
// Low 64 bits of a 64 x 64 product as three multiply-adds (one wide, two accumulating into its
// high word); the compiler's own expansion is four instructions (two independent cross products,
// the wide one and an add): one fewer issue slot per multiply matters more here than the shorter
// chain (FB_MUL64_PLAIN keeps the plain product).
__host__ __device__ __forceinline__ uint64_t mul64lo(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__) && !defined(FB_MUL64_PLAIN)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, lo, hi;\n\t"
        ".reg .u64 t;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "mul.wide.u32 t, a0, b0;\n\t"
        "mov.b64 {lo, hi}, t;\n\t"
        "mad.lo.u32 hi, a0, b1, hi;\n\t"
        "mad.lo.u32 hi, a1, b0, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
#else
    return a * b;
#endif
}
__host__ __device__ __forceinline__ uint64_t pcg_output(uint64_t hi, uint64_t lo) {
    uint64_t h = hi;
    h ^= h >> 32;
    h = mul64lo(h, PCG_CHEAP_MULT);
    h ^= h >> 48;
    h = mul64lo(h, lo | 1u);
    return h;
}
// state = state * CHEAP_MULT + inc (mod 2^128)
__host__ __device__ __forceinline__ void pcg_step(uint64_t& hi, uint64_t& lo, uint64_t ihi,
                                                  uint64_t ilo) {
#ifdef __CUDA_ARCH__
    // 128 x 64 -> 128 multiply-add as one carry chain over 32-bit limbs: seven multiplies
    // (ptxas fuses the lo/hi pairs into IMAD.WIDE) and no compare/select carry fix-ups,
    // which moves the step from the ALU pipe to the FMA pipe.
    const uint32_t s0 = (uint32_t)lo, s1 = (uint32_t)(lo >> 32), s2 = (uint32_t)hi, s3 = (uint32_t)(hi >> 32);
    const uint32_t i0 = (uint32_t)ilo, i1 = (uint32_t)(ilo >> 32), i2 = (uint32_t)ihi, i3 = (uint32_t)(ihi >> 32);
    const uint32_t m0 = (uint32_t)PCG_CHEAP_MULT, m1 = (uint32_t)(PCG_CHEAP_MULT >> 32);
    uint32_t r0, r1, r2, r3;
    asm("mad.lo.cc.u32   %0, %4, %8, %10;\n\t"   // r0  = lo(s0 m0) + i0
        "madc.hi.cc.u32  %1, %4, %8, %11;\n\t"   // r1  = hi(s0 m0) + i1 + c
        "madc.lo.cc.u32  %2, %5, %9, %12;\n\t"   // r2  = lo(s1 m1) + i2 + c
        "madc.hi.u32     %3, %5, %9, %13;\n\t"   // r3  = hi(s1 m1) + i3 + c
        "mad.lo.cc.u32   %1, %4, %9, %1;\n\t"    // r1 += lo(s0 m1)
        "madc.hi.cc.u32  %2, %4, %9, %2;\n\t"    // r2 += hi(s0 m1) + c
        "madc.lo.u32     %3, %6, %9, %3;\n\t"    // r3 += lo(s2 m1) + c
        "mad.lo.cc.u32   %1, %5, %8, %1;\n\t"    // r1 += lo(s1 m0)
        "madc.hi.cc.u32  %2, %5, %8, %2;\n\t"    // r2 += hi(s1 m0) + c
        "madc.lo.u32     %3, %7, %8, %3;\n\t"    // r3 += lo(s3 m0) + c
        "mad.lo.cc.u32   %2, %6, %8, %2;\n\t"    // r2 += lo(s2 m0)
        "madc.hi.u32     %3, %6, %8, %3;\n\t"    // r3 += hi(s2 m0) + c
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(m0), "r"(m1), "r"(i0), "r"(i1), "r"(i2), "r"(i3));
    lo = (uint64_t)r0 | ((uint64_t)r1 << 32);
    hi = (uint64_t)r2 | ((uint64_t)r3 << 32);
#else
    uint64_t plo = lo * PCG_CHEAP_MULT;
    uint64_t phi = FB_UMULHI(lo, PCG_CHEAP_MULT) + hi * PCG_CHEAP_MULT;
    uint64_t nlo = plo + ilo;
    hi = phi + ihi + (nlo < plo ? 1u : 0u);
    lo = nlo;
#endif
}

// Sequential view with NumPy's persistent 32-bit half buffer (pcg64_cm_next32).
struct PcgStream {
    Pcg g;
    uint32_t saved;
    bool has32;
    __host__ __device__ __forceinline__ uint64_t next64() {
        uint64_t o = pcg_output(g.hi, g.lo);
        pcg_step(g.hi, g.lo, g.ihi, g.ilo);
        return o;
    }
    __host__ __device__ __forceinline__ uint32_t next32() {
        if (has32) {
            has32 = false;
            return saved;
        }
        uint64_t n = next64();
        has32 = true;
        saved = (uint32_t)(n >> 32);
        return (uint32_t)n;
    }
    // Generator.integers(1, 7): Lemire on 32-bit draws, range 6.  NumPy redraws while
    // leftover < threshold = (2^32 - 6) % 6 = 4 (buffered_bounded_lemire_uint32; its
    // outer `leftover < 6` test is implied).  Returns 0..5; `words` counts the fresh
    // 64-bit outputs this die consumed.
    __host__ __device__ __forceinline__ uint32_t die0(uint32_t& words) {
        uint64_t m;
        do {
            words += has32 ? 0u : 1u;
            m = (uint64_t)next32() * 6u;
        } while ((uint32_t)m < 4u);
        return (uint32_t)(m >> 32);
    }
    // random_interval(max) for max < 2^32: masked rejection on 32-bit draws.
    __host__ __device__ __forceinline__ uint32_t interval(uint32_t max, uint32_t mask) {
        uint32_t v;
        do {
            v = next32() & mask;
        } while (v > max);
        return v;
    }
};

}  // namespace fb
