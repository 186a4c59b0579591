// scoring.cuh — face-histogram score lookup, smart discards and the keep/bank rule.
//
// Replaces, for one roll:
//   SCORE_TABLE / _evaluate_nb     src/farkle/game/scoring_lookup.py:123-172,244-278
//   decide_smart_discards          src/farkle/game/scoring.py:303-467
//   apply_discards                 src/farkle/game/scoring.py:548-578
//   _decide_continue               src/farkle/simulation/strategies.py:125-162
//
// Layout of the lookup in shared memory (15,136 bytes per CTA):
//   idxA[512] u8   packed 3-bit counts of faces 1,2,3 -> combo index 0..83
//   idxB[512] u8   packed 3-bit counts of faces 4,5,6 -> combo index 0..83
//   tab[84*84] u16 score/50 (7 bits) | used (3) | single_fives (2) | single_ones (2)
// A roll's histogram h = sum 1 << 3*(face-1) indexes it as
//   tab[idxA[h & 511] * 84 + idxB[h >> 9]].
#pragma once
#include <cstdint>

#include "../../include/farkle_b200.h"

namespace fb {

constexpr int LUT_COMBOS = 84;  // multisets of <= 6 dice over 3 faces = C(9,3)
constexpr int LUT_IDX = 512;
constexpr int LUT_TAB = LUT_COMBOS * LUT_COMBOS;
constexpr int LUT_BYTES = 2 * LUT_IDX + 2 * LUT_TAB;  // 15,136

struct ScoreLut {
    uint8_t idxA[LUT_IDX];
    uint8_t idxB[LUT_IDX];
    uint16_t tab[LUT_TAB];
};
static_assert(sizeof(ScoreLut) == LUT_BYTES, "lut layout");

struct RollScore {
    int score, used, sf, so;
};

// Rules of _evaluate_nb (scoring_lookup.py:123-172) on a count vector.
inline RollScore host_evaluate_counts(const int cin[6]) {
    int c[6];
    for (int i = 0; i < 6; i++) c[i] = cin[i];
    int ones = 0, pairs = 0, trips = 0, four = 0;
    for (int f = 0; f < 6; f++) {
        ones += c[f] == 1;
        pairs += c[f] == 2;
        trips += c[f] == 3;
        four += c[f] == 4;
    }
    if (ones == 6) return {1500, 6, 0, 0};
    if (pairs == 3) return {1500, 6, 0, 0};
    if (trips == 2) return {2500, 6, 0, 0};
    if (four && pairs) return {1500, 6, 0, 0};
    RollScore r{0, 0, 0, 0};
    for (int f = 0; f < 6; f++) {
        if (c[f] >= 3) {
            static const int kind[7] = {0, 0, 0, 0, 1000, 2000, 3000};
            r.score += c[f] == 3 ? (f == 0 ? 300 : 100 * (f + 1)) : kind[c[f]];
            r.used += c[f];
            c[f] = 0;
        }
    }
    r.so = c[0];
    r.sf = c[4];
    r.score += 100 * r.so + 50 * r.sf;
    r.used += r.so + r.sf;
    return r;
}

inline void host_build_lut(ScoreLut& lut) {
    int combo[LUT_COMBOS][3];
    int n = 0;
    for (int i = 0; i < LUT_IDX; i++) lut.idxA[i] = lut.idxB[i] = 0xFF;
    for (int a = 0; a <= 6; a++)
        for (int b = 0; a + b <= 6; b++)
            for (int c = 0; a + b + c <= 6; c++) {
                combo[n][0] = a; combo[n][1] = b; combo[n][2] = c;
                lut.idxA[a | (b << 3) | (c << 6)] = (uint8_t)n;
                lut.idxB[a | (b << 3) | (c << 6)] = (uint8_t)n;
                n++;
            }
    for (int i = 0; i < LUT_COMBOS; i++)
        for (int j = 0; j < LUT_COMBOS; j++) {
            int c[6] = {combo[i][0], combo[i][1], combo[i][2], combo[j][0], combo[j][1], combo[j][2]};
            uint16_t e = 0;
            if (c[0] + c[1] + c[2] + c[3] + c[4] + c[5] <= 6) {
                RollScore r = host_evaluate_counts(c);
                e = (uint16_t)((r.score / 50) | (r.used << 7) | (r.sf << 10) | (r.so << 12));
            }
            lut.tab[i * LUT_COMBOS + j] = e;
        }
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lut_lookup(const ScoreLut* lut, uint32_t hist) {
    const uint32_t a = lut->idxA[hist & 511u];
    const uint32_t b = lut->idxB[hist >> 9];
    return lut->tab[a * LUT_COMBOS + b];
}

// Strategy parameters as the kernels keep them: p0 = score_threshold,
// p1 = (uint16)dice_threshold | flags << 16.
__device__ __forceinline__ int strat_dice_threshold(uint32_t p1) { return (int)(int16_t)(p1 & 0xffffu); }
__device__ __forceinline__ bool strat_flag(uint32_t p1, uint32_t f) { return (p1 >> 16) & f; }

// Closed form of decide_smart_discards (scoring.py:369-467): the candidate
// multisets that survive _select_candidate's filter are exactly "drop a of the
// sf lone fives and b of the so lone ones"; re-scoring such a candidate through
// the table gives score-50a-100b with used-a-b dice (no special 6-dice pattern
// can appear because used != n).  Enumeration order and the strict '>' keep the
// reference's tie-breaking.  Returns d5 | d1 << 8.
__device__ __forceinline__ uint32_t smart_discards(int score, int used, int sf, int so, int n,
                                                   int ts, int st, uint32_t p1) {
    if (!strat_flag(p1, FB_SF_SMART_FIVE) || used == n || (sf | so) == 0) return 0u;
    const int dt = strat_dice_threshold(p1);
    const bool cs = strat_flag(p1, FB_SF_CONSIDER_SCORE);
    const bool cd = strat_flag(p1, FB_SF_CONSIDER_DICE);
    const bool both = cs && cd && strat_flag(p1, FB_SF_REQUIRE_BOTH);
    const bool fav_score = strat_flag(p1, FB_SF_FAVOR_SCORE);
    const int bmax = strat_flag(p1, FB_SF_SMART_ONE) ? so : 0;
    int best = -1;
    uint32_t pick = 0u;
    for (int a = 0; a <= sf; a++) {
        for (int b = 0; b <= bmax; b++) {
            const int cand = score - 50 * a - 100 * b;
            if (cand == 0) continue;  // score_lister drops non-scoring candidates
            const int sa = ts + cand;
            const int dl = n - used + a + b;
            const bool hit_s = cs && sa >= st;
            const bool hit_d = cd && dl <= dt;
            const bool bank = both ? (hit_s && hit_d) : (hit_s || hit_d);
            if (bank) continue;
            const int key = fav_score ? ((sa << 3) | dl) : ((dl << 27) | sa);
            if (key > best) {
                best = key;
                pick = (uint32_t)a | ((uint32_t)b << 8);
            }
        }
    }
    return pick;
}

// _decide_continue (strategies.py:125-162).
__device__ __forceinline__ bool decide_continue(int ts, int dice, int st, uint32_t p1) {
    const bool cs = strat_flag(p1, FB_SF_CONSIDER_SCORE);
    const bool cd = strat_flag(p1, FB_SF_CONSIDER_DICE);
    const bool want_s = cs && ts < st;
    const bool want_d = cd && dice > strat_dice_threshold(p1);
    if (cs && cd) return strat_flag(p1, FB_SF_REQUIRE_BOTH) ? (want_s || want_d) : (want_s && want_d);
    return cs ? want_s : (cd ? want_d : false);
}
#endif  // __CUDACC__

}  // namespace fb
