// scoring.cuh — face-histogram score lookup, smart discards and the keep/bank rule.
//
// Replaces, for one roll:
//   SCORE_TABLE / _evaluate_nb     src/farkle/game/scoring_lookup.py:123-172,244-278
//   decide_smart_discards          src/farkle/game/scoring.py:303-467
//   apply_discards                 src/farkle/game/scoring.py:548-578
//   _decide_continue               src/farkle/simulation/strategies.py:125-162
//
// Layout of the lookup in shared memory (18,592 bytes per CTA):
//   idxA[512] u8   packed 3-bit counts of faces 1,2,3 -> combo index 0..83
//   idxB[512] u8   packed 3-bit counts of faces 4,5,6 -> combo index 0..83
//   tab[84*84] u16 score/50 (7 bits) | used (3) | single_fives (2) | single_ones (2)
//   disc[3456] u8  smart-discard decision, see disc_index()
// A roll's histogram h = sum 1 << 3*(face-1) indexes it as
//   tab[idxA[h & 511] * 84 + idxB[h >> 9]].
#pragma once
#include <cstdint>

#include "../../include/farkle_b200.h"

namespace fb {

constexpr int LUT_COMBOS = 84;  // multisets of <= 6 dice over 3 faces = C(9,3)
constexpr int LUT_IDX = 512;
constexpr int LUT_TAB = LUT_COMBOS * LUT_COMBOS;
// discard table: [favor_score 2][require_both 2][all_singles 2][sf 3][bmax 3][xs 8][yd 6]
constexpr int DISC_INNER = 2 * 3 * 3 * 8 * 6;  // entries per (favor, both) pair = 864
constexpr int LUT_DISC = 4 * DISC_INNER;       // 3,456
constexpr int LUT_BYTES = 2 * LUT_IDX + 2 * LUT_TAB + LUT_DISC;  // 18,592

struct ScoreLut {
    uint8_t idxA[LUT_IDX];
    uint8_t idxB[LUT_IDX];
    uint16_t tab[LUT_TAB];
    uint8_t disc[LUT_DISC];
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
    // Smart-discard table.  A candidate "drop a lone fives and b lone ones" loses
    // u = a + 2b units of 50 points and frees D = a + b dice; whether it must bank depends
    // only on u < xs (score threshold still met) and D < yd (dice threshold still met),
    // and the preference key is (-u, D) or (D, -u).  So the whole search of
    // decide_smart_discards (scoring.py:303-467) is a function of seven small integers.
    for (int fav = 0; fav < 2; fav++)
        for (int both = 0; both < 2; both++)
            for (int excl = 0; excl < 2; excl++)
                for (int sf = 0; sf < 3; sf++)
                    for (int bm = 0; bm < 3; bm++)
                        for (int xs = 0; xs < 8; xs++)
                            for (int yd = 0; yd < 6; yd++) {
                                int best_k1 = -100, best_k2 = -100, pick = 0;
                                bool have = false;
                                for (int a = 0; a <= sf; a++)
                                    for (int b = 0; b <= bm; b++) {
                                        if (excl && a == sf && b == bm) continue;  // candidate scores 0
                                        const int u = a + 2 * b, D = a + b;
                                        const bool hit_s = u < xs, hit_d = D < yd;
                                        if (both ? (hit_s && hit_d) : (hit_s || hit_d)) continue;
                                        const int k1 = fav ? -u : D, k2 = fav ? D : -u;
                                        if (!have || k1 > best_k1 || (k1 == best_k1 && k2 > best_k2)) {
                                            have = true;
                                            best_k1 = k1;
                                            best_k2 = k2;
                                            pick = a | (b << 2);
                                        }
                                    }
                                const int idx = (fav * 2 + both) * DISC_INNER +
                                                ((excl * 3 + sf) * 3 + bm) * 48 + xs * 6 + yd;
                                lut.disc[idx] = (uint8_t)pick;
                            }
}

// A strategy whose keep/bank rule can never say "bank": dice_left >= 1 always exceeds a
// dice threshold <= 0, and that alone keeps it rolling when dice are the only criterion or
// when both criteria must be met to stop (strategies.py:125-162).  Its turns end only in a
// farkle, so a table of such strategies always runs to the safety limit.  Scheduling hint only.
inline
#ifdef __CUDACC__
__host__ __device__
#endif
bool never_banks(int dice_threshold, uint32_t flags) {
    const bool cs = flags & FB_SF_CONSIDER_SCORE, cd = flags & FB_SF_CONSIDER_DICE;
    return cd && dice_threshold <= 0 && (!cs || (flags & FB_SF_REQUIRE_BOTH));
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

// Per-strategy part of the discard-table index: (favor_score * 2 + require_both) * 864.
__device__ __forceinline__ uint32_t disc_base(uint32_t p1) {
    const bool both = strat_flag(p1, FB_SF_CONSIDER_SCORE) && strat_flag(p1, FB_SF_CONSIDER_DICE) &&
                      strat_flag(p1, FB_SF_REQUIRE_BOTH);
    return ((strat_flag(p1, FB_SF_FAVOR_SCORE) ? 2u : 0u) + (both ? 1u : 0u)) * DISC_INNER;
}

// decide_smart_discards (scoring.py:369-467) as ONE table lookup, branch free.
// The candidates that survive _select_candidate's filter are exactly "drop a of the sf
// lone fives and b of the so lone ones" (b = 0 unless smart_one); re-scoring one through
// the table gives score-50a-100b with used-a-b dice (no special 6-dice pattern can appear
// because used != n).  With X = ts + score - score_threshold and Y = dice_threshold -
// (n - used):  hit_score <=> consider_score and 50(a+2b) <= X,  hit_dice <=>
// consider_dice and a+b <= Y.  xs / yd below count how many values of a+2b / a+b hit.
// Returns d5 | d1 << 2.
__device__ __forceinline__ uint32_t smart_discards(const ScoreLut* lut, uint32_t dbase, int score,
                                                   int used, int sf, int so, int n, int ts,
                                                   int st, uint32_t p1) {
    const bool on = strat_flag(p1, FB_SF_SMART_FIVE) && used != n;
    const int sfi = on ? sf : 0;
    const int bm = (on && strat_flag(p1, FB_SF_SMART_ONE)) ? so : 0;
    int xs = min(max(ts + score - st + 50, 0), 399);
    xs = (xs * 1311) >> 16;  // floor(xs / 50) for 0 <= xs <= 399
    xs = strat_flag(p1, FB_SF_CONSIDER_SCORE) ? xs : 0;
    int yd = min(max(strat_dice_threshold(p1) - (n - used) + 1, 0), 5);
    yd = strat_flag(p1, FB_SF_CONSIDER_DICE) ? yd : 0;
    const int excl = score == 50 * sfi + 100 * bm ? 1 : 0;  // the all-discard candidate scores 0
    return lut->disc[dbase + ((excl * 3 + sfi) * 3 + bm) * 48 + xs * 6 + yd];
}

// _decide_continue (strategies.py:125-162), branch free: with only one threshold
// considered the other "want" is false, so OR gives the single-threshold answer.
__device__ __forceinline__ bool decide_continue(int ts, int dice, int st, uint32_t p1) {
    const bool cs = strat_flag(p1, FB_SF_CONSIDER_SCORE);
    const bool cd = strat_flag(p1, FB_SF_CONSIDER_DICE);
    const bool want_s = cs && ts < st;
    const bool want_d = cd && dice > strat_dice_threshold(p1);
    const bool and_mode = cs && cd && !strat_flag(p1, FB_SF_REQUIRE_BOTH);
    return and_mode ? (want_s && want_d) : (want_s || want_d);
}
#endif  // __CUDACC__

}  // namespace fb
