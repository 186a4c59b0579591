// scoring.cuh — face-histogram score lookup, smart discards and the keep/bank rule.
//
// Replaces, for one roll:
//   SCORE_TABLE / _evaluate_nb     src/farkle/game/scoring_lookup.py:123-172,244-278
//   decide_smart_discards          src/farkle/game/scoring.py:303-467
//   apply_discards                 src/farkle/game/scoring.py:548-578
//   _decide_continue               src/farkle/simulation/strategies.py:125-162
//
// Layout of the lookup in shared memory (28,496 bytes per CTA):
//   rowA[512] u16     packed 3-bit counts of faces 1,2,3 -> offset of that combination's ROW in tab
//   colB[512] u8      packed 3-bit counts of faces 4,5,6 -> column inside the row
//   tab[3][924] u32   one copy per smart-discard variant of the strategy (0 none, 1 smart five,
//                     2 smart five + one): score/50 (7 bits) | used (3) | single_fives (2) |
//                     single_ones (2) | bits 16..25 the roll-dependent part of the discard-table
//                     index, premultiplied (see disc_index)
//   disc[16*864] u8   smart-discard decision
//   hist3[512] u32    three queued face codes -> packed face counts (play.cuh, face queue)
// A roll's histogram h = sum 1 << 3*(face-1) indexes it as
//   tab[variant][rowA[h & 511] + colB[h >> 9]].
// Only the 924 multisets of at most six dice exist, so the table is triangular: the 84 (c4,c5,c6)
// combinations are numbered by ascending dice count, a (c1,c2,c3) combination that uses s dice owns
// a row of C(9-s,3) entries (the combinations that fit in the remaining 6-s dice are exactly a
// prefix of that numbering), and the rows are laid end to end: 84+3*56+6*35+10*20+15*10+21*4+28 =
// 924 entries instead of 84*84.  Same two index loads and one table load as the square table, and
// the row offset replaces the multiply by 84; what it buys is 73 KB of shared memory per CTA
// (used by the two-seat kernel for per-seat home slots, play.cuh).
#pragma once
#include <cstdint>

#include "../../include/farkle_b200.h"

namespace fb {

constexpr int LUT_COMBOS = 84;  // multisets of <= 6 dice over 3 faces = C(9,3)
constexpr int LUT_IDX = 512;
constexpr int LUT_TAB = 924;    // multisets of <= 6 dice over 6 faces = C(12,6)
constexpr int LUT_VARIANTS = 3;
// discard table: [consider_score 2][consider_dice 2][favor_score 2][require_both 2] x
//                [all_singles 2][sf 3][bm 3][xs 8][yd 6]
constexpr int DISC_INNER = 2 * 3 * 3 * 8 * 6;  // entries per strategy class = 864
constexpr int LUT_DISC = 16 * DISC_INNER;      // 13,824
constexpr int LUT_BYTES = 3 * LUT_IDX + 4 * LUT_VARIANTS * LUT_TAB + LUT_DISC + 4 * LUT_IDX;  // 28,496
constexpr int LUT_OFF_COLB = 2 * LUT_IDX;
constexpr int LUT_OFF_TAB = 3 * LUT_IDX;
constexpr int LUT_OFF_DISC = LUT_OFF_TAB + 4 * LUT_VARIANTS * LUT_TAB;
constexpr int LUT_OFF_HIST3 = LUT_OFF_DISC + LUT_DISC;

struct ScoreLut {
    uint16_t rowA[LUT_IDX];
    uint8_t colB[LUT_IDX];
    uint32_t tab[LUT_VARIANTS * LUT_TAB];
    uint8_t disc[LUT_DISC];
    uint32_t hist3[LUT_IDX];  // three 3-bit face codes (6, 7 = no die) -> their packed face counts
};
static_assert(sizeof(ScoreLut) == LUT_BYTES, "lut layout");

// Strategy constants as the kernels keep them per seat (SeatImm, written at seeding):
//   st_d  score threshold for BOTH the keep/bank rule and the discard search, -2^30 when the
//         strategy does not consider score ("turn_score < st_d" is then never true; the discard
//         sub-table of such a strategy ignores the score axis)
//   dt_d  dice threshold, 127 when the strategy does not consider dice
//   kf    KF_* flags
//   dbase discard sub-table offset, tab_off score-table variant offset (elements)
constexpr uint32_t KF_AND_MODE = 1u;  // both considered and NOT require_both: stop unless both say go
constexpr uint32_t KF_AUTO_HOT = 2u;
constexpr uint32_t KF_RUN_UP = 4u;
constexpr int ST_NOT_CONSIDERED = -(1 << 30);
constexpr int DT_NOT_CONSIDERED = 127;

struct SeatConsts {
    int st_d, dt_d;
    uint32_t kf, dbase, tab_off;
};
inline
#ifdef __CUDACC__
__host__ __device__
#endif
SeatConsts seat_consts(int score_threshold, int dice_threshold, uint32_t flags) {
    const bool cs = flags & FB_SF_CONSIDER_SCORE, cd = flags & FB_SF_CONSIDER_DICE;
    const bool rb = flags & FB_SF_REQUIRE_BOTH, fav = flags & FB_SF_FAVOR_SCORE;
    const bool both = cs && cd && rb;
    SeatConsts c;
    c.st_d = cs ? score_threshold : ST_NOT_CONSIDERED;
    c.dt_d = cd ? dice_threshold : DT_NOT_CONSIDERED;
    c.kf = ((cs && cd && !rb) ? KF_AND_MODE : 0u) | ((flags & FB_SF_AUTO_HOT_DICE) ? KF_AUTO_HOT : 0u) |
           ((flags & FB_SF_RUN_UP_SCORE) ? KF_RUN_UP : 0u);
    c.dbase = (uint32_t)((((cs ? 2 : 0) + (cd ? 1 : 0)) * 2 + (fav ? 1 : 0)) * 2 + (both ? 1 : 0)) * DISC_INNER;
    const int variant = (flags & FB_SF_SMART_FIVE) ? ((flags & FB_SF_SMART_ONE) ? 2 : 1) : 0;
    c.tab_off = (uint32_t)variant * LUT_TAB;
    return c;
}

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
    // the 84 three-face combinations, numbered by ascending dice count (then lexicographically)
    int combo[LUT_COMBOS][3], csum[LUT_COMBOS];
    int n = 0;
    for (int s = 0; s <= 6; s++)
        for (int a = 0; a <= s; a++)
            for (int b = 0; a + b <= s; b++) {
                combo[n][0] = a; combo[n][1] = b; combo[n][2] = s - a - b;
                csum[n] = s;
                n++;
            }
    // rows: combination i of faces 1-3 (using csum[i] dice) owns the columns j with csum[j] <= 6 - csum[i]
    int row_len[7], row_off[LUT_COMBOS];
    for (int s = 0; s <= 6; s++) {
        row_len[s] = 0;
        for (int j = 0; j < LUT_COMBOS; j++) row_len[s] += csum[j] <= 6 - s;
    }
    int total = 0;
    for (int i = 0; i < LUT_COMBOS; i++) {
        row_off[i] = total;
        total += row_len[csum[i]];
    }
    // (total == LUT_TAB: 924)
    for (int i = 0; i < LUT_IDX; i++) {
        lut.rowA[i] = 0;
        lut.colB[i] = 0;
        uint32_t h = 0;
        for (int d = 0; d < 3; d++) {
            const int code = (i >> (3 * d)) & 7;
            if (code < 6) h += 1u << (3 * code);
        }
        lut.hist3[i] = h;
    }
    for (int i = 0; i < LUT_COMBOS; i++) {
        const int key = combo[i][0] | (combo[i][1] << 3) | (combo[i][2] << 6);
        lut.rowA[key] = (uint16_t)row_off[i];
        lut.colB[key] = (uint8_t)i;
    }
    for (int v = 0; v < LUT_VARIANTS; v++)
        for (int i = 0; i < LUT_COMBOS; i++)
            for (int j = 0; j < row_len[csum[i]]; j++) {
                int c[6] = {combo[i][0], combo[i][1], combo[i][2], combo[j][0], combo[j][1], combo[j][2]};
                const int nd = c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
                uint32_t e = 0;
                {
                    RollScore r = host_evaluate_counts(c);
                    e = (uint32_t)((r.score / 50) | (r.used << 7) | (r.sf << 10) | (r.so << 12));
                    // decide_smart_discards (scoring.py:369-467): candidates exist only with
                    // smart_five, unused dice left over and lone fives/ones to give back
                    const bool on = v >= 1 && r.used != nd;
                    const int sfi = on ? r.sf : 0, bm = (on && v == 2) ? r.so : 0;
                    const int excl = r.score == 50 * sfi + 100 * bm ? 1 : 0;  // all-discard scores 0
                    e |= (uint32_t)(((excl * 3 + sfi) * 3 + bm) * 48) << 16;
                }
                lut.tab[v * LUT_TAB + row_off[i] + j] = e;
            }
    // Smart-discard table.  A candidate "drop a lone fives and b lone ones" loses
    // u = a + 2b units of 50 points and frees D = a + b dice; whether it must bank depends
    // only on u < xs (score threshold still met) and D < yd (dice threshold still met),
    // and the preference key is (-u, D) or (D, -u).  So the whole search of
    // decide_smart_discards (scoring.py:303-467) is a function of nine small integers.
    for (int cs = 0; cs < 2; cs++)
    for (int cd = 0; cd < 2; cd++)
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
                                        const bool hit_s = cs && u < xs, hit_d = cd && D < yd;
                                        if (both ? (hit_s && hit_d) : (hit_s || hit_d)) continue;
                                        const int k1 = fav ? -u : D, k2 = fav ? D : -u;
                                        if (!have || k1 > best_k1 || (k1 == best_k1 && k2 > best_k2)) {
                                            have = true;
                                            best_k1 = k1;
                                            best_k2 = k2;
                                            pick = a | (b << 2);
                                        }
                                    }
                                const int idx = ((((cs * 2 + cd) * 2 + fav) * 2 + both) * DISC_INNER) +
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
__device__ __forceinline__ uint32_t lut_lookup(const ScoreLut* lut, uint32_t tab_off, uint32_t hist) {
    const uint32_t a = lut->rowA[hist & 511u];
    const uint32_t b = lut->colB[hist >> 9];
    return lut->tab[tab_off + a + b];
}

// decide_smart_discards (scoring.py:369-467) as ONE table lookup, branch free.
// The candidates that survive _select_candidate's filter are exactly "drop a of the sf
// lone fives and b of the so lone ones" (b = 0 unless smart_one); re-scoring one through
// the table gives score-50a-100b with used-a-b dice (no special 6-dice pattern can appear
// because used != n).  With X = ts + score - score_threshold and Y = dice_threshold -
// (n - used):  hit_score <=> consider_score and 50(a+2b) <= X,  hit_dice <=>
// consider_dice and a+b <= Y.  xs / yd below count how many values of a+2b / a+b hit; the
// (sf, so, all-singles) part of the index comes premultiplied from the score-table entry e
// of the strategy's variant, the consider/favor/both part from the seat's dbase.
// Returns d5 | d1 << 2.
__device__ __forceinline__ uint32_t smart_discards(const ScoreLut* lut, uint32_t dbase, uint32_t e, int n,
                                                   int ts, int st_d, int dt_d) {
    const int score = (int)(e & 127u) * 50, used = (int)((e >> 7) & 7u);
    int xs = min(max(ts + score - st_d + 50, 0), 399);
    xs = (xs * 1311) >> 16;  // floor(xs / 50) for 0 <= xs <= 399
    const int yd = min(max(dt_d - (n - used) + 1, 0), 5);
    return lut->disc[dbase + (e >> 16) + (uint32_t)(xs * 6 + yd)];
}

// Shared-memory twins of the two lookups above for play_kernel: `lut_s` is the 32-bit
// shared-space address of the ScoreLut copy, kept in one register for the whole kernel (taking the
// generic pointer instead makes ptxas rebuild the shared-window base from SR_CgaCtaId every roll).
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lut_lookup_s(uint32_t lut_s, uint32_t tab_off, uint32_t hist) {
    const uint32_t a = lds_u16(lut_s + 2u * (hist & 511u));
    const uint32_t b = lds_u8(lut_s + LUT_OFF_COLB + (hist >> 9));
    return lds_u32(lut_s + LUT_OFF_TAB + 4u * (tab_off + a + b));
}
__device__ __forceinline__ uint32_t smart_discards_s(uint32_t lut_s, uint32_t dbase, uint32_t e, int n, int ts,
                                                     int st_d, int dt_d) {
    const int score = (int)(e & 127u) * 50, used = (int)((e >> 7) & 7u);
    int xs = min(max(ts + score - st_d + 50, 0), 399);
    xs = (xs * 1311) >> 16;  // floor(xs / 50) for 0 <= xs <= 399
    const int yd = min(max(dt_d - (n - used) + 1, 0), 5);
    return lds_u8(lut_s + LUT_OFF_DISC + dbase + (e >> 16) + (uint32_t)(xs * 6 + yd));
}

// _decide_continue (strategies.py:125-162), branch free on the seat constants: a threshold
// that is not considered can never "want" to go on, so OR gives the single-threshold answer.
__device__ __forceinline__ bool decide_continue(int ts, int dice, int st_d, int dt_d, uint32_t kf) {
    const bool want_s = ts < st_d;
    const bool want_d = dice > dt_d;
    return (kf & KF_AND_MODE) ? (want_s && want_d) : (want_s || want_d);
}
#endif  // __CUDACC__

}  // namespace fb
