// Host check of the score lookup builder (farkle_ii_b200/csrc/scoring.cuh): every multiset of at
// most six dice must be reachable through rowA/colB and carry the fields host_evaluate_counts gives.
// Compiled and run by tests/test_host_logic.py::test_score_lookup_table_layout (no GPU needed).
#include <cstdio>
#include <set>

#include "../farkle_ii_b200/csrc/scoring.cuh"

int main() {
    static fb::ScoreLut lut;
    fb::host_build_lut(lut);
    int checked = 0, bad = 0;
    std::set<unsigned> slots;
    int c[6];
    for (c[0] = 0; c[0] <= 6; c[0]++)
    for (c[1] = 0; c[0] + c[1] <= 6; c[1]++)
    for (c[2] = 0; c[0] + c[1] + c[2] <= 6; c[2]++)
    for (c[3] = 0; c[0] + c[1] + c[2] + c[3] <= 6; c[3]++)
    for (c[4] = 0; c[0] + c[1] + c[2] + c[3] + c[4] <= 6; c[4]++)
    for (c[5] = 0; c[0] + c[1] + c[2] + c[3] + c[4] + c[5] <= 6; c[5]++) {
        unsigned hist = 0;
        for (int f = 0; f < 6; f++) hist |= (unsigned)c[f] << (3 * f);
        const unsigned slot = lut.rowA[hist & 511u] + lut.colB[hist >> 9];
        if (slot >= (unsigned)fb::LUT_TAB) { bad++; continue; }
        slots.insert(slot);
        const fb::RollScore r = fb::host_evaluate_counts(c);
        const int nd = c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
        for (int v = 0; v < fb::LUT_VARIANTS; v++) {
            const unsigned e = lut.tab[v * fb::LUT_TAB + slot];
            const bool on = v >= 1 && r.used != nd;
            const int sfi = on ? r.sf : 0, bm = (on && v == 2) ? r.so : 0;
            const int excl = r.score == 50 * sfi + 100 * bm ? 1 : 0;
            const unsigned want = (unsigned)((r.score / 50) | (r.used << 7) | (r.sf << 10) | (r.so << 12)) |
                                  ((unsigned)(((excl * 3 + sfi) * 3 + bm) * 48) << 16);
            bad += e != want;
        }
        checked++;
    }
    // hist3: three queued 3-bit face codes (6, 7 = no die) -> their packed face counts; two lookups
    // give the histogram of a roll of up to six dice (play.cuh, face queue)
    for (unsigned i = 0; i < 512; i++) {
        unsigned want = 0;
        for (int d = 0; d < 3; d++) {
            const unsigned code = (i >> (3 * d)) & 7u;
            if (code < 6u) want += 1u << (3 * code);
        }
        bad += lut.hist3[i] != want;
    }
    printf("%d %d %zu %d\n", checked, bad, slots.size(), (int)sizeof(fb::ScoreLut));
    return bad != 0;
}
