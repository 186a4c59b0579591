/* farkle_oracle.c — CPU restatement of the Farkle_II simulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in farkle_ii_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the reference
 * (/root/reference/src) in the build container and freezes its outputs
 * (SeedSequence words, PCG64DXSM states, dice, permutations, the 923-entry
 * score table, default_score sweeps, whole-game rows, tallies, H2H blocks);
 * tests/test_oracle_golden.py checks this file against those fixtures and
 * against the reference's own known-answer values.
 *
 * Third-party arithmetic on the path: NumPy (pinned `numpy>=1.26` in the
 * reference's pyproject.toml:22; 2.3.5 in the container) — SeedSequence,
 * PCG64DXSM, Generator.integers (Lemire, 32-bit buffered), Generator.permutation.
 * Their published algorithms are restated below; call sites in the reference:
 * src/farkle/utils/random.py:156,188,225, src/farkle/game/engine.py:101,
 * src/farkle/simulation/run_tournament.py:318.
 *
 * The discard search deliberately follows the reference's structure
 * (enumerate post-discard multisets, re-score each through the table, filter,
 * pick by key) instead of the closed form the CUDA kernel uses, so that the
 * two are independent derivations.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/farkle_b200.h"

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------- */
/* SeedSequence (numpy/random/bit_generator.pyx; used at utils/random.py:156)  */
/* ------------------------------------------------------------------------- */
#define SS_INIT_A 0x43b0d7e5u
#define SS_MULT_A 0x931e8875u
#define SS_INIT_B 0x8b51f9ddu
#define SS_MULT_B 0x58f38dedu
#define SS_MIX_L 0xca01f9ddu
#define SS_MIX_R 0x4973f715u
#define SS_XSHIFT 16

static inline uint32_t ss_hashmix(uint32_t v, uint32_t* hc) {
    v ^= *hc;
    *hc *= SS_MULT_A;
    v *= *hc;
    v ^= v >> SS_XSHIFT;
    return v;
}
static inline uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = SS_MIX_L * x - SS_MIX_R * y;
    r ^= r >> SS_XSHIFT;
    return r;
}
static void ss_pool(const uint32_t* entropy, int n, uint32_t pool[4]) {
    uint32_t hc = SS_INIT_A;
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < n ? entropy[i] : 0u, &hc);
    for (int s = 0; s < 4; s++)
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], &hc));
    for (int s = 4; s < n; s++)
        for (int d = 0; d < 4; d++) pool[d] = ss_mix(pool[d], ss_hashmix(entropy[s], &hc));
}
static void ss_generate(const uint32_t pool[4], int n_words, uint32_t* out) {
    uint32_t hc = SS_INIT_B;
    for (int i = 0; i < n_words; i++) {
        uint32_t v = pool[i & 3];
        v ^= hc;
        hc *= SS_MULT_B;
        v *= hc;
        v ^= v >> SS_XSHIFT;
        out[i] = v;
    }
}

void fo_seedseq_generate(const uint32_t* entropy, int n_entropy, int n_words, uint32_t* out) {
    uint32_t pool[4];
    ss_pool(entropy, n_entropy, pool);
    ss_generate(pool, n_words, out);
}

/* coordinate_entropy — src/farkle/utils/random.py:80-124 (18 uint32 words). */
typedef struct {
    uint64_t purpose, root_seed, k, shuffle_index, pair_id, order, game_index, seat_index,
        replicate_index;
} fo_coord_t;

static void coord_entropy(const fo_coord_t* c, uint32_t e[18]) {
    const uint64_t v[8] = {c->root_seed, c->k,          c->shuffle_index, c->pair_id,
                           c->order,     c->game_index, c->seat_index,    c->replicate_index};
    e[0] = 2u; /* RNG_SCHEME_VERSION, utils/random.py:13 */
    e[1] = (uint32_t)c->purpose;
    for (int i = 0; i < 8; i++) {
        e[2 + 2 * i] = (uint32_t)(v[i] & 0xffffffffu);
        e[3 + 2 * i] = (uint32_t)(v[i] >> 32);
    }
}

void fo_coordinate_entropy(const uint64_t coords[9], uint32_t out[18]) {
    fo_coord_t c = {coords[0], coords[1], coords[2], coords[3], coords[4],
                    coords[5], coords[6], coords[7], coords[8]};
    coord_entropy(&c, out);
}

/* coordinate_seed — src/farkle/utils/random.py:191-225. */
static uint64_t coord_seed(const fo_coord_t* c, int as_u32) {
    uint32_t e[18], w[2];
    coord_entropy(c, e);
    fo_seedseq_generate(e, 18, 2, w);
    return as_u32 ? (uint64_t)w[0] : ((uint64_t)w[0] | ((uint64_t)w[1] << 32));
}
uint64_t fo_coordinate_seed(const uint64_t coords[9], int as_u32) {
    fo_coord_t c = {coords[0], coords[1], coords[2], coords[3], coords[4],
                    coords[5], coords[6], coords[7], coords[8]};
    return coord_seed(&c, as_u32);
}

/* ------------------------------------------------------------------------- */
/* PCG64DXSM (numpy/random/src/pcg64) — coordinate_rng, utils/random.py:159-188 */
/* ------------------------------------------------------------------------- */
#define PCG_CHEAP_MULT 0xda942042e4dd58b5ULL
static const u128 PCG_DEFAULT_MULT =
    ((u128)2549297995355413924ULL << 64) | (u128)4865540595714422341ULL;

typedef struct {
    u128 state, inc;
    int has32;
    uint32_t saved;
    uint64_t words, dice; /* work counters: fresh 64-bit outputs / dice drawn */
    uint64_t rejects;     /* halves the Lemire test threw away (fo_scan_rejected_halves) */
} pcg_t;

static void pcg_seed_words(pcg_t* g, const uint64_t w[4]) {
    u128 initstate = ((u128)w[0] << 64) | w[1];
    u128 initseq = ((u128)w[2] << 64) | w[3];
    g->inc = (initseq << 1) | 1u;
    g->state = 0;
    g->state = g->state * PCG_DEFAULT_MULT + g->inc;
    g->state += initstate;
    g->state = g->state * PCG_DEFAULT_MULT + g->inc;
    g->has32 = 0;
    g->saved = 0;
    g->words = 0;
    g->dice = 0;
}
static void pcg_seed_coord(pcg_t* g, const fo_coord_t* c) {
    uint32_t e[18], s[8];
    uint64_t w[4];
    coord_entropy(c, e);
    fo_seedseq_generate(e, 18, 8, s);
    for (int j = 0; j < 4; j++) w[j] = (uint64_t)s[2 * j] | ((uint64_t)s[2 * j + 1] << 32);
    pcg_seed_words(g, w);
}
static inline uint64_t pcg_next64(pcg_t* g) {
    uint64_t hi = (uint64_t)(g->state >> 64);
    uint64_t lo = (uint64_t)g->state | 1u;
    g->state = g->state * (u128)PCG_CHEAP_MULT + g->inc;
    hi ^= hi >> 32;
    hi *= PCG_CHEAP_MULT;
    hi ^= hi >> 48;
    hi *= lo;
    return hi;
}
static inline uint32_t pcg_next32(pcg_t* g) {
    if (g->has32) {
        g->has32 = 0;
        return g->saved;
    }
    uint64_t n = pcg_next64(g);
    g->has32 = 1;
    g->saved = (uint32_t)(n >> 32);
    return (uint32_t)n;
}

/* Generator.integers(1, 7) — Lemire 32-bit path; game/engine.py:101. */
/* Test knob shared with the CUDA library (FB_TEST_LEMIRE_THR=t, 4 <= t <= 2^31): a half whose low
 * product word is below t is re-drawn.  NumPy's threshold is 4 (four 32-bit values in 2^32); the
 * knob makes rejected halves frequent so that the tests can exercise the kernel's handling of
 * them.  Refreshed at every exported play call (refresh_roll_limit). */
static uint32_t g_lemire_thr = 4u;
static void refresh_roll_limit(void);
static inline int pcg_die(pcg_t* g) {
    const uint32_t rng_excl = 6;
    if (!g->has32) g->words++;
    uint64_t m = (uint64_t)pcg_next32(g) * rng_excl;
    uint32_t left = (uint32_t)m;
    const uint32_t thr = g_lemire_thr; /* (0xffffffff - 5) % 6 == 4 without the knob */
    if (left < (thr > rng_excl ? thr : rng_excl)) {
        while (left < thr) {
            g->rejects++;
            if (!g->has32) g->words++;
            m = (uint64_t)pcg_next32(g) * rng_excl;
            left = (uint32_t)m;
        }
    }
    g->dice++;
    return 1 + (int)(m >> 32);
}

void fo_seed_stream(const uint64_t coords[9], uint64_t out[4]) {
    fo_coord_t c = {coords[0], coords[1], coords[2], coords[3], coords[4],
                    coords[5], coords[6], coords[7], coords[8]};
    pcg_t g;
    pcg_seed_coord(&g, &c);
    out[0] = (uint64_t)(g.state >> 64);
    out[1] = (uint64_t)g.state;
    out[2] = (uint64_t)(g.inc >> 64);
    out[3] = (uint64_t)g.inc;
}

/* Rolls from an explicit generator state (has32/saved included so crafted
 * states can exercise the Lemire rejection loop). */
void fo_roll_dice_state(const uint64_t state_inc[4], int has32, uint32_t saved,
                        const int32_t* n_dice, int n_rolls, uint8_t* faces_out) {
    pcg_t g;
    g.state = ((u128)state_inc[0] << 64) | state_inc[1];
    g.inc = ((u128)state_inc[2] << 64) | state_inc[3];
    g.has32 = has32;
    g.saved = saved;
    g.words = g.dice = g.rejects = 0;
    refresh_roll_limit();
    for (int r = 0; r < n_rolls; r++)
        for (int i = 0; i < 6; i++) faces_out[r * 6 + i] = i < n_dice[r] ? (uint8_t)pcg_die(&g) : 0;
}

/* Generator.permutation(N) — run_tournament.py:312-318. */
static void pcg_permutation(pcg_t* g, int n, int32_t* a) {
    for (int i = 0; i < n; i++) a[i] = i;
    for (int i = n - 1; i >= 1; i--) {
        uint32_t mask = (uint32_t)i;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        do {
            v = pcg_next32(g) & mask;
        } while (v > (uint32_t)i);
        int32_t t = a[i];
        a[i] = a[v];
        a[v] = t;
    }
}
void fo_permutation(uint64_t root_seed, uint64_t k, uint64_t shuffle_index, int n, int32_t* out) {
    fo_coord_t c = {FB_PURPOSE_SHUFFLE_PERMUTATION, root_seed, k, shuffle_index, 0, 0, 0, 0, 0};
    pcg_t g;
    pcg_seed_coord(&g, &c);
    pcg_permutation(&g, n, out);
}

/* ------------------------------------------------------------------------- */
/* Scoring — game/scoring_lookup.py:123-172                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int score, used, sf, so;
} score_t;

static score_t evaluate_counts(const int c_in[6]) {
    int c[6];
    memcpy(c, c_in, sizeof c);
    score_t r = {0, 0, 0, 0};
    int all_one = 1, pairs = 0, trips = 0, has4 = 0, has2 = 0;
    for (int f = 0; f < 6; f++) {
        if (c[f] != 1) all_one = 0;
        if (c[f] == 2) { pairs++; has2 = 1; }
        if (c[f] == 3) trips++;
        if (c[f] == 4) has4 = 1;
    }
    if (all_one) { r.score = 1500; r.used = 6; return r; }          /* _straight       :27-38 */
    if (pairs == 3) { r.score = 1500; r.used = 6; return r; }       /* _three_pairs    :41-53 */
    if (trips == 2) { r.score = 2500; r.used = 6; return r; }       /* _two_triplets   :56-68 */
    if (has4 && has2) { r.score = 1500; r.used = 6; return r; }     /* _four_kind_plus_pair :71-82 */
    for (int f = 0; f < 6; f++) {                                    /* _apply_sets     :85-115 */
        int n = c[f];
        if (n >= 3) {
            int pts = n == 3 ? (f == 0 ? 300 : (f + 1) * 100) : n == 4 ? 1000 : n == 5 ? 2000 : 3000;
            r.score += pts;
            r.used += n;
            c[f] = 0;
        }
    }
    r.so = c[0];
    r.sf = c[4];
    r.score += r.so * 100 + r.sf * 50;
    r.used += r.so + r.sf;
    return r;
}

void fo_evaluate_counts(const int32_t counts[6], int32_t out[4]) {
    int c[6];
    for (int i = 0; i < 6; i++) c[i] = counts[i];
    score_t r = evaluate_counts(c);
    out[0] = r.score; out[1] = r.used; out[2] = r.sf; out[3] = r.so;
}

typedef struct {
    int score_threshold, dice_threshold;
    int smart_five, smart_one, consider_score, consider_dice, require_both, auto_hot_dice,
        run_up_score, favor_score;
} strat_t;

static strat_t strat_unpack(const fb_strategy_t* s) {
    strat_t r;
    r.score_threshold = s->score_threshold;
    r.dice_threshold = s->dice_threshold;
    r.smart_five = !!(s->flags & FB_SF_SMART_FIVE);
    r.smart_one = !!(s->flags & FB_SF_SMART_ONE);
    r.consider_score = !!(s->flags & FB_SF_CONSIDER_SCORE);
    r.consider_dice = !!(s->flags & FB_SF_CONSIDER_DICE);
    r.require_both = !!(s->flags & FB_SF_REQUIRE_BOTH);
    r.auto_hot_dice = !!(s->flags & FB_SF_AUTO_HOT_DICE);
    r.run_up_score = !!(s->flags & FB_SF_RUN_UP_SCORE);
    r.favor_score = !!(s->flags & FB_SF_FAVOR_SCORE);
    return r;
}

/* _must_bank — game/scoring.py:283-300 */
static int must_bank(int score_after, int dice_left_after, const strat_t* s) {
    int hit_score = s->consider_score ? (score_after >= s->score_threshold) : 0;
    int hit_dice = s->consider_dice ? (dice_left_after <= s->dice_threshold) : 0;
    return (s->consider_score && s->consider_dice && s->require_both) ? (hit_score && hit_dice)
                                                                      : (hit_score || hit_dice);
}

/* decide_smart_discards — game/scoring.py:369-467 via generate_sequences
 * (:196-236), score_lister (:242-275) and _select_candidate (:303-366). */
static void decide_discards(const int counts[6], const score_t* raw, int n, int turn_score_pre,
                            const strat_t* s, int* d5, int* d1) {
    *d5 = 0;
    *d1 = 0;
    if (!s->smart_five || raw->used == n || (raw->sf == 0 && raw->so == 0)) return;
    int max_fives = counts[4];
    int max_ones = s->smart_one ? counts[0] : 0;
    int have_best = 0, best_a = 0, best_b = 0, best_sf = raw->sf, best_so = raw->so;
    for (int drop5 = 0; drop5 <= max_fives; drop5++) {
        for (int drop1 = 0; drop1 <= max_ones; drop1++) {
            int nc[6];
            memcpy(nc, counts, sizeof nc);
            nc[4] -= drop5;
            nc[0] -= drop1;
            score_t cand = evaluate_counts(nc);
            if (cand.score == 0) continue;                 /* score_lister :261-262 */
            if (drop5 > raw->sf || drop1 > raw->so) continue; /* :330-333 */
            int score_after = turn_score_pre + cand.score;
            int dice_left_after = n - cand.used;
            if (must_bank(score_after, dice_left_after, s)) continue;
            int ka = s->favor_score ? score_after : dice_left_after;
            int kb = s->favor_score ? dice_left_after : score_after;
            if (!have_best || ka > best_a || (ka == best_a && kb > best_b)) {
                have_best = 1;
                best_a = ka;
                best_b = kb;
                best_sf = cand.sf;
                best_so = cand.so;
            }
        }
    }
    if (!have_best) return;
    *d5 = raw->sf - best_sf;
    *d1 = raw->so - best_so;
}

/* default_score(..., return_discards=True) — game/scoring.py:618-693 */
static void default_score(const uint8_t* faces, int n, int turn_score_pre, const strat_t* s,
                          int out[5]) {
    int counts[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) counts[faces[i] - 1]++;
    score_t raw = evaluate_counts(counts);
    int d5, d1;
    decide_discards(counts, &raw, n, turn_score_pre, s, &d5, &d1);
    int final_score = raw.score - 50 * d5 - 100 * d1; /* apply_discards :548-578 */
    int final_used = raw.used - d5 - d1;
    out[0] = final_score;
    out[1] = final_used;
    out[2] = n - final_used;
    out[3] = d5;
    out[4] = d1;
}

void fo_default_score(const uint8_t faces[6], int32_t turn_score_pre, const fb_strategy_t* strat,
                      int32_t out[5]) {
    int n = 0;
    uint8_t f[6];
    for (int i = 0; i < 6; i++)
        if (faces[i]) f[n++] = faces[i];
    strat_t s = strat_unpack(strat);
    int o[5];
    default_score(f, n, turn_score_pre, &s, o);
    for (int i = 0; i < 5; i++) out[i] = o[i];
}

/* ------------------------------------------------------------------------- */
/* Player turn and game — game/engine.py:208-273, 436-550                     */
/* ------------------------------------------------------------------------- */
#define ROLL_LIMIT 1000 /* game/engine.py:36 */
/* Test knob shared with the CUDA library (FB_TEST_ROLL_LIMIT=n, 1 <= n <= 1000): lets the tests reach
 * the RuntimeError path of engine.py:242-243, which real play practically never does. */
static int g_roll_limit = ROLL_LIMIT; /* refreshed at every exported play call */
static void refresh_roll_limit(void) {
    const char* e = getenv("FB_TEST_ROLL_LIMIT");
    const int v = e ? atoi(e) : ROLL_LIMIT;
    g_roll_limit = (v >= 1 && v <= ROLL_LIMIT) ? v : ROLL_LIMIT;
    const char* t = getenv("FB_TEST_LEMIRE_THR");
    const unsigned long thr = t ? strtoul(t, NULL, 0) : 4ul;
    g_lemire_thr = (thr >= 4ul && thr <= 0x80000000ul) ? (uint32_t)thr : 4u;
}

typedef struct {
    pcg_t rng;
    strat_t strat;
    int strategy_id;
    int score, has_scored;
    int n_turns, n_farkles, n_rolls, highest_turn, sf_uses, n_sf_dice, so_uses, n_so_dice,
        n_hot_dice;
} player_t;

/* _decide_continue — simulation/strategies.py:125-162 */
static int decide_continue(int turn_score, int dice_left, const strat_t* s) {
    int want_s = s->consider_score && turn_score < s->score_threshold;
    int want_d = s->consider_dice && dice_left > s->dice_threshold;
    if (s->consider_score && s->consider_dice)
        return s->require_both ? (want_s || want_d) : (want_s && want_d);
    if (s->consider_score) return want_s;
    if (s->consider_dice) return want_d;
    return 0;
}

/* ThresholdStrategy.decide — simulation/strategies.py:212-275 */
static int strategy_decide(const player_t* p, int turn_score, int dice_left, int final_round,
                           int score_to_beat, int running_total) {
    if (!p->has_scored && turn_score < 500) return 1;
    if (final_round) {
        if (running_total <= score_to_beat) return 1;
        if (!p->strat.run_up_score) return 0;
    }
    return decide_continue(turn_score, dice_left, &p->strat);
}

/* FarklePlayer.take_turn — game/engine.py:208-273.  Returns 1 on ROLL_LIMIT. */
static int take_turn(player_t* p, int final_round, int score_to_beat, uint64_t* rolls_ctr) {
    p->n_turns++;
    int dice = 6, turn_score = 0, rolls_this_turn = 0;
    const int limit = g_roll_limit;
    while (dice > 0) {
        if (rolls_this_turn >= limit) return 1;
        uint8_t roll[6];
        p->n_rolls++; /* _roll :85-101 */
        (*rolls_ctr)++;
        for (int i = 0; i < dice; i++) roll[i] = (uint8_t)pcg_die(&p->rng);
        rolls_this_turn++;
        int o[5]; /* _score_roll :103-147 */
        default_score(roll, dice, turn_score, &p->strat, o);
        int pts = o[0], used = o[1], reroll = o[2], d5 = o[3], d1 = o[4];
        if (pts == 0) {
            p->n_farkles++;
            turn_score = 0;
            break;
        }
        if (d5 > 0) { p->sf_uses++; p->n_sf_dice += d5; }
        if (d1 > 0) { p->so_uses++; p->n_so_dice += d1; }
        dice = (used == dice && reroll == 0) ? 6 : reroll;
        turn_score += pts;
        if (p->strat.auto_hot_dice && dice == 6) { /* _apply_hot_dice :149-154 */
            p->n_hot_dice++;
            continue;
        }
        /* _should_continue :156-205 */
        int running_total = p->score + turn_score;
        if (final_round && running_total > score_to_beat && !p->strat.run_up_score) break;
        int keep = strategy_decide(p, turn_score, dice, final_round, score_to_beat, running_total);
        if (final_round && running_total <= score_to_beat) keep = 1;
        if (!keep) break;
    }
    if (!p->has_scored && turn_score >= 500) p->has_scored = 1;
    if (p->has_scored) {
        p->score += turn_score;
        if (turn_score > p->highest_turn) p->highest_turn = turn_score;
    }
    return 0;
}

typedef struct {
    uint64_t rolls, dice, words, turns, rejects;
} work_t;

/* FarkleGame.play + _run_final_round — game/engine.py:436-550, flattened into
 * the compact row of include/farkle_b200.h. */
static void play_game(const fo_coord_t* seat_coord_base, int k, const fb_strategy_t* strats,
                      const int32_t* strat_ids, int target_score, int max_rounds,
                      uint64_t game_seed, uint32_t ordinal, uint8_t* row_out, work_t* work) {
    player_t pl[FB_MAX_PLAYERS];
    memset(pl, 0, sizeof pl);
    for (int s = 0; s < k; s++) {
        fo_coord_t c = *seat_coord_base;
        c.seat_index = (uint64_t)s;
        pcg_seed_coord(&pl[s].rng, &c);
        pl[s].strat = strat_unpack(&strats[s]);
        pl[s].strategy_id = strat_ids ? strat_ids[s] : s;
    }
    uint64_t rolls = 0;
    int final_round = 0, score_to_beat = target_score, rounds = 0, err = 0;
    while (rounds < max_rounds && !err) {
        rounds++;
        for (int i = 0; i < k && !err; i++) {
            err |= take_turn(&pl[i], final_round, score_to_beat, &rolls);
            if (err) break;
            if (!final_round && pl[i].score >= target_score) {
                final_round = 1;
                score_to_beat = pl[i].score;
                for (int j = 0; j < k && !err; j++) {
                    if (j == i) continue;
                    err |= take_turn(&pl[j], 1, score_to_beat, &rolls);
                    if (pl[j].score > score_to_beat) score_to_beat = pl[j].score;
                }
                break;
            }
        }
        if (final_round) break;
    }
    int safety = (!final_round) && rounds >= max_rounds;
    int winner = 0xFF;
    if (!safety) {
        winner = 0;
        for (int s = 1; s < k; s++)
            if (pl[s].score > pl[winner].score) winner = s; /* stable: ties -> lower seat */
    }
    fb_row_header_t* h = (fb_row_header_t*)row_out;
    fb_row_seat_t* seats = (fb_row_seat_t*)(row_out + sizeof(fb_row_header_t));
    uint8_t flags = safety ? FB_ROW_SAFETY_LIMIT : 0;
    if (err) flags |= FB_ROW_ROLL_LIMIT;
    uint64_t turns = 0;
    for (int s = 0; s < k; s++) {
        const player_t* p = &pl[s];
        if (p->n_rolls > 32767 || p->highest_turn > 32767 || p->n_sf_dice > 32767 ||
            p->n_so_dice > 32767 || p->n_turns > 32767 || p->n_hot_dice > 32767)
            flags |= FB_ROW_I16_OVERFLOW;
        seats[s].score = p->score;
        seats[s].strategy = p->strategy_id;
        seats[s].highest_turn = p->highest_turn;
        seats[s].farkles = (uint16_t)p->n_farkles;
        seats[s].rolls = (uint16_t)p->n_rolls;
        seats[s].n_turns = (uint16_t)p->n_turns;
        seats[s].hot_dice = (uint16_t)p->n_hot_dice;
        seats[s].smart_five_uses = (uint16_t)p->sf_uses;
        seats[s].n_smart_five_dice = (uint16_t)p->n_sf_dice;
        seats[s].smart_one_uses = (uint16_t)p->so_uses;
        seats[s].n_smart_one_dice = (uint16_t)p->n_so_dice;
        turns += (uint64_t)p->n_turns;
    }
    h->game_seed = game_seed;
    h->game_ordinal = ordinal;
    h->n_rounds = (uint16_t)rounds;
    h->winner_seat = (uint8_t)winner;
    h->flags = flags;
    if (work) {
        work->rolls += rolls;
        for (int s = 0; s < k; s++) {
            work->dice += pl[s].rng.dice;
            work->words += pl[s].rng.words;
            work->rejects += pl[s].rng.rejects;
        }
        work->turns += turns;
    }
}

static size_t row_stride(int k) {
    size_t b = sizeof(fb_row_header_t) + (size_t)k * sizeof(fb_row_seat_t);
    return (b + 15u) & ~(size_t)15u;
}
size_t fo_row_stride(int k) { return row_stride(k); }

/* Tally one finished row — OutcomeCounter.record_row + the winner metric sums
 * of _play_one_shuffle, run_tournament.py:177-195,375-391. */
static void tally_row(const uint8_t* row, int k, int64_t* tallies /*[ids][26]*/, int64_t* totals) {
    const fb_row_header_t* h = (const fb_row_header_t*)row;
    const fb_row_seat_t* seats = (const fb_row_seat_t*)(row + sizeof(fb_row_header_t));
    int safety = h->flags & FB_ROW_SAFETY_LIMIT;
    if (tallies) {
        for (int s = 0; s < k; s++) {
            int64_t* t = tallies + (size_t)seats[s].strategy * FB_TALLY_WIDTH;
            t[1] += 1;
            t[safety ? 3 : 2] += 1;
        }
        if (!safety) {
            const fb_row_seat_t* w = &seats[h->winner_seat];
            int64_t* t = tallies + (size_t)w->strategy * FB_TALLY_WIDTH;
            const int64_t m[FB_N_METRICS] = {w->score,           h->n_rounds,
                                             w->farkles,         w->rolls,
                                             w->highest_turn,    w->smart_five_uses,
                                             w->n_smart_five_dice, w->smart_one_uses,
                                             w->n_smart_one_dice, w->hot_dice,
                                             0 /* winner_hit_max_rounds is False when completed */};
            t[0] += 1;
            for (int i = 0; i < FB_N_METRICS; i++) {
                t[4 + i] += m[i];
                t[4 + FB_N_METRICS + i] += m[i] * m[i];
            }
        }
    }
    if (totals) {
        totals[0] += 1;
        totals[safety ? 2 : 1] += 1;
        if (h->flags & (FB_ROW_ROLL_LIMIT | FB_ROW_I16_OVERFLOW)) totals[7] += 1;
        if (!safety) totals[8 + h->winner_seat] += 1;
    }
}

/* ------------------------------------------------------------------------- */
/* Tournament — run_tournament.py:301-393 (_play_one_shuffle) per shuffle      */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t root_seed;
    int k;
    uint64_t shuffle0;
    int n_shuffles;
    const fb_strategy_t* strategies;
    const int32_t* strategy_ids;
    int n_strategies, n_tally_ids;
    int target_score, max_rounds;
    const uint64_t* ov_shuffle;
    const uint32_t* ov_game;
    const int32_t* ov_max_rounds;
    int n_overrides;
    int shuffles_per_slot;
    int want_game_seeds;
    uint8_t* rows;
    /* per-thread outputs */
    int64_t* tallies;
    int64_t totals[FB_TOTALS_WIDTH];
    int thread, n_threads;
} tour_job_t;

static void* tour_worker(void* arg) {
    tour_job_t* j = (tour_job_t*)arg;
    const int k = j->k, n = j->n_strategies, gps = n / k;
    const size_t stride = row_stride(k);
    int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t* rowbuf = (uint8_t*)calloc(1, stride);
    work_t work = {0, 0, 0, 0, 0};
    for (int si = j->thread; si < j->n_shuffles; si += j->n_threads) {
        uint64_t shuffle = j->shuffle0 + (uint64_t)si;
        fo_permutation(j->root_seed, (uint64_t)k, shuffle, n, perm);
        int slot = j->shuffles_per_slot > 0 ? si / j->shuffles_per_slot : 0;
        int64_t* tl =
            j->tallies ? j->tallies + (size_t)slot * (size_t)j->n_tally_ids * FB_TALLY_WIDTH : NULL;
        for (int g = 0; g < gps; g++) {
            fb_strategy_t st[FB_MAX_PLAYERS];
            int32_t ids[FB_MAX_PLAYERS];
            for (int s = 0; s < k; s++) {
                int idx = perm[g * k + s];
                st[s] = j->strategies[idx];
                ids[s] = j->strategy_ids ? j->strategy_ids[idx] : idx;
            }
            int max_rounds = j->max_rounds;
            for (int o = 0; o < j->n_overrides; o++)
                if (j->ov_shuffle[o] == shuffle && j->ov_game[o] == (uint32_t)g) {
                    max_rounds = j->ov_max_rounds[o];
                    break;
                }
            fo_coord_t c = {FB_PURPOSE_TOURNAMENT_PLAYER, j->root_seed, (uint64_t)k, shuffle, 0, 0,
                            (uint64_t)g, 0, 0};
            uint64_t gseed = 0;
            if (j->want_game_seeds) {
                fo_coord_t gc = c;
                gc.purpose = FB_PURPOSE_TOURNAMENT_GAME;
                gseed = coord_seed(&gc, 1);
            }
            uint32_t ordinal = (uint32_t)((uint64_t)si * (uint64_t)gps + (uint64_t)g);
            uint8_t* row = j->rows ? j->rows + (size_t)ordinal * stride : rowbuf;
            memset(row, 0, stride);
            play_game(&c, k, st, ids, j->target_score, max_rounds, gseed, ordinal, row, &work);
            tally_row(row, k, tl, j->totals);
        }
    }
    j->totals[3] += (int64_t)work.rolls;
    j->totals[4] += (int64_t)work.dice;
    j->totals[5] += (int64_t)work.words;
    j->totals[6] += (int64_t)work.turns;
    free(perm);
    free(rowbuf);
    return NULL;
}

/* Same contract as fb_play_tournament, host buffers, n_threads workers over
 * shuffles (the reference parallelises over chunks of shuffles:
 * run_tournament.py:1576-1586).  tallies/totals are ACCUMULATED into. */
int fo_play_tournament(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                       const fb_strategy_t* strategies, const int32_t* strategy_ids,
                       int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                       const uint64_t* ov_shuffle, const uint32_t* ov_game,
                       const int32_t* ov_max_rounds, int n_overrides, int shuffles_per_slot,
                       int64_t* tallies, int64_t* totals, void* rows, int want_game_seeds,
                       int n_threads) {
    refresh_roll_limit();
    if (k < 1 || k > FB_MAX_PLAYERS || n_strategies % k != 0 || n_shuffles < 0) return -2;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_shuffles && n_shuffles > 0) n_threads = n_shuffles;
    int n_slots = shuffles_per_slot > 0 ? (n_shuffles + shuffles_per_slot - 1) / shuffles_per_slot : 1;
    size_t tally_len = (size_t)n_slots * (size_t)n_tally_ids * FB_TALLY_WIDTH;
    tour_job_t* jobs = (tour_job_t*)calloc((size_t)n_threads, sizeof(tour_job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        tour_job_t* j = &jobs[t];
        j->root_seed = root_seed; j->k = k; j->shuffle0 = shuffle0; j->n_shuffles = n_shuffles;
        j->strategies = strategies; j->strategy_ids = strategy_ids;
        j->n_strategies = n_strategies; j->n_tally_ids = n_tally_ids;
        j->target_score = target_score; j->max_rounds = max_rounds;
        j->ov_shuffle = ov_shuffle; j->ov_game = ov_game; j->ov_max_rounds = ov_max_rounds;
        j->n_overrides = n_overrides; j->shuffles_per_slot = shuffles_per_slot;
        j->want_game_seeds = want_game_seeds; j->rows = (uint8_t*)rows;
        j->thread = t; j->n_threads = n_threads;
        j->tallies = tallies ? (t == 0 ? tallies : (int64_t*)calloc(tally_len, sizeof(int64_t))) : NULL;
    }
    /* thread 0's job runs on the calling thread after the others start */
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, tour_worker, &jobs[t]);
    tour_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < n_threads; t++) {
        if (totals)
            for (int i = 0; i < FB_TOTALS_WIDTH; i++) totals[i] += jobs[t].totals[i];
        if (t > 0 && tallies) {
            for (size_t i = 0; i < tally_len; i++) tallies[i] += jobs[t].tallies[i];
            free(jobs[t].tallies);
        }
    }
    free(jobs);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Explicit-coordinate games — simulation.py:576-655 (_play_game)             */
/* ------------------------------------------------------------------------- */
int fo_play_games(const uint64_t* coords /*[n][7]*/, uint64_t n_games, int k,
                  const fb_strategy_t* seat_strategies, const int32_t* seat_strategy_ids,
                  const int32_t* target_score_v, int32_t target_score, const int32_t* max_rounds_v,
                  int32_t max_rounds, void* rows, int64_t* totals) {
    refresh_roll_limit();
    if (k < 1 || k > FB_MAX_PLAYERS) return -2;
    const size_t stride = row_stride(k);
    uint8_t* rowbuf = (uint8_t*)calloc(1, stride);
    work_t work = {0, 0, 0, 0, 0};
    for (uint64_t i = 0; i < n_games; i++) {
        const uint64_t* cc = coords + i * 7;
        fo_coord_t c = {cc[0], cc[1], cc[2], cc[3], cc[4], cc[5], cc[6], 0, 0};
        uint8_t* row = rows ? (uint8_t*)rows + i * stride : rowbuf;
        memset(row, 0, stride);
        play_game(&c, k, seat_strategies + i * (uint64_t)k,
                  seat_strategy_ids ? seat_strategy_ids + i * (uint64_t)k : NULL,
                  target_score_v ? target_score_v[i] : target_score,
                  max_rounds_v ? max_rounds_v[i] : max_rounds, 0, (uint32_t)i, row, &work);
        tally_row(row, k, NULL, totals);
    }
    if (totals) {
        totals[3] += (int64_t)work.rolls;
        totals[4] += (int64_t)work.dice;
        totals[5] += (int64_t)work.words;
        totals[6] += (int64_t)work.turns;
    }
    free(rowbuf);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* H2H block — analysis/h2h_schedule.py:1149-1243                              */
/* progress = {attempted, completed, safety, wins_seat1, wins_seat2}, updated  */
/* in place exactly like the reference loop (sequential, early stop).          */
/* ------------------------------------------------------------------------- */
int fo_play_h2h_block(uint64_t root_seed, uint64_t pair_id, int order, const fb_strategy_t* seat1,
                      const fb_strategy_t* seat2, int32_t n_completed_required,
                      int32_t max_attempts, int32_t chunk_games, int32_t target_score,
                      int32_t max_rounds, int32_t progress[5], uint8_t* outcome_out) {
    refresh_roll_limit();
    int attempted = progress[0], completed = progress[1], safety = progress[2];
    int w1 = progress[3], w2 = progress[4];
    int stop = attempted + chunk_games < max_attempts ? attempted + chunk_games : max_attempts;
    fb_strategy_t st[2] = {*seat1, *seat2};
    uint8_t row[128];
    int start = attempted;
    for (int a = start; a < stop; a++) {
        if (completed >= n_completed_required) break;
        fo_coord_t c = {FB_PURPOSE_H2H_PLAYER, root_seed, 2, 0, pair_id, (uint64_t)order,
                        (uint64_t)a, 0, 0};
        memset(row, 0, sizeof row);
        play_game(&c, 2, st, NULL, target_score, max_rounds, 0, (uint32_t)a, row, NULL);
        const fb_row_header_t* h = (const fb_row_header_t*)row;
        attempted++;
        uint8_t oc;
        if (h->flags & FB_ROW_SAFETY_LIMIT) { safety++; oc = 0; }
        else { completed++; if (h->winner_seat == 0) { w1++; oc = 1; } else { w2++; oc = 2; } }
        if (outcome_out) outcome_out[a - start] = oc;
    }
    progress[0] = attempted; progress[1] = completed; progress[2] = safety;
    progress[3] = w1; progress[4] = w2;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Test-fixture helper: which games of a tournament cell meet a REJECTED half   */
/* (Lemire leftover < 4: four 32-bit values in 2^32, about seven games of an    */
/* 11 M-game cell)?  tests/golden/make_golden_rejects.py uses it to pick the     */
/* games that exercise the rejection paths of the CUDA kernel.  out[i] =         */
/* {shuffle_index, game_index, rejected halves}; returns the number found        */
/* (at most cap are stored).  Single-threaded per call: callers split the        */
/* shuffle range.                                                                 */
/* ------------------------------------------------------------------------- */
int64_t fo_scan_rejected_halves(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                                const fb_strategy_t* strategies, int n_strategies,
                                int32_t target_score, int32_t max_rounds, uint64_t* out, int64_t cap) {
    refresh_roll_limit();
    if (k < 1 || k > FB_MAX_PLAYERS || n_strategies % k != 0 || n_shuffles < 0) return -2;
    const int gps = n_strategies / k;
    int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_strategies);
    uint8_t row[16 + 32 * FB_MAX_PLAYERS];
    int64_t found = 0;
    for (int si = 0; si < n_shuffles; si++) {
        const uint64_t shuffle = shuffle0 + (uint64_t)si;
        fo_permutation(root_seed, (uint64_t)k, shuffle, n_strategies, perm);
        for (int g = 0; g < gps; g++) {
            fb_strategy_t st[FB_MAX_PLAYERS];
            for (int s = 0; s < k; s++) st[s] = strategies[perm[g * k + s]];
            fo_coord_t c = {FB_PURPOSE_TOURNAMENT_PLAYER, root_seed, (uint64_t)k, shuffle, 0, 0,
                            (uint64_t)g, 0, 0};
            work_t work = {0, 0, 0, 0, 0};
            memset(row, 0, sizeof row);
            play_game(&c, k, st, NULL, target_score, max_rounds, 0, 0, row, &work);
            if (work.rejects) {
                if (found < cap) {
                    out[3 * found] = shuffle;
                    out[3 * found + 1] = (uint64_t)g;
                    out[3 * found + 2] = work.rejects;
                }
                found++;
            }
        }
    }
    free(perm);
    return found;
}

