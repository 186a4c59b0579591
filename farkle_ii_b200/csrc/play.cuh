// play.cuh — the persistent one-game-per-lane Farkle state machine.
//
// Replaces the interpreter loops of the reference:
//   FarklePlayer.take_turn / _score_roll / _apply_hot_dice / _should_continue
//                                         src/farkle/game/engine.py:103-273
//   ThresholdStrategy.decide              src/farkle/simulation/strategies.py:212-275
//   FarkleGame.play / _run_final_round    src/farkle/game/engine.py:436-550
//   _play_game row flattening             src/farkle/simulation/simulation.py:576-655
//   OutcomeCounter.record_row + winner metric sums
//                                         src/farkle/simulation/run_tournament.py:177-195,375-391
//
// Execution model
//   * one game per lane; the loop body is ONE ROLL, so turn / seat / round /
//     final-round changes are state transitions inside a warp-uniform loop;
//   * the active seat lives in registers, parked seats in a per-warp shared
//     memory store laid out [word][33] (stride 33 => the owner lane's column
//     access and the cooperative row access of the epilogue are both
//     bank-conflict free);
//   * persistent CTAs (one per SM): a lane whose game ended takes the next game
//     ordinal from a global counter (lane refill), so the 100x spread of game
//     lengths does not idle the warp;
//   * game end is handled by the whole warp for one lane at a time: lane s ranks
//     seat s, lane w writes word w of the compact row (one coalesced store),
//     lanes 0..22 issue the winner's tally REDs in a single instruction.
#pragma once
#include <cstdint>

#include "rng.cuh"
#include "scoring.cuh"

namespace fb {

constexpr int ROLL_LIMIT = 1000;  // src/farkle/game/engine.py:36

// parked seat record, 32-bit words
enum SeatWord {
    W_LO0 = 0, W_LO1, W_HI0, W_HI1,  // PCG state (lo, hi)
    W_ILO0, W_ILO1, W_IHI0, W_IHI1,  // PCG increment
    W_SAVED,                          // buffered high half of the last 64-bit draw
    W_SCORE,
    W_HIGH,   // highest_turn | has32 << 30 | has_scored << 31
    W_FR,     // n_farkles | n_rolls << 16
    W_TH,     // n_turns | n_hot_dice << 16
    W_SF,     // smart_five_uses | n_smart_five_dice << 16
    W_SO,     // smart_one_uses | n_smart_one_dice << 16
    W_P0,     // score_threshold
    W_P1,     // (u16)dice_threshold | flags << 16
    SEAT_WORDS
};
constexpr int STORE_STRIDE = 33;
constexpr uint32_t HIGH_MASK = 0x3fffffffu;

struct PlayParams {
    const uint32_t* seat_state;       // [n_games*k][8]: lo, hi, ilo, ihi as u32 pairs
    const int32_t* seat_strat;        // [n_games*k] index into `strategies`
    const fb_strategy_t* strategies;  // table
    const int32_t* strategy_ids;      // id of table entry (ids_mode 1)
    int ids_mode;                     // 0 id = index, 1 id = strategy_ids[index], 2 id = seat
    const uint64_t* game_seed;        // [n_games] or nullptr
    const int32_t* limits;            // [n_games][2] {target, max_rounds} or nullptr
    int32_t target_score, max_rounds;
    uint32_t n_games;
    int k;
    uint32_t games_per_slot;  // 0 = one tally slot
    int n_tally_ids;
    unsigned long long* tallies;  // [slots][ids][26] or nullptr
    unsigned long long* totals;   // [FB_TOTALS_WIDTH] or nullptr
    uint32_t* rows;               // or nullptr
    int row_words;
    uint8_t* outcome;  // [n_games] or nullptr
    unsigned int* counter;
};

enum LaneStatus { ST_NEED = 0, ST_LOAD = 1, ST_PLAY = 2, ST_DEAD = 3 };

constexpr uint32_t FULL = 0xffffffffu;

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) play_kernel(const PlayParams P, const ScoreLut* __restrict__ lut_g) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ unsigned long long s_tot[FB_TOTALS_WIDTH];
    ScoreLut* lut = reinterpret_cast<ScoreLut*>(smem_raw);
    uint32_t* store = reinterpret_cast<uint32_t*>(smem_raw + ((LUT_BYTES + 15) & ~15));

    for (int i = threadIdx.x; i < LUT_BYTES / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(lut)[i] = reinterpret_cast<const uint32_t*>(lut_g)[i];
    if (threadIdx.x < FB_TOTALS_WIDTH) s_tot[threadIdx.x] = 0ull;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = P.k;
    const int seat_words = k * SEAT_WORDS;
    uint32_t* ws = store + (size_t)warp * seat_words * STORE_STRIDE;
    const uint32_t lt_mask = (1u << lane) - 1u;

#define SEATW(seat_, w_) ws[((seat_) * SEAT_WORDS + (w_)) * STORE_STRIDE + lane]

    // ---- per-lane game state ------------------------------------------------
    int status = ST_NEED;
    bool have_result = false, newgame = false;
    uint32_t g = 0, err = 0;
    int seat = 0, round = 0, trigger = -1, stb = 0, target = 0, max_rounds = 0;
    // active seat
    Pcg rng{0, 0, 0, 0};
    uint32_t saved = 0;
    bool has32 = false, has_scored = false;
    int score = 0, highest = 0, st = 0;
    uint32_t c_fr = 0, c_th = 0, c_sf = 0, c_so = 0, p1 = 0, dbase = 0;
    // active turn
    int ts = 0, dice = 6, rolls_turn = 0;
    // work/total accumulators (reduced once at kernel end)
    uint32_t a_done = 0, a_safe = 0, a_err = 0, a_swins = 0, a_dice = 0, a_words = 0;
    unsigned long long a_rolls = 0, a_turns = 0;

    for (;;) {
        // ================= R: game end (cooperative) + lane refill =============
        const uint32_t need = __ballot_sync(FULL, status == ST_NEED);
        if (need) {
            __syncwarp();  // parked seat records of finished lanes are visible to the warp
            uint32_t fin = __ballot_sync(FULL, status == ST_NEED && have_result);
            while (fin) {
                const int f = __ffs(fin) - 1;
                fin &= fin - 1;
                const uint32_t gf = __shfl_sync(FULL, g, f);
                const int roundf = __shfl_sync(FULL, round, f);
                const int trigf = __shfl_sync(FULL, trigger, f);
                const uint32_t errf = __shfl_sync(FULL, err, f);
                const bool safety = trigf < 0;
                const uint32_t* col = ws + f;  // column of lane f
                // winner = highest score, ties to the lower seat (stable sort, engine.py:483)
                int sc = 0;
                int sid = 0;
                if (lane < k) {
                    sc = (int)col[(lane * SEAT_WORDS + W_SCORE) * STORE_STRIDE];
                    const uint32_t idx = P.seat_strat ? (uint32_t)P.seat_strat[(size_t)gf * k + lane]
                                                      : gf * (uint32_t)k + lane;
                    sid = P.ids_mode == 0 ? (int)idx : (P.ids_mode == 1 ? P.strategy_ids[idx] : lane);
                }
                const uint32_t key = lane < k ? (((uint32_t)sc << 4) | (uint32_t)(15 - lane)) : 0u;
                const uint32_t best = __reduce_max_sync(FULL, key);
                const int winner = safety ? 0xFF : 15 - (int)(best & 15u);
                uint32_t i16 = 0;
                if (lane < k) {
                    const uint32_t fr = col[(lane * SEAT_WORDS + W_FR) * STORE_STRIDE];
                    const uint32_t th = col[(lane * SEAT_WORDS + W_TH) * STORE_STRIDE];
                    const uint32_t sfw = col[(lane * SEAT_WORDS + W_SF) * STORE_STRIDE];
                    const uint32_t sow = col[(lane * SEAT_WORDS + W_SO) * STORE_STRIDE];
                    const uint32_t hi = col[(lane * SEAT_WORDS + W_HIGH) * STORE_STRIDE] & HIGH_MASK;
                    a_rolls += fr >> 16;
                    a_turns += th & 0xffffu;
                    i16 = ((fr >> 16) > 32767u) | (hi > 32767u) | ((sfw >> 16) > 32767u) |
                          ((sow >> 16) > 32767u) | ((th & 0xffffu) > 32767u) | ((th >> 16) > 32767u);
                }
                const bool any16 = __any_sync(FULL, i16 != 0);
                const uint32_t flags = (safety ? FB_ROW_SAFETY_LIMIT : 0u) | errf |
                                       (any16 ? FB_ROW_I16_OVERFLOW : 0u);
                if (lane == 0) {
                    a_done += 1;
                    a_safe += safety ? 1u : 0u;
                    a_err += (flags & (FB_ROW_ROLL_LIMIT | FB_ROW_I16_OVERFLOW)) ? 1u : 0u;
                    if (P.outcome)
                        P.outcome[gf] = (uint8_t)((safety ? 0 : winner + 1) |
                                                  ((flags & ~FB_ROW_SAFETY_LIMIT) ? 0x80 : 0));
                }
                if (lane == winner) a_swins += 1;

                // ---- compact row: lane w writes word w (coalesced) ----
                if (P.rows) {
                    uint32_t* row = P.rows + (size_t)gf * P.row_words;
                    for (int base = 0; base < P.row_words; base += 32) {
                        const int w = base + lane;
                        const int t = w - 4;
                        const int s = t >= 0 ? t / 7 : 0;
                        const int j = t - s * 7;
                        const int sidv = __shfl_sync(FULL, sid, s & 31);
                        uint32_t val = 0;
                        if (w < 4) {
                            if (w < 2) {
                                const uint64_t gs = P.game_seed ? P.game_seed[gf] : 0ull;
                                val = w == 0 ? (uint32_t)gs : (uint32_t)(gs >> 32);
                            } else if (w == 2) {
                                val = gf;
                            } else {
                                val = (uint32_t)roundf | ((uint32_t)winner << 16) | (flags << 24);
                            }
                        } else if (s < k) {
                            const uint32_t* sp = col + (s * SEAT_WORDS) * STORE_STRIDE;
                            switch (j) {
                                case 0: val = sp[W_SCORE * STORE_STRIDE]; break;
                                case 1: val = (uint32_t)sidv; break;
                                case 2: val = sp[W_HIGH * STORE_STRIDE] & HIGH_MASK; break;
                                case 3: val = sp[W_FR * STORE_STRIDE]; break;
                                case 4: val = sp[W_TH * STORE_STRIDE]; break;
                                case 5: val = sp[W_SF * STORE_STRIDE]; break;
                                default: val = sp[W_SO * STORE_STRIDE]; break;
                            }
                        }
                        if (w < P.row_words) row[w] = val;
                    }
                }
                // ---- tallies: REDs spread over the lanes ----
                if (P.tallies) {
                    const uint32_t slot = P.games_per_slot ? gf / P.games_per_slot : 0u;
                    unsigned long long* T =
                        P.tallies + (size_t)slot * (size_t)P.n_tally_ids * FB_TALLY_WIDTH;
                    if (lane < k) {
                        atomicAdd(&T[(size_t)sid * FB_TALLY_WIDTH + 1], 1ull);
                        atomicAdd(&T[(size_t)sid * FB_TALLY_WIDTH + (safety ? 3 : 2)], 1ull);
                    }
                    if (!safety) {
                        const int wsid = __shfl_sync(FULL, sid, winner);
                        const uint32_t* sp = col + (winner * SEAT_WORDS) * STORE_STRIDE;
                        const int j = lane < FB_N_METRICS ? lane : lane - FB_N_METRICS;
                        unsigned long long v = 0;
                        switch (j) {  // METRIC_LABELS order, run_tournament.py:109-121
                            case 0: v = sp[W_SCORE * STORE_STRIDE]; break;
                            case 1: v = (unsigned)roundf; break;
                            case 2: v = sp[W_FR * STORE_STRIDE] & 0xffffu; break;
                            case 3: v = sp[W_FR * STORE_STRIDE] >> 16; break;
                            case 4: v = sp[W_HIGH * STORE_STRIDE] & HIGH_MASK; break;
                            case 5: v = sp[W_SF * STORE_STRIDE] & 0xffffu; break;
                            case 6: v = sp[W_SF * STORE_STRIDE] >> 16; break;
                            case 7: v = sp[W_SO * STORE_STRIDE] & 0xffffu; break;
                            case 8: v = sp[W_SO * STORE_STRIDE] >> 16; break;
                            case 9: v = sp[W_TH * STORE_STRIDE] >> 16; break;
                            default: v = 0; break;  // winner_hit_max_rounds: False when completed
                        }
                        unsigned long long* Tw = T + (size_t)wsid * FB_TALLY_WIDTH;
                        if (lane < FB_N_METRICS) atomicAdd(&Tw[4 + j], v);
                        else if (lane < 2 * FB_N_METRICS) atomicAdd(&Tw[4 + FB_N_METRICS + j], v * v);
                        else if (lane == 2 * FB_N_METRICS) atomicAdd(&Tw[0], 1ull);
                    }
                }
            }
            // ---- refill: contiguous ordinals for the lanes that need a game ----
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.counter, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, 0);
            if (status == ST_NEED) {
                have_result = false;
                const uint32_t ng = base + (uint32_t)__popc(need & lt_mask);
                if (ng < P.n_games) {
                    g = ng;
                    status = ST_LOAD;
                    newgame = true;
                } else {
                    status = ST_DEAD;
                }
            }
            uint32_t fresh = __ballot_sync(FULL, newgame);
            while (fresh) {
                const int f = __ffs(fresh) - 1;
                fresh &= fresh - 1;
                const uint32_t gf = __shfl_sync(FULL, g, f);
                for (int base_w = 0; base_w < seat_words; base_w += 32) {
                    const int w = base_w + lane;
                    if (w < seat_words) {
                        const int s = w / SEAT_WORDS;
                        const int fld = w - s * SEAT_WORDS;
                        uint32_t val = 0;
                        if (fld < 8) {
                            val = P.seat_state[((size_t)gf * k + s) * 8 + fld];
                        } else if (fld >= W_P0) {
                            const uint32_t idx = P.seat_strat ? (uint32_t)P.seat_strat[(size_t)gf * k + s]
                                                              : gf * (uint32_t)k + s;
                            const uint2 sv = reinterpret_cast<const uint2*>(P.strategies)[idx];
                            val = fld == W_P0 ? sv.x : sv.y;
                        }
                        ws[w * STORE_STRIDE + f] = val;
                    }
                }
            }
            __syncwarp();
            if (newgame) {
                newgame = false;
                seat = 0;
                round = 0;
                trigger = -1;
                err = 0;
                target = P.limits ? P.limits[2 * (size_t)g] : P.target_score;
                max_rounds = P.limits ? P.limits[2 * (size_t)g + 1] : P.max_rounds;
                stb = target;
                if (max_rounds <= 0) {  // while rounds < max_rounds never runs (engine.py:455)
                    status = ST_NEED;
                    have_result = true;
                }
            }
        }
        if (__all_sync(FULL, status == ST_DEAD)) break;

        // ================= L: seat the next player, start the turn ==============
        if (status == ST_LOAD) {
            if (trigger < 0 && seat == 0) round++;
            rng.lo = (uint64_t)SEATW(seat, W_LO0) | ((uint64_t)SEATW(seat, W_LO1) << 32);
            rng.hi = (uint64_t)SEATW(seat, W_HI0) | ((uint64_t)SEATW(seat, W_HI1) << 32);
            rng.ilo = (uint64_t)SEATW(seat, W_ILO0) | ((uint64_t)SEATW(seat, W_ILO1) << 32);
            rng.ihi = (uint64_t)SEATW(seat, W_IHI0) | ((uint64_t)SEATW(seat, W_IHI1) << 32);
            saved = SEATW(seat, W_SAVED);
            score = (int)SEATW(seat, W_SCORE);
            const uint32_t hw = SEATW(seat, W_HIGH);
            highest = (int)(hw & HIGH_MASK);
            has32 = (hw >> 30) & 1u;
            has_scored = hw >> 31;
            c_fr = SEATW(seat, W_FR);
            c_th = SEATW(seat, W_TH) + 1u;  // n_turns += 1 (engine.py:237)
            c_sf = SEATW(seat, W_SF);
            c_so = SEATW(seat, W_SO);
            st = (int)SEATW(seat, W_P0);
            p1 = SEATW(seat, W_P1);
            dbase = disc_base(p1);
            dice = 6;
            ts = 0;
            rolls_turn = 0;
            status = ST_PLAY;
        }

        // ================= P: one roll (straight-line, no divergent branches) =====
        if (status == ST_PLAY) {
            const int n = dice;
            // -- dice: n consecutive 32-bit halves of the seat's stream (engine.py:101).
            // Halves H0..H6 = [buffered half, lo/hi of up to three fresh 64-bit outputs];
            // die i reads H[i + p].  All three outputs are always computed; the state only
            // advances past the nw words this roll really consumes.
            const uint32_t p = has32 ? 0u : 1u;
            const int nw = (n + (int)p) >> 1;
            uint64_t shi = rng.hi, slo = rng.lo;
            const uint64_t o1 = pcg_output(shi, slo);
            {
                uint64_t th = shi, tl = slo;
                pcg_step(th, tl, rng.ihi, rng.ilo);
                const bool c = nw > 0;
                shi = c ? th : shi;
                slo = c ? tl : slo;
            }
            const uint64_t o2 = pcg_output(shi, slo);
            {
                uint64_t th = shi, tl = slo;
                pcg_step(th, tl, rng.ihi, rng.ilo);
                const bool c = nw > 1;
                shi = c ? th : shi;
                slo = c ? tl : slo;
            }
            const uint64_t o3 = pcg_output(shi, slo);
            {
                uint64_t th = shi, tl = slo;
                pcg_step(th, tl, rng.ihi, rng.ilo);
                const bool c = nw > 2;
                shi = c ? th : shi;
                slo = c ? tl : slo;
            }
            const uint32_t H1 = (uint32_t)o1, H2 = (uint32_t)(o1 >> 32);
            const uint32_t H3 = (uint32_t)o2, H4 = (uint32_t)(o2 >> 32);
            const uint32_t H5 = (uint32_t)o3, H6 = (uint32_t)(o3 >> 32);
            uint32_t hist = 0;
            uint32_t minlo = 0xffffffffu;  // Lemire leftover; < 4 means NumPy redraws
#define FB_DIE(i_, Ha_, Hb_)                                       \
    {                                                              \
        const uint32_t u_ = p ? (Hb_) : (Ha_);                     \
        const uint64_t m_ = (uint64_t)u_ * 6u;                     \
        minlo = min(minlo, (uint32_t)m_);                          \
        if ((i_) < n) hist += 1u << (3u * (uint32_t)(m_ >> 32));   \
    }
            FB_DIE(0, saved, H1)
            FB_DIE(1, H1, H2)
            FB_DIE(2, H2, H3)
            FB_DIE(3, H3, H4)
            FB_DIE(4, H4, H5)
            FB_DIE(5, H5, H6)
#undef FB_DIE
            const uint32_t q = p + (uint32_t)n;  // index of the first unread half
            bool nhas = (q & 1u) == 0u;
            uint32_t nsaved = q == 2u ? H2 : (q == 4u ? H4 : H6);
            uint32_t words = (uint32_t)nw;
            if (minlo < 4u) {
                // A draw with leftover < 4 somewhere in the window (4 in 2^32 per die; unused
                // slots can only add false alarms): replay this roll draw by draw.
                PcgStream s{rng, saved, has32};
                hist = 0;
                words = 0;
                for (int i = 0; i < n; i++) hist += 1u << (3u * s.die0(words));
                shi = s.g.hi;
                slo = s.g.lo;
                nhas = s.has32;
                nsaved = s.saved;
            }
            rng.hi = shi;
            rng.lo = slo;
            has32 = nhas;
            saved = nsaved;
            a_dice += (uint32_t)n;
            a_words += words;

            // -- score the roll (engine.py:103-147), discards, counters
            const uint32_t e = lut_lookup(lut, hist);
            const int rscore = (int)(e & 127u) * 50;
            const int used0 = (int)((e >> 7) & 7u);
            const int sf = (int)((e >> 10) & 3u), so = (int)((e >> 12) & 3u);
            const bool farkle = rscore == 0;
            const uint32_t dd = smart_discards(lut, dbase, rscore, used0, sf, so, n, ts, st, p1);
            const uint32_t d5 = dd & 3u, d1 = dd >> 2;  // both 0 on a farkle
            const int pts = rscore - 50 * (int)d5 - 100 * (int)d1;
            const int used = used0 - (int)d5 - (int)d1;
            c_sf += (d5 << 16) + ((d5 + 1u) >> 1);  // uses += (d5 > 0), dice += d5
            c_so += (d1 << 16) + ((d1 + 1u) >> 1);
            c_fr += 0x10000u + (farkle ? 1u : 0u);  // n_rolls += 1, n_farkles += farkle
            rolls_turn++;
            const int ndice = used == n ? 6 : n - used;
            const int ts2 = ts + pts;
            // -- hot dice (engine.py:149-154), then _should_continue (engine.py:156-205) and
            //    ThresholdStrategy.decide (strategies.py:212-275)
            const bool hot = !farkle && ndice == 6 && strat_flag(p1, FB_SF_AUTO_HOT_DICE);
            c_th += hot ? 0x10000u : 0u;
            const bool fin = trigger >= 0;
            const int rt = score + ts2;
            const bool behind = fin && rt <= stb;
            const bool stop_ahead = fin && rt > stb && !strat_flag(p1, FB_SF_RUN_UP_SCORE);
            const bool gate = !has_scored && ts2 < 500;
            const bool keep = !stop_ahead && (gate || behind || decide_continue(ts2, ndice, st, p1));
            bool turn_over = farkle || (!hot && !keep);
            ts = farkle ? 0 : ts2;
            dice = ndice;
            if (!turn_over && rolls_turn >= ROLL_LIMIT) {  // engine.py:242-243 raises
                err |= FB_ROW_ROLL_LIMIT;
                turn_over = true;
            }

            // ================= T: bank, park the seat, pick the next one ========
            if (turn_over) {
                if (!has_scored && ts >= 500) has_scored = true;
                if (has_scored) {
                    score += ts;
                    highest = max(highest, ts);
                }
                SEATW(seat, W_LO0) = (uint32_t)rng.lo;
                SEATW(seat, W_LO1) = (uint32_t)(rng.lo >> 32);
                SEATW(seat, W_HI0) = (uint32_t)rng.hi;
                SEATW(seat, W_HI1) = (uint32_t)(rng.hi >> 32);
                SEATW(seat, W_SAVED) = saved;
                SEATW(seat, W_SCORE) = (uint32_t)score;
                SEATW(seat, W_HIGH) = (uint32_t)highest | ((uint32_t)has32 << 30) | ((uint32_t)has_scored << 31);
                SEATW(seat, W_FR) = c_fr;
                SEATW(seat, W_TH) = c_th;
                SEATW(seat, W_SF) = c_sf;
                SEATW(seat, W_SO) = c_so;
                bool over;
                if (trigger < 0) {
                    if (score >= target) {  // first trigger starts the final round (engine.py:466-471)
                        trigger = seat;
                        stb = score;
                        seat = seat == 0 ? 1 : 0;
                        over = seat >= k;
                    } else {
                        seat++;
                        over = false;
                        if (seat == k) {
                            seat = 0;
                            over = round >= max_rounds;
                        }
                    }
                } else {
                    if (score > stb) stb = score;  // engine.py:547-548
                    seat++;
                    if (seat == trigger) seat++;
                    over = seat >= k;
                }
                if (err) over = true;
                if (over) {
                    status = ST_NEED;
                    have_result = true;
                } else {
                    status = ST_LOAD;
                }
            }
        }
    }
#undef SEATW

    // ---- totals: warp shuffle -> shared memory -> one global RED per CTA ------
    unsigned long long v[8] = {a_done, (unsigned long long)a_done - a_safe, a_safe, a_rolls,
                               a_dice,  a_words, a_turns, a_err};
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(FULL, v[i], o);
        if (lane == 0 && v[i]) atomicAdd(&s_tot[i], v[i]);
    }
    if (lane < k && a_swins) atomicAdd(&s_tot[8 + lane], (unsigned long long)a_swins);
    __syncthreads();
    if (P.totals && threadIdx.x < FB_TOTALS_WIDTH && s_tot[threadIdx.x])
        atomicAdd(&P.totals[threadIdx.x], s_tot[threadIdx.x]);
}

}  // namespace fb
