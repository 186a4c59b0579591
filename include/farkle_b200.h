/* farkle_b200.h — C ABI of the B200-native Farkle Monte-Carlo engine.
 *
 * This is the drop-in boundary for ONE hot path of Isaac-McPadden/Farkle_II:
 * "simulate whole k-player games for every (strategy, seed, k) cell and reduce
 * them to per-strategy tallies".  The reference has no FFI of its own; the path
 * sits behind Python callables.  Each entry point below names the reference
 * callable (file:line under /root/reference) whose work it replaces.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative fb_status otherwise;
 *    fb_last_error() returns a thread-local message for the last failure.
 *  - No exceptions, no torch types, no ownership transfer: the caller owns every
 *    buffer.  Pointers suffixed _dev are device pointers on the device given to
 *    fb_init(); pointers suffixed _host are host pointers (pinned or pageable).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls only enqueue work; they do not synchronise unless stated.
 *  - There is NO CPU fallback: without a CUDA device every compute call fails
 *    with FB_ERR_NO_DEVICE.
 */
#ifndef FARKLE_B200_H
#define FARKLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------- */
enum fb_status {
    FB_OK = 0,
    FB_ERR_NO_DEVICE = -1,   /* no CUDA device / fb_init not called          */
    FB_ERR_BAD_ARG = -2,     /* argument outside the documented domain        */
    FB_ERR_CUDA = -3,        /* a CUDA runtime call failed (see last_error)   */
    FB_ERR_WORKSPACE = -4,   /* caller workspace too small                    */
    FB_ERR_INTERNAL = -5     /* an internal consistency check failed          */
};

/* ---- RNG purposes (src/farkle/utils/random.py:18-37) --------------------- */
enum fb_purpose {
    FB_PURPOSE_INDEXED_SEED = 1,
    FB_PURPOSE_PLAYER = 10,
    FB_PURPOSE_TOURNAMENT_SHUFFLE = 100,
    FB_PURPOSE_SHUFFLE_PERMUTATION = 101,
    FB_PURPOSE_TOURNAMENT_GAME = 102,
    FB_PURPOSE_TOURNAMENT_PLAYER = 103,
    FB_PURPOSE_H2H_GAME = 202,
    FB_PURPOSE_H2H_PLAYER = 203
};

/* ---- strategy table entry -------------------------------------------------
 * One ThresholdStrategy (src/farkle/simulation/strategies.py:165-290), 8 bytes.
 */
typedef struct fb_strategy {
    int32_t score_threshold;
    int16_t dice_threshold;
    uint16_t flags; /* FB_SF_* */
} fb_strategy_t;

#define FB_SF_SMART_FIVE 0x01u
#define FB_SF_SMART_ONE 0x02u
#define FB_SF_CONSIDER_SCORE 0x04u
#define FB_SF_CONSIDER_DICE 0x08u
#define FB_SF_REQUIRE_BOTH 0x10u
#define FB_SF_AUTO_HOT_DICE 0x20u
#define FB_SF_RUN_UP_SCORE 0x40u
#define FB_SF_FAVOR_SCORE 0x80u /* FavorDiceOrScore.SCORE; clear = DICE */

/* ---- compact per-game row --------------------------------------------------
 * The ingest-ready row of src/farkle/simulation/simulation.py:576-655 (schema
 * src/farkle/utils/schema_helpers.py:23-60) in fixed-width form.  Everything
 * the reference row holds is either here or derivable on the host from it
 * (ranks, margins, seat_ranks, strings).  A row occupies fb_row_stride(k)
 * bytes: one header followed by k seat records, padded to 16 bytes.
 */
typedef struct fb_row_header {
    uint64_t game_seed;   /* diagnostic fingerprint: u32 (purpose 102) for
                             tournaments, u64 (purpose 202) for H2H,
                             0 when fingerprints were not requested          */
    uint32_t game_ordinal; /* index of the game inside the launch            */
    uint16_t n_rounds;
    uint8_t winner_seat;  /* 0-based seat of rank 1; 0xFF = safety limit     */
    uint8_t flags;        /* FB_ROW_*                                        */
} fb_row_header_t;

#define FB_ROW_SAFETY_LIMIT 0x01u /* termination_status == "safety_limit"   */
#define FB_ROW_ROLL_LIMIT 0x02u   /* a turn hit ROLL_LIMIT (engine.py:242): the
                                     reference raises RuntimeError            */
#define FB_ROW_I16_OVERFLOW 0x04u /* a counter left the int16 range of the
                                     Arrow schema (the reference's Arrow
                                     conversion raises)                       */

typedef struct fb_row_seat {
    int32_t score;
    int32_t strategy;     /* strategy id seated here                         */
    int32_t highest_turn;
    uint16_t farkles;
    uint16_t rolls;
    uint16_t n_turns;
    uint16_t hot_dice;
    uint16_t smart_five_uses;
    uint16_t n_smart_five_dice;
    uint16_t smart_one_uses;
    uint16_t n_smart_one_dice;
} fb_row_seat_t;

/* ---- tallies ---------------------------------------------------------------
 * Per strategy id, int64[FB_TALLY_WIDTH]
 * (src/farkle/simulation/run_tournament.py:109-139,165-230,375-391):
 *   0 wins                      1 attempted_exposures
 *   2 completed_exposures       3 safety_limit_exposures
 *   4..14  sum of METRIC_LABELS[i] over games this strategy won
 *   15..25 sum of squares of the same
 * The reference stores the sums as Python floats; every addend is an integer
 * and every total is < 2^53, so float(int64) is bit-identical.
 */
#define FB_TALLY_WIDTH 26
#define FB_N_METRICS 11

/* Optional per-seat tallies, int64[n_slots][n_tally_ids][k][FB_SEAT_TALLY_WIDTH]
 * (the counts src/farkle/analysis/seat_analysis.py:166-229 re-derives from rows):
 *   0 raw_wins  1 raw_exposures  2 raw_completed_exposures
 *   3 raw_safety_limit_exposures        of strategy id x 0-based seat.        */
#define FB_SEAT_TALLY_WIDTH 4

/* Optional RNG lag statistics of the strategy groups (replaces the external sort + online
 * accumulator of src/farkle/analysis/rng_diagnostics.py:2032-2077 for "strategy" groups):
 * int64[n_strategies][n_lags][FB_LAG_WIDTH], indexed by TABLE position (not strategy id).
 * The sequence of table entry i is its seat exposure in shuffle shuffle0, shuffle0+1, ... (the
 * reference's order by (root_seed, k, shuffle_index, game_index, seat_index): a strategy is
 * seated once per shuffle); x is the observation `lag` places earlier, y the current one.
 *   0       lagged pairs
 *   1..5    win indicator: sum x, sum y, sum x^2, sum y^2, sum x*y
 *   6..10   n_rounds:      sum x, sum y, sum x^2, sum y^2, sum x*y
 * Edges, uint32[n_strategies][2][max_lag] with max_lag = the largest lag: the first and the last
 * min(max_lag, n_shuffles) observations of the launch, each n_rounds | win << 16, so that the
 * pairs straddling two launches of one cell can be added by the caller.           */
#define FB_LAG_WIDTH 11
#define FB_MAX_LAGS 8
#define FB_MAX_LAG 4096

/* "Matchup" groups of the same stage: the games of one cell that seat the same multiset of
 * strategies, in (shuffle_index, game_index) order, observation = n_rounds (one per game;
 * src/farkle/analysis/rng_diagnostics.py:1870-1901).  Only groups with at least
 * matchup_min_observations games are written (the reference's eligibility rule is
 * min(lags) + 2), in no particular order:
 *   matchup_participants_dev int32 [capacity][k]   sorted strategy ids of the group
 *   matchup_count_dev        uint32[capacity]      observations (games)
 *   matchup_stats_dev        int64 [capacity][n_lags][FB_MATCHUP_LAG_WIDTH]
 *                            lagged pairs, sum x, sum y, sum x^2, sum y^2, sum x*y of n_rounds
 * capacity >= n_games / matchup_min_observations always suffices.                 */
#define FB_MATCHUP_LAG_WIDTH 6

typedef struct fb_lag_request {
    const int32_t* lags;          /* HOST array: n_lags distinct lags in [1, FB_MAX_LAG]   */
    int32_t n_lags;               /* 1..FB_MAX_LAGS; 0 = no lag statistics at all            */
    int32_t matchup_min_observations; /* 0 = no matchup groups                               */
    int64_t* strategy_stats_dev;  /* accumulated into (caller zeroes); NULL = skip strategy groups */
    uint32_t* strategy_edges_dev; /* overwritten; required with strategy_stats_dev           */
    uint64_t matchup_capacity;    /* groups the three matchup buffers can hold               */
    int32_t* matchup_participants_dev;
    uint32_t* matchup_count_dev;
    int64_t* matchup_stats_dev;   /* zeroed by the call                                      */
    void* scratch_dev;            /* >= fb_matchup_scratch_bytes(n_games) when matchups are on */
    size_t scratch_bytes;
    int64_t* n_matchups_host;     /* HOST out: groups written; the call synchronises `stream` */
    /* Independent of the lags (n_lags may be 0 when only this is wanted): uint32[n_tally_ids][4],
     * overwritten -- the first exposure ordinal (shuffle - shuffle0) * n_strategies + game * k + seat
     * at which the id  0: won  1: was seated  2: was seated in a completed game  3: was seated in a
     * safety-limit game;  0xFFFFFFFF = never.  It is the insertion order of the reference's
     * Counter / dict keys (run_tournament.py:177-195,375-391), which a byte-identical checkpoint
     * pickle needs.                                                                           */
    uint32_t* first_seen_dev;
    /* Unconditional all-player sufficient statistics per (deterministic batch, strategy)
     * (src/farkle/analysis/all_player_metrics.py:31-98,262-340: what the reference's metrics stage
     * re-derives by re-reading every curated row).  int64 [slots][n_tally_ids][FB_ALLP_WIDTH],
     * OVERWRITTEN for the slots of this launch; needs shuffles_per_slot > 0 (slot = batch).
     * Columns: 0 exposures  1 completed  2 safety-limit  3 wins  4 turn/round mismatches
     *   5,6 sum / square sum of final score   7,8 of n_turns   9,10 of (n_turns - n_rounds)
     *   11 + 3b .. 13 + 3b  observations, sum, square sum of behaviour b in the reference's order
     *   (rank, loss_margin, rolls, farkles, highest_turn, hot_dice, smart_five_uses,
     *   n_smart_five_dice, smart_one_uses, n_smart_one_dice; rank and loss_margin are observed in
     *   completed games only)
     *   41..44  IEEE-754 double BIT PATTERNS: sum and square sum of score / n_turns, then of
     *   score / n_rounds, added in shuffle order with round-to-nearest divisions, products and
     *   additions, i.e. the exact value of the reference's float64 accumulator.              */
    int64_t* all_player_dev;
} fb_lag_request_t;
#define FB_ALLP_WIDTH 45
#define FB_ALLP_BEHAVIOURS 10

/* Per-launch totals, int64[FB_TOTALS_WIDTH]:
 *   0 games_attempted  1 games_completed  2 games_safety_limit
 *   3 rolls  4 dice  5 rng_words (64-bit PCG64DXSM outputs consumed by dice)
 *   6 turns  7 error rows (FB_ROW_ROLL_LIMIT | FB_ROW_I16_OVERFLOW)
 *   8..8+FB_MAX_PLAYERS-1  wins by 0-based seat
 */
#define FB_MAX_PLAYERS 12
/* Largest max_rounds a launch accepts: n_rounds is an int16 row column
 * (utils/schema_helpers.py:23-42) and 16 bits of the game header.  Scalar arguments above it
 * return FB_ERR_BAD_ARG; per-game values above it are cut to FB_MAX_ROUNDS + 1 and a game
 * that actually plays that many rounds is reported with FB_ROW_I16_OVERFLOW.             */
#define FB_MAX_ROUNDS 32767
#define FB_TOTALS_WIDTH (8 + FB_MAX_PLAYERS)

/* ---- library ---------------------------------------------------------------*/
int fb_abi_version(void);
const char* fb_last_error(void);

/* Make `device` current for the calling thread (cudaSetDevice) and build its context: score
 * lookup table, jump-ahead constants, cached streams / buffers.  Idempotent per device.  A process
 * may initialise several devices; every other entry point works on the context of the device
 * that is CURRENT for the calling thread when it is called (as the CUDA runtime resolves streams
 * and allocations), so a host that drives several GPUs from one process calls cudaSetDevice(d)
 * (or fb_init(d)) before the calls meant for device d.  Pointers and the stream passed to a call
 * must belong to that device.  Calls on different devices do not serialise each other.        */
int fb_init(int device);
/* sm_count / clock_khz may be NULL. */
int fb_device_info(int* sm_count, int* clock_khz, int* cc_major, int* cc_minor);

size_t fb_row_stride(int k);
/* Bytes of device workspace one launch over n_games k-player games needs. */
size_t fb_workspace_bytes(int k, uint64_t n_games);

/* ---- building blocks (exposed for parity tests and the host mirror) ------- */

/* SeedSequence(entropy).generate_state(n_words, uint32) for n_streams entropy
 * rows of n_entropy uint32 words each.  Replaces the NumPy call made at
 * src/farkle/utils/random.py:156,225.                                        */
int fb_seedseq_generate(const uint32_t* entropy_dev, int n_entropy, uint64_t n_streams,
                        int n_words, uint32_t* out_dev, void* stream);

/* coordinate_seed(purpose, ...) fingerprints (src/farkle/utils/random.py:191-225)
 * for index i in [0,n): the coordinate named by `vary` (0 shuffle_index,
 * 1 game_index, 2 pair_id) is base + i, the others are fixed.  out is uint64;
 * as_u32 != 0 returns the uint32 fingerprint (zero-extended).               */
int fb_coordinate_seeds(uint32_t purpose, uint64_t root_seed, uint64_t k, uint64_t shuffle_index,
                        uint64_t pair_id, uint64_t order, uint64_t game_index, int vary,
                        uint64_t base, uint64_t n, int as_u32, uint64_t* out_dev, void* stream);

/* PCG64DXSM (state, inc) of coordinate_rng(...) (src/farkle/utils/random.py:159-188)
 * for n explicit coordinate rows coords[i] = {purpose, root_seed, k,
 * shuffle_index, pair_id, order, game_index, seat_index, replicate_index}.
 * out[i] = {state_hi, state_lo, inc_hi, inc_lo}.                            */
int fb_seed_streams(const uint64_t* coords_dev /*[n][9]*/, uint64_t n,
                    uint64_t* state_inc_out_dev /*[n][4]*/, void* stream);

/* FarklePlayer._roll (src/farkle/game/engine.py:85-101): for each of n streams
 * (state_inc as produced by fb_seed_streams) draw n_rolls rolls of n_dice[r]
 * dice; faces_out[i][r][6] (unused slots 0).  half_buffer_dev[i] = {has_uint32,
 * uinteger} is NumPy's buffered 32-bit half at the start (NULL = empty); with
 * crafted states this is the hook that tests the Lemire rejection branch.    */
int fb_roll_dice(const uint64_t* state_inc_dev, const uint32_t* half_buffer_dev, uint64_t n,
                 const int32_t* n_dice_dev, int n_rolls, uint8_t* faces_out_dev, void* stream);

/* default_score(..., return_discards=True) (src/farkle/game/scoring.py:618-693)
 * for n rolls: faces[i][6] (0 = no die), turn_score_pre[i], strategy[i].
 * out[i] = {final_score, final_used, final_reroll, discard_fives, discard_ones}. */
int fb_default_score(const uint8_t* faces_dev, const int32_t* turn_score_pre_dev,
                     const fb_strategy_t* strategy_dev, uint64_t n, int32_t* out_dev,
                     void* stream);

/* Generator.permutation(n_strategies) of the SHUFFLE_PERMUTATION stream
 * (src/farkle/simulation/run_tournament.py:312-318) for shuffles
 * shuffle0 .. shuffle0+n_shuffles-1.  perm_out[j][n_strategies].            */
int fb_permute_shuffles(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                        int n_strategies, int32_t* perm_out_dev, void* stream);

/* ---- the hot path --------------------------------------------------------- */

/* Tournament shuffles.  Replaces _play_one_shuffle / _run_chunk /
 * _run_chunk_metrics (src/farkle/simulation/run_tournament.py:301-585) for
 * shuffles shuffle0 .. shuffle0+n_shuffles-1 of the (root_seed, k) cell:
 * permutation -> n_strategies/k games per shuffle -> rows + tallies.
 *
 *  strategies_dev    [n_strategies] table; the id of entry i is strategy_ids_dev[i]
 *                    (NULL = identity) and must lie in [0, n_tally_ids).
 *  target_score      10,000 in production (GameProfile default).
 *  max_rounds        200 in production.
 *  override_*        n_overrides per-game max_rounds overrides
 *                    (src/farkle/simulation/game_profile.py:142-160), keyed by
 *                    (shuffle_index, game_index); may be NULL/0.
 *  shuffles_per_slot tallies of shuffle s go to slot (s - shuffle0) /
 *                    shuffles_per_slot; 0 = a single slot.
 *  tallies_dev       int64 [n_slots][n_tally_ids][FB_TALLY_WIDTH], ACCUMULATED into
 *                    (caller zeroes); may be NULL.
 *  totals_dev        int64 [FB_TOTALS_WIDTH], accumulated into; may be NULL.
 *  rows_dev          n_shuffles * (n_strategies/k) rows of fb_row_stride(k)
 *                    bytes in (shuffle, game) order; may be NULL.
 *  want_game_seeds   also compute the purpose-102 fingerprints for rows.
 *  workspace_dev     >= fb_workspace_bytes(k, n_shuffles * (n_strategies/k))
 *                    + 2 * align256(n_shuffles * n_strategies * 4) bytes (the
 *                    permutations and their inverses).
 */
int fb_play_tournament(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                       const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                       int n_strategies, int n_tally_ids, int32_t target_score,
                       int32_t max_rounds, const uint64_t* override_shuffle_dev,
                       const uint32_t* override_game_dev, const int32_t* override_max_rounds_dev,
                       int n_overrides, int shuffles_per_slot, int64_t* tallies_dev,
                       int64_t* totals_dev, void* rows_dev, int want_game_seeds,
                       void* workspace_dev, size_t workspace_bytes, void* stream);

/* fb_play_tournament plus the per-seat tallies (seat_tallies_dev accumulated into, caller
 * zeroes; NULL = plain fb_play_tournament; needs tallies_dev).                */
int fb_play_tournament_seats(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                             const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                             int n_strategies, int n_tally_ids, int32_t target_score,
                             int32_t max_rounds, const uint64_t* override_shuffle_dev,
                             const uint32_t* override_game_dev,
                             const int32_t* override_max_rounds_dev, int n_overrides,
                             int shuffles_per_slot, int64_t* tallies_dev, int64_t* totals_dev,
                             void* rows_dev, int want_game_seeds, int64_t* seat_tallies_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream);

/* fb_play_tournament_seats plus the RNG lag statistics (layouts above) of the strategy groups
 * and / or the matchup groups of the launch, and / or the first-seen ordinals.  Needs
 * tallies_dev (the winner marks and inverse permutations come with it).  lag == NULL is plain
 * fb_play_tournament_seats.                                                          */
int fb_play_tournament_lags(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                            const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                            int n_strategies, int n_tally_ids, int32_t target_score,
                            int32_t max_rounds, const uint64_t* override_shuffle_dev,
                            const uint32_t* override_game_dev,
                            const int32_t* override_max_rounds_dev, int n_overrides,
                            int shuffles_per_slot, int64_t* tallies_dev, int64_t* totals_dev,
                            void* rows_dev, int want_game_seeds, int64_t* seat_tallies_dev,
                            const fb_lag_request_t* lag, void* workspace_dev,
                            size_t workspace_bytes, void* stream);

/* Scratch bytes the matchup grouping of n_games games needs (sort buffers, segment tables). */
/* ---- several (root, k) cells per call, pipelined --------------------------------------
 * One cell of a tournament run: shuffles shuffle0 .. shuffle0 + n_shuffles - 1 of (root_seed, k)
 * (the unit `run_tournament` is called with once per k, simulation/runner.py:1326-1754; a rank's
 * share of a cell in a multi-GPU run).  tallies_dev / totals_dev are accumulated into, as in
 * fb_play_tournament; either may be NULL.                                                     */
typedef struct fb_cell {
    uint64_t root_seed;
    uint64_t shuffle0;
    int32_t k;
    int32_t n_shuffles;
    int64_t* tallies_dev; /* int64 [slots][n_tally_ids][FB_TALLY_WIDTH] */
    int64_t* totals_dev;  /* int64 [FB_TOTALS_WIDTH] */
} fb_cell_t;
/* Workspace for fb_play_tournament_cells: two slots, each large enough for the largest cell. */
size_t fb_cells_workspace_bytes(const fb_cell_t* cells, int n_cells, int n_strategies);
/* Plays cells[0 .. n_cells) in order on `stream` with the results of fb_play_tournament called
 * once per cell, but pipelined: the permutations and seat seeding of cell i+1 run on an internal
 * stream, in the other workspace slot, released by the end of cell i's play_kernel, so that they
 * run beside the finish / tally passes of cell i (they only touch the workspace; nothing can run
 * beside the persistent play_kernel itself).  n_ahead = 1 additionally PREPARES
 * cells[n_cells] without playing it; the next call on the same workspace whose first cell equals
 * it (same strategies, limits) starts playing at once.  A runner that knows its cell list passes
 * it whole; one that is driven cell by cell passes the next cell as look-ahead.  No rows, seat
 * tallies, overrides or lag statistics on this entry point (use fb_play_tournament_lags).      */
int fb_play_tournament_cells(const fb_cell_t* cells, int n_cells, int n_ahead,
                             const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                             int n_strategies, int n_tally_ids, int32_t target_score,
                             int32_t max_rounds, int shuffles_per_slot, void* workspace_dev,
                             size_t workspace_bytes, void* stream);

size_t fb_matchup_scratch_bytes(uint64_t n_games);

/* Head-to-head attempts.  Replaces the attempt loop of
 * _simulate_block_from_manifest (src/farkle/analysis/h2h_schedule.py:1149-1243)
 * for n_blocks blocks: block b plays attempts attempt0[b] .. attempt0[b] +
 * n_attempts[b] - 1 of (root_seed, pair_id[b], order[b]) with seat1/seat2
 * strategies.  outcome_out holds one byte per attempt in block-major order:
 * 0 safety limit, 1 P1 won, 2 P2 won, |0x80 if the row carries an error flag.
 * The reference's early stop ("break once games_completed >= target") is a
 * prefix property of this sequence; fb_h2h_resolve applies it.
 * workspace_dev >= fb_workspace_bytes(2, total_attempts)
 *                  + align256((n_blocks + 1) * 8) + align256(n_blocks * 16) bytes.
 */
int fb_play_h2h(uint64_t root_seed, int n_blocks, const uint64_t* pair_id_dev,
                const uint8_t* order_dev, const fb_strategy_t* seat1_dev,
                const fb_strategy_t* seat2_dev, const uint32_t* attempt0_dev,
                const uint32_t* n_attempts_dev, uint64_t total_attempts, int32_t target_score,
                int32_t max_rounds, uint8_t* outcome_out_dev, void* rows_dev,
                int64_t* totals_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Apply the early-stop rule to the outcomes of fb_play_h2h.  progress[b] =
 * {games_attempted, games_completed, games_safety_limit, wins_seat1,
 * wins_seat2} is read as the state before these attempts and updated in place
 * exactly as the reference loop would after consuming the attempt prefix.   */
int fb_h2h_resolve(int n_blocks, const uint32_t* n_attempts_dev, const uint8_t* outcome_dev,
                   const int32_t* n_completed_required_dev, int32_t* progress_dev /*[n][5]*/,
                   void* stream);

/* Games at explicit coordinates.  Replaces _play_game
 * (src/farkle/simulation/simulation.py:576-655) and the loops of
 * simulate_many_games (:658-722): game i seats seat_strategies[i][0..k-1]
 * and draws from streams coords[i] = {purpose, root_seed, k, shuffle_index,
 * pair_id, order, game_index} + seat_index.                                 */
int fb_play_games(const uint64_t* coords_dev /*[n][7]*/, uint64_t n_games, int k,
                  const fb_strategy_t* seat_strategies_dev /*[n][k]*/,
                  const int32_t* seat_strategy_ids_dev /*[n][k], NULL = 0..k-1*/,
                  const int32_t* target_score_dev /*[n] or NULL*/, int32_t target_score,
                  const int32_t* max_rounds_dev /*[n] or NULL*/, int32_t max_rounds,
                  void* rows_dev, int64_t* totals_dev, void* workspace_dev,
                  size_t workspace_bytes, void* stream);

/* Host-buffer convenience for the reference-facing plug-in: H2D strategy
 * table, fb_play_tournament, D2H tallies/totals(/rows), synchronised before
 * returning.  All pointers are HOST pointers; rows_host may be NULL.
 * tallies_host/totals_host are overwritten (not accumulated).  With rows_host
 * (pinned memory recommended) a large range is played in up to four chunks of
 * whole tally slots and the rows of one chunk are copied to the host on a
 * second stream while the next chunk is being played; game_ordinal still runs
 * over the whole call.                                                       */
int fb_run_tournament_host(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                           const fb_strategy_t* strategies_host,
                           const int32_t* strategy_ids_host, int n_strategies, int n_tally_ids,
                           int32_t target_score, int32_t max_rounds, int shuffles_per_slot,
                           int64_t* tallies_host, int64_t* totals_host, void* rows_host,
                           int want_game_seeds);

/* Timing hook: milliseconds the play kernel of the most recent hot-path call on
 * this thread took, measured with CUDA events on the launching stream.  Blocks
 * until that kernel has finished.  Returns < 0 if nothing was launched.      */
float fb_last_play_kernel_ms(void);
/* The same for the most recent launches of this thread, newest first: fills out_ms with up
 * to max_entries durations (the library keeps the last 64) and returns how many were
 * written (negative fb_status on error).  Blocks until those kernels have finished.      */
int fb_play_kernel_ms_history(float* out_ms, int max_entries);
/* Timeline hook.  fb_timeline(1) clears the log and starts recording a CUDA timing event on the
 * launching stream behind every kernel of the tournament path (permute, seed, play_kernel, finish,
 * gather); fb_timeline(0) stops and clears.  fb_timeline_dump waits for the recorded events and
 * writes one line per mark, "<lane> <name> <ms since the first mark>\n" (lane 0 = the caller's
 * stream, 1 = the preparation stream of fb_play_tournament_cells), returning the number of lines
 * written (negative fb_status on error).  Off by default: no events, no cost.  Process-wide.   */
int fb_timeline(int enable);
int fb_timeline_dump(char* out, size_t capacity);
/* Roofline probe: register-only kernels of independent 32-bit integer chains on every SM
 * (1,024 threads per SM, no memory), returning measured lane-instructions per second.
 * Synchronous.  fb_measure_issue_peak reports the best of the mixed-pipe variants
 * (half IMAD on the FMA-heavy pipe, half LOP3/IADD3 on the ALU pipe): the issue peak the
 * path is held against (MEASURED_PEAKS.json has no integer peak).
 * fb_measure_issue_peak_variant runs one variant: 0 chain-major mad/xor/mad/add (each
 * instruction depends on its predecessor), 1 op-major over 8 chains, 2 op-major over 16
 * chains with the pipes interleaved, 3 xor/add chains (LOP3 + IMAD.IADD alternating: the
 * one that reaches ~0.98 instructions per cycle and scheduler), 4 mad only (one pipe).   */
int fb_measure_issue_peak(int iters, double* lane_ops_per_second);
int fb_measure_issue_peak_variant(int variant, int iters, double* lane_ops_per_second);
/* Number of kernels this library has launched since fb_init (all threads).  */
uint64_t fb_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FARKLE_B200_H */
