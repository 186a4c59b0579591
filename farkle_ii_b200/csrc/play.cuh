// play.cuh — the persistent one-game-per-lane Farkle state machine + the dense finish pass.
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
// Data layout in HBM: one 80-byte Seat record per (game, seat), written by the seed kernels
//   line 0  PCG state (16 B)
//   line 1  face queue lo | score | highest_turn|queue length in bits<<24|queue holds a rejected
//           half<<30|has_scored<<31 | farkles|rolls<<16
//   line 2  turns|hot<<16 | smart-five uses|dice<<16 | smart-one uses|dice<<16 | face queue hi
//   line 3  PCG increment (16 B)
//   line 4  st_d | kf|dt_d<<16 | dbase|tab_off<<16 | strategy table index   (seat_consts())
//   game header 4 B  n_rounds | flags << 16 | HDR_LONG, written when the game ends; the seed
//                    kernels pre-set HDR_LONG on games that are certain to run to the safety
//                    limit so that play_kernel starts them first (longest-first scheduling)
// The active seat lives in registers.  Generic k: a turn switch stores lines 0-2 of the outgoing
// seat (three 16-byte st.cg) and takes the incoming seat from this lane's shared-memory staging
// slots, which a cp.async prefetch filled during the turn that just ended; the records of the
// games in flight stay L2 resident.  Two seats (K2): each seat has a home slot in the lane's
// shared memory for the whole game (a switch stores three mutable lines to one slot and loads
// five lines from the other); global memory is touched when a game starts and when it ends.
// The generator state of line 0 runs AHEAD of the reference's by the halves already turned into
// the face codes of the seat's queue (lines 1-2): only that seat's rolls read the stream, in
// order, so every die is the reference's die (the G block of play_kernel).
//
// play_kernel   one game per lane, loop body = ONE ROLL, straight-line; persistent CTAs
//               (one per SM); a lane whose game ended takes the next game from its warp's
//               queue (lane refill), so the 100x spread of game lengths does not idle the
//               warp.  It only plays: at game end it writes the header.
// finish_kernel one thread per finished game, dense and lane-parallel: ranks the seats,
//               marks the winner in the header, emits the compact row (every byte of a row
//               is written by one thread, L2 merges the sectors) and the launch totals.
// tally_gather_kernel
//               the per-strategy tallies: exposures and winner metrics by gather over the
//               inverse permutation (few atomics).
#pragma once
#include <climits>
#include <cstdint>

#include "rng.cuh"
#include "scoring.cuh"

namespace fb {

constexpr int ROLL_LIMIT = 1000;  // src/farkle/game/engine.py:36
constexpr uint32_t HIGH_MASK = 0x00ffffffu;  // highest_turn (< 3,000 points x 1,000 rolls)
constexpr uint32_t HW_REJ = 1u << 30;  // the seat's face queue holds a rejected half (code 6)
// Lemire threshold of Generator.integers(1, 7): a half whose low product word is below it is
// re-drawn; NumPy's value is (2^32 - 6) % 6 = 4.  Test builds raise it (scripts/build_variant.py,
// -DFB_LEMIRE_THR=0x40000000u: one half in four rejected) so that the rejection paths of the face
// queue run all the time; the oracle has the matching run-time knob FB_TEST_LEMIRE_THR.
#ifndef FB_LEMIRE_THR
#define FB_LEMIRE_THR 4u
#endif
// Face queue (the dice of a seat's stream that are already drawn): 3-bit codes, next die in the low
// bits, 0..5 a face, 6 a half NumPy's Lemire test rejects (it is skipped, but it was read).
// The queue length is kept in BITS (3 per code): it is the shift count of an insert as it stands.
#ifndef FB_FQ_WORDS
#define FB_FQ_WORDS 3
#endif
constexpr int FQ_WORDS = FB_FQ_WORDS;            // 64-bit outputs per top-up
constexpr int FQ_CAP_BITS = 63;                  // 21 codes in one 64-bit word
constexpr int FQ_GEN_BITS = 6 * FQ_WORDS;        // two codes per output
constexpr uint32_t HW_LEN_SHIFT = 24;            // the record keeps the length in bits 24..29 of line 1 word 2
constexpr uint32_t HW_LEN_MASK = 63u << HW_LEN_SHIFT;
constexpr int FQ_IDLE = 63;  // queue length of a lane that is not playing: never short, never topped up
static_assert(FQ_WORDS == 3, "one top-up must cover the longest roll (6 codes) for every lane it serves");
constexpr uint32_t HW_SCORED = 1u << 31;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t HDR_LONG = 1u << 31;

struct __align__(16) Seat {
    uint4 state;  // PCG state lo0, lo1, hi0, hi1
    uint4 a;      // face queue lo | score | hw | farkles|rolls<<16
    uint4 b;      // turns|hot<<16 | sf uses|dice<<16 | so uses|dice<<16 | face queue hi
    uint4 inc;    // PCG increment
    uint4 cst;    // st_d | kf|dt_d<<16 | dbase|tab_off<<16 | strategy table index
};
static_assert(sizeof(Seat) == 80, "seat record layout");
constexpr uint32_t CST_DT_SHIFT = 16;  // cst.y: kf in the low half, dt_d (signed) in the high half

struct PlayParams {
    Seat* seats;            // [n_games*k]
    uint32_t* header;       // [n_games]
    const int32_t* limits;  // [n_games][2] {target, max_rounds} or nullptr
    int32_t target_score, max_rounds;
    int32_t roll_limit;  // ROLL_LIMIT, or the FB_TEST_ROLL_LIMIT test knob (launch_play)
    uint32_t n_games;
    int k;
    unsigned long long* totals;  // [FB_TOTALS_WIDTH] or nullptr (dice / rng words only)
    unsigned int* counter;       // [0] next ordinal (zeroed per launch), [1] number of HDR_LONG games
    const uint32_t* long_list;   // [counter[1]] ordinals of the HDR_LONG games
};

enum LaneStatus { ST_NEED = 0, ST_PLAY = 1, ST_DEAD = 2 };

// 16-byte asynchronous global -> shared copy (LDGSTS) and its completion wait.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(smem_addr)
                 : "memory");
    return v;
}
// Staging area behind the tables: line f of lane t at stage0 + f * STAGE_STRIDE + t * 16, so a
// warp's 16-byte accesses to one line are contiguous (conflict free) and the five lines of a
// lane are immediate offsets from one address register.
constexpr int PLAY_THREADS = 1024;
constexpr uint32_t STAGE_STRIDE = PLAY_THREADS * 16u;
constexpr uint32_t QUEUE_CAP = 64;  // games per warp queue (a top-up adds <= 32 to < 32)
// Generic kernel: 5 staging lines per lane (the prefetched record of the seat that plays next).
// K2 kernel: 10 lines per lane, a HOME SLOT of five lines for each seat -- lines 5*s .. 5*s+2 the
// mutable part of seat s, lines 5*s+3, 5*s+4 its constants (increment, strategy constants) -- so a
// two-seat game lives in registers + shared memory from its first roll to its last, and a turn
// switch stores the outgoing seat to ITS slot before loading the incoming one from the other slot:
// no swap through temporaries, the loads land in the registers the stores just freed.  (The compact
// score table, scoring.cuh, is what makes 160 KB of slots fit beside the tables.)
__host__ __device__ constexpr uint32_t stage_lines(bool k2) { return k2 ? 10u : 5u; }
__host__ __device__ constexpr size_t play_queue_offset(bool k2) { return (size_t)((LUT_BYTES + 15) & ~15) + stage_lines(k2) * STAGE_STRIDE; }
__host__ __device__ constexpr size_t play_smem_bytes(bool k2) { return play_queue_offset(k2) + (PLAY_THREADS / 32) * QUEUE_CAP * 4u; }
static_assert(play_smem_bytes(true) <= 227u * 1024u, "K2 staging must fit the 227 KB of one CTA");
__device__ __forceinline__ void sts128(uint32_t smem_addr, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
// u * 6 as its two 32-bit words (one IMAD.WIDE)
__device__ __forceinline__ void mul_wide6(uint32_t u, uint32_t& lo, uint32_t& hi) {
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, 6; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(u));
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// K2: two-seat games (every H2H block and the k=2 tournament cells).  The other seat always plays
// next, so the turn switch needs no seat-order logic and both seats fit the lane's registers plus
// seven shared-memory lines (stage_lines).
template <bool LIMITS, bool K2>
__global__ void __launch_bounds__(PLAY_THREADS, 1) play_kernel(const PlayParams P, const ScoreLut* __restrict__ lut_g) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ unsigned long long s_tot[2];
    ScoreLut* lut = reinterpret_cast<ScoreLut*>(smem_raw);
    uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    asm volatile("" : "+r"(lut_s));  // opaque: keep it in a register instead of recomputing it
    const uint32_t stage = lut_s + (uint32_t)((LUT_BYTES + 15) & ~15) + threadIdx.x * 16u;
    for (int i = threadIdx.x; i < LUT_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(lut)[i] = reinterpret_cast<const uint4*>(lut_g)[i];
    if (threadIdx.x < 2) s_tot[threadIdx.x] = 0ull;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int k = K2 ? 2 : P.k;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // Ordinals [0, n_long) walk the long list, [n_long, n_long + n_games) walk every game and
    // skip the ones the long list already covered.
    const uint32_t n_long = P.counter[1];
    const uint32_t n_ordinals = P.n_games + n_long;

    // per-warp game queue (warp-uniform registers; entries = game | HDR_LONG)
    const uint32_t gq = lut_s + (uint32_t)play_queue_offset(K2) + (threadIdx.x >> 5) * (QUEUE_CAP * 4u);
    uint32_t q_head = 0, q_tail = 0, q_seen = 0;  // q_seen: the global counter as last observed
    const uint32_t q_share = 2u * gridDim.x * (blockDim.x >> 5);  // twice the number of warps
    bool q_dry = false;

    // ---- per-lane game state ------------------------------------------------
    int status = ST_NEED;
    uint32_t g = 0, err = 0;
    int seat = 0, nseat = 0, round = 0, trigger = -1, stb = 0;
    int target = P.target_score, max_rounds = P.max_rounds;
    // active seat in registers
    Pcg rng{0, 0, 0, 0};
    uint32_t hw = 0;
    uint64_t fq = 0;  // face queue of the active seat
    int fq_bits = FQ_IDLE;  // 3 x the number of queued codes (FQ_IDLE while the lane has no game)
    int score = 0, st_d = 0, dt_d = 0;
    uint32_t c_fr = 0, c_th = 0, c_sf = 0, c_so = 0, kf = 0, dbase = 0, tab_off = 0;
    // active turn
    int ts = 0, dice = 6, rolls_turn = 0;
    uint32_t a_dice = 0, a_words = 0;

    // Start the turn of player `seat` from its record lines (engine.py:229-240), then predict
    // who plays next if this turn neither triggers the final round nor ends the game — seat
    // order within a round, or the final-round order that skips the trigger seat
    // (engine.py:459-471,533-548); >= k means nobody — and prefetch that record into the
    // staging slots so its L2 latency overlaps this whole turn instead of stalling the warp.
    // (The predicted record was last written by this lane, earlier in program order.)
    auto start_turn = [&](const uint4 m0, const uint4 m1, const uint4 m2, const uint4 i0, const uint4 i1) {
        rng.lo = (uint64_t)m0.x | ((uint64_t)m0.y << 32);
        rng.hi = (uint64_t)m0.z | ((uint64_t)m0.w << 32);
        rng.ilo = (uint64_t)i0.x | ((uint64_t)i0.y << 32);
        rng.ihi = (uint64_t)i0.z | ((uint64_t)i0.w << 32);
        fq = (uint64_t)m1.x | ((uint64_t)m2.w << 32);
        fq_bits = (int)((m1.z & HW_LEN_MASK) >> HW_LEN_SHIFT);
        score = (int)m1.y;
        hw = m1.z & ~HW_LEN_MASK;  // (the field stays clear in the register: the store ORs the length in)
        c_fr = m1.w;
        c_th = m2.x + 1u;  // n_turns += 1 (engine.py:237)
        c_sf = m2.y;
        c_so = m2.z;
        st_d = (int)i1.x;
        kf = i1.y;  // KF_* bits live in the low half; the high half is never tested
        dt_d = (int)i1.y >> CST_DT_SHIFT;
        dbase = i1.z & 0xffffu;
        tab_off = i1.z >> 16;
        dice = 6;
        ts = 0;
        rolls_turn = 0;
        if (!K2) {
            int ns = seat + 1;
            if (trigger < 0) ns = ns == k ? 0 : ns;
            else if (ns == trigger) ns++;
            nseat = ns;
            // Past the last seat of the final round nobody plays next: fetch the last record
            // anyway (valid memory, never consumed) rather than branch around the copies.
            const uint4* np = reinterpret_cast<const uint4*>(P.seats + (g * (uint32_t)k + (uint32_t)min(ns, k - 1)));
            cp_async16(stage, np);
            cp_async16(stage + STAGE_STRIDE, np + 1);
            cp_async16(stage + 2u * STAGE_STRIDE, np + 2);
            cp_async16(stage + 3u * STAGE_STRIDE, np + 3);
            cp_async16(stage + 4u * STAGE_STRIDE, np + 4);
        }
    };
    auto start_turn_from_l2 = [&]() {
        const uint4* sp = reinterpret_cast<const uint4*>(P.seats + (g * (uint32_t)k + (uint32_t)seat));
        if (K2) {
            // A two-seat game starts: seat 0 goes to registers, seat 1 and the constants of seat 0 go
            // to the home slots; after that the game touches global memory again only when it ends.
            const uint4 m0 = __ldcg(sp), m1 = __ldcg(sp + 1), m2 = __ldcg(sp + 2), i0 = __ldcg(sp + 3), i1 = __ldcg(sp + 4);
            const uint4 q0 = __ldcg(sp + 5), q1 = __ldcg(sp + 6), q2 = __ldcg(sp + 7), j0 = __ldcg(sp + 8), j1 = __ldcg(sp + 9);
            sts128(stage + 3u * STAGE_STRIDE, i0);
            sts128(stage + 4u * STAGE_STRIDE, i1);
            sts128(stage + 5u * STAGE_STRIDE, q0);
            sts128(stage + 6u * STAGE_STRIDE, q1);
            sts128(stage + 7u * STAGE_STRIDE, q2);
            sts128(stage + 8u * STAGE_STRIDE, j0);
            sts128(stage + 9u * STAGE_STRIDE, j1);
            start_turn(m0, m1, m2, i0, i1);
            return;
        }
        // An unconsumed prefetch of this lane (previous game, or a seat order the prediction
        // missed) must have landed before the staging slots are targeted again: cp.async
        // copies of one thread are not ordered among themselves.
        cp_async_wait_all();
        const uint4 m0 = __ldcg(sp), m1 = __ldcg(sp + 1), m2 = __ldcg(sp + 2), i0 = __ldcg(sp + 3), i1 = __ldcg(sp + 4);
        start_turn(m0, m1, m2, i0, i1);
    };
    auto start_turn_staged = [&]() {
        cp_async_wait_all();
        const uint4 m0 = lds128(stage), m1 = lds128(stage + STAGE_STRIDE);
        const uint4 m2 = lds128(stage + 2u * STAGE_STRIDE), i0 = lds128(stage + 3u * STAGE_STRIDE);
        const uint4 i1 = lds128(stage + 4u * STAGE_STRIDE);
        start_turn(m0, m1, m2, i0, i1);
    };

    // One loop iteration; returns true when every lane of the warp is out of work.
    auto roll_step = [&]() -> bool {
        // ================= R: lane refill =====================================
        // Games come from a per-warp queue in shared memory.  When the queue cannot serve the
        // lanes that need a game, the warp reserves 32 ordinals with ONE atomic, every lane
        // resolves its ordinal to a game (long list first, then all games minus the ones the
        // list covered), prefetches that game's first seat record into L2 and enqueues it.  The
        // atomic, the ordinal->game load and the first-touch DRAM latency of the record are
        // thus paid one top-up ahead of use instead of stalling the warp at every game start.
        const uint32_t need = __ballot_sync(FULL, status == ST_NEED);
        if (need) {
            const uint32_t want = (uint32_t)__popc(need);
            if (q_tail - q_head < want && !q_dry) {
                // guided chunk size: 32 while plenty is left, shrinking to 1 so that no warp is
                // left holding a queue of games while the others have run dry
                const uint32_t left = n_ordinals - min(n_ordinals, q_seen);
                const uint32_t chunk = min(32u, max(max(want, 1u), left / q_share));
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(P.counter, chunk);
                base = __shfl_sync(FULL, base, 0);
                q_seen = base + chunk;
                const uint32_t ng = base + (uint32_t)lane;
                bool ok = (uint32_t)lane < chunk && ng < n_ordinals;
                uint32_t gq_entry = 0;
                if (ok) {
                    if (ng < n_long) {
                        gq_entry = P.long_list[ng] | HDR_LONG;
                    } else {
                        gq_entry = ng - n_long;
                        ok = !(__ldcg(&P.header[gq_entry]) & HDR_LONG);  // already played from the list
                    }
                }
                if (ok) {
                    const char* rp = reinterpret_cast<const char*>(P.seats + (size_t)(gq_entry & ~HDR_LONG) * k);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 64));
                }
                const uint32_t m = __ballot_sync(FULL, ok);
                if (ok) sts_u32(gq + ((q_tail + (uint32_t)__popc(m & lt_mask)) & (QUEUE_CAP - 1u)) * 4u, gq_entry);
                q_tail += (uint32_t)__popc(m);
                q_dry = base + chunk >= n_ordinals;
                __syncwarp();
            }
            const uint32_t have = min(want, q_tail - q_head);
            if (status == ST_NEED) {
                const uint32_t rank = (uint32_t)__popc(need & lt_mask);
                if (rank < have) {
                    const uint32_t entry = lds_u32(gq + ((q_head + rank) & (QUEUE_CAP - 1u)) * 4u);
                    g = entry & ~HDR_LONG;
                    err = entry & HDR_LONG;  // carried into the header (bit 31), not a row flag
                    seat = 0;
                    round = 1;
                    trigger = -1;
                    if (LIMITS) {
                        target = P.limits[2 * (size_t)g];
                        max_rounds = P.limits[2 * (size_t)g + 1];
                    }
                    stb = target;
                    if (max_rounds <= 0) {  // `while rounds < max_rounds` never runs (engine.py:455)
                        P.header[g] = ((uint32_t)FB_ROW_SAFETY_LIMIT << 16) | err;
                    } else {
                        start_turn_from_l2();
                        status = ST_PLAY;
                    }
                } else if (q_dry) {
                    status = ST_DEAD;  // (have == q_tail - q_head here: the queue is empty too)
                }
            }
            q_head += have;
            // lanes only die here, so this is the one place the warp can run out of work
            if (__all_sync(FULL, status == ST_DEAD)) return true;
        }

        // ================= G: top up the face queues ===========================
        // A seat's dice are a pure function of its own stream of 32-bit halves, read in order
        // (engine.py:85-101: Generator.integers(1, 7, size=n) draws n Lemire-bounded values from
        // the buffered halves, re-drawing a half whose low product word is < 4), and nothing else
        // reads that stream.  So the halves can be turned into faces AHEAD of the rolls that use
        // them: every seat keeps a queue of 3-bit face codes, and the warp tops its queues up by
        // three 64-bit outputs (six codes) whenever some lane's queue is shorter than its next roll.
        // That is 2.2 outputs computed per roll instead of three (2.13 are consumed), no per-word
        // state selects, and no per-die window arithmetic: a roll takes its n codes off the queue
        // with two shifts.
        // A top-up adds six codes, the longest roll, to every lane it serves (the lanes with room
        // for them), and a lane it does not serve holds more than that: one top-up per iteration
        // at most.  Lanes without a game carry the length FQ_IDLE and are neither short nor served.
        const int need_bits = 3 * dice;
        if (__any_sync(FULL, fq_bits < need_bits)) {
            if (fq_bits <= FQ_CAP_BITS - FQ_GEN_BITS) {
                uint64_t shi = rng.hi, slo = rng.lo;
                // Lemire per half: face = high word of half * 6, rejected when the low word is < 4
                // (4 in 2^32); codes packed two per output, first-read half lowest
                uint32_t packed = 0, minl = 0xffffffffu;
#pragma unroll
                for (int w = 0; w < FQ_WORDS; w++) {
                    const uint64_t o = pcg_output(shi, slo);
                    pcg_step(shi, slo, rng.ihi, rng.ilo);
                    uint32_t l0, f0, l1, f1;
                    mul_wide6((uint32_t)o, l0, f0);
                    mul_wide6((uint32_t)(o >> 32), l1, f1);
                    minl = min(minl, min(l0, l1));
                    packed += (f0 + (f1 << 3)) << (6 * w);
                }
                if (minl < FB_LEMIRE_THR) {  // redo the batch half by half, code 6 for a rejected half
                    shi = rng.hi;
                    slo = rng.lo;
                    packed = 0;
                    for (int w = 0; w < FQ_WORDS; w++) {
                        const uint64_t o = pcg_output(shi, slo);
                        pcg_step(shi, slo, rng.ihi, rng.ilo);
                        const uint32_t u0 = (uint32_t)o, u1 = (uint32_t)(o >> 32);
                        const uint32_t c0 = u0 * 6u < FB_LEMIRE_THR ? 6u : __umulhi(u0, 6u);
                        const uint32_t c1 = u1 * 6u < FB_LEMIRE_THR ? 6u : __umulhi(u1, 6u);
                        packed |= (c0 | (c1 << 3)) << (6 * w);
                    }
                    hw |= HW_REJ;
                }
                rng.hi = shi;
                rng.lo = slo;
                fq |= (uint64_t)packed << (uint32_t)fq_bits;
                fq_bits += FQ_GEN_BITS;
                a_words += (uint32_t)FQ_WORDS;
            }
        }

        // ================= P: one roll (straight-line, no divergent branches) =====
        if (status == ST_PLAY) {
            const int n = dice;
            uint32_t hist;
            if (!(hw & HW_REJ)) {
                // the next n codes; the slots beyond them read as code 7, which counts nothing
                const uint32_t bits = (uint32_t)fq | (0xffffffffu << (uint32_t)need_bits);
                hist = lds_u32(lut_s + LUT_OFF_HIST3 + 4u * (bits & 511u)) +
                       lds_u32(lut_s + LUT_OFF_HIST3 + 4u * ((bits >> 9) & 511u));
                fq >>= (uint32_t)need_bits;
                fq_bits -= need_bits;
            } else {
                // A rejected half somewhere in the queue: take the dice one code at a time,
                // skipping it (and drawing more when the queue runs dry).
                hist = 0;
                for (int taken = 0; taken < n;) {
                    if (fq_bits == 0) {
                        const uint64_t o = pcg_output(rng.hi, rng.lo);
                        pcg_step(rng.hi, rng.lo, rng.ihi, rng.ilo);
                        const uint32_t u0 = (uint32_t)o, u1 = (uint32_t)(o >> 32);
                        const uint32_t c0 = u0 * 6u < FB_LEMIRE_THR ? 6u : __umulhi(u0, 6u);
                        const uint32_t c1 = u1 * 6u < FB_LEMIRE_THR ? 6u : __umulhi(u1, 6u);
                        fq = c0 | (c1 << 3);
                        fq_bits = 6;
                        a_words += 1u;
                    }
                    const uint32_t code = (uint32_t)fq & 7u;
                    fq >>= 3;
                    fq_bits -= 3;
                    if (code < 6u) {
                        hist += 1u << (3u * code);
                        taken++;
                    }
                }
                bool rej = false;
                uint64_t t = fq;
                for (int i = 0; i < fq_bits; i += 3, t >>= 3) rej = rej || ((uint32_t)t & 7u) == 6u;
                hw = rej ? hw : (hw & ~HW_REJ);
            }
            a_dice += (uint32_t)n;

            // -- score the roll (engine.py:103-147), discards, counters
            const uint32_t e = lut_lookup_s(lut_s, tab_off, hist);
            const int rscore = (int)(e & 127u) * 50;
            const int used0 = (int)((e >> 7) & 7u);
            const bool farkle = rscore == 0;
            const uint32_t dd = smart_discards_s(lut_s, dbase, e, n, ts, st_d, dt_d);
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
            const bool hot = !farkle && ndice == 6 && (kf & KF_AUTO_HOT);
            c_th += hot ? 0x10000u : 0u;
            const bool fin = trigger >= 0;
            const int rt = score + ts2;
            const bool behind = fin && rt <= stb;
            const bool stop_ahead = fin && rt > stb && !(kf & KF_RUN_UP);
            const bool gate = !(hw & HW_SCORED) && ts2 < 500;
            const bool keep = !stop_ahead && (gate || behind || decide_continue(ts2, ndice, st_d, dt_d, kf));
            bool turn_over = farkle || (!hot && !keep);
            ts = farkle ? 0 : ts2;
            dice = ndice;
            if (!turn_over && rolls_turn >= P.roll_limit) {  // engine.py:242-243 raises
                err |= FB_ROW_ROLL_LIMIT;
                turn_over = true;
            }

            // ================= T: bank, park the seat, seat the next one ========
            if (turn_over) {
                hw |= ts >= 500 ? HW_SCORED : 0u;  // entry turn (engine.py:266-267)
                const int banked = (hw & HW_SCORED) ? ts : 0;  // select, not a branch: ts >= 0
                score += banked;
                hw = (hw & ~HIGH_MASK) | max(hw & HIGH_MASK, (uint32_t)banked);
                const uint4 l0 = make_uint4((uint32_t)rng.lo, (uint32_t)(rng.lo >> 32), (uint32_t)rng.hi,
                                            (uint32_t)(rng.hi >> 32));
                const uint4 l1 = make_uint4((uint32_t)fq, (uint32_t)score, hw | ((uint32_t)fq_bits << HW_LEN_SHIFT), c_fr);
                const uint4 l2 = make_uint4(c_th, c_sf, c_so, (uint32_t)(fq >> 32));
                uint4* sp = reinterpret_cast<uint4*>(P.seats + (g * (uint32_t)k + (uint32_t)seat));
                if (!K2) {
                    __stcg(sp, l0);
                    __stcg(sp + 1, l1);
                    __stcg(sp + 2, l2);
                }
                if (K2) {
                    // Branch free.  fin: the seat that did not trigger has had its final turn;
                    // else this turn may trigger the final round (engine.py:466-471); else seat 1
                    // closes the round, and the game unless the safety limit allows another one.
                    const bool trig_now = !fin && score >= target;
                    const bool closes = !fin && !trig_now && seat == 1;
                    const bool capped = closes && round >= max_rounds;
                    trigger = trig_now ? seat : trigger;
                    stb = trig_now ? score : stb;
                    round += (closes && !capped) ? 1 : 0;
                    const bool over = fin || capped;
                    const uint32_t home = stage + 5u * (uint32_t)seat * STAGE_STRIDE;          // this seat's slot
                    const uint32_t other = stage + 5u * (uint32_t)(seat ^ 1) * STAGE_STRIDE;  // the other seat's
                    if (over || (err & FB_ROW_ROLL_LIMIT)) {
                        // the game ends: the counters of both seats (lines 1-2; nobody reads a
                        // finished game's generator state) go back to global memory for the finish pass
                        const uint4 p1 = lds128(other + STAGE_STRIDE);
                        const uint4 p2 = lds128(other + 2u * STAGE_STRIDE);
                        __stcg(sp + 1, l1);
                        __stcg(sp + 2, l2);
                        uint4* op = reinterpret_cast<uint4*>(P.seats + (g * 2u + (uint32_t)(seat ^ 1)));
                        __stcg(op + 1, p1);
                        __stcg(op + 2, p2);
                        P.header[g] = (uint32_t)round | (err & HDR_LONG) |
                                      (((trigger < 0 ? FB_ROW_SAFETY_LIMIT : 0u) | (err & 0xffu)) << 16);
                        status = ST_NEED;
                        fq_bits = FQ_IDLE;
                    } else {
                        // park this seat in its home slot, then seat the other one from its own
                        sts128(home, l0);
                        sts128(home + STAGE_STRIDE, l1);
                        sts128(home + 2u * STAGE_STRIDE, l2);
                        seat ^= 1;
                        start_turn(lds128(other), lds128(other + STAGE_STRIDE), lds128(other + 2u * STAGE_STRIDE),
                                   lds128(other + 3u * STAGE_STRIDE), lds128(other + 4u * STAGE_STRIDE));
                    }
                } else {
                    // Who plays next.  Without a trigger event it is the seat predicted (and
                    // prefetched) at the start of this turn.
                    // Branch free: final round (engine.py:533-548) | this turn triggers it
                    // (engine.py:466-471) | next seat, or next round unless the safety limit is
                    // reached (engine.py:455).
                    const bool trig_now = !fin && score >= target;
                    const bool plain = !fin && !trig_now;
                    stb = fin ? max(stb, score) : (trig_now ? score : stb);
                    trigger = trig_now ? seat : trigger;
                    const int next = trig_now ? (seat == 0 ? 1 : 0) : nseat;
                    const bool closes = plain && next == 0;
                    const bool capped = closes && round >= max_rounds;
                    round += (closes && !capped) ? 1 : 0;
                    const bool over = capped || (!plain && next >= k);
                    if (over || (err & FB_ROW_ROLL_LIMIT)) {
                        P.header[g] = (uint32_t)round | (err & HDR_LONG) |
                                      (((trigger < 0 ? FB_ROW_SAFETY_LIMIT : 0u) | (err & 0xffu)) << 16);
                        status = ST_NEED;
                        fq_bits = FQ_IDLE;
                    } else {
                        // A trigger event (or k == 1) makes the prediction miss: restage the right
                        // record (the outstanding copies into the same slots must land first) and
                        // take the common path; rare, so its exposed L2 latency does not matter.
                        if (next != nseat || k == 1) {
                            cp_async_wait_all();
                            const uint4* rp = reinterpret_cast<const uint4*>(P.seats + (g * (uint32_t)k + (uint32_t)next));
                            cp_async16(stage, rp);
                            cp_async16(stage + STAGE_STRIDE, rp + 1);
                            cp_async16(stage + 2u * STAGE_STRIDE, rp + 2);
                            cp_async16(stage + 3u * STAGE_STRIDE, rp + 3);
                            cp_async16(stage + 4u * STAGE_STRIDE, rp + 4);
                        }
                        seat = next;
                        start_turn_staged();
                    }
                }
            }
        }
        return false;
    };
    // Two copies of the body per trip: the loop-carried seat state then alternates between two
    // register sets instead of being moved back at the end of every iteration.
    for (;;) {
        if (roll_step()) break;
        if (roll_step()) break;
    }

    cp_async_wait_all();  // no prefetch may still be in flight when the CTA's shared memory is released

    // ---- work counters: warp shuffle -> shared memory -> one global RED per CTA ----
    unsigned long long v[2] = {a_dice, a_words};
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(FULL, v[i], o);
        if (lane == 0 && v[i]) atomicAdd(&s_tot[i], v[i]);
    }
    __syncthreads();
    if (P.totals && threadIdx.x < 2 && s_tot[threadIdx.x])
        atomicAdd(&P.totals[4 + threadIdx.x], s_tot[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// finish pass
// ---------------------------------------------------------------------------
struct FinishParams {
    const Seat* seats;
    uint32_t* header;
    const int32_t* strategy_ids;  // id of table entry (ids_mode 1)
    const int32_t* perm;          // table index of record r (tournaments: the permutations); nullptr = r
    int ids_mode;                 // 0 id = table index, 1 id = strategy_ids[index], 2 id = seat
    const uint64_t* game_seed;    // [n_games] or nullptr
    uint32_t n_games;
    int k;
    uint32_t games_per_slot;  // 0 = one tally slot
    int n_tally_ids;
    int mark_winner;              // 1: write winner seat + 1 into header bits 24..27 for the gather pass
    unsigned long long* tallies;  // [slots][ids][26] or nullptr (filled by tally_gather_kernel)
    unsigned long long* totals;   // [FB_TOTALS_WIDTH] or nullptr
    uint32_t* rows;               // or nullptr
    uint32_t ordinal_base;        // game_ordinal of the launch's first game (chunked host calls)
    int row_words;
    uint8_t* outcome;  // [n_games] or nullptr
};

// Rows mode stages the CTA's 256 rows in dynamic shared memory and writes them out as one
// contiguous run of 16-byte stores (coalesced); dynamic shared memory = 256 * row bytes then.
__global__ void __launch_bounds__(256) finish_kernel(const FinishParams F) {
    extern __shared__ __align__(16) uint32_t row_tile[];
    __shared__ unsigned long long s_tot[FB_TOTALS_WIDTH];
    if (threadIdx.x < FB_TOTALS_WIDTH) s_tot[threadIdx.x] = 0ull;
    __syncthreads();
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = F.k;
    unsigned long long t_rolls = 0, t_turns = 0;
    uint32_t t_done = 0, t_safe = 0, t_err = 0, t_ahead = 0;
    int winner = -1;
    if (g < F.n_games) {
        const uint32_t hdr = F.header[g];
        const uint32_t rounds = hdr & 0xffffu;
        uint32_t flags = (hdr >> 16) & 0xffu;
        const bool safety = flags & FB_ROW_SAFETY_LIMIT;
        if (rounds > (uint32_t)FB_MAX_ROUNDS) flags |= FB_ROW_I16_OVERFLOW;  // n_rounds is an int16 column
        const Seat* seats = F.seats + (size_t)g * k;
        // pass 1: winner = highest score, ties to the lower seat (stable sort, engine.py:483)
        int best = -1;
        for (int s = 0; s < k; s++) {
            const uint4 a = __ldcg(&seats[s].a);
            const uint4 b = __ldcg(&seats[s].b);
            const int sc = (int)a.y;
            if (sc > best) {
                best = sc;
                winner = s;
            }
            t_rolls += a.w >> 16;
            t_turns += b.x & 0xffffu;
            t_ahead += ((a.z & HW_LEN_MASK) >> HW_LEN_SHIFT) / 6u;  // whole outputs (two codes) drawn ahead, never read
            if ((a.w >> 16) > 32767u || (a.z & HIGH_MASK) > 32767u || (b.y >> 16) > 32767u ||
                (b.z >> 16) > 32767u || (b.x & 0xffffu) > 32767u || (b.x >> 16) > 32767u)
                flags |= FB_ROW_I16_OVERFLOW;
        }
        if (safety) winner = -1;
        t_done = 1;
        t_safe = safety ? 1u : 0u;
        t_err = (flags & (FB_ROW_ROLL_LIMIT | FB_ROW_I16_OVERFLOW)) ? 1u : 0u;
        if (F.outcome)
            F.outcome[g] = (uint8_t)((safety ? 0 : winner + 1) | ((flags & ~FB_ROW_SAFETY_LIMIT) ? 0x80 : 0));
        if (F.mark_winner) F.header[g] = hdr | ((uint32_t)(winner + 1) << 24);
        uint32_t* row = F.rows ? row_tile + (size_t)threadIdx.x * F.row_words : nullptr;
        if (row) {
            const uint64_t gs = F.game_seed ? F.game_seed[g] : 0ull;
            reinterpret_cast<uint4*>(row)[0] =
                make_uint4((uint32_t)gs, (uint32_t)(gs >> 32), F.ordinal_base + g,
                           rounds | ((uint32_t)(safety ? 0xFF : winner) << 16) | (flags << 24));
        }
        // pass 2: the compact row
        for (int s = 0; row && s < k; s++) {
            const uint4 a = __ldcg(&seats[s].a);
            const uint4 b = __ldcg(&seats[s].b);
            const uint32_t rec = g * (uint32_t)k + (uint32_t)s;
            const uint32_t idx = F.ids_mode == 2 ? 0u : (F.perm ? (uint32_t)F.perm[rec] : rec);
            const int sid = F.ids_mode == 0 ? (int)idx : (F.ids_mode == 1 ? F.strategy_ids[idx] : s);
            uint32_t* w = row + 4 + s * 7;
            w[0] = a.y;
            w[1] = (uint32_t)sid;
            w[2] = a.z & HIGH_MASK;
            w[3] = a.w;
            w[4] = b.x;
            w[5] = b.y;
            w[6] = b.z;
        }
        if (row) {
            for (int w = 4 + 7 * k; w < F.row_words; w++) row[w] = 0u;  // padding
        }
    }
    if (F.rows) {  // the tile is one contiguous piece of the row array
        __syncthreads();
        const uint32_t first = blockIdx.x * blockDim.x;
        const uint32_t n_rows = min((uint32_t)blockDim.x, F.n_games - first);
        uint4* dst = reinterpret_cast<uint4*>(F.rows + (size_t)first * F.row_words);
        const uint4* src = reinterpret_cast<const uint4*>(row_tile);
        const uint32_t n_vec = n_rows * (uint32_t)F.row_words / 4u;
        for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = src[i];
    }
    // ---- totals: warp shuffle -> shared memory -> one global RED per CTA ----
    if (F.totals) {
        const int lane = threadIdx.x & 31;
        unsigned long long v[7] = {t_done, (unsigned long long)t_done - t_safe, t_safe, t_rolls, t_turns, t_err, t_ahead};
#pragma unroll
        for (int i = 0; i < 7; i++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(FULL, v[i], o);
        }
        if (lane == 0) {
            if (v[0]) atomicAdd(&s_tot[0], v[0]);
            if (v[1]) atomicAdd(&s_tot[1], v[1]);
            if (v[2]) atomicAdd(&s_tot[2], v[2]);
            if (v[3]) atomicAdd(&s_tot[3], v[3]);
            if (v[4]) atomicAdd(&s_tot[6], v[4]);
            if (v[5]) atomicAdd(&s_tot[7], v[5]);
            // totals[5] counts the 64-bit outputs the reference's generators would have PRODUCED:
            // play_kernel adds the ones it computed, this takes back the ones still queued
            if (v[6]) atomicAdd(&s_tot[5], 0ull - v[6]);
        }
        // wins by seat, warp-aggregated: a 64-bit shared-memory add is a compare-and-swap loop, and
        // 256 threads adding 1 to the same two or three counters made those loops the kernel's top
        // stall (short scoreboard 18.6 warps per issue cycle, profiles/r01_finish_kernel.md)
        for (int s = 0; s < k; s++) {
            const uint32_t won = __ballot_sync(FULL, winner == s);
            if (lane == 0 && won) atomicAdd(&s_tot[8 + s], (unsigned long long)__popc(won));
        }
        __syncthreads();
        if (threadIdx.x < FB_TOTALS_WIDTH && s_tot[threadIdx.x])
            atomicAdd(&F.totals[threadIdx.x], s_tot[threadIdx.x]);
    }
}

// Tallies by gather (run_tournament.py:336-353,375-391).  Every strategy is seated exactly once per
// shuffle, so thread (strategy i, chunk c) walks the chunk's shuffles, finds its game through the
// inverse permutation, and — if it won — adds the winner metrics from its seat record into
// registers; one RED per non-zero column at the end (~25x fewer L2 atomics than adding per game).
// Exposures need no game data at all: attempted = the chunk's shuffle count, completed = that
// minus the safety-limit games met on the way.
struct GatherParams {
    const Seat* seats;
    const uint32_t* header;  // rounds | flags << 16 | (winner seat + 1) << 24
    const int32_t* inv;      // [n_shuffles][n_strategies] position of strategy i in shuffle j
    const int32_t* strategy_ids;
    int n_strategies, n_tally_ids, n_shuffles, k;
    uint32_t gps;
    int chunk;    // shuffles per thread
    int slotted;  // 1: chunk c -> tally slot c, 0: single slot
    unsigned long long* tallies;
    unsigned long long* seat_tallies;  // [slots][ids][k][FB_SEAT_TALLY_WIDTH] or nullptr
    uint32_t* first_seen;  // [ids][4] first ordinal of: win, exposure, completed exposure, safety exposure
};

__global__ void __launch_bounds__(128) tally_gather_kernel(const GatherParams G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.n_strategies) return;
    const int c = blockIdx.y;
    const int j0 = c * G.chunk, j1 = min(j0 + G.chunk, G.n_shuffles);
    unsigned long long wins = 0, safety = 0, sum[10], sq[10];
    uint32_t first[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
    for (int m = 0; m < 10; m++) sum[m] = sq[m] = 0ull;
    for (int j = j0; j < j1; j++) {
        const uint32_t pos = (uint32_t)G.inv[(size_t)j * G.n_strategies + i];
        const uint32_t gi = pos / (uint32_t)G.k, seat = pos - gi * (uint32_t)G.k;
        const uint32_t game = (uint32_t)j * G.gps + gi;
        const uint32_t hdr = __ldg(&G.header[game]);
        const bool is_safety = (hdr >> 16) & FB_ROW_SAFETY_LIMIT;
        const bool won = !is_safety && ((hdr >> 24) & 15u) == seat + 1u;
        if (G.first_seen) {  // j ascends, so the first hit inside the chunk is the chunk's minimum
            const uint32_t ord = (uint32_t)j * (uint32_t)G.n_strategies + pos;
            if (won) first[0] = min(first[0], ord);
            first[1] = min(first[1], ord);
            first[is_safety ? 3 : 2] = min(first[is_safety ? 3 : 2], ord);
        }
        if (G.seat_tallies) {  // per (strategy, seat): wins, exposures, completed, safety limit
            const int sid_s = G.strategy_ids ? G.strategy_ids[i] : i;
            unsigned long long* S = G.seat_tallies +
                (((size_t)(G.slotted ? c : 0) * G.n_tally_ids + sid_s) * G.k + seat) * FB_SEAT_TALLY_WIDTH;
            if (won) atomicAdd(&S[0], 1ull);
            atomicAdd(&S[1], 1ull);
            atomicAdd(&S[is_safety ? 3 : 2], 1ull);
        }
        if (is_safety) {
            safety++;
        } else if (won) {
            const Seat* s = G.seats + ((size_t)game * G.k + seat);
            const uint4 a = __ldg(&s->a), b = __ldg(&s->b);
            // METRIC_LABELS order, run_tournament.py:109-121 (winner_hit_max_rounds stays 0)
            const unsigned long long v[10] = {a.y, hdr & 0xffffu, a.w & 0xffffu, a.w >> 16, a.z & HIGH_MASK,
                                              b.y & 0xffffu, b.y >> 16, b.z & 0xffffu, b.z >> 16, b.x >> 16};
            wins++;
#pragma unroll
            for (int m = 0; m < 10; m++) {
                sum[m] += v[m];
                sq[m] += v[m] * v[m];
            }
        }
    }
    const int sid = G.strategy_ids ? G.strategy_ids[i] : i;
    if (G.first_seen) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (first[q] != 0xffffffffu) atomicMin(&G.first_seen[(size_t)sid * 4 + q], first[q]);
    }
    unsigned long long* T = G.tallies + ((size_t)(G.slotted ? c : 0) * G.n_tally_ids + sid) * FB_TALLY_WIDTH;
    if (wins) atomicAdd(&T[0], wins);
    const unsigned long long seated = (unsigned long long)(j1 - j0);
    if (seated) atomicAdd(&T[1], seated);
    if (seated - safety) atomicAdd(&T[2], seated - safety);
    if (safety) atomicAdd(&T[3], safety);
#pragma unroll
    for (int m = 0; m < 10; m++) {
        if (sum[m]) {
            atomicAdd(&T[4 + m], sum[m]);
            atomicAdd(&T[4 + FB_N_METRICS + m], sq[m]);
        }
    }
}

// ---------------------------------------------------------------------------
// Unconditional all-player statistics (analysis/all_player_metrics.py:262-340)
// ---------------------------------------------------------------------------
// One thread per (strategy, deterministic batch): the strategy's exposures of the batch are its
// games of the batch's shuffles, in shuffle order -- the order in which the reference's unbuffered
// np.add.at meets them in the curated rows -- so the float64 sums of score / n_turns and
// score / n_rounds are reproduced bit for bit by adding in that order with round-to-nearest
// operations; everything else is integer.  Plain stores: a (slot, id) cell has one writer.
struct AllPlayerParams {
    const Seat* seats;
    const uint32_t* header;  // rounds | flags << 16 | (winner seat + 1) << 24
    const int32_t* inv;
    const int32_t* strategy_ids;
    int n_strategies, n_tally_ids, n_shuffles, k;
    uint32_t gps;
    int per_slot;  // shuffles per slot (> 0)
    long long* out;  // [slots][n_tally_ids][FB_ALLP_WIDTH]
};

__global__ void __launch_bounds__(128) allplayer_gather_kernel(const AllPlayerParams A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_strategies) return;
    const int c = blockIdx.y;
    const int j0 = c * A.per_slot, j1 = min(j0 + A.per_slot, A.n_shuffles);
    long long core[11];
    long long beh[FB_ALLP_BEHAVIOURS][3];
#pragma unroll
    for (int f = 0; f < 11; f++) core[f] = 0;
#pragma unroll
    for (int b = 0; b < FB_ALLP_BEHAVIOURS; b++) beh[b][0] = beh[b][1] = beh[b][2] = 0;
    double exact = 0.0, exact2 = 0.0, proxy = 0.0, proxy2 = 0.0;
    for (int j = j0; j < j1; j++) {
        const uint32_t pos = (uint32_t)A.inv[(size_t)j * A.n_strategies + i];
        const uint32_t gi = pos / (uint32_t)A.k, seat = pos - gi * (uint32_t)A.k;
        const uint32_t game = (uint32_t)j * A.gps + gi;
        const uint32_t hdr = __ldg(&A.header[game]);
        const bool safety = (hdr >> 16) & FB_ROW_SAFETY_LIMIT;
        const long long rounds = hdr & 0xffffu;
        const Seat* table = A.seats + (size_t)game * A.k;
        const uint4 a = __ldg(&table[seat].a), b = __ldg(&table[seat].b);
        const long long score = (int)a.y, turns = b.x & 0xffffu;
        const bool won = !safety && ((hdr >> 24) & 15u) == seat + 1u;
        core[0] += 1;
        core[safety ? 2 : 1] += 1;
        core[3] += won ? 1 : 0;
        const long long tmr = turns - rounds;
        core[4] += tmr != 0 ? 1 : 0;
        core[5] += score;
        core[6] += score * score;
        core[7] += turns;
        core[8] += turns * turns;
        core[9] += tmr;
        core[10] += tmr * tmr;
        const double sd = (double)score;
        const double e = turns ? __ddiv_rn(sd, (double)turns) : 0.0;
        const double p = rounds ? __ddiv_rn(sd, (double)rounds) : 0.0;
        exact = __dadd_rn(exact, e);
        exact2 = __dadd_rn(exact2, __dmul_rn(e, e));
        proxy = __dadd_rn(proxy, p);
        proxy2 = __dadd_rn(proxy2, __dmul_rn(p, p));
        long long v[FB_ALLP_BEHAVIOURS];
        bool seen_rank = false;
        v[0] = v[1] = 0;
        if (!safety) {  // rank = place in the stable sort by (-score, seat); margin to the winner
            int ahead = 0;
            int best = INT_MIN;
            for (int s = 0; s < A.k; s++) {
                const int sc = (int)__ldg(&table[s].a).y;
                best = max(best, sc);
                ahead += (sc > (int)score || (sc == (int)score && (uint32_t)s < seat)) ? 1 : 0;
            }
            v[0] = ahead + 1;
            v[1] = (long long)best - score;
            seen_rank = true;
        }
        v[2] = a.w >> 16;           // rolls
        v[3] = a.w & 0xffffu;       // farkles
        v[4] = a.z & HIGH_MASK;     // highest_turn
        v[5] = b.x >> 16;           // hot_dice
        v[6] = b.y & 0xffffu;       // smart_five_uses
        v[7] = b.y >> 16;           // n_smart_five_dice
        v[8] = b.z & 0xffffu;       // smart_one_uses
        v[9] = b.z >> 16;           // n_smart_one_dice
#pragma unroll
        for (int q = 0; q < FB_ALLP_BEHAVIOURS; q++) {
            const bool present = q >= 2 || seen_rank;
            beh[q][0] += present ? 1 : 0;
            beh[q][1] += present ? v[q] : 0;
            beh[q][2] += present ? v[q] * v[q] : 0;
        }
    }
    const int sid = A.strategy_ids ? A.strategy_ids[i] : i;
    long long* O = A.out + ((size_t)c * A.n_tally_ids + sid) * FB_ALLP_WIDTH;
#pragma unroll
    for (int f = 0; f < 11; f++) O[f] = core[f];
#pragma unroll
    for (int q = 0; q < FB_ALLP_BEHAVIOURS; q++) {
        O[11 + 3 * q] = beh[q][0];
        O[12 + 3 * q] = beh[q][1];
        O[13 + 3 * q] = beh[q][2];
    }
    O[41] = __double_as_longlong(exact);
    O[42] = __double_as_longlong(exact2);
    O[43] = __double_as_longlong(proxy);
    O[44] = __double_as_longlong(proxy2);
}

// ---------------------------------------------------------------------------
// RNG lag diagnostics, strategy groups (analysis/rng_diagnostics.py:2032-2077,1870-1901)
// ---------------------------------------------------------------------------
// The reference sorts every seat exposure of a (strategy, k) group by its RNG coordinate
// (root, k, shuffle, game, seat) and feeds win indicator and n_rounds to an online lagged-pair
// accumulator.  A strategy is seated exactly once per shuffle, so inside one (root, k) cell its
// sequence is simply "its game of shuffle 0, 1, 2, ...": observation j of strategy i is found
// through the inverse permutation, like the winner tallies.  All six sums are small integers, so
// int64 accumulation equals the reference's float64 accumulation exactly.
struct LagParams {
    const uint32_t* header;  // rounds | flags << 16 | (winner seat + 1) << 24
    const int32_t* inv;      // [n_shuffles][n_strategies]
    int n_strategies, n_shuffles, k;
    uint32_t gps;
    int chunk;  // shuffles per thread
    int n_lags, max_lag;
    int lags[FB_MAX_LAGS];
    unsigned long long* stats;  // [n_strategies][n_lags][FB_LAG_WIDTH], accumulated into
    uint32_t* edges;            // [n_strategies][2][max_lag]: first / last observations of the launch
};

// observation j of strategy i: n_rounds | win << 16
__device__ __forceinline__ uint32_t lag_observation(const LagParams& L, int i, int j) {
    const uint32_t pos = (uint32_t)L.inv[(size_t)j * L.n_strategies + i];
    const uint32_t gi = pos / (uint32_t)L.k, seat = pos - gi * (uint32_t)L.k;
    const uint32_t hdr = __ldg(&L.header[(uint32_t)j * L.gps + gi]);
    const bool won = !((hdr >> 16) & FB_ROW_SAFETY_LIMIT) && ((hdr >> 24) & 15u) == seat + 1u;
    return (hdr & 0xffffu) | (won ? 0x10000u : 0u);
}

// thread (strategy i, chunk c of the shuffles, lag z): the pairs (j - lag, j) with j in the chunk
__global__ void __launch_bounds__(128) lag_gather_kernel(const LagParams L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n_strategies) return;
    const int lag = L.lags[blockIdx.z];
    const int j0 = max(blockIdx.y * L.chunk, lag), j1 = min((blockIdx.y + 1) * L.chunk, L.n_shuffles);
    if (j0 >= j1) return;
    unsigned long long wx = 0, wy = 0, wxy = 0, rx = 0, ry = 0, rx2 = 0, ry2 = 0, rxy = 0;
    for (int j = j0; j < j1; j++) {
        const uint32_t x = lag_observation(L, i, j - lag), y = lag_observation(L, i, j);
        const unsigned long long xr = x & 0xffffu, yr = y & 0xffffu;
        const uint32_t xw = x >> 16, yw = y >> 16;
        wx += xw;
        wy += yw;
        wxy += xw & yw;
        rx += xr;
        ry += yr;
        rx2 += xr * xr;
        ry2 += yr * yr;
        rxy += xr * yr;
    }
    unsigned long long* S = L.stats + ((size_t)i * L.n_lags + blockIdx.z) * FB_LAG_WIDTH;
    atomicAdd(&S[0], (unsigned long long)(j1 - j0));
    if (wx) { atomicAdd(&S[1], wx); atomicAdd(&S[3], wx); }  // x^2 == x for an indicator
    if (wy) { atomicAdd(&S[2], wy); atomicAdd(&S[4], wy); }
    if (wxy) atomicAdd(&S[5], wxy);
    atomicAdd(&S[6], rx);
    atomicAdd(&S[7], ry);
    atomicAdd(&S[8], rx2);
    atomicAdd(&S[9], ry2);
    atomicAdd(&S[10], rxy);
}

// The first and last min(max_lag, n_shuffles) observations of every strategy: what a caller needs
// to add the pairs that straddle two launches of one cell (batches, ranks).
__global__ void lag_edges_kernel(const LagParams L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    const int cnt = min(L.max_lag, L.n_shuffles);
    if (i >= L.n_strategies || m >= cnt) return;
    uint32_t* E = L.edges + (size_t)i * 2 * L.max_lag;
    E[m] = lag_observation(L, i, m);
    E[L.max_lag + m] = lag_observation(L, i, L.n_shuffles - cnt + m);
}

}  // namespace fb
