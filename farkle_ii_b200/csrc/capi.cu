// capi.cu — C ABI of include/farkle_b200.h plus the prepare / building-block kernels.
//
// Kernel inventory (all sm_100a, integer-issue bound, no tensor-core work):
//   seed_tournament_kernel  coordinate_rng for every (game, seat) of a run of shuffles
//   seed_h2h_kernel         same for H2H attempts
//   seed_explicit_kernel    same for explicit coordinates
//   permute_warp_kernel     Generator.permutation, one warp per shuffle, jump-ahead draws
//   play_kernel             (play.cuh) the game state machine over L2-resident seat records
//   finish_kernel           (play.cuh) dense pass: winner, compact rows, tallies, totals
//   h2h_resolve_kernel      early-stop prefix rule over attempt outcomes
//   test kernels            seedseq / coordinate_seed / roll_dice / default_score
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/farkle_b200.h"
#include "matchup.cuh"
#include "play.cuh"
#include "rng.cuh"
#include "scoring.cuh"

using namespace fb;

// LCG jump-ahead constants for permute_warp_kernel: entry j holds A^j and
// 1 + A + ... + A^(j-1) (mod 2^128) for the cheap PCG64DXSM multiplier A.
struct JumpTable {
    uint64_t a_hi[33], a_lo[33], g_hi[33], g_lo[33];
};

// ---------------------------------------------------------------------------
// host state
// ---------------------------------------------------------------------------
namespace {

struct Ctx {
    int device = -1;
    int sm_count = 0, clock_khz = 0, cc_major = 0, cc_minor = 0;
    int max_smem_optin = 0;
    int l2_bytes = 0;
    ScoreLut* lut_dev = nullptr;
    JumpTable* jump_dev = nullptr;
    // cached buffers of fb_run_tournament_host
    void* host_ws = nullptr;
    size_t host_ws_bytes = 0;
    bool smem_opted = false;  // play_kernel dynamic shared memory opt-in done on this device
    bool finish_opted = false;
    // fb_run_tournament_host: compute / copy streams and the events that hand the row buffers over
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_played[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // fb_play_tournament_cells: preparation stream, hand-over events per workspace slot, and the
    // cell a previous call prepared ahead of time
    cudaStream_t s_prep = nullptr;
    cudaEvent_t ev_cells_entry = nullptr, ev_cell_prepared[2] = {nullptr, nullptr},
                ev_play_kernel_done[2] = {nullptr, nullptr};
    // set by fb_play_tournament_cells around a PLAY phase: launch_play records it right behind
    // play_kernel (before the finish pass) and clears it
    cudaEvent_t after_play_kernel = nullptr;
    struct Ahead {
        bool valid = false;
        uint64_t root_seed = 0, shuffle0 = 0;
        int k = 0, n_shuffles = 0, n_strategies = 0, slot = 0;
        int32_t target_score = 0, max_rounds = 0;
        const void* strategies = nullptr;
        const void* workspace = nullptr;
        size_t workspace_bytes = 0;
    } ahead;
};
// One context per CUDA device: a process may drive several GPUs (one engine each), and every entry
// point works on the context of the device that is CURRENT for the calling thread, the way the CUDA
// runtime itself resolves streams and allocations.  fb_init(device) makes `device` current and
// fills its slot.  The mutex of a context serialises only the calls that share its cached buffers
// (fb_run_tournament_host, fb_play_tournament_cells) on that device.
constexpr int MAX_DEVICES = 64;
Ctx g_ctxs[MAX_DEVICES];
std::mutex g_mus[MAX_DEVICES];
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}
#define g_ctx (g_ctxs[current_device()])
#define g_mu (g_mus[current_device()])
std::atomic<uint64_t> g_launches{0};

thread_local std::string t_err;
// CUDA-event pairs around the most recent play_kernel launches of this thread on each device (rings)
constexpr int EV_RING = 64;
struct EvRing {
    cudaEvent_t ev[EV_RING][2];
    bool made = false;
    uint64_t count = 0;  // launches recorded so far
};
thread_local EvRing t_rings[MAX_DEVICES];
#define t_ev (t_rings[current_device()].ev)
#define t_ev_made (t_rings[current_device()].made)
#define t_ev_count (t_rings[current_device()].count)

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}

#define FB_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(FB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

#define FB_REQUIRE_INIT()                                                        \
    do {                                                                         \
        if (g_ctx.device < 0)                                                    \
            return fail(FB_ERR_NO_DEVICE, "fb_init has not succeeded: no CUDA device bound"); \
    } while (0)

// Timeline hook (fb_timeline): when enabled, a timing event is recorded on the launching stream
// behind every kernel of the tournament path; fb_timeline_dump reports them relative to the first
// mark.  Off by default (no events, no cost).
struct TlMark {
    const char* name;
    int lane;  // 0 = caller's stream, 1 = preparation stream
    cudaEvent_t ev;
};
std::mutex g_tl_mu;
std::vector<TlMark> g_tl;
std::atomic<bool> g_tl_on{false};
inline void tl_mark(const char* name, cudaStream_t s, int lane = 0) {
    if (!g_tl_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_tl_mu);
    if (g_tl.size() >= 4096) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, s);
    g_tl.push_back({name, lane, ev});
}

inline int launch_check(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return FB_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
    Seat* seats;
    uint32_t* header;
    uint64_t* game_seed;
    int32_t* limits;
    unsigned int* counter;  // [0] play ordinal, [1] number of HDR_LONG games
    uint32_t* long_list;    // [n_games] worst case
    uint8_t* prefix;        // [n_shuffles or n_blocks] uint4 stream-prefix pools (<= n_games entries)
    uint8_t* extra;  // mode specific tail (perm / h2h tables)
    size_t extra_bytes;
};

size_t ws_core_bytes(int k, uint64_t n) {
    return align_up(n * (uint64_t)k * sizeof(Seat), 256) +
           align_up(n * 4, 256) + align_up(n * 8, 256) + align_up(n * 8, 256) + 256 + align_up(n * 4, 256) +
           align_up(n * 16, 256);
}

bool carve(void* base, size_t bytes, int k, uint64_t n, Workspace& w) {
    uint8_t* p = static_cast<uint8_t*>(base);
    const size_t core = ws_core_bytes(k, n);
    if (bytes < core) return false;
    w.seats = reinterpret_cast<Seat*>(p);
    p += align_up(n * (uint64_t)k * sizeof(Seat), 256);
    w.header = reinterpret_cast<uint32_t*>(p);
    p += align_up(n * 4, 256);
    w.game_seed = reinterpret_cast<uint64_t*>(p);
    p += align_up(n * 8, 256);
    w.limits = reinterpret_cast<int32_t*>(p);
    p += align_up(n * 8, 256);
    w.counter = reinterpret_cast<unsigned int*>(p);
    p += 256;
    w.long_list = reinterpret_cast<uint32_t*>(p);
    p += align_up(n * 4, 256);
    w.prefix = p;
    p += align_up(n * 16, 256);
    w.extra = p;
    w.extra_bytes = bytes - core;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------
// device helpers / kernels
// ---------------------------------------------------------------------------
// Fresh seat records: seeded stream, zero score / counters, the seat's strategy.
__device__ __forceinline__ void store_seat(Seat* seat, const Pcg& g, const fb_strategy_t* table,
                                           uint32_t strat_index) {
    const uint2 sv = reinterpret_cast<const uint2*>(table)[strat_index];
    const SeatConsts sc = seat_consts((int)sv.x, (int)(int16_t)(sv.y & 0xffffu), sv.y >> 16);
    seat->state = make_uint4((uint32_t)g.lo, (uint32_t)(g.lo >> 32), (uint32_t)g.hi, (uint32_t)(g.hi >> 32));
    seat->a = make_uint4(0u, 0u, 0u, 0u);
    seat->b = make_uint4(0u, 0u, 0u, 0u);
    seat->inc = make_uint4((uint32_t)g.ilo, (uint32_t)(g.ilo >> 32), (uint32_t)g.ihi, (uint32_t)(g.ihi >> 32));
    seat->cst = make_uint4((uint32_t)sc.st_d, sc.kf | ((uint32_t)sc.dt_d << CST_DT_SHIFT),
                           sc.dbase | (sc.tab_off << 16), strat_index);
}

// Longest-first scheduling hint: the lanes of a warp that hold a game whose seats can never
// bank append its ordinal to the long list (one atomic per warp) and every game's header is
// initialised.  Must be called by all 32 lanes of the warp.
__device__ __forceinline__ void publish_game(bool owns_game, bool is_long, uint32_t g, uint32_t* header,
                                             uint32_t* long_list, unsigned int* counter) {
    const uint32_t m = __ballot_sync(0xffffffffu, owns_game && is_long);
    if (owns_game) header[g] = is_long ? HDR_LONG : 0u;
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&counter[1], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (owns_game && is_long) long_list[base + __popc(m & ((1u << lane) - 1u))] = g;
    }
}
__device__ __forceinline__ bool entry_never_banks(const fb_strategy_t* table, uint32_t index) {
    const uint2 sv = reinterpret_cast<const uint2*>(table)[index];
    return never_banks((int)(int16_t)(sv.y & 0xffffu), sv.y >> 16);
}

__device__ __forceinline__ uint64_t coord_fingerprint(const Coord& c, bool as_u32) {
    uint32_t pool[4], w[2];
    ss_pool_coord(c, pool);
    ss_generate<2>(pool, w);
    return as_u32 ? (uint64_t)w[0] : ((uint64_t)w[0] | ((uint64_t)w[1] << 32));
}

// One thread per (game, seat) of shuffles shuffle0.. (run_tournament.py:336-364).
__global__ void __launch_bounds__(256) seed_tournament_kernel(
    uint64_t root, int k, uint64_t shuffle0, uint32_t gps, uint64_t n_games, const int32_t* perm,
    int n_strategies, int32_t target, int32_t max_rounds, const uint64_t* ov_shuffle,
    const uint32_t* ov_game, const int32_t* ov_rounds, int n_ov, int want_seeds,
    const fb_strategy_t* table, Seat* seats, uint64_t* game_seed, int32_t* limits,
    uint32_t* header, uint32_t* long_list, unsigned int* counter, const uint4* prefix) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_games * (uint64_t)k;
    const uint64_t g = live ? t / (uint64_t)k : 0;
    const uint32_t s = (uint32_t)(t - g * (uint64_t)k);
    const uint64_t sl = g / gps;
    const uint32_t gi = (uint32_t)(g - sl * gps);
    bool is_long = false;
    if (live) {
        Coord c{FB_PURPOSE_TOURNAMENT_PLAYER, root, (uint64_t)k, shuffle0 + sl, 0, 0, gi, s, 0};
        Pcg pg;
        const uint4 pre = prefix[sl];  // hash of the 12 coordinate words shared by the shuffle
        uint32_t pool[4] = {pre.x, pre.y, pre.z, pre.w};
        ss_pool_suffix(pool, gi, s, 0);
        pcg_seed_pool(pg, pool);
        const int32_t* seat_ids = perm + sl * (uint64_t)n_strategies + (uint64_t)gi * k;
        store_seat(seats + t, pg, table, (uint32_t)seat_ids[s]);
        if (s == 0) {
            is_long = true;
            for (int j = 0; j < k; j++) is_long = is_long && entry_never_banks(table, (uint32_t)seat_ids[j]);
            if (want_seeds) {
                Coord gc = c;
                gc.purpose = FB_PURPOSE_TOURNAMENT_GAME;
                gc.seat_index = 0;
                game_seed[g] = coord_fingerprint(gc, true);
            }
            if (limits) {
                int32_t mr = max_rounds;
                for (int o = 0; o < n_ov; o++)
                    if (ov_shuffle[o] == shuffle0 + sl && ov_game[o] == gi) {
                        mr = min(ov_rounds[o], FB_MAX_ROUNDS + 1);  // see pack_limits_kernel
                        break;
                    }
                limits[2 * g] = target;
                limits[2 * g + 1] = mr;
            }
        }
    }
    publish_game(live && s == 0, is_long, (uint32_t)g, header, long_list, counter);
}

// Exclusive prefix sum of n_attempts (single block; n_blocks is at most ~1e5).
__global__ void h2h_offsets_kernel(const uint32_t* n_attempts, int n_blocks, uint64_t* offsets) {
    __shared__ uint64_t carry;
    __shared__ uint64_t part[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        uint64_t v = i < n_blocks ? n_attempts[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < blockDim.x; o <<= 1) {
            uint64_t add = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
            __syncthreads();
            part[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_blocks) offsets[i] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_blocks] = carry;
}

// Shared 12-word stream prefix of every H2H block (root, k = 2, pair_id, order).
__global__ void h2h_prefix_kernel(uint64_t root, int n_blocks, const uint64_t* pair_id, const uint8_t* order,
                                  uint4* prefix) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    Coord pc{FB_PURPOSE_H2H_PLAYER, root, 2, 0, pair_id[b], order[b], 0, 0, 0};
    uint32_t pool[4];
    ss_pool_prefix(pc, pool);
    prefix[b] = make_uint4(pool[0], pool[1], pool[2], pool[3]);
}

// One thread per (attempt, seat) (h2h_schedule.py:1169-1202).
__global__ void __launch_bounds__(256) seed_h2h_kernel(
    uint64_t root, int n_blocks, const uint64_t* pair_id, const uint8_t* order,
    const fb_strategy_t* seat1, const fb_strategy_t* seat2, const uint32_t* attempt0,
    const uint64_t* offsets, uint64_t total, int want_seeds, Seat* seats,
    uint64_t* game_seed, uint32_t* header, uint32_t* long_list, unsigned int* counter, const uint4* prefix) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < total * 2;
    const uint64_t g = live ? t >> 1 : 0;
    const uint32_t s = (uint32_t)(t & 1);
    bool is_long = false;
    if (live) {
        int lo = 0, hi = n_blocks - 1;  // last block with offsets[b] <= g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (offsets[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int b = lo;
        const uint64_t a = (uint64_t)attempt0[b] + (g - offsets[b]);
        Coord c{FB_PURPOSE_H2H_PLAYER, root, 2, 0, pair_id[b], order[b], a, s, 0};
        Pcg pg;
        const uint4 pre = prefix[b];
        uint32_t pool[4] = {pre.x, pre.y, pre.z, pre.w};
        ss_pool_suffix(pool, a, s, 0);
        pcg_seed_pool(pg, pool);
        store_seat(seats + t, pg, s ? seat2 : seat1, (uint32_t)b);
        if (s == 0) {
            is_long = entry_never_banks(seat1, (uint32_t)b) && entry_never_banks(seat2, (uint32_t)b);
            if (want_seeds) {
                Coord gc = c;
                gc.purpose = FB_PURPOSE_H2H_GAME;
                game_seed[g] = coord_fingerprint(gc, false);
            }
        }
    }
    publish_game(live && s == 0, is_long, (uint32_t)g, header, long_list, counter);
}

__global__ void __launch_bounds__(256) seed_explicit_kernel(const uint64_t* coords, uint64_t n_games, int k,
                                                            const fb_strategy_t* seat_table, Seat* seats,
                                                            uint32_t* header,
                                                            uint32_t* long_list, unsigned int* counter) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_games * (uint64_t)k;
    const uint64_t g = live ? t / (uint64_t)k : 0;
    const uint32_t s = (uint32_t)(t - g * (uint64_t)k);
    bool is_long = false;
    if (live) {
        const uint64_t* cc = coords + g * 7;
        Coord c{(uint32_t)cc[0], cc[1], cc[2], cc[3], cc[4], cc[5], cc[6], s, 0};
        Pcg pg;
        pcg_seed_coord(pg, c);
        store_seat(seats + t, pg, seat_table, (uint32_t)t);
        if (s == 0) {
            is_long = true;
            for (int j = 0; j < k; j++) is_long = is_long && entry_never_banks(seat_table, (uint32_t)(t + j));
        }
    }
    publish_game(live && s == 0, is_long, (uint32_t)g, header, long_list, counter);
}

__global__ void pack_limits_kernel(const int32_t* tv, int32_t t0, const int32_t* mv, int32_t m0,
                                   uint64_t n, int32_t* limits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    limits[2 * i] = tv ? tv[i] : t0;
    // n_rounds lives in 16 header bits and an int16 row column: a per-game limit beyond
    // FB_MAX_ROUNDS is cut to FB_MAX_ROUNDS + 1, and a game that really gets that far is reported
    // as an error row (FB_ROW_I16_OVERFLOW, finish_kernel) instead of wrapping silently.
    limits[2 * i + 1] = min(mv ? mv[i] : m0, FB_MAX_ROUNDS + 1);
}

// Tally ids must address [0, n_tally_ids): one flag for the whole table (checked on the host).
__global__ void check_ids_kernel(const int32_t* ids, int n, int n_tally_ids, unsigned int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (ids[i] < 0 || ids[i] >= n_tally_ids)) atomicOr(bad, 1u);
}

// Generator.permutation(n) per shuffle: Fisher-Yates from the top with masked
// rejection on buffered 32-bit draws (run_tournament.py:312-318).  The swap chain is
// sequential per shuffle but the draws are not.  Per batch, lane j of a lane group jumps the LCG
// ahead by j steps (S_j = A^j S + (A^j-1)/(A-1) inc, constants in JumpTable) and emits the j-th
// 64-bit output, i.e. 2 * PERM_LANES buffered 32-bit draws per batch; then every lane of the group
// replays the same accept/reject walk over them and one lane swaps in the shuffle's uint16 array in
// shared memory.
__device__ __forceinline__ void mul128(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo, uint64_t& rhi,
                                       uint64_t& rlo) {
    rlo = alo * blo;
    rhi = __umul64hi(alo, blo) + alo * bhi + ahi * blo;
}

// Sub-warp layout: PERM_LANES lanes share one shuffle, so a warp carries PERM_SHUFFLES different
// shuffles through the same instruction stream.  The kernel is LATENCY bound (one warp per CTA, a
// few warps per SM, bounded by the shared memory of the arrays: 4,300 shuffles of a 5,160-entry
// grid need two waves), so what counts is the length of the dependent chain per draw:
//   * the draws of batch b+1 are generated (four 128-bit multiplies and the output function per
//     lane, no dependence on the walk) in the same straight-line code as the walk over batch b,
//     into the other half of a double-buffered ring: the arithmetic fills the shared-memory
//     latency of the swaps instead of preceding them (a further split of the walk into a
//     register-only accept pass one batch ahead of the swaps gained nothing: with one warp per
//     scheduler the loop is bound by the ALU pipe's one warp instruction per two cycles, ~320
//     instructions per batch of 16 draws);
//   * the walk's own chain is v = h & mask -> accept = v <= i -> i -= accept -> mask halves when
//     i drops to mask >> 1 (compare + select, no count-leading-zeros per draw);
//   * a rejected draw costs no swap (the swap is predicated on accept), an accepted one is two
//     loads and two stores by ONE lane of the group (a swap is not idempotent).
// After the last swap (i == 0) further draws only "swap" a[0] with itself, so finished and unused
// groups ride along without a live flag in the chain.
constexpr int PERM_LANES = 8;
constexpr int PERM_SHUFFLES = 32 / PERM_LANES;  // per warp (= per CTA: one warp per CTA)
constexpr int PERM_RING = 2 * PERM_LANES;       // 32-bit draws per batch and shuffle

__host__ __device__ inline size_t perm_bytes_per_shuffle(int n) {
    return (((size_t)n * 2 + 15) & ~(size_t)15) + 2 * PERM_RING * 4;  // array + double-buffered ring
}

__global__ void __launch_bounds__(32) permute_warp_kernel(uint64_t root, int k, uint64_t shuffle0,
                                                          int n_shuffles, int n,
                                                          const JumpTable* __restrict__ jt,
                                                          int32_t* out, int32_t* inv, uint4* prefix) {
    extern __shared__ __align__(16) uint8_t perm_smem[];
    const int lane = threadIdx.x & 31;
    const int sub = lane / PERM_LANES, sl = lane % PERM_LANES;
    const int j = blockIdx.x * PERM_SHUFFLES + sub;  // shuffle handled by this lane group
    const bool valid = j < n_shuffles;
    const size_t per = perm_bytes_per_shuffle(n);
    uint16_t* a = reinterpret_cast<uint16_t*>(perm_smem + sub * per);
    uint32_t* ring = reinterpret_cast<uint32_t*>(perm_smem + sub * per + (per - 2 * PERM_RING * 4));
    for (int i = sl; i < n; i += PERM_LANES) a[i] = (uint16_t)i;
    Coord c{FB_PURPOSE_SHUFFLE_PERMUTATION, root, (uint64_t)k, shuffle0 + (uint64_t)j, 0, 0, 0, 0, 0};
    Pcg g;
    pcg_seed_coord(g, c);  // every lane of the group derives the same stream
    const uint64_t ahi = jt->a_hi[sl], alo = jt->a_lo[sl], ghi = jt->g_hi[sl], glo = jt->g_lo[sl];
    const uint64_t nahi = jt->a_hi[PERM_LANES], nalo = jt->a_lo[PERM_LANES];
    const uint64_t nghi = jt->g_hi[PERM_LANES], nglo = jt->g_lo[PERM_LANES];
    // batch: the lane's state S_sl, its output -> halves 2*sl, 2*sl+1 of ring buffer b; then S
    // advances by PERM_LANES steps
    auto generate = [&](int b) {
        uint64_t thi, tlo, uhi, ulo;
        mul128(ahi, alo, g.hi, g.lo, thi, tlo);
        mul128(ghi, glo, g.ihi, g.ilo, uhi, ulo);
        const uint64_t slo = tlo + ulo;
        const uint64_t shi = thi + uhi + (slo < tlo ? 1u : 0u);
        const uint64_t o = pcg_output(shi, slo);
        reinterpret_cast<uint2*>(ring + b * PERM_RING)[sl] = make_uint2((uint32_t)o, (uint32_t)(o >> 32));
        mul128(nahi, nalo, g.hi, g.lo, thi, tlo);
        mul128(nghi, nglo, g.ihi, g.ilo, uhi, ulo);
        g.lo = tlo + ulo;
        g.hi = thi + uhi + (g.lo < tlo ? 1u : 0u);
    };
    int i = valid ? n - 1 : 0;
    uint32_t mask = 0xffffffffu >> __clz(i | 1);
    int cur = 0;
    generate(0);
    __syncwarp();
    while (__any_sync(0xffffffffu, i >= 1)) {
        generate(cur ^ 1);  // the other buffer: its readers passed the __syncwarp below one trip ago
        const uint4* rp = reinterpret_cast<const uint4*>(ring + cur * PERM_RING);
        uint32_t hs[PERM_RING];
#pragma unroll
        for (int q = 0; q < PERM_RING / 4; q++) {
            const uint4 h = rp[q];
            hs[4 * q] = h.x;
            hs[4 * q + 1] = h.y;
            hs[4 * q + 2] = h.z;
            hs[4 * q + 3] = h.w;
        }
#pragma unroll
        for (int d = 0; d < PERM_RING; d++) {
            const uint32_t v = hs[d] & mask;
            const bool accept = v <= (uint32_t)i;
            if (sl == 0 && accept) {
                const uint16_t x = a[i], y = a[v];
                a[i] = y;
                a[v] = x;
            }
            i = max(i - (accept ? 1 : 0), 0);
            const uint32_t half = mask >> 1;
            mask = (uint32_t)i <= half ? half : mask;  // smallest 2^b - 1 that is >= i
        }
        __syncwarp();
        cur ^= 1;
    }
    __syncwarp();
    if (!valid) return;
    int32_t* dst = out + (size_t)j * n;
    for (int t = sl; t < n; t += PERM_LANES) dst[t] = (int32_t)a[t];
    if (inv) {  // inverse permutation: position of every strategy in this shuffle
        int32_t* idst = inv + (size_t)j * n;
        for (int t = sl; t < n; t += PERM_LANES) idst[a[t]] = t;
    }
    if (prefix && sl == 0) {  // shared 12-word prefix of this shuffle's seat streams (ss_pool_prefix)
        Coord pc{FB_PURPOSE_TOURNAMENT_PLAYER, root, (uint64_t)k, shuffle0 + (uint64_t)j, 0, 0, 0, 0, 0};
        uint32_t pool[4];
        ss_pool_prefix(pc, pool);
        prefix[j] = make_uint4(pool[0], pool[1], pool[2], pool[3]);
    }
}

// Fallback for grids too large for shared memory: the array lives in global memory.
__global__ void __launch_bounds__(128) permute_kernel(uint64_t root, int k, uint64_t shuffle0,
                                                      int n_shuffles, int n, int32_t* out, int32_t* inv,
                                                      uint4* prefix) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_shuffles) return;
    Coord c{FB_PURPOSE_SHUFFLE_PERMUTATION, root, (uint64_t)k, shuffle0 + (uint64_t)j, 0, 0, 0, 0, 0};
    PcgStream s;
    pcg_seed_coord(s.g, c);
    s.saved = 0;
    s.has32 = false;
    int32_t* a = out + (size_t)j * n;
    for (int i = 0; i < n; i++) a[i] = i;
    for (int i = n - 1; i >= 1; i--) {
        const uint32_t mask = 0xffffffffu >> __clz(i);
        const uint32_t v = s.interval((uint32_t)i, mask);
        const int32_t t = a[i];
        a[i] = a[v];
        a[v] = t;
    }
    if (inv) {
        int32_t* ia = inv + (size_t)j * n;
        for (int i = 0; i < n; i++) ia[a[i]] = i;
    }
    if (prefix) {
        Coord pc{FB_PURPOSE_TOURNAMENT_PLAYER, root, (uint64_t)k, shuffle0 + (uint64_t)j, 0, 0, 0, 0, 0};
        uint32_t pool[4];
        ss_pool_prefix(pc, pool);
        prefix[j] = make_uint4(pool[0], pool[1], pool[2], pool[3]);
    }
}

// Early-stop rule of _simulate_block_from_manifest (h2h_schedule.py:1167-1235):
// one warp per block walks the outcome bytes in order.
__global__ void h2h_resolve_kernel(int n_blocks, const uint32_t* n_attempts, const uint64_t* offsets,
                                   const uint8_t* outcome, const int32_t* required, int32_t* progress) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= n_blocks) return;
    int attempted = progress[b * 5 + 0], completed = progress[b * 5 + 1], safety = progress[b * 5 + 2];
    int w1 = progress[b * 5 + 3], w2 = progress[b * 5 + 4];
    const int target = required[b];
    const uint8_t* oc = outcome + offsets[b];
    const uint32_t n = n_attempts[b];
    for (uint32_t base = 0; base < n && completed < target; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t o = i < n ? (oc[i] & 0x7fu) : 0xffu;
        const uint32_t mc = __ballot_sync(FULL, o == 1u || o == 2u);
        const uint32_t valid = __ballot_sync(FULL, i < n);
        // take the shortest prefix of this 32-attempt window that reaches the target
        int take = __popc(valid);
        const int room = target - completed;
        if (__popc(mc) >= room) {
            // position of the room-th completed attempt
            uint32_t m = mc;
            for (int r = 1; r < room; r++) m &= m - 1;
            take = __ffs(m);
        }
        const uint32_t tm = take >= 32 ? FULL : ((1u << take) - 1u);
        attempted += take;
        completed += __popc(mc & tm);
        safety += __popc(__ballot_sync(FULL, o == 0u) & tm);
        w1 += __popc(__ballot_sync(FULL, o == 1u) & tm);
        w2 += __popc(__ballot_sync(FULL, o == 2u) & tm);
    }
    if (lane == 0) {
        progress[b * 5 + 0] = attempted; progress[b * 5 + 1] = completed; progress[b * 5 + 2] = safety;
        progress[b * 5 + 3] = w1; progress[b * 5 + 4] = w2;
    }
}

// ---- building-block / test kernels -----------------------------------------
__global__ void seedseq_kernel(const uint32_t* entropy, int n_entropy, uint64_t n_streams,
                               int n_words, uint32_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams) return;
    uint32_t pool[4];
    ss_pool_generic(entropy + i * n_entropy, n_entropy, pool);
    uint32_t hc = SS_INIT_B;
    for (int w = 0; w < n_words; w++) {
        uint32_t v = pool[w & 3] ^ hc;
        hc *= SS_MULT_B;
        v *= hc;
        v ^= v >> 16;
        out[i * n_words + w] = v;
    }
}

__global__ void coord_seeds_kernel(Coord c, int vary, uint64_t base, uint64_t n, int as_u32,
                                   uint64_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (vary == 0) c.shuffle_index = base + i;
    else if (vary == 1) c.game_index = base + i;
    else c.pair_id = base + i;
    out[i] = coord_fingerprint(c, as_u32 != 0);
}

__global__ void seed_streams_kernel(const uint64_t* coords, uint64_t n, uint64_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* cc = coords + i * 9;
    Coord c{(uint32_t)cc[0], cc[1], cc[2], cc[3], cc[4], cc[5], cc[6], cc[7], cc[8]};
    Pcg g;
    pcg_seed_coord(g, c);
    out[i * 4 + 0] = g.hi; out[i * 4 + 1] = g.lo; out[i * 4 + 2] = g.ihi; out[i * 4 + 3] = g.ilo;
}

// Same half-buffer roll code path as play_kernel's P phase, one stream per thread.
__global__ void roll_dice_kernel(const uint64_t* state_inc, const uint32_t* half_buffer, uint64_t n,
                                 const int32_t* n_dice, int n_rolls, uint8_t* faces) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pcg rng{state_inc[i * 4], state_inc[i * 4 + 1], state_inc[i * 4 + 2], state_inc[i * 4 + 3]};
    uint32_t saved = half_buffer ? half_buffer[i * 2 + 1] : 0u;
    bool has32 = half_buffer ? half_buffer[i * 2] != 0u : false;
    for (int r = 0; r < n_rolls; r++) {
        const int nd = n_dice[r];
        const uint32_t p = has32 ? 0u : 1u;
        const int nw = (nd + (int)p) >> 1;
        uint64_t shi = rng.hi, slo = rng.lo;
        uint32_t H[7] = {saved, 0, 0, 0, 0, 0, 0};
        for (int w = 0; w < nw; w++) {
            const uint64_t o = pcg_output(shi, slo);
            pcg_step(shi, slo, rng.ihi, rng.ilo);
            H[1 + 2 * w] = (uint32_t)o;
            H[2 + 2 * w] = (uint32_t)(o >> 32);
        }
        uint8_t f[6] = {0, 0, 0, 0, 0, 0};
        bool rej = false;
        for (int d = 0; d < nd; d++) {
            const uint64_t m = (uint64_t)H[d + p] * 6u;
            rej |= (uint32_t)m < 4u;
            f[d] = (uint8_t)(1u + (uint32_t)(m >> 32));
        }
        const uint32_t q = p + (uint32_t)nd;
        bool nhas = (q & 1u) == 0u;
        uint32_t nsaved = H[q <= 6 ? q : 6];
        if (rej) {
            PcgStream s{rng, saved, has32};
            uint32_t words = 0;
            for (int d = 0; d < nd; d++) f[d] = (uint8_t)(1u + s.die0(words));
            shi = s.g.hi; slo = s.g.lo; nhas = s.has32; nsaved = s.saved;
        }
        rng.hi = shi; rng.lo = slo; has32 = nhas; saved = nsaved;
        for (int d = 0; d < 6; d++) faces[(i * n_rolls + r) * 6 + d] = f[d];
    }
}

__global__ void default_score_kernel(const ScoreLut* lut, const uint8_t* faces, const int32_t* ts_pre,
                                     const fb_strategy_t* strat, uint64_t n, int32_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t hist = 0;
    int nd = 0;
    for (int d = 0; d < 6; d++) {
        const uint32_t f = faces[i * 6 + d];
        if (f) { hist += 1u << (3u * (f - 1u)); nd++; }
    }
    const uint2 sv = reinterpret_cast<const uint2*>(strat)[i];
    const SeatConsts sc = seat_consts((int)sv.x, (int)(int16_t)(sv.y & 0xffffu), sv.y >> 16);
    const uint32_t e = lut_lookup(lut, sc.tab_off, hist);
    const int rscore = (int)(e & 127u) * 50;
    int used = (int)((e >> 7) & 7u);
    const uint32_t dd = rscore ? smart_discards(lut, sc.dbase, e, nd, ts_pre[i], sc.st_d, sc.dt_d) : 0u;
    const int d5 = (int)(dd & 3u), d1 = (int)(dd >> 2);
    used -= d5 + d1;
    out[i * 5 + 0] = rscore - 50 * d5 - 100 * d1;
    out[i * 5 + 1] = used;
    out[i * 5 + 2] = nd - used;
    out[i * 5 + 3] = d5;
    out[i * 5 + 4] = d1;
}

// Integer-issue roofline probe: register-only chains of 32-bit integer instructions, no memory.
// sm_100 issues one warp instruction per cycle and scheduler, but each of the two integer pipes
// (FMA-heavy: IMAD; ALU: LOP3 / IADD3 / SHF) accepts one every second cycle, so the 1.0 IPC peak
// needs a 50/50 mix with neighbouring instructions independent of each other.
//   variant 0  8 chains, chain-major (mad, xor, mad, add of one chain back to back: every
//              instruction depends on the one before it; round 1's probe)
//   variant 1  8 chains, op-major (8 independent mads, 8 xors, 8 mads, 8 adds)
//   variant 2  16 chains, op-major, mad and ALU instructions interleaved one by one
//   variant 3  xor / add chains: ptxas emits LOP3 (ALU pipe) and IMAD.IADD (FMA pipe, two source
//              registers) alternately -- the mix that reaches ~0.98 IPC (profiles/r02_issue_peak.md)
//   variant 4  mad only (FMA-heavy pipe alone): the single-pipe rate, 0.5 IPC
template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) issue_peak_kernel(int iters, uint32_t seed, uint32_t* sink) {
    constexpr int NC = VARIANT == 2 ? 16 : 8;
    uint32_t x[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) x[i] = seed + threadIdx.x * (uint32_t)NC + i;
    const uint32_t m = seed | 1u, c = seed ^ 0x9e3779b9u;
    for (int it = 0; it < iters; it++) {
        if (VARIANT == 0) {
#pragma unroll
            for (int i = 0; i < NC; i++) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(c));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c), "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m));
            }
        } else if (VARIANT == 1) {
#pragma unroll
            for (int i = 0; i < NC; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(c));
#pragma unroll
            for (int i = 0; i < NC; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
#pragma unroll
            for (int i = 0; i < NC; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c), "r"(m));
#pragma unroll
            for (int i = 0; i < NC; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m));
        } else if (VARIANT == 2) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {  // chains 0-7 on the FMA pipe while 8-15 are on the ALU pipe
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[h ? i + 8 : i]) : "r"(m), "r"(c));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[h ? i : i + 8]) : "r"(c));
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[h ? i + 8 : i]) : "r"(c), "r"(m));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[h ? i : i + 8]) : "r"(m));
                }
            }
        } else if (VARIANT == 3) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
#pragma unroll
                for (int i = 0; i < NC; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
#pragma unroll
                for (int i = 0; i < NC; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; r++) {
#pragma unroll
                for (int i = 0; i < NC; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(c));
#pragma unroll
                for (int i = 0; i < NC; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c), "r"(m));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
}
constexpr int ISSUE_PEAK_VARIANTS = 5;
// lane instructions per thread and iteration of each variant
constexpr int ISSUE_PEAK_OPS[ISSUE_PEAK_VARIANTS] = {32, 32, 64, 32, 32};

// ---------------------------------------------------------------------------
// play launch
// ---------------------------------------------------------------------------
namespace {

int launch_play(const PlayParams& P_in, const FinishParams& F, cudaStream_t stream) {
    PlayParams P = P_in;
    P.roll_limit = ROLL_LIMIT;
    if (const char* env = getenv("FB_TEST_ROLL_LIMIT")) {  // test knob: reach the error path (tests/test_gpu_parity.py)
        const int lim = atoi(env);
        if (lim >= 1 && lim <= ROLL_LIMIT) P.roll_limit = lim;
    }
    // One persistent CTA per SM; shared memory holds only the lookup tables.
    int warps = 32;
    // Keep the seat records of the games in flight (lanes x k x 80 B on every SM) inside the L2:
    // beyond it every turn switch streams from HBM (k >= 10: 145 MB at full occupancy) and fewer
    // lanes are faster (k=12: 24 warps 13.7 ms vs 32 warps 15.3 ms for the 4,300-shuffle cell).
    if (g_ctx.l2_bytes > 0) {
        const double budget = 0.83 * (double)g_ctx.l2_bytes;
        const int fit = (int)(budget / ((double)g_ctx.sm_count * P.k * sizeof(Seat) * 32.0));
        warps = std::max(16, std::min(32, fit));
    }
    if (const char* env = getenv("FB_PLAY_WARPS")) {  // tuning knob: resident warps per SM
        const int cap = atoi(env);
        if (cap >= 1 && cap <= 32) warps = cap;
    }
    const uint64_t lanes_needed = P.n_games;
    int grid = g_ctx.sm_count;
    const uint64_t per_cta = (uint64_t)warps * 32;
    if ((uint64_t)grid * per_cta > lanes_needed) {  // small launch: fewer CTAs, then fewer warps
        grid = (int)((lanes_needed + per_cta - 1) / per_cta);
        if (grid < 1) grid = 1;
        if (grid == 1) {
            warps = (int)((lanes_needed + 31) / 32);
            if (warps < 1) warps = 1;
        }
    }
    const size_t smem = play_smem_bytes(P.k == 2);
    if (!g_ctx.smem_opted) {
        const int two = (int)play_smem_bytes(true), any = (int)play_smem_bytes(false);
        FB_CUDA(cudaFuncSetAttribute(play_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
        FB_CUDA(cudaFuncSetAttribute(play_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
        FB_CUDA(cudaFuncSetAttribute(play_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, any));
        FB_CUDA(cudaFuncSetAttribute(play_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, any));
        g_ctx.smem_opted = true;
    }
    if (!t_ev_made) {
        for (int i = 0; i < EV_RING; i++) {
            FB_CUDA(cudaEventCreate(&t_ev[i][0]));
            FB_CUDA(cudaEventCreate(&t_ev[i][1]));
        }
        t_ev_made = true;
    }
    cudaEvent_t* ev = t_ev[t_ev_count % EV_RING];
    tl_mark("play_begin", stream);
    FB_CUDA(cudaEventRecord(ev[0], stream));
    const dim3 block((unsigned)warps * 32u);
    if (P.k == 2) {
        if (P.limits) play_kernel<true, true><<<grid, block, smem, stream>>>(P, g_ctx.lut_dev);
        else play_kernel<false, true><<<grid, block, smem, stream>>>(P, g_ctx.lut_dev);
    } else {
        if (P.limits) play_kernel<true, false><<<grid, block, smem, stream>>>(P, g_ctx.lut_dev);
        else play_kernel<false, false><<<grid, block, smem, stream>>>(P, g_ctx.lut_dev);
    }
    int rc = launch_check("play_kernel");
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(ev[1], stream));
    t_ev_count++;
    tl_mark("play_kernel", stream);
    if (g_ctx.after_play_kernel) {
        FB_CUDA(cudaEventRecord(g_ctx.after_play_kernel, stream));
        g_ctx.after_play_kernel = nullptr;
    }
    const size_t tile = F.rows ? (size_t)256 * F.row_words * 4 : 0;  // <= 88 KB at k = 12
    if (tile > 40 * 1024 && !g_ctx.finish_opted) {
        FB_CUDA(cudaFuncSetAttribute(finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 352));
        g_ctx.finish_opted = true;
    }
    finish_kernel<<<(unsigned)((F.n_games + 255) / 256), 256, tile, stream>>>(F);
    rc = launch_check("finish_kernel");
    tl_mark("finish", stream);
    return rc;
}

inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int fb_abi_version(void) { return FB_ABI_VERSION; }
const char* fb_last_error(void) { return t_err.c_str(); }
uint64_t fb_kernel_launch_count(void) { return g_launches.load(); }

int fb_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(FB_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count || device >= MAX_DEVICES)
        return fail(FB_ERR_BAD_ARG, "device %d out of range [0,%d)", device, std::min(count, MAX_DEVICES));
    FB_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_mu);  // the lock and context of `device`, now current
    if (g_ctx.device == device && g_ctx.lut_dev) return FB_OK;
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    g_ctx.sm_count = prop.multiProcessorCount;
    g_ctx.cc_major = prop.major;
    g_ctx.cc_minor = prop.minor;
    FB_CUDA(cudaDeviceGetAttribute(&g_ctx.clock_khz, cudaDevAttrClockRate, device));
    FB_CUDA(cudaDeviceGetAttribute(&g_ctx.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    g_ctx.l2_bytes = prop.l2CacheSize;
    ScoreLut* host_lut = new ScoreLut;
    host_build_lut(*host_lut);
    if (g_ctx.lut_dev) cudaFree(g_ctx.lut_dev);
    FB_CUDA(cudaMalloc(&g_ctx.lut_dev, sizeof(ScoreLut)));
    FB_CUDA(cudaMemcpy(g_ctx.lut_dev, host_lut, sizeof(ScoreLut), cudaMemcpyHostToDevice));
    delete host_lut;
    {   // LCG jump-ahead constants for permute_warp_kernel: A^j and 1 + A + ... + A^(j-1) mod 2^128
        JumpTable jt;
        unsigned __int128 a = 1, gsum = 0;
        for (int j = 0; j <= 32; j++) {
            jt.a_hi[j] = (uint64_t)(a >> 64); jt.a_lo[j] = (uint64_t)a;
            jt.g_hi[j] = (uint64_t)(gsum >> 64); jt.g_lo[j] = (uint64_t)gsum;
            gsum += a;
            a *= (unsigned __int128)PCG_CHEAP_MULT;
        }
        if (g_ctx.jump_dev) cudaFree(g_ctx.jump_dev);
        FB_CUDA(cudaMalloc(&g_ctx.jump_dev, sizeof(JumpTable)));
        FB_CUDA(cudaMemcpy(g_ctx.jump_dev, &jt, sizeof(JumpTable), cudaMemcpyHostToDevice));
    }
    g_ctx.device = device;
    g_ctx.smem_opted = false;
    g_ctx.finish_opted = false;
    return FB_OK;
}

int fb_device_info(int* sm_count, int* clock_khz, int* cc_major, int* cc_minor) {
    FB_REQUIRE_INIT();
    if (sm_count) *sm_count = g_ctx.sm_count;
    if (clock_khz) *clock_khz = g_ctx.clock_khz;
    if (cc_major) *cc_major = g_ctx.cc_major;
    if (cc_minor) *cc_minor = g_ctx.cc_minor;
    return FB_OK;
}

size_t fb_row_stride(int k) {
    return (sizeof(fb_row_header_t) + (size_t)k * sizeof(fb_row_seat_t) + 15u) & ~(size_t)15u;
}

size_t fb_workspace_bytes(int k, uint64_t n_games) {
    if (k < 1 || k > FB_MAX_PLAYERS) return 0;
    // core + H2H block table / offsets tail allowance (callers add the perm buffer)
    return ws_core_bytes(k, n_games) + 4096;
}

int fb_seedseq_generate(const uint32_t* entropy_dev, int n_entropy, uint64_t n_streams, int n_words,
                        uint32_t* out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (n_entropy < 1 || n_words < 1) return fail(FB_ERR_BAD_ARG, "n_entropy and n_words must be >= 1");
    if (n_streams == 0) return FB_OK;
    seedseq_kernel<<<blocks_for(n_streams, 128), 128, 0, (cudaStream_t)stream>>>(entropy_dev, n_entropy,
                                                                                 n_streams, n_words, out_dev);
    return launch_check("seedseq_kernel");
}

int fb_coordinate_seeds(uint32_t purpose, uint64_t root_seed, uint64_t k, uint64_t shuffle_index,
                        uint64_t pair_id, uint64_t order, uint64_t game_index, int vary, uint64_t base,
                        uint64_t n, int as_u32, uint64_t* out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (vary < 0 || vary > 2) return fail(FB_ERR_BAD_ARG, "vary must be 0, 1 or 2");
    if (n == 0) return FB_OK;
    Coord c{purpose, root_seed, k, shuffle_index, pair_id, order, game_index, 0, 0};
    coord_seeds_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(c, vary, base, n, as_u32, out_dev);
    return launch_check("coord_seeds_kernel");
}

int fb_seed_streams(const uint64_t* coords_dev, uint64_t n, uint64_t* state_inc_out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (n == 0) return FB_OK;
    seed_streams_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(coords_dev, n, state_inc_out_dev);
    return launch_check("seed_streams_kernel");
}

int fb_roll_dice(const uint64_t* state_inc_dev, const uint32_t* half_buffer_dev, uint64_t n,
                 const int32_t* n_dice_dev, int n_rolls, uint8_t* faces_out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (n == 0 || n_rolls == 0) return FB_OK;
    roll_dice_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
        state_inc_dev, half_buffer_dev, n, n_dice_dev, n_rolls, faces_out_dev);
    return launch_check("roll_dice_kernel");
}

int fb_default_score(const uint8_t* faces_dev, const int32_t* turn_score_pre_dev,
                     const fb_strategy_t* strategy_dev, uint64_t n, int32_t* out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (n == 0) return FB_OK;
    default_score_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
        g_ctx.lut_dev, faces_dev, turn_score_pre_dev, strategy_dev, n, out_dev);
    return launch_check("default_score_kernel");
}

static int permute_shuffles(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles, int n_strategies,
                            int32_t* perm_out_dev, int32_t* inv_out_dev, uint4* prefix_out_dev, void* stream) {
    FB_REQUIRE_INIT();
    if (n_shuffles < 0 || n_strategies < 1) return fail(FB_ERR_BAD_ARG, "bad shuffle or strategy count");
    if (n_shuffles == 0) return FB_OK;
    const size_t smem = perm_bytes_per_shuffle(n_strategies) * PERM_SHUFFLES;
    if (n_strategies <= 65535 && smem <= (size_t)g_ctx.max_smem_optin - 1024) {
        FB_CUDA(cudaFuncSetAttribute(permute_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        permute_warp_kernel<<<(n_shuffles + PERM_SHUFFLES - 1) / PERM_SHUFFLES, 32, smem,
                              (cudaStream_t)stream>>>(root_seed, k, shuffle0, n_shuffles, n_strategies,
                                                      g_ctx.jump_dev, perm_out_dev, inv_out_dev, prefix_out_dev);
        return launch_check("permute_warp_kernel");
    }
    permute_kernel<<<blocks_for((uint64_t)n_shuffles, 128), 128, 0, (cudaStream_t)stream>>>(
        root_seed, k, shuffle0, n_shuffles, n_strategies, perm_out_dev, inv_out_dev, prefix_out_dev);
    return launch_check("permute_kernel");
}

int fb_permute_shuffles(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles, int n_strategies,
                        int32_t* perm_out_dev, void* stream) {
    return permute_shuffles(root_seed, k, shuffle0, n_shuffles, n_strategies, perm_out_dev, nullptr, nullptr, stream);
}

// ---- matchup groups (matchup.cuh) ------------------------------------------------------------
static size_t matchup_cub_bound(uint64_t n) { return align_up(2 * n + (8u << 20), 256); }

static size_t matchup_scratch_bytes(uint64_t n) {
    return 2 * align_up(n * 8, 256) + 6 * align_up(n * 4, 256) + align_up((n + 1) * 4, 256) + 256 +
           matchup_cub_bound(n);
}

static int run_matchups(const fb_lag_request_t& rq, const int* lags, int n_lags, const uint32_t* header,
                        const int32_t* perm, const int32_t* strategy_ids_dev, uint64_t n_games, int k,
                        int n_tally_ids, cudaStream_t stream) {
    const uint64_t n = n_games;
    if (rq.scratch_bytes < matchup_scratch_bytes(n))
        return fail(FB_ERR_WORKSPACE, "matchup scratch too small: need %zu bytes", matchup_scratch_bytes(n));
    uint8_t* p = static_cast<uint8_t*>(rq.scratch_dev);
    auto take = [&](size_t bytes) {
        uint8_t* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    uint64_t* key_a = reinterpret_cast<uint64_t*>(take(n * 8));
    uint64_t* key_b = reinterpret_cast<uint64_t*>(take(n * 8));
    uint32_t* game_a = reinterpret_cast<uint32_t*>(take(n * 4));
    uint32_t* game_b = reinterpret_cast<uint32_t*>(take(n * 4));
    MatchupParams M{};
    M.seg_flag = reinterpret_cast<uint32_t*>(take(n * 4));
    M.seg_id1 = reinterpret_cast<uint32_t*>(take(n * 4));
    M.elig = reinterpret_cast<uint32_t*>(take(n * 4));
    M.slot = reinterpret_cast<uint32_t*>(take(n * 4));
    M.seg_first = reinterpret_cast<uint32_t*>(take((n + 1) * 4));
    M.status = reinterpret_cast<uint32_t*>(take(256));
    void* cub_temp = p;
    const size_t cub_have = matchup_cub_bound(n);
    M.header = header;
    M.perm = perm;
    M.strategy_ids = strategy_ids_dev;
    M.n_games = (uint32_t)n;
    M.k = k;
    M.n_lags = n_lags;
    for (int z = 0; z < n_lags; z++) M.lags[z] = lags[z];
    M.min_obs = (uint32_t)rq.matchup_min_observations;
    int id_bits = 1;  // every id is < n_tally_ids
    while (id_bits < 32 && (1ll << id_bits) < (long long)n_tally_ids) id_bits++;
    M.id_bits = id_bits;
    const int key_bits = k <= 2 ? 2 * id_bits : 64;
    M.capacity = rq.matchup_capacity;
    M.participants = rq.matchup_participants_dev;
    M.count = rq.matchup_count_dev;
    M.stats = reinterpret_cast<unsigned long long*>(rq.matchup_stats_dev);
    FB_CUDA(cudaMemsetAsync(M.status, 0, 256, stream));
    FB_CUDA(cudaMemsetAsync(M.stats, 0, (size_t)rq.matchup_capacity * n_lags * FB_MATCHUP_LAG_WIDTH * 8, stream));
    const unsigned blocks = blocks_for(n, 256);
    matchup_key_kernel<<<blocks, 256, 0, stream>>>(M, key_a, game_a);
    int rc = launch_check("matchup_key_kernel");
    if (rc) return rc;
    cub::DoubleBuffer<uint64_t> keys(key_a, key_b);
    cub::DoubleBuffer<uint32_t> vals(game_a, game_b);
    size_t need = 0;
    FB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, vals, (int)n, 0, key_bits, stream));
    if (need > cub_have) return fail(FB_ERR_WORKSPACE, "radix sort needs %zu temporary bytes, %zu reserved", need, cub_have);
    FB_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, need, keys, vals, (int)n, 0, key_bits, stream));
    g_launches.fetch_add(1);
    M.key = keys.Current();
    M.game = vals.Current();
    matchup_flag_kernel<<<blocks, 256, 0, stream>>>(M);
    rc = launch_check("matchup_flag_kernel");
    if (rc) return rc;
    FB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, need, M.seg_flag, M.seg_id1, (int)n, stream));
    if (need > cub_have) return fail(FB_ERR_WORKSPACE, "scan needs %zu temporary bytes, %zu reserved", need, cub_have);
    FB_CUDA(cub::DeviceScan::InclusiveSum(cub_temp, need, M.seg_flag, M.seg_id1, (int)n, stream));
    matchup_first_kernel<<<blocks, 256, 0, stream>>>(M);
    rc = launch_check("matchup_first_kernel");
    if (rc) return rc;
    matchup_eligible_kernel<<<blocks, 256, 0, stream>>>(M);
    rc = launch_check("matchup_eligible_kernel");
    if (rc) return rc;
    FB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, M.elig, M.slot, (int)n, stream));
    if (need > cub_have) return fail(FB_ERR_WORKSPACE, "scan needs %zu temporary bytes, %zu reserved", need, cub_have);
    FB_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp, need, M.elig, M.slot, (int)n, stream));
    matchup_stats_kernel<<<blocks, 256, 0, stream>>>(M);
    rc = launch_check("matchup_stats_kernel");
    if (rc) return rc;
    uint32_t status[2] = {0, 0};
    FB_CUDA(cudaMemcpyAsync(status, M.status, sizeof(status), cudaMemcpyDeviceToHost, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    if (status[0]) return fail(FB_ERR_INTERNAL, "matchup key collision: two different matchups share a 64-bit key");
    if (status[1] > rq.matchup_capacity)
        return fail(FB_ERR_WORKSPACE, "matchup buffers hold %llu groups, the launch has %u",
                    (unsigned long long)rq.matchup_capacity, status[1]);
    *rq.n_matchups_host = (int64_t)status[1];
    return FB_OK;
}

// A tournament launch has two phases that fb_play_tournament_cells runs on different streams:
// PREPARE (permutations, seat seeding: touches only the workspace) and PLAY (play, finish, tally
// gather: reads the workspace, accumulates into the caller's tallies / totals / rows).
enum { PHASE_PREPARE = 1, PHASE_PLAY = 2, PHASE_ALL = 3 };

static int play_tournament_impl(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                                const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                                int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                                const uint64_t* override_shuffle_dev, const uint32_t* override_game_dev,
                                const int32_t* override_max_rounds_dev, int n_overrides, int shuffles_per_slot,
                                int64_t* tallies_dev, int64_t* totals_dev, void* rows_dev, int want_game_seeds,
                                void* workspace_dev, size_t workspace_bytes, void* stream_v,
                                uint32_t ordinal_base, int64_t* seat_tallies_dev = nullptr,
                                const fb_lag_request_t* lag = nullptr, bool ids_trusted = false,
                                int phase = PHASE_ALL) {
    FB_REQUIRE_INIT();
    if (seat_tallies_dev && !tallies_dev) return fail(FB_ERR_BAD_ARG, "seat tallies need tallies_dev");
    LagParams L{};
    if (lag) {
        if (!tallies_dev) return fail(FB_ERR_BAD_ARG, "lag statistics / first-seen ordinals need tallies_dev");
        if (lag->all_player_dev && shuffles_per_slot <= 0)
            return fail(FB_ERR_BAD_ARG, "all-player statistics are per deterministic batch: shuffles_per_slot > 0");
        const bool wants_lags = lag->strategy_stats_dev || lag->strategy_edges_dev || lag->matchup_min_observations != 0;
        if (lag->n_lags < 0 || lag->n_lags > FB_MAX_LAGS || (lag->n_lags > 0 && !lag->lags) ||
            (lag->n_lags == 0 && wants_lags))
            return fail(FB_ERR_BAD_ARG, "lag statistics need 1..%d lags", FB_MAX_LAGS);
        if ((lag->strategy_stats_dev != nullptr) != (lag->strategy_edges_dev != nullptr))
            return fail(FB_ERR_BAD_ARG, "strategy lag statistics need both the stats and the edges buffer");
        if (lag->matchup_min_observations < 0 ||
            (lag->matchup_min_observations > 0 &&
             (!lag->matchup_participants_dev || !lag->matchup_count_dev || !lag->matchup_stats_dev ||
              !lag->scratch_dev || !lag->n_matchups_host)))
            return fail(FB_ERR_BAD_ARG, "matchup lag statistics need all matchup buffers, scratch and the count pointer");
        for (int z = 0; z < lag->n_lags; z++) {
            if (lag->lags[z] < 1 || lag->lags[z] > FB_MAX_LAG)
                return fail(FB_ERR_BAD_ARG, "lag %d outside [1,%d]", lag->lags[z], FB_MAX_LAG);
            for (int y = 0; y < z; y++)
                if (lag->lags[y] == lag->lags[z]) return fail(FB_ERR_BAD_ARG, "duplicate lag %d", lag->lags[z]);
            L.lags[z] = lag->lags[z];
            L.max_lag = std::max(L.max_lag, lag->lags[z]);
        }
        L.n_lags = lag->n_lags;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (k < 1 || k > FB_MAX_PLAYERS) return fail(FB_ERR_BAD_ARG, "k=%d outside [1,%d]", k, FB_MAX_PLAYERS);
    if (n_strategies < k || n_strategies % k != 0)
        return fail(FB_ERR_BAD_ARG, "n_players must divide %d", n_strategies);  // run_tournament.py:274-275
    if (n_shuffles < 0 || shuffles_per_slot < 0 || n_overrides < 0) return fail(FB_ERR_BAD_ARG, "negative count");
    if (max_rounds > FB_MAX_ROUNDS)
        return fail(FB_ERR_BAD_ARG, "max_rounds=%d above %d (n_rounds is an int16 column)", max_rounds, FB_MAX_ROUNDS);
    if (tallies_dev) {
        if (n_tally_ids < 1) return fail(FB_ERR_BAD_ARG, "n_tally_ids=%d must be >= 1", n_tally_ids);
        if (!strategy_ids_dev && n_tally_ids < n_strategies)
            return fail(FB_ERR_BAD_ARG, "n_tally_ids=%d below n_strategies=%d with implicit ids", n_tally_ids, n_strategies);
        if (lag && lag->first_seen_dev && (uint64_t)n_shuffles * (uint64_t)n_strategies > 0x7fffffffull)
            return fail(FB_ERR_BAD_ARG, "first-seen ordinals need n_shuffles * n_strategies < 2^31");
    }
    if (n_shuffles == 0) {  // nothing played: the optional outputs still get their "empty" values
        if (lag && lag->first_seen_dev)
            FB_CUDA(cudaMemsetAsync(lag->first_seen_dev, 0xff, (size_t)n_tally_ids * 4 * sizeof(uint32_t), stream));
        if (lag && lag->n_matchups_host) *lag->n_matchups_host = 0;
        return FB_OK;
    }
    const uint32_t gps = (uint32_t)(n_strategies / k);
    const uint64_t n_games = (uint64_t)n_shuffles * gps;
    if (n_games * (uint64_t)k > 0xfffffff0ull || n_games >= 0x7ff00000ull)
        return fail(FB_ERR_BAD_ARG, "more than 2^32 seats or 2^31 games in one launch");
    Workspace w;
    const size_t perm_bytes = align_up((size_t)n_shuffles * n_strategies * 4, 256);
    if (!carve(workspace_dev, workspace_bytes, k, n_games, w) || w.extra_bytes < 2 * perm_bytes)
        return fail(FB_ERR_WORKSPACE, "workspace too small: need %zu bytes",
                    ws_core_bytes(k, n_games) + 2 * perm_bytes);
    int32_t* perm = reinterpret_cast<int32_t*>(w.extra);
    // (a cell prepared ahead does not know yet whether its play phase will want tallies)
    int32_t* inv = (tallies_dev || phase != PHASE_ALL) ? reinterpret_cast<int32_t*>(w.extra + perm_bytes) : nullptr;
    uint4* prefix = reinterpret_cast<uint4*>(w.prefix);
    int rc = FB_OK;
    int32_t* limits = n_overrides > 0 ? w.limits : nullptr;
    const int tl_lane = (g_ctx.s_prep && stream == g_ctx.s_prep) ? 1 : 0;
    if (phase & PHASE_PREPARE) {
    tl_mark("prepare_begin", stream, tl_lane);
    rc = permute_shuffles(root_seed, k, shuffle0, n_shuffles, n_strategies, perm, inv, prefix, stream);
    if (rc) return rc;
    tl_mark("permute", stream, tl_lane);
    FB_CUDA(cudaMemsetAsync(w.counter, 0, 4 * sizeof(unsigned int), stream));
    if (tallies_dev && strategy_ids_dev && !ids_trusted) {
        // explicit ids address the tally buffers: one pass over the table, one flag back
        check_ids_kernel<<<blocks_for((uint64_t)n_strategies, 256), 256, 0, stream>>>(strategy_ids_dev, n_strategies,
                                                                                   n_tally_ids, w.counter + 2);
        rc = launch_check("check_ids_kernel");
        if (rc) return rc;
        unsigned int bad = 0;
        FB_CUDA(cudaMemcpyAsync(&bad, w.counter + 2, sizeof bad, cudaMemcpyDeviceToHost, stream));
        FB_CUDA(cudaStreamSynchronize(stream));
        if (bad) return fail(FB_ERR_BAD_ARG, "a strategy id lies outside [0, n_tally_ids=%d)", n_tally_ids);
    }
    seed_tournament_kernel<<<blocks_for(n_games * k, 256), 256, 0, stream>>>(
        root_seed, k, shuffle0, gps, n_games, perm, n_strategies, target_score, max_rounds,
        override_shuffle_dev, override_game_dev, override_max_rounds_dev, n_overrides, want_game_seeds,
        strategies_dev, w.seats, w.game_seed, limits, w.header, w.long_list, w.counter, prefix);
    rc = launch_check("seed_tournament_kernel");
    if (rc) return rc;
    tl_mark("seed", stream, tl_lane);
    }  // PHASE_PREPARE
    if (!(phase & PHASE_PLAY)) return FB_OK;
    PlayParams P{};
    P.seats = w.seats;
    P.header = w.header;
    P.limits = limits;
    P.target_score = target_score;
    P.max_rounds = max_rounds;
    P.n_games = (uint32_t)n_games;
    P.k = k;
    P.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    P.counter = w.counter;
    P.long_list = w.long_list;
    FinishParams F{};
    F.seats = w.seats;
    F.header = w.header;
    F.strategy_ids = strategy_ids_dev;
    F.ids_mode = strategy_ids_dev ? 1 : 0;
    F.perm = perm;
    F.game_seed = want_game_seeds ? w.game_seed : nullptr;
    F.n_games = (uint32_t)n_games;
    F.k = k;
    F.games_per_slot = shuffles_per_slot > 0 ? (uint32_t)shuffles_per_slot * gps : 0u;
    F.n_tally_ids = n_tally_ids;
    F.tallies = reinterpret_cast<unsigned long long*>(tallies_dev);
    F.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    F.rows = reinterpret_cast<uint32_t*>(rows_dev);
    F.row_words = (int)(fb_row_stride(k) / 4);
    F.ordinal_base = ordinal_base;
    if (!F.tallies) return launch_play(P, F, stream);
    // tallies: exposures and winner metrics by gather after the finish pass
    F.mark_winner = 1;
    rc = launch_play(P, F, stream);
    if (rc) return rc;
    GatherParams G{};
    G.seats = w.seats;
    G.header = w.header;
    G.inv = inv;
    G.strategy_ids = strategy_ids_dev;
    G.n_strategies = n_strategies;
    G.n_tally_ids = n_tally_ids;
    G.n_shuffles = n_shuffles;
    G.k = k;
    G.gps = gps;
    G.chunk = shuffles_per_slot > 0 ? shuffles_per_slot : 43;
    G.slotted = shuffles_per_slot > 0;
    G.tallies = F.tallies;
    G.seat_tallies = reinterpret_cast<unsigned long long*>(seat_tallies_dev);
    G.first_seen = lag ? lag->first_seen_dev : nullptr;
    if (G.first_seen) FB_CUDA(cudaMemsetAsync(G.first_seen, 0xff, (size_t)n_tally_ids * 4 * sizeof(uint32_t), stream));
    const int n_chunks = (n_shuffles + G.chunk - 1) / G.chunk;
    tally_gather_kernel<<<dim3(blocks_for((uint64_t)n_strategies, 128), (unsigned)n_chunks), 128, 0, stream>>>(G);
    rc = launch_check("tally_gather_kernel");
    tl_mark("gather", stream, tl_lane);
    if (rc || !lag) return rc;
    if (lag->all_player_dev) {
        AllPlayerParams A{};
        A.seats = w.seats;
        A.header = w.header;
        A.inv = inv;
        A.strategy_ids = strategy_ids_dev;
        A.n_strategies = n_strategies;
        A.n_tally_ids = n_tally_ids;
        A.n_shuffles = n_shuffles;
        A.k = k;
        A.gps = gps;
        A.per_slot = shuffles_per_slot;
        A.out = reinterpret_cast<long long*>(lag->all_player_dev);
        const int a_slots = (n_shuffles + shuffles_per_slot - 1) / shuffles_per_slot;
        allplayer_gather_kernel<<<dim3(blocks_for((uint64_t)n_strategies, 128), (unsigned)a_slots), 128, 0, stream>>>(A);
        rc = launch_check("allplayer_gather_kernel");
        if (rc) return rc;
    }
    if (lag->strategy_stats_dev) {
        L.header = w.header;
        L.inv = inv;
        L.n_strategies = n_strategies;
        L.n_shuffles = n_shuffles;
        L.k = k;
        L.gps = gps;
        L.chunk = 43;
        L.stats = reinterpret_cast<unsigned long long*>(lag->strategy_stats_dev);
        L.edges = lag->strategy_edges_dev;
        const unsigned sblocks = blocks_for((uint64_t)n_strategies, 128);
        lag_gather_kernel<<<dim3(sblocks, (unsigned)((n_shuffles + L.chunk - 1) / L.chunk), (unsigned)L.n_lags), 128,
                            0, stream>>>(L);
        rc = launch_check("lag_gather_kernel");
        if (rc) return rc;
        lag_edges_kernel<<<dim3(sblocks, (unsigned)std::min(L.max_lag, n_shuffles)), 128, 0, stream>>>(L);
        rc = launch_check("lag_edges_kernel");
        if (rc) return rc;
    }
    if (lag->matchup_min_observations > 0)
        return run_matchups(*lag, L.lags, L.n_lags, w.header, perm, strategy_ids_dev, n_games, k, n_tally_ids, stream);
    return FB_OK;
}

int fb_play_tournament(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                       const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                       int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                       const uint64_t* override_shuffle_dev, const uint32_t* override_game_dev,
                       const int32_t* override_max_rounds_dev, int n_overrides, int shuffles_per_slot,
                       int64_t* tallies_dev, int64_t* totals_dev, void* rows_dev, int want_game_seeds,
                       void* workspace_dev, size_t workspace_bytes, void* stream_v) {
    return play_tournament_impl(root_seed, k, shuffle0, n_shuffles, strategies_dev, strategy_ids_dev, n_strategies,
                                n_tally_ids, target_score, max_rounds, override_shuffle_dev, override_game_dev,
                                override_max_rounds_dev, n_overrides, shuffles_per_slot, tallies_dev, totals_dev,
                                rows_dev, want_game_seeds, workspace_dev, workspace_bytes, stream_v, 0u);
}

int fb_play_tournament_seats(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                             const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                             int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                             const uint64_t* override_shuffle_dev, const uint32_t* override_game_dev,
                             const int32_t* override_max_rounds_dev, int n_overrides, int shuffles_per_slot,
                             int64_t* tallies_dev, int64_t* totals_dev, void* rows_dev, int want_game_seeds,
                             int64_t* seat_tallies_dev, void* workspace_dev, size_t workspace_bytes,
                             void* stream_v) {
    return play_tournament_impl(root_seed, k, shuffle0, n_shuffles, strategies_dev, strategy_ids_dev, n_strategies,
                                n_tally_ids, target_score, max_rounds, override_shuffle_dev, override_game_dev,
                                override_max_rounds_dev, n_overrides, shuffles_per_slot, tallies_dev, totals_dev,
                                rows_dev, want_game_seeds, workspace_dev, workspace_bytes, stream_v, 0u,
                                seat_tallies_dev);
}

int fb_play_tournament_lags(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                            const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                            int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                            const uint64_t* override_shuffle_dev, const uint32_t* override_game_dev,
                            const int32_t* override_max_rounds_dev, int n_overrides, int shuffles_per_slot,
                            int64_t* tallies_dev, int64_t* totals_dev, void* rows_dev, int want_game_seeds,
                            int64_t* seat_tallies_dev, const fb_lag_request_t* lag, void* workspace_dev,
                            size_t workspace_bytes, void* stream_v) {
    return play_tournament_impl(root_seed, k, shuffle0, n_shuffles, strategies_dev, strategy_ids_dev, n_strategies,
                                n_tally_ids, target_score, max_rounds, override_shuffle_dev, override_game_dev,
                                override_max_rounds_dev, n_overrides, shuffles_per_slot, tallies_dev, totals_dev,
                                rows_dev, want_game_seeds, workspace_dev, workspace_bytes, stream_v, 0u,
                                seat_tallies_dev, lag);
}

static size_t cell_slot_bytes(int k, uint64_t n_games, int n_shuffles, int n_strategies) {
    return align_up(ws_core_bytes(k, n_games) + 2 * align_up((size_t)n_shuffles * n_strategies * 4, 256), 256);
}

size_t fb_cells_workspace_bytes(const fb_cell_t* cells, int n_cells, int n_strategies) {
    size_t slot = 0;
    for (int i = 0; cells && i < n_cells; i++) {
        const fb_cell_t& c = cells[i];
        if (c.k < 1 || c.k > FB_MAX_PLAYERS || c.n_shuffles < 0 || n_strategies < c.k) return 0;
        slot = std::max(slot, cell_slot_bytes(c.k, (uint64_t)c.n_shuffles * (uint64_t)(n_strategies / c.k),
                                              c.n_shuffles, n_strategies));
    }
    return 2 * slot + 512;
}

int fb_play_tournament_cells(const fb_cell_t* cells, int n_cells, int n_ahead,
                             const fb_strategy_t* strategies_dev, const int32_t* strategy_ids_dev,
                             int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                             int shuffles_per_slot, void* workspace_dev, size_t workspace_bytes, void* stream_v) {
    FB_REQUIRE_INIT();
    if (n_cells < 0 || n_ahead < 0 || n_ahead > 1 || (n_cells + n_ahead > 0 && !cells))
        return fail(FB_ERR_BAD_ARG, "bad cell list (n_cells >= 0, n_ahead 0 or 1)");
    const int n_all = n_cells + n_ahead;
    if (n_all == 0) return FB_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_ctx.s_prep) {
        // Default priority, the same as the caller's stream.  Measured (scripts/timeline_step.py,
        // profiles/r02_timeline.md): the block scheduler places a kernel's CTAs only when no CTA
        // of a more urgent kernel is pending, so an urgent preparation stream holds the finish
        // pass back until the permutation kernel's second wave is placed (window 2.24 ms), a less
        // urgent one is held back by finish and gather (window 2.38 ms); at equal priority the
        // finish pass goes first and the permutation kernel runs beside the tally pass (2.03 ms).
        // FB_PREP_PRIO overrides (tuning knob).
        int prio = 0;
        if (const char* env = getenv("FB_PREP_PRIO")) prio = atoi(env);
        FB_CUDA(cudaStreamCreateWithPriority(&g_ctx.s_prep, cudaStreamNonBlocking, prio));
        FB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_cells_entry, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            FB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_cell_prepared[i], cudaEventDisableTiming));
            FB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_play_kernel_done[i], cudaEventDisableTiming));
        }
    }
    // two workspace slots: cell i is prepared in one while cell i-1 is played out of the other
    uint8_t* base = static_cast<uint8_t*>(workspace_dev);
    const size_t shift = (256 - (reinterpret_cast<uintptr_t>(base) & 255)) & 255;
    if (!base || workspace_bytes < shift + 512) return fail(FB_ERR_WORKSPACE, "workspace too small");
    const size_t slot_bytes = ((workspace_bytes - shift) / 2) & ~(size_t)255;
    uint8_t* slot_ptr[2] = {base + shift, base + shift + slot_bytes};
    for (int i = 0; i < n_all; i++) {
        const fb_cell_t& c = cells[i];
        if (c.k < 1 || c.k > FB_MAX_PLAYERS || c.n_shuffles < 0 || n_strategies < c.k || n_strategies % c.k)
            return fail(FB_ERR_BAD_ARG, "cell %d: bad k / shuffle count", i);
        const size_t need = cell_slot_bytes(c.k, (uint64_t)c.n_shuffles * (uint64_t)(n_strategies / c.k),
                                            c.n_shuffles, n_strategies);
        if (need > slot_bytes)
            return fail(FB_ERR_WORKSPACE, "workspace too small: cell %d needs %zu bytes per slot, %zu given "
                        "(fb_cells_workspace_bytes)", i, need, slot_bytes);
    }
    auto run = [&](int i, int slot, int phase, cudaStream_t s) {
        const fb_cell_t& c = cells[i];
        return play_tournament_impl(c.root_seed, c.k, c.shuffle0, c.n_shuffles, strategies_dev, strategy_ids_dev,
                                    n_strategies, n_tally_ids, target_score, max_rounds, nullptr, nullptr, nullptr, 0,
                                    shuffles_per_slot, c.tallies_dev, c.totals_dev, nullptr, 0, slot_ptr[slot],
                                    slot_bytes, s, 0u, nullptr, nullptr, /*ids_trusted=*/i > 0 || phase == PHASE_PLAY,
                                    phase);
    };
    // Was the first cell prepared by the previous call (its n_ahead cell)?
    Ctx::Ahead& ah = g_ctx.ahead;
    int slot0 = 0;
    bool first_ready = false;
    if (ah.valid && n_cells > 0) {
        const fb_cell_t& c = cells[0];
        first_ready = ah.root_seed == c.root_seed && ah.shuffle0 == c.shuffle0 && ah.k == c.k &&
                      ah.n_shuffles == c.n_shuffles && ah.n_strategies == n_strategies &&
                      ah.target_score == target_score && ah.max_rounds == max_rounds &&
                      ah.strategies == strategies_dev && ah.workspace == workspace_dev &&
                      ah.workspace_bytes == workspace_bytes;
        if (first_ready) slot0 = ah.slot;
    }
    ah.valid = false;
    int rc = FB_OK;
    // everything already queued on the caller's stream (table upload, zeroing of the tallies)
    // comes before the first preparation
    FB_CUDA(cudaEventRecord(g_ctx.ev_cells_entry, stream));
    FB_CUDA(cudaStreamWaitEvent(g_ctx.s_prep, g_ctx.ev_cells_entry, 0));
    // (explicit ids are validated in the PREPARE phase of the first cell; a cell prepared ahead was
    // validated by the call that prepared it)
    if (!first_ready) {
        rc = run(0, slot0, PHASE_PREPARE, g_ctx.s_prep);
        if (rc) return rc;
        FB_CUDA(cudaEventRecord(g_ctx.ev_cell_prepared[slot0], g_ctx.s_prep));
    }
    // play_kernel is persistent and fills every SM (all registers, all shared memory): nothing runs
    // beside it, and a small kernel that gets onto the SMs first keeps its CTAs out.  So the
    // preparation of cell i+1 is released by the END of cell i's play_kernel and runs beside cell
    // i's finish and tally passes (which only read slot i).  Measured gain: small -- sharing SMs,
    // the permutation kernel (209 KB of shared memory per SM) and the gathers (which want the L1)
    // each take about as long as both in sequence; what the second stream buys is the seed pass
    // under the tail of the tally pass (profiles/r02_timeline.md).  (Slot reuse: the passes of
    // cell i-1, which read the slot cell i+1 is prepared in, precede play_kernel(i) on `stream`.)
    for (int i = 0; i < n_cells; i++) {
        const int slot = (slot0 + i) & 1;
        FB_CUDA(cudaStreamWaitEvent(stream, g_ctx.ev_cell_prepared[slot], 0));
        const bool more = i + 1 < n_all;
        if (more) g_ctx.after_play_kernel = g_ctx.ev_play_kernel_done[slot];
        rc = run(i, slot, PHASE_PLAY, stream);
        const bool recorded = more && g_ctx.after_play_kernel == nullptr;
        g_ctx.after_play_kernel = nullptr;
        if (rc) return rc;
        if (more) {
            const int nslot = slot ^ 1;
            if (recorded) {
                FB_CUDA(cudaStreamWaitEvent(g_ctx.s_prep, g_ctx.ev_play_kernel_done[slot], 0));
            } else {  // nothing was played (empty cell): order behind whatever the stream holds
                FB_CUDA(cudaEventRecord(g_ctx.ev_play_kernel_done[slot], stream));
                FB_CUDA(cudaStreamWaitEvent(g_ctx.s_prep, g_ctx.ev_play_kernel_done[slot], 0));
            }
            rc = run(i + 1, nslot, PHASE_PREPARE, g_ctx.s_prep);
            if (rc) return rc;
            FB_CUDA(cudaEventRecord(g_ctx.ev_cell_prepared[nslot], g_ctx.s_prep));
        }
    }
    if (n_ahead == 1) {
        const fb_cell_t& c = cells[n_cells];
        ah.valid = true;
        ah.root_seed = c.root_seed; ah.shuffle0 = c.shuffle0; ah.k = c.k; ah.n_shuffles = c.n_shuffles;
        ah.n_strategies = n_strategies; ah.slot = (slot0 + n_cells) & 1;
        ah.target_score = target_score; ah.max_rounds = max_rounds;
        ah.strategies = strategies_dev; ah.workspace = workspace_dev; ah.workspace_bytes = workspace_bytes;
    }
    return FB_OK;
}

size_t fb_matchup_scratch_bytes(uint64_t n_games) { return matchup_scratch_bytes(n_games); }

int fb_play_h2h(uint64_t root_seed, int n_blocks, const uint64_t* pair_id_dev, const uint8_t* order_dev,
                const fb_strategy_t* seat1_dev, const fb_strategy_t* seat2_dev,
                const uint32_t* attempt0_dev, const uint32_t* n_attempts_dev, uint64_t total_attempts,
                int32_t target_score, int32_t max_rounds, uint8_t* outcome_out_dev, void* rows_dev,
                int64_t* totals_dev, void* workspace_dev, size_t workspace_bytes, void* stream_v) {
    FB_REQUIRE_INIT();
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (n_blocks < 0) return fail(FB_ERR_BAD_ARG, "negative block count");
    if (max_rounds > FB_MAX_ROUNDS)
        return fail(FB_ERR_BAD_ARG, "max_rounds=%d above %d (n_rounds is an int16 column)", max_rounds, FB_MAX_ROUNDS);
    if (n_blocks == 0 || total_attempts == 0) return FB_OK;
    if (total_attempts * 2 > 0xfffffff0ull) return fail(FB_ERR_BAD_ARG, "more than 2^31 attempts in one launch");
    Workspace w;
    const size_t off_bytes = align_up((size_t)(n_blocks + 1) * 8, 256);
    const size_t pre_bytes = align_up((size_t)n_blocks * 16, 256);
    if (!carve(workspace_dev, workspace_bytes, 2, total_attempts, w) || w.extra_bytes < off_bytes + pre_bytes)
        return fail(FB_ERR_WORKSPACE, "workspace too small: need %zu bytes",
                    ws_core_bytes(2, total_attempts) + off_bytes + pre_bytes);
    uint64_t* offsets = reinterpret_cast<uint64_t*>(w.extra);
    h2h_offsets_kernel<<<1, 1024, 0, stream>>>(n_attempts_dev, n_blocks, offsets);
    int rc = launch_check("h2h_offsets_kernel");
    if (rc) return rc;
    const int want_seeds = rows_dev != nullptr;
    FB_CUDA(cudaMemsetAsync(w.counter, 0, 2 * sizeof(unsigned int), stream));
    uint4* prefix = reinterpret_cast<uint4*>(w.extra + off_bytes);
    h2h_prefix_kernel<<<blocks_for((uint64_t)n_blocks, 128), 128, 0, stream>>>(root_seed, n_blocks, pair_id_dev,
                                                                               order_dev, prefix);
    rc = launch_check("h2h_prefix_kernel");
    if (rc) return rc;
    seed_h2h_kernel<<<blocks_for(total_attempts * 2, 256), 256, 0, stream>>>(
        root_seed, n_blocks, pair_id_dev, order_dev, seat1_dev, seat2_dev, attempt0_dev, offsets,
        total_attempts, want_seeds, w.seats, w.game_seed, w.header, w.long_list, w.counter, prefix);
    rc = launch_check("seed_h2h_kernel");
    if (rc) return rc;
    PlayParams P{};
    P.seats = w.seats;
    P.header = w.header;
    P.target_score = target_score;
    P.max_rounds = max_rounds;
    P.n_games = (uint32_t)total_attempts;
    P.k = 2;
    P.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    P.counter = w.counter;
    P.long_list = w.long_list;
    FinishParams F{};
    F.seats = w.seats;
    F.header = w.header;
    F.ids_mode = 2;
    F.game_seed = want_seeds ? w.game_seed : nullptr;
    F.n_games = (uint32_t)total_attempts;
    F.k = 2;
    F.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    F.rows = reinterpret_cast<uint32_t*>(rows_dev);
    F.row_words = (int)(fb_row_stride(2) / 4);
    F.outcome = outcome_out_dev;
    return launch_play(P, F, stream);
}

int fb_h2h_resolve(int n_blocks, const uint32_t* n_attempts_dev, const uint8_t* outcome_dev,
                   const int32_t* n_completed_required_dev, int32_t* progress_dev, void* stream_v) {
    FB_REQUIRE_INIT();
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (n_blocks <= 0) return n_blocks == 0 ? FB_OK : fail(FB_ERR_BAD_ARG, "negative block count");
    uint64_t* offsets = nullptr;
    FB_CUDA(cudaMallocAsync(&offsets, (size_t)(n_blocks + 1) * 8, stream));
    h2h_offsets_kernel<<<1, 1024, 0, stream>>>(n_attempts_dev, n_blocks, offsets);
    int rc = launch_check("h2h_offsets_kernel");
    if (!rc) {
        h2h_resolve_kernel<<<blocks_for((uint64_t)n_blocks * 32, 128), 128, 0, stream>>>(
            n_blocks, n_attempts_dev, offsets, outcome_dev, n_completed_required_dev, progress_dev);
        rc = launch_check("h2h_resolve_kernel");
    }
    cudaFreeAsync(offsets, stream);
    return rc;
}

int fb_play_games(const uint64_t* coords_dev, uint64_t n_games, int k,
                  const fb_strategy_t* seat_strategies_dev, const int32_t* seat_strategy_ids_dev,
                  const int32_t* target_score_dev, int32_t target_score, const int32_t* max_rounds_dev,
                  int32_t max_rounds, void* rows_dev, int64_t* totals_dev, void* workspace_dev,
                  size_t workspace_bytes, void* stream_v) {
    FB_REQUIRE_INIT();
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (k < 1 || k > FB_MAX_PLAYERS) return fail(FB_ERR_BAD_ARG, "k=%d outside [1,%d]", k, FB_MAX_PLAYERS);
    if (max_rounds > FB_MAX_ROUNDS)
        return fail(FB_ERR_BAD_ARG, "max_rounds=%d above %d (n_rounds is an int16 column)", max_rounds, FB_MAX_ROUNDS);
    if (n_games == 0) return FB_OK;
    if (n_games * (uint64_t)k > 0xfffffff0ull || n_games >= 0x7ff00000ull)
        return fail(FB_ERR_BAD_ARG, "more than 2^32 seats or 2^31 games in one launch");
    Workspace w;
    if (!carve(workspace_dev, workspace_bytes, k, n_games, w))
        return fail(FB_ERR_WORKSPACE, "workspace too small: need %zu bytes", ws_core_bytes(k, n_games));
    FB_CUDA(cudaMemsetAsync(w.counter, 0, 2 * sizeof(unsigned int), stream));
    seed_explicit_kernel<<<blocks_for(n_games * k, 256), 256, 0, stream>>>(
        coords_dev, n_games, k, seat_strategies_dev, w.seats, w.header, w.long_list, w.counter);
    int rc = launch_check("seed_explicit_kernel");
    if (rc) return rc;
    int32_t* limits = nullptr;
    if (target_score_dev || max_rounds_dev) {
        limits = w.limits;
        pack_limits_kernel<<<blocks_for(n_games, 256), 256, 0, stream>>>(target_score_dev, target_score,
                                                                         max_rounds_dev, max_rounds, n_games, limits);
        rc = launch_check("pack_limits_kernel");
        if (rc) return rc;
    }
    PlayParams P{};
    P.seats = w.seats;
    P.header = w.header;
    P.limits = limits;
    P.target_score = target_score;
    P.max_rounds = max_rounds;
    P.n_games = (uint32_t)n_games;
    P.k = k;
    P.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    P.counter = w.counter;
    P.long_list = w.long_list;
    FinishParams F{};
    F.seats = w.seats;
    F.header = w.header;
    F.strategy_ids = seat_strategy_ids_dev;
    F.ids_mode = seat_strategy_ids_dev ? 1 : 2;
    F.n_games = (uint32_t)n_games;
    F.k = k;
    F.totals = reinterpret_cast<unsigned long long*>(totals_dev);
    F.rows = reinterpret_cast<uint32_t*>(rows_dev);
    F.row_words = (int)(fb_row_stride(k) / 4);
    return launch_play(P, F, stream);
}

int fb_run_tournament_host(uint64_t root_seed, int k, uint64_t shuffle0, int n_shuffles,
                           const fb_strategy_t* strategies_host, const int32_t* strategy_ids_host,
                           int n_strategies, int n_tally_ids, int32_t target_score, int32_t max_rounds,
                           int shuffles_per_slot, int64_t* tallies_host, int64_t* totals_host,
                           void* rows_host, int want_game_seeds) {
    FB_REQUIRE_INIT();
    if (k < 1 || k > FB_MAX_PLAYERS || n_strategies < k || n_strategies % k != 0 || n_shuffles < 1)
        return fail(FB_ERR_BAD_ARG, "bad k / strategy / shuffle count");
    if (!strategies_host) return fail(FB_ERR_BAD_ARG, "strategies_host is required");
    if (max_rounds > FB_MAX_ROUNDS)
        return fail(FB_ERR_BAD_ARG, "max_rounds=%d above %d (n_rounds is an int16 column)", max_rounds, FB_MAX_ROUNDS);
    if (tallies_host) {
        if (n_tally_ids < 1 || (!strategy_ids_host && n_tally_ids < n_strategies))
            return fail(FB_ERR_BAD_ARG, "n_tally_ids=%d does not cover the ids in use", n_tally_ids);
        for (int i = 0; strategy_ids_host && i < n_strategies; i++)
            if (strategy_ids_host[i] < 0 || strategy_ids_host[i] >= n_tally_ids)
                return fail(FB_ERR_BAD_ARG, "strategy id %d (entry %d) outside [0, n_tally_ids=%d)",
                            strategy_ids_host[i], i, n_tally_ids);
    }
    const uint64_t gps = (uint64_t)(n_strategies / k);
    const int n_slots = shuffles_per_slot > 0 ? (n_shuffles + shuffles_per_slot - 1) / shuffles_per_slot : 1;
    // Rows mode streams the rows to the host while the next part of the range is being played:
    // the range is cut into four chunks of whole tally slots, two device row buffers alternate,
    // and the D2H copy of chunk i runs on a second stream under the kernels of i+1.  The chunks
    // shrink (44/30/18/8 %): the copy of the last one is the only part nothing overlaps, while
    // every extra launch costs a kernel start-up and tail, so few launches and a small last one.
    // FB_ROW_CHUNKS=n (1..64) forces n equal chunks instead.
    std::vector<int> starts{0};
    if (rows_host && (uint64_t)n_shuffles * gps >= (1u << 20)) {
        const int unit = shuffles_per_slot > 0 ? shuffles_per_slot : 1;
        auto cut = [&](double frac) {  // chunk boundary at frac of the range, on a slot boundary
            int at = (int)((double)n_shuffles * frac / unit + 0.5) * unit;
            if (at > starts.back() && at < n_shuffles) starts.push_back(at);
        };
        if (const char* e = getenv("FB_ROW_CHUNKS")) {
            const int parts = std::max(1, std::min(64, atoi(e)));
            for (int i = 1; i < parts; i++) cut((double)i / parts);
        } else {
            cut(0.44);
            cut(0.74);
            cut(0.92);
        }
    }
    starts.push_back(n_shuffles);
    const int n_chunks = (int)starts.size() - 1;
    int chunk = 0;  // the largest chunk sizes the buffers
    for (int c = 0; c < n_chunks; c++) chunk = std::max(chunk, starts[c + 1] - starts[c]);
    const uint64_t chunk_games = (uint64_t)chunk * gps;
    const size_t stride = fb_row_stride(k);
    const size_t strat_b = align_up((size_t)n_strategies * sizeof(fb_strategy_t), 256);
    const size_t ids_b = align_up((size_t)n_strategies * 4, 256);
    const size_t tally_b = align_up((size_t)n_slots * n_tally_ids * FB_TALLY_WIDTH * 8, 256);
    const size_t totals_b = 256;
    const size_t rows_b = rows_host ? align_up(chunk_games * stride, 256) : 0;  // one of two buffers
    const size_t ws_b = fb_workspace_bytes(k, chunk_games) + 2 * align_up((size_t)chunk * n_strategies * 4, 256);
    const size_t total = strat_b + ids_b + tally_b + totals_b + (n_chunks > 1 ? 2 : 1) * rows_b + ws_b;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_ctx.host_ws_bytes < total) {
        if (g_ctx.host_ws) cudaFree(g_ctx.host_ws);
        g_ctx.host_ws = nullptr;
        g_ctx.host_ws_bytes = 0;
        FB_CUDA(cudaMalloc(&g_ctx.host_ws, total));
        g_ctx.host_ws_bytes = total;
    }
    if (!g_ctx.s_compute) {
        FB_CUDA(cudaStreamCreateWithFlags(&g_ctx.s_compute, cudaStreamNonBlocking));
        FB_CUDA(cudaStreamCreateWithFlags(&g_ctx.s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            FB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_played[i], cudaEventDisableTiming));
            FB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_copied[i], cudaEventDisableTiming));
        }
    }
    uint8_t* p = static_cast<uint8_t*>(g_ctx.host_ws);
    fb_strategy_t* d_strat = reinterpret_cast<fb_strategy_t*>(p); p += strat_b;
    int32_t* d_ids = reinterpret_cast<int32_t*>(p); p += ids_b;
    int64_t* d_tally = reinterpret_cast<int64_t*>(p); p += tally_b;
    int64_t* d_totals = reinterpret_cast<int64_t*>(p); p += totals_b;
    uint8_t* d_rows[2] = {rows_host ? p : nullptr, nullptr};
    p += rows_b;
    if (n_chunks > 1) { d_rows[1] = p; p += rows_b; }
    void* d_ws = p;
    cudaStream_t stream = g_ctx.s_compute, copy = g_ctx.s_copy;
    FB_CUDA(cudaMemcpyAsync(d_strat, strategies_host, (size_t)n_strategies * sizeof(fb_strategy_t),
                            cudaMemcpyHostToDevice, stream));
    if (strategy_ids_host)
        FB_CUDA(cudaMemcpyAsync(d_ids, strategy_ids_host, (size_t)n_strategies * 4, cudaMemcpyHostToDevice, stream));
    FB_CUDA(cudaMemsetAsync(d_tally, 0, tally_b, stream));
    FB_CUDA(cudaMemsetAsync(d_totals, 0, totals_b, stream));
    for (int c = 0; c < n_chunks; c++) {
        const int s0 = starts[c], cnt = starts[c + 1] - s0;
        const int b = c & 1;
        if (c >= 2) FB_CUDA(cudaStreamWaitEvent(stream, g_ctx.ev_copied[b], 0));  // row buffer b is free again
        int64_t* tally_c = nullptr;
        if (tallies_host)
            tally_c = d_tally + (shuffles_per_slot > 0 ? (size_t)(s0 / shuffles_per_slot) * n_tally_ids * FB_TALLY_WIDTH : 0);
        int rc = play_tournament_impl(root_seed, k, shuffle0 + (uint64_t)s0, cnt, d_strat,
                                      strategy_ids_host ? d_ids : nullptr, n_strategies, n_tally_ids, target_score,
                                      max_rounds, nullptr, nullptr, nullptr, 0, shuffles_per_slot, tally_c, d_totals,
                                      d_rows[b], want_game_seeds, d_ws, ws_b, stream, (uint32_t)((uint64_t)s0 * gps),
                                      nullptr, nullptr, /*ids_trusted=*/true);
        if (rc) return rc;
        if (rows_host) {
            FB_CUDA(cudaEventRecord(g_ctx.ev_played[b], stream));
            FB_CUDA(cudaStreamWaitEvent(copy, g_ctx.ev_played[b], 0));
            FB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(rows_host) + (uint64_t)s0 * gps * stride, d_rows[b],
                                    (uint64_t)cnt * gps * stride, cudaMemcpyDeviceToHost, copy));
            FB_CUDA(cudaEventRecord(g_ctx.ev_copied[b], copy));
        }
    }
    if (tallies_host)
        FB_CUDA(cudaMemcpyAsync(tallies_host, d_tally, (size_t)n_slots * n_tally_ids * FB_TALLY_WIDTH * 8,
                                cudaMemcpyDeviceToHost, stream));
    if (totals_host)
        FB_CUDA(cudaMemcpyAsync(totals_host, d_totals, FB_TOTALS_WIDTH * 8, cudaMemcpyDeviceToHost, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    if (rows_host) FB_CUDA(cudaStreamSynchronize(copy));
    return FB_OK;
}

static void launch_issue_peak(int variant, int grid, int iters, uint32_t* sink) {
    switch (variant) {
        case 0: issue_peak_kernel<0><<<grid, 1024>>>(iters, 12345u, sink); break;
        case 1: issue_peak_kernel<1><<<grid, 1024>>>(iters, 12345u, sink); break;
        case 2: issue_peak_kernel<2><<<grid, 1024>>>(iters, 12345u, sink); break;
        case 3: issue_peak_kernel<3><<<grid, 1024>>>(iters, 12345u, sink); break;
        default: issue_peak_kernel<4><<<grid, 1024>>>(iters, 12345u, sink); break;
    }
}

int fb_measure_issue_peak_variant(int variant, int iters, double* lane_ops_per_second) {
    FB_REQUIRE_INIT();
    if (iters < 1 || !lane_ops_per_second) return fail(FB_ERR_BAD_ARG, "iters >= 1 and an output pointer are required");
    if (variant < 0 || variant >= ISSUE_PEAK_VARIANTS)
        return fail(FB_ERR_BAD_ARG, "probe variant %d outside [0,%d)", variant, ISSUE_PEAK_VARIANTS);
    uint32_t* sink = nullptr;
    FB_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    FB_CUDA(cudaEventCreate(&e0));
    FB_CUDA(cudaEventCreate(&e1));
    const int grid = g_ctx.sm_count;
    launch_issue_peak(variant, grid, iters / 8 + 1, sink);  // warm-up
    int rc = launch_check("issue_peak_kernel");
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(e0));
    launch_issue_peak(variant, grid, iters, sink);
    rc = launch_check("issue_peak_kernel");
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(e1));
    FB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    FB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *lane_ops_per_second = (double)grid * 1024.0 * (double)iters * (double)ISSUE_PEAK_OPS[variant] / (ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return FB_OK;
}

int fb_measure_issue_peak(int iters, double* lane_ops_per_second) {
    // the best of the mixed-pipe variants is the peak a real instruction stream can be held against
    double best = 0.0;
    for (int v = 0; v < 4; v++) {
        double r = 0.0;
        const int rc = fb_measure_issue_peak_variant(v, iters, &r);
        if (rc) return rc;
        best = std::max(best, r);
    }
    *lane_ops_per_second = best;
    return FB_OK;
}

int fb_play_kernel_ms_history(float* out_ms, int max_entries) {
    if (!out_ms || max_entries < 0) return fail(FB_ERR_BAD_ARG, "bad history buffer");
    const int n = (int)std::min<uint64_t>({(uint64_t)max_entries, t_ev_count, (uint64_t)EV_RING});
    for (int i = 0; i < n; i++) {  // most recent first
        cudaEvent_t* ev = t_ev[(t_ev_count - 1 - (uint64_t)i) % EV_RING];
        FB_CUDA(cudaEventSynchronize(ev[1]));
        FB_CUDA(cudaEventElapsedTime(&out_ms[i], ev[0], ev[1]));
    }
    return n;
}

int fb_timeline(int enable) {
    std::lock_guard<std::mutex> lock(g_tl_mu);
    for (TlMark& m : g_tl) cudaEventDestroy(m.ev);
    g_tl.clear();
    g_tl_on.store(enable != 0);
    return FB_OK;
}

int fb_timeline_dump(char* out, size_t capacity) {
    if (!out || capacity == 0) return fail(FB_ERR_BAD_ARG, "bad timeline buffer");
    std::lock_guard<std::mutex> lock(g_tl_mu);
    size_t used = 0;
    out[0] = 0;
    int written = 0;
    for (const TlMark& m : g_tl) {
        float ms = 0.0f;
        FB_CUDA(cudaEventSynchronize(m.ev));
        FB_CUDA(cudaEventElapsedTime(&ms, g_tl.front().ev, m.ev));
        const int n = snprintf(out + used, capacity - used, "%d %s %.4f\n", m.lane, m.name, ms);
        if (n < 0 || (size_t)n >= capacity - used) break;
        used += (size_t)n;
        written++;
    }
    return written;
}

float fb_last_play_kernel_ms(void) {
    float ms = -1.0f;
    return fb_play_kernel_ms_history(&ms, 1) == 1 ? ms : -1.0f;
}

}  // extern "C"
