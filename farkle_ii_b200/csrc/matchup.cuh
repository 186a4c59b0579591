// matchup.cuh — RNG lag statistics of the "matchup" groups of one tournament launch.
//
// Replaces, for one (root, k) cell, the count-route / eligibility / observation-route / external
// merge phases of the reference's rng_diagnostics stage for groups of type "matchup"
//   _observation_records / _observation_sort_fields   src/farkle/analysis/rng_diagnostics.py:1870-1961
//   _OnlineMetric.push                                 src/farkle/analysis/rng_diagnostics.py:2052-2064
// A matchup is the sorted multiset of strategy ids seated in a game; its sequence is the games
// with that multiset in (shuffle_index, game_index) order, one n_rounds observation per game.
//
// Pipeline (all on the launch's stream, data already in HBM from the play pass):
//   matchup_key_kernel     one thread per game: sort the k ids in registers, 64-bit key
//                          (k <= 2: the ids themselves packed into 2 * id_bits bits, injective, so
//                          the radix sort runs over those bits only; else a 64-bit mixing hash)
//   cub::DeviceRadixSort   (key, game ordinal) pairs; LSD radix sort is stable, ordinals start
//                          ascending, so equal keys stay in (shuffle, game) order
//   matchup_flag_kernel    segment starts; for hashed keys equal-key neighbours are compared id by
//                          id, a mismatch (hash collision) raises a flag and the call fails loudly
//   cub::DeviceScan        segment ids; first position of every segment; eligible segments
//                          (>= min observations) get a dense output slot by a second scan
//   matchup_stats_kernel   one thread per sorted position: lagged pairs inside its segment ->
//                          integer sums per (group, lag)
#pragma once
#include <cstdint>

namespace fb {

struct MatchupParams {
    const uint32_t* header;       // rounds | flags << 16 | ...
    const int32_t* perm;          // [n_games * k] table position seated at (game, seat)
    const int32_t* strategy_ids;  // id of table position, or nullptr = the position
    uint32_t n_games;
    int k;
    int n_lags;
    int lags[FB_MAX_LAGS];
    uint32_t min_obs;
    int id_bits;  // k <= 2: ids are < 2^id_bits and the key is ids[0] << id_bits | ids[1] (fewer sort passes)
    // scratch
    uint64_t* key;         // [n] sorted keys
    uint32_t* game;        // [n] game ordinal at sorted position
    uint32_t* seg_flag;    // [n] 1 at the first position of a segment
    uint32_t* seg_id1;     // [n] inclusive scan of seg_flag (segment id + 1)
    uint32_t* seg_first;   // [n + 1] first position of segment s; seg_first[n_seg] = n
    uint32_t* elig;        // [n] 1 if segment s has >= min_obs positions (0 beyond n_seg)
    uint32_t* slot;        // [n] exclusive scan of elig: output slot of segment s
    uint32_t* status;      // [0] collision flag, [1] number of eligible groups
    // outputs
    uint64_t capacity;
    int32_t* participants;        // [capacity][k]
    uint32_t* count;              // [capacity]
    unsigned long long* stats;    // [capacity][n_lags][FB_MATCHUP_LAG_WIDTH]
};

__device__ __forceinline__ void matchup_sorted_ids(const MatchupParams& M, uint32_t g, int32_t* ids) {
    const int k = M.k;
    for (int s = 0; s < k; s++) {
        const int32_t pos = M.perm[(size_t)g * k + s];
        int32_t v = M.strategy_ids ? M.strategy_ids[pos] : pos;
        int j = s;  // insertion sort, ascending
        while (j > 0 && ids[j - 1] > v) {
            ids[j] = ids[j - 1];
            j--;
        }
        ids[j] = v;
    }
}

__global__ void __launch_bounds__(256) matchup_key_kernel(const MatchupParams M, uint64_t* key_in, uint32_t* game_in) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M.n_games) return;
    int32_t ids[FB_MAX_PLAYERS];
    matchup_sorted_ids(M, g, ids);
    uint64_t key;
    if (M.k <= 2) {
        key = ((uint64_t)(uint32_t)ids[0] << M.id_bits) | (uint32_t)(M.k == 2 ? ids[1] : 0);
    } else {
        key = 0x9E3779B97F4A7C15ull;
        for (int s = 0; s < M.k; s++) {
            key = (key ^ (uint32_t)ids[s]) * 0xBF58476D1CE4E5B9ull;
            key ^= key >> 29;
        }
        key = (key ^ (key >> 32)) * 0x94D049BB133111EBull;
        key ^= key >> 31;
    }
    key_in[g] = key;
    game_in[g] = g;
}

__global__ void __launch_bounds__(256) matchup_flag_kernel(const MatchupParams M) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.n_games) return;
    bool start = p == 0 || M.key[p] != M.key[p - 1];
    if (!start && M.k > 2) {  // hashed key: make sure the neighbours really are the same matchup
        int32_t a[FB_MAX_PLAYERS], b[FB_MAX_PLAYERS];
        matchup_sorted_ids(M, M.game[p], a);
        matchup_sorted_ids(M, M.game[p - 1], b);
        for (int s = 0; s < M.k; s++)
            if (a[s] != b[s]) atomicOr(&M.status[0], 1u);
    }
    M.seg_flag[p] = start ? 1u : 0u;
}

__global__ void __launch_bounds__(256) matchup_first_kernel(const MatchupParams M) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.n_games) return;
    if (M.seg_flag[p]) M.seg_first[M.seg_id1[p] - 1u] = p;
    if (p == M.n_games - 1u) M.seg_first[M.seg_id1[p]] = M.n_games;
}

__global__ void __launch_bounds__(256) matchup_eligible_kernel(const MatchupParams M) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M.n_games) return;
    const uint32_t n_seg = M.seg_id1[M.n_games - 1u];
    M.elig[s] = (s < n_seg && M.seg_first[s + 1] - M.seg_first[s] >= M.min_obs) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) matchup_stats_kernel(const MatchupParams M) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.n_games) return;
    const uint32_t s = M.seg_id1[p] - 1u;
    if (p == M.n_games - 1u) M.status[1] = M.slot[s] + M.elig[s];  // s is the last segment here
    if (!M.elig[s]) return;
    const uint64_t e = M.slot[s];
    if (e >= M.capacity) return;  // reported through status[1] > capacity
    const uint32_t first = M.seg_first[s];
    if (p == first) {
        int32_t ids[FB_MAX_PLAYERS];
        matchup_sorted_ids(M, M.game[p], ids);
        for (int q = 0; q < M.k; q++) M.participants[e * M.k + q] = ids[q];
        M.count[e] = M.seg_first[s + 1] - first;
    }
    const unsigned long long y = M.header[M.game[p]] & 0xffffu;
    for (int z = 0; z < M.n_lags; z++) {
        const uint32_t lag = (uint32_t)M.lags[z];
        if (p < first + lag) continue;
        const unsigned long long x = M.header[M.game[p - lag]] & 0xffffu;
        unsigned long long* S = M.stats + (e * M.n_lags + z) * FB_MATCHUP_LAG_WIDTH;
        atomicAdd(&S[0], 1ull);
        atomicAdd(&S[1], x);
        atomicAdd(&S[2], y);
        atomicAdd(&S[3], x * x);
        atomicAdd(&S[4], y * y);
        atomicAdd(&S[5], x * y);
    }
}

}  // namespace fb
