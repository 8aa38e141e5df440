// radix_sort.cuh -- hand-written one-sweep LSD radix sort of (uint64 key, uint32 value) pairs.
//
// This is the sorting engine of the prefix-doubling suffix-array construction that replaces
// east/asts/easa.py:141-245 (_compute_suftab/_kark_sort/_radixpass).
//
// Structure (one launch per 8-bit digit, least significant first):
//   * the digit histograms of ALL passes are accumulated up-front by whoever generates the
//     keys (shared-memory histograms, flushed with global atomics) -- see rs_hist_add();
//   * k_rs_scan_hist turns them into exclusive bin offsets;
//   * k_rs_onesweep: each CTA takes the next tile (ticket from an atomic counter, so tile
//     order == scheduling order and look-back cannot deadlock), ranks its 4096 keys with
//     warp-level match_any multi-split + per-warp shared-memory counters, publishes its
//     per-digit counts, resolves the counts of the preceding tiles by decoupled look-back,
//     stages keys and values through shared memory so every digit run leaves as one
//     coalesced burst, and scatters.  Stable by construction.
#pragma once
#include "common.cuh"

namespace east {

constexpr int RS_MIN_TILE = 2048;  // smallest tile of any kernel variant (sizes the look-back scratch)
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_LOOKBACK = 8;   // predecessor status words fetched per look-back round trip

constexpr uint32_t RS_FLAG_AGG = 1u << 30;
constexpr uint32_t RS_FLAG_PREFIX = 2u << 30;
constexpr uint32_t RS_VALUE_MASK = (1u << 30) - 1;

static inline int rs_num_passes(int nbits) { return (nbits + 7) / 8; }
static inline int rs_num_tiles(int64_t n, int tile = RS_MIN_TILE) { return (int)((n + tile - 1) / tile); }
// bytes of scratch (status words + tile tickets) for sorting n pairs with `passes` passes
static inline size_t rs_scratch_bytes(int64_t n, int passes) {
    return ((size_t)rs_num_tiles(n) * 256 * passes + 64) * sizeof(uint32_t);
}

// Segmented mode: the sort runs inside independent segments (documents).  Every tile belongs to
// exactly one segment; digit offsets are per segment and the look-back chain restarts at every
// segment, so the segment id never has to be part of the sorted bits and no tile ever walks
// more than (segment size / tile) predecessors.
struct RsTileDesc {
    int32_t start;      // first element of the tile (absolute index into the key/value arrays)
    int32_t n;          // elements in the tile (<= TILE)
    int32_t seg;        // segment index (row of the per-segment histogram)
    int32_t seg_start;  // first element of the segment
};

#ifdef __CUDACC__

// Accumulate the digits of one key into a CTA-private histogram s_hist[passes][256].
// Warp-uniform digits (document id bits, high rank bits) are folded into one atomic.
// Digits below `uniform_from` are known to be spread (text symbols): plain atomics, no vote.
__device__ __forceinline__ void rs_hist_add(uint32_t *s_hist, uint64_t key, int passes, bool valid,
                                            int uniform_from = 0) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int p = 0; p < passes; ++p) {
        unsigned d = (unsigned)(key >> (8 * p)) & 255u;
        if (p < uniform_from) {
            if (valid) atomicAdd(&s_hist[p * 256 + d], 1u);
            continue;
        }
        unsigned d0 = __shfl_sync(full, d, 0);
        if (__all_sync(full, valid && d == d0)) {
            if (lane == 0) atomicAdd(&s_hist[p * 256 + d0], 32u);
        } else if (valid) {
            atomicAdd(&s_hist[p * 256 + d], 1u);
        }
    }
}

// flush a CTA-private histogram to the global one
__device__ __forceinline__ void rs_hist_flush(const uint32_t *s_hist, uint32_t *g_hist, int passes) {
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
        uint32_t v = s_hist[i];
        if (v) atomicAdd(&g_hist[i], v);
    }
}

// exclusive scan of each pass' 256 bins, in place.  grid = passes, block = 256
static __global__ void __launch_bounds__(256) k_rs_scan_hist(uint32_t *hist) {
    __shared__ uint32_t s_warp[8];
    uint32_t *h = hist + blockIdx.x * 256;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    uint32_t v = h[t], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < w; ++i) base += s_warp[i];
    h[t] = base + x - v;
}

// standalone histogram (used when keys were not produced by a fused generator)
static __global__ void __launch_bounds__(256) k_rs_hist(const uint64_t *__restrict__ keys, int32_t n,
                                                 int passes, uint32_t *g_hist) {
    __shared__ uint32_t s_hist[RS_MAX_PASSES * 256];
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((int64_t)n + 31) & ~31ll;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool valid = i < n;
        uint64_t k = valid ? keys[i] : 0;
        rs_hist_add(s_hist, k, passes, valid);
    }
    __syncthreads();
    rs_hist_flush(s_hist, g_hist, passes);
}

// One LSD pass.  hist_excl: this pass' 256 exclusive bin offsets.  status: tiles*256 words,
// zero-initialised.  ticket: zero-initialised tile counter.
// THREADS x ITEMS pairs per CTA; MINB = CTAs per SM the register allocation must allow.
template <int THREADS, int ITEMS>
struct RsCfg {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    // s_cnt[WARPS][256] | s_dstart[256] | s_gbase[256] | s_wsum[8] | s_tile (+pad) | s_keys[TILE]
    static constexpr int SMEM = (WARPS * 256 + 256 + 256 + 8 + 8) * 4 + TILE * 8;
};

template <int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_rs_onesweep(const uint64_t *__restrict__ kin, uint64_t *__restrict__ kout,
              const uint32_t *__restrict__ vin, uint32_t *__restrict__ vout, int32_t n, int shift,
              const uint32_t *__restrict__ hist_excl, volatile uint32_t *status, uint32_t *ticket,
              const RsTileDesc *__restrict__ descs, int hist_seg_stride) {
    using Cfg = RsCfg<THREADS, ITEMS>;
    constexpr int WARPS = Cfg::WARPS, TILE = Cfg::TILE;
    extern __shared__ __align__(16) uint32_t rs_smem[];
    uint32_t(*s_cnt)[256] = reinterpret_cast<uint32_t(*)[256]>(rs_smem);
    uint32_t *s_dstart = rs_smem + WARPS * 256;
    uint32_t *s_gbase = s_dstart + 256;
    uint32_t *s_wsum = s_gbase + 256;
    uint32_t *s_tile = s_wsum + 8;
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(s_tile + 8);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_keys);

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) *s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int i = 0; i < 8; ++i) s_cnt[w][lane + 32 * i] = 0;
    __syncthreads();
    const uint32_t tile = *s_tile;
    int64_t tile_base = (int64_t)tile * TILE;
    int tile_n;
    bool first = tile == 0;   // first tile of its segment: nothing to look back at
    uint32_t out_base = 0;    // where the segment starts in the output
    if (descs != nullptr) {
        const RsTileDesc d = descs[tile];
        tile_base = d.start;
        tile_n = d.n;
        first = d.start == d.seg_start;
        out_base = (uint32_t)d.seg_start;
        hist_excl += (size_t)d.seg * hist_seg_stride;
    } else {
        tile_n = (int)min((int64_t)TILE, (int64_t)n - tile_base);
    }

    // ---- load (warp-striped: item j of lane l is element w*32*ITEMS + j*32 + l of the tile)
    uint64_t key[ITEMS];
    uint32_t val[ITEMS];
    const int wbase = w * (32 * ITEMS) + lane;
    if (tile_n == TILE) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) key[j] = kin[tile_base + wbase + j * 32];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) val[j] = vin[tile_base + wbase + j * 32];
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            int o = wbase + j * 32;
            key[j] = (o < tile_n) ? kin[tile_base + o] : ~0ull;
            val[j] = (o < tile_n) ? vin[tile_base + o] : 0u;
        }
    }

    // ---- rank inside the warp (stable: items are visited in element order).
    // All lanes holding the same digit read the running counter, the HIGHEST of them writes it
    // back (its own rank + 1 is the group size): one MATCH and one POPC per key.
    uint32_t pos[ITEMS];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const unsigned d = (unsigned)(key[j] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = __popc(peers & lt);
        const uint32_t old = s_cnt[w][d];
        __syncwarp();
        if ((peers >> lane) == 1u) s_cnt[w][d] = old + below + 1u;  // no peer above this lane
        __syncwarp();
        pos[j] = old + below;
    }
    __syncthreads();

    // ---- per-digit exclusive prefix over warps, digit totals (thread t < 256 owns digit t)
    uint32_t total = 0;
    volatile uint32_t *my_status = status + (size_t)tile * 256 + (t & 255);
    if (t < 256) {
#pragma unroll
        for (int i = 0; i < WARPS; ++i) {
            uint32_t c = s_cnt[i][t];
            s_cnt[i][t] = total;
            total += c;
        }
        // publish the tile aggregate as early as possible
        if (first) *my_status = RS_FLAG_PREFIX | total;
        else *my_status = RS_FLAG_AGG | total;
    }
    // ---- exclusive scan of totals over the 256 digits -> local start of each digit run
    uint32_t x = total;
    if (t < 256) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_wsum[w] = x;
    }
    __syncthreads();
    if (t < 256) {
        uint32_t base = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) base += (i < w) ? s_wsum[i] : 0u;
        s_dstart[t] = base + x - total;

        // ---- decoupled look-back: how many keys with digit t precede this tile.  The status
        // words of RS_LOOKBACK predecessors are fetched together (independent loads, one L2 round
        // trip) and consumed in order; a word that is not published yet restarts the fetch there.
        uint32_t excl = 0;
        if (!first) {
            int64_t prev = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t sv[RS_LOOKBACK];
#pragma unroll
                for (int u = 0; u < RS_LOOKBACK; ++u) {
                    sv[u] = 2u << 30;  // before the first tile: an empty inclusive prefix
                    if (prev - u >= 0) sv[u] = status[(size_t)(prev - u) * 256 + t];
                }
#pragma unroll
                for (int u = 0; u < RS_LOOKBACK; ++u) {
                    if (done) break;
                    const uint32_t flag = sv[u] & ~RS_VALUE_MASK;
                    if (flag == 0) break;  // not there yet: refetch starting at this predecessor
                    excl += sv[u] & RS_VALUE_MASK;
                    --prev;
                    if (flag == RS_FLAG_PREFIX) done = true;
                }
            }
            *my_status = RS_FLAG_PREFIX | (excl + total);
        }
        s_gbase[t] = out_base + hist_excl[t] + excl - s_dstart[t];
    }
    __syncthreads();

    // ---- keys: scatter into shared memory at their tile-sorted position, then stream out
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        unsigned d = (unsigned)(key[j] >> shift) & 255u;
        pos[j] += s_dstart[d] + s_cnt[w][d];
        s_keys[pos[j]] = key[j];
    }
    __syncthreads();
    uint32_t dig[(ITEMS + 3) / 4];
#pragma unroll
    for (int i = 0; i < (ITEMS + 3) / 4; ++i) dig[i] = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        int p = t + i * THREADS;
        uint64_t k = s_keys[p];
        unsigned d = (unsigned)(k >> shift) & 255u;
        dig[i >> 2] |= d << (8 * (i & 3));
        if (p < tile_n) kout[s_gbase[d] + p] = k;
    }
    __syncthreads();
    // ---- values follow the same permutation
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) s_vals[pos[j]] = val[j];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        int p = t + i * THREADS;
        unsigned d = (dig[i >> 2] >> (8 * (i & 3))) & 255u;
        if (p < tile_n) vout[s_gbase[d] + p] = s_vals[p];
    }
}

#endif  // __CUDACC__

// Host driver.  Sorts n pairs on bits [0, nbits) of the key.  `hist` holds passes*256 counters
// that are either already accumulated (hist_ready) or computed here.  scratch: rs_scratch_bytes.
// Returns 0 if the result is in (ka, va), 1 if in (kb, vb).
int radix_sort_pairs(uint64_t *ka, uint64_t *kb, uint32_t *va, uint32_t *vb, int32_t n, int nbits,
                     uint32_t *hist, bool hist_ready, void *scratch, cudaStream_t s, int variant = 0);

// Segmented variant: descs[num_tiles] (device) describe tiles of RS_SEG_TILE elements that never
// cross a segment; hist[n_seg][passes][256] holds the per-segment digit counts (already accumulated).
constexpr int RS_SEG_TILE = 4096;
int radix_sort_pairs_segmented(uint64_t *ka, uint64_t *kb, uint32_t *va, uint32_t *vb, int nbits,
                               const RsTileDesc *descs, int num_tiles, int64_t n_total, uint32_t *hist, int n_seg,
                               void *scratch, cudaStream_t s);

}  // namespace east
