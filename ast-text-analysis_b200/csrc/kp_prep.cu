// kp_prep.cu -- everything the scorer derives from the keyphrases alone, computed on the device.
//
// The reference scores every suffix of every keyphrase against every text (east/applications.py:43-52 ->
// east/asts/easa.py:98-131).  A suffix result depends only on the code points from the suffix start to the end
// of its keyphrase, so identical query suffixes are walked ONCE per document; the distinct suffixes are visited
// in lexicographic order, so that the threads of a warp walk neighbouring SA intervals.  Round 1 did this on the
// host (hashing, std::sort, memcmp: 2.5 ms at 10^3 keyphrases, 0.6 s at 10^5 -- more than the device needs for
// the whole table).  Here:
//
//   k_kp_suffix_keys   one thread per keyphrase, backwards: 64-bit hash of every suffix, "contains a code point
//                      >= 0x0A00" flag, end of the suffix
//   k_kp_sortkeys      one thread per suffix: ONE 64-bit key = its first symbols (42 bits: 6 symbols of 7 bits for A-Z,
//                      3 of 12 bits for wide alphabets) above 22 hash bits
//   radix sort         the library's own onesweep LSD sort (8 passes) -> lexicographic by the first symbols, identical
//                      suffixes adjacent (equal key)
//   k_kp_mark          a suffix that equals its predecessor code point by code point (hash, length and a full
//                      comparison: a hash collision only costs a redundant walk, never a wrong twin) is a
//                      duplicate; every other one is the head of a group of identical suffixes
//   k_kp_scan_blocks   exclusive scan of the per-block head counts; the total is the number of distinct suffixes
//   k_kp_emit          position of every suffix's group in visiting order; one 16-byte record per distinct suffix
//   k_kp_encode        (needs the index: its dense alphabet) byte codes of the keyphrases and of the first 8
//                      symbols of every record
//
// Stage 1 (all but the last kernel) does not depend on the index: east_table_host runs it on a side stream while
// the text is still on its way to the device.
//
// Up to 64 Ki suffixes (10^3 keyphrases: 14 thousand) all of stage 1 is ONE kernel, k_kp_small, a cluster of 8 CTAs:
// the chain above is 17 launches of ~13 us each for 3 us of work, and the per-document kernel of a table call cannot
// start its first wave before the records exist.  The sort moves positions between two arrays in global memory (L2);
// the key bytes stay where they were written (one byte plane per pass); per-digit counts cross CTAs through
// distributed shared memory.  When the caller can tell which code table the index will have (a guessed alphabet, an
// existing index), the same kernel writes the dense codes as well and k_kp_encode has nothing left to do.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "kp_prep.h"
#include "radix_sort.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace east {

constexpr int KP_THREADS = 1024;
constexpr int KP_MAX_SYM_BITS = 12;   // code points of the texts are < 0x0A00: 12 bits hold code point + 1, larger ones are clamped

__global__ void __launch_bounds__(256)
k_kp_suffix_keys(const uint32_t *__restrict__ kp, const int32_t *__restrict__ off, int32_t K, uint64_t *__restrict__ hash,
                 uint32_t *__restrict__ vals, int32_t *__restrict__ send, uint8_t *__restrict__ weird) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        const int32_t b = off[k], e = off[k + 1];
        uint64_t h = 0x9e3779b97f4a7c15ull;
        uint8_t w = 0;
        for (int32_t p = e - 1; p >= b; --p) {
            const uint32_t cp = kp[p];
            h = h * 0x100000001b3ull + (uint64_t)cp + 0x632be59bd9b4e019ull;
            h ^= h >> 29;
            if (cp >= EAST_TERM_BASE) w = 1;
            hash[p] = h;
            vals[p] = (uint32_t)p;
            send[p] = e;
            weird[p] = w;
        }
    }
}

// ONE 64-bit sort key per suffix: its first symbols (sym_bits bits each: code point - smallest one + 1, clamped; 0 = past the end of the
// keyphrase) in the high lex_bits, hash bits below.  Sorted by it the suffixes are lexicographic by their first symbols
// and identical suffixes are adjacent; different suffixes that share the key only cost a redundant walk.
__global__ void __launch_bounds__(256)
k_kp_sortkeys(const uint32_t *__restrict__ kp, const int32_t *__restrict__ send, const uint64_t *__restrict__ hash, int32_t total,
              int sym_bits, uint32_t sym_min, int n_sym, int key_bits, uint64_t *__restrict__ keys) {
    const uint32_t top = (1u << sym_bits) - 1u;
    const int hash_bits = key_bits - n_sym * sym_bits;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int32_t e = send[p];
        uint64_t w = 0ull;
        for (int q = 0; q < n_sym; ++q) w = (w << sym_bits) | (uint64_t)((p + q < e) ? min(kp[p + q] - sym_min + 1u, top) : 0u);
        keys[p] = (w << hash_bits) | (hash[p] >> (64 - hash_bits));
    }
}

// block-wide inclusive scan of one 0/1 flag per thread (KP_THREADS threads); returns the inclusive count, *total = block sum
__device__ __forceinline__ uint32_t kp_block_scan(uint32_t flag, uint32_t *s_warp, uint32_t *total) {
    static_assert(KP_THREADS == 1024, "one lane per warp total");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag != 0u);
    const uint32_t incl = __popc(ballot & (0xffffffffu >> (31 - lane)));
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t x = s_warp[lane];   // every warp scans the 32 warp totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    const uint32_t below = __shfl_sync(0xffffffffu, x, (warp + 31) & 31);
    return (warp ? below : 0u) + incl;
}

__global__ void __launch_bounds__(KP_THREADS)
k_kp_mark(const uint32_t *__restrict__ kp, const uint64_t *__restrict__ hash, const int32_t *__restrict__ send,
          const uint32_t *__restrict__ vals, int32_t total, int dedup, uint8_t *__restrict__ flags, uint32_t *__restrict__ bsum) {
    __shared__ uint32_t s_warp[KP_THREADS / 32];
    const int i = blockIdx.x * KP_THREADS + threadIdx.x;
    uint32_t head = 0;
    if (i < total) {
        head = 1;
        if (dedup && i > 0) {
            const int32_t p = (int32_t)vals[i], q = (int32_t)vals[i - 1];
            const int32_t len = send[p] - p;
            if (hash[p] == hash[q] && send[q] - q == len) {
                bool same = true;
                for (int32_t x = 0; x < len && same; ++x) same = kp[p + x] == kp[q + x];
                if (same) head = 0;
            }
        }
        flags[i] = (uint8_t)head;
    }
    uint32_t sum;
    kp_block_scan(head, s_warp, &sum);
    if (threadIdx.x == 0) bsum[blockIdx.x] = sum;
}

// exclusive scan of bsum[nb] in place by one CTA; *n_uniq = total
__global__ void __launch_bounds__(KP_THREADS)
k_kp_scan_blocks(uint32_t *__restrict__ bsum, int32_t nb, uint32_t *__restrict__ n_uniq) {
    __shared__ uint32_t s_part[KP_THREADS];
    const int t = threadIdx.x;
    const int per = (nb + KP_THREADS - 1) / KP_THREADS;
    const int b0 = min(nb, t * per), b1 = min(nb, b0 + per);
    uint32_t sum = 0;
    for (int i = b0; i < b1; ++i) sum += bsum[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int i = 0; i < KP_THREADS; ++i) { const uint32_t v = s_part[i]; s_part[i] = run; run += v; }
        *n_uniq = run;
    }
    __syncthreads();
    uint32_t run = s_part[t];
    for (int i = b0; i < b1; ++i) { const uint32_t v = bsum[i]; bsum[i] = run; run += v; }
}

__global__ void __launch_bounds__(KP_THREADS)
k_kp_emit(const int32_t *__restrict__ send, const uint8_t *__restrict__ weird, const uint32_t *__restrict__ vals,
          const uint8_t *__restrict__ flags, const uint32_t *__restrict__ bsum, int32_t total, int32_t *__restrict__ uniq_of,
          SufRec *__restrict__ recs) {
    __shared__ uint32_t s_warp[KP_THREADS / 32];
    const int i = blockIdx.x * KP_THREADS + threadIdx.x;
    const uint32_t head = (i < total) ? flags[i] : 0u;
    uint32_t sum;
    const uint32_t incl = kp_block_scan(head, s_warp, &sum);
    if (i >= total) return;
    const uint32_t pos = bsum[blockIdx.x] + incl - 1u;   // position of the suffix's group in visiting order
    const int32_t p = (int32_t)vals[i];
    uniq_of[p] = (int32_t)pos;
    if (head) {
        SufRec r;
        r.q8_first = 0ull;
        r.sidx = p;
        r.len = (uint16_t)(send[p] - p);
        r.generic = weird[p];
        r.pad = weird[p];      // what stage 1 knows; `generic` is rewritten by every k_kp_encode
        recs[pos] = r;
    }
}

struct KpCodeTable { uint32_t w[EAST_TERM_BASE / 4]; };   // the code table of an index as a kernel parameter

// ---- stage 1 in one kernel: a cluster of 8 CTAs (total <= KP_SMALL_MAX)
constexpr int KP_CL_CTAS = 8;
constexpr int KP_SMALL_MAX = 65536;
constexpr int KP_CL_THREADS = KP_CL_CTAS * KP_THREADS;
constexpr int KP_CL_PER_CTA = KP_SMALL_MAX / KP_CL_CTAS;   // sorted positions one CTA marks and emits
constexpr int KP_CARRY = 2;                                // rounds of 32 a warp / of 1024 a CTA keeps in registers between two loops

// Data one CTA writes and another reads (hashes, key bytes, the two position arrays) lives in global memory and is
// read with ld.global.cg: L1 is not coherent between the SMs of the cluster; cluster.sync() orders the phases.  Only
// the per-digit counts cross CTAs through distributed shared memory.
__global__ void __cluster_dims__(KP_CL_CTAS, 1, 1) __launch_bounds__(KP_THREADS, 1)
k_kp_small(const uint32_t *__restrict__ kp, const int32_t *__restrict__ off, int32_t K, int32_t total, int dedup, int sym_bits, uint32_t sym_min,
           int n_sym, int key_bits, uint64_t *hash, int32_t *send, uint8_t *weird, uint8_t *planes, uint32_t *idx_a, uint32_t *idx_b,
           int32_t *__restrict__ uniq_of, SufRec *recs, uint32_t *__restrict__ n_uniq, unsigned long long *phase_clk,
           KpCodeTable tab, uint8_t *__restrict__ q8 /* NULL: no dense codes yet (k_kp_encode follows) */) {
    cg::cluster_group cluster = cg::this_cluster();
    long long t_prev = clock64();
#define KP_STAMP(k)                                                                     \
    do {                                                                                \
        if (phase_clk && blockIdx.x == 0 && threadIdx.x == 0) {                         \
            const long long t_now = clock64();                                          \
            phase_clk[k] += (unsigned long long)(t_now - t_prev);                       \
            t_prev = t_now;                                                             \
        }                                                                               \
    } while (0)
    __shared__ uint32_t cursor[(KP_THREADS / 32) * 256];   // [warp][digit]
    __shared__ uint32_t s_cta_tot[256];                    // read by the other CTAs of the cluster
    __shared__ uint32_t s_digit_base[256];
    __shared__ uint32_t s_warp[KP_THREADS / 32];
    __shared__ uint32_t s_heads;                           // read by the other CTAs of the cluster
    __shared__ uint8_t s_flag[KP_CL_PER_CTA];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = (int)cluster.block_rank();
    const int gtid = cta * KP_THREADS + tid, gwarp = cta * (KP_THREADS / 32) + warp;

    // hashes of all suffixes, keyphrase by keyphrase from its end (k_kp_suffix_keys); the code points are fetched 16 at
    // a time ahead of the serial hash chain
    for (int k = tid * KP_CL_CTAS + cta; k < K; k += KP_CL_THREADS) {   // few keyphrases: spread over all CTAs
        const int32_t b = off[k], e = off[k + 1];
        uint64_t h = 0x9e3779b97f4a7c15ull;
        uint8_t w = 0;
        for (int32_t hi = e; hi > b; hi -= 16) {
            uint32_t c[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = (hi - 1 - j >= b) ? kp[hi - 1 - j] : 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int32_t p = hi - 1 - j;
                if (p >= b) {
                    h = h * 0x100000001b3ull + (uint64_t)c[j] + 0x632be59bd9b4e019ull;
                    h ^= h >> 29;
                    if (c[j] >= EAST_TERM_BASE) w = 1;
                    hash[p] = h;
                    send[p] = e;
                    weird[p] = w;
                }
            }
        }
    }
    cluster.sync();
    KP_STAMP(0);

    const uint32_t *order = nullptr;   // nullptr: suffix order
    if (dedup) {
        // the sort key of k_kp_sortkeys, byte j of it to plane j
        const int passes = (key_bits + 7) / 8;
        const uint32_t top = (1u << sym_bits) - 1u;
        const int hash_bits = key_bits - n_sym * sym_bits;
        for (int p = gtid; p < total; p += KP_CL_THREADS) {
            const int32_t e = __ldcg(send + p);
            uint64_t w = 0ull;
            for (int q = 0; q < n_sym; ++q) w = (w << sym_bits) | (uint64_t)((p + q < e) ? min(kp[p + q] - sym_min + 1u, top) : 0u);
            const uint64_t key = (w << hash_bits) | (__ldcg(hash + p) >> (64 - hash_bits));
            for (int j = 0; j < passes; ++j) planes[(size_t)j * total + p] = (uint8_t)(key >> (8 * j));
            idx_a[p] = (uint32_t)p;
        }
        cluster.sync();
        KP_STAMP(1);
        // LSD radix sort of the positions, 8 bits per pass.  Warp g of the cluster's 256 owns the g-th segment of the
        // current order: its digit counts, scanned over (digit, CTA, warp), are the first output slots of its members;
        // inside the segment the members go out 32 at a time in order (match_any ranks equal digits): stable passes.
        const int seg = (((total + KP_CL_THREADS / 32 - 1) / (KP_CL_THREADS / 32)) + 31) & ~31;
        const int s0 = min(total, gwarp * seg), s1 = min(total, s0 + seg);
        uint32_t *my = cursor + warp * 256;
        uint32_t *src = idx_a, *dst = idx_b;
        for (int pass = 0; pass < passes; ++pass) {
            const uint8_t *pl = planes + (size_t)pass * total;
            for (int d = lane; d < 256; d += 32) my[d] = 0u;
            __syncwarp();
            // the first KP_CARRY rounds of the segment stay in registers for the scatter loop (10^3 keyphrases: all of it)
            uint32_t rv[KP_CARRY], rd[KP_CARRY];
#pragma unroll
            for (int j = 0; j < KP_CARRY; ++j) rv[j] = (s0 + 32 * j + lane < s1) ? __ldcg(src + s0 + 32 * j + lane) : 0u;
#pragma unroll
            for (int j = 0; j < KP_CARRY; ++j) rd[j] = (s0 + 32 * j + lane < s1) ? (uint32_t)__ldcg(pl + rv[j]) : (0x100u | (uint32_t)lane);
#pragma unroll
            for (int j = 0; j < KP_CARRY; ++j) if (s0 + 32 * j + lane < s1) atomicAdd(&my[rd[j]], 1u);
            for (int i = s0 + 32 * KP_CARRY + lane; i < s1; i += 32) atomicAdd(&my[__ldcg(pl + __ldcg(src + i))], 1u);
            __syncthreads();
            KP_STAMP(2);
            uint32_t all = 0, before = 0, x = 0;
            if (tid < 256) {
                uint32_t run = 0;
                for (int w = 0; w < KP_THREADS / 32; ++w) { const uint32_t v = cursor[w * 256 + tid]; cursor[w * 256 + tid] = run; run += v; }
                s_cta_tot[tid] = run;
            }
            cluster.sync();
            KP_STAMP(3);
            if (tid < 256) {
                for (int c = 0; c < KP_CL_CTAS; ++c) {
                    const uint32_t v = cluster.map_shared_rank(s_cta_tot, c)[tid];
                    all += v;
                    if (c < cta) before += v;
                }
                x = all;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (lane == 31) s_warp[warp] = x;
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t base = x - all;
                for (int w = 0; w < warp; ++w) base += s_warp[w];
                s_digit_base[tid] = base + before;
            }
            __syncthreads();
            KP_STAMP(4);
            auto scatter32 = [&](bool valid, uint32_t v, uint32_t d) {
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                if (valid) dst[s_digit_base[d] + my[d] + rank] = v;
                __syncwarp();
                if (valid && rank == 0u) my[d] += __popc(peers);
                __syncwarp();
            };
#pragma unroll
            for (int j = 0; j < KP_CARRY; ++j)
                if (s0 + 32 * j < s1) scatter32(s0 + 32 * j + lane < s1, rv[j], rd[j]);
            for (int i0 = s0 + 32 * KP_CARRY; i0 < s1; i0 += 32) {
                const bool valid = i0 + lane < s1;
                const uint32_t v = valid ? __ldcg(src + i0 + lane) : 0u;
                scatter32(valid, v, valid ? (uint32_t)__ldcg(pl + v) : (0x100u | (uint32_t)lane));
            }
            KP_STAMP(5);
            cluster.sync();   // the pass is in global memory; s_cta_tot may be rewritten
            KP_STAMP(6);
            uint32_t *t = src; src = dst; dst = t;
        }
        order = src;
    }

    // heads of the groups of identical suffixes (k_kp_mark): CTA c takes the c-th stretch of the sorted positions
    const int per = (total + KP_CL_CTAS - 1) / KP_CL_CTAS;
    const int r0 = min(total, cta * per), r1 = min(total, r0 + per);
    uint32_t heads = 0;
    int32_t cp_[KP_CARRY], ce_[KP_CARRY];   // suffix and its end for the first KP_CARRY chunks (emit re-reads the others)
    int chunk = 0;
    for (int i0 = r0; i0 < r1; i0 += KP_THREADS, ++chunk) {
        const int i = i0 + tid;
        uint32_t head = 0;
        int32_t p = 0, e = 0;
        if (i < r1) {
            head = 1;
            p = order ? (int32_t)__ldcg(order + i) : i;
            const int32_t q = (dedup && i > 0) ? (int32_t)__ldcg(order + i - 1) : 0;
            e = __ldcg(send + p);
            if (dedup && i > 0) {
                const int32_t len = e - p;
                if (__ldcg(hash + p) == __ldcg(hash + q) && __ldcg(send + q) - q == len) {
                    bool same = true;
                    for (int32_t x = 0; x < len && same; x += 4) {   // four code points per round trip
                        uint32_t a[4], c[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { a[j] = (x + j < len) ? kp[p + x + j] : 0u; c[j] = (x + j < len) ? kp[q + x + j] : 0u; }
#pragma unroll
                        for (int j = 0; j < 4; ++j) same = same && a[j] == c[j];
                    }
                    if (same) head = 0;
                }
            }
            s_flag[i - r0] = (uint8_t)head;
        }
#pragma unroll
        for (int j = 0; j < KP_CARRY; ++j) if (chunk == j) { cp_[j] = p; ce_[j] = e; }
        heads += (uint32_t)__syncthreads_count((int)head);
    }
    if (tid == 0) s_heads = heads;
    cluster.sync();
    KP_STAMP(7);
    // positions in visiting order, one record per group (k_kp_scan_blocks, k_kp_emit)
    uint32_t running = 0, n_all = 0;
    for (int c = 0; c < KP_CL_CTAS; ++c) {
        const uint32_t v = *cluster.map_shared_rank(&s_heads, c);
        if (c < cta) running += v;
        n_all += v;
    }
    if (cta == KP_CL_CTAS - 1 && tid == 0) *n_uniq = n_all;
    chunk = 0;
    for (int i0 = r0; i0 < r1; i0 += KP_THREADS, ++chunk) {
        const int i = i0 + tid;
        const uint32_t head = (i < r1) ? (uint32_t)s_flag[i - r0] : 0u;
        uint32_t sum;
        const uint32_t incl = kp_block_scan(head, s_warp, &sum);
        if (i < r1) {
            int32_t p = 0, e = 0;
            bool have = false;
#pragma unroll
            for (int j = 0; j < KP_CARRY; ++j) if (chunk == j) { p = cp_[j]; e = ce_[j]; have = true; }
            if (!have) { p = order ? (int32_t)__ldcg(order + i) : i; e = __ldcg(send + p); }
            const uint32_t pos = running + incl - 1u;
            uniq_of[p] = (int32_t)pos;
            if (head) {
                const uint8_t w = __ldcg(weird + p);
                SufRec r;
                r.q8_first = 0ull;
                r.sidx = p;
                r.len = (uint16_t)(e - p);
                r.generic = w;
                r.pad = w;
                recs[pos] = r;
            }
        }
        running += sum;
        __syncthreads();
    }
    cluster.sync();   // no CTA leaves while another may still read its shared memory; every record is in global memory
    KP_STAMP(8);
    // the dense codes for the code table the index is expected to have (k_kp_encode's work, one launch less on the way
    // to the first wave of the per-document kernel); shared memory from here on is this CTA's own
    if (q8) {
        uint32_t *s_tab = cursor;
        for (int i = tid; i < (int)(EAST_TERM_BASE / 4); i += KP_THREADS) s_tab[i] = tab.w[i];
        __syncthreads();
        const uint8_t *table = reinterpret_cast<const uint8_t *>(s_tab);
        for (int i = gtid; i < total + 16; i += KP_CL_THREADS) {
            uint8_t c = 0;
            if (i < total) { const uint32_t cp = kp[i]; if (cp < EAST_TERM_BASE) c = table[cp]; }
            q8[i] = c;
        }
        for (int i = gtid; i < (int)n_all; i += KP_CL_THREADS) {
            uint4 raw = __ldcg(reinterpret_cast<const uint4 *>(recs) + i);
            const int32_t sidx = (int32_t)raw.z;
            const int len = (int)(raw.w & 0xffffu);
            uint64_t first = 0ull;
            for (int q = 0; q < 8 && q < len; ++q) {
                const uint32_t cp = kp[sidx + q];
                first |= (uint64_t)(cp < EAST_TERM_BASE ? table[cp] : 0) << (8 * q);
            }
            raw.x = (uint32_t)first; raw.y = (uint32_t)(first >> 32);
            reinterpret_cast<uint4 *>(recs)[i] = raw;
        }
        KP_STAMP(9);
    }
#undef KP_STAMP
}

// index-dependent part: dense byte codes (0 = the code point does not occur in the batch, or is >= 0x0A00).  The code
// table of the index comes BY VALUE (2.5 KB of kernel parameters, staged in shared memory): an upload would be one more
// operation in the chain the first wave of the per-document kernel waits for, behind the text on the host link.
__global__ void __launch_bounds__(256)
k_kp_encode(const uint32_t *__restrict__ kp, int32_t total, KpCodeTable tab, int fast /* 0: the index has no fast path */,
            const uint32_t *__restrict__ n_uniq, SufRec *__restrict__ recs, uint8_t *__restrict__ q8) {
    __shared__ uint32_t s_tab[EAST_TERM_BASE / 4];
    if (fast) for (int i = threadIdx.x; i < (int)(EAST_TERM_BASE / 4); i += blockDim.x) s_tab[i] = tab.w[i];
    __syncthreads();
    const uint8_t *table = reinterpret_cast<const uint8_t *>(s_tab);
    const int32_t nu = (int32_t)*n_uniq;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(total + 16, nu); i += gridDim.x * blockDim.x) {
        if (fast && i < total + 16) {
            uint8_t c = 0;
            if (i < total) { const uint32_t cp = kp[i]; if (cp < EAST_TERM_BASE) c = table[cp]; }
            q8[i] = c;   // the scorer reads the queries 8 bytes at a time: 16 zero bytes of slack
        }
        if (i < nu) {
            SufRec r = recs[i];
            r.q8_first = 0ull;
            if (fast) {
                for (int q = 0; q < 8 && q < (int)r.len; ++q) {
                    const uint32_t cp = kp[r.sidx + q];
                    r.q8_first |= (uint64_t)(cp < EAST_TERM_BASE ? table[cp] : 0) << (8 * q);
                }
                r.generic = r.pad;
            } else {
                r.generic = 1;
            }
            recs[i] = r;
        }
    }
}

// pinned words the distinct-suffix counts are copied to: a small ring per thread, so that two preparations in flight on
// one thread (a table call's own + the cached one of a score call) never share a word
static thread_local uint32_t *g_pinned_words = nullptr;
static thread_local unsigned g_pinned_next = 0;
constexpr unsigned KP_PINNED_WORDS = 64;

// Symbols of the sort key: code point - smallest code point + 1 (0 = past the end), as many bits as the largest needs.
// The host has the keyphrases; without a host copy: 12 bits from 0 (code points of the texts are < 0x0A00).
static int kp_symbol_bits(const uint32_t *kp_host, int32_t total, uint32_t *sym_min) {
    *sym_min = 0;
    if (!kp_host || total <= 0) return KP_MAX_SYM_BITS;
    uint32_t mn = kp_host[0], mx = kp_host[0];
    for (int32_t p = 1; p < total; ++p) { mn = std::min(mn, kp_host[p]); mx = std::max(mx, kp_host[p]); }
    *sym_min = mn;
    return std::min(KP_MAX_SYM_BITS, std::max(1, bits_for((uint64_t)(mx - mn) + 1)));
}

static void kp_stage1_done(KpDevice &kp, cudaStream_t s) {
    if (!g_pinned_words) EAST_CUDA(cudaHostAlloc((void **)&g_pinned_words, sizeof(uint32_t) * KP_PINNED_WORDS, cudaHostAllocDefault));
    kp.n_uniq_host = g_pinned_words + (g_pinned_next++ % KP_PINNED_WORDS);
    EAST_CUDA(cudaMemcpyAsync(kp.n_uniq_host, kp.d_n_uniq.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (!kp.done) EAST_CUDA(cudaEventCreateWithFlags(&kp.done, cudaEventDisableTiming));
    EAST_CUDA(cudaEventRecord(kp.done, s));
}

void kp_stage1(KpDevice &kp, const uint32_t *kp_dev, const uint32_t *kp_host, const int64_t *kp_off, int32_t K, bool dedup, cudaStream_t s,
               int32_t small_max, const uint8_t *likely_code_table) {
    const int64_t total64 = kp_off[K];
    const int32_t total = (int32_t)total64;
    kp.total = total; kp.K = K; kp.dedup = dedup; kp.n_uniq = -1;
    kp.off32.resize((size_t)K + 1);
    for (int32_t k = 0; k <= K; ++k) kp.off32[(size_t)k] = (int32_t)kp_off[k];
    kp.d_off = DevBuf<int32_t>((size_t)K + 1, s);
    kp.d_uniq_of = DevBuf<int32_t>((size_t)total, s);
    kp.d_recs = DevBuf<SufRec>((size_t)total, s);
    kp.d_n_uniq = DevBuf<uint32_t>(1, s);
    EAST_CUDA(cudaMemcpyAsync(kp.d_off.p, kp.off32.data(), sizeof(int32_t) * ((size_t)K + 1), cudaMemcpyHostToDevice, s));

    static const int key_bits = getenv("EAST_KP_KEY_BITS") ? atoi(getenv("EAST_KP_KEY_BITS")) : 64;
    if (total <= std::min(small_max, KP_SMALL_MAX)) {
        // 40-bit keys here (5 passes): four 5-bit symbols for A-Z above 20 hash bits (at least 16).  The first three
        // symbols open a walk with table lookups; what the order beyond them buys is below the noise of a step.
        static const int small_key_bits = getenv("EAST_KP_KEY_BITS") ? key_bits : 40;
        uint32_t sym_min = 0;
        const int sym_bits = dedup ? kp_symbol_bits(kp_host, total, &sym_min) : KP_MAX_SYM_BITS;
        const int n_sym = std::max(1, (small_key_bits - 16) / sym_bits);
        const int passes = (small_key_bits + 7) / 8;
        const size_t t8 = ((size_t)total + 7) & ~(size_t)7;
        // hash | send | idx_a | idx_b | weird | planes, in 8-byte words
        DevBuf<uint64_t> scratch(t8 + 3 * (t8 / 2) + t8 / 8 + (size_t)passes * (t8 / 8) + 8, s);
        uint64_t *hash = scratch.p;
        int32_t *send = reinterpret_cast<int32_t *>(hash + t8);
        uint32_t *idx_a = reinterpret_cast<uint32_t *>(send + t8), *idx_b = idx_a + t8;
        uint8_t *weird = reinterpret_cast<uint8_t *>(idx_b + t8);
        uint8_t *planes = weird + t8;
        // EAST_KP_STAMPS=1: SM clocks of CTA 0 per phase, printed after a synchronize (debugging aid; not for timed runs)
        static const bool stamps = getenv("EAST_KP_STAMPS") != nullptr;
        DevBuf<unsigned long long> clk;
        if (stamps) { clk = DevBuf<unsigned long long>(16, s); EAST_CUDA(cudaMemsetAsync(clk.p, 0, 16 * sizeof(unsigned long long), s)); }
        KpCodeTable tab;
        kp.encoded_for.clear();
        if (likely_code_table) {
            memcpy(tab.w, likely_code_table, EAST_TERM_BASE);
            kp.encoded_for.assign(likely_code_table, likely_code_table + EAST_TERM_BASE);
            if (!kp.d_q8.p) kp.d_q8 = DevBuf<uint8_t>((size_t)total + 16, s);
        }
        EAST_LAUNCH(k_kp_small, KP_CL_CTAS, KP_THREADS, 0, s, kp_dev, kp.d_off.p, K, total, dedup ? 1 : 0, sym_bits, sym_min, n_sym, small_key_bits, hash, send,
                    weird, planes, idx_a, idx_b, kp.d_uniq_of.p, kp.d_recs.p, kp.d_n_uniq.p, clk.p, tab,
                    likely_code_table ? kp.d_q8.p : (uint8_t *)nullptr);
        if (stamps) {
            unsigned long long h[16];
            EAST_CUDA(cudaMemcpyAsync(h, clk.p, sizeof(h), cudaMemcpyDeviceToHost, s));
            EAST_CUDA(cudaStreamSynchronize(s));
            fprintf(stderr, "[east] k_kp_small clocks: hash %llu keys %llu | hist %llu sync %llu scan %llu scatter %llu sync %llu | mark %llu emit %llu encode %llu\n",
                    h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
        }
        kp_stage1_done(kp, s);
        return;
    }
    kp.encoded_for.clear();
    DevBuf<uint64_t> hash((size_t)total, s), keys_a, keys_b;
    DevBuf<uint32_t> vals_a((size_t)total, s), vals_b;
    DevBuf<int32_t> send((size_t)total, s);
    DevBuf<uint8_t> weird((size_t)total, s), flags((size_t)total, s);
    const int nb = (total + KP_THREADS - 1) / KP_THREADS;
    DevBuf<uint32_t> bsum((size_t)nb, s);
    EAST_LAUNCH(k_kp_suffix_keys, grid_for(K, 256, 8), 256, 0, s, kp_dev, kp.d_off.p, K, hash.p, vals_a.p, send.p, weird.p);
    const uint32_t *order = vals_a.p;
    if (dedup) {
        uint32_t sym_min = 0;
        const int sym_bits = kp_symbol_bits(kp_host, total, &sym_min);
        const int n_sym = std::max(1, (key_bits - 20) / sym_bits);
        keys_a = DevBuf<uint64_t>((size_t)total, s); keys_b = DevBuf<uint64_t>((size_t)total, s);
        vals_b = DevBuf<uint32_t>((size_t)total, s);
        DevBuf<uint32_t> hist(256 * RS_MAX_PASSES, s);
        DevBuf<uint8_t> scratch(rs_scratch_bytes(total, RS_MAX_PASSES), s);
        EAST_LAUNCH(k_kp_sortkeys, grid_for(total, 256, 8), 256, 0, s, kp_dev, send.p, hash.p, total, sym_bits, sym_min, n_sym, key_bits, keys_a.p);
        const int cur = radix_sort_pairs(keys_a.p, keys_b.p, vals_a.p, vals_b.p, total, key_bits, hist.p, false, scratch.p, s);
        order = cur ? vals_b.p : vals_a.p;
    }
    EAST_LAUNCH(k_kp_mark, nb, KP_THREADS, 0, s, kp_dev, hash.p, send.p, order, total, dedup ? 1 : 0, flags.p, bsum.p);
    EAST_LAUNCH(k_kp_scan_blocks, 1, KP_THREADS, 0, s, bsum.p, nb, kp.d_n_uniq.p);
    EAST_LAUNCH(k_kp_emit, nb, KP_THREADS, 0, s, send.p, weird.p, order, flags.p, bsum.p, total, kp.d_uniq_of.p, kp.d_recs.p);
    kp_stage1_done(kp, s);
    // the scratch buffers above go back to the pool in stream order
}

void kp_stage2(KpDevice &kp, const uint32_t *kp_dev, const uint8_t *code_table_host, cudaStream_t s) {
    EAST_CUDA(cudaStreamWaitEvent(s, kp.done, 0));
    const bool fast = code_table_host != nullptr;
    KpCodeTable tab;
    if (fast) {
        if (!kp.d_q8.p) kp.d_q8 = DevBuf<uint8_t>((size_t)kp.total + 16, s);
        memcpy(tab.w, code_table_host, EAST_TERM_BASE);
    }
    const bool encoded = fast && kp.encoded_for.size() == (size_t)EAST_TERM_BASE && memcmp(kp.encoded_for.data(), code_table_host, EAST_TERM_BASE) == 0;
    if (fast) kp.encoded_for.assign(code_table_host, code_table_host + EAST_TERM_BASE); else kp.encoded_for.clear();
    if (!encoded) EAST_LAUNCH(k_kp_encode, grid_for(kp.total + 16, 256, 8), 256, 0, s, kp_dev, kp.total, tab, fast ? 1 : 0, kp.d_n_uniq.p, kp.d_recs.p, kp.d_q8.p);
    if (kp.n_uniq < 0) {
        EAST_CUDA(cudaEventSynchronize(kp.done));
        kp.n_uniq = (int64_t)*kp.n_uniq_host;
    }
}

KpDevice::~KpDevice() { if (done) cudaEventDestroy(done); }

}  // namespace east
