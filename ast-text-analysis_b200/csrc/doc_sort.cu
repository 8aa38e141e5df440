// doc_sort.cu -- suffix array of every SMALL document by one CTA that keeps the document in
// shared memory (B200: 227 KB per CTA hold the text, the bucket counters and the sort scratch).
//
// Replaces east/asts/easa.py:141-245 (_compute_suftab) for batches whose documents have at most
// 65 535 code points each (BASELINE configs 1, 2, 4, 5: 10-50 KB texts) on the terminator-class
// fast path.  Larger documents keep the global prefix-doubling sort of sa_build.cu.
//
// Order being computed (equal to the code point order of the reference because dense codes are
// monotone and every terminator 0x0A00+i sorts above all text): compare symbol by symbol, all
// terminators share the top code `term`; two suffixes that agree up to and including their first
// terminator differ only in WHICH terminator they reach = string index = text position.
//
// Phases of one CTA (document of n code points, b bits per symbol, G = symbols per bucket id):
//   1  stage the byte-coded text in shared memory (128-bit loads)
//   2  histogram of the G-gram bucket ids (window cut after the first terminator): packed 16-bit
//      shared-memory counters
//   3  exclusive scan -> bucket starts; bitmap of the ranks where a bucket starts; list of the
//      buckets larger than a warp; the scorer's 2-gram table (first rank of every 2-gram)
//   4  scatter every suffix into its bucket (shared-memory cursor atomics, 4-byte global stores
//      that stay in L2)
//   5a buckets of <= 32 suffixes: one warp per window of 32 ranks ranks every suffix inside its own
//      bucket by counting the smaller ones; keys are the next WS symbols packed into 64 bits, ties go
//      to a byte-wise SWAR comparison of the shared-memory text
//   5b larger buckets: bitonic sort in shared memory by groups of 4 warps (<= 1024 suffixes) or by
//      the whole CTA (<= 8192); anything larger raises the overflow flag and the host falls back to
//      the global sort.
#include "sa_build.h"

namespace east {

constexpr int DS_THREADS = 1024;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_GROUP = 128;                       // threads of a medium-bucket sorting group
constexpr int DS_NGROUPS = DS_THREADS / DS_GROUP;   // 8
constexpr int DS_MED_MAX = 1024;                    // largest bucket a group sorts
constexpr int DS_BIG_MAX = 8192;                    // largest bucket the CTA sorts
constexpr int DS_SCR_BYTES = DS_BIG_MAX * 10;       // u64 key + u16 position per element: 80 KB
constexpr int DS_BIG_CAP = 2048;                    // >= 65535 / 33 buckets can be larger than a warp
constexpr int DS_WIN_BYTES = 64 * 8 + 64 * 4;       // per-warp window scratch: 64 keys + 64 positions

struct DocSortParams {
    const uint8_t *t8;        // byte-coded text of the batch (n + 128 bytes allocated)
    const int32_t *doc_off;
    int32_t *sa;              // out: global text positions in suffix order, doc-major
    uint32_t *bkt;            // out (optional): first global rank of every (document, 2-gram)
    uint32_t *overflow;       // out: set when a bucket exceeds DS_BIG_MAX
    int b, G, WS;             // bits per symbol, symbols per bucket id, symbols per 64-bit key word
    uint32_t term;            // terminator class code
    int text_cap;             // bytes reserved for the staged text
    int bits_words;           // words of the bucket-start bitmap
};

// 8 bytes of shared memory at an arbitrary byte offset from a 16-byte aligned base
__device__ __forceinline__ uint64_t ds_lds8(const uint8_t *base, int o) {
    const uint64_t *q = reinterpret_cast<const uint64_t *>(base + (o & ~7));
    const int sh = (o & 7) * 8;
    const uint64_t lo = q[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (q[1] << (64 - sh));
}

__device__ __forceinline__ uint64_t ds_haszero(uint64_t v) {
    return (v - 0x0101010101010101ull) & ~v & 0x8080808080808080ull;
}

// 8 consecutive byte codes (first symbol in the lowest byte) -> 8*b bits, first symbol on top
__device__ __forceinline__ uint64_t ds_pack8(uint64_t x, int b) {
    x = ((x & 0x00ff00ff00ff00ffull) << b) | ((x >> 8) & 0x00ff00ff00ff00ffull);
    x = ((x & 0x0000ffff0000ffffull) << (2 * b)) | ((x >> 16) & 0x0000ffff0000ffffull);
    return ((x & 0x00000000ffffffffull) << (4 * b)) | (x >> 32);
}

// bucket id of the suffix at byte offset o: its first G symbols, cut after the first terminator
__device__ __forceinline__ uint32_t ds_bucket(const uint8_t *s_raw, int o, uint64_t term8, int b, int G) {
    uint64_t x = ds_lds8(s_raw, o);
    const uint64_t z = ds_haszero(x ^ term8);
    if (z) {
        const int k = (__ffsll((long long)z) - 1) >> 3;  // byte index of the first terminator
        if (k < 7) x &= (1ull << (8 * (k + 1))) - 1ull;
    }
    uint32_t id = 0;
    for (int g = 0; g < G; ++g) id = (id << b) | (uint32_t)((x >> (8 * g)) & 0xffull);
    return id;
}

// sort key of the suffix at byte offset o: symbols [G, G+WS) of its cut window in the upper bits,
// bit 0 = "the window [0, G+WS) contains the terminator" (then equal keys are ordered by position)
__device__ __forceinline__ uint64_t ds_key(const uint8_t *s_raw, int o, uint64_t term8, int b, int G, int WS) {
    uint64_t x0 = ds_lds8(s_raw, o), x1 = ds_lds8(s_raw, o + 8);
    const uint64_t z0 = ds_haszero(x0 ^ term8);
    int tl = 16;  // offset of the first terminator among the 16 symbols (16 = none)
    if (z0) {
        tl = (__ffsll((long long)z0) - 1) >> 3;
    } else {
        const uint64_t z1 = ds_haszero(x1 ^ term8);
        if (z1) tl = 8 + ((__ffsll((long long)z1) - 1) >> 3);
    }
    if (tl < 7) { x0 &= (1ull << (8 * (tl + 1))) - 1ull; x1 = 0; }
    else if (tl == 7) x1 = 0;
    else if (tl < 15) x1 &= (1ull << (8 * (tl - 7))) - 1ull;
    // bytes G .. G+15 of the window
    const uint64_t y0 = (x0 >> (8 * G)) | (x1 << (64 - 8 * G));   // 1 <= G <= 7
    const uint64_t y1 = x1 >> (8 * G);
    uint64_t w;
    if (WS > 8) w = (ds_pack8(y0, b) << (b * (WS - 8))) | (ds_pack8(y1, b) >> (b * (16 - WS)));
    else w = ds_pack8(y0, b) >> (b * (8 - WS));
    return (w << 1) | (tl < G + WS ? 1ull : 0ull);
}

// suffixes at byte offsets oi, oj agree on [0, from) and have no terminator there: is i < j ?
__device__ __forceinline__ bool ds_deep_less(const uint8_t *s_raw, int oi, int oj, int from, uint64_t term8) {
    int o = from;
    while (true) {
        const uint64_t x = ds_lds8(s_raw, oi + o), y = ds_lds8(s_raw, oj + o);
        const uint64_t stop = (x ^ y) | ds_haszero(x ^ term8);
        if (stop) {
            const int sh = (__ffsll((long long)stop) - 1) & ~7;
            const uint32_t bx = (uint32_t)(x >> sh) & 0xffu, by = (uint32_t)(y >> sh) & 0xffu;
            if (bx != by) return bx < by;
            return oi < oj;  // both reach their terminator here: the earlier string comes first
        }
        o += 8;
    }
}

// (key, position) order; padding entries carry key ~0 and compare by position only
__device__ __forceinline__ bool ds_less(const uint8_t *s_raw, int shift, uint64_t ka, uint32_t pa, uint64_t kb,
                                        uint32_t pb, int from, uint64_t term8) {
    if (ka != kb) return ka < kb;
    if ((ka & 1ull) || ka == ~0ull) return pa < pb;
    return ds_deep_less(s_raw, shift + (int)pa, shift + (int)pb, from, term8);
}

__device__ __forceinline__ void ds_group_sync(int id, int nthr) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}

// bitonic sort of P (power of two) elements in shared memory by `nthr` threads (index tid).
// sync_id: 0 = __syncthreads (whole CTA), else named barrier id for nthr threads
__device__ void ds_bitonic(uint64_t *keys, uint16_t *pos, int P, int tid, int nthr, int sync_id,
                           const uint8_t *s_raw, int shift, int from, uint64_t term8) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nthr) {
                // t-th compare-exchange of this stage: partner indices differ in bit j
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const bool up = (lo & k) == 0;
                const uint64_t ka = keys[lo], kb = keys[hi];
                const uint32_t pa = pos[lo], pb = pos[hi];
                const bool b_lt_a = ds_less(s_raw, shift, kb, pb, ka, pa, from, term8);
                if (b_lt_a == up) {
                    keys[lo] = kb; keys[hi] = ka;
                    pos[lo] = (uint16_t)pb; pos[hi] = (uint16_t)pa;
                }
            }
            if (sync_id == 0) __syncthreads();
            else ds_group_sync(sync_id, nthr);
        }
    }
}

// sort one bucket [start, start+size) of the document's suffix array with `nthr` threads
__device__ void ds_sort_bucket(int32_t *sa_doc, int32_t base, int start, int size, uint64_t *keys, uint16_t *pos,
                               int tid, int nthr, int sync_id, const uint8_t *s_raw, int shift,
                               const DocSortParams &p, uint64_t term8) {
    int P = 64;
    while (P < size) P <<= 1;
    for (int e = tid; e < P; e += nthr) {
        if (e < size) {
            const int li = sa_doc[start + e] - base;
            keys[e] = ds_key(s_raw, shift + li, term8, p.b, p.G, p.WS);
            pos[e] = (uint16_t)li;
        } else {
            keys[e] = ~0ull;
            pos[e] = (uint16_t)(e & 0xffff);
        }
    }
    if (sync_id == 0) __syncthreads();
    else ds_group_sync(sync_id, nthr);
    ds_bitonic(keys, pos, P, tid, nthr, sync_id, s_raw, shift, p.G + p.WS, term8);
    for (int e = tid; e < size; e += nthr) sa_doc[start + e] = base + (int32_t)pos[e];
    if (sync_id == 0) __syncthreads();
    else ds_group_sync(sync_id, nthr);
}

__global__ void __launch_bounds__(DS_THREADS, 1)
k_doc_suffix_sort(DocSortParams p) {
    extern __shared__ __align__(16) uint8_t ds_smem[];
    uint8_t *s_raw = ds_smem;                                                    // staged text
    uint32_t *s_scr = reinterpret_cast<uint32_t *>(ds_smem + p.text_cap);        // counters, later sort scratch
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(ds_smem + p.text_cap + DS_SCR_BYTES);
    uint32_t *s_big = s_bits + p.bits_words;
    __shared__ uint32_t s_warp_sum[DS_WARPS];
    __shared__ uint32_t s_nbig;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int doc = blockIdx.x;
    const int32_t base = p.doc_off[doc];
    const int n = p.doc_off[doc + 1] - base;
    const int b = p.b, G = p.G, WS = p.WS;
    const uint64_t term8 = 0x0101010101010101ull * (uint64_t)p.term;
    const int NB = 1 << (G * b);      // bucket ids
    const int NW = NB >> 1;           // counter words (two 16-bit counters each)
    int32_t *sa_doc = p.sa + base;    // plain pointer: written and re-read by this CTA

    // ---- phase 1: stage the text, clear counters / bitmap
    const int32_t a0 = base & ~15;
    const int shift = base - a0;
    {
        const int nbytes = (shift + n + 48 + 15) & ~15;
        for (int o = tid * 16; o < nbytes; o += DS_THREADS * 16)
            *reinterpret_cast<uint4 *>(s_raw + o) = *reinterpret_cast<const uint4 *>(p.t8 + a0 + o);
        for (int i = tid; i < NW; i += DS_THREADS) s_scr[i] = 0;
        for (int i = tid; i < p.bits_words; i += DS_THREADS) s_bits[i] = 0;
        if (tid == 0) s_nbig = 0;
    }
    __syncthreads();

    // ---- phase 2: bucket histogram
    for (int i = tid; i < n; i += DS_THREADS) {
        const uint32_t id = ds_bucket(s_raw, shift + i, term8, b, G);
        atomicAdd(&s_scr[id >> 1], (id & 1u) ? 0x10000u : 1u);
    }
    __syncthreads();

    // ---- phase 3: exclusive scan of the counters (each thread owns CW consecutive words)
    {
        const int CW = (NW + DS_THREADS - 1) / DS_THREADS;
        const int w0 = tid * CW, w1 = min(NW, w0 + CW);
        uint32_t sum = 0;
        for (int w = w0; w < w1; ++w) { const uint32_t v = s_scr[w]; sum += (v & 0xffffu) + (v >> 16); }
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp_sum[warp] = x;
        __syncthreads();
        uint32_t run = x - sum;
        for (int i = 0; i < warp; ++i) run += s_warp_sum[i];
        const int sub = (G - 2) * b;                 // a 2-gram owns 2^sub consecutive bucket ids
        const uint32_t sub_mask = (1u << sub) - 1u;
        uint32_t *bkt_row = p.bkt ? p.bkt + ((size_t)doc << (2 * b)) : nullptr;
        for (int w = w0; w < w1; ++w) {
            const uint32_t v = s_scr[w];
            uint32_t packed = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t id = 2u * (uint32_t)w + (uint32_t)h;
                const uint32_t c = h ? (v >> 16) : (v & 0xffffu);
                if (bkt_row && (id & sub_mask) == 0u) bkt_row[id >> sub] = (uint32_t)base + run;
                if (c) {
                    atomicOr(&s_bits[run >> 5], 1u << (run & 31));
                    if (c > 32u) {
                        if (c > (uint32_t)DS_BIG_MAX) atomicOr(p.overflow, 1u);
                        else {
                            const uint32_t slot = atomicAdd(&s_nbig, 1u);
                            if (slot < (uint32_t)DS_BIG_CAP) s_big[slot] = (run << 16) | c;
                        }
                    }
                }
                packed |= (run & 0xffffu) << (16 * h);
                run += c;
            }
            s_scr[w] = packed;  // bucket start, used as the scatter cursor
        }
        if (tid == 0) atomicOr(&s_bits[n >> 5], 1u << (n & 31));  // sentinel: "a bucket starts at rank n"
    }
    __syncthreads();

    // ---- phase 4: scatter the suffixes into their buckets (arbitrary order inside a bucket)
    for (int i = tid; i < n; i += DS_THREADS) {
        const uint32_t id = ds_bucket(s_raw, shift + i, term8, b, G);
        const uint32_t old = atomicAdd(&s_scr[id >> 1], (id & 1u) ? 0x10000u : 1u);
        const uint32_t r = (id & 1u) ? (old >> 16) : (old & 0xffffu);
        sa_doc[r] = base + i;
    }
    __syncthreads();

    // ---- phase 5a: windows of 32 ranks; every bucket of <= 32 suffixes that STARTS in the window
    {
        uint64_t *wkeys = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_scr) + warp * DS_WIN_BYTES);
        uint32_t *wpos = reinterpret_cast<uint32_t *>(wkeys + 64);
        const int nwin = (n + 31) >> 5;
        for (int w = warp; w < nwin; w += DS_WARPS) {
            uint32_t word = s_bits[w];
            const bool last_win = w == (n >> 5);              // the window that holds the sentinel bit of rank n
            if (last_win) word &= (1u << (n & 31)) - 1u;
            if (word == 0u) continue;
            const int r0 = 32 * w + (__ffs(word) - 1);
            int r1 = n;   // first bucket start at or after the end of the window (the sentinel at the latest)
            if (!last_win) {
                int ww = w + 1;
                uint32_t x;
                while ((x = s_bits[ww]) == 0u) ++ww;
                r1 = 32 * ww + (__ffs(x) - 1);
            }
            const int last_start = 32 * w + (31 - __clz(word));
            int nseg = __popc(word);
            if (r1 - last_start > 32) { r1 = last_start; --nseg; }  // a large bucket: phase 5b
            const int len = r1 - r0;   // <= 63
            if (len <= nseg) continue; // only singletons
            uint64_t key[2];
            int li[2], sb[2], se[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int x = lane + 32 * s;
                key[s] = 0; li[s] = 0; sb[s] = 0; se[s] = 0;
                if (x < len) {
                    const int r = r0 + x;
                    li[s] = sa_doc[r] - base;
                    const int wp = r - 32 * w;  // 0..63
                    const uint32_t le = (wp < 32) ? (word & (0xffffffffu >> (31 - wp))) : word;
                    sb[s] = 32 * w + (31 - __clz(le)) - r0;
                    const uint32_t gt = (wp < 31) ? (word & (0xfffffffeu << wp)) : 0u;
                    se[s] = gt ? (32 * w + (__ffs(gt) - 1) - r0) : len;
                    if (se[s] > len) se[s] = len;
                    if (se[s] - sb[s] > 1) {
                        key[s] = ds_key(s_raw, shift + li[s], term8, b, G, WS);
                        wkeys[x] = key[s];
                    }
                    wpos[x] = (uint32_t)li[s];
                }
            }
            __syncwarp();
            int out[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int x = lane + 32 * s;
                out[s] = x;
                if (x < len && se[s] - sb[s] > 1) {
                    int below = 0;
                    for (int c = sb[s]; c < se[s]; ++c) {
                        if (c == x) continue;
                        const uint64_t kc = wkeys[c];
                        if (kc < key[s]) ++below;
                        else if (kc == key[s]) {
                            const uint32_t pc = wpos[c];
                            const bool less = (kc & 1ull) ? (pc < (uint32_t)li[s])
                                                          : ds_deep_less(s_raw, shift + (int)pc, shift + li[s], G + WS, term8);
                            if (less) ++below;
                        }
                    }
                    out[s] = sb[s] + below;
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int x = lane + 32 * s;
                if (x < len && se[s] - sb[s] > 1) sa_doc[r0 + out[s]] = base + li[s];
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- phase 5b: buckets of more than 32 suffixes
    {
        const int nbig = min((int)s_nbig, DS_BIG_CAP);
        // medium: one group of 4 warps per bucket
        const int g = tid / DS_GROUP, gt = tid % DS_GROUP;
        uint64_t *gkeys = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_scr) + g * (DS_MED_MAX * 10));
        uint16_t *gpos = reinterpret_cast<uint16_t *>(gkeys + DS_MED_MAX);
        bool any_large = false;
        for (int e = g; e < nbig; e += DS_NGROUPS) {
            const uint32_t ent = s_big[e];
            const int start = (int)(ent >> 16);
            const int size = (int)(ent & 0xffffu);
            if (size > DS_MED_MAX) { any_large = true; continue; }
            ds_sort_bucket(sa_doc, base, start, size, gkeys, gpos, gt, DS_GROUP, 1 + g, s_raw, shift, p, term8);
        }
        // large: the whole CTA, one bucket after the other (rare)
        if (__syncthreads_or(any_large ? 1 : 0)) {
            uint64_t *bkeys = reinterpret_cast<uint64_t *>(s_scr);
            uint16_t *bpos = reinterpret_cast<uint16_t *>(bkeys + DS_BIG_MAX);
            for (int e = 0; e < nbig; ++e) {
                const uint32_t ent = s_big[e];
                const int start = (int)(ent >> 16);
                const int size = (int)(ent & 0xffffu);
                if (size <= DS_MED_MAX) continue;
                ds_sort_bucket(sa_doc, base, start, size, bkeys, bpos, tid, DS_THREADS, 0, s_raw, shift, p, term8);
            }
        }
    }
}

// Host side: can the batch take the per-document path, and with which parameters
bool doc_sort_plan(int sigma, int32_t max_doc_n, DocSortPlan &plan) {
    const int b = bits_for((uint64_t)sigma + 1);
    if (b < 1 || b > 7) return false;
    if (max_doc_n > 65535) return false;
    int G = 15 / b;
    if (G > 7) G = 7;
    if (G < 2) return false;
    int WS = 63 / b;
    if (WS > 16 - G) WS = 16 - G;
    plan.b = b; plan.G = G; plan.WS = WS;
    plan.text_cap = (max_doc_n + 15 + 48 + 16 + 15) & ~15;
    plan.bits_words = (max_doc_n >> 5) + 3;
    plan.smem = (size_t)plan.text_cap + DS_SCR_BYTES + sizeof(uint32_t) * ((size_t)plan.bits_words + DS_BIG_CAP);
    return plan.smem <= (size_t)220 * 1024;
}

void doc_sort_launch(const DocSortPlan &plan, const uint8_t *t8, const int32_t *doc_off, int n_docs, int64_t n_total,
                     uint32_t term, int32_t *sa, uint32_t *bkt, uint32_t *overflow, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        EAST_CUDA(cudaFuncSetAttribute(k_doc_suffix_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured = true;
    }
    DocSortParams p;
    p.t8 = t8; p.doc_off = doc_off; p.sa = sa; p.bkt = bkt; p.overflow = overflow;
    p.b = plan.b; p.G = plan.G; p.WS = plan.WS; p.term = term;
    p.text_cap = plan.text_cap; p.bits_words = plan.bits_words;
    EAST_BYTES(9.0 * (double)n_total);  // byte text in, suffix array out + one re-read (L2-resident scatter)
    EAST_LAUNCH(k_doc_suffix_sort, n_docs, DS_THREADS, plan.smem, s, p);
}

}  // namespace east
