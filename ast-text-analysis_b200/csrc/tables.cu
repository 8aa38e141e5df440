// tables.cu -- LCP table, child table and annotation table of a batch of documents.
//
//   LCP        replaces east/asts/easa.py:247-266 (_compute_lcptab, Kasai et al.)
//   up/down    replaces east/asts/easa.py:268-287 (_compute_childtab)
//   next       replaces east/asts/easa.py:289-304 (_compute_childtab_next_l_index)
//   ann        replaces east/asts/easa.py:306-331 (_compute_anntab + bottom-up traversal)
//
// The reference computes the three child arrays and the annotation with sequential stack
// passes.  Here every rank p works independently from two neighbours in the LCP array,
//     q = PSE(p) = largest  q < p with lcp[q] <= lcp[p]
//     e = NSV(p) = smallest e > p with lcp[e] <  lcp[p]   (document end if none)
// and the tables follow in closed form (proved equal to the stack code in DESIGN.md and
// checked against the oracle):
//     lcp[q] == lcp[p]                      ->  next[q] = p
//     lcp[q] <  lcp[p]  (p is the first l-index of the interval [q .. e-1]):
//                                               ann[p] = e - q            (= leaves below the node)
//         e inside the document and lcp[q] <= lcp[e]  ->  up[e]   = p
//         e inside the document and lcp[e] <= lcp[q]  ->  down[q] = p
//     first rank of a document              ->  ann = n - m   (easa.py:329)
// q and e are found with a min-pyramid over the LCP array (level k+1 = min of 32 entries of
// level k; level 1 is produced by the LCP kernel itself): a search scans at most 32 entries per
// level going up and 32 per level coming down, so no rank -- not even the first l-indices of the
// root's children, whose neighbours are thousands of ranks away -- does a long linear scan.
// Table VALUES are ranks local to the document (0 = none, as in the reference); table
// POSITIONS are global ranks of the batch.
#include "sa_build.h"

namespace east {

constexpr int TB_THREADS = 256;
constexpr int MAX_LEVELS = 8;  // 32^7 > 2^30

// Document of rank/position x for the lanes of one warp that hold 32 consecutive x: lane 0 and
// lane 31 binary-search the whole table once, every lane then searches only [d_first, d_last]
// (0-1 steps for real collections).  No shared memory, no block barrier.
__device__ __forceinline__ int warp_doc_of(const int32_t *__restrict__ doc_off, int D, int64_t x, int64_t n) {
    const int lane = threadIdx.x & 31;
    int d = 0;
    if (lane == 0 || lane == 31) d = doc_of(doc_off, D, (int32_t)min(x, n - 1));
    int lo = __shfl_sync(0xffffffffu, d, 0), hi = __shfl_sync(0xffffffffu, d, 31);
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(doc_off + mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct MinPyramid {
    const int32_t *lv[MAX_LEVELS];  // lv[0] = lcp
    int32_t size[MAX_LEVELS];
    int levels;                     // number of valid entries in lv[]
};

// 8 text bytes starting at an arbitrary address: two aligned 64-bit loads + funnel shift
// (the byte text is allocated with 64 bytes of slack, so the second load never faults)
__device__ __forceinline__ uint64_t load8(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint64_t *q = (const uint64_t *)(a & ~(uintptr_t)7);
    const int sh = (int)(a & 7) * 8;
    const uint64_t lo = q[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (q[1] << (64 - sh));
}

// FAST: compare dense byte codes 8 at a time.  All terminators share one byte code, so the
// common prefix also ends at the first terminator of either suffix (distinct terminators differ).
template <bool FAST>
__global__ void __launch_bounds__(TB_THREADS)
k_lcp(const uint32_t *__restrict__ T, const uint8_t *__restrict__ T8, uint64_t term8,
      const int32_t *__restrict__ sa, const int32_t *__restrict__ doc_off,
      int D, int32_t n, int32_t *__restrict__ lcp, int32_t *__restrict__ min1) {
    const int64_t stride = (int64_t)gridDim.x * TB_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * TB_THREADS; base < n; base += stride) {
        const int64_t r = base + threadIdx.x;
        int32_t h = 0x7fffffff;
        const int lo = warp_doc_of(doc_off, D, r, n);
        if (r < n) {
            const int32_t start = __ldg(doc_off + lo), end = __ldg(doc_off + lo + 1);
            h = 0;
            if (r > start) {
                const int32_t i = sa[r - 1], j = sa[r];
                if (FAST) {
                    while (true) {
                        const uint64_t x = load8(T8 + i + h), y = load8(T8 + j + h);
                        const uint64_t diff = x ^ y;
                        const uint64_t t = x ^ term8;  // zero byte <=> terminator in suffix i
                        const uint64_t tz = (t - 0x0101010101010101ull) & ~t & 0x8080808080808080ull;
                        const uint64_t stop = diff | tz;
                        if (stop) { h += (__ffsll((long long)stop) - 1) >> 3; break; }
                        h += 8;
                    }
                } else {
                    const int32_t lim = end - max(i, j);
                    while (h < lim && T[i + h] == T[j + h]) ++h;
                }
            }
            lcp[r] = h;
        }
        // level 1 of the min pyramid: one entry per 32 consecutive ranks (one warp)
        const int32_t wmin = __reduce_min_sync(0xffffffffu, h);
        if ((threadIdx.x & 31) == 0 && base + (threadIdx.x & ~31) < n) min1[(base + threadIdx.x) >> 5] = wmin;
    }
}

__global__ void __launch_bounds__(256)
k_min32(const int32_t *__restrict__ in, int32_t n_in, int32_t *__restrict__ out, int32_t n_out) {
    // one warp per output: coalesced read of its 32 inputs
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_out) return;
    const int64_t i = warp * 32 + lane;
    int32_t v = (i < n_in) ? in[i] : 0x7fffffff;
    v = __reduce_min_sync(0xffffffffu, v);
    if (lane == 0) out[warp] = v;
}

// largest q < p with lcp[q] <= l.  Exists: the first rank of the document has lcp 0.
__device__ __forceinline__ int32_t prev_le(const MinPyramid &M, int32_t p, int32_t l, int level = 0) {
    int32_t idx = p;  // index AT `level` of the group that contains the query rank
    int32_t j;
    while (true) {
        const int32_t gs = idx & ~31;
        const int32_t *a = M.lv[level];
        for (j = idx - 1; j >= gs; --j)
            if (a[j] <= l) goto found;
        idx >>= 5;
        ++level;
        if (level >= M.levels) return 0;  // unreachable for well-formed input
    }
found:
    while (level > 0) {
        --level;
        const int32_t *a = M.lv[level];
        int32_t c = min(j * 32 + 31, M.size[level] - 1);
        while (a[c] > l) --c;  // the group minimum is <= l, so this stops inside the group
        j = c;
    }
    return j;
}

// smallest e > p with lcp[e] < l, or n if there is none
__device__ __forceinline__ int32_t next_lt(const MinPyramid &M, int32_t p, int32_t l, int32_t n, int level = 0) {
    int32_t idx = p;  // index AT `level` of the group that contains the query rank
    int32_t j;
    while (true) {
        const int32_t ge = min((idx | 31) + 1, M.size[level]);
        const int32_t *a = M.lv[level];
        for (j = idx + 1; j < ge; ++j)
            if (a[j] < l) goto found;
        idx >>= 5;
        ++level;
        if (level >= M.levels) return n;
    }
found:
    while (level > 0) {
        --level;
        const int32_t *a = M.lv[level];
        int32_t c = j * 32;
        while (a[c] >= l) ++c;
        j = c;
    }
    return j;
}

__global__ void __launch_bounds__(TB_THREADS)
k_child_ann(MinPyramid M, const int32_t *__restrict__ doc_off, const int32_t *__restrict__ doc_m, int D,
            int32_t n, int32_t *__restrict__ up, int32_t *__restrict__ down, int32_t *__restrict__ next,
            int32_t *__restrict__ ann) {
    const int32_t *__restrict__ lcp = M.lv[0];
    const int64_t stride = (int64_t)gridDim.x * TB_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * TB_THREADS; base < n; base += stride) {
        const int32_t p = (int32_t)(base + threadIdx.x);
        const int lo = warp_doc_of(doc_off, D, base + threadIdx.x, n);
        if (p >= n) continue;
        const int32_t start = __ldg(doc_off + lo), end = __ldg(doc_off + lo + 1);
        if (p == start) {
            ann[p] = (end - start) - __ldg(doc_m + lo);
            continue;
        }
        const int32_t l = lcp[p];
        const int32_t q = prev_le(M, p, l);
        const int32_t lq = lcp[q];
        const int32_t p_local = p - start;
        if (lq == l) {
            next[q] = p_local;
            continue;
        }
        // here l > lq >= 0; the first rank of the next document has lcp 0 < l, so e <= end
        const int32_t e = next_lt(M, p, l, n);
        ann[p] = e - q;
        if (e < end) {
            const int32_t le = lcp[e];
            if (lq <= le) up[e] = p_local;
            if (le <= lq) down[q] = p_local;
        }
    }
}

// Warp-cooperative version: one warp owns 32 consecutive ranks = one level-0 group of the pyramid.
// PSE/NSV inside the group are found with 32 shuffle rounds (no divergence); ranks whose neighbour
// lies outside the group look at the 32 level-1 minima of the enclosing 1024-rank block, again by
// shuffles over one coalesced row; only what is still unresolved falls back to the per-thread
// pyramid search from level 2.
__global__ void __launch_bounds__(TB_THREADS)
k_child_ann_warp(MinPyramid M, const int32_t *__restrict__ doc_off, const int32_t *__restrict__ doc_m, int D,
                 int32_t n, int32_t *__restrict__ up, int32_t *__restrict__ down, int32_t *__restrict__ next,
                 int32_t *__restrict__ ann) {
    const int32_t *__restrict__ lcp = M.lv[0];
    const int32_t *__restrict__ lv1 = M.lv[1];
    const int32_t size1 = M.size[1];
    const int lane = threadIdx.x & 31;
    const int32_t INF = 0x7fffffff;
    const int64_t stride = (int64_t)gridDim.x * TB_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * TB_THREADS; base < n; base += stride) {
        const int32_t p = (int32_t)(base + threadIdx.x);
        const bool valid = p < n;
        int32_t start = 0, end = 0, dsel = 0;
        const int lo = warp_doc_of(doc_off, D, base + threadIdx.x, n);
        if (valid) {
            dsel = lo;
            start = __ldg(doc_off + lo); end = __ldg(doc_off + lo + 1);
        }
        const int32_t g = p >> 5;            // level-1 index of this warp's group (uniform in the warp)
        const int32_t v = valid ? lcp[p] : INF;
        // ---- inside the group
        int pse = -1, nsv = -1;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int32_t vj = __shfl_sync(0xffffffffu, v, j);
            if (j < lane && vj <= v) pse = j;
            if (j > lane && nsv < 0 && vj < v) nsv = j;
        }
        const bool is_root = valid && p == start;
        if (is_root) ann[p] = (end - start) - __ldg(doc_m + dsel);
        const bool active = valid && !is_root;
        int32_t q = (pse >= 0) ? (g << 5) + pse : -1;
        // ---- level-1 row of the enclosing 1024-rank block (one coalesced load per warp)
        const int32_t gs1 = g & ~31, gi = g & 31;
        const int32_t row = (gs1 + lane < size1) ? __ldg(lv1 + gs1 + lane) : INF;
        if (__any_sync(0xffffffffu, active && q < 0)) {
            int cand = -1;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int32_t mj = __shfl_sync(0xffffffffu, row, j);
                if (j < gi && mj <= v) cand = j;
            }
            if (active && q < 0) {
                if (cand >= 0) {
                    int32_t c = ((gs1 + cand) << 5) + 31;
                    while (lcp[c] > v) --c;
                    q = c;
                } else {
                    q = prev_le(M, g >> 5, v, 2);
                }
            }
        }
        int32_t lq = 0;
        bool first = false;
        if (active) {
            lq = lcp[q];
            if (lq == v) next[q] = p - start;
            else first = true;  // p is the first l-index of [q .. e-1]
        }
        // ---- NSV for the first l-indices
        int32_t e = (nsv >= 0) ? (g << 5) + nsv : -1;
        if (__any_sync(0xffffffffu, first && e < 0)) {
            int cand = -1;
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                const int32_t mj = __shfl_sync(0xffffffffu, row, j);
                if (j > gi && mj < v) cand = j;
            }
            if (first && e < 0) {
                if (cand >= 0) {
                    int32_t c = (gs1 + cand) << 5;
                    while (lcp[c] >= v) ++c;
                    e = c;
                } else {
                    e = next_lt(M, g >> 5, v, n, 2);
                }
            }
        }
        if (first) {
            ann[p] = e - q;
            if (e < end) {
                const int32_t le = lcp[e];
                if (lq <= le) up[e] = p - start;
                if (le <= lq) down[q] = p - start;
            }
        }
    }
}

void build_lcp_tables(const uint32_t *text, const uint8_t *t8, int term_code, const int32_t *sa, const int32_t *doc_off, const int32_t *doc_m,
                      int n_docs, int32_t n, int32_t *lcp, int32_t *up, int32_t *down, int32_t *next,
                      int32_t *ann, StageTimer &tm, cudaStream_t s, int child_variant) {
    // pyramid storage: sizes n/32, n/1024, ...
    MinPyramid M;
    M.lv[0] = lcp;
    M.size[0] = n;
    size_t total = 0;
    int levels = 1;
    for (int32_t sz = n; sz > 1 && levels < MAX_LEVELS; ++levels) {
        sz = (sz + 31) / 32;
        M.size[levels] = sz;
        total += (size_t)sz;
    }
    DevBuf<int32_t> pyr(total + 1, s);
    {
        size_t off = 0;
        for (int k = 1; k < levels; ++k) { M.lv[k] = pyr.p + off; off += (size_t)M.size[k]; }
        for (int k = levels; k < MAX_LEVELS; ++k) { M.lv[k] = nullptr; M.size[k] = 0; }
    }
    M.levels = levels;

    tm.mark("lcp");
    int32_t *min1 = levels > 1 ? const_cast<int32_t *>(M.lv[1]) : pyr.p;
    if (t8) {
        const uint64_t term8 = 0x0101010101010101ull * (uint64_t)(term_code & 0xff);
        EAST_BYTES(10.0 * n);  // SA in, LCP out, >= one text byte per suffix of each compared pair
        EAST_LAUNCH(k_lcp<true>, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, text, t8, term8, sa, doc_off,
                    n_docs, n, lcp, min1);
    } else {
        EAST_BYTES(16.0 * n);  // SA in, LCP out, >= one text word per suffix of each compared pair
        EAST_LAUNCH(k_lcp<false>, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, text, t8, 0ull, sa, doc_off,
                    n_docs, n, lcp, min1);
    }
    for (int k = 2; k < levels; ++k)
        EAST_LAUNCH(k_min32, (int)(((int64_t)M.size[k] * 32 + 255) / 256), 256, 0, s, M.lv[k - 1], M.size[k - 1],
                    const_cast<int32_t *>(M.lv[k]), M.size[k]);

    tm.mark("child_ann");
    EAST_CUDA(cudaMemsetAsync(up, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(down, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(next, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(ann, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_BYTES(8.0 * n);   // LCP in; annotation + child entries out (sparse)
    if (levels >= 3 && child_variant == 1) {
        EAST_LAUNCH(k_child_ann_warp, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, M, doc_off, doc_m, n_docs, n,
                    up, down, next, ann);
    } else {  // fewer than 1025 ranks: the plain per-thread pyramid search
        EAST_LAUNCH(k_child_ann, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, M, doc_off, doc_m, n_docs, n,
                    up, down, next, ann);
    }
}

}  // namespace east
