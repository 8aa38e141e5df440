// score_walk.cuh -- the suffix walks of the scorer (device functions), shared by the batched scorer kernel
// (score.cu) and the per-document build kernel (doc_sort.cu), which scores its document right after indexing it.
//
// RO = the suffix array, the per-rank key words and the bucket tables are read-only for the calling kernel
// (loads through the non-coherent path); the per-document kernel has just written them itself and reads them
// with plain loads after a CTA barrier.  The byte text and the queries are read-only for both.
#pragma once
#include <cstdint>

namespace east {

template <bool RO, typename T>
__device__ __forceinline__ T walk_ld(const T *p) {
    if (RO) return __ldg(p);
    return *p;
}

// 8 bytes at an arbitrary global address (two aligned 64-bit loads + funnel shift; the buffers carry slack)
__device__ __forceinline__ uint64_t load8u(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint64_t *q = (const uint64_t *)(a & ~(uintptr_t)7);
    const int sh = (int)(a & 7) * 8;
    const uint64_t lo = q[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (q[1] << (64 - sh));
}

// character of suffix rank r at depth d as a comparable value: 0 = "no character" (sorts first)
template <bool RO>
__device__ __forceinline__ uint64_t sym_at_(const uint32_t *__restrict__ T, const int32_t *sa,
                                           int32_t r, int32_t d, int32_t end) {
    int32_t p = walk_ld<RO>(sa + r) + d;
    return (p < end) ? (uint64_t)T[p] + 1ull : 0ull;
}

// PROBES: also count the ALGORITHMIC BYTES read: 8 per (SA word, text word) probe,
// 5 per (SA word, text byte) probe and 8 per bucket-table lookup on the fast path
#define sym_at(T, sa, r, d, end) (PROBES ? (probes += 8, sym_at_<RO>(T, sa, r, d, end)) : sym_at_<RO>(T, sa, r, d, end))
template <bool PROBES, bool RO>
__device__ __forceinline__ double score_one_suffix(const uint32_t *__restrict__ T, const int32_t *sa,
                                                   int32_t start, int32_t end, int32_t m,
                                                   const uint32_t *__restrict__ q, int32_t len, int normalized,
                                                   unsigned long long &probes) {
    int32_t lo = start, hi = end - 1;
    int32_t parent_f = (end - start) - m;
    int32_t d = 0, nodes = 0;
    double frac = 0.0;
    while (d < len) {
        const uint64_t c = (uint64_t)q[d] + 1ull;
        int32_t nlo, nhi;
        if (lo == hi) {
            if (sym_at(T, sa, lo, d, end) != c) break;
            nlo = lo; nhi = hi;
        } else {
            const uint64_t clo = sym_at(T, sa, lo, d, end);
            const uint64_t chi = sym_at(T, sa, hi, d, end);
            if (c < clo || c > chi) break;
            // lower bound: first rank in [lo, hi] whose symbol is >= c
            if (clo == c) {
                nlo = lo;
            } else {
                int32_t a = lo, b = hi;  // sym(a) < c <= sym(b)
                while (b - a > 1) {
                    int32_t mid = a + ((b - a) >> 1);
                    if (sym_at(T, sa, mid, d, end) < c) a = mid; else b = mid;
                }
                nlo = b;
                if (b != hi && sym_at(T, sa, b, d, end) != c) break;
                if (b == hi && chi != c) break;
            }
            // upper bound: last rank in [nlo, hi] whose symbol is <= c
            if (chi == c) {
                nhi = hi;
            } else {
                int32_t a = nlo, b = hi;  // sym(a) == c < sym(b)
                while (b - a > 1) {
                    int32_t mid = a + ((b - a) >> 1);
                    if (sym_at(T, sa, mid, d, end) <= c) a = mid; else b = mid;
                }
                nhi = a;
            }
        }
        const int32_t size = nhi - nlo + 1;
        if (d == 0 || size != hi - lo + 1) {
            ++nodes;
            frac = frac + (double)size / (double)parent_f;
        }
        lo = nlo; hi = nhi; parent_f = size; ++d;
    }
    if (d == 0) return 0.0;
    double r = (frac + (double)d) - (double)nodes;
    if (normalized) r = r / (double)d;
    return r;
}
#undef sym_at

// Fast path walk: depth 0 and 1 come from the 2-gram bucket table built during round 0 of the
// suffix sort (two table reads each instead of two binary searches over the largest intervals);
// deeper levels narrow by binary search over (SA word, text BYTE) probes.  Query symbols are dense
// codes (0 = absent from the batch: cannot match).  Same arithmetic, same order as the generic walk.
// SaT / rank_base: the per-document kernel walks a copy of its document in shared memory -- 16-bit positions
// local to the document, indexed by ranks local to the document (the bucket rows hold global ranks: minus rank_base).
template <bool PROBES, bool RO, typename SaT = int32_t>
__device__ __forceinline__ double score_one_suffix_fast(const uint8_t *__restrict__ T8, const SaT *sa,
                                                        const uint32_t *sk, const uint32_t *row, const uint32_t *row3, int b,
                                                        int32_t start, int32_t end, int32_t m,
                                                        const uint8_t *__restrict__ q, uint64_t qw, int32_t len, int normalized,
                                                        unsigned long long &probes, int32_t rank_base = 0) {
    // query symbol at depth d: the first 8 travel in the suffix record, deeper ones are read from the byte-coded queries
#define QSYM(d) ((d) < 8 ? (uint32_t)(qw >> (8 * (d))) & 0xffu : (uint32_t)q[d])
    // symbol of suffix rank r at depth d >= 2: depths 2..5 come from the per-rank key word (one load instead of
    // the dependent SA -> text pair), deeper ones from the text
#define SYM8(r, d) (PROBES ? (probes += 5, (uint32_t)T8[(int32_t)walk_ld<RO>(sa + (r)) + (d)]) \
                           : ((sk != nullptr && (d) < 6) ? ((walk_ld<RO>(sk + (r)) >> (8 * ((d) - 2))) & 0xffu) \
                                                         : (uint32_t)T8[(int32_t)walk_ld<RO>(sa + (r)) + (d)]))
    int32_t parent_f = (end - start) - m;
    const uint32_t c0 = QSYM(0);
    if (c0 == 0) return 0.0;
    // the table lookups of depths 0, 1 and 2 depend on the query only: all six loads are issued together
    // (one memory round trip instead of three dependent ones), then the depths are replayed in order
    const uint32_t c1 = len > 1 ? QSYM(1) : 0u, c2 = len > 2 ? QSYM(2) : 0u;
    const uint32_t x = (c0 << b) | c1, x3 = (x << b) | c2;
    const int32_t lo0 = (int32_t)walk_ld<RO>(row + (c0 << b)) - rank_base, hi0 = (int32_t)walk_ld<RO>(row + ((c0 + 1) << b)) - 1 - rank_base;
    const int32_t lo1 = (int32_t)walk_ld<RO>(row + x) - rank_base, hi1 = (int32_t)walk_ld<RO>(row + x + 1) - 1 - rank_base;
    int32_t lo2 = 0, hi2 = -1;
    if (row3 != nullptr) { lo2 = (int32_t)walk_ld<RO>(row3 + x3) - rank_base; hi2 = (int32_t)walk_ld<RO>(row3 + x3 + 1) - 1 - rank_base; }
    int32_t lo = lo0, hi = hi0;
    if (PROBES) probes += 8;
    if (hi < lo) return 0.0;
    int32_t size = hi - lo + 1;
    int32_t d = 1, nodes = 1;
    double frac = (double)size / (double)parent_f;
    parent_f = size;
    if (c1 != 0) {
        const int32_t nlo = lo1, nhi = hi1;
        if (PROBES) probes += 8;
        if (nhi >= nlo) {
            size = nhi - nlo + 1;
            if (size != hi - lo + 1) {
                ++nodes;
                frac = frac + (double)size / (double)parent_f;
            }
            lo = nlo; hi = nhi; parent_f = size; d = 2;
            if (row3 != nullptr && c2 != 0) {
                // depth 2 from the 3-gram table of the per-document build (its buckets): one lookup instead of
                // the binary search over the largest intervals of the walk
                if (PROBES) probes += 8;
                if (hi2 < lo2) len = 2;   // no such 3-gram: the walk ends here
                else {
                    size = hi2 - lo2 + 1;
                    if (size != hi - lo + 1) {
                        ++nodes;
                        frac = frac + (double)size / (double)parent_f;
                    }
                    lo = lo2; hi = hi2; parent_f = size; d = 3;
                }
            }
            while (d < len) {
                const uint32_t c = QSYM(d);
                if (c == 0) break;
                int32_t nl, nh;
                if (lo == hi) {
                    if (!PROBES) {
                        // One suffix left: every further depth keeps the interval (no new node, frac unchanged), so
                        // the rest of the walk is the length of the common prefix of the query and that suffix --
                        // one SA load and 8 symbols per step instead of one probe per depth.
                        const uint8_t *tp = T8 + (int32_t)walk_ld<RO>(sa + lo);
                        while (d < len) {
                            const uint64_t diff = load8u(tp + d) ^ load8u(q + d);
                            const int same = diff ? ((__ffsll((long long)diff) - 1) >> 3) : 8;
                            d += min(same, len - d);
                            if (same < 8) break;
                        }
                        break;
                    }
                    if (SYM8(lo, d) != c) break;
                    nl = lo; nh = hi;
                } else {
                    const uint32_t clo = SYM8(lo, d), chi = SYM8(hi, d);
                    if (c < clo || c > chi) break;
                    if (clo == c) {
                        nl = lo;
                    } else {
                        int32_t a = lo, e = hi;  // sym(a) < c <= sym(e)
                        while (e - a > 1) {
                            const int32_t mid = a + ((e - a) >> 1);
                            if (SYM8(mid, d) < c) a = mid; else e = mid;
                        }
                        nl = e;
                        if (e != hi && SYM8(e, d) != c) break;
                        if (e == hi && chi != c) break;
                    }
                    if (chi == c) {
                        nh = hi;
                    } else {
                        int32_t a = nl, e = hi;  // sym(a) == c < sym(e)
                        while (e - a > 1) {
                            const int32_t mid = a + ((e - a) >> 1);
                            if (SYM8(mid, d) <= c) a = mid; else e = mid;
                        }
                        nh = a;
                    }
                }
                size = nh - nl + 1;
                if (size != hi - lo + 1) {
                    ++nodes;
                    frac = frac + (double)size / (double)parent_f;
                }
                lo = nl; hi = nh; parent_f = size; ++d;
            }
        }
    }
    double r = (frac + (double)d) - (double)nodes;
    if (normalized) r = r / (double)d;
    return r;
#undef SYM8
#undef QSYM
}

}  // namespace east
