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
// Table VALUES are ranks local to the document (0 = none, as in the reference); table
// POSITIONS are global ranks of the batch.
#include "sa_build.h"

namespace east {

constexpr int TB_THREADS = 256;

__global__ void __launch_bounds__(TB_THREADS)
k_lcp(const uint32_t *__restrict__ T, const int32_t *__restrict__ sa, const int32_t *__restrict__ doc_off,
      int D, int32_t n, int32_t *__restrict__ lcp) {
    __shared__ int s_dlo, s_dhi;
    const int64_t stride = (int64_t)gridDim.x * TB_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * TB_THREADS; base < n; base += stride) {
        __syncthreads();
        if (threadIdx.x == 0) s_dlo = doc_of(doc_off, D, (int32_t)base);
        if (threadIdx.x == 32) s_dhi = doc_of(doc_off, D, (int32_t)min(base + TB_THREADS, (int64_t)n) - 1);
        __syncthreads();
        const int64_t r = base + threadIdx.x;
        if (r >= n) continue;
        int lo = s_dlo, hi = s_dhi;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (__ldg(doc_off + mid) <= r) lo = mid; else hi = mid - 1;
        }
        const int32_t start = __ldg(doc_off + lo), end = __ldg(doc_off + lo + 1);
        int32_t h = 0;
        if (r > start) {
            const int32_t i = sa[r - 1], j = sa[r];
            const int32_t lim = end - max(i, j);
            while (h < lim && T[i + h] == T[j + h]) ++h;
        }
        lcp[r] = h;
    }
}

void build_lcp(const uint32_t *text, const int32_t *sa, const int32_t *doc_off, int n_docs, int32_t n,
               int32_t *lcp, cudaStream_t s) {
    EAST_BYTES(16.0 * n);  // SA in, LCP out, >= one text word per suffix of each compared pair
    EAST_LAUNCH(k_lcp, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, text, sa, doc_off, n_docs, n, lcp);
}

__global__ void __launch_bounds__(TB_THREADS)
k_child_ann(const int32_t *__restrict__ lcp, const int32_t *__restrict__ doc_off,
            const int32_t *__restrict__ doc_m, int D, int32_t n, int32_t *__restrict__ up,
            int32_t *__restrict__ down, int32_t *__restrict__ next, int32_t *__restrict__ ann) {
    __shared__ int s_dlo, s_dhi;
    const int64_t stride = (int64_t)gridDim.x * TB_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * TB_THREADS; base < n; base += stride) {
        __syncthreads();
        if (threadIdx.x == 0) s_dlo = doc_of(doc_off, D, (int32_t)base);
        if (threadIdx.x == 32) s_dhi = doc_of(doc_off, D, (int32_t)min(base + TB_THREADS, (int64_t)n) - 1);
        __syncthreads();
        const int64_t p = base + threadIdx.x;
        if (p >= n) continue;
        int lo = s_dlo, hi = s_dhi;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (__ldg(doc_off + mid) <= p) lo = mid; else hi = mid - 1;
        }
        const int32_t start = __ldg(doc_off + lo), end = __ldg(doc_off + lo + 1);
        if (p == start) {
            ann[p] = (end - start) - __ldg(doc_m + lo);
            continue;
        }
        const int32_t l = lcp[p];
        int64_t q = p - 1;
        while (lcp[q] > l) --q;  // stops at `start` at the latest: lcp[start] == 0
        const int32_t lq = lcp[q];
        const int32_t p_local = (int32_t)(p - start);
        if (lq == l) {
            next[q] = p_local;
            continue;
        }
        int64_t e = p + 1;
        while (e < end && lcp[e] >= l) ++e;
        ann[p] = (int32_t)(e - q);
        if (e < end) {
            const int32_t le = lcp[e];
            if (lq <= le) up[e] = p_local;
            if (le <= lq) down[q] = p_local;
        }
    }
}

void build_child_ann(const int32_t *lcp, const int32_t *doc_off, const int32_t *doc_m, int n_docs,
                     int32_t n, int32_t *up, int32_t *down, int32_t *next, int32_t *ann, cudaStream_t s) {
    EAST_CUDA(cudaMemsetAsync(up, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(down, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(next, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_CUDA(cudaMemsetAsync(ann, 0, sizeof(int32_t) * (size_t)n, s));
    EAST_BYTES(8.0 * n);   // LCP in; annotation + child entries out (sparse)
    EAST_LAUNCH(k_child_ann, grid_for(n, TB_THREADS, 16), TB_THREADS, 0, s, lcp, doc_off, doc_m, n_docs, n,
                up, down, next, ann);
}

}  // namespace east
