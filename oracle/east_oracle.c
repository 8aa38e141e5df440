/*
 * east_oracle.c -- CPU restatement of EAST's Enhanced Annotated Suffix Array (EASA).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library, and only as the
 * checker (or as the timed CPU baseline) -- never as a product code path.
 *
 * Parity status: PINNED.  Checked bit-for-bit against the reference itself
 * (py3-patched copy made by oracle/make_ref.py) on the committed golden vectors in
 * tests/golden/ and, when oracle/_ref exists, on random collections (tests/test_oracle.py).
 *
 * Every function cites the reference lines (relative to /root/reference) it follows.
 * All arrays are int32 here (int64 in the reference); values are identical.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EAST_TERMINATOR_BASE 0x0A00u /* east/consts.py:24 UNICODE_SPECIAL_SYMBOLS_START */

/* ------------------------------------------------------------------------------------------
 * east/asts/utils.py:25-40 make_unique_endings + east/asts/easa.py:19 "".join(...)
 * chars: concatenated code points of the m strings, str_off[m+1] their boundaries.
 * out must hold str_off[m] + m code points.  Returns n.
 * ---------------------------------------------------------------------------------------- */
int64_t oracle_pack(const uint32_t *chars, const int64_t *str_off, int32_t m, uint32_t *out)
{
    int64_t w = 0;
    for (int32_t s = 0; s < m; ++s) {
        for (int64_t p = str_off[s]; p < str_off[s + 1]; ++p)
            out[w++] = chars[p];
        out[w++] = EAST_TERMINATOR_BASE + (uint32_t)s;
    }
    return w;
}

/* ------------------------------------------------------------------------------------------
 * east/asts/easa.py:141-245 _compute_suftab (DC3 there).  The suffix array of T under code
 * point order is unique (the last code point of T is a terminator that occurs once), so any
 * correct construction gives the same array.  Here: Manber-Myers prefix doubling with
 * counting sorts -- deliberately a different algorithm from both the reference (DC3) and
 * the CUDA path (packed-key radix doubling).
 * A suffix that runs off the end of T compares smaller (never decides for well-formed T).
 * ---------------------------------------------------------------------------------------- */
static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

int oracle_suffix_array(const uint32_t *T, int32_t n, int32_t *sa)
{
    if (n <= 0) return 0;
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *sa2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    if (!rank || !tmp || !sa2 || !cnt) { free(rank); free(tmp); free(sa2); free(cnt); return -1; }

    /* dense ranks of the code points */
    uint32_t maxc = 0;
    for (int32_t i = 0; i < n; ++i) if (T[i] > maxc) maxc = T[i];
    int32_t classes;
    if ((uint64_t)maxc < ((uint64_t)1 << 26)) {
        uint8_t *seen = (uint8_t *)calloc((size_t)maxc + 1, 1);
        int32_t *code = (int32_t *)malloc(sizeof(int32_t) * ((size_t)maxc + 1));
        if (!seen || !code) { free(seen); free(code); free(rank); free(tmp); free(sa2); free(cnt); return -1; }
        for (int32_t i = 0; i < n; ++i) seen[T[i]] = 1;
        int32_t c = 0;
        for (uint32_t v = 0; v <= maxc; ++v) if (seen[v]) code[v] = c++;
        classes = c;
        for (int32_t i = 0; i < n; ++i) rank[i] = code[T[i]];
        free(seen); free(code);
    } else {
        uint32_t *vals = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        if (!vals) { free(rank); free(tmp); free(sa2); free(cnt); return -1; }
        memcpy(vals, T, sizeof(uint32_t) * (size_t)n);
        qsort(vals, (size_t)n, sizeof(uint32_t), cmp_u32);
        int32_t u = 0;
        for (int32_t i = 0; i < n; ++i) if (i == 0 || vals[i] != vals[i - 1]) vals[u++] = vals[i];
        classes = u;
        for (int32_t i = 0; i < n; ++i) {
            int32_t lo = 0, hi = u - 1;
            while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (vals[mid] < T[i]) lo = mid + 1; else hi = mid; }
            rank[i] = lo;
        }
        free(vals);
    }
    /* counting sort by first code point */
    memset(cnt, 0, sizeof(int32_t) * ((size_t)n + 1));
    for (int32_t i = 0; i < n; ++i) cnt[rank[i] + 1]++;
    for (int32_t c = 0; c < classes && c < n; ++c) cnt[c + 1] += cnt[c];
    for (int32_t i = 0; i < n; ++i) sa[cnt[rank[i]]++] = i;

    for (int64_t h = 1;; h <<= 1) {
        if (classes >= n) break;
        /* order by second key rank[i+h]: suffixes with i+h >= n first, then by SA order */
        int32_t p = 0;
        for (int64_t i = n - h; i < n; ++i) if (i >= 0) sa2[p++] = (int32_t)i;
        for (int32_t j = 0; j < n; ++j) if (sa[j] >= h) sa2[p++] = (int32_t)(sa[j] - h);
        /* stable counting sort by first key rank[i] */
        memset(cnt, 0, sizeof(int32_t) * ((size_t)classes + 1));
        for (int32_t i = 0; i < n; ++i) cnt[rank[i] + 1]++;
        for (int32_t c = 0; c < classes; ++c) cnt[c + 1] += cnt[c];
        for (int32_t j = 0; j < n; ++j) sa[cnt[rank[sa2[j]]]++] = sa2[j];
        /* re-rank */
        int32_t c = 0;
        tmp[sa[0]] = 0;
        for (int32_t j = 1; j < n; ++j) {
            int32_t a = sa[j - 1], b = sa[j];
            int32_t a2 = (a + h < n) ? rank[a + h] : -1;
            int32_t b2 = (b + h < n) ? rank[b + h] : -1;
            if (rank[a] != rank[b] || a2 != b2) ++c;
            tmp[b] = c;
        }
        memcpy(rank, tmp, sizeof(int32_t) * (size_t)n);
        classes = c + 1;
        if (h > n) break;
    }
    free(rank); free(tmp); free(sa2); free(cnt);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * east/asts/easa.py:247-266 _compute_lcptab (Kasai et al.)
 * ---------------------------------------------------------------------------------------- */
int oracle_lcp(const uint32_t *T, int32_t n, const int32_t *sa, int32_t *lcp)
{
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    if (!rank) return -1;
    for (int32_t i = 0; i < n; ++i) rank[sa[i]] = i;
    for (int32_t i = 0; i < n; ++i) lcp[i] = 0;
    int32_t h = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (rank[i] >= 1) {
            int32_t j = sa[rank[i] - 1];
            /* the reference relies on the unique last terminator to stop; the bound only
               guards malformed input */
            while (i + h < n && j + h < n && T[i + h] == T[j + h]) ++h;
            lcp[rank[i]] = h;
            if (h > 0) --h;
        }
    }
    free(rank);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * east/asts/easa.py:268-287 _compute_childtab (Abouelhoda et al.): one stack pass.
 * ---------------------------------------------------------------------------------------- */
int oracle_childtab(const int32_t *lcp, int32_t n, int32_t *up, int32_t *down)
{
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    if (!stack) return -1;
    int32_t sp = 0, last_index = -1;
    stack[sp++] = 0;
    for (int32_t i = 0; i < n; ++i) { up[i] = 0; down[i] = 0; }
    for (int32_t i = 0; i < n; ++i) {
        while (lcp[i] < lcp[stack[sp - 1]]) {
            last_index = stack[--sp];
            if (lcp[i] <= lcp[stack[sp - 1]] && lcp[stack[sp - 1]] != lcp[last_index])
                down[stack[sp - 1]] = last_index;
        }
        if (last_index != -1) {
            up[i] = last_index;
            last_index = -1;
        }
        stack[sp++] = i;
    }
    free(stack);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * east/asts/easa.py:289-304 _compute_childtab_next_l_index
 * ---------------------------------------------------------------------------------------- */
int oracle_next_l_index(const int32_t *lcp, int32_t n, int32_t *next)
{
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    if (!stack) return -1;
    int32_t sp = 0;
    stack[sp++] = 0;
    for (int32_t i = 0; i < n; ++i) next[i] = 0;
    for (int32_t i = 0; i < n; ++i) {
        while (lcp[i] < lcp[stack[sp - 1]]) --sp;
        if (lcp[i] == lcp[stack[sp - 1]]) {
            int32_t last_index = stack[--sp];
            next[last_index] = i;
        }
        stack[sp++] = i;
    }
    free(stack);
    return 0;
}

/* east/asts/utils.py:6-11 index(): linear scan for key from start (no bounds check there) */
static int32_t ref_index(const int32_t *array, int32_t n, int32_t key, int32_t start)
{
    int32_t i = start;
    while (i < n && array[i] != key) ++i;
    return i;
}

/* ------------------------------------------------------------------------------------------
 * east/asts/easa.py:306-331 _compute_anntab, driven by the bottom-up traversal of
 * east/asts/easa.py:57-85 traverse_depth_first_post_order, with east/asts/easa.py:333-338
 * _interval_index.  Each stack frame keeps what process_node needs from the children list:
 * a running cursor `i` and the running sum.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t l, i, j; int32_t have_children; int32_t cursor; int64_t acc; } frame_t;

static void ann_attach_child(frame_t *parent, const frame_t *child, int64_t child_ann)
{
    /* process_node body for one child (easa.py:319-323) */
    if (parent->cursor < child->i) parent->acc += child->i - parent->cursor;
    parent->acc += child_ann;
    parent->cursor = child->j + 1;
}

int oracle_anntab(const int32_t *lcp, int32_t n, int32_t m, int32_t *ann)
{
    frame_t *stack = (frame_t *)malloc(sizeof(frame_t) * ((size_t)n + 2));
    if (!stack) return -1;
    for (int32_t i = 0; i < n; ++i) ann[i] = 0;
    int32_t sp = 0;
    frame_t root = {0, 0, -1, 0, 0, 0};
    stack[sp++] = root;
    frame_t last; int have_last = 0;
    for (int32_t i = 1; i < n; ++i) {
        int32_t lb = i - 1;
        while (lcp[i] < stack[sp - 1].l) {
            stack[sp - 1].j = i - 1;
            last = stack[--sp]; have_last = 1;
            /* callback(last): finish process_node (easa.py:324-325) */
            {
                int64_t acc = last.acc;
                if (last.cursor <= last.j) acc += last.j - last.cursor + 1;
                ann[ref_index(lcp, n, last.l, last.i)] += (int32_t)acc;
            }
            lb = last.i;
            if (lcp[i] <= stack[sp - 1].l) {
                ann_attach_child(&stack[sp - 1], &last, ann[ref_index(lcp, n, last.l, last.i)]);
                have_last = 0;
            }
        }
        if (lcp[i] > stack[sp - 1].l) {
            frame_t f = {lcp[i], lb, -1, 0, lb, 0};
            if (have_last) {
                ann_attach_child(&f, &last, ann[ref_index(lcp, n, last.l, last.i)]);
                have_last = 0;
            }
            stack[sp++] = f;
        }
    }
    /* easa.py:84-85: only the stack top is reported at the end */
    stack[sp - 1].j = n - 1;
    {
        frame_t top = stack[sp - 1];
        int64_t acc = top.acc;
        if (top.cursor <= top.j) acc += top.j - top.cursor + 1;
        ann[ref_index(lcp, n, top.l, top.i)] += (int32_t)acc;
    }
    /* easa.py:329 remove the degenerate first-level terminator leaves */
    ann[0] -= m;
    free(stack);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Scoring: east/asts/easa.py:91-139 _score, :379-400 _get_child_interval,
 * :349-356 _lcp_value, :340-347 _annotation/_is_leaf, east/asts/utils.py:14-22 match_strings.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint32_t *T; int32_t n, m;
    const int32_t *sa, *lcp, *up, *down, *next, *ann;
    int64_t probes; /* child-table cells + text cells touched, for the byte model */
    int error;
} easa_t;

typedef struct { int32_t l, i, j; int valid; } interval_t;

static int32_t lcp_value(easa_t *e, int32_t i, int32_t j)
{
    int32_t n = e->n;
    if ((i == 0 || i == n - 1) && j == n - 1) return 0;
    if (j + 1 >= n) { e->error = 1; return 0; } /* reference raises IndexError here */
    if (i < e->up[j + 1] && e->up[j + 1] <= j) return e->lcp[e->up[j + 1]];
    return e->lcp[e->down[i]];
}

static interval_t get_child_interval(easa_t *e, int32_t i, int32_t j, uint32_t ch)
{
    interval_t none = {0, 0, 0, 0};
    if (i == j) return none;
    int32_t n = e->n;
    int32_t l = lcp_value(e, i, j);
    int32_t i1;
    if (e->error) return none;
    if (i == 0 && j == n - 1) {
        i1 = 0;
    } else {
        if (i < e->up[j + 1]) i1 = e->up[j + 1]; else i1 = e->down[i];
        e->probes++;
        if ((int64_t)e->sa[i] + l < n && e->T[e->sa[i] + l] == ch) {
            interval_t r = {lcp_value(e, i, i1 - 1), i, i1 - 1, 1};
            return r;
        }
    }
    while (e->next[i1] != 0) {
        int32_t i2 = e->next[i1];
        e->probes++;
        if ((int64_t)e->sa[i1] + l < n && e->T[e->sa[i1] + l] == ch) {
            interval_t r = {lcp_value(e, i1, i2 - 1), i1, i2 - 1, 1};
            return r;
        }
        i1 = i2;
    }
    e->probes++;
    if ((int64_t)e->sa[i1] + l < n && e->T[e->sa[i1] + l] == ch) {
        interval_t r = {lcp_value(e, i1, j), i1, j, 1};
        return r;
    }
    return none;
}

static int32_t annotation(easa_t *e, interval_t v)
{
    if (v.i == v.j) return 1;
    return e->ann[ref_index(e->lcp, e->n, v.l, v.i)];
}

/* suffix_scores: NULL or array of L doubles (per suffix start; the reference's dict keyed by
 * suffix string is rebuilt on the Python side).  Returns the score; *status = 0 ok,
 * 1 = ZeroDivisionError (empty query), 2 = IndexError path of the reference. */
double oracle_score(const uint32_t *T, int32_t n, int32_t m,
                    const int32_t *sa, const int32_t *lcp, const int32_t *up, const int32_t *down,
                    const int32_t *next, const int32_t *ann,
                    const uint32_t *q, int32_t L, int normalized,
                    double *suffix_scores, int64_t *probes, int *status)
{
    easa_t e = {T, n, m, sa, lcp, up, down, next, ann, 0, 0};
    if (status) *status = 0;
    if (L <= 0) { if (status) *status = 1; return 0.0; }
    double result = 0.0;
    interval_t root = {0, 0, n - 1, 1};
    for (int32_t s = 0; s < L; ++s) {
        const uint32_t *suffix = q + s;
        int32_t rem = L - s;
        double suffix_score = 0.0, suffix_result = 0.0;
        int32_t matched_chars = 0, nodes_matched = 0;
        interval_t parent = root;
        interval_t child = get_child_interval(&e, parent.i, parent.j, suffix[0]);
        while (child.valid) {
            nodes_matched++;
            int64_t substr_start = (int64_t)sa[child.i] + parent.l;
            int64_t substr_end;
            if (child.i == child.j) substr_end = n;
            else substr_end = substr_start + child.l - parent.l;
            int32_t match = 0;
            int64_t span = substr_end - substr_start;
            int32_t lim = rem < span ? rem : (int32_t)span;
            while (match < lim && suffix[match] == T[substr_start + match]) { ++match; e.probes++; }
            suffix_score += (double)annotation(&e, child) / (double)annotation(&e, parent);
            matched_chars += match;
            suffix += match; rem -= match;
            if (rem > 0 && match == span) {
                parent = child;
                child = get_child_interval(&e, parent.i, parent.j, suffix[0]);
            } else {
                break;
            }
        }
        if (matched_chars) {
            suffix_result = (suffix_score + matched_chars) - nodes_matched;
            if (normalized) suffix_result /= matched_chars;
            result += suffix_result;
        }
        if (suffix_scores) suffix_scores[s] = suffix_result;
    }
    result /= L;
    if (probes) *probes = e.probes;
    if (e.error && status) *status = 2;
    return result;
}

/* Convenience: full build of one document (pipeline of east/asts/easa.py:16-24). */
int oracle_build(const uint32_t *T, int32_t n, int32_t m,
                 int32_t *sa, int32_t *lcp, int32_t *up, int32_t *down, int32_t *next, int32_t *ann)
{
    if (oracle_suffix_array(T, n, sa)) return -1;
    if (oracle_lcp(T, n, sa, lcp)) return -1;
    if (oracle_childtab(lcp, n, up, down)) return -1;
    if (oracle_next_l_index(lcp, n, next)) return -1;
    if (oracle_anntab(lcp, n, m, ann)) return -1;
    return 0;
}

/* Score K keyphrases against one built document (relevance.py:51-53 loop body). */
int oracle_score_many(const uint32_t *T, int32_t n, int32_t m,
                      const int32_t *sa, const int32_t *lcp, const int32_t *up, const int32_t *down,
                      const int32_t *next, const int32_t *ann,
                      const uint32_t *kp, const int64_t *kp_off, int32_t K, int normalized,
                      double *out, int64_t *probes_total)
{
    int64_t tot = 0;
    int rc = 0;
    for (int32_t k = 0; k < K; ++k) {
        int64_t p = 0; int st = 0;
        out[k] = oracle_score(T, n, m, sa, lcp, up, down, next, ann,
                              kp + kp_off[k], (int32_t)(kp_off[k + 1] - kp_off[k]), normalized,
                              NULL, &p, &st);
        tot += p;
        if (st) rc = st;
    }
    if (probes_total) *probes_total = tot;
    return rc;
}
