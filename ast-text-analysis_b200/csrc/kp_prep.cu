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
#include <cstdlib>
#include "kp_prep.h"
#include "radix_sort.cuh"

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

// ONE 64-bit sort key per suffix: its first symbols (sym_bits bits each: code point + 1, clamped; 0 = past the end of the
// keyphrase) in the high lex_bits, hash bits below.  Sorted by it the suffixes are lexicographic by their first symbols
// and identical suffixes are adjacent; different suffixes that share the key only cost a redundant walk.
__global__ void __launch_bounds__(256)
k_kp_sortkeys(const uint32_t *__restrict__ kp, const int32_t *__restrict__ send, const uint64_t *__restrict__ hash, int32_t total,
              int sym_bits, int n_sym, int key_bits, uint64_t *__restrict__ keys) {
    const uint32_t top = (1u << sym_bits) - 1u;
    const int hash_bits = key_bits - n_sym * sym_bits;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int32_t e = send[p];
        uint64_t w = 0ull;
        for (int q = 0; q < n_sym; ++q) w = (w << sym_bits) | (uint64_t)((p + q < e) ? min(kp[p + q] + 1u, top) : 0u);
        keys[p] = (w << hash_bits) | (hash[p] >> (64 - hash_bits));
    }
}

// block-wide inclusive scan of one 0/1 flag per thread (KP_THREADS threads); returns the inclusive count, *total = block sum
__device__ __forceinline__ uint32_t kp_block_scan(uint32_t flag, uint32_t *s_warp, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag != 0u);
    const uint32_t incl = __popc(ballot & (0xffffffffu >> (31 - lane)));
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, sum = 0;
    for (int i = 0; i < KP_THREADS / 32; ++i) { const uint32_t v = s_warp[i]; if (i < warp) base += v; sum += v; }
    *total = sum;
    return base + incl;
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

// index-dependent part: dense byte codes (0 = the code point does not occur in the batch, or is >= 0x0A00)
__global__ void __launch_bounds__(256)
k_kp_encode(const uint32_t *__restrict__ kp, int32_t total, const uint8_t *__restrict__ table /* NULL: no fast path */,
            const uint32_t *__restrict__ n_uniq, SufRec *__restrict__ recs, uint8_t *__restrict__ q8) {
    const int32_t nu = (int32_t)*n_uniq;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(total + 16, nu); i += gridDim.x * blockDim.x) {
        if (table && i < total + 16) {
            uint8_t c = 0;
            if (i < total) { const uint32_t cp = kp[i]; if (cp < EAST_TERM_BASE) c = table[cp]; }
            q8[i] = c;   // the scorer reads the queries 8 bytes at a time: 16 zero bytes of slack
        }
        if (i < nu) {
            SufRec r = recs[i];
            r.q8_first = 0ull;
            if (table) {
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

void kp_stage1(KpDevice &kp, const uint32_t *kp_dev, const uint32_t *kp_host, const int64_t *kp_off, int32_t K, bool dedup, cudaStream_t s) {
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

    DevBuf<uint64_t> hash((size_t)total, s), keys_a, keys_b;
    DevBuf<uint32_t> vals_a((size_t)total, s), vals_b;
    DevBuf<int32_t> send((size_t)total, s);
    DevBuf<uint8_t> weird((size_t)total, s), flags((size_t)total, s);
    const int nb = (total + KP_THREADS - 1) / KP_THREADS;
    DevBuf<uint32_t> bsum((size_t)nb, s);
    EAST_LAUNCH(k_kp_suffix_keys, grid_for(K, 256, 8), 256, 0, s, kp_dev, kp.d_off.p, K, hash.p, vals_a.p, send.p, weird.p);
    const uint32_t *order = vals_a.p;
    if (dedup) {
        // symbol width from the largest code point (the host has the keyphrases; without a host copy: the full 12 bits);
        // 42 bits of symbols (6 seven-bit symbols for A-Z) + 22 hash bits: one 8-pass sort
        int sym_bits = KP_MAX_SYM_BITS;
        if (kp_host) {
            uint32_t mx = 0;
            for (int32_t p = 0; p < total; ++p) mx = std::max(mx, kp_host[p]);
            sym_bits = std::min(KP_MAX_SYM_BITS, std::max(1, bits_for((uint64_t)mx + 1)));
        }
        static const int key_bits = getenv("EAST_KP_KEY_BITS") ? atoi(getenv("EAST_KP_KEY_BITS")) : 64;
        const int n_sym = std::max(1, (key_bits - 20) / sym_bits);
        keys_a = DevBuf<uint64_t>((size_t)total, s); keys_b = DevBuf<uint64_t>((size_t)total, s);
        vals_b = DevBuf<uint32_t>((size_t)total, s);
        DevBuf<uint32_t> hist(256 * RS_MAX_PASSES, s);
        DevBuf<uint8_t> scratch(rs_scratch_bytes(total, RS_MAX_PASSES), s);
        EAST_LAUNCH(k_kp_sortkeys, grid_for(total, 256, 8), 256, 0, s, kp_dev, send.p, hash.p, total, sym_bits, n_sym, key_bits, keys_a.p);
        const int cur = radix_sort_pairs(keys_a.p, keys_b.p, vals_a.p, vals_b.p, total, key_bits, hist.p, false, scratch.p, s);
        order = cur ? vals_b.p : vals_a.p;
    }
    EAST_LAUNCH(k_kp_mark, nb, KP_THREADS, 0, s, kp_dev, hash.p, send.p, order, total, dedup ? 1 : 0, flags.p, bsum.p);
    EAST_LAUNCH(k_kp_scan_blocks, 1, KP_THREADS, 0, s, bsum.p, nb, kp.d_n_uniq.p);
    EAST_LAUNCH(k_kp_emit, nb, KP_THREADS, 0, s, send.p, weird.p, order, flags.p, bsum.p, total, kp.d_uniq_of.p, kp.d_recs.p);
    if (!g_pinned_words) EAST_CUDA(cudaHostAlloc((void **)&g_pinned_words, sizeof(uint32_t) * KP_PINNED_WORDS, cudaHostAllocDefault));
    kp.n_uniq_host = g_pinned_words + (g_pinned_next++ % KP_PINNED_WORDS);
    EAST_CUDA(cudaMemcpyAsync(kp.n_uniq_host, kp.d_n_uniq.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (!kp.done) EAST_CUDA(cudaEventCreateWithFlags(&kp.done, cudaEventDisableTiming));
    EAST_CUDA(cudaEventRecord(kp.done, s));
    // the scratch buffers above go back to the pool in stream order
}

void kp_stage2(KpDevice &kp, const uint32_t *kp_dev, const uint8_t *code_table_host, cudaStream_t s) {
    EAST_CUDA(cudaStreamWaitEvent(s, kp.done, 0));
    const bool fast = code_table_host != nullptr;
    if (fast) {
        if (!kp.d_table.p) kp.d_table = DevBuf<uint8_t>(EAST_TERM_BASE, s);
        if (!kp.d_q8.p) kp.d_q8 = DevBuf<uint8_t>((size_t)kp.total + 16, s);
        EAST_CUDA(cudaMemcpyAsync(kp.d_table.p, code_table_host, EAST_TERM_BASE, cudaMemcpyHostToDevice, s));
    }
    EAST_LAUNCH(k_kp_encode, grid_for(kp.total + 16, 256, 8), 256, 0, s, kp_dev, kp.total, fast ? kp.d_table.p : (const uint8_t *)nullptr,
                kp.d_n_uniq.p, kp.d_recs.p, kp.d_q8.p);
    if (kp.n_uniq < 0) {
        EAST_CUDA(cudaEventSynchronize(kp.done));
        kp.n_uniq = (int64_t)*kp.n_uniq_host;
    }
}

KpDevice::~KpDevice() { if (done) cudaEventDestroy(done); }

}  // namespace east
