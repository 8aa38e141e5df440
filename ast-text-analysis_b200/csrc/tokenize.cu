// tokenize.cu -- the host preprocessing of the reference on the device, for ASCII and Cyrillic texts.
//
// Replaces east/utils.py:31-79 (prepare_text, tokenize, text_to_strings_collection) + east/asts/utils.py:25-40
// (make_unique_endings) + the join of east/asts/easa.py:19 for a batch of raw UTF-8 texts: upper-case, maximal runs of
// [\w'] characters, tokens of fewer than 3 characters and all-digit tokens dropped, every 3 consecutive tokens joined
// into one string, string i followed by the terminator 0x0A00 + i; a text without a usable token becomes [" "].
// Output = the packed uint32 documents the build consumes.  Host Python needs 1.8 s for the 1 000 x 50 KB documents the
// device then indexes and scores in 4 ms; this kernel pair needs tens of microseconds.
//
// Character set handled here: U+0000-007F and the Cyrillic block U+0400-045F (2-byte UTF-8, lead bytes D0 / D1):
//   word characters   0-9 A-Z a-z _ ' and every Cyrillic letter of the block
//   upper()           a-z -> A-Z;  U+0430-044F -> U+0410-042F;  U+0450-045F -> U+0400-040F  (all 1:1)
//   isdigit()         0-9
// plus, as separators only (not word characters, no case mapping): the controls and signs of U+0080-00BF that Python does
// not count as alphanumeric (no-break space, guillemets, section / degree / copyright signs ...), General Punctuation
// U+2000-206F (dashes, typographic quotation marks, ellipsis), the numero sign U+2116 and the byte order mark U+FEFF --
// what Russian and English running text actually contains besides letters.
// Any other byte (another lead byte, a stray continuation byte, invalid UTF-8) flags the text: the caller sends the
// collection through the host preprocessing instead (Python's full Unicode tables).
//
// One CTA per text, thread t owns the tokens that START in its chunk of ceil(bytes / threads) consecutive bytes and
// walks each of them to its end (also beyond the chunk): length and all-digit flag decide whether the token is kept.
//   pass 1   kept tokens / kept characters per thread; block-wide exclusive scan -> index j of every kept token among the
//            text's kept tokens and the number of kept characters before it
//   k_tokenize<false> stops here and reports the size of the packed document (n, m); the host turns the sizes into
//            document offsets (they are also the build's arguments)
//   pass 2   (k_tokenize<true>) token j goes to offset chars_before(j) + j / 3 of its document (one terminator per
//            completed group of 3), followed by the terminator 0x0A00 + j / 3 when it closes a group or the text
#include "sa_build.h"

namespace east {

constexpr int TK_THREADS = 512;

struct TkChar { uint32_t cp; int len; bool word, digit, bad; };

// U+0080 + i, i = 0 .. 63: bit i set = the character is NOT a word character and has no case mapping (controls, no-break
// space, currency and typographic signs, the guillemets ...); clear = ª ² ³ µ ¹ º ¼ ½ ¾, which Python counts as
// alphanumeric (or upper-cases into another block): those flag the text.  Checked against Python's re / str.upper for
// every code point in tests/test_gpu_size.py.
constexpr uint64_t TK_LATIN1_SEPARATORS = 0x89d3fbffffffffffull;

// class of a byte: 1 word character, 2 digit, 4 lower-case letter (ASCII); 0x80 = part of a multi-byte sequence.
// One shared-memory lookup instead of a dozen comparisons per byte (the walks decode every kept character three times).
constexpr uint32_t TKC_WORD = 1u, TKC_DIGIT = 2u, TKC_LOWER = 4u, TKC_MULTI = 0x80u;
__device__ __forceinline__ void tk_fill_classes(uint8_t *lut) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) {
        uint32_t cls = 0;
        if (b >= 0x80) cls = TKC_MULTI;
        else {
            const bool lower = b >= 'a' && b <= 'z', digit = b >= '0' && b <= '9';
            if (lower || digit || (b >= 'A' && b <= 'Z') || b == '_' || b == '\'') cls |= TKC_WORD;
            if (digit) cls |= TKC_DIGIT;
            if (lower) cls |= TKC_LOWER;
        }
        lut[b] = (uint8_t)cls;
    }
}

// the code point that starts at byte i (i is not a continuation byte), upper-cased
__device__ __forceinline__ TkChar tk_decode(const uint8_t *__restrict__ p, int64_t i, int64_t end, const uint8_t *lut) {
    TkChar c;
    const uint32_t b = p[i];
    c.len = 1; c.bad = false; c.digit = false;
    const uint32_t cls = lut[b];
    if (!(cls & TKC_MULTI)) {
        c.digit = (cls & TKC_DIGIT) != 0u;
        c.word = (cls & TKC_WORD) != 0u;
        c.cp = b - ((cls & TKC_LOWER) << 3);   // lower case: - 32
        return c;
    }
    const uint32_t b1 = (i + 1 < end) ? p[i + 1] : 0u;
    if ((b == 0xd0u || b == 0xd1u) && (b1 & 0xc0u) == 0x80u) {          // U+0400 .. 047F: the Cyrillic block up to 045F
        uint32_t cp = ((b & 0x1fu) << 6) | (b1 & 0x3fu);
        c.len = 2;
        c.word = true;
        if (cp > 0x45fu) c.bad = true;
        else if (cp >= 0x450u) cp -= 0x50u;
        else if (cp >= 0x430u) cp -= 0x20u;
        c.cp = cp;
        return c;
    }
    if (b == 0xc2u && (b1 & 0xc0u) == 0x80u) {                           // U+0080 .. 00BF: separators only
        c.len = 2; c.word = false; c.cp = 0x80u + (b1 & 0x3fu);
        c.bad = ((TK_LATIN1_SEPARATORS >> (b1 & 0x3fu)) & 1ull) == 0ull;
        return c;
    }
    const uint32_t b2 = (i + 2 < end) ? p[i + 2] : 0u;
    if ((b == 0xe2u || b == 0xefu) && (b1 & 0xc0u) == 0x80u && (b2 & 0xc0u) == 0x80u) {
        // three bytes: General Punctuation U+2000 .. 206F (dashes, quotation marks, ellipsis, spaces), the numero sign
        // U+2116 and the byte order mark U+FEFF -- none of them a word character, none with a case mapping
        const uint32_t cp = ((b & 0x0fu) << 12) | ((b1 & 0x3fu) << 6) | (b2 & 0x3fu);
        c.len = 3; c.word = false; c.cp = cp;
        c.bad = !((cp >= 0x2000u && cp <= 0x206fu) || cp == 0x2116u || cp == 0xfeffu);
        return c;
    }
    c.bad = true; c.word = false; c.cp = 0xfffdu;
    return c;
}

// is the code point that ENDS at byte i - 1 a word character (i > begin)
__device__ __forceinline__ bool tk_prev_is_word(const uint8_t *__restrict__ p, int64_t i, int64_t begin, const uint8_t *lut) {
    const uint32_t b = p[i - 1];
    const uint32_t cls = lut[b];
    if (!(cls & TKC_MULTI)) return (cls & TKC_WORD) != 0u;
    // the last byte of a multi-byte character: only the Cyrillic letters (lead bytes D0 / D1, two bytes) are word characters
    if (i - 2 < begin) return false;
    const uint32_t lead = p[i - 2];
    return lead == 0xd0u || lead == 0xd1u;
}

// a continuation byte at i: does a lead byte the tokenizer accepts claim it (anything else flags the text)
__device__ __forceinline__ bool tk_continuation_ok(const uint8_t *__restrict__ p, int64_t i, int64_t begin) {
    if (i - 1 < begin) return false;
    const uint32_t a = p[i - 1];
    if (a >= 0xc0u) return true;                                         // second byte of a 2- or 3-byte sequence
    if ((a & 0xc0u) != 0x80u || i - 2 < begin) return false;
    return p[i - 2] >= 0xe0u;                                            // third byte of a 3-byte sequence
}

// block-wide exclusive scan of two counters per thread; totals to everybody
__device__ __forceinline__ void tk_scan2(uint32_t a, uint32_t b, uint32_t *s_wa, uint32_t *s_wb, uint32_t &ea, uint32_t &eb,
                                         uint32_t &ta, uint32_t &tb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t xa = a, xb = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ya; xb += yb; }
    }
    if (lane == 31) { s_wa[warp] = xa; s_wb[warp] = xb; }
    __syncthreads();
    uint32_t ba = 0, bb = 0;
    ta = 0; tb = 0;
    for (int w = 0; w < TK_THREADS / 32; ++w) {
        if (w < warp) { ba += s_wa[w]; bb += s_wb[w]; }
        ta += s_wa[w]; tb += s_wb[w];
    }
    ea = ba + xa - a; eb = bb + xb - b;
}

template <bool EMIT>
__global__ void __launch_bounds__(TK_THREADS)
k_tokenize(const uint8_t *__restrict__ raw, const int64_t *__restrict__ raw_off, int32_t n_texts,
           int32_t *__restrict__ sizes /* [n_texts][3]: n, m, unsupported */, const int64_t *__restrict__ doc_off,
           uint32_t *__restrict__ text) {
    __shared__ uint32_t s_wa[TK_THREADS / 32], s_wb[TK_THREADS / 32];
    __shared__ uint8_t s_cls[256];
    tk_fill_classes(s_cls);
    __syncthreads();
    const int d = blockIdx.x;
    const int64_t begin = raw_off[d], end = raw_off[d + 1];
    const int64_t len = end - begin;
    const int64_t per = (len + TK_THREADS - 1) / TK_THREADS;
    const int64_t c0 = begin + min(len, (int64_t)threadIdx.x * per), c1 = begin + min(len, ((int64_t)threadIdx.x + 1) * per);
    const uint8_t *__restrict__ p = raw;

    // ---- pass 1: the tokens that start in [c0, c1)
    uint32_t kept_tokens = 0, kept_chars = 0;
    bool bad = false;
    for (int64_t i = c0; i < c1;) {
        if ((p[i] & 0xc0u) == 0x80u) {   // continuation byte: belongs to the code point before it
            if (!tk_continuation_ok(p, i, begin)) bad = true;
            ++i;
            continue;
        }
        TkChar c = tk_decode(p, i, end, s_cls);
        bad = bad || c.bad;
        if (!c.word || (i > begin && tk_prev_is_word(p, i, begin, s_cls))) { i += c.len; continue; }
        // a token starts here: walk it to its end
        uint32_t tl = 0;
        bool all_digits = true;
        int64_t e = i;
        while (e < end) {
            c = tk_decode(p, e, end, s_cls);
            if (!c.word) break;
            bad = bad || c.bad;
            ++tl;
            all_digits = all_digits && c.digit;
            e += c.len;
        }
        if (tl > 2u && !all_digits) { ++kept_tokens; kept_chars += tl; }
        i = e;
    }
    uint32_t j0, ch0, J, CH;
    tk_scan2(kept_tokens, kept_chars, s_wa, s_wb, j0, ch0, J, CH);
    if (!EMIT) {
        if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) sizes[3 * d + 2] = 1;
        if (threadIdx.x == 0) {
            const uint32_t m = J ? (J + 2u) / 3u : 1u;   // no usable token: [" "]
            sizes[3 * d + 0] = (int32_t)(J ? CH + m : 2u);
            sizes[3 * d + 1] = (int32_t)m;
        }
        return;
    }
    // ---- pass 2: emit
    uint32_t *out = text + doc_off[d];
    if (J == 0u) {
        if (threadIdx.x == 0) { out[0] = (uint32_t)' '; out[1] = EAST_TERM_BASE; }
        return;
    }
    uint32_t j = j0, ch = ch0;
    for (int64_t i = c0; i < c1;) {
        if ((p[i] & 0xc0u) == 0x80u) { ++i; continue; }
        TkChar c = tk_decode(p, i, end, s_cls);
        if (!c.word || (i > begin && tk_prev_is_word(p, i, begin, s_cls))) { i += c.len; continue; }
        uint32_t tl = 0;
        bool all_digits = true;
        int64_t e = i;
        while (e < end) {
            c = tk_decode(p, e, end, s_cls);
            if (!c.word) break;
            ++tl;
            all_digits = all_digits && c.digit;
            e += c.len;
        }
        if (tl > 2u && !all_digits) {
            uint32_t *dst = out + ch + j / 3u;
            int64_t q = i;
            for (uint32_t x = 0; x < tl; ++x) {
                c = tk_decode(p, q, end, s_cls);
                dst[x] = c.cp;
                q += c.len;
            }
            if (j % 3u == 2u || j == J - 1u) dst[tl] = EAST_TERM_BASE + j / 3u;
            ++j;
            ch += tl;
        }
        i = e;
    }
}

void tokenize_texts(const uint8_t *raw_dev, const int64_t *raw_off_dev, int32_t n_texts, int32_t *sizes_dev, cudaStream_t s) {
    EAST_CUDA(cudaMemsetAsync(sizes_dev, 0, sizeof(int32_t) * 3 * (size_t)n_texts, s));
    EAST_LAUNCH(k_tokenize<false>, n_texts, TK_THREADS, 0, s, raw_dev, raw_off_dev, n_texts, sizes_dev, (const int64_t *)nullptr,
                (uint32_t *)nullptr);
}

void tokenize_emit(const uint8_t *raw_dev, const int64_t *raw_off_dev, int32_t n_texts, const int64_t *doc_off_dev, uint32_t *text_dev,
                   cudaStream_t s) {
    EAST_LAUNCH(k_tokenize<true>, n_texts, TK_THREADS, 0, s, raw_dev, raw_off_dev, n_texts, (int32_t *)nullptr, doc_off_dev, text_dev);
}

}  // namespace east
