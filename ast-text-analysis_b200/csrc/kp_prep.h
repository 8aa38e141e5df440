// kp_prep.h -- device-side preparation of the keyphrases for the scorer (kp_prep.cu).
#pragma once
#include "sa_build.h"

namespace east {

struct KpDevice {
    int32_t total = 0, K = 0;
    bool dedup = false;
    int64_t n_uniq = -1;               // host copy, valid after kp_stage2
    std::vector<int32_t> off32;        // K + 1 (source of the upload)
    DevBuf<int32_t> d_off, d_uniq_of;  // K + 1 offsets; per suffix: position of its distinct twin in visiting order
    DevBuf<SufRec> d_recs;             // one record per distinct suffix in visiting order (n_uniq of `total` used)
    DevBuf<uint8_t> d_q8;              // dense codes of the keyphrases
    std::vector<uint8_t> encoded_for;  // the code table stage 1 has already made the dense codes for (one-kernel variant), or empty
    DevBuf<uint32_t> d_n_uniq;
    uint32_t *n_uniq_host = nullptr;   // pinned word the count is copied to
    cudaEvent_t done = nullptr;        // stage 1 complete (recorded on its stream)
    KpDevice() {}
    KpDevice(const KpDevice &) = delete;
    KpDevice &operator=(const KpDevice &) = delete;
    ~KpDevice();
};

// Stage 1 (independent of any index): hashes, lexicographic order, groups of identical suffixes -> d_off, d_uniq_of,
// d_recs (without dense codes), the distinct count on its way to the host.  Everything is queued on `s`; nothing waits.
// dedup = false: every suffix is its own group, in suffix order (per-suffix results wanted).
// Up to small_max suffixes (at most 64 Ki) the whole stage is one kernel (a cluster of 8 CTAs).
void kp_stage1(KpDevice &kp, const uint32_t *kp_dev, const uint32_t *kp_host /* optional */, const int64_t *kp_off_host, int32_t K, bool dedup,
               cudaStream_t s, int32_t small_max = 1 << 30, const uint8_t *likely_code_table = nullptr);
// likely_code_table: the code table the index is expected to have (a guess, or the table of an existing index): the
// one-kernel variant then writes the dense codes too and kp_stage2 for the same table has nothing left to launch.
// Stage 2 (per index alphabet; may be repeated): dense byte codes.  code_table_host = NULL: the index has no fast
// path, every suffix takes the generic walk.  Queued on `s` behind stage 1; returns with kp.n_uniq known.
void kp_stage2(KpDevice &kp, const uint32_t *kp_dev, const uint8_t *code_table_host, cudaStream_t s);

}  // namespace east
