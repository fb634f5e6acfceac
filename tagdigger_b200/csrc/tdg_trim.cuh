// Trim decision of the barcode splitter: where does the genomic part of a read
// end?  One warp per read, positions and candidates spread over the lanes and
// combined with warp ballots (sm_100a).
//
// Replaces findAdapterSeq, /root/reference/tagdigger_fun.py:1251-1283, with the
// tables of build_adapter_tree (:1208-1249) prepared by the host
// (tagdigger_b200/trimming.py):
//   rs0 = sequence.find(fullsite0, searchstart); rs1 = sequence.find(fullsite1, searchstart)
//   neither: the reversed read is looked up in a trie of reversed adapter
//            prefixes -> "the read ENDS WITH a prefix of (site remnant + adapter)"
//            -> a negative slice index, else 999
//   else:    the earlier site wins (the rare cutter on a tie): position + len(site)
// The trie's reachable set is prefix-free, so at most one candidate can match and
// the walk becomes one string compare per candidate length.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tdg {

constexpr int TRIM_THREADS = 256;          // 8 reads per CTA
constexpr int TRIM_NONE = 999;             // the reference's "no 3' trim" value

struct TrimCand {            // one reachable adapter prefix of one barcode
    uint16_t len;            // characters of the prefix
    int16_t  idx;            // slice index the reference returns for it
    uint32_t which;          // 0: common-cutter string, 1: this barcode's rare-cutter string
};

struct TrimArgs {
    const uint8_t *seqs;             // concatenated sequence lines (already stripped; any case)
    const unsigned long long *off;   // [n + 1]
    const int32_t *bar;              // [n] barcode index of each read
    const uint32_t *start;           // [n] searchstart (barcode length + cut-site length)
    uint32_t n;
    int32_t *out;                    // [n] slice2
    const uint8_t *site0, *site1;    // full restriction sites (upper case)
    uint32_t len0, len1;
    const uint8_t *a0;               // common-cutter remnant + adapter
    const uint8_t *a1;               // concatenated per-barcode rare-cutter remnant + adapter
    const uint32_t *a1_off;          // [nbar + 1]
    const TrimCand *cand;            // concatenated per-barcode candidate lists
    const uint32_t *cand_off;        // [nbar + 1]
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t fold_upper(uint32_t c)      // str.upper() for ASCII
{
    return (c >= 'a' && c <= 'z') ? c - 32u : c;
}

__global__ void __launch_bounds__(TRIM_THREADS) trim_kernel(const TrimArgs t)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * (TRIM_THREADS / 32) + (threadIdx.x >> 5);
    if (r >= t.n) return;
    const uint8_t *s = t.seqs + t.off[r];
    const uint32_t n = (uint32_t)(t.off[r + 1] - t.off[r]);
    const uint32_t start = t.start[r];
    int32_t result = TRIM_NONE;
    bool done = false;

    // ---- full restriction sites, 32 positions per step --------------------------
    for (uint32_t base = start; base < n && !done; base += 32) {
        const uint32_t p = base + lane;
        bool m0 = p + t.len0 <= n && t.len0 > 0, m1 = p + t.len1 <= n && t.len1 > 0;
        for (uint32_t i = 0; m0 && i < t.len0; i++) m0 = fold_upper(s[p + i]) == t.site0[i];
        for (uint32_t i = 0; m1 && i < t.len1; i++) m1 = fold_upper(s[p + i]) == t.site1[i];
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, m0), b1 = __ballot_sync(0xFFFFFFFFu, m1);
        if (b0 | b1) {
            const uint32_t p0 = b0 ? base + (uint32_t)__ffs(b0) - 1u : 0xFFFFFFFFu;
            const uint32_t p1 = b1 ? base + (uint32_t)__ffs(b1) - 1u : 0xFFFFFFFFu;
            result = p0 < p1 ? (int32_t)(p0 + t.len0) : (int32_t)(p1 + t.len1);
            done = true;
        }
    }
    // ---- adapter at the very end of the read -------------------------------------
    if (!done) {
        const int32_t b = t.bar[r];
        const uint32_t lo = t.cand_off[b], hi = t.cand_off[b + 1];
        const uint8_t *a1 = t.a1 + t.a1_off[b];
        for (uint32_t cbase = lo; cbase < hi && !done; cbase += 32) {
            const uint32_t c = cbase + lane;
            bool hit = false;
            int32_t idx = 0;
            if (c < hi) {
                const TrimCand tc = t.cand[c];
                const uint32_t L = tc.len;
                if (L <= n) {
                    const uint8_t *a = tc.which ? a1 : t.a0;
                    const uint8_t *tail = s + (n - L);
                    hit = true;
                    for (uint32_t i = 0; hit && i < L; i++) hit = fold_upper(tail[i]) == a[i];
                    idx = tc.idx;
                }
            }
            const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
            if (hits) {
                result = __shfl_sync(0xFFFFFFFFu, idx, __ffs(hits) - 1);
                done = true;
            }
        }
    }
    if (lane == 0) t.out[r] = result;
}

#endif  // __CUDACC__

}  // namespace tdg
