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
#include "tdg_match.h"

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

// 32-character window read byte by byte from global memory (any alignment)
struct GlobalBytes {
    const uint8_t *p;
    uint32_t limit;
    __device__ __forceinline__ void load8(uint32_t off, uint32_t w[8]) const
    {
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            uint32_t v = 0;
            for (int k = 0; k < 4; k++) {
                uint32_t o = off + 4 * i + k;
                uint32_t c = o < limit ? p[o] : 0u;
                v |= c << (8 * k);
            }
            w[i] = v;
        }
    }
};

__device__ __forceinline__ uint32_t fold_upper(uint32_t c)      // str.upper() for ASCII
{
    return (c >= 'a' && c <= 'z') ? c - 32u : c;
}

// The decision for one read, computed by a whole warp (all lanes return the result).
__device__ __forceinline__ int32_t trim_decide(const TrimArgs &t, const uint8_t *s, uint32_t n, uint32_t start, int32_t b,
                                               uint32_t lane)
{
    int32_t result = TRIM_NONE;
    bool done = false;

    // ---- full restriction sites, 32 positions per step --------------------------
    for (uint32_t base = start; base < n && !done; base += 32) {
        const uint32_t p = base + lane;
        bool m0 = p + t.len0 <= n, m1 = p + t.len1 <= n;
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
    return result;
}

__global__ void __launch_bounds__(TRIM_THREADS) trim_kernel(const TrimArgs t)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * (TRIM_THREADS / 32) + (threadIdx.x >> 5);
    if (r >= t.n) return;
    const uint8_t *s = t.seqs + t.off[r];
    const uint32_t n = (uint32_t)(t.off[r + 1] - t.off[r]);
    const int32_t result = trim_decide(t, s, n, t.start[r], t.bar[r], lane);
    if (lane == 0) t.out[r] = result;
}

// Barcode splitter decisions for one batch of sequence lines: which barcode (the
// reference's sequence_index_lookup(sequence, barcuttree), tagdigger_fun.py:1333) and,
// for reads that have one, where to cut (findAdapterSeq with searchstart = barcode
// length + cut-site length, :1337-1339).  bar_out = -1: no barcode.
struct SplitArgs {
    TrimArgs t;                      // t.bar / t.start unused; t.out = slice2
    const BarTable *bar;             // barcode+cutsite table of the file (rows = barcode indices)
    const uint32_t *bar_len;         // [nbar] barcode lengths
    uint32_t cutlen;
    int32_t *bar_out;                // [n]
};

__global__ void __launch_bounds__(TRIM_THREADS) split_kernel(const SplitArgs a)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * (TRIM_THREADS / 32) + (threadIdx.x >> 5);
    if (r >= a.t.n) return;
    const uint8_t *s = a.t.seqs + a.t.off[r];
    const uint32_t n = (uint32_t)(a.t.off[r + 1] - a.t.off[r]);
    // barcode + cut site at the start of the read: the counting path's lookup, barcode part
    int32_t b = -1;
    {
        GlobalBytes f;
        f.p = s;
        f.limit = n;
        const BarEntry *bent = (const BarEntry *)((const uint8_t *)a.bar + sizeof(BarTable));
        b = match_barcode(f, a.bar, bent);
    }
    int32_t cut = TRIM_NONE;
    if (b >= 0) cut = trim_decide(a.t, s, n, a.bar_len[b] + a.cutlen, b, lane);
    if (lane == 0) {
        a.bar_out[r] = b;
        a.t.out[r] = cut;
    }
}

// Per-read results of the counting path's matcher for a batch of (stripped) sequence
// lines: row = barcode index or -1, col = tag column or -1.  Used where the host has to
// see individual reads (find_tags_fastq with tassel_tagcount=True adds a per-read weight
// taken from the header line, tagdigger_fun.py:251-253,264-265).  One thread per read.
struct MatchArgs {
    const uint8_t *seqs;
    const unsigned long long *off;   // [n + 1]
    uint32_t n;
    const BarTable *bar;
    TagTable tags;
    int32_t *row_out, *col_out;      // [n]
};

__global__ void __launch_bounds__(128) match_kernel(const MatchArgs a)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    GlobalBytes f;
    f.p = a.seqs + a.off[r];
    f.limit = (uint32_t)(a.off[r + 1] - a.off[r]);
    const BarEntry *bent = (const BarEntry *)((const uint8_t *)a.bar + sizeof(BarTable));
    const MatchResult m = match_line(f, a.bar, bent, a.tags);
    a.row_out[r] = m.row;
    a.col_out[r] = m.col;
}

#endif  // __CUDACC__

}  // namespace tdg
