// Packed pattern tables and the per-read match, shared by the CUDA kernel and
// the host-side table self test (same code, compiled for both).
//
// Reference semantics restated here (tagdigger_fun.py):
//   :115-134  sequence_index_lookup -- the unique stored pattern that prefixes
//             the query; fails on end of sequence or on any non-ACGT character;
//   :256-267  line.strip().upper(); barcode+cutsite lookup at offset 0; tag
//             lookup at barcutlen[barcode]; counts[bar][tag] += 1.
// Because the effective pattern sets are prefix-free (the host reproduces the
// trie builder's conflict rules, tagdigger_b200/matchset.py), "walk the trie"
// is replaced by exact-match probes of 2-bit packed keys.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TDG_HD __host__ __device__ __forceinline__
#else
#define TDG_HD inline
#endif

namespace tdg {

// 2-bit code of a base: (c >> 1) & 3  ->  A=0 C=1 T=2 G=3 (case-insensitive).
// Base i of a packed word lives in bits [2i+1 : 2i] (little-endian base order).
TDG_HD uint64_t lowmask(uint32_t nbases)   // mask of the first nbases (<=32) bases
{
    return nbases >= 32 ? ~0ull : ((1ull << (2 * nbases)) - 1ull);
}

struct BarEntry {            // 16 bytes
    uint64_t key;            // packed pattern (<= 32 bases)
    int32_t  row;            // matrix row
    uint16_t len;            // pattern length in bases
    uint16_t tag_off;        // where the tag comparison starts, from line start
};

struct BarTable {            // header of the device-resident barcode table
    uint32_t bucket[257];    // entries of bucket b (first 4 bases) are [bucket[b], bucket[b+1])
    uint32_t any_base;       // TDG_ANY_BASE: the single empty pattern
    int32_t  any_row;
    uint32_t any_tag_off;
    uint32_t n_entries;
    uint32_t max_len;        // longest pattern
    uint32_t max_tag_off;
    uint32_t pad;
    // followed by BarEntry[n_entries] (16-byte aligned: sizeof(BarTable) % 16 == 0)
};
static_assert(sizeof(BarTable) % 16 == 0, "BarEntry array must stay 16-byte aligned");

struct TagEntry {            // 32 bytes = one L2 sector
    uint64_t k0, k1;         // first 64 bases
    uint32_t len;            // TDG_EMPTY_LEN = empty slot
    int32_t  col;
    uint32_t ext;            // index into ext words for bases 64.. (len > 64)
    uint32_t pad;
};
#define TDG_EMPTY_LEN 0xFFFFFFFFu
// Flag in TagEntry::len of the FIRST slot of an even-aligned slot pair: some key whose
// probe sequence passes through this pair was stored beyond it, so a lookup that found
// both slots occupied without a match must go on.  Clear (the common case: a marker's
// two alleles fill their pair exactly) ends the lookup after one 64-byte line.
#define TDG_LEN_MORE  0x80000000u
#define TDG_LEN_MASK  0x7FFFFFFFu
#define TDG_MAX_CLASSES 4

struct TagClass {
    uint32_t K;              // prefix length hashed (1..32); every tag in the class has len >= K
    uint32_t base;           // first slot of this class in the entry array
    uint32_t mask;           // slots - 1 (power of two)
    uint32_t pad;
};

struct TagTable {
    const TagEntry *entries;
    const uint64_t *ext;
    TagClass cls[TDG_MAX_CLASSES];
    uint32_t n_classes;
    uint32_t any_base;       // single empty tag
    int32_t  any_col;
    uint32_t max_len;
    uint32_t min_len;
    uint32_t pad;
};

TDG_HD uint32_t tag_hash(uint64_t prefix)
{
    uint64_t x = prefix * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(x >> 32) ^ (uint32_t)x;
}
// First slot of a probe sequence: always even, so that one 64-byte line holds the
// first two slots of every sequence (the kernel fetches both in one round trip).
TDG_HD uint32_t tag_slot(uint64_t prefix, uint32_t mask)
{
    return tag_hash(prefix) & mask & ~1u;
}

// ---------------------------------------------------------------------------
// Turning characters into 2-bit codes, 4 at a time.
//   w: 4 consecutive characters (little endian).  Returns the 8-bit packed codes
//   and sets bad to a word whose byte k is non-zero iff character k is not one
//   of ACGTacgt.
TDG_HD uint32_t codes4(uint32_t w, uint32_t &bad)
{
    uint32_t cf = w & 0xDFDFDFDFu;                    // fold case (exact for ASCII letters)
    uint32_t c2 = (cf >> 1) & 0x03030303u;            // A0 C1 T2 G3
    uint32_t m  = (c2 >> 1) & ~c2 & 0x01010101u;      // 1 where code == 2 (T)
    uint32_t e  = 0x41414141u + 2u * c2 + 15u * m;    // the letter that code stands for
    bad = e ^ cf;
    return (c2 * 0x01041040u) >> 24;                  // c0 | c1<<2 | c2<<4 | c3<<6
}

// A Fetch provides
//   void load8(uint32_t off, uint32_t w[8]) -- the 32 characters that start off
//        bytes after the (whitespace-stripped) start of the line, any alignment;
//        characters beyond `limit` may hold anything;
//   uint32_t limit -- number of characters available from the line start up to
//        the end of the data.
template <class Fetch>
TDG_HD uint64_t pack32(const Fetch &f, uint32_t off, uint32_t &nvalid)
{
    uint32_t w[8];
    f.load8(off, w);
    uint64_t key = 0;
    uint32_t v = 32;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        uint32_t bad;
        uint32_t b = codes4(w[i], bad);
        key |= (uint64_t)b << (8 * i);
        if (bad != 0) {
            // index of the first bad character in this word
            uint32_t k = (bad & 0xFFu) ? 0 : (bad & 0xFF00u) ? 1 : (bad & 0xFF0000u) ? 2 : 3;
            v = 4 * i + k;
        }
    }
    uint32_t room = f.limit > off ? f.limit - off : 0;
    nvalid = v < room ? v : room;
    return key;
}

struct MatchResult {
    int32_t row;     // >= 0 when barcode+cutsite matched
    int32_t col;     // >= 0 when a tag matched as well
};

// Barcode + cut site at the start of the line: the row of the unique pattern that
// prefixes it (-1 if none); tag_off receives where the tag comparison starts.
template <class Fetch>
TDG_HD int32_t match_barcode(const Fetch &f, const BarTable *bar, const BarEntry *bent, uint32_t *tag_off_out = nullptr)
{
    uint32_t v0;
    uint64_t key0 = pack32(f, 0, v0);
    if (bar->any_base) {
        if (v0 < 1) return -1;
        if (tag_off_out) *tag_off_out = bar->any_tag_off;
        return bar->any_row;
    }
    // bucket = first four bases; a read with fewer valid bases looks in the
    // bucket its valid prefix padded with A would fall into (patterns
    // shorter than four bases are listed in every bucket they can start)
    uint32_t b = (uint32_t)key0 & 0xFFu & (uint32_t)lowmask(v0 < 4 ? v0 : 4);
    uint32_t lo = bar->bucket[b], hi = bar->bucket[b + 1];
    for (uint32_t e = lo; e < hi; e++) {
        BarEntry be = bent[e];
        if (be.len <= v0 && ((key0 ^ be.key) & lowmask(be.len)) == 0) {
            if (tag_off_out) *tag_off_out = be.tag_off;
            return be.row;
        }
    }
    return -1;
}

template <class Fetch>
TDG_HD MatchResult match_line(const Fetch &f, const BarTable *bar, const BarEntry *bent,
                              const TagTable &tt)
{
    MatchResult r;
    r.row = -1;
    r.col = -1;
    uint32_t tag_off = 0;
    r.row = match_barcode(f, bar, bent, &tag_off);
    if (r.row < 0) return r;
    // ---- tag at tag_off ----
    uint32_t tv0;
    uint64_t t0 = pack32(f, tag_off, tv0);
    if (tt.any_base) {
        if (tv0 >= 1) r.col = tt.any_col;
        return r;
    }
    if (tv0 < (tt.min_len < 32 ? tt.min_len : 32)) return r;
    uint32_t tv1 = 0;
    uint64_t t1 = 0;
    if (tv0 == 32 && tt.max_len > 32) t1 = pack32(f, tag_off + 32, tv1);
    uint32_t V = tv0 + tv1;                       // valid bases among the first 64
    for (uint32_t c = 0; c < tt.n_classes; c++) {
        TagClass tc = tt.cls[c];
        if (V < tc.K) continue;
        uint64_t km = lowmask(tc.K);
        uint64_t pre = t0 & km;
        uint32_t h = tag_slot(pre, tc.mask);
        for (;;) {
            TagEntry te = tt.entries[tc.base + h];
            if (te.len == TDG_EMPTY_LEN) break;
            if ((te.k0 & km) == pre) {
                uint32_t L = te.len & TDG_LEN_MASK;
                uint32_t L64 = L < 64 ? L : 64;
                bool ok = L64 <= V;
                if (ok) {
                    uint32_t a = L64 < 32 ? L64 : 32;
                    ok = ((t0 ^ te.k0) & lowmask(a)) == 0;
                    if (ok && L64 > 32) ok = ((t1 ^ te.k1) & lowmask(L64 - 32)) == 0;
                }
                if (ok && L > 64) {
                    uint32_t rest = L - 64, w = 0;
                    while (ok && rest > 0) {
                        uint32_t xv;
                        uint64_t x = pack32(f, tag_off + 64 + 32 * w, xv);
                        uint32_t a = rest < 32 ? rest : 32;
                        ok = xv >= a && ((x ^ tt.ext[te.ext + w]) & lowmask(a)) == 0;
                        rest -= a;
                        w++;
                    }
                }
                if (ok) { r.col = te.col; return r; }   // prefix-free set: at most one match
            }
            h = (h + 1) & tc.mask;
        }
    }
    return r;
}

// str.strip() whitespace (ASCII part) that can PRECEDE the sequence inside a
// line.  '\n' and '\r' are whitespace too, but they end the line (universal
// newlines): the scan for the first character must stop at them, so that an
// empty or blank line stays empty instead of running into the next line.
TDG_HD bool is_lead_space(uint32_t c)
{
    return c == 0x09 || c == 0x0b || c == 0x0c || (c >= 0x1c && c <= 0x20);
}

// Length in bytes of the UTF-8 encoded Unicode whitespace character that starts
// at p (0 if none).  str.strip() also removes these (the reference reads its
// files in text mode); they are all two or three bytes long:
//   U+0085 U+00A0 U+1680 U+2000-200A U+2028 U+2029 U+202F U+205F U+3000.
TDG_HD uint32_t utf8_space(uint32_t b0, uint32_t b1, uint32_t b2)
{
    if (b0 == 0xC2) return (b1 == 0x85 || b1 == 0xA0) ? 2u : 0u;
    if (b0 == 0xE1) return (b1 == 0x9A && b2 == 0x80) ? 3u : 0u;
    if (b0 == 0xE2) {
        if (b1 == 0x80) return ((b2 >= 0x80 && b2 <= 0x8A) || b2 == 0xA8 || b2 == 0xA9 || b2 == 0xAF) ? 3u : 0u;
        if (b1 == 0x81) return b2 == 0x9F ? 3u : 0u;
        return 0u;
    }
    if (b0 == 0xE3) return (b1 == 0x80 && b2 == 0x80) ? 3u : 0u;
    return 0u;
}

}  // namespace tdg
