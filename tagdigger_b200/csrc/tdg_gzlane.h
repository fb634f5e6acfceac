// One inflater per LANE: the device side of the gzip feed (csrc/tdg_gzdev.cuh) and its CPU test
// harness (tests/native/gzlane_check.cpp) share this file.
//
// find_tags_fastq opens 'gz' files with gzip.open(f, 'rt') (/root/reference/tagdigger_fun.py:240-241):
// one deflate stream, no index.  tdg_pgz.h inflates such a stream on host threads by entering it
// speculatively at block starts and writing 16-bit symbols (a reference into the unknown 32 KiB
// before the entry point becomes a marker); this file is the same idea shaped for a GPU, where
// there are thousands of slow lanes instead of sixteen fast cores:
//   * the compressed stream is cut into chunks, ONE LANE per chunk;
//   * a lane's Huffman tables live in its own slice of shared memory, interleaved word by word
//     with the slices of the other 31 lanes of its warp (STRIDE = 32), so that a table look-up of
//     a whole warp never has a bank conflict whatever the lanes index;
//   * entries are 16 bits: a 10-bit primary table for literal/length codes, an 8-bit one for
//     distance codes, no subtables -- longer codes (rare) are decoded canonically, bit by bit,
//     from the per-length counts and the symbols in code order;
//   * length / distance bases come from arithmetic, not from tables.
// Everything that decides whether a chunk's output is USED (the chain from chunk to chunk, member
// trailers, CRC) is host logic in tdg_gzdev.cuh; a lane only reports where it started, where it
// stopped and why.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define TDG_GZ_HD __host__ __device__ __forceinline__
#define TDG_GZ_FN __host__ __device__
#else
#define TDG_GZ_HD inline
#define TDG_GZ_FN
#endif

namespace tdg {
namespace gzl {

constexpr uint32_t WIN = 32768;
#ifndef TDG_GZ_LIT_ROOT
#define TDG_GZ_LIT_ROOT 9
#endif
#ifndef TDG_GZ_DIST_ROOT
#define TDG_GZ_DIST_ROOT 8
#endif
constexpr int LIT_ROOT = TDG_GZ_LIT_ROOT, DIST_ROOT = TDG_GZ_DIST_ROOT, PRE_ROOT = 7;
static_assert(LIT_ROOT >= 7 && LIT_ROOT <= 12 && DIST_ROOT >= 6 && DIST_ROOT <= 10, "table roots");

// A lane's HOT memory (shared memory on the device), in 16-bit units; every region starts on a
// 32-bit word.  What bounds the lanes per SM is this: 1,760 bytes per lane = 128 lanes = 4 warps.
constexpr uint32_t O_LIT = 0;                              // primary literal/length table (code-length table while a header is read)
constexpr uint32_t O_DIST = O_LIT + (1u << LIT_ROOT);      // primary distance table (scratch while the other codes are built)
constexpr uint32_t O_LCNT = O_DIST + (1u << DIST_ROOT);    // 16: codes per length; [0]: the first code of length root + 1
constexpr uint32_t O_DCNT = O_LCNT + 16;                   // 16
constexpr uint32_t O_LENS = O_DCNT + 16;                   // 320 x 4 bits: code lengths of the header being read
constexpr uint32_t LANE_U16 = O_LENS + 80;
static_assert(LANE_U16 % 2 == 0, "a lane's slice is a whole number of 32-bit words");
// A lane's COLD memory (global memory on the device): the symbols whose codes are longer than the
// roots, in code order -- touched only when such a code turns up (under 1 % of the symbols).
constexpr uint32_t C_LSORT = 0;                            // 288
constexpr uint32_t C_DSORT = 288;                          // 32
constexpr uint32_t COLD_U16 = 320;

// What a lane writes: TOKENS, 16 bits each -- a literal (0..255), or a match as two slots,
// T_LEN | length (3..258) followed by T_DIST | (distance - 1).  A lane never reads what it has
// produced, so its steps cost instruction latency only; copying the matches (whose sources are
// mostly out of L2: thousands of lanes times 32 KiB of history each) is left to gz_expand, where a
// warp per chunk has all the memory-level parallelism a lane lacks.
constexpr uint16_t T_LEN = 0x4000, T_DIST = 0x8000;

constexpr uint16_t E_LONG = 0x8000;           // primary entry: the code is longer than the root
// other entries: symbol << 4 | code length; 0 = no code here

enum : uint32_t {
    F_FOUND = 1,       // decoding started (a block start was accepted)
    F_FINAL = 2,       // stopped behind the last block of a member (end_bit = the bit after its end-of-block code)
    F_ERROR = 4,       // invalid data after end_bit
    F_SPACE = 8,       // the symbol buffer was full
    F_INPUT = 16,      // ran out of compressed bytes
};

struct Meta {                 // what a lane reports about its chunk
    uint64_t start_bit;       // where decoding started (absolute bit position in the file)
    uint64_t end_bit;         // the block boundary at which out_len symbols were complete
    uint32_t out_len;
    uint32_t flags;
    uint32_t min_pre;         // smallest index into the 32 KiB before the chunk that a match named directly (WIN: none)
    uint32_t ntok;            // token slots that hold those out_len symbols (expand_tokens)
};

template <int STRIDE>
struct Mem {
    uint16_t *base;           // the lane's first 16-bit unit of hot memory
    uint16_t *cold;           // the lane's cold memory (COLD_U16 units, not interleaved)
    TDG_GZ_HD uint16_t &h(uint32_t i) const { return base[((i >> 1) * STRIDE) * 2 + (i & 1u)]; }
    TDG_GZ_HD uint8_t &b(uint32_t off16, uint32_t i) const
    {
        const uint32_t byte = off16 * 2 + i;
        return ((uint8_t *)base)[((byte >> 2) * STRIDE) * 4 + (byte & 3u)];
    }
    // arrays of 4-bit code lengths
    TDG_GZ_HD uint32_t nib(uint32_t off16, uint32_t i) const { return (b(off16, i >> 1) >> ((i & 1u) * 4)) & 15u; }
    TDG_GZ_HD void set_nib(uint32_t off16, uint32_t i, uint32_t v) const
    {
        uint8_t &x = b(off16, i >> 1);
        x = (uint8_t)((i & 1u) ? ((x & 0x0Fu) | (v << 4)) : ((x & 0xF0u) | v));
    }
};

// LSB-first bit reader over 32-bit words (the buffer is padded: words past nwords read as zero).
// The word after the ones in the buffer is always on its way already (`ahead`): a lane's refill
// never waits for memory unless it drains 32 bits faster than a load takes.
struct Bits {
    const uint32_t *in;
    uint64_t nwords;
    uint64_t w;               // index of the word in `ahead`
    uint64_t buf;
    uint32_t ahead;
    int cnt;
    TDG_GZ_HD uint32_t word(uint64_t i) const { return i < nwords ? in[i] : 0u; }
    TDG_GZ_HD void fill()     // afterwards at least 32 bits are in the buffer
    {
        if (cnt <= 32) {
            buf |= (uint64_t)ahead << cnt;
            cnt += 32;
            w++;
            ahead = word(w);
        }
    }
    TDG_GZ_HD void seek(uint64_t bit)
    {
        w = bit >> 5;
        ahead = word(w);
        buf = 0;
        cnt = 0;
        fill();
        drop((int)(bit & 31));
    }
    TDG_GZ_HD void drop(int n)
    {
        buf >>= n;
        cnt -= n;
    }
    TDG_GZ_HD uint32_t take(int n)
    {
        const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        drop(n);
        return v;
    }
    TDG_GZ_HD uint64_t pos() const { return w * 32 - (uint64_t)cnt; }
};

// the low `len` bits of c in reverse order
TDG_GZ_HD uint32_t bitrev(uint32_t c, int len)
{
#if defined(__CUDA_ARCH__)
    return __brev(c) >> (32 - len);
#else
    uint32_t r = 0;
    for (int b = 0; b < len; b++) r |= ((c >> b) & 1u) << (len - 1 - b);
    return r;
#endif
}

// Builds the decoding structures of one canonical code from the n code lengths that start at
// nibble `first` of the lane's array of 4-bit lengths at 16-bit offset o_lens; o_sort: where in
// the lane's cold memory the symbols with codes longer than the root are listed.  Acceptance as in zlib's
// inflate_table: over-subscribed and incomplete sets are refused, except the incomplete set that
// is a single 1-bit code (allow_single); a set without any code is accepted and decodes nothing.
// o_scr: 32 16-bit units of lane memory that are free while the code is built (the next code
// value and the next slot in the long-code list, per length): one pass over the symbols.
template <int STRIDE>
TDG_GZ_FN bool build_code(const Mem<STRIDE> m, uint32_t o_lens, uint32_t first, uint32_t n, int root, uint32_t o_tab,
                          uint32_t o_sort, uint32_t o_cnt, uint32_t o_scr, bool allow_single, bool check_only = false)
{
    for (uint32_t l = 0; l < 16; l++) m.h(o_cnt + l) = 0;
    for (uint32_t i = 0; i < n; i++) m.h(o_cnt + m.nib(o_lens, first + i))++;
    m.h(o_cnt) = 0;
    int max = 15;
    while (max >= 1 && !m.h(o_cnt + max)) max--;
    const uint32_t psize = 1u << root;
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        left <<= 1;
        left -= (int)m.h(o_cnt + len);
        if (left < 0) return false;
    }
    if (max > 0 && left > 0 && !(allow_single && max == 1)) return false;
    if (check_only) return true;                             // (the scan only asks whether the header is one)
    if (left > 0)                                            // an incomplete (or empty) set leaves holes: "no code here"
        for (uint32_t j = 0; j < psize; j++) m.h(o_tab + j) = 0;
    if (max == 0) return true;
    {
        uint32_t code = 0, slot = 0, first_long = 0;
        for (int len = 1; len <= 15; len++) {
            code = (code + m.h(o_cnt + len - 1)) << 1;      // first code of this length
            m.h(o_scr + len) = (uint16_t)code;
            m.h(o_scr + 16 + len) = (uint16_t)slot;
            if (len > root) slot += m.h(o_cnt + len);
            if (len == root + 1) first_long = code;
        }
        m.h(o_cnt) = (uint16_t)first_long;                  // (the count of unused symbols is of no use: decode_long starts here)
    }
    for (uint32_t s = 0; s < n; s++) {
        const int len = (int)m.nib(o_lens, first + s);
        if (!len) continue;
        const uint32_t c = m.h(o_scr + len);
        m.h(o_scr + len) = (uint16_t)(c + 1);
        if (len <= root) {
            // every table index that ends in the (bit-reversed) code
            const uint16_t e = (uint16_t)(s << 4 | (uint32_t)len);
            for (uint32_t j = bitrev(c, len); j < psize; j += 1u << len) m.h(o_tab + j) = e;
        } else {
            // a longer code: mark its root-bit prefix, list the symbol in code order
            m.h(o_tab + bitrev(c >> (len - root), root)) = E_LONG;
            const uint32_t at = m.h(o_scr + 16 + len);
            m.h(o_scr + 16 + len) = (uint16_t)(at + 1);
            m.cold[o_sort + at] = (uint16_t)s;
        }
    }
    return true;
}

// A code longer than the root: its first `root` bits are the table index that led here; the
// rest is read bit by bit (canonical order: the first code of each length follows from the
// counts).  Returns symbol << 4 | length, 0 when the bits are no code.
template <int STRIDE>
TDG_GZ_FN uint32_t decode_long(const Mem<STRIDE> m, uint64_t buf, int root, uint32_t o_sort, uint32_t o_cnt)
{
    uint32_t code = bitrev((uint32_t)buf & ((1u << root) - 1u), root) << 1;
    uint32_t first = m.h(o_cnt);                             // first code of length root + 1
    buf >>= root;
    uint32_t index = 0;
    for (int len = root + 1; len <= 15; len++) {
        code |= (uint32_t)(buf & 1u);
        buf >>= 1;
        const uint32_t count = m.h(o_cnt + len);
        if (code < first + count) return (uint32_t)m.cold[o_sort + index + (code - first)] << 4 | (uint32_t)len;
        index += count;
        first = (first + count) << 1;
        code <<= 1;
    }
    return 0;
}

// A lane as a STATE MACHINE: step() does one small piece of work -- a symbol or two of a Huffman
// block, 64 bytes of a stored block, a block header, the next candidate -- and returns.  The
// kernel steps all 32 lanes of a warp in lock step behind a vote, so the lanes stay CONVERGED on
// the hot path (symbols) however differently their blocks are cut; a lane that runs its whole
// chunk in one call tree diverges from its neighbours at the first block end and the warp
// executes one lane at a time from then on (measured: 30 x slower).
enum : uint32_t { S_CAND = 0, S_HEADER = 1, S_HUFF = 2, S_STORED = 3, S_DONE = 4 };

template <int STRIDE>
struct Lane {
    Mem<STRIDE> m;
    Bits bits;
    uint64_t in_bits;         // valid bits of the buffer
    uint64_t wmax;            // a word index beyond this means the input is exhausted for sure
    uint16_t *out;            // token slots
    uint32_t cap;             // symbols the chunk may produce
    uint32_t tokcap;          // token slots `out` holds (literal-heavy data needs about two per compressed byte)
    uint32_t o;               // symbols produced
    uint32_t t;               // token slots written
    uint32_t hist;            // how far back a distance may reach beyond o (WIN: unknown prehistory)
    uint32_t min_pre;
    bool fixed_loaded;
    // the job
    bool known;
    uint32_t known_hist;
    uint64_t search_base, stop_bit;
    const uint32_t *cand;
    uint32_t ncand, c;
    // progress
    uint32_t state;
    bool probation;           // a candidate start is on trial until its first block is done and the header behind it is sane
    bool final_block;
    uint32_t stored_left, blocks_done;
    uint64_t start, last_bit;
    uint32_t last_o, last_t;
    Meta *r;

    TDG_GZ_HD bool exhausted() const { return bits.pos() > in_bits; }

    TDG_GZ_FN void init(Mem<STRIDE> mem, const uint32_t *in, uint64_t nwords, uint64_t inbits, bool known_start, uint64_t base,
                        const uint32_t *cands, uint32_t ncands, uint64_t stop, uint32_t history, uint16_t *dst, uint32_t dst_cap,
                        uint32_t sym_cap, Meta *report)
    {
        m = mem;
        bits.in = in;
        bits.nwords = nwords;
        bits.w = 0;
        bits.buf = 0;
        bits.ahead = 0;
        bits.cnt = 0;
        in_bits = inbits;
        wmax = (inbits + 31) / 32 + 2;
        out = dst;
        tokcap = dst_cap;
        cap = sym_cap;
        o = 0;
        t = 0;
        hist = WIN;
        min_pre = WIN;
        fixed_loaded = false;
        known = known_start;
        known_hist = history;
        search_base = base;
        stop_bit = stop;
        cand = cands;
        ncand = ncands;
        c = 0;
        state = S_CAND;
        probation = false;
        final_block = false;
        stored_left = blocks_done = 0;
        start = last_bit = base;
        last_o = last_t = 0;
        r = report;
        r->start_bit = base;
        r->end_bit = base;
        r->out_len = 0;
        r->flags = 0;
        r->min_pre = WIN;
        r->ntok = 0;
    }

    TDG_GZ_FN void finish(uint32_t flags)
    {
        r->start_bit = start;
        r->end_bit = last_bit;
        r->out_len = last_o;
        r->ntok = last_t;
        r->flags = F_FOUND | flags;
        r->min_pre = min_pre;
        state = S_DONE;
    }
    // decoding cannot go on: a candidate on trial is dropped, anything else is reported with the last boundary passed
    TDG_GZ_FN void fail(uint32_t why)
    {
        if (probation) state = S_CAND;
        else finish(why);
    }

    TDG_GZ_FN bool load_fixed()
    {
        for (uint32_t i = 0; i < 144; i++) m.set_nib(O_LENS, i, 8);
        for (uint32_t i = 144; i < 256; i++) m.set_nib(O_LENS, i, 9);
        for (uint32_t i = 256; i < 280; i++) m.set_nib(O_LENS, i, 7);
        for (uint32_t i = 280; i < 288; i++) m.set_nib(O_LENS, i, 8);
        if (!build_code<STRIDE>(m, O_LENS, 0, 288, LIT_ROOT, O_LIT, C_LSORT, O_LCNT, O_DIST, true)) return false;
        for (uint32_t i = 0; i < 32; i++) m.set_nib(O_LENS, i, 5);
        return build_code<STRIDE>(m, O_LENS, 0, 32, DIST_ROOT, O_DIST, C_DSORT, O_DCNT, O_LENS + 32, true);
    }

    // dynamic block header at the reader's position (behind the three block bits)
    TDG_GZ_FN bool read_dynamic(bool check_only = false)
    {
        bits.fill();
        const uint32_t hlit = bits.take(5) + 257, hdist = bits.take(5) + 1, hclen = bits.take(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        // order of the code-length code's lengths: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, five bits each
        const uint64_t ord0 = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 |
                              10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
        const uint64_t ord1 = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
        // the code-length code: its lengths sit in the (still unused) distance table, its decoding
        // table borrows the literal/length region -- both are rebuilt below
        for (uint32_t i = 0; i < 10; i++) m.b(O_DIST, i) = 0;
        for (uint32_t i = 0; i < hclen; i++) {
            bits.fill();
            const uint32_t sym = (uint32_t)((i < 12 ? ord0 >> (5 * i) : ord1 >> (5 * (i - 12))) & 31u);
            m.set_nib(O_DIST, sym, bits.take(3));
        }
        if (!build_code<STRIDE>(m, O_DIST, 0, 19, PRE_ROOT, O_LIT, C_LSORT, O_LCNT, O_DIST + 32, false)) return false;
        const uint32_t total = hlit + hdist;
        uint32_t i = 0;
        uint32_t prev = 0;
        // Kraft sums of the two codes as their lengths arrive (in units of 2^-15): random bits
        // over-subscribe a code within a few dozen lengths, long before the header has been read
        uint32_t kraft_l = 0, kraft_d = 0;
        while (i < total) {
            bits.fill();
            const uint32_t e = m.h(O_LIT + ((uint32_t)bits.buf & ((1u << PRE_ROOT) - 1u)));
            if (e == 0 || (e & E_LONG)) return false;       // (codes of the code-length code have at most 7 bits)
            bits.drop((int)(e & 15u));
            const uint32_t sym = e >> 4;
            if (sym < 16) {
                m.set_nib(O_LENS, i, sym);
                if (sym) {
                    if (i < hlit) kraft_l += 32768u >> sym;
                    else kraft_d += 32768u >> sym;
                    if (kraft_l > 32768u || kraft_d > 32768u) return false;
                }
                prev = sym;
                i++;
                continue;
            }
            uint32_t rep, val = 0;
            if (sym == 16) {
                if (i == 0) return false;
                val = prev;
                rep = 3 + bits.take(2);
            } else if (sym == 17) {
                rep = 3 + bits.take(3);
            } else {
                rep = 11 + bits.take(7);
            }
            if (i + rep > total) return false;
            while (rep--) {
                if (val) {
                    if (i < hlit) kraft_l += 32768u >> val;
                    else kraft_d += 32768u >> val;
                }
                m.set_nib(O_LENS, i++, val);
            }
            if (kraft_l > 32768u || kraft_d > 32768u) return false;
            prev = val;
        }
        if (exhausted()) return false;
        if (m.nib(O_LENS, 256) == 0) return false;          // no end-of-block code
        if (!build_code<STRIDE>(m, O_LENS, 0, hlit, LIT_ROOT, O_LIT, C_LSORT, O_LCNT, O_DIST, true, check_only)) return false;
        if (!build_code<STRIDE>(m, O_LENS, hlit, hdist, DIST_ROOT, O_DIST, C_DSORT, O_DCNT, O_LENS, true, check_only)) return false;
        fixed_loaded = false;
        return true;
    }

    TDG_GZ_FN void block_end()
    {
        if (exhausted()) {
            fail(F_INPUT);
            return;
        }
        blocks_done++;
        if (final_block) {
            last_bit = bits.pos();
            last_o = o;
            last_t = t;
            finish(F_FINAL);
            return;
        }
        state = S_HEADER;
    }

    // S_CAND: the next place to start from
    TDG_GZ_FN void step_cand()
    {
        if (known) {
            if (c++) {                                       // (a known start is never on trial: this is not reached)
                state = S_DONE;
                return;
            }
            start = search_base;
            hist = known_hist;
            probation = false;
        } else {
            if (c >= ncand) {                                // nothing held: F_FOUND stays clear
                state = S_DONE;
                return;
            }
            start = search_base + cand[c++];
            hist = WIN;
            probation = true;
        }
        bits.seek(start);
        o = 0;
        t = 0;
        min_pre = WIN;
        fixed_loaded = false;
        blocks_done = 0;
        state = S_HEADER;
    }

    // S_HEADER: at a block boundary.  A lane stops at the first boundary at or behind stop_bit that
    // the NEXT lane's scan can find: one in front of a non-final dynamic block.
    TDG_GZ_FN void step_header()
    {
        const uint64_t bp = bits.pos();
        last_bit = bp;
        last_o = o;
        last_t = t;
        if (bp + 3 > in_bits) {
            fail(F_INPUT);
            return;
        }
        bits.fill();
        const uint32_t h = bits.take(3);
        const uint32_t type = h >> 1;
        final_block = (h & 1u) != 0;
        if (type == 3) {
            fail(F_ERROR);
            return;
        }
        if (type == 2) {
            const bool at_stop = bp >= stop_bit && !final_block;
            if (at_stop && !probation) {
                finish(0);
                return;
            }
            if (!read_dynamic()) {
                fail(exhausted() ? F_INPUT : F_ERROR);
                return;
            }
            if (blocks_done) probation = false;              // the header behind the first block is sane
            if (at_stop) {
                finish(0);
                return;
            }
            state = S_HUFF;
            return;
        }
        if (type == 1) {
            if (!fixed_loaded) {
                if (!load_fixed()) {
                    fail(F_ERROR);
                    return;
                }
                fixed_loaded = true;
            }
            if (blocks_done) probation = false;
            state = S_HUFF;
            return;
        }
        bits.drop(bits.cnt & 7);
        bits.fill();
        const uint32_t len = bits.take(16);
        bits.fill();
        const uint32_t nlen = bits.take(16);
        if ((len ^ 0xffffu) != nlen) {
            fail(F_ERROR);
            return;
        }
        if (bits.pos() + (uint64_t)len * 8 > in_bits) {
            fail(F_INPUT);
            return;
        }
        if (o + len + 260 > cap || t + len + 8 > tokcap) {
            fail(F_SPACE);
            return;
        }
        if (blocks_done) probation = false;
        stored_left = len;
        state = S_STORED;
    }

    TDG_GZ_FN void step_stored()
    {
        uint32_t n = stored_left < 64 ? stored_left : 64;
        stored_left -= n;
        while (n--) {
            bits.fill();
            out[t++] = (uint16_t)bits.take(8);
            o++;
        }
        if (stored_left == 0) block_end();
    }

    // S_HUFF: one literal/length symbol (and a second literal when it is there for the taking), or a whole match
    TDG_GZ_FN void step_huff()
    {
        if (o + 260 > cap || t + 8 > tokcap) {
            fail(F_SPACE);
            return;
        }
        if (bits.w > wmax) {
            fail(F_INPUT);
            return;
        }
        bits.fill();
        uint32_t e = m.h(O_LIT + ((uint32_t)bits.buf & ((1u << LIT_ROOT) - 1u)));
        if (e & E_LONG) e = decode_long<STRIDE>(m, bits.buf, LIT_ROOT, C_LSORT, O_LCNT);
        if (e == 0) {
            fail(F_ERROR);
            return;
        }
        bits.drop((int)(e & 15u));
        const uint32_t sym = e >> 4;
        if (sym < 256) {
            out[t++] = (uint16_t)sym;
            o++;
            // more literals without another fill, while the bits left cover a table index (a code
            // that is not longer than the root is decided by those bits alone)
#pragma unroll
            for (int more = 0; more < 2; more++) {
                if (bits.cnt < LIT_ROOT) return;
                e = m.h(O_LIT + ((uint32_t)bits.buf & ((1u << LIT_ROOT) - 1u)));
                if (e == 0 || e >= (256u << 4)) return;      // not a short literal code: the next step looks again
                bits.drop((int)(e & 15u));
                out[t++] = (uint16_t)(e >> 4);
                o++;
            }
            return;
        }
        if (sym == 256) {
            block_end();
            return;
        }
        if (sym > 285) {
            fail(F_ERROR);
            return;
        }
        uint32_t len;
        {
            const uint32_t idx = sym - 257;
            if (idx < 8) len = 3 + idx;
            else if (idx == 28) len = 258;
            else {
                const int eb = (int)((idx - 4) >> 2);
                len = 3 + ((4 + (idx & 3u)) << eb) + bits.take(eb);
            }
        }
        bits.fill();
        uint32_t f = m.h(O_DIST + ((uint32_t)bits.buf & ((1u << DIST_ROOT) - 1u)));
        if (f & E_LONG) f = decode_long<STRIDE>(m, bits.buf, DIST_ROOT, C_DSORT, O_DCNT);
        if (f == 0) {
            fail(F_ERROR);
            return;
        }
        bits.drop((int)(f & 15u));
        const uint32_t ds = f >> 4;
        if (ds > 29) {
            fail(F_ERROR);
            return;
        }
        uint32_t d;
        if (ds < 4) d = 1 + ds;
        else {
            const int eb = (int)(ds >> 1) - 1;
            d = 1 + ((2 + (ds & 1u)) << eb) + bits.take(eb);
        }
        if (d > o) {
            if (d - o > hist) {                              // too far back
                fail(F_ERROR);
                return;
            }
            const uint32_t pre = WIN - (d - o);
            if (pre < min_pre) min_pre = pre;
        }
        out[t] = (uint16_t)(T_LEN | len);
        out[t + 1] = (uint16_t)(T_DIST | (d - 1));
        t += 2;
        o += len;
    }

    TDG_GZ_FN void step()
    {
        if (state == S_HUFF) step_huff();
        else if (state == S_STORED) step_stored();
        else if (state == S_HEADER) step_header();
        else if (state == S_CAND) step_cand();
    }
};

// Cheap test of bit position p as the start of a non-final dynamic block: block bits, HLIT / HDIST
// in range, and the code-length code exactly complete.  lo = the 64 bits at p, hi = the 32 after
// them.  kraft3[v] = sum over the three 3-bit lengths in v of (128 >> length), zero lengths adding
// nothing (512 entries, shared).
TDG_GZ_HD bool quick_test(uint64_t lo, uint32_t hi, const uint8_t *kraft3)
{
    if (((uint32_t)lo & 7u) != 4u) return false;
    if ((((uint32_t)lo >> 3) & 31u) > 29u || (((uint32_t)lo >> 8) & 31u) > 29u) return false;
    const uint32_t n = (((uint32_t)lo >> 13) & 15u) + 4u;           // lengths present: 4..19
    uint64_t v = lo >> 17 | (uint64_t)hi << 47;                      // 57 bits of lengths (bits 17..73 of the header)
    if (n < 19) v &= (1ull << (3 * n)) - 1ull;
    else v &= (1ull << 57) - 1ull;
    uint32_t sum = 0;
    for (int j = 0; j < 7; j++) sum += kraft3[(uint32_t)(v >> (9 * j)) & 511u];
    return sum == 128u;
}

inline void make_kraft3(uint8_t *t)
{
    for (uint32_t v = 0; v < 512; v++) {
        uint32_t s = 0;
        for (int k = 0; k < 3; k++) {
            const uint32_t l = (v >> (3 * k)) & 7u;
            if (l) s += 128u >> l;
        }
        t[v] = (uint8_t)s;
    }
}

// Tokens -> symbols, one after the other (the CPU harness; on the device gz_expand does this with
// a warp per chunk).  sym[j] for j < 0 is the unknown text before the chunk: marker 256 + (WIN + j).
TDG_GZ_FN inline uint32_t expand_tokens(const uint16_t *tok, uint32_t ntok, uint16_t *sym)
{
    uint32_t o = 0;
    for (uint32_t i = 0; i < ntok; i++) {
        const uint16_t s = tok[i];
        if (s < 256) {
            sym[o++] = s;
            continue;
        }
        const uint32_t len = s & 0x1FFu, d = (uint32_t)(tok[++i] & 0x7FFFu) + 1u;
        int32_t j = (int32_t)o - (int32_t)d;
        for (uint32_t k = 0; k < len; k++, j++) sym[o + k] = j < 0 ? (uint16_t)(256 + WIN + j) : sym[j];
        o += len;
    }
    return o;
}

// Does a dynamic block header parse at `bit` (both Huffman codes valid, an end-of-block code
// present)?  The scan's second test for a position that passed quick_test.
template <int STRIDE>
TDG_GZ_FN bool header_parses(Mem<STRIDE> m, const uint32_t *in, uint64_t nwords, uint64_t in_bits, uint64_t bit)
{
    Lane<STRIDE> z;
    Meta dummy;
    z.init(m, in, nwords, in_bits, true, bit, nullptr, 0, 0, 0, nullptr, 0, 0, &dummy);
    z.bits.seek(bit);
    z.bits.fill();
    z.bits.drop(3);
    return z.read_dynamic(true);
}

// The scan's second test, lean enough for 32 candidates at a time (one per lane, 384 bytes of
// interleaved shared memory each): does a dynamic block header parse at `bit`?  Same verdict as
// Lane::read_dynamic, but nothing is kept -- the code-length code gets its table, the lengths it
// decodes are only summed up (Kraft sums of the two codes as they arrive: random bits
// over-subscribe a code within a few dozen lengths), and the end-of-block code must exist.
constexpr uint32_t V_TAB = 0;                 // 128: table of the code-length code
constexpr uint32_t V_CNT = 128;               // 16
constexpr uint32_t V_LENS = 144;              // 19 x 4 bits (10 units), padded to 16
constexpr uint32_t V_SCR = 160;               // 32
constexpr uint32_t VAL_U16 = 192;

template <int STRIDE>
TDG_GZ_FN bool header_check(const Mem<STRIDE> m, const uint32_t *in, uint64_t nwords, uint64_t in_bits, uint64_t bit)
{
    Bits bits;
    bits.in = in;
    bits.nwords = nwords;
    bits.seek(bit);
    bits.fill();
    bits.drop(3);
    bits.fill();
    const uint32_t hlit = bits.take(5) + 257, hdist = bits.take(5) + 1, hclen = bits.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    const uint64_t ord0 = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 |
                          10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
    const uint64_t ord1 = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    for (uint32_t i = 0; i < 10; i++) m.b(V_LENS, i) = 0;
    for (uint32_t i = 0; i < hclen; i++) {
        bits.fill();
        const uint32_t sym = (uint32_t)((i < 12 ? ord0 >> (5 * i) : ord1 >> (5 * (i - 12))) & 31u);
        m.set_nib(V_LENS, sym, bits.take(3));
    }
    if (!build_code<STRIDE>(m, V_LENS, 0, 19, PRE_ROOT, V_TAB, 0, V_CNT, V_SCR, false)) return false;
    const uint32_t total = hlit + hdist;
    uint32_t i = 0, prev = 0, eob = 0;
    uint32_t kraft_l = 0, kraft_d = 0, codes_l = 0, codes_d = 0, max_l = 0, max_d = 0;
    while (i < total) {
        bits.fill();
        const uint32_t e = m.h(V_TAB + ((uint32_t)bits.buf & ((1u << PRE_ROOT) - 1u)));
        if (e == 0 || (e & E_LONG)) return false;
        bits.drop((int)(e & 15u));
        const uint32_t sym = e >> 4;
        uint32_t rep = 1, val = sym;
        if (sym >= 16) {
            val = 0;
            if (sym == 16) {
                if (i == 0) return false;
                val = prev;
                rep = 3 + bits.take(2);
            } else if (sym == 17) {
                rep = 3 + bits.take(3);
            } else {
                rep = 11 + bits.take(7);
            }
            if (i + rep > total) return false;
        }
        prev = val;
        if (val) {
            // how many of the rep lengths fall to the literal/length code
            const uint32_t to_l = i >= hlit ? 0u : (i + rep <= hlit ? rep : hlit - i);
            kraft_l += to_l * (32768u >> val);
            kraft_d += (rep - to_l) * (32768u >> val);
            if (kraft_l > 32768u || kraft_d > 32768u) return false;
            codes_l += to_l;
            codes_d += rep - to_l;
            if (to_l && val > max_l) max_l = val;
            if (rep > to_l && val > max_d) max_d = val;
            if (i <= 256 && 256 < i + rep) eob = val;
        }
        i += rep;
    }
    if (bits.pos() > in_bits) return false;
    if (eob == 0) return false;
    // an incomplete code is accepted only when it is a single 1-bit code (zlib's inflate_table)
    if (kraft_l < 32768u && max_l != 1) return false;
    if (codes_d && kraft_d < 32768u && max_d != 1) return false;
    return true;
}

// A lane's whole job for its chunk, one call (the CPU harness; the kernel steps its lanes itself):
// try the candidate block starts `cand[0..ncand)` (bit offsets from search_base, ascending;
// positions where the scan saw a dynamic block header parse) until one holds -- its first block
// decodes and the header behind that block is sane -- then inflate up to the block boundary at or
// beyond stop_bit that stands in front of a non-final dynamic block.  `known`: start exactly at
// search_base with `hist` bytes of real history (the first chunk of a round); nothing is tried,
// every defect is real.  Bit positions are relative to the buffer (`in`).
template <int STRIDE>
TDG_GZ_FN void run_chunk(Mem<STRIDE> m, const uint32_t *in, uint64_t nwords, uint64_t in_bits, bool known, uint64_t search_base,
                         const uint32_t *cand, uint32_t ncand, uint64_t stop_bit, uint32_t hist, uint16_t *out, uint32_t tokcap,
                         uint32_t symcap, Meta &r)
{
    Lane<STRIDE> z;
    z.init(m, in, nwords, in_bits, known, search_base, cand, ncand, stop_bit, hist, out, tokcap, symcap, &r);
    while (z.state != S_DONE) z.step();
}

}  // namespace gzl
}  // namespace tdg
