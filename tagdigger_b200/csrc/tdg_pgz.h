// Parallel inflate of ONE ordinary gzip stream on host threads.  Plain C++, no CUDA.
//
// find_tags_fastq reads 'gz' files with gzip.open(f, 'rt') (/root/reference/tagdigger_fun.py:240-241):
// one thread, one deflate stream, no index.  A deflate stream cannot be entered in the middle by a
// stock inflater because (a) block starts are not byte aligned or marked and (b) every block may
// refer to the 32 KiB of text before it.  This reader does both speculatively:
//   1. the compressed file is cut into chunks of a nominal size; one thread per chunk looks for
//      the first bit position at or after its nominal offset where a non-final dynamic-Huffman
//      block header parses (complete code-length code, complete literal/length code with an
//      end-of-block symbol) and whose first block decodes to its end;
//   2. from there it inflates into 16-BIT symbols: a reference into the unknown 32 KiB before the
//      chunk yields the marker 256 + (index into that window) instead of a byte, and markers are
//      copied around like bytes;
//   3. a short serial pass chains the chunks: chunk k is accepted only if the inflater of chunk
//      k-1 arrived, at a block boundary, on exactly the bit where chunk k started (so a start
//      that was not a real block start can never contribute output); gaps (stored / fixed /
//      final blocks the finder skips, chunks whose buffer filled up) are inflated on the spot
//      with the now known window; the window is handed from chunk to chunk (32 Ki look-ups each);
//   4. markers are replaced and the CRC-32 of every piece is taken in parallel while the bytes
//      are written to their final place; the pieces' CRCs are folded with crc32_combine and
//      checked against every member trailer (CRC32, ISIZE) like gzip does.
// Anything this reader does not want to judge itself (corrupt data, header flags it does not
// parse, garbage after a member, a truncated file) makes the caller fall back to zlib at the
// uncompressed offset delivered so far, so that error behaviour stays that of the zlib path.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace tdg {
namespace pgz {

constexpr uint32_t WIN = 32768;
constexpr int LIT_PRIMARY = 10, DIST_PRIMARY = 8, PRE_PRIMARY = 7;
constexpr uint32_t K_INVALID = 0, K_LIT = 1, K_LEN = 2, K_EOB = 3, K_SUB = 4, K_DIST = 5;

// table entry: bits 0-3 code bits to drop, 4-7 extra bits (K_SUB: index bits of the subtable),
// 8-10 kind, 16-31 value (literal, length base, distance base, subtable offset)
static inline uint32_t mk(uint32_t kind, uint32_t value, uint32_t extra) { return value << 16 | kind << 8 | extra << 4; }
static inline uint32_t e_kind(uint32_t e) { return (e >> 8) & 7; }
static inline uint32_t e_extra(uint32_t e) { return (e >> 4) & 15; }

struct SymbolTemplates {
    uint32_t lit[288], dist[32], pre[19];
    SymbolTemplates()
    {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (uint32_t s = 0; s < 256; s++) lit[s] = mk(K_LIT, s, 0);
        lit[256] = mk(K_EOB, 0, 0);
        for (uint32_t s = 257; s < 286; s++) lit[s] = mk(K_LEN, lbase[s - 257], lext[s - 257]);
        lit[286] = lit[287] = mk(K_INVALID, 0, 0);
        for (uint32_t s = 0; s < 30; s++) dist[s] = mk(K_DIST, dbase[s], dext[s]);
        dist[30] = dist[31] = mk(K_INVALID, 0, 0);
        for (uint32_t s = 0; s < 19; s++) pre[s] = mk(K_LIT, s, 0);
    }
};

static inline const SymbolTemplates &templates()
{
    static const SymbolTemplates t;
    return t;
}

// Canonical Huffman decoding table, indexed by the next bits of the stream (LSB first).  Same
// acceptance rules as zlib's inflate_table: over-subscribed sets and incomplete sets are refused,
// except the incomplete set that consists of a single 1-bit code (allow_single).
static bool build_table(const uint8_t *lens, int n, int primary, const uint32_t *tmpl, uint32_t *table, bool allow_single)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int max = 15;
    while (max >= 1 && !count[max]) max--;
    const uint32_t psize = 1u << primary;
    memset(table, 0, psize * sizeof(uint32_t));
    if (max == 0) return true;                       // no codes at all: every look-up is invalid
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return false;
    }
    if (left > 0 && !(allow_single && max == 1)) return false;
    uint32_t next[16];
    uint32_t code = 0;
    for (int len = 1; len <= 15; len++) {
        code = (code + (uint32_t)count[len - 1]) << 1;
        next[len] = code;
    }
    const int subbits = max > primary ? max - primary : 0;
    uint32_t used = psize;
    for (int s = 0; s < n; s++) {
        const int len = lens[s];
        if (!len) continue;
        uint32_t c = next[len]++;
        uint32_t rev = 0;
        for (int b = 0; b < len; b++) rev |= ((c >> b) & 1u) << (len - 1 - b);
        if (len <= primary) {
            const uint32_t e = tmpl[s] | (uint32_t)len;
            for (uint32_t j = rev; j < psize; j += 1u << len) table[j] = e;
        } else {
            const uint32_t prefix = rev & (psize - 1);
            if (e_kind(table[prefix]) != K_SUB) {
                table[prefix] = mk(K_SUB, used, (uint32_t)subbits) | (uint32_t)primary;
                memset(table + used, 0, sizeof(uint32_t) << subbits);
                used += 1u << subbits;
            }
            uint32_t *sub = table + (table[prefix] >> 16);
            const uint32_t e = tmpl[s] | (uint32_t)(len - primary);
            for (uint32_t j = rev >> primary; j < (1u << subbits); j += 1u << (len - primary)) sub[j] = e;
        }
    }
    return true;
}

enum Status { ST_STOP, ST_SPACE, ST_MEMBER_END, ST_ERROR };

// Resumable inflater that writes 16-bit symbols behind a 32 Ki-symbol "prehistory".
struct Inflater {
    const uint8_t *in = nullptr;
    size_t in_size = 0;
    size_t pos = 0;            // next byte to load
    uint64_t bitbuf = 0;
    int bitcnt = 0;
    uint16_t *out = nullptr;   // symbols; [0, WIN) is the window before the chunk
    size_t o = WIN, cap = 0;
    size_t member_start = 0;   // smallest index a distance may reach
    enum { HEADER, STORED, HUFF } state = HEADER;
    bool final_block = false;
    uint32_t stored_left = 0;
    uint64_t last_bit = 0;     // the last block boundary passed: bit position, output index, member start
    size_t last_o = WIN, last_ms = 0;
    int fixed_loaded = 0;
    uint32_t lit[(1 << LIT_PRIMARY) + 288 * 32];
    uint32_t dist[(1 << DIST_PRIMARY) + 32 * 128];
    uint32_t pre[1 << PRE_PRIMARY];

    uint64_t bitpos() const { return (uint64_t)pos * 8 - (uint64_t)bitcnt; }
    bool past_end() const { return pos > in_size && bitpos() > (uint64_t)in_size * 8; }

    void seek(uint64_t bit)
    {
        pos = (size_t)(bit >> 3);
        bitbuf = 0;
        bitcnt = 0;
        state = HEADER;
        refill();
        drop((int)(bit & 7));
    }
    inline void refill()
    {
        if (pos + 8 <= in_size) {
            uint64_t w;
            memcpy(&w, in + pos, 8);
            bitbuf |= w << bitcnt;
            pos += (size_t)((63 - bitcnt) >> 3);
            bitcnt |= 56;
        } else {
            while (bitcnt <= 56) {
                if (pos < in_size) bitbuf |= (uint64_t)in[pos] << bitcnt;
                pos++;
                bitcnt += 8;
            }
        }
    }
    inline void drop(int n)
    {
        bitbuf >>= n;
        bitcnt -= n;
    }
    inline uint32_t take(int n)
    {
        uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1));
        drop(n);
        return v;
    }
    void align_byte() { drop(bitcnt & 7); }
    size_t bytepos() const { return (size_t)(bitpos() >> 3); }     // after align_byte()

    bool load_fixed()
    {
        uint8_t l[288], d[32];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        for (int i = 0; i < 32; i++) d[i] = 5;
        const SymbolTemplates &t = templates();
        return build_table(l, 288, LIT_PRIMARY, t.lit, lit, true) && build_table(d, 32, DIST_PRIMARY, t.dist, dist, true);
    }

    bool read_dynamic()
    {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const SymbolTemplates &t = templates();
        refill();
        const uint32_t hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pl[19] = {0};
        for (uint32_t i = 0; i < hclen; i++) {
            if (bitcnt < 3) refill();
            pl[order[i]] = (uint8_t)take(3);
        }
        if (!build_table(pl, 19, PRE_PRIMARY, t.pre, pre, false)) return false;
        uint8_t lens[288 + 32];
        const uint32_t total = hlit + hdist;
        uint32_t i = 0;
        while (i < total) {
            refill();
            const uint32_t e = pre[bitbuf & ((1u << PRE_PRIMARY) - 1)];
            if (e_kind(e) != K_LIT) return false;
            drop((int)(e & 15));
            const uint32_t sym = e >> 16;
            if (sym < 16) {
                lens[i++] = (uint8_t)sym;
                continue;
            }
            uint32_t rep;
            uint8_t val = 0;
            if (sym == 16) {
                if (i == 0) return false;
                val = lens[i - 1];
                rep = 3 + take(2);
            } else if (sym == 17) {
                rep = 3 + take(3);
            } else {
                rep = 11 + take(7);
            }
            if (i + rep > total) return false;
            while (rep--) lens[i++] = val;
        }
        if (past_end()) return false;
        if (lens[256] == 0) return false;
        if (!build_table(lens, (int)hlit, LIT_PRIMARY, t.lit, lit, true)) return false;
        if (!build_table(lens + hlit, (int)hdist, DIST_PRIMARY, t.dist, dist, true)) return false;
        fixed_loaded = 0;
        return true;
    }

    // the symbols of one Huffman block up to its end-of-block code; ST_STOP here means "block done"
    Status huff_body()
    {
        uint16_t *const dst = out;
        size_t at = o;
        const size_t ms = member_start;
        for (;;) {
            if (at + 272 > cap) {
                o = at;
                return ST_SPACE;
            }
            refill();
            if (pos > in_size && past_end()) {
                o = at;
                return ST_ERROR;
            }
            uint32_t e = lit[bitbuf & ((1u << LIT_PRIMARY) - 1)];
            if (e_kind(e) == K_SUB) {
                drop(LIT_PRIMARY);
                e = lit[(e >> 16) + (uint32_t)(bitbuf & ((1u << e_extra(e)) - 1))];
            }
            drop((int)(e & 15));
            uint32_t kind = e_kind(e);
            if (kind == K_LIT) {
                dst[at++] = (uint16_t)(e >> 16);
                // a second literal without another refill (>= 41 bits are left)
                e = lit[bitbuf & ((1u << LIT_PRIMARY) - 1)];
                if (e_kind(e) != K_LIT) continue;
                drop((int)(e & 15));
                dst[at++] = (uint16_t)(e >> 16);
                continue;
            }
            if (kind == K_LEN) {
                const uint32_t len = (e >> 16) + take((int)e_extra(e));
                uint32_t f = dist[bitbuf & ((1u << DIST_PRIMARY) - 1)];
                if (e_kind(f) == K_SUB) {
                    drop(DIST_PRIMARY);
                    f = dist[(f >> 16) + (uint32_t)(bitbuf & ((1u << e_extra(f)) - 1))];
                }
                drop((int)(f & 15));
                if (e_kind(f) != K_DIST) {
                    o = at;
                    return ST_ERROR;
                }
                const uint32_t d = (f >> 16) + take((int)e_extra(f));
                if (d > at - ms) {
                    o = at;
                    return ST_ERROR;
                }
                uint16_t *q = dst + at;
                const uint16_t *s = q - d;
                if (d >= 4) {
                    for (uint32_t k = 0; k < len; k += 4) memcpy(q + k, s + k, 8);     // may write 3 symbols of slack
                } else {
                    for (uint32_t k = 0; k < len; k++) q[k] = s[k];
                }
                at += len;
                continue;
            }
            o = at;
            if (kind == K_EOB) return past_end() ? ST_ERROR : ST_STOP;
            return ST_ERROR;
        }
    }

    // Inflates blocks until a block boundary at or beyond stop_bit (ST_STOP), the end of a
    // member's last block (ST_MEMBER_END; the caller reads trailer and next header), a full
    // output buffer (ST_SPACE, resumable) or invalid data (ST_ERROR).
    Status run(uint64_t stop_bit)
    {
        for (;;) {
            if (state == HEADER) {
                const uint64_t bp = bitpos();
                if (bp >= stop_bit) return ST_STOP;
                last_bit = bp;
                last_o = o;
                last_ms = member_start;
                refill();
                const uint32_t h = take(3);
                final_block = h & 1;
                const uint32_t type = h >> 1;
                if (type == 0) {
                    align_byte();
                    refill();
                    const uint32_t len = take(16), nlen = take(16);
                    if ((len ^ 0xffffu) != nlen || past_end()) return ST_ERROR;
                    // continue with whole bytes straight from the input
                    pos = bytepos();
                    bitbuf = 0;
                    bitcnt = 0;
                    stored_left = len;
                    state = STORED;
                } else if (type == 1) {
                    if (!fixed_loaded) {
                        if (!load_fixed()) return ST_ERROR;
                        fixed_loaded = 1;
                    }
                    state = HUFF;
                } else if (type == 2) {
                    if (!read_dynamic()) return ST_ERROR;
                    state = HUFF;
                } else {
                    return ST_ERROR;
                }
                if (past_end()) return ST_ERROR;
            }
            if (state == STORED) {
                while (stored_left) {
                    if (o + 272 > cap) return ST_SPACE;
                    const size_t n = std::min<size_t>(stored_left, cap - 272 - o + 1);
                    if (pos + n > in_size) return ST_ERROR;
                    for (size_t k = 0; k < n; k++) out[o + k] = in[pos + k];
                    o += n;
                    pos += n;
                    stored_left -= (uint32_t)n;
                }
            } else {
                const Status s = huff_body();
                if (s != ST_STOP) return s;
            }
            state = HEADER;
            if (final_block) return ST_MEMBER_END;
        }
    }
};

// Parses a gzip member header at byte p.  Returns the offset of the deflate data, 0 when the
// header is not one this reader handles (wrong magic, FHCRC, reserved flags, truncated).
static inline size_t gzip_header(const uint8_t *in, size_t n, size_t p)
{
    if (p + 10 > n || in[p] != 0x1f || in[p + 1] != 0x8b || in[p + 2] != 8) return 0;
    const uint32_t flg = in[p + 3];
    if (flg & 0xe2) return 0;
    size_t q = p + 10;
    if (flg & 4) {
        if (q + 2 > n) return 0;
        const size_t xlen = in[q] | (in[q + 1] << 8);
        q += 2 + xlen;
        if (q > n) return 0;
    }
    for (int f = 8; f <= 16; f <<= 1) {
        if (!(flg & f)) continue;
        while (q < n && in[q]) q++;
        if (q >= n) return 0;
        q++;
    }
    return q < n ? q : 0;
}

struct MemberEnd {
    size_t at;              // index in the chunk's symbols where the member ends
    uint32_t crc, isize;
};

// symbol buffer whose pages stay untouched until written (no zero fill)
struct SymBuf {
    uint16_t *p = nullptr;
    size_t n = 0;
    SymBuf() {}
    SymBuf(const SymBuf &) = delete;
    SymBuf &operator=(const SymBuf &) = delete;
    ~SymBuf() { free(p); }
    uint16_t *data() { return p; }
    size_t size() const { return n; }
    uint16_t &operator[](size_t i) { return p[i]; }
    void resize(size_t m)           // grows only, keeps the contents
    {
        if (m <= n) return;
        uint16_t *q = (uint16_t *)realloc(p, m * sizeof(uint16_t));
        if (!q) throw std::bad_alloc();
        p = q;
        n = m;
    }
};

struct Chunk {
    SymBuf buf;
    size_t len = WIN;            // valid symbols: [WIN, len)
    bool growable = false, found = false, eof = false, error = false;
    uint64_t start_bit = 0, end_bit = 0;
    size_t ms_end = 0;           // member_start at the end (how much history the next block may use)
    std::vector<MemberEnd> mends;
    std::vector<uint8_t> win_before;     // resolved window in front of the chunk (WIN bytes)

    void prepare(size_t cap_symbols, bool grow)
    {
        if (buf.size() < cap_symbols) buf.resize(cap_symbols);
        growable = grow;
        len = WIN;
        found = eof = error = false;
        mends.clear();
    }
    void unknown_window()
    {
        for (uint32_t i = 0; i < WIN; i++) buf[i] = (uint16_t)(256 + i);
    }
    void known_window(const uint8_t *w)
    {
        for (uint32_t i = 0; i < WIN; i++) buf[i] = w[i];
    }
};

// Drives z (positioned at a block header, writing into c) until a boundary at or after stop_bit.
static void drive(Inflater &z, Chunk &c, uint64_t stop_bit)
{
    z.out = c.buf.data();
    z.cap = c.buf.size();
    auto truncate = [&]() {
        c.len = z.last_o;
        c.end_bit = z.last_bit;
        c.ms_end = z.last_ms;
        while (!c.mends.empty() && c.mends.back().at > c.len) c.mends.pop_back();
    };
    for (;;) {
        const Status s = z.run(stop_bit);
        if (s == ST_STOP) {
            c.len = z.o;
            c.end_bit = z.bitpos();
            c.ms_end = z.member_start;
            return;
        }
        if (s == ST_SPACE) {
            if (c.growable) {
                c.buf.resize(c.buf.size() * 2);
                z.out = c.buf.data();
                z.cap = c.buf.size();
                continue;
            }
            truncate();
            return;
        }
        if (s == ST_ERROR) {
            truncate();
            c.error = true;
            return;
        }
        // end of a member: CRC32, ISIZE, then the end of the file or another member
        z.align_byte();
        z.refill();
        const uint32_t crc = z.take(32);
        z.refill();
        const uint32_t isize = z.take(32);
        if (z.past_end()) {
            truncate();
            c.error = true;
            return;
        }
        c.mends.push_back(MemberEnd{z.o, crc, isize});
        const size_t p = z.bytepos();
        // the end of the file, or trailing bytes that do not start with the gzip magic (zlib's
        // gzread ignores such garbage; so does this reader)
        if (p + 1 >= z.in_size || z.in[p] != 0x1f || z.in[p + 1] != 0x8b) {
            c.len = z.o;
            c.end_bit = (uint64_t)z.in_size * 8;
            c.ms_end = z.o;
            c.eof = true;
            return;
        }
        const size_t q = gzip_header(z.in, z.in_size, p);
        if (!q) {
            c.len = z.o;                 // everything up to the trailer is good; what follows is for zlib to judge
            c.end_bit = (uint64_t)p * 8;
            c.ms_end = z.o;
            c.error = true;
            return;
        }
        z.seek((uint64_t)q * 8);
        z.member_start = z.o;
    }
}

// Looks for a block start in [from_bit, to_bit) and, when one is found, inflates from it.
static void speculate(Inflater &z, Chunk &c, uint64_t from_bit, uint64_t to_bit, uint64_t stop_bit)
{
    c.unknown_window();
    z.out = c.buf.data();
    z.cap = c.buf.size();
    const uint8_t *in = z.in;
    const size_t n = z.in_size;
    for (uint64_t b = from_bit; b < to_bit; b++) {
        const size_t p = (size_t)(b >> 3);
        if (p + 8 > n) break;                            // the last bytes of a file hold no dynamic block worth finding
        uint64_t w;
        memcpy(&w, in + p, 8);
        w >>= (b & 7);
        // BFINAL = 0, BTYPE = 2 (bits 1-2 = 0,1), HLIT <= 29, HDIST <= 29
        if ((w & 7) != 4 || ((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29) continue;
        {   // Kraft sum of the code-length code must be exactly 1
            const uint32_t hclen = (uint32_t)((w >> 13) & 15) + 4;
            uint64_t v = w >> 17;                        // 40 valid bits: 13 lengths
            uint32_t sum = 0;
            uint32_t take_n = hclen < 13 ? hclen : 13;
            for (uint32_t i = 0; i < take_n; i++) {
                const uint32_t l = (uint32_t)(v & 7);
                v >>= 3;
                if (l) sum += 128u >> l;
            }
            if (sum > 128 || (hclen <= 13 && sum != 128)) continue;
        }
        z.seek(b);
        z.o = WIN;
        z.member_start = 0;
        z.last_bit = b;
        z.last_o = WIN;
        if (z.run(b + 1) != ST_STOP) continue;           // exactly one block
        // the next header must at least be sane
        {
            const size_t spos = z.pos;
            const uint64_t sbuf = z.bitbuf;
            const int scnt = z.bitcnt;
            z.refill();
            const uint32_t h = z.take(3);
            bool ok = true;
            if ((h >> 1) == 3) ok = false;
            else if ((h >> 1) == 0) {
                z.align_byte();
                z.refill();
                const uint32_t len = z.take(16), nlen = z.take(16);
                ok = (len ^ 0xffffu) == nlen;
            } else if ((h >> 1) == 2) ok = z.read_dynamic();
            if (z.past_end()) ok = false;
            z.pos = spos;
            z.bitbuf = sbuf;
            z.bitcnt = scnt;
            z.state = Inflater::HEADER;
            if (!ok) continue;
        }
        c.found = true;
        c.start_bit = b;
        drive(z, c, stop_bit);
        if (c.len == WIN && c.end_bit <= b) c.found = false;     // nothing usable
        return;
    }
}

// dst[i] = byte of symbol src[i] (markers looked up in win); returns the CRC-32 of the bytes
static inline uint32_t resolve(const uint16_t *src, size_t n, const uint8_t *win, uint8_t *dst)
{
    size_t i = 0;
#if defined(__SSE2__)
    const __m128i hi = _mm_set1_epi16((short)0xff00);
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i *)(src + i + 8));
        if (_mm_movemask_epi8(_mm_cmpeq_epi16(_mm_and_si128(_mm_or_si128(a, b), hi), _mm_setzero_si128())) == 0xffff) {
            _mm_storeu_si128((__m128i *)(dst + i), _mm_packus_epi16(a, b));
        } else {
            for (size_t k = i; k < i + 16; k++) {
                const uint16_t s = src[k];
                dst[k] = s < 256 ? (uint8_t)s : win[s - 256];
            }
        }
    }
#endif
    for (; i < n; i++) {
        const uint16_t s = src[i];
        dst[i] = s < 256 ? (uint8_t)s : win[s - 256];
    }
    return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, (uInt)n);
}

class Reader {
public:
    // data: the whole compressed file (memory mapped by the caller).  False when the first
    // header is not one this reader handles (the caller then uses zlib from the start).
    bool open(const uint8_t *data, size_t size, int threads, size_t chunk_bytes)
    {
        in_ = data;
        size_ = size;
        threads_ = std::max(1, threads);
        chunk_ = std::max<size_t>(chunk_bytes, 4096);
        const size_t q = gzip_header(in_, size_, 0);
        if (!q) return false;
        pos_bit_ = (uint64_t)q * 8;
        hist_ = 0;
        window_.assign(WIN, 0);
        crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
        member_len_ = 0;
        return true;
    }

    // Continues a stream that another inflater (the device feed, csrc/tdg_gzdev.cuh) has brought to
    // the block boundary at bit `pos_bit`: `window` holds the WIN bytes in front of it (the last
    // `hist` of them belong to the member being inflated), `crc` / `member_len` are the member's
    // running CRC-32 and length, `delivered` the uncompressed offset reached.
    void resume(const uint8_t *data, size_t size, int threads, size_t chunk_bytes, uint64_t pos_bit, const uint8_t *window,
                size_t hist, uint32_t crc, uint64_t member_len, uint64_t delivered)
    {
        in_ = data;
        size_ = size;
        threads_ = std::max(1, threads);
        chunk_ = std::max<size_t>(chunk_bytes, 4096);
        pos_bit_ = pos_bit;
        hist_ = std::min<size_t>(hist, WIN);
        window_.assign(window, window + WIN);
        crc_ = crc;
        member_len_ = member_len;
        delivered_ = delivered;
        eof_ = fallback_ = bad_check_ = false;
        segs_.clear();
        seg_ = 0;
    }

    // Up to cap bytes into p.  >0 bytes delivered; 0 end of file; -1 the caller must continue
    // with zlib at uncompressed offset delivered(); -2 a member's CRC32 / ISIZE did not match.
    long long read(uint8_t *p, size_t cap)
    {
        size_t got = 0;
        while (got < cap) {
            if (seg_ == segs_.size()) {
                if (bad_check_) return got ? (long long)got : -2;
                if (fallback_) return got ? (long long)got : -1;
                if (eof_) break;
                round();
                continue;
            }
            got += drain(p + got, cap - got);
            if (bad_check_) return -2;
        }
        return (long long)got;
    }

    uint64_t delivered() const { return delivered_; }

private:
    struct Seg {
        int chunk;
        size_t a, b;            // symbols [a, b) of the chunk
        int mend;               // index of the member end that follows this segment, or -1
    };

    Chunk &chunk_at(size_t i)
    {
        while (pool_.size() <= i) pool_.emplace_back(new Chunk());
        return *pool_[i];
    }

    void round()
    {
        segs_.clear();
        seg_ = 0;
        accepted_ = 0;
        const size_t T = (size_t)threads_;
        const size_t base = (size_t)(pos_bit_ >> 3) / chunk_ * chunk_;          // nominal grid
        const size_t spec_cap = WIN + chunk_ * 16 + 512;
        while (infl_.size() < T + 1) infl_.emplace_back(new Inflater());
        chunk_at(T);                                       // the pool must not grow while the threads run
        for (auto &z : infl_) {
            z->in = in_;
            z->in_size = size_;
        }
        auto nominal = [&](size_t k) { return (uint64_t)std::min<size_t>(base + k * chunk_, size_) * 8; };
        // ---- phase A: chunk 0 from the exact position, the others speculatively
        std::vector<std::thread> th;
        auto work = [&](size_t k) {
            Inflater &z = *infl_[k];
            Chunk &c = chunk_at(k);
            if (k == 0) {
                c.prepare(spec_cap, true);
                start_known(z, c);
                drive(z, c, nominal(1));
                c.found = true;
            } else {
                c.prepare(spec_cap, false);
                if (nominal(k) < nominal(k + 1)) speculate(z, c, nominal(k), nominal(k + 1), nominal(k + 1));
            }
        };
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t k = 1; k < T; k++) th.emplace_back(work, k);
        work(0);
        for (auto &x : th) x.join();
        const auto t1 = std::chrono::steady_clock::now();
        // ---- phase B: chain
        size_t extra = T;                                  // pool slots for gap chunks
        for (size_t k = 0; k < T && !eof_ && !fallback_; k++) {
            Chunk *c = &chunk_at(k);
            bool take = k == 0;
            if (k > 0 && c->found && c->start_bit >= pos_bit_) {
                if (pos_bit_ < c->start_bit && !gap(extra, c->start_bit)) break;
                take = pos_bit_ == c->start_bit;
            }
            if (take) {
                accepted_++;
                accept((int)k);
            } else if (pos_bit_ < nominal(k + 1)) {
                if (!gap(extra, nominal(k + 1))) break;
            }
        }
        // a round must make progress even when everything was refused
        if (segs_.empty() && !eof_ && !fallback_ && pos_bit_ < nominal(T)) gap(extra, nominal(T));
        if (debug_) {
            const auto t2 = std::chrono::steady_clock::now();
            size_t found = 0;
            for (size_t k = 1; k < T; k++) found += chunk_at(k).found;
            fprintf(stderr, "pgz round: A %.1f ms, B %.1f ms, found %zu/%zu, accepted %zu, gaps %zu\n",
                    std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(t2 - t1).count(), found, T - 1, accepted_, extra - T);
        }
    }

    void start_known(Inflater &z, Chunk &c)
    {
        c.known_window(window_.data());
        z.seek(pos_bit_);
        z.o = WIN;
        z.member_start = WIN - hist_;
        z.last_bit = pos_bit_;
        z.last_o = WIN;
        z.last_ms = z.member_start;
        c.start_bit = pos_bit_;
    }

    // inflate serially from pos_bit_ up to a block boundary at or after stop_bit
    bool gap(size_t &slot, uint64_t stop_bit)
    {
        Chunk &c = chunk_at(slot);
        Inflater &z = *infl_[threads_];
        c.prepare(WIN + chunk_ * 8 + 512, true);
        start_known(z, c);
        drive(z, c, stop_bit);
        c.found = true;
        accept((int)slot);
        slot++;
        return !eof_ && !fallback_;
    }

    // chunk i continues the stream at pos_bit_: fix its window, queue its symbols
    void accept(int i)
    {
        Chunk &c = *pool_[(size_t)i];
        c.win_before = window_;
        // window after the chunk: the last WIN symbols of prehistory + data, resolved
        {
            std::vector<uint8_t> nw(WIN);
            const uint16_t *s = c.buf.data() + c.len - WIN;
            for (uint32_t j = 0; j < WIN; j++) nw[j] = s[j] < 256 ? (uint8_t)s[j] : window_[s[j] - 256];
            window_.swap(nw);
        }
        // bytes of the new window that belong to the member being inflated
        if (c.ms_end >= WIN) hist_ = std::min<size_t>(WIN, c.len - c.ms_end);      // a member began inside the chunk
        else hist_ = std::min<size_t>(WIN, hist_ + (c.len - WIN));
        size_t a = WIN;
        for (size_t m = 0; m < c.mends.size(); m++) {
            segs_.push_back(Seg{i, a, c.mends[m].at, (int)m});
            a = c.mends[m].at;
        }
        if (a < c.len) segs_.push_back(Seg{i, a, c.len, -1});
        pos_bit_ = c.end_bit;
        if (c.eof) eof_ = true;
        if (c.error) fallback_ = true;
    }

    // resolve queued segments into p (at most cap bytes), in parallel; fold and check CRCs
    size_t drain(uint8_t *p, size_t cap)
    {
        struct Task {
            const uint16_t *src;
            const uint8_t *win;
            uint8_t *dst;
            size_t n;
            uint32_t crc;
            int chunk, mend;
        };
        std::vector<Task> tasks;
        const auto d0 = std::chrono::steady_clock::now();
        size_t used = 0;
        const size_t piece = (size_t)1 << 20;
        while (seg_ < segs_.size() && (used < cap || segs_[seg_].a == segs_[seg_].b)) {
            Seg &s = segs_[seg_];
            Chunk &c = *pool_[(size_t)s.chunk];
            const size_t n = std::min(std::min(s.b - s.a, cap - used), piece);
            const bool done = s.a + n == s.b;
            tasks.push_back(Task{c.buf.data() + s.a, c.win_before.data(), p + used, n, 0, s.chunk, done ? s.mend : -1});
            used += n;
            s.a += n;
            if (done) seg_++;
        }
        const int nt = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, tasks.size()));
        auto work = [&](int t) {
            for (size_t i = (size_t)t; i < tasks.size(); i += (size_t)nt) {
                Task &k = tasks[i];
                k.crc = resolve(k.src, k.n, k.win, k.dst);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        for (const Task &k : tasks) {
            if (k.n) crc_ = (uint32_t)crc32_combine(crc_, k.crc, (z_off_t)k.n);
            member_len_ += k.n;
            if (k.mend >= 0) {
                const MemberEnd &m = pool_[(size_t)k.chunk]->mends[(size_t)k.mend];
                if (crc_ != m.crc || (uint32_t)member_len_ != m.isize) bad_check_ = true;
                crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
                member_len_ = 0;
            }
        }
        delivered_ += used;
        if (debug_)
            fprintf(stderr, "pgz drain: %zu bytes, %zu tasks, %.1f ms\n", used, tasks.size(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - d0).count());
        return used;
    }

    const uint8_t *in_ = nullptr;
    size_t size_ = 0, chunk_ = 0;
    int threads_ = 1;
    uint64_t pos_bit_ = 0;           // exact position of the next block header
    size_t hist_ = 0;                // bytes of window_ that belong to the current member
    std::vector<uint8_t> window_;    // the last WIN bytes delivered/queued
    bool eof_ = false, fallback_ = false, bad_check_ = false;
    uint32_t crc_ = 0;
    uint64_t member_len_ = 0, delivered_ = 0;
    std::vector<std::unique_ptr<Chunk>> pool_;
    std::vector<std::unique_ptr<Inflater>> infl_;
    std::vector<Seg> segs_;
    size_t seg_ = 0;
    size_t accepted_ = 0;
    bool debug_ = getenv("TDG_PGZ_DEBUG") != nullptr;
};

}  // namespace pgz
}  // namespace tdg
