// Host-side text checks of the file feed (plain C++, no CUDA).
//
// The reference opens its FASTQ files in TEXT mode (open(f, 'r') / gzip.open(f, 'rt'),
// /root/reference/tagdigger_fun.py:240-243): the bytes are decoded as UTF-8 (errors='strict'),
// so a file with an invalid byte sequence raises UnicodeDecodeError instead of being counted, and
// lines are split with universal newlines.  The counting kernel works on raw bytes; these helpers
// let tdg_count_file keep both behaviours:
//   - utf8_*      the validity rules of Python's UTF-8 decoder, resumable across buffers;
//   - LineLimit   where the maxreads'th read ends (tagdigger_fun.py:272-273: the loop stops
//                 there, so nothing after that point is read, decoded or inflated).
#pragma once
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "tdg_pool.h"

namespace tdg {

// Does p[0..n) hold a byte >= 0x80?  (FASTQ text almost never does: then there is nothing to validate.)
inline bool has_high_bit(const uint8_t *p, size_t n)
{
    size_t i = 0;
    uint64_t acc = 0;
    for (; i < n && ((uintptr_t)(p + i) & 7u); i++) acc |= p[i];
    for (; i + 64 <= n; i += 64) {
        const uint64_t *w = (const uint64_t *)(p + i);
        acc |= (w[0] | w[1]) | (w[2] | w[3]) | (w[4] | w[5]) | (w[6] | w[7]);
    }
    for (; i < n; i++) acc |= p[i];
    return (acc & 0x8080808080808080ull) != 0;
}

inline bool has_high_bit_mt(const uint8_t *p, size_t n, int threads)
{
    int nt = (int)std::min<size_t>((size_t)std::max(1, threads), std::max<size_t>(1, n >> 22));
    if (nt <= 1) return has_high_bit(p, n);
    std::vector<char> hit(nt, 0);
    auto work = [&](int t) {
        size_t lo = n / nt * t, hi = t == nt - 1 ? n : n / nt * (t + 1);
        hit[t] = has_high_bit(p + lo, hi - lo) ? 1 : 0;
    };
    Pool::get().run(nt, work);
    for (char h : hit)
        if (h) return true;
    return false;
}

// Resumable UTF-8 validation with the rules of CPython's decoder (= RFC 3629: no overlong forms,
// no surrogates, nothing above U+10FFFF).
struct Utf8State {
    uint32_t need = 0;        // continuation bytes still expected
    uint8_t lo = 0x80, hi = 0xBF;   // allowed range of the NEXT continuation byte
    uint64_t offset = 0;      // bytes validated so far (for the error message)
    uint64_t seq_start = 0;   // offset of the lead byte of the sequence in progress
};

// Returns -1 when p[0..n) continues a valid stream, else the offset IN THE STREAM of the byte
// that starts the invalid sequence.
inline long long utf8_feed(Utf8State &st, const uint8_t *p, size_t n)
{
    size_t i = 0;
    while (i < n) {
        if (st.need == 0) {
            // ASCII run, eight bytes at a time
            while (i + 8 <= n) {
                uint64_t w;
                memcpy(&w, p + i, 8);
                if (w & 0x8080808080808080ull) break;
                i += 8;
            }
            if (i >= n) break;
            const uint8_t c = p[i];
            if (c < 0x80) { i++; continue; }
            st.seq_start = st.offset + i;
            st.lo = 0x80;
            st.hi = 0xBF;
            if (c >= 0xC2 && c <= 0xDF) st.need = 1;
            else if (c == 0xE0) { st.need = 2; st.lo = 0xA0; }
            else if (c == 0xED) { st.need = 2; st.hi = 0x9F; }
            else if (c >= 0xE1 && c <= 0xEF) st.need = 2;
            else if (c == 0xF0) { st.need = 3; st.lo = 0x90; }
            else if (c == 0xF4) { st.need = 3; st.hi = 0x8F; }
            else if (c >= 0xF1 && c <= 0xF3) st.need = 3;
            else return (long long)st.seq_start;            // 80..BF, C0, C1, F5..FF: invalid start byte
            i++;
        } else {
            const uint8_t c = p[i];
            if (c < st.lo || c > st.hi) return (long long)st.seq_start;   // invalid continuation byte
            st.lo = 0x80;
            st.hi = 0xBF;
            st.need--;
            i++;
        }
    }
    st.offset += n;
    return -1;
}

// End of the stream: a sequence cut short is an error too ("unexpected end of data").
inline long long utf8_finish(const Utf8State &st) { return st.need ? (long long)st.seq_start : -1; }

// Where does line end number `remaining` fall?  Universal newlines: '\n', '\r\n' and a lone '\r'
// each end one line.  feed() returns the number of bytes of p[0..n) up to and including that line
// end (n when it lies beyond this buffer) and sets `reached` when the count is complete.
struct LineLimit {
    uint64_t remaining = 0;
    bool prev_cr = false;
    bool reached = false;

    size_t feed(const uint8_t *p, size_t n)
    {
        if (reached) return 0;
        size_t i = 0;
        if (prev_cr && n) {
            prev_cr = false;
            if (p[0] != '\n') {                        // the '\r' that ended the previous buffer ended a line
                if (--remaining == 0) { reached = true; return 0; }
            }
        }
        const bool any_cr = memchr(p, '\r', n) != nullptr;
        if (!any_cr) {
            while (i < n) {
                const uint8_t *q = (const uint8_t *)memchr(p + i, '\n', n - i);
                if (!q) return n;
                i = (size_t)(q - p) + 1;
                if (--remaining == 0) { reached = true; return i; }
            }
            return n;
        }
        for (; i < n; i++) {
            const uint8_t c = p[i];
            if (c == '\n') {
                if (--remaining == 0) { reached = true; return i + 1; }
            } else if (c == '\r') {
                if (i + 1 == n) { prev_cr = true; return n; }
                if (p[i + 1] != '\n') {
                    if (--remaining == 0) { reached = true; return i + 1; }
                }
            }
        }
        return n;
    }
};

}  // namespace tdg
