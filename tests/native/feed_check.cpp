// TEST HARNESS, not product code: drives the product's host feeder (tagdigger_b200/csrc/tdg_feed.h)
// without a GPU, so that CPU tests can compare the bytes it delivers with Python's own reading of
// the same file.  Built into tests/native/libfeed_check.so by tests/feed_check.py.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../tagdigger_b200/csrc/tdg_feed.h"

#include "../../tagdigger_b200/csrc/tdg_text.h"

static std::string g_err;

extern "C" {

// Reads the whole file through Feeder in chunks of `chunk` bytes.  Returns the number of bytes
// written to out (at most cap), or a negative code (message via fck_error).  *mode receives the
// feeder mode at the end (0 plain, 1 plain sequential, 2 BGZF, 3 zlib).
long long fck_read(const char *path, int gz, size_t chunk, uint8_t *out, size_t cap, int *mode)
{
    tdg::Feeder f;
    int rc = f.open(path, gz != 0);
    if (rc) { g_err = f.error(); return rc; }
    std::vector<uint8_t> buf(chunk);
    size_t total = 0;
    for (;;) {
        long long r = f.fill(buf.data(), chunk);
        if (r < 0) { g_err = f.error(); return r; }
        if (r == 0) break;
        if (total + (size_t)r > cap) { g_err = "output buffer too small"; return -100; }
        memcpy(out + total, buf.data(), (size_t)r);
        total += (size_t)r;
    }
    if (mode) *mode = (int)f.mode();
    return (long long)total;
}

const char *fck_error(void) { return g_err.c_str(); }

// tdg_text.h: UTF-8 validation of `data` fed in pieces of `piece` bytes.  Returns -1 (valid) or
// the stream offset of the lead byte of the first invalid sequence.
long long fck_utf8(const uint8_t *data, size_t n, size_t piece)
{
    tdg::Utf8State st;
    for (size_t at = 0; at < n; at += piece) {
        size_t m = n - at < piece ? n - at : piece;
        if (!tdg::has_high_bit(data + at, m) && st.need == 0) { st.offset += m; continue; }   // as tdg_count_file does
        long long bad = tdg::utf8_feed(st, data + at, m);
        if (bad >= 0) return bad;
    }
    return tdg::utf8_finish(st);
}

// tdg_text.h: bytes of `data` (fed in pieces) up to and including line end number `lines`;
// n + 1 when the data holds fewer line ends.
long long fck_line_limit(const uint8_t *data, size_t n, size_t piece, uint64_t lines)
{
    tdg::LineLimit ll;
    ll.remaining = lines;
    size_t used = 0;
    for (size_t at = 0; at < n; at += piece) {
        size_t m = n - at < piece ? n - at : piece;
        used += ll.feed(data + at, m);
        if (ll.reached) return (long long)used;
    }
    return (long long)n + 1;
}

}  // extern "C"
