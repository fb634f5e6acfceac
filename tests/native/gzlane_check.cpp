// TEST HARNESS, not product code: runs the device-side gzip feed's per-lane inflater
// (tagdigger_b200/csrc/tdg_gzlane.h, compiled for the host with STRIDE = 1) and its round / chain
// logic (tdg_gzchain.h) on the CPU, one "lane" after the other, so that the code the GPU kernel
// executes per lane can be held to zlib without a GPU.  Built by tests/gzlane_check.py.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../tagdigger_b200/csrc/tdg_gzchain.h"

using namespace tdg;

namespace {

// candidates of chunk k: bit offsets from `from` (relative to the buffer) that pass quick_test and
// whose dynamic block header parses; the search stops at the first max_cand of them
void scan(const std::vector<uint8_t> &buf, uint64_t nwords, uint64_t in_bits, uint64_t from, uint64_t to, const uint8_t *kraft3,
          uint16_t *mem, uint16_t *cold, std::vector<uint32_t> &cand, size_t max_cand)
{
    cand.clear();
    for (uint64_t p = from; p < to && cand.size() < max_cand; p++) {
        const size_t byte = (size_t)(p >> 3);
        if (byte + 16 > buf.size()) break;
        uint64_t a, b;
        memcpy(&a, buf.data() + byte, 8);
        memcpy(&b, buf.data() + byte + 8, 8);
        const int sh = (int)(p & 7);
        const uint64_t lo = sh ? (a >> sh | b << (64 - sh)) : a;
        const uint32_t hi = (uint32_t)(b >> sh);
        if (!gzl::quick_test(lo, hi, kraft3)) continue;
        if (gzl::header_check<1>(gzl::Mem<1>{mem, cold}, (const uint32_t *)buf.data(), nwords, in_bits, p)) cand.push_back((uint32_t)(p - from));
    }
}

}  // namespace

extern "C" {

// Inflates the gzip file image gz[0..n) the way the device feed does.  Returns the number of bytes
// written to out, or: -10 first header not handled, -2 CRC / length mismatch, -1 the stream needs
// zlib (out holds the bytes delivered so far: info[4]), -100 out too small.
// info: [0] rounds, [1] chunks accepted, [2] handover (0 none, 1 resumed by tdg_pgz, 2 straight to zlib),
//       [3] chunks the host inflated in place of a lane, [4] bytes delivered before zlib was asked for, [5] chunks run
long long gzl_inflate(const uint8_t *gz, size_t n, size_t chunk, uint32_t max_chunks, size_t search_bytes, uint32_t symcap,
                      uint8_t *out, size_t cap, long long *info, char *why, size_t why_cap, uint32_t blind_every)
{
    for (int i = 0; i < 6; i++) info[i] = 0;
    if (why && why_cap) why[0] = 0;
    gzc::Stream st;
    if (!st.open(gz, n)) return -10;
    uint8_t kraft3[512];
    gzl::make_kraft3(kraft3);
    std::vector<uint8_t> window(gzl::WIN, 0);
    std::vector<uint16_t> mem(gzl::LANE_U16), cold(gzl::COLD_U16);
    std::vector<std::vector<uint16_t>> syms;
    std::vector<uint16_t> toks;
    std::vector<gzl::Meta> meta;
    std::vector<uint32_t> cand;
    size_t total = 0;
    while (!st.eof && !st.handover) {
        const gzc::Round r = st.plan(chunk, max_chunks);
        info[0]++;
        const size_t nb = r.buf_end - r.buf_off;
        std::vector<uint8_t> buf(((nb + 3) / 4) * 4 + 64, 0);
        memcpy(buf.data(), gz + r.buf_off, nb);
        const uint64_t nwords = (nb + 3) / 4;
        const uint64_t base_bit = (uint64_t)r.buf_off * 8;
        if (syms.size() < r.nchunks) syms.resize(r.nchunks);
        meta.assign(r.nchunks, gzl::Meta());
        for (uint32_t k = 0; k < r.nchunks; k++) {
            syms[k].resize(symcap);
            toks.resize(symcap);
            gzl::Mem<1> m{mem.data(), cold.data()};
            const uint64_t stop = r.nominal(k + 1, n) - base_bit;
            if (k == 0) {
                gzl::run_chunk<1>(m, (const uint32_t *)buf.data(), nwords, (uint64_t)nb * 8, true, r.pos_bit - base_bit, nullptr, 0,
                                  stop, r.hist, toks.data(), symcap, symcap, meta[k]);
            } else {
                const uint64_t from = r.nominal(k, n) - base_bit;
                const uint64_t to = std::min<uint64_t>(from + (uint64_t)search_bytes * 8, stop);
                scan(buf, nwords, (uint64_t)nb * 8, from, to, kraft3, mem.data(), cold.data(), cand, 2);
                if (blind_every && k % blind_every == 0) cand.clear();          // (test: a lane that finds no start; the host fills in)
                gzl::run_chunk<1>(m, (const uint32_t *)buf.data(), nwords, (uint64_t)nb * 8, false, from, cand.data(),
                                  (uint32_t)cand.size(), stop, 0, toks.data(), symcap, symcap, meta[k]);
            }
            if (gzl::expand_tokens(toks.data(), meta[k].ntok, syms[k].data()) != meta[k].out_len) return -50;
            meta[k].start_bit += base_bit;
            meta[k].end_bit += base_bit;
            info[5]++;
        }
        const gzc::Outcome o = st.chain(r, meta.data(), symcap);
        for (const gzc::Repair &rp : o.repairs) {
            std::copy(rp.syms.begin(), rp.syms.end(), syms[rp.chunk].begin());
            info[2]++;
        }
        info[1] += o.accepted;
        // resolve the accepted chunks against the window handed from chunk to chunk
        const size_t text_at = total;
        for (uint32_t k = 0; k < o.accepted; k++) {
            const uint32_t len = o.lens[k];
            if (total + len > cap) return -100;
            for (uint32_t i = 0; i < len; i++) {
                const uint16_t s = syms[k][i];
                out[total + i] = s < 256 ? (uint8_t)s : window[s - 256];
            }
            total += len;
            if (len >= gzl::WIN) memcpy(window.data(), out + total - gzl::WIN, gzl::WIN);
            else {
                memmove(window.data(), window.data() + len, gzl::WIN - len);
                memcpy(window.data() + gzl::WIN - len, out + total - len, len);
            }
        }
        const uint64_t text_len = total - text_at;
        const uint32_t text_crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), out + text_at, (uInt)text_len);
        if (!st.advance(r, o, text_len, text_crc)) return -2;
    }
    const long long repairs = info[2];
    info[2] = 0;
    info[3] = repairs;
    if (st.handover) {
        if (why && why_cap) {
            strncpy(why, st.why, why_cap - 1);
            why[why_cap - 1] = 0;
        }
        info[4] = (long long)total;
        if (st.to_zlib) {
            info[2] = 2;
            return -1;
        }
        info[2] = 1;
        pgz::Reader rd;
        rd.resume(gz, n, 2, 1 << 16, st.pos_bit, window.data(), st.hist, st.crc, st.member_len, st.delivered);
        for (;;) {
            if (total == cap) return -100;
            const long long got = rd.read(out + total, cap - total);
            if (got == 0) break;
            if (got < 0) {
                info[4] = (long long)rd.delivered();
                return got;              // -1 zlib needed, -2 bad check
            }
            total += (size_t)got;
        }
    }
    return (long long)total;
}

}  // extern "C"
