// Host logic of the device-side gzip feed: which chunks of a round continue the stream.
// Plain C++ (no CUDA): used by csrc/tdg_gzdev.cuh and by the CPU harness tests/native/gzlane_check.cpp.
//
// A ROUND: the compressed file (gzip.open(f) of /root/reference/tagdigger_fun.py:240-241) is cut at
// multiples of `chunk` bytes; chunk 0 of a round starts at the exact bit the stream has reached
// (pos_bit), every other chunk at the first block start its lane could confirm behind its nominal
// offset.  Lane k inflates up to the first block boundary at or behind the nominal start of chunk
// k + 1.  The output of chunk k is part of the file's text if and only if chunk k - 1 is, and the
// inflater of chunk k - 1 stopped -- at a block boundary -- on exactly the bit chunk k started
// from: then lane k did what a sequential inflater would have done from there.  A false block
// start can therefore shorten a round, never change the text.  What this code does not want to
// judge (invalid data, a buffer that filled up, a header it does not parse, too little progress)
// ends the device feed with a HANDOVER: the host reader (tdg_pgz.h) resumes at the state reached
// -- bit position, last 32 KiB, running CRC -- and whatever is wrong with the file surfaces there,
// on the code path that decides the error behaviour for every gzip file.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "tdg_gzlane.h"
#include "tdg_pgz.h"

namespace tdg {
namespace gzc {

// A chunk the lanes could not deliver (no block start confirmed in it, or one that was not the
// stream's) does not end the round: the HOST inflates it -- one chunk is half a millisecond of
// zlib-class work here, and a lane's worth of waiting (tens of milliseconds) on the device --
// into the same 16-bit symbols, from the bit the lane before it stopped at to the boundary the
// next lane must have started from.
struct Repair {
    uint32_t chunk;
    std::vector<uint16_t> syms;
};
constexpr size_t MAX_REPAIRS = 16;

// symbols of the blocks from pos_bit up to the first boundary at or behind stop_bit that stands in
// front of a non-final dynamic block (the lanes' stopping rule); false when that cannot be done
// here (invalid data, the end of a member, the end of the input)
inline bool host_gap(const uint8_t *in, size_t size, uint64_t pos_bit, uint64_t stop_bit, uint64_t hist, std::vector<uint16_t> &syms,
                     uint64_t &end_bit)
{
    std::unique_ptr<pgz::Inflater> z(new pgz::Inflater());
    pgz::Chunk c;
    z->in = in;
    z->in_size = size;
    c.prepare(pgz::WIN + ((size_t)1 << 20), true);
    c.unknown_window();
    z->out = c.buf.data();
    z->cap = c.buf.size();
    z->seek(pos_bit);
    z->o = pgz::WIN;
    z->member_start = pgz::WIN - (size_t)std::min<uint64_t>(pgz::WIN, hist);
    uint64_t stop = stop_bit;
    for (;;) {
        const pgz::Status st = z->run(stop);
        if (st == pgz::ST_SPACE) {
            if (c.buf.size() > ((size_t)64 << 20)) return false;
            c.buf.resize(c.buf.size() * 2);
            z->out = c.buf.data();
            z->cap = c.buf.size();
            continue;
        }
        if (st != pgz::ST_STOP) return false;
        // at a boundary: is the next block one a lane's scan accepts?
        const uint64_t bp = z->bitpos();
        if (bp + 3 > (uint64_t)size * 8) return false;
        const uint32_t h = (uint32_t)((in[bp >> 3] | (uint32_t)in[(bp >> 3) + 1 < size ? (bp >> 3) + 1 : (bp >> 3)] << 8) >> (bp & 7)) & 7u;
        if (h == 4u) {
            end_bit = bp;
            syms.assign(c.buf.data() + pgz::WIN, c.buf.data() + z->o);
            return true;
        }
        stop = bp + 1;                                       // run through this block as well
    }
}

struct Round {
    size_t chunk = 0;          // nominal chunk size in bytes (multiple of 16)
    size_t grid = 0;           // file offset of the nominal start of chunk 0 (multiple of chunk)
    uint32_t nchunks = 0;
    size_t buf_off = 0;        // the bytes [buf_off, buf_end) of the file are what the lanes see
    size_t buf_end = 0;
    uint64_t pos_bit = 0;      // exact start of chunk 0
    uint32_t hist = 0;         // bytes of real history in front of chunk 0 that belong to its member
    uint64_t nominal(size_t k, size_t size) const { return (uint64_t)std::min<size_t>(grid + k * chunk, size) * 8; }
};

struct Outcome {
    uint32_t accepted = 0;             // chunks 0..accepted-1 continue the stream
    std::vector<uint64_t> text_off;    // [accepted + 1] offsets of their bytes in the round's text
    std::vector<uint32_t> lens;        // [accepted] symbols each of them contributes (0: the lane before ran through the whole chunk)
    uint64_t end_bit = 0;              // where the last of them stopped
    std::vector<Repair> repairs;       // chunks whose symbols the host made (their lens entry says how many)
    bool member_end = false;           // the last accepted chunk ends a member (trailer follows at end_bit)
    bool handover = false;             // the host reader must take over behind the accepted chunks
    const char *why = "";
};

struct Stream {
    const uint8_t *in = nullptr;       // the whole compressed file
    size_t size = 0;
    uint64_t pos_bit = 0;
    uint32_t hist = 0;
    uint32_t crc = 0;                  // CRC-32 of the current member's bytes so far
    uint64_t member_len = 0, delivered = 0;
    bool eof = false, handover = false, bad_check = false;
    bool to_zlib = false;              // handover without a resumable position: zlib re-reads the file up to `delivered`
    int poor_rounds = 0;
    const char *why = "";

    bool open(const uint8_t *data, size_t n)
    {
        in = data;
        size = n;
        const size_t q = pgz::gzip_header(in, size, 0);
        if (!q) return false;
        pos_bit = (uint64_t)q * 8;
        hist = 0;
        crc = (uint32_t)crc32(0L, Z_NULL, 0);
        member_len = delivered = 0;
        eof = handover = bad_check = to_zlib = false;
        poor_rounds = 0;
        return true;
    }

    Round plan(size_t chunk, uint32_t max_chunks) const
    {
        Round r;
        r.chunk = chunk;
        r.grid = (size_t)(pos_bit >> 3) / chunk * chunk;
        const size_t left = size - r.grid;
        r.nchunks = (uint32_t)std::min<size_t>(max_chunks, (left + chunk - 1) / chunk);
        r.buf_off = r.grid;
        r.buf_end = std::min(size, r.grid + ((size_t)r.nchunks + 1) * chunk);      // one chunk of slack: the last lane finishes its block
        r.pos_bit = pos_bit;
        r.hist = hist;
        return r;
    }

    // meta[k]: lane k's report with ABSOLUTE bit positions
    Outcome chain(const Round &r, const gzl::Meta *meta, size_t sym_cap) const
    {
        Outcome o;
        o.text_off.push_back(0);
        uint64_t pos = r.pos_bit;
        uint64_t h = r.hist;
        o.end_bit = pos;
        for (uint32_t k = 0; k < r.nchunks; k++) {
            const gzl::Meta &c = meta[k];
            if (k > 0 && pos >= r.nominal(k + 1, size)) {
                // no block starts inside this chunk: the lane before it went through all of it
                o.accepted = k + 1;
                o.text_off.push_back(o.text_off.back());
                o.lens.push_back(0);
                continue;
            }
            if (!(c.flags & gzl::F_FOUND) || c.start_bit != pos) {
                // the lane of this chunk did not start where the stream is: the host fills in
                if (o.repairs.size() >= MAX_REPAIRS || k == 0) break;
                Repair rp;
                rp.chunk = k;
                uint64_t e = 0;
                if (!host_gap(in, size, pos, r.nominal(k + 1, size), h, rp.syms, e)) break;
                if (rp.syms.size() > sym_cap) break;
                o.accepted = k + 1;
                o.text_off.push_back(o.text_off.back() + rp.syms.size());
                o.lens.push_back((uint32_t)rp.syms.size());
                h += rp.syms.size();
                pos = e;
                o.end_bit = pos;
                o.repairs.push_back(std::move(rp));
                continue;
            }
            if (k > 0 && c.min_pre < gzl::WIN - std::min<uint64_t>(gzl::WIN, h)) {
                // a distance reaches in front of its member: zlib calls that invalid
                o.handover = true;
                o.why = "distance too far back";
                break;
            }
            if (c.end_bit == c.start_bit && !(c.flags & gzl::F_FINAL)) {
                // not one block done
                if (c.flags & (gzl::F_ERROR | gzl::F_SPACE)) {
                    o.handover = true;
                    o.why = (c.flags & gzl::F_ERROR) ? "invalid data" : "symbol buffer full";
                } else if ((c.flags & gzl::F_INPUT) && r.buf_end == size) {
                    o.handover = true;
                    o.why = "stream ends early";
                } else if (k == 0) {
                    o.handover = true;                      // a round must make progress
                    o.why = "no progress";
                }
                break;
            }
            o.accepted = k + 1;
            o.text_off.push_back(o.text_off.back() + c.out_len);
            o.lens.push_back(c.out_len);
            pos = c.end_bit;
            o.end_bit = pos;
            h += c.out_len;
            if (c.flags & gzl::F_FINAL) {
                o.member_end = true;
                break;
            }
            if (c.flags & (gzl::F_ERROR | gzl::F_SPACE)) {
                o.handover = true;
                o.why = (c.flags & gzl::F_ERROR) ? "invalid data" : "symbol buffer full";
                break;
            }
            if (c.flags & gzl::F_INPUT) {
                if (r.buf_end == size) {
                    o.handover = true;
                    o.why = "stream ends early";
                }
                break;
            }
        }
        if (o.accepted == 0 && !o.handover) {
            o.handover = true;
            o.why = "no progress";
        }
        return o;
    }

    // The accepted chunks' bytes (text_len of them, CRC-32 text_crc) are part of the file: move on.
    // Returns false when a member's trailer does not match (crc / length).
    bool advance(const Round &r, const Outcome &o, uint64_t text_len, uint32_t text_crc)
    {
        if (o.accepted) {
            if (text_len) crc = (uint32_t)crc32_combine(crc, text_crc, (z_off_t)text_len);
            member_len += text_len;
            delivered += text_len;
            hist = (uint32_t)std::min<uint64_t>(gzl::WIN, (uint64_t)hist + text_len);
            pos_bit = o.end_bit;
            // poor rounds: the lanes found little to do (huge blocks, stored data ...)
            const uint64_t span = pos_bit - r.pos_bit, want = r.nominal(r.nchunks, size) - r.pos_bit;
            if (!o.member_end && r.nchunks >= 8 && span * 4 < want) poor_rounds++;
            else poor_rounds = 0;
        }
        if (o.member_end) {
            const size_t t = (size_t)((pos_bit + 7) >> 3);
            if (t + 8 > size) {
                handover = to_zlib = true;                   // truncated trailer: zlib reports it
                why = "truncated trailer";
                return true;
            }
            uint32_t want_crc, want_len;
            memcpy(&want_crc, in + t, 4);
            memcpy(&want_len, in + t + 4, 4);
            if (want_crc != crc || want_len != (uint32_t)member_len) {
                bad_check = true;
                return false;
            }
            const size_t p = t + 8;
            crc = (uint32_t)crc32(0L, Z_NULL, 0);
            member_len = 0;
            hist = 0;
            // the end of the file, or bytes that do not start with the gzip magic (ignored, as gzread does)
            if (p + 1 >= size || in[p] != 0x1f || in[p + 1] != 0x8b) {
                eof = true;
                return true;
            }
            const size_t q = pgz::gzip_header(in, size, p);
            if (!q) {
                pos_bit = (uint64_t)p * 8;
                handover = to_zlib = true;                   // a header this code does not parse: zlib judges it
                why = "member header";
                return true;
            }
            pos_bit = (uint64_t)q * 8;
        }
        if (o.handover) {
            handover = true;
            why = o.why;
        } else if (poor_rounds >= 2) {
            handover = true;
            why = "little progress per round";
        }
        return true;
    }
};

}  // namespace gzc
}  // namespace tdg
