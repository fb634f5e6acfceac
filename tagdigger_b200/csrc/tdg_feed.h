// Host feed of tdg_count_file: the bytes of one FASTQ file, uncompressed, delivered in
// chunks into (pinned) buffers by several host threads.  Plain C++, no CUDA.
//
// Replaces the file side of find_tags_fastq (/root/reference/tagdigger_fun.py:240-243:
// gzip.open(f, 'rt') or open(f, 'r'), one Python thread, ~0.15 GB/s of inflate):
//   - plain files: the chunk is read with parallel pread() calls, one slice per thread;
//   - BGZF files (bgzip: gzip members of <= 64 KiB that carry their own size in a 'BC'
//     extra field): members are found without inflating and inflated in parallel, each
//     straight into its place in the chunk, CRC32 checked;
//   - any other gzip file of some size: speculative parallel inflate of the one deflate stream
//     (tdg_pgz.h: block-start search, 16-bit marker symbols for the unknown window, chained and
//     CRC-checked);
//   - small gzip files, one I/O thread, and everything tdg_pgz.h declines to judge (corrupt or
//     unusual streams): zlib's gzread on one thread, which also handles concatenated members
//     like Python's gzip.
// A BGZF file that turns into ordinary gzip members half way is continued sequentially.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "tdg_pgz.h"
#include "tdg_pool.h"

namespace tdg {

class Feeder {
public:
    enum Mode { PLAIN, PLAIN_SEQ, BGZF, GZ, PGZ };

    ~Feeder() { close(); }

    // 0 or a negative code (-3 I/O, -6 gzip) with `err` set
    int open(const char *path, bool gz)
    {
        path_ = path;
        threads_ = 16;
        if (const char *e = getenv("TDG_IO_THREADS")) threads_ = std::max(1, atoi(e));
        unsigned hw = std::thread::hardware_concurrency();
        if (hw && threads_ > (int)hw) threads_ = (int)hw;
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) return fail(-3, std::string("cannot open ") + path);
        struct stat st;
        bool regular = fstat(fd_, &st) == 0 && S_ISREG(st.st_mode);
        size_ = regular ? (uint64_t)st.st_size : 0;
        if (!gz) {
            mode_ = regular ? PLAIN : PLAIN_SEQ;
            return 0;
        }
        mode_ = GZ;
        if (regular && threads_ > 1) {
            uint8_t head[64];
            ssize_t n = pread(fd_, head, sizeof head, 0);
            uint32_t csize = 0, hlen = 0;
            if (n >= 18 && bgzf_header(head, (size_t)n, csize, hlen)) mode_ = BGZF;
        }
        if (mode_ == GZ && regular && threads_ > 1) {
            // parallel inflate of an ordinary gzip stream (TDG_PGZ_MIN: smallest file, TDG_PGZ_CHUNK:
            // nominal compressed bytes per thread and round; both for tests)
            uint64_t min_size = (uint64_t)4 << 20;
            size_t pchunk = (size_t)2 << 20;
            if (const char *e = getenv("TDG_PGZ_MIN")) min_size = strtoull(e, nullptr, 10);
            if (const char *e = getenv("TDG_PGZ_CHUNK")) pchunk = (size_t)strtoull(e, nullptr, 10);
            if (size_ >= min_size && size_ > 18) {
                void *m = mmap(nullptr, (size_t)size_, PROT_READ, MAP_PRIVATE, fd_, 0);
                if (m != MAP_FAILED) {
                    map_ = (const uint8_t *)m;
                    madvise(m, (size_t)size_, MADV_SEQUENTIAL);
                    pgz_.reset(new pgz::Reader());
                    if (pgz_->open(map_, (size_t)size_, threads_, pchunk)) mode_ = PGZ;
                    else unmap();
                }
            }
        }
        if (mode_ == GZ) return open_gz(0);
        return 0;
    }

    // Continue a gzip file behind what another inflater (the device feed, tdg_gzdev.cuh) has
    // delivered: with `window` at the block boundary `pos_bit` (parallel reader), or -- window
    // null -- with zlib re-reading the file up to the uncompressed offset `delivered`.
    int open_resume(const char *path, uint64_t pos_bit, const uint8_t *window, size_t hist, uint32_t crc, uint64_t member_len,
                    uint64_t delivered)
    {
        path_ = path;
        threads_ = 16;
        if (const char *e = getenv("TDG_IO_THREADS")) threads_ = std::max(1, atoi(e));
        unsigned hw = std::thread::hardware_concurrency();
        if (hw && threads_ > (int)hw) threads_ = (int)hw;
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) return fail(-3, std::string("cannot open ") + path);
        struct stat st;
        if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) return fail(-3, std::string("not a regular file: ") + path);
        size_ = (uint64_t)st.st_size;
        mode_ = GZ;
        if (window) {
            size_t pchunk = (size_t)2 << 20;
            if (const char *e = getenv("TDG_PGZ_CHUNK")) pchunk = (size_t)strtoull(e, nullptr, 10);
            void *m = mmap(nullptr, (size_t)size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m != MAP_FAILED) {
                map_ = (const uint8_t *)m;
                pgz_.reset(new pgz::Reader());
                pgz_->resume(map_, (size_t)size_, threads_, pchunk, pos_bit, window, hist, crc, member_len, delivered);
                mode_ = PGZ;
                return 0;
            }
        }
        int rc = open_gz(delivered);
        gz_first_ = false;
        return rc;
    }

    // BGZF member at h: its size and the length of its header (public form of bgzf_header)
    static bool bgzf_member(const uint8_t *h, size_t n, uint32_t &csize, uint32_t &hlen) { return bgzf_header(h, n, csize, hlen); }

    // Continue a BGZF file at the member that starts at file offset `off` (`delivered`
    // uncompressed bytes went out already: the device feed's, tdg_gzdev.cuh).
    int open_resume_bgzf(const char *path, uint64_t off, uint64_t delivered)
    {
        int rc = open(path, true);
        if (rc) return rc;
        if (off == 0) return 0;
        if (mode_ != BGZF) {
            // (one thread, or a first header that is not BGZF after all): zlib from the uncompressed offset
            close();
            fd_ = ::open(path, O_RDONLY);
            if (fd_ < 0) return fail(-3, std::string("cannot open ") + path);
            mode_ = GZ;
            rc = open_gz(delivered);
            gz_first_ = false;
            return rc;
        }
        cbuf_.clear();
        cused_ = 0;
        cpos_ = off;
        delivered_ = delivered;
        return 0;
    }

    // does the file image start with a BGZF member?
    static bool is_bgzf(const uint8_t *h, size_t n)
    {
        uint32_t csize = 0, hlen = 0;
        return n >= 18 && bgzf_header(h, n, csize, hlen);
    }

    Mode mode() const { return mode_; }
    int threads() const { return threads_; }
    const std::string &error() const { return err_; }

    // Next chunk: up to cap bytes into p.  Returns the number of bytes (0 = end of file) or a
    // negative code with error() set.
    long long fill(uint8_t *p, size_t cap)
    {
        if (deferred_) {                  // the bytes before the defect went out with the previous call
            int code = deferred_;
            deferred_ = 0;
            return code;
        }
        switch (mode_) {
        case PLAIN: return fill_plain(p, cap);
        case PLAIN_SEQ: return fill_seq(p, cap);
        case BGZF: return fill_bgzf(p, cap);
        case PGZ: return fill_pgz(p, cap);
        default: return fill_gz(p, cap);
        }
    }

    void close()
    {
        if (zf_) gzclose(zf_);
        zf_ = nullptr;
        unmap();
        if (fd_ >= 0) ::close(fd_);
        fd_ = -1;
    }

private:
    struct Block {
        size_t at;        // offset of the member inside cbuf_
        uint32_t csize;   // whole member
        uint32_t hlen;    // header length (deflate data starts here)
        uint32_t isize;   // uncompressed size
        size_t out;       // offset in the output chunk
    };

    int fail(int code, const std::string &msg)
    {
        err_ = msg;
        return code;
    }

    // A defect `got` bytes into a chunk: hand out the good bytes first and report the error with
    // the next call, as a reader that walks the stream line by line would meet them (a caller
    // that stops inside those bytes -- maxreads -- never sees the error, like the reference).
    long long fail_after(size_t got, int code, const std::string &msg)
    {
        err_ = msg;
        if (got == 0) return code;
        deferred_ = code;
        return (long long)got;
    }

    // Is this the header of a BGZF member?  csize = size of the whole member, hlen = header bytes.
    static bool bgzf_header(const uint8_t *h, size_t n, uint32_t &csize, uint32_t &hlen)
    {
        if (n < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
        if (h[3] & ~4u) return false;                       // name/comment/hcrc flags: not what bgzip writes
        uint32_t xlen = h[10] | (h[11] << 8);
        if (12 + xlen > n) return false;
        for (uint32_t p = 12; p + 4 <= 12 + xlen;) {
            uint32_t slen = h[p + 2] | (h[p + 3] << 8);
            if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) {
                csize = (uint32_t)(h[p + 4] | (h[p + 5] << 8)) + 1u;
                hlen = 12 + xlen;
                return csize >= hlen + 8;
            }
            p += 4 + slen;
        }
        return false;
    }

    long long fill_plain(uint8_t *p, size_t cap)
    {
        if (pos_ >= size_) return fill_seq_at(p, cap);      // the file may have grown: plain reads from here
        size_t n = (size_t)std::min<uint64_t>(cap, size_ - pos_);
        int nt = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, n >> 20));   // >= 1 MiB per thread (pooled threads: a loop costs microseconds to start)
        std::vector<long long> got(nt, 0);
        auto work = [&](int t) {
            size_t lo = n / nt * t, hi = t == nt - 1 ? n : n / nt * (t + 1);
            size_t done = lo;
            while (done < hi) {
                ssize_t r = pread(fd_, p + done, hi - done, (off_t)(pos_ + done));
                if (r < 0) { got[t] = -1; return; }
                if (r == 0) break;
                done += (size_t)r;
            }
            got[t] = (long long)(done - lo);
        };
        Pool::get().run(nt, work);
        size_t total = 0;
        for (int t = 0; t < nt; t++) {
            if (got[t] < 0) return fail(-3, "read error on " + path_);
            size_t want = (t == nt - 1 ? n : n / nt * (t + 1)) - n / nt * t;
            total += (size_t)got[t];
            if ((size_t)got[t] < want) break;               // the file shrank: what follows is not contiguous
        }
        pos_ += total;
        return (long long)total;
    }

    long long fill_seq_at(uint8_t *p, size_t cap)
    {
        size_t done = 0;
        while (done < cap) {
            ssize_t r = pread(fd_, p + done, cap - done, (off_t)(pos_ + done));
            if (r < 0) return fail(-3, "read error on " + path_);
            if (r == 0) break;
            done += (size_t)r;
        }
        pos_ += done;
        return (long long)done;
    }

    long long fill_seq(uint8_t *p, size_t cap)          // pipes and other non-regular files
    {
        size_t done = 0;
        while (done < cap) {
            ssize_t r = read(fd_, p + done, cap - done);
            if (r < 0) return fail(-3, "read error on " + path_);
            if (r == 0) break;
            done += (size_t)r;
        }
        return (long long)done;
    }

    int open_gz(uint64_t skip)
    {
        zf_ = gzopen(path_.c_str(), "rb");
        if (!zf_) return fail(-3, "cannot open " + path_);
        gzbuffer(zf_, 1 << 20);
        gz_first_ = skip == 0;
        if (skip && gzseek(zf_, (z_off_t)skip, SEEK_SET) < 0) return fail(-6, "gzip error in " + path_ + ": seek failed");
        return 0;
    }

    long long fill_gz(uint8_t *p, size_t cap)
    {
        size_t got = 0;
        while (got < cap) {
            unsigned want = (unsigned)std::min<size_t>(cap - got, 1u << 30);
            int r = gzread(zf_, p + got, want);
            if (r < 0) {
                int en = 0;
                const char *m = gzerror(zf_, &en);
                return fail_after(got, -6, "gzip error in " + path_ + ": " + (m ? m : "?"));
            }
            if (r == 0) {
                // a stream that stops before its end-of-stream marker or inside its trailer: gzread
                // hands out what it has and only notes Z_BUF_ERROR; Python's gzip raises EOFError
                int en = 0;
                const char *m = gzerror(zf_, &en);
                if (en == Z_BUF_ERROR)
                    return fail_after(got, -6, "gzip error in " + path_ + ": " + (m && *m ? m : "unexpected end of file"));
                break;
            }
            got += (size_t)r;
            if (gz_first_) {
                gz_first_ = false;
                if (gzdirect(zf_)) return fail(-6, "Not a gzipped file: " + path_);   // the reference's gzip.open raises
            }
        }
        return (long long)got;
    }

    void unmap()
    {
        pgz_.reset();
        if (map_) munmap(const_cast<uint8_t *>(map_), (size_t)size_);
        map_ = nullptr;
    }

    long long fill_pgz(uint8_t *p, size_t cap)
    {
        long long r;
        try {
            r = pgz_->read(p, cap);
        } catch (const std::bad_alloc &) {
            return fail(-3, "out of memory while inflating " + path_);
        }
        if (r >= 0) return r;
        if (r == -2) return fail(-6, "gzip error in " + path_ + ": incorrect data check");
        // something tdg_pgz.h leaves to zlib: continue sequentially at the offset reached
        const uint64_t at = pgz_->delivered();
        unmap();
        mode_ = GZ;
        int rc = open_gz(at);
        if (rc) return rc;
        gz_first_ = false;
        return fill_gz(p, cap);
    }

    // make sure cbuf_ holds file bytes [cpos_ + at, cpos_ + at + need) if the file has them
    bool window(size_t &at, size_t need)
    {
        if (at + need <= cbuf_.size()) return true;
        // drop what has been consumed, then read ahead
        if (cused_ > 0) {
            cbuf_.erase(cbuf_.begin(), cbuf_.begin() + (long)cused_);
            cpos_ += cused_;
            at -= cused_;
            for (auto &b : blocks_) b.at -= cused_;
            cused_ = 0;
        }
        uint64_t have_to = cpos_ + cbuf_.size();
        if (have_to >= size_) return at + need <= cbuf_.size();
        size_t want = std::max<size_t>(need, (size_t)16 << 20);
        want = (size_t)std::min<uint64_t>(want, size_ - have_to);
        size_t old = cbuf_.size();
        cbuf_.resize(old + want);
        size_t done = 0;
        while (done < want) {
            ssize_t r = pread(fd_, cbuf_.data() + old + done, want - done, (off_t)(have_to + done));
            if (r <= 0) break;
            done += (size_t)r;
        }
        cbuf_.resize(old + done);
        return at + need <= cbuf_.size();
    }

    long long fill_bgzf(uint8_t *p, size_t cap)
    {
        // ---- collect members while their output fits the chunk
        blocks_.clear();
        size_t at = cused_, out = 0;
        bool foreign = false;
        for (;;) {
            if (cpos_ + at >= size_) break;                          // end of file
            if (!window(at, 18)) { foreign = true; break; }          // a tail too short for a member: let zlib judge it
            uint32_t csize = 0, hlen = 0;
            size_t avail = cbuf_.size() - at;
            if (!bgzf_header(cbuf_.data() + at, std::min<size_t>(avail, 1024), csize, hlen)) { foreign = true; break; }
            if (!window(at, csize)) { foreign = true; break; }       // truncated member
            const uint8_t *m = cbuf_.data() + at;
            uint32_t isize = m[csize - 4] | (m[csize - 3] << 8) | (m[csize - 2] << 16) | ((uint32_t)m[csize - 1] << 24);
            if (isize > 65536) { foreign = true; break; }
            if (out + isize > cap) {
                if (blocks_.empty()) foreign = true;                 // a chunk smaller than one member: zlib path
                break;
            }
            blocks_.push_back(Block{at, csize, hlen, isize, out});
            at += csize;
            out += isize;
        }
        if (blocks_.empty() && foreign) {
            // not (or no longer) BGZF: the rest of the file goes through zlib sequentially,
            // from the uncompressed position reached so far
            mode_ = GZ;
            int rc = open_gz(delivered_);
            if (rc) return rc;
            gz_first_ = false;
            return fill_gz(p, cap);
        }
        // ---- inflate them in parallel
        int nt = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, blocks_.size() / 4));
        std::vector<int> bad(nt, 0);
        auto work = [&](int t) {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) { bad[t] = 1; return; }
            for (size_t i = (size_t)t; i < blocks_.size(); i += (size_t)nt) {
                const Block &b = blocks_[i];
                const uint8_t *m = cbuf_.data() + b.at;
                inflateReset(&z);
                z.next_in = const_cast<Bytef *>(m + b.hlen);
                z.avail_in = b.csize - b.hlen - 8;
                z.next_out = p + b.out;
                z.avail_out = b.isize;
                int r = inflate(&z, Z_FINISH);
                uint32_t crc = m[b.csize - 8] | (m[b.csize - 7] << 8) | (m[b.csize - 6] << 16) | ((uint32_t)m[b.csize - 5] << 24);
                if (r != Z_STREAM_END || z.avail_out != 0 || z.avail_in != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), p + b.out, b.isize) != crc) {
                    bad[t] = 1;
                    break;
                }
            }
            inflateEnd(&z);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        for (int t = 0; t < nt; t++)
            if (bad[t]) return fail(-6, "gzip error in " + path_ + ": corrupt BGZF member");
        cused_ = at;
        delivered_ += out;
        if (out == 0 && cpos_ + at < size_ && !foreign) return fill_bgzf(p, cap);   // only empty members so far (EOF markers)
        return (long long)out;
    }

    std::string path_, err_;
    Mode mode_ = PLAIN;
    int fd_ = -1, threads_ = 16;
    int deferred_ = 0;
    uint64_t size_ = 0, pos_ = 0;
    gzFile zf_ = nullptr;
    bool gz_first_ = true;
    // BGZF: a window of the compressed file
    std::vector<uint8_t> cbuf_;
    uint64_t cpos_ = 0;          // file offset of cbuf_[0]
    size_t cused_ = 0;           // bytes of cbuf_ already turned into output
    uint64_t delivered_ = 0;     // uncompressed bytes handed out so far
    std::vector<Block> blocks_;
    // ordinary gzip, inflated in parallel
    const uint8_t *map_ = nullptr;
    std::unique_ptr<pgz::Reader> pgz_;
};

}  // namespace tdg
