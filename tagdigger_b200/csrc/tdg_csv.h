// Multithreaded CSV output of the count matrix and of the diploid genotype table
// (host code; no CUDA).  Replaces the per-cell Python work of writeCounts and
// writeDiploidGeno (/root/reference/tagdigger_fun.py:1100-1111, :1144-1180) for matrices
// that arrive as int32 arrays: config 4's 384 x 500,000 matrix is 192 M cells of text.
// The bytes are those of Python's csv.writer (default dialect): the caller passes the
// header line and the already CSV-escaped row labels; cells are plain decimal integers
// (or 0/1/2/empty genotype calls), fields separated by ',', rows ended by "\r\n".
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace tdg {

inline char *put_int(char *p, int32_t v)
{
    char tmp[12];
    int n = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// rows [r0, r1) of the table into out (appended)
template <class Cell>
inline void format_rows(std::string &out, uint32_t r0, uint32_t r1, uint32_t ncell, const char *labels,
                        const uint64_t *label_off, Cell cell)
{
    std::vector<char> line((size_t)ncell * 12 + 16);
    for (uint32_t r = r0; r < r1; r++) {
        out.append(labels + label_off[r], (size_t)(label_off[r + 1] - label_off[r]));
        char *p = line.data();
        for (uint32_t c = 0; c < ncell; c++) {
            *p++ = ',';
            p = cell(p, r, c);
        }
        *p++ = '\r';
        *p++ = '\n';
        out.append(line.data(), (size_t)(p - line.data()));
    }
}

// Writes header + rows to path.  Returns "" or an error message.
template <class Cell>
inline std::string write_table(const char *path, const char *header, size_t header_len, uint32_t rows, uint32_t ncell,
                               const char *labels, const uint64_t *label_off, int threads, Cell cell)
{
    FILE *fp = fopen(path, "wb");
    if (!fp) return std::string("cannot open ") + path + " for writing";
    bool ok = fwrite(header, 1, header_len, fp) == header_len;
    if (threads < 1) threads = 1;
    // blocks of rows sized so that a round of blocks stays within ~256 MiB of text
    size_t per_row = (size_t)ncell * 4 + 64;
    uint32_t block = (uint32_t)std::max<size_t>(1, ((size_t)32 << 20) / per_row);
    for (uint32_t base = 0; base < rows && ok; base += block * (uint32_t)threads) {
        int nt = 0;
        std::vector<std::string> bufs((size_t)threads);
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) {
            uint64_t r0 = (uint64_t)base + (uint64_t)t * block;
            if (r0 >= rows) break;
            uint32_t r1 = (uint32_t)std::min<uint64_t>(rows, r0 + block);
            nt++;
            th.emplace_back([&, t, r0, r1]() { format_rows(bufs[(size_t)t], (uint32_t)r0, r1, ncell, labels, label_off, cell); });
        }
        for (auto &x : th) x.join();
        for (int t = 0; t < nt && ok; t++) ok = fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), fp) == bufs[(size_t)t].size();
    }
    if (fclose(fp) != 0) ok = false;
    return ok ? "" : std::string("write error on ") + path;
}

}  // namespace tdg
