// Device side of the gzip feed: the kernels that inflate ONE ordinary gzip stream on the GPU.
//
// find_tags_fastq opens 'gz' files with gzip.open(f, 'rt') (/root/reference/tagdigger_fun.py:240-241).
// On the host that inflate is the wall of the whole path (tdg_pgz.h: 2.8 GB/s of text on sixteen
// threads); on the device the compressed bytes cross PCIe instead of the text, and the text is
// born in HBM where count_kernel reads it.  A round (tdg_gzchain.h) runs four kernels:
//
//   gz_scan     one WARP per chunk looks for the first dynamic block header behind the chunk's
//               nominal offset: 32 bit positions per step pass the cheap test (block bits, HLIT /
//               HDIST, code-length code exactly complete -- a table of Kraft sums for three
//               lengths at a time); a position that passes has its header parsed by its lane;
//   gz_decode   one LANE per chunk inflates from there (tdg_gzlane.h) into 16-bit symbols, up to
//               the block boundary at or behind the next chunk's nominal offset, and reports
//               where it started and stopped.  A lane is a state machine and the 32 lanes of a
//               warp take their steps together behind a vote, so they stay converged on the
//               symbol path however their blocks are cut.  Two warps per SM: a lane's tables take
//               3.5 KB of shared memory, interleaved with those of the other lanes of its warp;
//   (host)      the chain decides which chunks continue the stream (tdg_gzchain.h);
//   gz_windows  one CTA hands the last 32 KiB from accepted chunk to accepted chunk, resolving
//               markers against the window before (shared memory, 32 symbols per thread);
//   gz_resolve  every accepted symbol becomes its byte at its place in the round's text; the CTA
//               of a 16 KiB piece also takes the piece's raw CRC-32 (64 bytes per thread, then a
//               tree of carry-less multiplications by x^(8 * 64 * 2^j)) and notes whether any byte
//               has its high bit set (text mode: such a file needs UTF-8 validation).
// The host folds the piece CRCs into the member's CRC-32 and checks it against the trailer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tdg_gzlane.h"

namespace tdg {
namespace gzd {

constexpr int DEC_THREADS = 64;                      // lanes (= chunks) per CTA of gz_decode: two warps
constexpr size_t DEC_SMEM = (size_t)DEC_THREADS * gzl::LANE_U16 * 2;
constexpr int SCAN_WARPS = 8;
constexpr size_t SCAN_SMEM = (size_t)SCAN_WARPS * gzl::LANE_U16 * 2;
constexpr uint32_t MAXC = 2;                         // block starts the scan keeps per chunk
constexpr uint32_t PIECE = 16384;                    // bytes of text per CTA of gz_resolve
constexpr int RES_THREADS = 256;
constexpr uint32_t SUB = PIECE / RES_THREADS;        // bytes per thread in the CRC
constexpr int WIN_THREADS = 1024;
static_assert(SUB == 64, "the CRC tree's first operator is x^(8*64)");

struct RoundArgs {
    const uint32_t *in;          // compressed bytes of the round from the chunk grid on, zero padded
    uint64_t nwords, in_bits;
    uint32_t nchunks;
    uint64_t chunk_bytes, file_left;     // file_left: bytes from the grid to the end of the file
    uint64_t pos_rel;            // exact start of chunk 0, in bits from the grid
    uint64_t base_bit;           // bit position of the grid in the file
    uint32_t hist;
    uint32_t symcap;
    uint32_t *cand, *ncand;      // [nchunks][MAXC], [nchunks]
    uint16_t *syms;              // [nchunks][symcap]
    gzl::Meta *meta;
    const uint8_t *kraft3;
};

__device__ __forceinline__ uint64_t nominal_rel(const RoundArgs &a, uint64_t k)
{
    const uint64_t b = k * a.chunk_bytes;
    return (b < a.file_left ? b : a.file_left) * 8;
}

__global__ void __launch_bounds__(SCAN_WARPS * 32) gz_scan(const RoundArgs a)
{
    extern __shared__ uint16_t s_lane[];
    __shared__ uint8_t s_kraft[512];
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s_kraft[i] = a.kraft3[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x * SCAN_WARPS + warp + 1;       // chunk 0 starts at a known bit
    if (k >= a.nchunks) return;
    const uint64_t from = nominal_rel(a, k), to = nominal_rel(a, k + 1);
    const gzl::Mem<1> m{s_lane + warp * gzl::LANE_U16};
    uint32_t n = 0;
    for (uint64_t base = from; base < to && n < MAXC; base += 32) {
        const uint64_t idx = base >> 5;                          // `from` is a multiple of 128 bits
        const uint32_t w0 = a.in[idx], w1 = a.in[idx + 1], w2 = a.in[idx + 2], w3 = a.in[idx + 3];
        const uint64_t lo64 = (uint64_t)w1 << 32 | w0, hi64 = (uint64_t)w3 << 32 | w2;
        const uint64_t lo = lane ? (lo64 >> lane | hi64 << (64 - lane)) : lo64;
        const uint32_t hi = (uint32_t)(hi64 >> lane);
        const uint64_t p = base + lane;
        const bool ok = p < to && gzl::quick_test(lo, hi, s_kraft);
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, ok);
        while (mask && n < MAXC) {
            const uint32_t l = (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            bool v = false;
            if (lane == l) v = gzl::header_parses<1>(m, a.in, a.nwords, a.in_bits, p);
            if (__ballot_sync(0xFFFFFFFFu, v)) {
                if (lane == 0) a.cand[(size_t)k * MAXC + n] = (uint32_t)(base + l - from);
                n++;
            }
        }
    }
    if (lane == 0) a.ncand[k] = n;
}

__global__ void __launch_bounds__(DEC_THREADS) gz_decode(const RoundArgs a)
{
    extern __shared__ uint16_t s_lane[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x * DEC_THREADS + threadIdx.x;
    const bool mine = k < a.nchunks;
    gzl::Lane<32> z;
    gzl::Meta r;
    z.state = gzl::S_DONE;
    if (mine) {
        const gzl::Mem<32> m{s_lane + (size_t)warp * gzl::LANE_U16 * 32 + lane * 2};
        const uint64_t stop = nominal_rel(a, k + 1);
        uint16_t *out = a.syms + (size_t)k * a.symcap;
        if (k == 0) z.init(m, a.in, a.nwords, a.in_bits, true, a.pos_rel, nullptr, 0, stop, a.hist, out, a.symcap, &r);
        else z.init(m, a.in, a.nwords, a.in_bits, false, nominal_rel(a, k), a.cand + (size_t)k * MAXC, a.ncand[k], stop, 0, out, a.symcap, &r);
    }
    // every lane of the warp takes every step together: the vote is the point of reconvergence
    while (__any_sync(0xFFFFFFFFu, z.state != gzl::S_DONE)) z.step();
    if (mine) {
        r.start_bit += a.base_bit;
        r.end_bit += a.base_bit;
        a.meta[k] = r;
    }
}

// windows[0] holds the 32 KiB in front of the round; windows[k + 1] = the 32 KiB behind accepted chunk k
struct WinArgs {
    const uint16_t *syms;
    uint32_t symcap;
    const uint32_t *out_len;     // [accepted]
    uint32_t accepted;
    uint8_t *windows;            // [accepted + 1][WIN]
};

__global__ void __launch_bounds__(WIN_THREADS) gz_windows(const WinArgs a)
{
    extern __shared__ uint8_t s_w[];                 // two windows
    constexpr uint32_t PER = gzl::WIN / WIN_THREADS; // 32 entries per thread
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < gzl::WIN; i += WIN_THREADS) s_w[i] = a.windows[i];
    // the last WIN symbols of a chunk do not depend on the chain: the next chunk's are loaded
    // while this chunk's are resolved (a step is then shared-memory work only)
    uint32_t cur_s[PER / 2], nxt_s[PER / 2];          // two symbols per register
    auto load = [&](uint32_t k, uint32_t *dst) {
        const uint32_t len = a.out_len[k];
        const uint16_t *src = a.syms + (size_t)k * a.symcap;
        const uint32_t keep = len >= gzl::WIN ? 0u : gzl::WIN - len;
#pragma unroll
        for (uint32_t j = 0; j < PER; j += 2) {
            const uint32_t i0 = tid + j * WIN_THREADS, i1 = i0 + WIN_THREADS;
            const uint32_t s0 = i0 < keep ? 0u : src[len - (gzl::WIN - i0)];
            const uint32_t s1 = i1 < keep ? 0u : src[len - (gzl::WIN - i1)];
            dst[j / 2] = s0 | s1 << 16;
        }
    };
    if (a.accepted) load(0, cur_s);
    __syncthreads();
    for (uint32_t k = 0; k < a.accepted; k++) {
        const uint8_t *cur = s_w + (k & 1u) * gzl::WIN;
        uint8_t *nxt = s_w + ((k & 1u) ^ 1u) * gzl::WIN;
        const uint32_t len = a.out_len[k];
        uint8_t *dst = a.windows + (size_t)(k + 1) * gzl::WIN;
        const uint32_t keep = len >= gzl::WIN ? 0u : gzl::WIN - len;     // bytes of the old window that stay
        if (k + 1 < a.accepted) load(k + 1, nxt_s);
#pragma unroll
        for (uint32_t j = 0; j < PER; j++) {
            const uint32_t i = tid + j * WIN_THREADS;
            const uint32_t s = (j & 1u) ? cur_s[j / 2] >> 16 : cur_s[j / 2] & 0xFFFFu;
            const uint8_t v = i < keep ? cur[i + len] : (s < 256 ? (uint8_t)s : cur[s - 256]);
            nxt[i] = v;
            dst[i] = v;
        }
#pragma unroll
        for (uint32_t j = 0; j < PER / 2; j++) cur_s[j] = nxt_s[j];
        __syncthreads();
    }
}

struct ResArgs {
    const uint16_t *syms;
    uint32_t symcap;
    const uint64_t *text_off;    // [accepted + 1]
    uint32_t accepted;
    const uint8_t *windows;
    uint8_t *text;               // where byte 0 of the round's text goes
    uint64_t text_len;
    uint32_t *crc;               // [pieces] raw CRC-32 (register starts at 0, no final inversion)
    uint32_t *flag;              // OR of all bytes
    const uint32_t *table;       // the 256-entry table of the reflected CRC-32 polynomial
    uint32_t op[8];              // x^(8 * 64 * 2^j) modulo the polynomial, j = 0..7 (zlib's crc32_combine_gen)
};

// a(x) * b(x) modulo the CRC-32 polynomial, reflected representation (bit 31 = x^0)
__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 4
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}

__global__ void __launch_bounds__(RES_THREADS) gz_resolve(const ResArgs a)
{
    __shared__ uint8_t s_txt[PIECE + PIECE / 64 * 4];        // 64-byte rows padded to 68: the CRC's word reads spread over the banks
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_part[RES_THREADS];
    __shared__ uint32_t s_first;
    const uint32_t tid = threadIdx.x;
    const uint64_t g0 = (uint64_t)blockIdx.x * PIECE;
    const uint32_t n = (uint32_t)(a.text_len - g0 < PIECE ? a.text_len - g0 : PIECE);
    const uint32_t shift = PIECE - n;                        // a short piece sits at the END of its rows: leading zeros leave a raw CRC alone
    s_tab[tid] = a.table[tid];
    if (tid == 0) {
        uint32_t lo = 0, hi = a.accepted;                    // last chunk whose text starts at or before g0
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) / 2;
            if (a.text_off[mid] <= g0) lo = mid; else hi = mid;
        }
        s_first = lo;
    }
    __syncthreads();
    uint32_t c = s_first;
    uint64_t c_end = a.text_off[c + 1];
    uint32_t any = 0;
    for (uint32_t i = tid; i < n; i += RES_THREADS) {
        const uint64_t g = g0 + i;
        while (g >= c_end) c_end = a.text_off[++c + 1];
        const uint16_t s = a.syms[(size_t)c * a.symcap + (g - a.text_off[c])];
        const uint8_t v = s < 256 ? (uint8_t)s : a.windows[(size_t)c * gzl::WIN + (s - 256)];
        a.text[g] = v;
        const uint32_t u = i + shift;
        s_txt[u + (u >> 6) * 4] = v;
        any |= v;
    }
    __syncthreads();
    uint32_t reg = 0;
    {
        const uint32_t u0 = tid * SUB;
        const uint8_t *row = s_txt + u0 + tid * 4;
        for (uint32_t j = 0; j < SUB; j++)
            if (u0 + j >= shift) reg = s_tab[(reg ^ row[j]) & 0xFFu] ^ (reg >> 8);
    }
    s_part[tid] = reg;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t stride = 1u << j;
        if ((tid & (2 * stride - 1)) == 0) s_part[tid] = gf_mul(a.op[j], s_part[tid]) ^ s_part[tid + stride];
        __syncthreads();
    }
    if (tid == 0) a.crc[blockIdx.x] = s_part[0];
    any = __reduce_or_sync(0xFFFFFFFFu, any);
    if ((tid & 31u) == 0 && (any & 0x80u)) atomicOr(a.flag, 0x80u);
}

}  // namespace gzd
}  // namespace tdg
