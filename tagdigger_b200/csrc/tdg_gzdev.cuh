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
//               lengths at a time); positions that pass are queued and their headers parsed 32 at a
//               time, one per lane;
//   gz_decode   one LANE per chunk inflates from there (tdg_gzlane.h) into TOKENS (literals and
//               length / distance pairs), up to the block boundary at or behind the next chunk's
//               nominal offset, and reports where it started and stopped.  A lane is a state machine and the 32 lanes of a
//               warp take their steps together behind a vote, so they stay converged on the
//               symbol path however their blocks are cut.  Four warps per SM: a lane's hot tables take
//               1.7 KB of shared memory, interleaved with those of the other lanes of its warp;
//   (host)      the chain decides which chunks continue the stream (tdg_gzchain.h);
//   gz_expand   one WARP per accepted chunk turns its tokens into 16-bit symbols: a byte, or -- for
//               a reference into the unknown 32 KiB before the chunk -- the marker 256 + index;
//   gz_ptr_*    the 32 KiB in front of every accepted chunk, all chunks at once: every entry is a
//               byte or a pointer into the window before, and log2(chunks) passes of pointer
//               jumping leave bytes only;
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

#ifndef TDG_GZ_WARPS
#define TDG_GZ_WARPS 4
#endif
constexpr int DEC_THREADS = 32 * TDG_GZ_WARPS;       // lanes (= chunks) per CTA of gz_decode, one CTA per SM
constexpr size_t DEC_SMEM = (size_t)DEC_THREADS * gzl::LANE_U16 * 2;
constexpr int SCAN_WARPS = 8;
constexpr size_t SCAN_SMEM = (size_t)SCAN_WARPS * 32 * gzl::VAL_U16 * 2;     // 32 validator slices per warp
constexpr uint32_t MAXC = 1;                         // block starts the scan keeps per chunk (a second one would cost a second block's worth of scanning)
constexpr uint32_t PIECE = 16384;                    // bytes of text per CTA of gz_resolve
constexpr int RES_THREADS = 256;
constexpr uint32_t SUB = PIECE / RES_THREADS;        // bytes per thread in the CRC
constexpr int WIN_THREADS = 1024;
static_assert(DEC_SMEM <= 232448, "a CTA's lanes must fit the 227 KB of shared memory");
static_assert(SUB == 64, "the CRC tree's first operator is x^(8*64)");

struct RoundArgs {
    const uint32_t *in;          // compressed bytes of the round from the chunk grid on, zero padded
    uint64_t nwords, in_bits;
    uint32_t nchunks;
    uint64_t chunk_bytes, file_left;     // file_left: bytes from the grid to the end of the file
    uint64_t pos_rel;            // exact start of chunk 0, in bits from the grid
    uint64_t base_bit;           // bit position of the grid in the file
    uint32_t hist;
    uint32_t symcap, tokcap;     // symbols a chunk may produce; token slots per chunk in `syms`
    uint32_t *cand, *ncand;      // [nchunks][MAXC], [nchunks]
    uint16_t *syms;              // [nchunks][tokcap] token slots
    gzl::Meta *meta;
    const uint8_t *kraft3;
    uint16_t *cold;              // [nchunks][COLD_U16]: a lane's list of long-code symbols (the scan's warps use it first)
    // BGZF: every chunk is a whole member -- lane k inflates [m_start[k], m_end[k]) (bits from `in`) to its final block
    const uint64_t *m_start, *m_end;
};

__device__ __forceinline__ uint64_t nominal_rel(const RoundArgs &a, uint64_t k)
{
    const uint64_t b = k * a.chunk_bytes;
    return (b < a.file_left ? b : a.file_left) * 8;
}

__global__ void __launch_bounds__(SCAN_WARPS * 32) gz_scan(const RoundArgs a)
{
    extern __shared__ uint16_t s_lane[];                         // per warp: 32 interleaved validator slices
    __shared__ uint8_t s_kraft[512];
    __shared__ uint32_t s_queue[SCAN_WARPS][64];                 // positions that passed the cheap test, in order
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s_kraft[i] = a.kraft3[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x * SCAN_WARPS + warp + 1;       // chunk 0 starts at a known bit
    if (k >= a.nchunks) return;
    const uint64_t from = nominal_rel(a, k), to = nominal_rel(a, k + 1);
    const gzl::Mem<32> m{s_lane + (size_t)warp * gzl::VAL_U16 * 32 + lane * 2, nullptr};
    uint32_t *queue = s_queue[warp];
    uint32_t qn = 0, n = 0;
    // About one position in 300 passes the cheap test and a full look at one costs a lane ~2,000
    // instructions: the positions are queued and looked at 32 at a time, one per lane.
    for (uint64_t base = from; n < MAXC; base += 32) {
        const bool more = base < to;
        if (more) {
            const uint64_t idx = base >> 5;                      // `from` is a multiple of 128 bits
            const uint32_t w0 = a.in[idx], w1 = a.in[idx + 1], w2 = a.in[idx + 2], w3 = a.in[idx + 3];
            const uint64_t lo64 = (uint64_t)w1 << 32 | w0, hi64 = (uint64_t)w3 << 32 | w2;
            const uint64_t lo = lane ? (lo64 >> lane | hi64 << (64 - lane)) : lo64;
            const uint32_t hi = (uint32_t)(hi64 >> lane);
            const bool ok = base + lane < to && gzl::quick_test(lo, hi, s_kraft);
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, ok);
            if (ok) queue[qn + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(base + lane - from);
            qn += __popc(mask);
            __syncwarp();
        }
        if (qn >= 32 || (!more && qn)) {
            const uint32_t take = qn < 32 ? qn : 32;
            bool v = false;
            if (lane < take) v = gzl::header_check<32>(m, a.in, a.nwords, a.in_bits, from + queue[lane]);
            uint32_t good = __ballot_sync(0xFFFFFFFFu, v);
            while (good && n < MAXC) {
                const uint32_t l = (uint32_t)__ffs((int)good) - 1u;
                good &= good - 1u;
                if (lane == 0) a.cand[(size_t)k * MAXC + n] = queue[l];
                n++;
            }
            // drop the positions looked at
            const uint32_t rest = qn - take;
            const uint32_t moved = lane < rest ? queue[take + lane] : 0u;
            __syncwarp();
            if (lane < rest) queue[lane] = moved;
            qn = rest;
            __syncwarp();
        }
        if (!more && qn == 0) break;
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
        const gzl::Mem<32> m{s_lane + (size_t)warp * gzl::LANE_U16 * 32 + lane * 2, a.cold + (size_t)k * gzl::COLD_U16};
        const uint64_t stop = nominal_rel(a, k + 1);
        uint16_t *out = a.syms + (size_t)k * a.tokcap;
        if (a.m_start) z.init(m, a.in, a.nwords, a.m_end[k], true, a.m_start[k], nullptr, 0, ~0ull, 0, out, a.tokcap, a.symcap, &r);
        else if (k == 0) z.init(m, a.in, a.nwords, a.in_bits, true, a.pos_rel, nullptr, 0, stop, a.hist, out, a.tokcap, a.symcap, &r);
        else z.init(m, a.in, a.nwords, a.in_bits, false, nominal_rel(a, k), a.cand + (size_t)k * MAXC, a.ncand[k], stop, 0, out, a.tokcap, a.symcap, &r);
    }
    // every lane of the warp takes every step together: the vote is the point of reconvergence
    while (__any_sync(0xFFFFFFFFu, z.state != gzl::S_DONE)) z.step();
    if (mine) {
        r.start_bit += a.base_bit;
        r.end_bit += a.base_bit;
        a.meta[k] = r;
    }
}

// Tokens -> symbols, one WARP per accepted chunk.  32 token slots per step: a prefix sum of the
// lengths places every token; literals are stored at once; the step's matches are copied one after
// the other, each by the whole warp -- symbol k of a match is the symbol (k mod distance) of the
// distance symbols in front of it, all of which are written before the match begins, so the 32
// lanes never wait for each other inside a match (and overlapping matches need no special case).
struct ExpandArgs {
    const uint16_t *tok;         // [nchunks][tokcap]
    uint16_t *syms;              // [accepted][symcap]
    uint32_t tokcap, symcap;
    const uint32_t *ntok;        // [accepted] token slots of each accepted chunk (0: nothing to do)
    uint32_t accepted;
};
constexpr int EXP_WARPS = 8;

__global__ void __launch_bounds__(EXP_WARPS * 32) gz_expand(const ExpandArgs a)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t k = blockIdx.x * EXP_WARPS + (threadIdx.x >> 5);
    if (k >= a.accepted) return;
    const uint32_t ntok = a.ntok[k];
    const uint16_t *tok = a.tok + (size_t)k * a.tokcap;
    uint16_t *sym = a.syms + (size_t)k * a.symcap;
    uint32_t o = 0;
    for (uint32_t base = 0; base < ntok; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t s = i < ntok ? tok[i] : (uint32_t)gzl::T_DIST;           // beyond the end: a slot of no length
        const uint32_t nx = i + 1 < ntok ? tok[i + 1] : 0u;
        const bool lit = s < 256u, head = (s & 0xC000u) == gzl::T_LEN;
        const uint32_t len = lit ? 1u : (head ? (s & 0x1FFu) : 0u);
        uint32_t off = len;                                                     // inclusive prefix sum
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
            if ((int)lane >= d) off += v;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, off, 31);
        off = o + off - len;
        if (lit) sym[off] = (uint16_t)s;
        uint32_t mm = __ballot_sync(0xFFFFFFFFu, head);
        while (mm) {
            const int src = __ffs((int)mm) - 1;
            mm &= mm - 1u;
            const uint32_t mlen = __shfl_sync(0xFFFFFFFFu, len, src);
            const uint32_t dist = (__shfl_sync(0xFFFFFFFFu, nx, src) & 0x7FFFu) + 1u;
            const uint32_t dst = __shfl_sync(0xFFFFFFFFu, off, src);
            __syncwarp();                                                       // what the lanes stored so far is visible
            const int32_t from = (int32_t)dst - (int32_t)dist;
            for (uint32_t kk = lane; kk < mlen; kk += 32) {
                const int32_t j = from + (int32_t)(kk < dist ? kk : kk % dist);
                sym[dst + kk] = j < 0 ? (uint16_t)(256 + gzl::WIN + j) : sym[j];
            }
        }
        __syncwarp();
        o += total;
    }
}

// The 32 KiB in front of every accepted chunk.  R_0 is the window in front of the round (known
// bytes), R_q the window behind accepted chunk q - 1: its entry i is the chunk's symbol WIN - i
// from the end -- a byte, or a marker = entry s - 256 of R_{q-1} -- or, for a chunk shorter than
// WIN, entry i + len of R_{q-1}.  Headers copied from record to record make these chains as long
// as the round, so they are shortened by POINTER JUMPING, in place, all chunks at once: an entry
// is a byte (bit 31 set) or a pointer (q' << 15 | i'); a pass replaces every pointer by what it
// points to; ceil(log2(chunks)) + 1 passes leave bytes only.  (Handing the window from chunk to
// chunk in one CTA took 6.4 us per chunk: 57 ms for the 8,861 chunks of a 1.2 GB round.)
constexpr uint32_t P_BYTE = 0x80000000u;

struct WinArgs {
    const uint16_t *syms;
    uint32_t symcap;
    const uint32_t *out_len;     // [accepted]
    uint32_t accepted;
    const uint8_t *window_in;    // R_0
    uint32_t *ptrs;              // [accepted + 1][WIN]
    uint8_t *done;               // [accepted + 1] R_q holds bytes only
    uint8_t *window_out;         // the bytes of R_accepted (gz_ptr_take)
};

__global__ void __launch_bounds__(WIN_THREADS) gz_ptr_init(const WinArgs a)
{
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    uint32_t *dst = a.ptrs + (size_t)q * gzl::WIN;
    if (q == 0) {
        for (uint32_t i = tid; i < gzl::WIN; i += WIN_THREADS) dst[i] = P_BYTE | a.window_in[i];
        if (tid == 0) a.done[0] = 1;
        return;
    }
    const uint32_t len = a.out_len[q - 1];
    const uint16_t *src = a.syms + (size_t)(q - 1) * a.symcap;
    const uint32_t keep = len >= gzl::WIN ? 0u : gzl::WIN - len;         // entries that are the old window, moved up
    const uint32_t up = (q - 1) << 15;
    for (uint32_t i = tid; i < gzl::WIN; i += WIN_THREADS) {
        uint32_t e;
        if (i < keep) e = up | (i + len);
        else {
            const uint32_t s = src[len - (gzl::WIN - i)];
            e = s < 256 ? (P_BYTE | s) : (up | (s - 256));
        }
        dst[i] = e;
    }
    if (tid == 0) a.done[q] = 0;
}

__global__ void __launch_bounds__(WIN_THREADS) gz_ptr_jump(const WinArgs a)
{
    const uint32_t q = blockIdx.x + 1, tid = threadIdx.x;
    if (a.done[q]) return;
    uint32_t *mine = a.ptrs + (size_t)q * gzl::WIN;
    int left = 0;
    for (uint32_t i = tid; i < gzl::WIN; i += WIN_THREADS) {
        const uint32_t e = mine[i];
        if (e & P_BYTE) continue;
        // whatever the entry pointed to says now is true of this entry too (in-place passes may see
        // a target before or after its own update: both are right, the later one is shorter)
        const uint32_t t = a.ptrs[(size_t)(e >> 15) * gzl::WIN + (e & 0x7FFFu)];
        mine[i] = t;
        left |= !(t & P_BYTE);
    }
    left = __syncthreads_or(left);
    if (tid == 0 && !left) a.done[q] = 1;
}

__global__ void __launch_bounds__(WIN_THREADS) gz_ptr_take(const WinArgs a)
{
    const uint32_t *src = a.ptrs + (size_t)a.accepted * gzl::WIN;
    for (uint32_t i = blockIdx.x * WIN_THREADS + threadIdx.x; i < gzl::WIN; i += gridDim.x * WIN_THREADS) a.window_out[i] = (uint8_t)src[i];
}

struct ResArgs {
    const uint16_t *syms;
    uint32_t symcap;
    const uint64_t *text_off;    // [accepted + 1]
    uint32_t accepted;
    const uint32_t *ptrs;        // the windows in front of the accepted chunks (bytes by now)
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
        const uint8_t v = s < 256 ? (uint8_t)s : (uint8_t)a.ptrs[(size_t)c * gzl::WIN + (s - 256)];
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

// The raw CRCs of 256 consecutive FULL pieces become one (the host then folds one value per 4 MiB of
// text instead of one per 16 KiB).  A group that has fewer pieces (the last one) sits against the end
// of its frame: the empty slots in front are zero and stay zero under the shifts.
struct CrcFoldArgs {
    const uint32_t *piece;       // [nfull]
    uint32_t nfull;
    uint32_t *group;             // [ceil(nfull / 256)]
    uint32_t op[8];              // x^(8 * PIECE * 2^j) modulo the polynomial
};

__global__ void __launch_bounds__(256) gz_crc_fold(const CrcFoldArgs a)
{
    __shared__ uint32_t s_part[256];
    const uint32_t tid = threadIdx.x;
    const uint32_t first = blockIdx.x * 256u;
    const uint32_t cnt = a.nfull - first < 256u ? a.nfull - first : 256u;
    const uint32_t shift = 256u - cnt;
    s_part[tid] = tid >= shift ? a.piece[first + (tid - shift)] : 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t stride = 1u << j;
        if ((tid & (2 * stride - 1)) == 0) s_part[tid] = gf_mul(a.op[j], s_part[tid]) ^ s_part[tid + stride];
        __syncthreads();
    }
    if (tid == 0) a.group[blockIdx.x] = s_part[0];
}

// BGZF: the CRC-32 of every member's bytes (at most 64 KiB each), one warp per member.  Lane l takes
// the l-th 2 KiB of the member laid against the END of a 64 KiB frame (the lanes in front of a short
// member have nothing: a register that is still zero stays zero under the shifts); the lane that
// holds the member's first byte starts from 0xFFFFFFFF as the CRC asks.
struct MemberCrcArgs {
    const uint16_t *syms;        // [n][stride] the members' bytes as 16-bit symbols
    uint32_t stride;
    const uint32_t *lens;
    uint32_t n;
    uint32_t *crc;
    const uint32_t *table;
    uint32_t op[5];              // x^(8 * 2048 * 2^j) modulo the polynomial
};

__global__ void __launch_bounds__(256) gz_member_crc(const MemberCrcArgs a)
{
    __shared__ uint32_t s_tab[256];
    s_tab[threadIdx.x] = a.table[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= a.n) return;
    const uint32_t len = a.lens[k] < 65536u ? a.lens[k] : 65536u;
    const uint32_t shift = 65536u - len;
    const uint16_t *src = a.syms + (size_t)k * a.stride;
    uint32_t reg = 0;
    const uint32_t u0 = lane * 2048u;
    for (uint32_t j = 0; j < 2048u; j++) {
        const uint32_t u = u0 + j;
        if (u < shift) continue;
        const uint32_t i = u - shift;
        if (i == 0) reg = 0xFFFFFFFFu;
        reg = s_tab[(reg ^ src[i]) & 0xFFu] ^ (reg >> 8);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, reg, 1u << j);
        if ((lane & ((2u << j) - 1u)) == 0) reg = gf_mul(a.op[j], reg) ^ other;
    }
    if (lane == 0) a.crc[k] = len ? reg ^ 0xFFFFFFFFu : 0u;
}

}  // namespace gzd
}  // namespace tdg
