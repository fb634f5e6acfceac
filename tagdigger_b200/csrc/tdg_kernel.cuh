// The fused counting kernel: FASTQ line splitting, whitespace strip / case fold,
// 2-bit packing, barcode+cutsite lookup, tag probe and count update in ONE pass
// over the raw bytes (sm_100a).
//
// Replaces the loop of find_tags_fastq, /root/reference/tagdigger_fun.py:250-274:
//   for line in fqcon:                      -> line ends found 64 bytes per thread (SWAR)
//       if lineindex % 4 == 1:              -> see "line numbering" below
//           line1 = line.strip().upper()    -> leading-whitespace skip + case fold
//           sequence_index_lookup(x2)       -> tdg_match.h (packed exact-match probes)
//           mycounts[bar][tag] += 1         -> warp-aggregated red.global.add.s32
//
// Data movement: a persistent grid; each CTA draws SEGMENTS (runs of consecutive
// tiles) from a global ticket counter and streams their tiles through a ring of
// shared-memory buffers filled by the TMA unit (cp.async.bulk + mbarrier
// complete_tx).  Every byte of the stream is read from HBM exactly once.
//
// Line numbering.  Which lines are sequence lines is decided by the GLOBAL line
// index (lineindex % 4 == 1 counted from the start of the file), which a CTA
// that starts in the middle of the file cannot know without every byte before
// it.  The kernel therefore runs speculatively and verifies:
//   1. count pass (this kernel, mode MAIN): each segment numbers its lines from a
//      GUESS of its first line's index mod 4, read off the FASTQ structure of its
//      first lines ('@' line, '+' two lines later, equal sequence/quality length),
//      and records how many lines it saw.  Segment 0 knows its true index.
//   2. verify_kernel: a prefix sum over the per-segment line counts gives every
//      segment's true first line index; segments whose guess was wrong (or that
//      reach past the read limit) go on a fix list.  For real FASTQ it is empty.
//   3. fix pass (this kernel, mode FIX): listed segments are counted again with
//      the guessed numbering and weight -1, then with the true numbering, the
//      read limit and weight +1.
// The result is exact for ANY byte stream; only the speed depends on the guess.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdg_match.h"

namespace tdg {

#ifndef TDG_THREADS
#define TDG_THREADS 128
#endif
#ifndef TDG_STAGES
#define TDG_STAGES 3
#endif

constexpr int      THREADS = TDG_THREADS;
constexpr int      WARPS = THREADS / 32;
constexpr uint32_t SPAN = 64;                      // bytes per thread in the line scan
constexpr uint32_t TILE = THREADS * SPAN;          // bytes per tile
constexpr uint32_t HALO = 512;                     // == TDG_HALO_BYTES
constexpr uint32_t STAGE_BYTES = TILE + HALO;
constexpr uint32_t STAGE_STRIDE = STAGE_BYTES + 128;   // slack for unaligned word reads
constexpr int      STAGES = TDG_STAGES;
constexpr uint32_t STARTS_CAP = TILE / 8;          // line starts kept per emission window
constexpr uint32_t BAR_SMEM_MAX = 16384;           // barcode tables up to this size are copied to smem
constexpr uint32_t GUESS_LINES = 16;               // lines inspected for the FASTQ structure guess

enum { PREV_NONE = 0, PREV_LF = 1, PREV_CR = 2, PREV_OTHER = 3 };
enum { MODE_MAIN = 0, MODE_FIX = 1 };

struct LineState {            // where a chunk starts in its file
    unsigned long long next_line;   // index the next line START will receive
    uint32_t prev_kind;             // PREV_*
    uint32_t pad;
};

struct SegInfo {              // written by the count pass, one per segment
    uint32_t lines;           // line starts numbered by this segment
    uint32_t guess;           // assumed (index of its first line start) mod 4
};

struct FixEntry {
    uint32_t seg;
    uint32_t guess;
    unsigned long long true_first;
};

struct ChunkArgs {
    const uint8_t *bytes;           // 16-byte aligned; allocation >= round_up(n, TILE) + HALO
    unsigned long long n;
    uint32_t num_tiles;
    uint32_t seg_tiles;             // tiles per segment
    uint32_t num_segs;
    uint32_t mode;                  // MODE_*
    uint32_t use_arg_state;         // 1: (line_base, prev_kind) below; 0: *state_in
    uint32_t prev_kind;
    unsigned long long line_base;
    const LineState *state_in;
    unsigned long long *ticket;     // zeroed before every launch
    SegInfo *seginfo;               // [num_segs]
    uint32_t *last_kind;            // PREV_* of the last byte of the chunk
    const FixEntry *fix;            // MODE_FIX: the list, and
    const uint32_t *n_fix;          //           its length
    unsigned long long reads_limit;
    const BarTable *bar;
    uint32_t bar_bytes;             // header + entries
    uint32_t cols;
    TagTable tags;
    int32_t *matrix;
    unsigned long long *totals;     // [3]: reads, barcode+cutsite hits, tag hits
};

struct VerifyArgs {
    uint32_t num_segs;
    uint32_t use_arg_state;
    uint32_t prev_kind;
    uint32_t make_fixes;
    unsigned long long line_base;
    unsigned long long reads_limit;
    const LineState *state_in;
    LineState *state_out;
    const SegInfo *seginfo;
    const uint32_t *last_kind;
    FixEntry *fix;
    uint32_t *n_fix;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 16 bytes -> 16-bit mask of bytes < 0x20 (every byte must be < 0x80).
// Per word: bit 7 of (w + 0x60) is clear iff the byte is < 0x20; the multiply
// gathers the four flags into the top nibble (no partial products collide).
__device__ __forceinline__ uint32_t ctl_mask16(uint4 q)
{
    uint32_t acc = 0, f;
    f = ~(q.w + 0x60606060u) & 0x80808080u; acc = __funnelshift_l(f * 0x00204081u, acc, 4);
    f = ~(q.z + 0x60606060u) & 0x80808080u; acc = __funnelshift_l(f * 0x00204081u, acc, 4);
    f = ~(q.y + 0x60606060u) & 0x80808080u; acc = __funnelshift_l(f * 0x00204081u, acc, 4);
    f = ~(q.x + 0x60606060u) & 0x80808080u; acc = __funnelshift_l(f * 0x00204081u, acc, 4);
    return acc;
}

// Unaligned 32-character window out of a shared-memory stage buffer.
struct SmemFetch {
    const uint8_t *p;     // first character of the stripped line (shared memory)
    uint32_t limit;
    __device__ __forceinline__ void load8(uint32_t off, uint32_t w[8]) const
    {
        const uint8_t *q = p + off;
        uint32_t sh = ((uint32_t)(uintptr_t)q & 3u) * 8u;
        const uint32_t *a = (const uint32_t *)((uintptr_t)q & ~(uintptr_t)3);
        uint32_t prev = a[0];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t nxt = a[i + 1];
            w[i] = __funnelshift_r(prev, nxt, sh);
            prev = nxt;
        }
    }
};

// The same window read byte by byte from global memory (lines that run past
// the staged halo: very long leading whitespace or very long tags).
struct GlobalFetch {
    const uint8_t *p;
    uint32_t limit;
    __device__ __forceinline__ void load8(uint32_t off, uint32_t w[8]) const
    {
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            uint32_t v = 0;
            for (int k = 0; k < 4; k++) {
                uint32_t o = off + 4 * i + k;
                uint32_t c = o < limit ? p[o] : 0u;
                v |= c << (8 * k);
            }
            w[i] = v;
        }
    }
};

template <bool MATCH>
__global__ void __launch_bounds__(THREADS) count_kernel(const ChunkArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *stage_base = smem;
    uint16_t *starts = (uint16_t *)(smem + STAGES * STAGE_STRIDE);
    uint8_t *bar_smem = (uint8_t *)(starts + STARTS_CAP);

    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ uint32_t s_item[STAGES];            // work item of the tile in each stage (or NONE)
    __shared__ uint32_t s_tix[STAGES];             // tile index inside its segment
    __shared__ uint32_t s_warp_cnt[WARPS];
    __shared__ uint32_t s_guess;
    __shared__ unsigned long long s_tot[2];

    constexpr uint32_t NONE = 0xFFFFFFFFu;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    // ---- chunk-level state -------------------------------------------------
    unsigned long long line_base;
    uint32_t prev_kind;
    if (a.use_arg_state) {
        line_base = a.line_base;
        prev_kind = a.prev_kind;
    } else {
        line_base = a.state_in->next_line;
        prev_kind = a.state_in->prev_kind;
    }
    const uint32_t n_items = a.mode == MODE_FIX ? 2u * *a.n_fix : a.num_segs;
    if (n_items == 0) return;

    const BarTable *bar = a.bar;
    if (MATCH) {
        if (a.bar_bytes <= BAR_SMEM_MAX) {
            const uint4 *src = (const uint4 *)a.bar;
            uint4 *dst = (uint4 *)bar_smem;
            for (uint32_t i = tid; i < a.bar_bytes / 16; i += THREADS) dst[i] = src[i];
            bar = (const BarTable *)bar_smem;
        }
    }
    const BarEntry *bent = (const BarEntry *)((const uint8_t *)bar + sizeof(BarTable));

    // ---- producer state (thread 0): the next tile to request ------------------
    uint32_t p_item = NONE, p_seg = 0, p_tix = 0, p_ntiles = 0;
    auto produce = [&](int s) {
        // called by thread 0 only: pick the next tile and start its copy into stage s
        if (p_tix == p_ntiles) {
            unsigned long long t = atomicAdd(a.ticket, 1ull);
            if (t < n_items) {
                p_item = (uint32_t)t;
                p_seg = a.mode == MODE_FIX ? a.fix[p_item >> 1].seg : p_item;
                uint32_t first_tile = p_seg * a.seg_tiles;
                uint32_t left = a.num_tiles - first_tile;
                p_ntiles = left < a.seg_tiles ? left : a.seg_tiles;
                p_tix = 0;
            } else {
                p_item = NONE;
                p_ntiles = 0;
                p_tix = 0;
            }
        }
        s_item[s] = p_item;
        s_tix[s] = p_tix;
        if (p_item != NONE) {
            size_t tile = (size_t)p_seg * a.seg_tiles + p_tix;
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            bulk_g2s(stage_base + s * STAGE_STRIDE, a.bytes + tile * TILE, STAGE_BYTES, &full_bar[s]);
            p_tix++;
        }
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
        s_tot[0] = s_tot[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < STAGES; s++) produce(s);
    }
    __syncthreads();

    // bytes a match may touch past the stripped line start (fast path bound)
    uint32_t need = 0;
    if (MATCH) {
        need = bar->max_tag_off + a.tags.max_len + 36u;
        if (bar->max_len + 36u > need) need = bar->max_len + 36u;
    }
    long long my_reads = 0;               // thread 0 only
    int32_t my_bar = 0, my_tag = 0;

    // ---- per-item state (uniform across the CTA) -------------------------------
    uint32_t seg = 0, seg_ntiles = 0, seg_lines = 0, guess = 0;
    unsigned long long seg_first = 0;     // (assumed) index of the segment's first line start
    unsigned long long limit = ~0ull;
    int32_t weight = 1;
    bool need_guess = false;

    for (uint32_t it = 0;; it++) {
        const uint32_t s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        const uint32_t item = s_item[s];
        if (item == NONE) break;
        const uint32_t tix = s_tix[s];
        mbar_wait(&full_bar[s], parity);

        if (tix == 0) {                    // a new segment starts
            seg_lines = 0;
            limit = ~0ull;
            weight = 1;
            need_guess = false;
            if (a.mode == MODE_FIX) {
                FixEntry fe = a.fix[item >> 1];
                seg = fe.seg;
                if (item & 1u) { seg_first = fe.true_first; limit = a.reads_limit; }
                else           { seg_first = fe.guess; weight = -1; }
            } else {
                seg = item;
                if (seg == 0) seg_first = line_base;       // known exactly
                else { seg_first = 0; need_guess = true; }
            }
            guess = (uint32_t)(seg_first & 3ull);
            uint32_t first_tile = seg * a.seg_tiles;
            uint32_t left = a.num_tiles - first_tile;
            seg_ntiles = left < a.seg_tiles ? left : a.seg_tiles;
        }
        const uint32_t t = seg * a.seg_tiles + tix;         // tile index in the chunk

        const uint8_t *buf = stage_base + s * STAGE_STRIDE;
        const unsigned long long tile_off = (unsigned long long)t * TILE;
        const unsigned long long avail = a.n - tile_off;            // bytes from tile start to chunk end
        const uint32_t valid = avail < TILE ? (uint32_t)avail : TILE;

        // An implicit line end just before byte 0 of the chunk?
        uint32_t extra = 0;
        if (t == 0) {
            if (prev_kind == PREV_NONE || prev_kind == PREV_LF) extra = 1;
            else if (prev_kind == PREV_CR && buf[0] != '\n') extra = 1;
        }

        // ---- phase A: line-end mask of my 64 bytes --------------------------
        uint32_t mlo = 0, mhi = 0;          // bit i: a line ends at byte tid*64 + i
        uint32_t hi_or = 0;
        {
            const uint4 *src = (const uint4 *)(buf + tid * SPAN);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t q = (i + (lane >> 1)) & 3u;      // conflict-free piece order
                uint4 v = src[q];
                hi_or |= v.x | v.y | v.z | v.w;
                uint32_t m16 = ctl_mask16(v) << ((q & 1u) * 16u);
                if (q & 2u) mhi |= m16; else mlo |= m16;
            }
        }
        // every candidate must really be '\n' and every byte ASCII, otherwise
        // the whole tile is redone exactly below
        uint32_t bad = hi_or & 0x80808080u;
        {
            uint32_t m = mlo;
            while (m) { uint32_t b = __ffs(m) - 1; m &= m - 1; bad |= (buf[tid * SPAN + b] != '\n'); }
            m = mhi;
            while (m) { uint32_t b = __ffs(m) - 1; m &= m - 1; bad |= (buf[tid * SPAN + 32 + b] != '\n'); }
        }
        const bool last_tile = t == a.num_tiles - 1;
        auto clip_last = [&]() {
            // The line that would start right after the last byte of the chunk
            // is numbered by the NEXT chunk (PREV_LF), and bytes at and after n
            // do not exist: keep line ends at p < valid - 1 only.
            uint32_t lim = valid - 1;
            uint32_t first = tid * SPAN;
            if (first + 64 > lim) {
                uint32_t keep = lim > first ? lim - first : 0;          // 0..63
                uint64_t km = (1ull << keep) - 1ull;
                mlo &= (uint32_t)km;
                mhi &= (uint32_t)(km >> 32);
            }
        };
        if (last_tile) clip_last();
        uint32_t cnt = __popc(mlo) + __popc(mhi);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp_cnt[warp] = incl;
        if (__syncthreads_or(bad)) {
            // exact path: '\n' ends a line; '\r' ends one unless a '\n' follows
            // (Python universal newlines); anything else is content.
            mlo = mhi = 0;
            for (uint32_t i = 0; i < SPAN; i++) {
                uint32_t p = tid * SPAN + i;
                uint32_t c = buf[p];
                bool end = c == '\n';
                if (c == '\r') {
                    // the byte after the last byte of the chunk is unknown: pending
                    end = (p + 1 < avail) && buf[p + 1] != '\n';
                }
                if (end) { if (i < 32) mlo |= 1u << i; else mhi |= 1u << (i - 32); }
            }
            if (last_tile) clip_last();
            cnt = __popc(mlo) + __popc(mhi);
            incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            __syncthreads();                 // everyone has read the fast-path counts
            if (lane == 31) s_warp_cnt[warp] = incl;
            __syncthreads();
        }
        uint32_t before = extra, total = extra;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t c = s_warp_cnt[w];
            if (w < (int)warp) before += c;
            total += c;
        }
        const uint32_t my_rank0 = before + incl - cnt;      // rank of my first line end in the tile

        if (last_tile && tid == 0) {
            uint32_t c = buf[valid - 1];
            *a.last_kind = c == '\n' ? PREV_LF : (c == '\r' ? PREV_CR : PREV_OTHER);
        }

        if (MATCH) {
            // ---- emission + matching, STARTS_CAP line starts at a time --------
            for (uint32_t wlo = 0; wlo < total; wlo += STARTS_CAP) {
                {
                    uint32_t rank = my_rank0;
                    uint32_t m = mlo;
                    while (m) {
                        uint32_t b = __ffs(m) - 1; m &= m - 1;
                        uint32_t k = rank - wlo;
                        if (k < STARTS_CAP) starts[k] = (uint16_t)(tid * SPAN + b + 1);
                        rank++;
                    }
                    m = mhi;
                    while (m) {
                        uint32_t b = __ffs(m) - 1; m &= m - 1;
                        uint32_t k = rank - wlo;
                        if (k < STARTS_CAP) starts[k] = (uint16_t)(tid * SPAN + 32 + b + 1);
                        rank++;
                    }
                    if (tid == 0 && extra && wlo == 0) starts[0] = 0;
                }
                __syncthreads();             // starts[] is ready

                if (need_guess) {
                    // First lines of a segment whose position in the file is not
                    // known yet: find a line that looks like a FASTQ header
                    // ('@', then '+' two lines on, sequence and quality lines of
                    // equal length).  Any answer is acceptable -- a wrong one is
                    // found and repaired by verify_kernel + the fix pass.
                    if (tid == 0) {
                        uint32_t g = 0;
                        uint32_t have = total < STARTS_CAP ? total : STARTS_CAP;
                        for (uint32_t k = 0; k + 4 < have && k < GUESS_LINES; k++) {
                            uint32_t p0 = starts[k], p1 = starts[k + 1], p2 = starts[k + 2], p3 = starts[k + 3],
                                     p4 = starts[k + 4];
                            if (buf[p0] == '@' && buf[p2] == '+' && p2 - p1 == p4 - p3) {
                                g = (4u - (k & 3u)) & 3u;          // line k has index 0 mod 4
                                break;
                            }
                        }
                        s_guess = g;
                    }
                    __syncthreads();
                    guess = s_guess;
                    seg_first = guess;
                    need_guess = false;
                }

                const unsigned long long first_line = seg_first + seg_lines;   // index of rank 0 of this tile
                // sequence lines: index % 4 == 1
                uint32_t r0 = (uint32_t)((1ull - first_line) & 3ull);
                uint32_t whi = wlo + STARTS_CAP < total ? wlo + STARTS_CAP : total;
                // first r >= wlo with r % 4 == r0 % 4
                uint32_t rbeg = wlo + ((r0 - wlo) & 3u);
                for (uint32_t r = rbeg + 4 * tid; r < whi; r += 4 * THREADS) {
                    unsigned long long read_idx = (first_line + r) >> 2;
                    bool live = read_idx < limit;
                    uint32_t pos = starts[r - wlo];       // always < avail (see clip_last)
                    int32_t row = -1, col = -1;
                    if (live) {
                        // leading whitespace (str.strip)
                        const uint32_t staged = avail < STAGE_BYTES ? (uint32_t)avail : STAGE_BYTES;
                        while (pos < staged) {
                            uint32_t c = buf[pos];
                            if (is_lead_space(c)) { pos++; continue; }
                            if (c >= 0xC2 && c <= 0xE3 && pos + 2 < staged) {
                                uint32_t u = utf8_space(c, buf[pos + 1], buf[pos + 2]);
                                if (u) { pos += u; continue; }
                            }
                            break;
                        }
                        MatchResult mr;
                        if (pos + need <= STAGE_BYTES) {
                            SmemFetch f;
                            f.p = buf + pos;
                            unsigned long long room = avail - pos;
                            f.limit = room > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)room;
                            mr = match_line(f, bar, bent, a.tags);
                        } else {
                            // rare: continue from global memory
                            const uint8_t *g = a.bytes + tile_off;
                            unsigned long long gp = pos;
                            while (gp < avail) {
                                uint32_t c = g[gp];
                                if (is_lead_space(c)) { gp++; continue; }
                                if (c >= 0xC2 && c <= 0xE3 && gp + 2 < avail) {
                                    uint32_t u = utf8_space(c, g[gp + 1], g[gp + 2]);
                                    if (u) { gp += u; continue; }
                                }
                                break;
                            }
                            GlobalFetch f;
                            f.p = g + gp;
                            unsigned long long room = avail - gp;
                            f.limit = room > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)room;
                            mr = match_line(f, bar, bent, a.tags);
                        }
                        row = mr.row;
                        col = mr.col;
                    }
                    if (row >= 0) my_bar += weight;
                    if (col >= 0) {
                        my_tag += weight;
                        // warp-aggregated count update: one red per distinct cell
                        unsigned long long cell = (unsigned long long)(uint32_t)row * a.cols + (uint32_t)col;
                        uint32_t act = __activemask();
                        uint32_t peers = __match_any_sync(act, cell);
                        if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&a.matrix[cell], weight * __popc(peers));
                    }
                }
                if (tid == 0) {
                    // reads in this window with read index below the limit
                    // (read indices grow with r, so they form a prefix)
                    long long nreads = 0;
                    if (rbeg < whi) {
                        unsigned long long cntr = (whi - rbeg + 3) / 4;
                        unsigned long long first_idx = (first_line + rbeg) >> 2;
                        if (first_idx < limit) {
                            unsigned long long room = limit - first_idx;
                            nreads = (long long)(cntr < room ? cntr : room);
                        }
                    }
                    my_reads += weight * nreads;
                }
                __syncthreads();             // starts[] (and finally the stage) may be reused
            }
            if (total == 0) __syncthreads();
        } else {
            __syncthreads();
        }
        seg_lines += total;

        if (tid == 0) {
            if (a.mode == MODE_MAIN && tix == seg_ntiles - 1) {
                SegInfo si;
                si.lines = seg_lines;
                si.guess = guess;
                a.seginfo[seg] = si;
            }
            // ---- refill this stage --------------------------------------------
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            produce(s);
        }
    }

    if (MATCH) {
        // totals: one set of atomics per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_bar += __shfl_xor_sync(0xFFFFFFFFu, my_bar, o);
            my_tag += __shfl_xor_sync(0xFFFFFFFFu, my_tag, o);
        }
        if (lane == 0) {
            atomicAdd(&s_tot[0], (unsigned long long)(long long)my_bar);
            atomicAdd(&s_tot[1], (unsigned long long)(long long)my_tag);
        }
        __syncthreads();
        if (tid == 0) {
            if (my_reads) atomicAdd(&a.totals[0], (unsigned long long)my_reads);
            if (s_tot[0]) atomicAdd(&a.totals[1], s_tot[0]);
            if (s_tot[1]) atomicAdd(&a.totals[2], s_tot[1]);
        }
    }
}

// One CTA: prefix sum over the per-segment line counts, next chunk's state,
// and the list of segments the fix pass must redo.
constexpr int VERIFY_THREADS = 1024;
__global__ void __launch_bounds__(VERIFY_THREADS) verify_kernel(const VerifyArgs v)
{
    __shared__ unsigned long long s_sum[VERIFY_THREADS];
    const uint32_t tid = threadIdx.x;
    unsigned long long line_base;
    if (v.use_arg_state) line_base = v.line_base; else line_base = v.state_in->next_line;
    uint32_t per = (v.num_segs + VERIFY_THREADS - 1) / VERIFY_THREADS;
    uint32_t lo = tid * per, hi = lo + per < v.num_segs ? lo + per : v.num_segs;
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += v.seginfo[i].lines;
    s_sum[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partial sums
    for (int d = 1; d < VERIFY_THREADS; d <<= 1) {
        unsigned long long x = tid >= (uint32_t)d ? s_sum[tid - d] : 0;
        __syncthreads();
        s_sum[tid] += x;
        __syncthreads();
    }
    unsigned long long first = line_base + s_sum[tid] - sum;       // true first index of segment lo
    if (v.make_fixes) {
        for (uint32_t i = lo; i < hi; i++) {
            SegInfo si = v.seginfo[i];
            bool wrong = ((first ^ si.guess) & 3ull) != 0;
            // could any read of this segment reach the limit?
            bool past = si.lines && ((first + si.lines - 1) >> 2) >= v.reads_limit;
            if (si.lines && (wrong || past)) {
                uint32_t k = atomicAdd(v.n_fix, 1u);
                FixEntry fe;
                fe.seg = i;
                fe.guess = si.guess;
                fe.true_first = first;
                v.fix[k] = fe;
            }
            first += si.lines;
        }
    }
    if (tid == VERIFY_THREADS - 1) {
        v.state_out->next_line = line_base + s_sum[tid];
        v.state_out->prev_kind = *v.last_kind;
        v.state_out->pad = 0;
    }
}

#endif  // __CUDACC__

}  // namespace tdg
