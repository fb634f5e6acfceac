// The fused counting kernel: FASTQ line splitting, whitespace strip / case fold,
// 2-bit packing, barcode+cutsite lookup, tag probe and count update in ONE pass
// over the raw bytes (sm_100a).
//
// Replaces the loop of find_tags_fastq, /root/reference/tagdigger_fun.py:250-274:
//   for line in fqcon:                      -> line ends found 64 bytes per thread (SWAR)
//       if lineindex % 4 == 1:              -> global line index by single-pass
//                                              decoupled look-back over tiles
//           line1 = line.strip().upper()    -> leading-whitespace skip + case fold
//           sequence_index_lookup(x2)       -> tdg_match.h (packed exact-match probes)
//           mycounts[bar][tag] += 1         -> warp-aggregated red.global.add.s32
//
// Data movement: a persistent grid; each CTA draws tile numbers from a global
// ticket counter (tickets are handed out in order, which is what makes the
// look-back deadlock free) and keeps a ring of STAGES shared-memory buffers
// filled by the TMA unit (cp.async.bulk + mbarrier complete_tx).  Every byte of
// the stream is read from HBM exactly once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdg_match.h"

namespace tdg {

constexpr uint32_t TILE = 16384;          // == TDG_TILE_BYTES
constexpr uint32_t HALO = 512;            // == TDG_HALO_BYTES
constexpr uint32_t STAGE_BYTES = TILE + HALO;
constexpr uint32_t STAGE_STRIDE = STAGE_BYTES + 128;   // slack for unaligned word reads
constexpr int      STAGES = 3;
constexpr int      THREADS = 256;
constexpr int      WARPS = THREADS / 32;
constexpr uint32_t SPAN = TILE / THREADS; // bytes per thread in the line scan (64)
constexpr uint32_t STARTS_CAP = 2048;     // line starts kept per emission window
constexpr uint32_t BAR_SMEM_MAX = 16384;  // barcode tables up to this size are copied to smem
static_assert(SPAN == 64, "the scan below is written for 64 bytes per thread");

enum { PREV_NONE = 0, PREV_LF = 1, PREV_CR = 2, PREV_OTHER = 3 };

struct LineState {            // where a chunk starts in its file
    unsigned long long next_line;   // index the next line START will receive
    uint32_t prev_kind;             // PREV_*
    uint32_t pad;
};

struct ChunkArgs {
    const uint8_t *bytes;           // 16-byte aligned; allocation >= round_up(n, TILE) + HALO
    unsigned long long n;
    uint32_t num_tiles;
    uint32_t use_arg_state;         // 1: (line_base, prev_kind) below; 0: *state_in
    unsigned long long line_base;
    uint32_t prev_kind;
    uint32_t match;                 // 0: count lines only
    const LineState *state_in;
    LineState *state_out;
    unsigned long long *desc;       // [num_tiles], zeroed; desc[-1..] see layout below
    unsigned long long *ticket;     // zeroed
    unsigned long long reads_limit;
    const BarTable *bar;
    uint32_t bar_bytes;             // header + entries
    uint32_t cols;
    TagTable tags;
    int32_t *matrix;
    unsigned long long *totals;     // [4]: reads, barcode+cutsite hits, tag hits, (unused)
};

// descriptor word: [63:62] flag, [61:0] value
constexpr unsigned long long FLAG_AGG = 1ull << 62;
constexpr unsigned long long FLAG_PRE = 2ull << 62;
constexpr unsigned long long VAL_MASK = (1ull << 62) - 1;

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
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
    __shared__ uint32_t s_tile[STAGES];
    __shared__ uint32_t s_warp_cnt[WARPS];
    __shared__ unsigned long long s_prefix;        // line starts before this tile (chunk relative)
    __shared__ unsigned long long s_tot[3];

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

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
        s_tot[0] = s_tot[1] = s_tot[2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < STAGES; s++) {
            unsigned long long t = atomicAdd(a.ticket, 1ull);
            uint32_t tt = t < a.num_tiles ? (uint32_t)t : 0xFFFFFFFFu;
            s_tile[s] = tt;
            if (tt != 0xFFFFFFFFu) {
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                bulk_g2s(stage_base + s * STAGE_STRIDE, a.bytes + (size_t)tt * TILE, STAGE_BYTES, &full_bar[s]);
            }
        }
    }
    __syncthreads();

    // bytes a match may touch past the stripped line start (fast path bound)
    uint32_t need = 0;
    if (MATCH) {
        need = bar->max_tag_off + a.tags.max_len + 36u;
        if (bar->max_len + 36u > need) need = bar->max_len + 36u;
    }
    unsigned long long my_reads = 0;      // thread 0 only
    uint32_t my_bar = 0, my_tag = 0;

    for (uint32_t it = 0;; it++) {
        const uint32_t s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        const uint32_t t = s_tile[s];
        if (t == 0xFFFFFFFFu) break;
        mbar_wait(&full_bar[s], parity);

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
        }
        if (t == a.num_tiles - 1) {
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
        }
        const uint32_t cnt = __popc(mlo) + __popc(mhi);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp_cnt[warp] = incl;
        __syncthreads();
        uint32_t before = extra, total = extra;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t c = s_warp_cnt[w];
            if (w < (int)warp) before += c;
            total += c;
        }
        const uint32_t my_rank0 = before + incl - cnt;      // rank of my first line end in the tile

        // ---- look-back (warp 0) overlapped with emission (other warps first) --
        if (warp == 0) {
            unsigned long long excl = 0;
            if (t > 0) {
                if (lane == 0) st_desc(&a.desc[t], FLAG_AGG | (unsigned long long)total);
                long long idx = (long long)t - 1;
                for (;;) {
                    long long mine = idx - (long long)lane;
                    unsigned long long d = mine >= 0 ? ld_desc(&a.desc[mine]) : FLAG_PRE;
                    while (__any_sync(0xFFFFFFFFu, (d >> 62) == 0)) {
                        if ((d >> 62) == 0) d = ld_desc(&a.desc[mine]);
                    }
                    uint32_t pm = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                    uint32_t upto = pm ? (uint32_t)(__ffs(pm) - 1) : 31u;    // lanes 0..upto contribute
                    unsigned long long v = lane <= upto ? (d & VAL_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                    excl += v;
                    if (pm) break;
                    idx -= 32;
                }
            }
            if (lane == 0) {
                st_desc(&a.desc[t], FLAG_PRE | (excl + total));
                s_prefix = excl;
                if (t == a.num_tiles - 1) {
                    uint32_t c = buf[valid - 1];
                    a.state_out->next_line = line_base + excl + total;
                    a.state_out->prev_kind = c == '\n' ? PREV_LF : (c == '\r' ? PREV_CR : PREV_OTHER);
                }
            }
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
                __syncthreads();             // starts[] and s_prefix are ready

                const unsigned long long first_line = line_base + s_prefix;   // index of rank 0
                // sequence lines: index % 4 == 1
                uint32_t r0 = (uint32_t)((1ull - first_line) & 3ull);
                uint32_t whi = wlo + STARTS_CAP < total ? wlo + STARTS_CAP : total;
                // first r >= wlo with r % 4 == r0 % 4
                uint32_t rbeg = wlo + ((r0 - wlo) & 3u);
                for (uint32_t r = rbeg + 4 * tid; r < whi; r += 4 * THREADS) {
                    unsigned long long read_idx = (first_line + r) >> 2;
                    bool live = read_idx < a.reads_limit;
                    uint32_t pos = starts[r - wlo];       // always < avail (see the last-tile mask)
                    int32_t row = -1, col = -1;
                    if (live) {
                        // leading whitespace (str.strip)
                        const uint32_t staged = avail < STAGE_BYTES ? (uint32_t)avail : STAGE_BYTES;
                        while (pos < staged) {
                            uint32_t c = buf[pos];
                            if (is_space(c)) { pos++; continue; }
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
                                if (is_space(c)) { gp++; continue; }
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
                    if (row >= 0) my_bar++;
                    if (col >= 0) {
                        my_tag++;
                        // warp-aggregated count update: one red per distinct cell
                        unsigned long long cell = (unsigned long long)(uint32_t)row * a.cols + (uint32_t)col;
                        uint32_t act = __activemask();
                        uint32_t peers = __match_any_sync(act, cell);
                        if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&a.matrix[cell], __popc(peers));
                    }
                }
                if (tid == 0) {
                    // reads in this window with read index below the limit
                    // (read indices grow with r, so they form a prefix)
                    unsigned long long nreads = 0;
                    if (rbeg < whi) {
                        unsigned long long cntr = (whi - rbeg + 3) / 4;
                        unsigned long long first_idx = (first_line + rbeg) >> 2;
                        if (first_idx < a.reads_limit) {
                            unsigned long long room = a.reads_limit - first_idx;
                            nreads = cntr < room ? cntr : room;
                        }
                    }
                    my_reads += nreads;
                }
                __syncthreads();             // starts[] (and finally the stage) may be reused
            }
            if (total == 0) __syncthreads();
        } else {
            __syncthreads();
        }

        // ---- refill this stage ----------------------------------------------
        if (tid == 0) {
            unsigned long long t2 = atomicAdd(a.ticket, 1ull);
            uint32_t tt = t2 < a.num_tiles ? (uint32_t)t2 : 0xFFFFFFFFu;
            s_tile[s] = tt;
            if (tt != 0xFFFFFFFFu) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                bulk_g2s(stage_base + s * STAGE_STRIDE, a.bytes + (size_t)tt * TILE, STAGE_BYTES, &full_bar[s]);
            }
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
            atomicAdd(&s_tot[1], (unsigned long long)my_bar);
            atomicAdd(&s_tot[2], (unsigned long long)my_tag);
        }
        __syncthreads();
        if (tid == 0) {
            if (my_reads) atomicAdd(&a.totals[0], my_reads);
            if (s_tot[1]) atomicAdd(&a.totals[1], s_tot[1]);
            if (s_tot[2]) atomicAdd(&a.totals[2], s_tot[2]);
        }
    }
}

constexpr size_t count_kernel_smem()
{
    return (size_t)STAGES * STAGE_STRIDE + STARTS_CAP * sizeof(uint16_t) + BAR_SMEM_MAX;
}

#endif  // __CUDACC__

}  // namespace tdg
