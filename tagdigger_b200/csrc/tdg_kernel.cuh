// The fused counting kernel: FASTQ line splitting, whitespace strip / case fold,
// 2-bit packing, barcode+cutsite lookup, tag probe and count update in ONE pass
// over the raw bytes (sm_100a).
//
// Replaces the loop of find_tags_fastq, /root/reference/tagdigger_fun.py:250-274:
//   for line in fqcon:                      -> line ends found 16 bytes at a time (SWAR)
//       if lineindex % 4 == 1:              -> see "line numbering" below
//           line1 = line.strip().upper()    -> leading-whitespace skip + case fold
//           sequence_index_lookup(x2)       -> packed exact-match probes (prefix-free sets)
//           mycounts[bar][tag] += 1         -> warp-aggregated red.global.add.s32
//
// Execution model: every WARP is an independent pipeline.  A warp draws SEGMENTS
// (runs of consecutive tiles; long ones first, short ones at the end of a launch) from
// a global ticket counter -- the next ticket one tile ahead of its use -- and streams
// their tiles through its own ring of TWO shared-memory stages filled by the TMA unit
// (cp.async.bulk + mbarrier complete_tx, L2 evict-first).  There is no CTA-wide barrier
// after the prologue, and every byte of the stream is read from HBM exactly once.
//
// A tile is sized to be ONE BATCH: 7,680 bytes (+ up to 256 of halo) hold about 30
// reads of 250 bytes, one per lane of the matcher.  Per tile a warp
//   1. scans: each lane looks at its 240 contiguous bytes, 15 x 128-bit shared loads
//      (an odd number of 16-byte units per lane: consecutive lanes start in different
//      bank groups, so the loads are conflict free); per 4 bytes an add, a PRMT with
//      sign replication and a byte dot product leave the mask of the line-end
//      CANDIDATES, and 1.5 LOP3 check that every candidate is a line feed (otherwise
//      every candidate is classified exactly: '\n', '\r\n', lone '\r').  The masks stay
//      in registers;
//   2. ranks the line ends with three ballots, which gives every line start its index
//      in the file; a lane's span is handled as two halves of <= 128 bytes, each of
//      which holds at most one start of a SEQUENCE line in ordinary FASTQ -- selected
//      without a loop and written to the tile's queue slot given by its rank;
//   3. matches the tile's starts, one per lane (batch_front): pack 2 bits per base
//      straight from the staged bytes, barcode bucket lookup in shared memory, 128-bit
//      key, the loads of one hash probe of the L2-resident tag table.  The stage is
//      then refilled; the compare and the warp vote (batch_back) run in the middle of
//      the NEXT tile's scan, the warp-aggregated red.global.add.s32 at its end, so that
//      the probe's round trip and the vote hide behind the scan.
// Tiles that are not ordinary (chunk edges, lines shorter than half a span, the read
// limit of the fix pass) take a general walk over the same masks; reads with leading
// whitespace or non-ASCII text and table shapes outside the packed matcher's envelope
// take match_general.  Tags longer than the 128-bit key (up to 160 bases) use the long
// form, count_kernel<true, true>: the key is probed as usual and bases 64.. are
// compared against the table's side array.  Warp-uniform producer and segment state
// lives in registers: 128 registers, 14 warps per SM (two stages of 7.9 KB each).
//
// Line numbering.  Which lines are sequence lines is decided by the GLOBAL line
// index (lineindex % 4 == 1 counted from the start of the file), which a warp that
// starts in the middle of the file cannot know without every byte before it.  The
// kernel therefore runs speculatively and verifies:
//   1. count pass (MODE_MAIN): each segment numbers its lines from a GUESS of its
//      first line's index mod 4, read off the FASTQ structure of its first lines
//      ('@' line, '+' two lines later, equal sequence/quality length), and records
//      how many lines it saw.  Segment 0 knows its true index.
//   2. verify_kernel: a prefix sum over the per-segment line counts gives every
//      segment's true first line index; segments whose guess was wrong (or that
//      reach past the read limit) go on a fix list.  For real FASTQ it is empty.
//   3. fix pass (MODE_FIX): listed segments are counted again with the guessed
//      numbering and weight -1, then with the true numbering, the read limit and
//      weight +1.
// The result is exact for ANY byte stream; only the speed depends on the guess.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "tdg_match.h"

namespace tdg {

#ifndef TDG_WARPS
#define TDG_WARPS 14
#endif
#ifndef TDG_CHUNKS
#define TDG_CHUNKS 15
#endif
#ifndef TDG_HALO
#define TDG_HALO 256
#endif

constexpr int      WARPS = TDG_WARPS;              // independent pipelines per CTA (one CTA per SM)
constexpr int      THREADS = WARPS * 32;
constexpr uint32_t CHUNKS = TDG_CHUNKS;            // 16-byte pieces per lane per tile; odd: see scan
constexpr uint32_t SPAN = CHUNKS * 16;             // 240 bytes per lane
constexpr uint32_t MWORDS = (SPAN + 31) / 32;      // 32-bit mask words per lane
constexpr uint32_t HALF_WORDS = MWORDS / 2;        // a lane's span is handled as two halves (<= 128 bytes each)
constexpr uint32_t TILE = 32 * SPAN;               // 7,680 bytes per warp tile: about one batch of 32 reads
constexpr uint32_t HALO = TDG_HALO;                // bytes staged past a tile (<= TDG_HALO_BYTES)
constexpr uint32_t STAGE = TILE + HALO;            // one ring stage
#ifndef TDG_STAGES
#define TDG_STAGES 2
#endif
constexpr int      STAGES = TDG_STAGES;            // the tile being processed + the one(s) in flight
constexpr uint32_t RING = STAGES * STAGE;          // bytes of shared memory per warp
constexpr uint32_t QCAP = 64;                      // sequence-line starts a tile can queue on the common path
constexpr uint32_t BAR_SMEM_MAX = 10240;           // barcode tables up to this size (384-plex: 7.2 KB) are copied to smem
constexpr uint32_t GUESS_LINES = 16;               // lines inspected for the FASTQ structure guess
constexpr uint32_t FAST_WORDS_MAX = 24;            // 4-character words the fast matcher packs per read (tags <= 64 bases)
constexpr uint32_t LONG_WORDS_MAX = 48;            // ... in its long-tag form (tags <= 160 bases: key + 96 bases of tail)
constexpr uint32_t LONG_TAIL_MAX = 96;             // bases past the 128-bit key that the long form compares
constexpr uint32_t TIX_LAST = 0x80000000u;         // tile metadata: last tile of its segment
static_assert(CHUNKS % 2 == 1, "an odd chunk count keeps the 128-bit scan loads free of bank conflicts");
static_assert(TILE % 16 == 0 && STAGE % 16 == 0, "tiles must keep the 16-byte alignment TMA needs");
static_assert(RING < 65536, "queue entries are 16-bit offsets into a warp's ring");
static_assert((MWORDS - HALF_WORDS) * 32 <= 128 && HALF_WORDS * 32 <= 128, "a half must not exceed 128 bytes");

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
    const uint8_t *bytes;           // 16-byte aligned; allocation >= round_up(n, 16)
    unsigned long long n;
    uint32_t num_tiles;
    uint32_t seg_tiles;             // tiles per segment, segments [0, n_big)
    uint32_t n_big;                 // the launch ends in SHORT segments (even finish): segments >= n_big have
    uint32_t seg_small;             //   seg_small tiles each
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
    uint32_t halo_bytes;            // bytes copied past each tile (multiple of 16, <= HALO)
    uint32_t fast_words;            // > 0: tables fit the fast matcher, which packs this many words
    uint32_t need;                  // bytes a match may touch past the stripped line start
    uint32_t ulen;                  // > 0: every tag has this length (<= 64), and
    uint32_t um[4];                 //      these are its four 32-bit compare masks
    uint32_t bar_in_smem;           // the barcode table is copied into shared memory
    unsigned long long tag_km;      // mask of the first K bases of a key (class 0)
    TagTable tags;
    int32_t *matrix;
    int32_t *replicas;              // [n_replicas - 1][cells] extra copies of a SMALL matrix (or null)
    uint32_t n_replicas;            // warps spread their updates over the copies: hot cells stop serialising in L2
    uint32_t cells;
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
    unsigned long long *run_total;  // [VERIFY_MAX_CTAS * 32] lines per run of segments (verify_sums -> verify_kernel)
};

// Can the fast matcher serve these tables?  Returns the number of 4-character
// words it has to pack per read (0 = use the general matcher).  Host and device
// agree on this through ChunkArgs::fast_words.
inline uint32_t fast_words_for(const BarTable *bar, const TagTable &tt)
{
    if (bar->any_base || tt.any_base) return 0;
    if (bar->max_len > 16 || bar->max_tag_off > 28) return 0;
    if (tt.n_classes != 1 || tt.max_len > 64 + LONG_TAIL_MAX || tt.min_len < 1) return 0;
    uint32_t chars = 3 + bar->max_tag_off + tt.max_len;       // 3: worst misalignment of the line start
    if (chars < 3 + 16) chars = 3 + 16;                       // the barcode key is always 16 bases
    uint32_t nw = (chars + 3) / 4;
    return nw <= (tt.max_len > 64 ? LONG_WORDS_MAX : FAST_WORDS_MAX) ? nw : 0;
}
// Tags longer than the 128-bit key take the long form of the fast matcher (count_kernel<true, true>).
inline bool fast_is_long(const TagTable &tt) { return tt.max_len > 64; }

#if defined(__CUDACC__)

struct alignas(16) WarpShared {   // per-warp scratch in shared memory
    uint16_t q[QCAP];                // sequence-line starts of the current tile: offsets into the warp's ring
    uint16_t gs[GUESS_LINES + 8];    // first line starts of a segment (structure guess)
};
static_assert(sizeof(WarpShared) % 16 == 0, "control blocks must keep the barcode table 16-byte aligned");

constexpr size_t SMEM_FIXED = (size_t)WARPS * RING + (size_t)WARPS * sizeof(WarpShared);

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
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    // the FASTQ bytes are read once: let them leave L2 first, so that the tag table and the count
    // matrix (probed at random) stay resident
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// Line-end candidates, 16 bytes at a time.  b + 0x60 has its top bit set for every byte in
// 0x20..0x9F, i.e. "certainly no control character"; PRMT's sign replication turns that bit
// into 0xFF / 0x00 bytes, and a signed-by-unsigned byte dot product with the weights
// 1,2,4,..,128, started at 255, leaves the 8-bit mask of the CANDIDATES of eight bytes:
// three instructions per four bytes (IADD, PRMT, IDP.4A).
// The add carries between bytes, so the test is one-sided on purpose:
//   * '\n' (0x0A) and '\r' (0x0D) are ALWAYS candidates: with or without a carry from the
//     byte below, b + 0x60 (+1) stays under 0x80 for every b <= 0x1E;
//   * a byte >= 0xA0 wraps around and becomes a candidate although it is no control
//     character, and 0x1F right above such a byte is missed (it ends no line: nothing lost).
// The candidate walk checks that every candidate really is '\n' and switches the warp to the
// exact classifier otherwise (see `classify`), so the line ends found are exact for all
// byte values.
__device__ __forceinline__ uint32_t notctl4(uint32_t w)
{
    uint32_t r;                                              // 0xFF where bit 7 of the sum is set
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w + 0x60606060u), "r"(0u), "r"(0xBA98u));
    return r;
}
__device__ __forceinline__ int32_t dp4a_su(uint32_t a_signed_bytes, uint32_t b_unsigned_bytes, int32_t c)
{
    int32_t d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_signed_bytes), "r"(b_unsigned_bytes), "r"(c));
    return d;
}
// `dev` collects (candidate byte) ^ '\n' over all candidates: it stays zero exactly when every
// candidate is a line feed, which is all the common path needs to know about them.
__device__ __forceinline__ uint32_t ctl_mask16(uint4 q, uint32_t &dev)
{
    const uint32_t nx = notctl4(q.x), ny = notctl4(q.y), nz = notctl4(q.z), nw = notctl4(q.w);
    const int32_t lo = dp4a_su(ny, 0x80402010u, dp4a_su(nx, 0x08040201u, 255));   // bytes 0-7
    const int32_t hi = dp4a_su(nw, 0x80402010u, dp4a_su(nz, 0x08040201u, 255));   // bytes 8-15
    dev |= ((q.x ^ 0x0A0A0A0Au) & ~nx) | ((q.y ^ 0x0A0A0A0Au) & ~ny);
    dev |= ((q.z ^ 0x0A0A0A0Au) & ~nz) | ((q.w ^ 0x0A0A0A0Au) & ~nw);
    return (uint32_t)hi * 256u + (uint32_t)lo;
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

// General matcher for one line (any table shape, leading whitespace, lines cut
// by the end of the chunk, lines that leave the staged bytes).  Rare: kept out
// of line so that the fast path stays small.
//   buf/pos: the line start inside a stage; staged: bytes of the stage that are
//   valid (tile + halo, clipped to the chunk); gtile: the same tile in global
//   memory; avail: bytes from the tile start to the end of the chunk.
__device__ __noinline__ MatchResult match_general(const uint8_t *buf, uint32_t pos, uint32_t staged,
                                                  const uint8_t *gtile, unsigned long long avail, uint32_t need,
                                                  const BarTable *bar, const BarEntry *bent, const TagTable *tt)
{
    // leading whitespace (str.strip)
    while (pos < staged) {
        uint32_t c = buf[pos];
        if (is_lead_space(c)) { pos++; continue; }
        if (c >= 0xC2 && c <= 0xE3 && pos + 2 < staged) {
            uint32_t u = utf8_space(c, buf[pos + 1], buf[pos + 2]);
            if (u) { pos += u; continue; }
        }
        break;
    }
    if (pos + need <= staged || staged == avail) {
        SmemFetch f;
        f.p = buf + pos;
        unsigned long long room = avail - pos;
        f.limit = room > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)room;
        return match_line(f, bar, bent, *tt);
    }
    // continue from global memory
    unsigned long long gp = pos;
    while (gp < avail) {
        uint32_t c = gtile[gp];
        if (is_lead_space(c)) { gp++; continue; }
        if (c >= 0xC2 && c <= 0xE3 && gp + 2 < avail) {
            uint32_t u = utf8_space(c, gtile[gp + 1], gtile[gp + 2]);
            if (u) { gp += u; continue; }
        }
        break;
    }
    GlobalFetch f;
    f.p = gtile + gp;
    unsigned long long room = avail - gp;
    f.limit = room > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)room;
    return match_line(f, bar, bent, *tt);
}

// One 4-character word of the fast matcher: 2-bit codes gathered in the top
// byte of the result; `bad` non-zero iff a character is not one of ACGTacgt.
__device__ __forceinline__ uint32_t pack_word(uint32_t w, uint32_t &bad)
{
    uint32_t c2 = (w >> 1) & 0x03030303u;             // A0 C1 T2 G3 (bits 1-2 of the character)
    uint32_t m = (c2 >> 1) & ~c2 & 0x01010101u;       // 1 where code == 2 (T)
    // With bits 1-2 accounted for by the code and bit 5 (case) ignored, the remaining bits
    // 0,3,4,6,7 of a base are 0x41 for A/C/G and 0x50 for T.
    bad = (w & 0xD9D9D9D9u) ^ (0x41414141u ^ (m * 0x11u));
    return c2 * 0x01041040u;                          // byte 3 = c0 | c1<<2 | c2<<4 | c3<<6
}

__device__ __forceinline__ uint32_t lowmask32(uint32_t nbases)    // first nbases (<= 16) bases of a 32-bit word
{
    return nbases >= 16 ? 0xFFFFFFFFu : ((1u << (2 * nbases)) - 1u);
}

template <bool MATCH, bool LONG = false>
__global__ void __launch_bounds__(THREADS, 1) count_kernel(const __grid_constant__ ChunkArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[WARPS][STAGES];

    constexpr uint32_t NONE = 0xFFFFFFFFu;
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    uint8_t *const wbase = smem + (size_t)warp * RING;
    WarpShared *const ws = (WarpShared *)(smem + (size_t)WARPS * RING) + warp;
    uint8_t *const bar_smem = smem + SMEM_FIXED;

    // ---- chunk-level state -------------------------------------------------
    const uint32_t prev_kind = a.use_arg_state ? a.prev_kind : a.state_in->prev_kind;
    if ((a.mode == MODE_FIX ? *a.n_fix : a.num_segs) == 0) return;

    const BarTable *bar = a.bar;
    if (MATCH) {
        if (a.bar_in_smem) {
            const uint4 *src = (const uint4 *)a.bar;
            uint4 *dst = (uint4 *)bar_smem;
            for (uint32_t i = tid; i < (a.bar_bytes + 15u) / 16u; i += THREADS) dst[i] = src[i];
            bar = (const BarTable *)bar_smem;
        }
    }
    const BarEntry *bent = (const BarEntry *)((const uint8_t *)bar + sizeof(BarTable));
    if (lane == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();           // the only CTA-wide barrier: table copy + barrier init

    const uint32_t copy_bytes = TILE + a.halo_bytes;

    // ---- producer: request the next tile into stage st -----------------------------------------
    // All of its state is warp-uniform and lives in registers; every lane runs the (uniform)
    // arithmetic, lane 0 alone draws the ticket and issues the copy.
    uint32_t p_item = NONE, p_seg = 0, p_tix = 0, p_ntiles = 0;
    // The ticket of the NEXT segment is drawn one tile ahead (when the current segment's last tile
    // is requested): the atomic's round trip to L2 is over by the time the ticket is needed.
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd((unsigned int *)a.ticket, 1u);
    auto produce = [&](uint32_t st, uint32_t &m_tile, uint32_t &m_item, uint32_t &m_tixf) {
        if (p_tix == p_ntiles) {                                   // the next segment
            const uint32_t n_items = a.mode == MODE_FIX ? 2u * *a.n_fix : a.num_segs;
            const uint32_t t = __shfl_sync(FULL, ticket, 0);
            p_tix = 0;
            if (t < n_items) {
                p_item = t;
                const uint32_t seg = a.mode == MODE_FIX ? a.fix[t >> 1].seg : t;
                // first tile and length of the segment: long segments first, short ones at the end
                const bool big = seg < a.n_big;
                const uint32_t want = big ? a.seg_tiles : a.seg_small;
                p_seg = big ? seg * a.seg_tiles : a.n_big * a.seg_tiles + (seg - a.n_big) * a.seg_small;
                const uint32_t left = a.num_tiles - p_seg;
                p_ntiles = left < want ? left : want;
            } else {
                p_item = NONE;
                p_ntiles = NONE;                                   // never equal to p_tix again: no more tickets
            }
            // The ticket and (fix pass) the list entry are LOADED values: pin them down inside this rare
            // branch.  Otherwise the wait for them is placed at their first use after the branch, where
            // it runs every tile and -- scoreboards being few and shared -- also waits for the probe
            // loads that batch_front has in flight.
            asm volatile("" ::"r"(p_item), "r"(p_seg), "r"(p_ntiles));
        }
        m_tile = p_seg + p_tix;                  // (p_seg: first tile of the segment)
        m_item = p_item;
        m_tixf = p_tix + 1 == p_ntiles ? (p_tix | TIX_LAST) : p_tix;
        if (p_item != NONE) {
            if (lane == 0) {
                const unsigned long long off = (unsigned long long)m_tile * TILE;
                uint32_t bytes = copy_bytes;
                if (__builtin_expect(m_tile + 2 >= a.num_tiles, 0)) {    // only the last tiles can run past the data
                    const unsigned long long left = a.n - off;
                    if (left < bytes) bytes = ((uint32_t)left + 15u) & ~15u;
                }
                mbar_expect_tx(&full_bar[warp][st], bytes);
                bulk_g2s(wbase + st * STAGE, a.bytes + off, bytes, &full_bar[warp][st]);
            }
            p_tix++;
            if (p_tix == p_ntiles && lane == 0) ticket = atomicAdd((unsigned int *)a.ticket, 1u);
        }
    };

    // ---- matcher set-up: everything uniform comes straight from the kernel arguments ----
    const uint32_t nw = MATCH ? a.fast_words : 0;       // 0: general matcher only
    int32_t *wmatrix = a.matrix;                        // the copy of the count matrix this warp updates
    if (MATCH && a.n_replicas > 1) {
        const uint32_t rep = (blockIdx.x * WARPS + warp) % a.n_replicas;
        if (rep) wmatrix = a.replicas + (size_t)(rep - 1) * a.cells;
    }
    const uint4 *tag_entries = (const uint4 *)a.tags.entries;
    int32_t my_bar = 0, my_tag = 0;

    // ---- per-segment state (uniform across the warp) -----------------------------------------
    uint32_t seg_lines = 0;               // line starts numbered so far
    uint32_t seg_reads = 0;               // reads numbered so far (those below the limit)
    uint32_t phase = 0;                   // (assumed) index of the segment's first line start, mod 4
    bool has_limit = false;               // fix pass, true numbering: a.reads_limit applies
    int32_t weight = 1;
    bool need_guess = false;
    bool classify_first = false;          // sticky: this input has control characters other than '\n'

    // metadata of the tile being processed (cur) and of the ones in flight (nxt, oldest first)
    uint32_t cur_tile = 0, cur_item = NONE, cur_tixf = 0;
    uint32_t nxt_tile[STAGES - 1], nxt_item[STAGES - 1], nxt_tixf[STAGES - 1];

    // 128-bit compare of a table entry with the read's key over the entry's length
    auto tag_differs = [&](const uint4 &k, uint32_t L, uint32_t T0, uint32_t T1, uint32_t T2, uint32_t T3) -> uint32_t {
        if (a.ulen) return ((k.x ^ T0) & a.um[0]) | ((k.y ^ T1) & a.um[1]) | ((k.z ^ T2) & a.um[2]) | ((k.w ^ T3) & a.um[3]);
        return ((k.x ^ T0) & lowmask32(L)) | ((k.y ^ T1) & lowmask32(L > 16 ? L - 16 : 0)) |
               ((k.z ^ T2) & lowmask32(L > 32 ? L - 32 : 0)) | ((k.w ^ T3) & lowmask32(L > 48 ? L - 48 : 0));
    };

    // ---- matching, software-pipelined ------------------------------------------------
    // batch_front takes up to 32 queued line starts of the CURRENT tile, one per lane: pack,
    // barcode lookup, tag key, hash -- and ISSUES the loads of the first two table slots.
    // batch_back, called when the next tile is opened, compares, finishes rare longer probe
    // sequences and votes (MATCH.ANY); red_retire, after that tile's scan, issues the reds.
    // The L2 round trip of the probe hides behind the refill, the MATCH behind the scan.
    bool pb_pending = false;              // uniform: a batch is between front and back
    int32_t pb_row = -1, pb_col = -1;     // per lane
    bool pb_probe = false;                // per lane: slots loaded, compare still to do
    uint32_t pb_T0 = 0, pb_T1 = 0, pb_T2 = 0, pb_T3 = 0, pb_h = 0, pb_V = 0xFFFFFFFFu, pb_end = 0;
    uint4 pb_k0 = make_uint4(0, 0, 0, 0), pb_k1 = pb_k0;
    uint2 pb_m0 = make_uint2(0, 0), pb_m1 = pb_m0;      // len | flags, column (the first half of an entry's second 16 bytes)
    uint32_t pb_x0 = 0, pb_x1 = 0;                      // LONG: where the entries' bases 64.. sit in the side array
    uint32_t pb_TL[LONG ? LONG_TAIL_MAX / 16 : 1];      // LONG: bases 64..159 of the read's tag window
    constexpr uint32_t NG = (LONG ? LONG_WORDS_MAX : FAST_WORDS_MAX) / 4;     // groups of four packed words

    auto batch_front = [&](uint32_t qoff, uint32_t nb, uint32_t st) {
        uint32_t off = st * STAGE;
        const bool have = lane < nb;
        if (have) off = ws->q[qoff + lane];                       // lanes without a read work on the stage start, results dropped
        pb_row = -1;
        pb_col = -1;
        pb_probe = false;
        // first character of the line, out of the aligned word the pack starts with
        const uint32_t c0 = (*(const uint32_t *)(wbase + (off & ~3u)) >> (8u * (off & 3u))) & 0xFFu;
        // the general matcher takes: table shapes outside the fast envelope, the last tiles of a
        // chunk (lines may be cut by the end of the data), leading whitespace, non-ASCII
        // (evaluated without short-circuit branches: the warp must stay converged for the pack)
        const bool lead = (c0 == 0x09) | (c0 == 0x0b) | (c0 == 0x0c) | ((c0 - 0x1cu) <= 4u) | (c0 >= 0x80);
        const bool slow = have & ((nw == 0) | (cur_tile + 2 >= a.num_tiles) | lead);
        __syncwarp();
        if (nw != 0) {
            // ---- fast matcher, ALL lanes in lock step (the results of `slow` lanes and of lanes
            // without a read are discarded): pack the first 4*nw characters, from the aligned
            // word that holds the line start, once, 2 bits per base
            const uint32_t sh = off & 3u;
            const uint32_t *wp = (const uint32_t *)(wbase + (off & ~3u));
            uint32_t P[NG + 1];
            uint32_t GB[NG];                  // non-zero: words 4g..4g+3 hold a character outside ACGTacgt
            auto pack_group = [&](uint32_t g) {
                uint32_t x[4], gbad = 0;
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    uint32_t bad;
                    x[k] = pack_word(wp[4 * g + k], bad);
                    if (g == 0 && k == 0) bad &= 0xFFFFFFFFu << (8u * sh);   // bytes before the line start
                    gbad |= bad;          // words past nw may flag too: V below ignores them
                }
                GB[g] = gbad;
                P[g] = __byte_perm(__byte_perm(x[0], x[1], 0x0073), __byte_perm(x[2], x[3], 0x0073), 0x5410);
            };
            // groups 0..3 (64 characters) hold the barcode, the cut site and the first 32 bases
            // of the tag, which is all the hash needs: the probe loads go out before the rest
#pragma unroll
            for (uint32_t g = 0; g < NG; g++) { P[g] = 0; GB[g] = 0; }
            P[NG] = 0;
            // (always four groups: the halo covers them whatever the table shape)
            pack_group(0);
            pack_group(1);
            // ---- barcode + cut site: first 16 bases, bucket = first 4
            const uint32_t key0 = __funnelshift_r(P[0], P[1], 2u * sh);
            uint32_t tag_off = 0, blen = 0;
            int32_t row = -1;
            {
                // entry: key (low half), compare mask (tdg_tables.h puts it in the key's unused high
                // half), row, len | tag_off << 16.  A bucket rarely holds more than two patterns:
                // its first two entries are compared without a loop (the table ends in two zero
                // entries, so these loads never leave it), the rest in a rare divergent loop.
                const uint32_t b = key0 & 0xFFu;
                const uint32_t lo = bar->bucket[b], hi = bar->bucket[b + 1];
                const uint4 be0 = *(const uint4 *)(bent + lo), be1 = *(const uint4 *)(bent + lo + 1);
                const bool h0 = (lo < hi) & (((key0 ^ be0.x) & be0.y) == 0);
                const bool h1 = (lo + 1 < hi) & (((key0 ^ be1.x) & be1.y) == 0);
                uint32_t rz = h0 ? be0.z : be1.z, rw = h0 ? be0.w : be1.w;
                bool hit = h0 | h1;
                if (!hit & (lo + 2 < hi)) {
                    for (uint32_t e = lo + 2; e < hi; e++) {
                        const uint4 be = *(const uint4 *)(bent + e);
                        if (((key0 ^ be.x) & be.y) == 0) { rz = be.z; rw = be.w; hit = true; break; }
                    }
                }
                if (hit) {
                    row = (int32_t)rz;
                    blen = rw & 0xFFFFu;
                    tag_off = rw >> 16;
                }
            }
            __syncwarp();             // the bucket walks have different lengths: reconverge here
            pack_group(2);            // (the bucket's shared-memory loads overlap this packing)
            pack_group(3);
            // ---- tag key at tag_off: the first 32 bases give the slot; fetch the first two
            // slots of the probe sequence right away (speculatively: validity is checked below)
            const uint32_t toff = sh + tag_off;
            const uint32_t bit = (toff & 15u) * 2u;
            const bool up = (toff >> 4) != 0;          // toff <= 31
            {
                const uint32_t Q0 = up ? P[1] : P[0], Q1 = up ? P[2] : P[1], Q2 = up ? P[3] : P[2];
                pb_T0 = __funnelshift_r(Q0, Q1, bit);
                pb_T1 = __funnelshift_r(Q1, Q2, bit);
            }
            const uint64_t pre = (((uint64_t)pb_T1 << 32) | pb_T0) & a.tag_km;
            pb_h = tag_slot(pre, a.tags.cls[0].mask);            // even: both slots share a 64-byte line
            const bool want = have & !slow & (row >= 0);
            {
                // unconditional (lanes that do not probe read slot 0): a predicated load would be
                // staged through temporaries and copied, and the copy waits for the data at once
                const uint4 *e0 = tag_entries + 2 * (size_t)(a.tags.cls[0].base + (want ? pb_h : 0u));
                pb_k0 = __ldg(e0);
                if (LONG) {
                    const uint4 f0 = __ldg(e0 + 1), f1 = __ldg(e0 + 3);     // len, column, tail index
                    pb_m0 = make_uint2(f0.x, f0.y);
                    pb_m1 = make_uint2(f1.x, f1.y);
                    pb_x0 = f0.z;
                    pb_x1 = f1.z;
                } else {
                    pb_m0 = __ldg((const uint2 *)(e0 + 1));
                    pb_m1 = __ldg((const uint2 *)(e0 + 3));
                }
                pb_k1 = __ldg(e0 + 2);
            }
            // ---- the rest of the tag (bases 32..) while the loads are in flight
#pragma unroll
            for (uint32_t g = 4; g < NG; g++)
                if (nw > 4 * g) pack_group(g);
            {
                const uint32_t Q2 = up ? P[3] : P[2], Q3 = up ? P[4] : P[3], Q4 = up ? P[5] : P[4];
                pb_T2 = __funnelshift_r(Q2, Q3, bit);
                pb_T3 = __funnelshift_r(Q3, Q4, bit);
            }
            if (LONG) {
#pragma unroll
                for (uint32_t j = 0; j < LONG_TAIL_MAX / 16; j++) {
                    const uint32_t Qa = up ? P[5 + j] : P[4 + j], Qb = up ? P[6 + j] : P[5 + j];
                    pb_TL[j] = __funnelshift_r(Qa, Qb, bit);
                }
            }
            // some character is not a base: matches stand only if they end before it
            uint32_t V = 0xFFFFFFFFu;                  // valid bases from the line start
            uint32_t anybad = 0;
#pragma unroll
            for (uint32_t g = 0; g < NG; g++) anybad |= GB[g];
            if (anybad != 0) {
                // The first flagged group of four words brackets the first such character:
                // [Vmin, Vmax].  Mostly that decides already -- everything a match needs lies
                // before the group, or some of it surely lies behind -- and V = Vmin gives the
                // right answer; only a lane whose match could end INSIDE the group looks for the
                // exact position.
                uint32_t i0 = 0;
#pragma unroll
                for (int g = (int)NG - 1; g >= 0; g--)
                    if (GB[g] != 0) i0 = 4u * g;
                const uint32_t Vmin = i0 ? 4u * i0 - sh : 0u, Vmax = 4u * i0 + 15u - sh;
                const uint32_t end_max = tag_off + a.tags.max_len, end_min = tag_off + a.tags.min_len;
                // (the barcode+cutsite decision feeds a total of its own, so it is bracketed by itself)
                const bool bar_amb = (blen > Vmin) & (blen <= Vmax);
                const bool tag_amb = (blen <= Vmin) & (end_max > Vmin) & (end_min <= Vmax);
                V = Vmin;
                if (want & (bar_amb | tag_amb)) {
                    uint32_t firstbad = 0, at = 0;
#pragma unroll
                    for (int k = 3; k >= 0; k--) {
                        uint32_t bad;
                        (void)pack_word(wp[i0 + k], bad);
                        if (i0 + k == 0) bad &= 0xFFFFFFFFu << (8u * sh);
                        if (bad != 0 && i0 + k < nw) { firstbad = bad; at = i0 + k; }
                    }
                    V = firstbad ? 4u * at + ((uint32_t)(__ffs(firstbad) - 1) >> 3) - sh : 0xFFFFFFFFu;
                }
            }
            __syncwarp();
            pb_probe = want & (blen <= V);
            if (pb_probe) {
                pb_row = row;
                pb_V = V;
                pb_end = tag_off;
            }
        }
        if (slow) {
            const uint32_t p = off - st * STAGE;
            const unsigned long long tile_off = (unsigned long long)cur_tile * TILE;
            const unsigned long long avail = a.n - tile_off;
            const uint32_t staged = avail < copy_bytes ? (uint32_t)avail : copy_bytes;
            MatchResult mr = match_general(wbase + st * STAGE, p, staged, a.bytes + tile_off, avail, a.need, bar, bent, &a.tags);
            pb_row = mr.row;
            pb_col = mr.col;
        }
        __syncwarp();
        pb_pending = true;
    };

    // The count update of a batch is issued a scan later than its vote: MATCH.ANY takes a while,
    // and this way nothing waits for it.
    uint32_t pr_cell = NONE, pr_peers = 0;         // (retired within the segment of their batch: `weight` still applies)
    auto red_retire = [&]() {
        // warp-aggregated: one red per distinct cell
        if (pr_cell != NONE && lane == (uint32_t)(__ffs(pr_peers) - 1))
            asm volatile("red.global.add.s32 [%0], %1;" ::"l"(wmatrix + pr_cell), "r"(weight * (int32_t)__popc(pr_peers)) : "memory");
        pr_cell = NONE;
    };

    auto batch_back = [&]() {
        red_retire();
        {
            // the first slot pair of the probe sequence, without branches (every lane computes,
            // probing lanes use the result)
            const bool e0 = pb_m0.x == TDG_EMPTY_LEN, e1 = pb_m1.x == TDG_EMPTY_LEN;
            const uint32_t L0 = pb_m0.x & TDG_LEN_MASK;
            bool hit0 = !e0 & (tag_differs(pb_k0, L0, pb_T0, pb_T1, pb_T2, pb_T3) == 0);
            bool hit1 = !e0 & !e1 & (LONG | !hit0) & (tag_differs(pb_k1, pb_m1.x, pb_T0, pb_T1, pb_T2, pb_T3) == 0);
            if (LONG) {
                // Tags longer than the key: bases 64.. of both candidates come from the side array (one
                // more round trip to L2, for both slots at once -- two alleles that differ past base 64
                // share a key) and are compared with the read's packed tail.
                const uint64_t *x0 = a.tags.ext + ((hit0 & pb_probe & (L0 > 64u)) ? pb_x0 : 0u);
                const uint64_t *x1 = a.tags.ext + ((hit1 & pb_probe & (pb_m1.x > 64u)) ? pb_x1 : 0u);
                uint2 t0[LONG_TAIL_MAX / 32], t1[LONG_TAIL_MAX / 32];
#pragma unroll
                for (uint32_t j = 0; j < LONG_TAIL_MAX / 32; j++) {
                    t0[j] = __ldg((const uint2 *)(x0 + j));
                    t1[j] = __ldg((const uint2 *)(x1 + j));
                }
                auto tail_differs = [&](const uint2 (&t)[LONG_TAIL_MAX / 32], uint32_t L) -> uint32_t {
                    const uint32_t rest = L > 64u ? L - 64u : 0u;
                    uint32_t d = 0;
#pragma unroll
                    for (uint32_t j = 0; j < LONG_TAIL_MAX / 32; j++) {
                        d |= (t[j].x ^ pb_TL[2 * j]) & lowmask32(rest > 32u * j ? rest - 32u * j : 0u);
                        d |= (t[j].y ^ pb_TL[2 * j + 1]) & lowmask32(rest > 32u * j + 16u ? rest - 32u * j - 16u : 0u);
                    }
                    return d;
                };
                hit0 = hit0 & (tail_differs(t0, L0) == 0);
                hit1 = hit1 & !hit0 & (tail_differs(t1, pb_m1.x) == 0);
            }
            int32_t col = hit0 ? (int32_t)pb_m0.y : (hit1 ? (int32_t)pb_m1.y : -1);
            uint32_t tlen = hit0 ? L0 : pb_m1.x;
            // (LONG, rare path below) bases 64.. of the entry whose second half is at `half`
            auto long_tail_ok = [&](const uint4 *half, uint32_t L) -> bool {
                if (L <= 64u) return true;
                const uint64_t *x = a.tags.ext + __ldg(half).z;
                const uint32_t rest = L - 64u;
                uint32_t d = 0;
#pragma unroll
                for (uint32_t j = 0; j < LONG_TAIL_MAX / 32; j++) {
                    const uint2 t = __ldg((const uint2 *)(x + j));
                    d |= (t.x ^ pb_TL[LONG ? 2 * j : 0]) & lowmask32(rest > 32u * j ? rest - 32u * j : 0u);
                    d |= (t.y ^ pb_TL[LONG ? 2 * j + 1 : 0]) & lowmask32(rest > 32u * j + 16u ? rest - 32u * j - 16u : 0u);
                }
                return d == 0;
            };
            // rare: both slots taken by other keys and something was stored beyond this pair
            if (pb_probe & !e0 & !e1 & !hit0 & !hit1 & ((pb_m0.x & TDG_LEN_MORE) != 0)) {
                uint32_t h = pb_h;
                for (;;) {
                    h = (h + 2) & a.tags.cls[0].mask;                 // slots come in even-aligned pairs
                    const uint4 *e = tag_entries + 2 * (size_t)(a.tags.cls[0].base + h);
                    const uint4 k0 = __ldg(e), k1 = __ldg(e + 2);
                    const uint2 m0 = __ldg((const uint2 *)(e + 1)), m1 = __ldg((const uint2 *)(e + 3));
                    if (m0.x == TDG_EMPTY_LEN) break;
                    const uint32_t l0 = m0.x & TDG_LEN_MASK;
                    if (tag_differs(k0, l0, pb_T0, pb_T1, pb_T2, pb_T3) == 0 && (!LONG || long_tail_ok(e + 1, l0))) {
                        col = (int32_t)m0.y; tlen = l0; break;
                    }
                    if (m1.x == TDG_EMPTY_LEN) break;
                    if (tag_differs(k1, m1.x, pb_T0, pb_T1, pb_T2, pb_T3) == 0 && (!LONG || long_tail_ok(e + 3, m1.x))) {
                        col = (int32_t)m1.y; tlen = m1.x; break;
                    }
                    if (!(m0.x & TDG_LEN_MORE)) break;
                }
            }
            __syncwarp();
            if (pb_probe) pb_col = (col >= 0 && pb_end + tlen > pb_V) ? -1 : col;     // -1: the tag runs into a non-base
        }
        if (pb_row >= 0) my_bar += weight;
        uint32_t cell = NONE;
        if (pb_col >= 0) {
            my_tag += weight;
            cell = (uint32_t)pb_row * a.cols + (uint32_t)pb_col;
        }
        pr_cell = cell;
        pr_peers = __match_any_sync(FULL, cell);
        pb_pending = false;
    };

    // ---- the tile loop ------------------------------------------------------------------------
    // Rotated: an iteration FINISHES the tile that the previous iteration opened (ranks, emission,
    // batch_front), refills its stage, and then OPENS the next one (batch_back, wait for its bytes,
    // scan, red_retire).  Only straight-line code sits between batch_front's probe loads and
    // batch_back, and between the vote and the reds: ptxas waits for everything outstanding at
    // loop headers.
    uint32_t s = 0, parity = 0;
    uint32_t mk[MWORDS];                  // line-end (candidate) masks of my SPAN bytes of the open tile
#pragma unroll
    for (uint32_t j = 0; j < MWORDS; j++) mk[j] = 0;
    uint32_t scan_dev = 0;
    bool opened = false;

    produce(0, cur_tile, cur_item, cur_tixf);
#pragma unroll
    for (int k = 0; k < STAGES - 1; k++) produce(k + 1, nxt_tile[k], nxt_item[k], nxt_tixf[k]);

    for (;;) {
        if (opened) {
            const uint8_t *const buf = wbase + s * STAGE;
            const uint32_t sbase = s * STAGE;
            const uint32_t t = cur_tile;
            const bool seg_end = (cur_tixf & TIX_LAST) != 0;
            // Only a chunk's last tiles can be cut by the end of the data, only its first tile can
            // follow an implicit line end: everything in between takes neither branch.
            const bool edge = (t == 0) | (t + 2 >= a.num_tiles);
            uint32_t avail = 0xFFFFFFFFu;         // bytes from the tile start to the end of the chunk, saturated
            uint32_t extra = 0;                   // 1: the tile's first byte starts a line (chunk start)
            if (edge) {
                if (t + 2 >= a.num_tiles) {
                    const unsigned long long avail64 = a.n - (unsigned long long)t * TILE;
                    avail = avail64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)avail64;
                }
                if (t == 0) {
                    if (prev_kind == PREV_NONE || prev_kind == PREV_LF) extra = 1;
                    else if (prev_kind == PREV_CR && buf[0] != '\n') extra = 1;
                }
            }
            // Exact line ends: '\n' ends a line; '\r' ends one unless a '\n' follows (Python
            // universal newlines); every other control character is content.  The scan's masks
            // hold CANDIDATES; they are exact as long as every candidate is a line feed.  The
            // first tile that has another one (a '\r', a tab, a byte >= 0xA0 ...) switches the warp
            // to checking every candidate (sticky: such files have them everywhere).
            if (classify_first | __any_sync(FULL, scan_dev != 0)) {
                classify_first = true;
#pragma unroll
                for (uint32_t j = 0; j < MWORDS; j++) {
                    uint32_t m = mk[j], keepm = m;
                    while (m) {
                        const uint32_t b = __ffs(m) - 1u;
                        m &= m - 1u;
                        const uint32_t p = lane * SPAN + 32 * j + b;
                        const uint32_t c = buf[p];
                        bool end = c == '\n';
                        // the byte after the last byte of the chunk is unknown: pending
                        if (c == '\r') end = (p + 1 < avail) && buf[p + 1] != '\n';
                        if (!end) keepm &= ~(1u << b);
                    }
                    mk[j] = keepm;
                }
                __syncwarp();
            }
            uint32_t cntA = 0, cntB = 0;          // line ends in the two halves of my span
#pragma unroll
            for (uint32_t j = 0; j < MWORDS; j++) {
                if (j < HALF_WORDS) cntA += __popc(mk[j]); else cntB += __popc(mk[j]);
            }
            const uint32_t cnt = cntA + cntB;

            // ---- ranks: inclusive prefix sum of cnt over the lanes.  Counts are small: one ballot
            // per bit of the count (independent of each other) instead of five dependent shuffles.
            uint32_t incl, total;
            if (!__any_sync(FULL, cnt >= 8u)) {
                const uint32_t upto = 0xFFFFFFFFu >> (31u - lane);          // lanes 0..lane
                const uint32_t b0 = __ballot_sync(FULL, (cnt & 1u) != 0), b1 = __ballot_sync(FULL, (cnt & 2u) != 0),
                               b2 = __ballot_sync(FULL, (cnt & 4u) != 0);
                incl = __popc(b0 & upto) + 2u * __popc(b1 & upto) + 4u * __popc(b2 & upto);
                total = extra + __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
            } else {
                incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t o = __shfl_up_sync(FULL, incl, d);
                    if (lane >= (uint32_t)d) incl += o;
                }
                total = extra + __shfl_sync(FULL, incl, 31);
            }
            uint32_t nlive = 0;
            if (MATCH) {
                const uint32_t rhoA = extra + incl - cnt;       // rank of the line start after my first line end
                if (need_guess) {
                    // First lines of a segment whose position in the file is not known yet:
                    // find a line that looks like a FASTQ header ('@', then '+' two lines on,
                    // sequence and quality lines of equal length).  Any answer is acceptable --
                    // a wrong one is found and repaired by verify_kernel + the fix pass.
                    if (rhoA < GUESS_LINES + 5) {
                        uint32_t r = rhoA;
#pragma unroll
                        for (uint32_t j = 0; j < MWORDS; j++) {
                            uint32_t m = mk[j];
                            while (m && r < GUESS_LINES + 5) {
                                const uint32_t b = __ffs(m) - 1u;
                                m &= m - 1u;
                                ws->gs[r] = (uint16_t)(lane * SPAN + 32 * j + b + 1);
                                r++;
                            }
                        }
                    }
                    __syncwarp();
                    const uint32_t have = total < GUESS_LINES + 5 ? total : GUESS_LINES + 5;
                    bool hit = false;
                    if (lane < GUESS_LINES && lane + 4 < have) {
                        uint32_t p0 = ws->gs[lane], p1 = ws->gs[lane + 1], p2 = ws->gs[lane + 2], p3 = ws->gs[lane + 3],
                                 p4 = ws->gs[lane + 4];
                        hit = buf[p0] == '@' && buf[p2] == '+' && p2 - p1 == p4 - p3;
                    }
                    const uint32_t hits = __ballot_sync(FULL, hit);
                    phase = hits ? ((4u - ((uint32_t)(__ffs(hits) - 1) & 3u)) & 3u) : 0u;   // that line has index 0 mod 4
                    need_guess = false;
                }
                // ---- emission: the starts of sequence lines (index % 4 == 1) ----------------------
                // rank 0 of this tile has index F = (first index of the segment) + seg_lines
                const uint32_t a4 = (1u - (phase + seg_lines)) & 3u;         // first rank that is a sequence line
                const uint32_t nq = total > a4 ? (total - a4 + 3u) >> 2 : 0u;
                nlive = nq;                                                  // those below the read limit (a prefix)
                if (has_limit) {                                             // only the fix pass applies a limit
                    const unsigned long long seg_first = a.fix[cur_item >> 1].true_first;   // index of the segment's first line start
                    const unsigned long long first_idx = (seg_first + seg_lines + a4) >> 2;
                    nlive = 0;
                    if (nq && first_idx < a.reads_limit) {
                        unsigned long long room = a.reads_limit - first_idx;
                        nlive = nq < room ? nq : (uint32_t)room;
                    }
                }
                // Each half of my span (HALF_WORDS mask words, then the rest; <= 128 bytes each) holds
                // at most one such start in ordinary FASTQ: candidate number `skip` of the half.
                const uint32_t rhoB = rhoA + cntA;
                const uint32_t skipA = (a4 - rhoA) & 3u, skipB = (a4 - rhoB) & 3u;
                // The common tile: the middle of a chunk, no read limit, no half with two starts.
                const bool fast = !(edge | has_limit | (nq > QCAP)) &&
                                  !__any_sync(FULL, (cntA > skipA + 4u) | (cntB > skipB + 4u));
                if (fast) {
                    // candidate number `skip` of a half, found by counting through its mask words
                    uint32_t mA = 0, pA = 0, rA = skipA, seen = 0;
#pragma unroll
                    for (uint32_t j = 0; j < HALF_WORDS; j++) {
                        const uint32_t w = mk[j];
                        if (skipA >= seen) { mA = w; pA = 32u * j; rA = skipA - seen; }
                        seen += __popc(w);
                    }
                    uint32_t mB = 0, pB = 0, rB = skipB;
                    seen = 0;
#pragma unroll
                    for (uint32_t j = HALF_WORDS; j < MWORDS; j++) {
                        const uint32_t w = mk[j];
                        if (skipB >= seen) { mB = w; pB = 32u * j; rB = skipB - seen; }
                        seen += __popc(w);
                    }
                    if (rA >= 1u) mA &= mA - 1u;
                    if (rA >= 2u) mA &= mA - 1u;
                    if (rA >= 3u) mA &= mA - 1u;
                    if (rB >= 1u) mB &= mB - 1u;
                    if (rB >= 2u) mB &= mB - 1u;
                    if (rB >= 3u) mB &= mB - 1u;
                    const uint32_t mine = sbase + lane * SPAN;
                    if (cntA > skipA) ws->q[(rhoA + skipA - a4) >> 2] = (uint16_t)(mine + pA + __ffs(mA));
                    if (cntB > skipB) ws->q[(rhoB + skipB - a4) >> 2] = (uint16_t)(mine + pB + __ffs(mB));
                }
                // ---- batches of up to 32 starts, one per lane.  The common tile has one (about
                // 30 reads of 250 bytes), whose probe loads stay in flight until the next tile's
                // scan is half done; the loop is left through a forward branch (ptxas waits for
                // everything outstanding at loop headers).
                for (uint32_t w0 = 0; w0 < nlive;) {
                    const uint32_t room = nlive - w0 < 32u ? nlive - w0 : 32u;
                    if (!fast) {
                        // any other tile (chunk edges, very short lines, the read limit of the fix
                        // pass): walk my candidates; ordinals w0 .. w0+room go to q[0 .. room)
                        const uint32_t skip0 = skipA, jj0 = (rhoA + skipA - a4) >> 2;
                        uint32_t skip = skip0, jj = jj0;
#pragma unroll
                        for (uint32_t j = 0; j < MWORDS; j++) {
                            uint32_t m = mk[j];
                            while (m) {
                                const uint32_t b = __ffs(m) - 1u;
                                m &= m - 1u;
                                if (skip == 0) {
                                    if (jj - w0 < room) ws->q[jj - w0] = (uint16_t)(sbase + lane * SPAN + 32 * j + b + 1);
                                    jj++;
                                    skip = 3;
                                } else {
                                    skip--;
                                }
                            }
                        }
                        // the line that starts with the chunk's first byte
                        if (extra != 0 && lane == 0 && a4 == 0 && w0 == 0) ws->q[0] = (uint16_t)sbase;
                    }
                    __syncwarp();
                    if (pb_pending) batch_back();
                    batch_front(fast ? w0 : 0u, room, s);
                    w0 += room;
                    if (w0 >= nlive) break;
                }
            }
            seg_lines += total;
            seg_reads += nlive;

            if (seg_end) {                   // uniform, once per segment: nothing may stay in flight
                if (MATCH) {
                    if (pb_pending) batch_back();
                    red_retire();
                }
                if (lane == 0) {
                    // reads numbered by this segment (signed: the fix pass subtracts)
                    if (MATCH && seg_reads) atomicAdd(&a.totals[0], (unsigned long long)(weight * (long long)seg_reads));
                    if (a.mode == MODE_MAIN) {
                        SegInfo si;
                        si.lines = seg_lines;
                        si.guess = phase;
                        a.seginfo[cur_item] = si;          // (main pass: the work item is the segment)
                    }
                }
                seg_reads = 0;
            }

            // ---- refill this tile's stage (nothing points into it any more) --------------------
            __syncwarp();
            if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            cur_tile = nxt_tile[0];
            cur_item = nxt_item[0];
            cur_tixf = nxt_tixf[0];
#pragma unroll
            for (int k = 0; k + 1 < STAGES - 1; k++) {
                nxt_tile[k] = nxt_tile[k + 1];
                nxt_item[k] = nxt_item[k + 1];
                nxt_tixf[k] = nxt_tixf[k + 1];
            }
            produce(s, nxt_tile[STAGES - 2], nxt_item[STAGES - 2], nxt_tixf[STAGES - 2]);
            if (++s == STAGES) { s = 0; parity ^= 1u; }
        }

        // ---- open the next tile ------------------------------------------------------
        if (cur_item == NONE) break;
        if (!mbar_test(&full_bar[warp][s], parity)) mbar_wait(&full_bar[warp][s], parity);   // usually there already

        if ((cur_tixf & ~TIX_LAST) == 0) {     // a new segment starts
            seg_lines = 0;
            has_limit = false;
            weight = 1;
            need_guess = false;
            if (a.mode == MODE_FIX) {
                const FixEntry fe = a.fix[cur_item >> 1];
                if (cur_item & 1u) { phase = (uint32_t)fe.true_first & 3u; has_limit = true; }
                else               { phase = fe.guess & 3u; weight = -1; }
            } else if (cur_item == 0) {
                // known exactly
                phase = (uint32_t)(a.use_arg_state ? a.line_base : a.state_in->next_line) & 3u;
            } else {
                phase = 0;
                need_guess = MATCH;
            }
            asm volatile("" ::"r"(phase));     // a loaded value: see produce
        }

        // ---- scan: control-character mask of my SPAN bytes ----------------------
        // (lane l reads 16-byte units CHUNKS*l + i: with CHUNKS odd, eight consecutive
        // lanes hit eight different bank groups, so every 128-bit load is conflict free)
        // The pending batch is finished in the middle of the scan: its probe loads went out before
        // the refill and have had half a scan more to come back, and the vote (MATCH.ANY) issued
        // at the end of batch_back has the other half to finish before red_retire.
        {
            const uint4 *src = (const uint4 *)(wbase + s * STAGE + lane * SPAN);
            scan_dev = 0;
#pragma unroll
            for (uint32_t i = 0; i < CHUNKS; i++) {
                if (MATCH && i == (CHUNKS + 1) / 2) {
                    if (pb_pending) batch_back();
                }
                uint32_t m16 = ctl_mask16(src[i], scan_dev);
                if (i & 1u) mk[i >> 1] |= m16 << 16; else mk[i >> 1] = m16;
            }
        }
        if (cur_tile == a.num_tiles - 1) {
            // The line that would start right after the last byte of the chunk is
            // numbered by the NEXT chunk (PREV_LF), and bytes at and after n do not
            // exist: keep line ends at p < valid - 1 only.
            const unsigned long long avail64 = a.n - (unsigned long long)cur_tile * TILE;
            const uint32_t valid = avail64 < TILE ? (uint32_t)avail64 : TILE;
            const uint32_t lim = valid - 1;
            const uint32_t first = lane * SPAN;
            const uint32_t keep = lim > first ? lim - first : 0;
#pragma unroll
            for (uint32_t j = 0; j < MWORDS; j++) {
                uint32_t nb = keep > 32 * j ? keep - 32 * j : 0;
                if (nb < 32) mk[j] &= (1u << nb) - 1u;
            }
            if (lane == 0) {
                uint32_t c = (wbase + s * STAGE)[valid - 1];
                *a.last_kind = c == '\n' ? PREV_LF : (c == '\r' ? PREV_CR : PREV_OTHER);
            }
        }
        if (MATCH) red_retire();
        opened = true;
    }

    // totals: one set of atomics per warp
    if (MATCH) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_bar += __shfl_xor_sync(FULL, my_bar, o);
            my_tag += __shfl_xor_sync(FULL, my_tag, o);
        }
    }
    if (lane == 0) {
        if (MATCH) {
            if (my_bar) atomicAdd(&a.totals[1], (unsigned long long)(long long)my_bar);
            if (my_tag) atomicAdd(&a.totals[2], (unsigned long long)(long long)my_tag);
        }
    }
}

// Adds the extra copies of a small count matrix into the matrix and clears them.
__global__ void __launch_bounds__(256) fold_kernel(int32_t *matrix, int32_t *replicas, uint32_t cells, uint32_t extra)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        int32_t sum = 0;
        for (uint32_t r = 0; r < extra; r++) {
            int32_t v = replicas[(size_t)r * cells + i];
            if (v) { sum += v; replicas[(size_t)r * cells + i] = 0; }
        }
        if (sum) matrix[i] += sum;
    }
}

// Smallest cell of the count matrix (overflow guard: counts only grow, so a negative cell is a
// cell that passed INT32_MAX; the reference counts with unbounded Python integers).
__global__ void __launch_bounds__(256) min_kernel(const int32_t *matrix, uint32_t cells, int32_t *out)
{
    int32_t m = 0x7FFFFFFF;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        const int32_t v = matrix[i];
        m = v < m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t v = __shfl_xor_sync(0xFFFFFFFFu, m, o);
        m = v < m ? v : m;
    }
    if ((threadIdx.x & 31u) == 0) atomicMin(out, m);
}

// Prefix sum over the per-segment line counts, next chunk's state, and the list of segments the
// fix pass must redo.  Two launches of a few CTAs: every warp takes a contiguous run of segments
// and walks it 32 entries at a time (coalesced loads).  verify_sums leaves the runs' totals in
// `run_total`; verify_kernel adds up the totals of the runs before its own and gives every segment
// its true first line index by a warp scan per round.  (One CTA doing both took 70 us for the
// 170 k segments of a 200 M-read launch -- 5 % of an eighth of that launch.)
constexpr int VERIFY_THREADS = 1024;
constexpr uint32_t VERIFY_MAX_CTAS = 32;
__device__ __forceinline__ void verify_run(const VerifyArgs &v, uint32_t &lo, uint32_t &hi)
{
    const uint32_t nruns = gridDim.x * (VERIFY_THREADS / 32);
    const uint32_t run = blockIdx.x * (VERIFY_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t per = ((v.num_segs + nruns - 1) / nruns + 31u) & ~31u;    // segments per run: whole rounds
    lo = (unsigned long long)run * per < v.num_segs ? run * per : v.num_segs;
    hi = (unsigned long long)lo + per < v.num_segs ? lo + per : v.num_segs;
}

__global__ void __launch_bounds__(VERIFY_THREADS) verify_sums(const VerifyArgs v)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t lo, hi;
    verify_run(v, lo, hi);
    unsigned long long sum = 0;
    for (uint32_t i = lo + lane; i < hi; i += 128) {
        uint32_t part[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) part[k] = i + 32 * k < hi ? __ldg(&v.seginfo[i + 32 * k].lines) : 0u;
        sum += (unsigned long long)part[0] + part[1] + part[2] + part[3];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) v.run_total[blockIdx.x * (VERIFY_THREADS / 32) + (threadIdx.x >> 5)] = sum;
}

__global__ void __launch_bounds__(VERIFY_THREADS) verify_kernel(const VerifyArgs v)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const unsigned long long line_base = v.use_arg_state ? v.line_base : v.state_in->next_line;
    const uint32_t nruns = gridDim.x * (VERIFY_THREADS / 32);
    const uint32_t run = blockIdx.x * (VERIFY_THREADS / 32) + (tid >> 5);
    uint32_t lo, hi;
    verify_run(v, lo, hi);

    unsigned long long before = 0, total = 0;                      // lines in the runs before mine / in all
    for (uint32_t w = lane; w < nruns; w += 32) {
        const unsigned long long t = v.run_total[w];
        if (w < run) before += t;
        total += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
        total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
    }

    if (v.make_fixes) {
        unsigned long long first = line_base + before;             // true first index of the round's first segment
        for (uint32_t i0 = lo; i0 < hi; i0 += 128) {
            // four rounds' loads go out together (a round's scan must not wait a memory latency)
            uint2 part[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t i = i0 + 32 * k + lane;
                part[k] = i < hi ? __ldg((const uint2 *)&v.seginfo[i]) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t i = i0 + 32 * k + lane;
                const uint32_t lines = part[k].x, guess = part[k].y;
                uint32_t incl = lines;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += o;
                }
                const unsigned long long mine = first + (incl - lines);
                const bool wrong = ((mine ^ guess) & 3ull) != 0;
                // could any read of this segment reach the limit?
                const bool past = lines && ((mine + lines - 1) >> 2) >= v.reads_limit;
                if (lines && (wrong || past)) {
                    const uint32_t slot = atomicAdd(v.n_fix, 1u);
                    FixEntry fe;
                    fe.seg = i;
                    fe.guess = guess;
                    fe.true_first = mine;
                    v.fix[slot] = fe;
                }
                first += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        v.state_out->next_line = line_base + total;
        v.state_out->prev_kind = *v.last_kind;
        v.state_out->pad = 0;
    }
}

#endif  // __CUDACC__

}  // namespace tdg
