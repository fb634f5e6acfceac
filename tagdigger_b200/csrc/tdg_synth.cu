// Synthetic GBS FASTQ generator on the device (bench / test support, not part
// of the counting path).  Produces, for reads [first_read, first_read+nreads),
// the deterministic byte image described in SURVEY.md section 8(d): variable
// length headers, barcode + cut site + known tag + random tail reads, reads
// with unknown inserts, reads without a barcode, near misses, N's, lower-case
// reads, short reads and quality lines that begin with '@' or '+'.  Every byte
// is a pure function of (seed, read number, byte position), so any shard of
// the job can be generated on any GPU.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <stdint.h>

#include <string>

namespace {

struct Params {
    uint64_t seed;
    uint32_t nbar, ntags;
    uint32_t bar_stride, tag_stride;   // bytes per row of the barcode / tag matrices
    uint32_t readlen, cutlen;
    uint32_t p_hit, p_unknown, p_nobar;   // cumulative 32-bit thresholds: kind 0,1,2 (else 3 = near miss)
    uint32_t p_n, p_lower, p_short, p_qual_at;
};

struct Tables {
    const uint8_t *bar;       // [nbar][bar_stride]
    const uint32_t *bar_len;  // [nbar]
    const uint8_t *tag;       // [ntags][tag_stride]  (tags include the cut site)
    const uint32_t *tag_len;  // [ntags]
    const uint32_t *tag_cdf;  // [ntags] cumulative popularity thresholds
    uint8_t cut[16];
};

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}

struct Rec {
    uint32_t kind, bidx, tidx, hl, seqlen;
    uint32_t bl, tl, span;
    uint32_t has_n, npos, lower, nm_pos, nm_sub, qa, qa_char;
    uint32_t salt;
    bool counted;      // must be counted by construction
};

__device__ inline Rec make_rec(const Params &P, const Tables &T, uint64_t r)
{
    Rec c;
    uint64_t h0 = mix64(P.seed ^ (r * 0xD1342543DE82EF95ull));
    uint64_t h1 = mix64(h0), h2 = mix64(h1), h3 = mix64(h2), h4 = mix64(h3);
    uint32_t u = (uint32_t)h0;
    c.kind = u < P.p_hit ? 0 : u < P.p_unknown ? 1 : u < P.p_nobar ? 2 : 3;
    c.bidx = (uint32_t)(h0 >> 32) % P.nbar;
    // tag by popularity: first index with cdf >= draw
    uint32_t d = (uint32_t)h1;
    uint32_t lo = 0, hi = P.ntags - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (T.tag_cdf[mid] >= d) hi = mid; else lo = mid + 1;
    }
    c.tidx = lo;
    c.hl = 30 + (uint32_t)(h1 >> 32) % 31;
    c.seqlen = P.readlen;
    bool is_short = (uint32_t)h2 < P.p_short;
    if (is_short) {
        uint32_t k = (uint32_t)(h2 >> 32) % 3;
        c.seqlen = k == 0 ? 3 : k == 1 ? 11 : 19;
    }
    c.has_n = (uint32_t)h3 < P.p_n;
    c.npos = (uint32_t)(h3 >> 32) % P.readlen;
    c.lower = (uint32_t)h4 < P.p_lower;
    c.bl = T.bar_len[c.bidx];
    c.tl = T.tag_len[c.tidx];
    c.span = c.bl + c.tl;                     // tags include the cut site
    uint64_t h5 = mix64(h4);
    uint32_t lim = c.span < P.readlen ? c.span : P.readlen;
    c.nm_pos = (uint32_t)h5 % lim;
    c.nm_sub = 1 + (uint32_t)(h5 >> 32) % 3;
    uint64_t h6 = mix64(h5);
    c.qa = (uint32_t)h6 < P.p_qual_at;
    c.qa_char = (h6 >> 32) & 1 ? '@' : '+';
    c.salt = (uint32_t)(h6 >> 33) ^ (uint32_t)h0;
    bool mutated = (c.kind == 3) || (c.has_n && c.npos < c.span) || (c.seqlen < c.span) || (c.span > P.readlen);
    c.counted = c.kind == 0 && !mutated;
    return c;
}

__device__ inline uint32_t base_code_of(uint8_t ch)
{
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3;
}

__device__ inline uint8_t seq_char(const Params &P, const Tables &T, const Rec &c, uint32_t s)
{
    const char acgt[4] = {'A', 'C', 'G', 'T'};
    uint8_t ch = acgt[mix32(c.salt ^ (s * 0x9E3779B1u) ^ 0x51ED27u) & 3u];
    if (c.kind != 2) {
        if (s < c.bl) ch = T.bar[(size_t)c.bidx * P.bar_stride + s];
        else if (s < c.bl + P.cutlen) ch = T.cut[s - c.bl];
    }
    if (c.kind == 0 || c.kind == 3) {
        if (s >= c.bl && s - c.bl < c.tl) ch = T.tag[(size_t)c.tidx * P.tag_stride + (s - c.bl)];
    }
    if (c.kind == 3 && s == c.nm_pos) ch = acgt[(base_code_of(ch) + c.nm_sub) & 3u];
    if (c.has_n && s == c.npos) ch = 'N';
    if (c.lower) ch |= 0x20;
    return ch;
}

__global__ void lengths_kernel(Params P, Tables T, uint64_t first, uint64_t n, uint64_t *len, int32_t *expected)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rec c = make_rec(P, T, first + i);
    len[i] = c.hl + 2 * c.seqlen + 5;
    if (expected && c.counted) atomicAdd(&expected[(size_t)c.bidx * P.ntags + c.tidx], 1);
}

__global__ void fill_kernel(Params P, Tables T, uint64_t first, uint64_t n, const uint64_t *off, uint8_t *out)
{
    const char hdr_alpha[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789:";
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31u;
    if (w >= n) return;
    Rec c = make_rec(P, T, first + w);
    uint8_t *dst = out + off[w];
    uint32_t reclen = c.hl + 2 * c.seqlen + 5;
    uint32_t s0 = c.hl + 1, s1 = s0 + c.seqlen;      // sequence [s0, s1), then '\n', '+', '\n'
    uint32_t q0 = s1 + 3, q1 = q0 + c.seqlen;        // quality [q0, q1), then '\n'
    for (uint32_t j = lane; j < reclen; j += 32) {
        uint8_t ch;
        if (j < c.hl) ch = j == 0 ? '@' : hdr_alpha[mix32(c.salt ^ (j * 0x85EBCA77u)) % 37u];
        else if (j == c.hl) ch = '\n';
        else if (j < s1) ch = seq_char(P, T, c, j - s0);
        else if (j == s1) ch = '\n';
        else if (j == s1 + 1) ch = '+';
        else if (j == s1 + 2) ch = '\n';
        else if (j < q1) {
            uint32_t s = j - q0;
            ch = 35 + mix32(c.salt ^ (s * 0xC2B2AE3Du) ^ 0xA5A5A5u) % 40u;
            if (s == 0 && c.qa) ch = (uint8_t)c.qa_char;
        } else ch = '\n';
        dst[j] = ch;
    }
}

std::string g_err;

}  // namespace

extern "C" {

const char *tdgs_last_error(void) { return g_err.c_str(); }

// Generates reads [first, first+n) on `device`.  All table pointers are HOST
// pointers; `probs` holds the eight 32-bit thresholds in the order of Params.
// *d_out receives a cudaMalloc'ed image padded for tdg_count_device, *nbytes its
// length; d_expected (device int32 [nbar*ntags], may be NULL) is incremented for
// every read that must be counted by construction.
int tdgs_generate(int device, uint64_t seed, uint64_t first, uint64_t n, uint32_t readlen, const char *cutsite,
                  const uint8_t *bar, const uint32_t *bar_len, uint32_t nbar, uint32_t bar_stride,
                  const uint8_t *tag, const uint32_t *tag_len, const uint32_t *tag_cdf, uint32_t ntags,
                  uint32_t tag_stride, const uint32_t *probs, size_t pad_tile, size_t pad_halo, void **d_out,
                  uint64_t *nbytes, int32_t *d_expected)
{
#define CKS(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);       \
            return -1;                                                         \
        }                                                                      \
    } while (0)
    CKS(cudaSetDevice(device));
    Params P;
    P.seed = seed;
    P.nbar = nbar;
    P.ntags = ntags;
    P.bar_stride = bar_stride;
    P.tag_stride = tag_stride;
    P.readlen = readlen;
    P.cutlen = 0;
    Tables T;
    while (cutsite[P.cutlen] && P.cutlen < 16) { T.cut[P.cutlen] = (uint8_t)cutsite[P.cutlen]; P.cutlen++; }
    P.p_hit = probs[0]; P.p_unknown = probs[1]; P.p_nobar = probs[2];
    P.p_n = probs[3]; P.p_lower = probs[4]; P.p_short = probs[5]; P.p_qual_at = probs[6];
    uint8_t *dbar = nullptr, *dtag = nullptr;
    uint32_t *dbl = nullptr, *dtl = nullptr, *dcdf = nullptr;
    uint64_t *doff = nullptr, *dlen = nullptr;   // 64-bit lengths: the scan accumulates in its input type
    void *tmp = nullptr;
    CKS(cudaMalloc(&dbar, (size_t)nbar * bar_stride));
    CKS(cudaMalloc(&dtag, (size_t)ntags * tag_stride));
    CKS(cudaMalloc(&dbl, nbar * 4));
    CKS(cudaMalloc(&dtl, ntags * 4));
    CKS(cudaMalloc(&dcdf, ntags * 4));
    CKS(cudaMemcpy(dbar, bar, (size_t)nbar * bar_stride, cudaMemcpyHostToDevice));
    CKS(cudaMemcpy(dtag, tag, (size_t)ntags * tag_stride, cudaMemcpyHostToDevice));
    CKS(cudaMemcpy(dbl, bar_len, nbar * 4, cudaMemcpyHostToDevice));
    CKS(cudaMemcpy(dtl, tag_len, ntags * 4, cudaMemcpyHostToDevice));
    CKS(cudaMemcpy(dcdf, tag_cdf, ntags * 4, cudaMemcpyHostToDevice));
    T.bar = dbar; T.bar_len = dbl; T.tag = dtag; T.tag_len = dtl; T.tag_cdf = dcdf;
    CKS(cudaMalloc(&dlen, (n + 1) * sizeof(uint64_t)));
    CKS(cudaMalloc(&doff, (n + 1) * sizeof(uint64_t)));
    CKS(cudaMemset(dlen + n, 0, sizeof(uint64_t)));
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (n) lengths_kernel<<<blocks, 256>>>(P, T, first, n, dlen, d_expected);
    CKS(cudaGetLastError());
    size_t tmp_bytes = 0;
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, dlen, doff, n + 1));
    CKS(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    CKS(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, dlen, doff, n + 1));
    uint64_t total = 0;
    CKS(cudaMemcpy(&total, doff + n, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    size_t cap = (total + pad_tile - 1) / pad_tile * pad_tile + pad_halo;
    uint8_t *out = nullptr;
    CKS(cudaMalloc(&out, cap ? cap : 1));
    CKS(cudaMemset(out + total, '\n', cap - total));
    unsigned long long warps = n;
    unsigned fblocks = (unsigned)((warps * 32 + 255) / 256);
    if (n) fill_kernel<<<fblocks, 256>>>(P, T, first, n, doff, out);
    CKS(cudaGetLastError());
    CKS(cudaDeviceSynchronize());
    cudaFree(tmp); cudaFree(dlen); cudaFree(doff);
    cudaFree(dbar); cudaFree(dtag); cudaFree(dbl); cudaFree(dtl); cudaFree(dcdf);
    *d_out = out;
    *nbytes = total;
    return 0;
}

void tdgs_free(int device, void *p)
{
    cudaSetDevice(device);
    cudaFree(p);
}

}  // extern "C"
