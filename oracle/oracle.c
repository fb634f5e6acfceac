/*
 * C restatement of TagDigger's per-read counting loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the fast half of the oracle (see oracle/tagdigger_oracle.py for
 * the header that explains how the oracle is pinned to the reference).  It is
 * linked only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg; nothing under tagdigger_b200/ may load it.
 *
 * It follows, in /root/reference/tagdigger_fun.py:
 *   :71-113   trie build and its conflict rules      -> orc_trie_build
 *   :115-134  per-base trie walk                      -> orc_trie_lookup
 *   :245-277  line loop, lineindex % 4 == 1, strip().upper(), two lookups,
 *             counts[bar][tag] += 1, maxreads         -> orc_count
 * with Python's text-mode universal newlines ("\n", "\r\n", lone "\r") and
 * str.strip() whitespace (ASCII 0x09-0x0d, 0x1c-0x1f, 0x20, plus the Unicode
 * whitespace characters in their UTF-8 encoding) restated by hand.  Other bytes
 * >= 0x80 are ordinary non-ACGT characters (input is assumed to be valid UTF-8,
 * which the reference needs too: it decodes the file in text mode).
 *
 * Before it is trusted it is compared with the Python oracle and the golden
 * vectors in tests/test_oracle_c.py.
 */

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t *kids;      /* 4 per node; 0 = absent */
    int32_t *leaf;      /* stored index or -1 */
    int32_t  nnodes, cap;
    int      any_base;      /* the single-empty-pattern tree, :109-110 */
    int      root_is_leaf;  /* first of several patterns is empty */
} otrie;

static int base_code(uint8_t c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default:  return -1;
    }
}

static int32_t new_node(otrie *t)
{
    if (t->nnodes == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->kids = (int32_t *)realloc(t->kids, sizeof(int32_t) * 4 * (size_t)t->cap);
        t->leaf = (int32_t *)realloc(t->leaf, sizeof(int32_t) * (size_t)t->cap);
    }
    int32_t n = t->nnodes++;
    t->kids[4 * n] = t->kids[4 * n + 1] = t->kids[4 * n + 2] = t->kids[4 * n + 3] = 0;
    t->leaf[n] = -1;
    return n;
}

void orc_trie_free(otrie *t)
{
    if (!t) return;
    free(t->kids);
    free(t->leaf);
    free(t);
}

/* returns 0 ok; 1 = AssertionError (*problem = index printed by the reference);
 * 2 = IndexError (empty pattern list). */
int orc_trie_build(const char *chars, const uint64_t *off, uint32_t n, uint32_t numseq,
                   otrie **out, int64_t *problem)
{
    *out = NULL;
    otrie *t = (otrie *)calloc(1, sizeof(otrie));
    new_node(t);
    if (numseq == 1 && n == 1 && off[1] == off[0]) {
        t->any_base = 1;
        *out = t;
        return 0;
    }
    if (n == 0) { orc_trie_free(t); return 2; }

    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * n);
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) perm[i] = i;
    /* explicit stack of (node, depth, begin, end) ranges of perm */
    size_t scap = 1024, sp = 0;
    uint64_t *stack = (uint64_t *)malloc(sizeof(uint64_t) * 4 * scap);
#define PUSH(a, b, c, d) do { if (sp == scap) { scap *= 2; stack = (uint64_t *)realloc(stack, sizeof(uint64_t) * 4 * scap); } \
        stack[4 * sp] = (a); stack[4 * sp + 1] = (b); stack[4 * sp + 2] = (c); stack[4 * sp + 3] = (d); sp++; } while (0)
    PUSH(0, 0, 0, n);
    int rc = 0;
    while (sp) {
        sp--;
        int32_t node = (int32_t)stack[4 * sp];
        uint64_t depth = stack[4 * sp + 1];
        uint32_t b = (uint32_t)stack[4 * sp + 2], e = (uint32_t)stack[4 * sp + 3];
        uint32_t first = perm[b];
        if (off[first + 1] - off[first] == depth) {        /* :76-77 */
            t->leaf[node] = (int32_t)(first % numseq);
            if (node == 0) t->root_is_leaf = 1;
            continue;
        }
        uint32_t cnt[4] = {0, 0, 0, 0};
        for (uint32_t i = b; i < e; i++) {                 /* :81-85 */
            uint32_t m = perm[i];
            if (off[m + 1] - off[m] <= depth) { rc = 1; *problem = m % numseq; goto done; }
            int c = base_code((uint8_t)chars[off[m] + depth]);
            if (c < 0) c = 3;                              /* "ACGT".find -> -1 -> last bucket */
            cnt[c]++;
        }
        uint32_t start[4], pos[4];
        start[0] = b;
        for (int c = 1; c < 4; c++) start[c] = start[c - 1] + cnt[c - 1];
        memcpy(pos, start, sizeof(pos));
        for (uint32_t i = b; i < e; i++) {
            uint32_t m = perm[i];
            int c = base_code((uint8_t)chars[off[m] + depth]);
            if (c < 0) c = 3;
            tmp[pos[c]++] = m;
        }
        memcpy(perm + b, tmp + b, sizeof(uint32_t) * (e - b));
        /* children visited A,C,G,T: push in reverse */
        int32_t child[4] = {0, 0, 0, 0};
        for (int c = 0; c < 4; c++)
            if (cnt[c]) { child[c] = new_node(t); t->kids[4 * node + c] = child[c]; }
        for (int c = 3; c >= 0; c--)
            if (cnt[c]) PUSH((uint64_t)child[c], depth + 1, start[c], start[c] + cnt[c]);
    }
done:
    free(perm); free(tmp); free(stack);
    if (rc) { orc_trie_free(t); return rc; }
    *out = t;
    return 0;
}

/* -1 no match; -2 / -3: the reference would raise IndexError / TypeError */
int32_t orc_trie_lookup(const otrie *t, const uint8_t *s, size_t len)
{
    if (t->any_base) return (len > 0 && base_code(s[0]) >= 0) ? 0 : -1;
    int32_t node = 0;
    for (size_t i = 0; i < len; i++) {
        int c = base_code(s[i]);
        if (c < 0) return -1;
        if (t->root_is_leaf) return c == 1 ? -3 : -2;
        node = t->kids[4 * node + c];
        if (node == 0) return -1;
        if (t->leaf[node] >= 0) return t->leaf[node];
    }
    return -1;
}

static int is_space(uint8_t c)
{
    return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20);
}

/* Length of the UTF-8 encoded Unicode whitespace character at s[0..avail), or 0.
 * str.strip() removes these too once the file has been decoded as UTF-8:
 * U+0085 U+00A0 U+1680 U+2000..U+200A U+2028 U+2029 U+202F U+205F U+3000. */
static size_t uni_space(const uint8_t *s, size_t avail)
{
    if (avail >= 2 && s[0] == 0xC2 && (s[1] == 0x85 || s[1] == 0xA0)) return 2;
    if (avail < 3) return 0;
    if (s[0] == 0xE1 && s[1] == 0x9A && s[2] == 0x80) return 3;
    if (s[0] == 0xE2 && s[1] == 0x80 &&
        ((s[2] >= 0x80 && s[2] <= 0x8A) || s[2] == 0xA8 || s[2] == 0xA9 || s[2] == 0xAF)) return 3;
    if (s[0] == 0xE2 && s[1] == 0x81 && s[2] == 0x9F) return 3;
    if (s[0] == 0xE3 && s[1] == 0x80 && s[2] == 0x80) return 3;
    return 0;
}

/* Counts reads of one FASTQ byte image.  `first_line` is the index of the line
 * that starts at bytes[0] (0 at file start; lets callers shard a file at line
 * boundaries).  `reads_before` is how many sequence lines precede this image.
 * returns 0 ok, -2/-3 as orc_trie_lookup. */
int orc_count(const uint8_t *bytes, size_t n, const otrie *bar, const otrie *tag,
              const uint32_t *offsets, uint32_t ntags, double maxreads,
              uint64_t first_line, uint64_t reads_before,
              int64_t *counts, uint64_t totals[3])
{
    size_t cap = 1024;
    uint8_t *buf = (uint8_t *)malloc(cap);
    uint64_t line = first_line, reads = reads_before, nbar = 0, ntag = 0;
    size_t i = 0;
    int rc = 0;
    while (i < n) {
        size_t e = i;
        while (e < n && bytes[e] != '\n' && bytes[e] != '\r') e++;
        if ((line & 3) == 1) {
            reads++;
            size_t a = i, b = e;
            for (;;) {                                     /* str.strip(), :256 */
                size_t u;
                if (a < b && is_space(bytes[a])) a++;
                else if (a < b && (u = uni_space(bytes + a, b - a)) != 0) a += u;
                else break;
            }
            for (;;) {
                if (b > a && is_space(bytes[b - 1])) b--;
                else if (b >= a + 2 && uni_space(bytes + b - 2, 2) == 2) b -= 2;
                else if (b >= a + 3 && uni_space(bytes + b - 3, 3) == 3) b -= 3;
                else break;
            }
            size_t L = b - a;
            if (L > cap) { cap = L * 2; buf = (uint8_t *)realloc(buf, cap); }
            for (size_t k = 0; k < L; k++) {
                uint8_t c = bytes[a + k];
                buf[k] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
            }
            int32_t bi = orc_trie_lookup(bar, buf, L);
            if (bi < -1) { rc = bi; break; }
            if (bi >= 0) {
                nbar++;
                size_t o = offsets[bi];
                int32_t ti = (o <= L) ? orc_trie_lookup(tag, buf + o, L - o) : -1;
                if (ti < -1) { rc = ti; break; }
                if (ti >= 0) { ntag++; counts[(size_t)bi * ntags + (size_t)ti]++; }
            }
            if ((double)reads >= maxreads) break;
        }
        line++;
        if (e < n) {
            if (bytes[e] == '\r' && e + 1 < n && bytes[e + 1] == '\n') e++;
            e++;
        }
        i = e;
    }
    free(buf);
    totals[0] = reads - reads_before; totals[1] = nbar; totals[2] = ntag;
    return rc;
}

/* Number of text lines Python would iterate in this image. */
uint64_t orc_count_lines(const uint8_t *bytes, size_t n)
{
    uint64_t lines = 0;
    size_t i = 0;
    while (i < n) {
        size_t e = i;
        while (e < n && bytes[e] != '\n' && bytes[e] != '\r') e++;
        lines++;
        if (e < n) {
            if (bytes[e] == '\r' && e + 1 < n && bytes[e + 1] == '\n') e++;
            e++;
        }
        i = e;
    }
    return lines;
}
