// Host-side builders of the packed barcode and tag tables (plain C++, no CUDA).
//
// Input is the EFFECTIVE pattern set: what the reference's trie builder
// (build_sequence_tree, /root/reference/tagdigger_fun.py:71-113) leaves
// reachable, computed by tagdigger_b200/matchset.py.  Such a set is prefix-free,
// which is what makes exact-match hashing equivalent to the trie walk; the
// builders verify it and refuse anything else.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "tdg_match.h"

namespace tdg {

inline int base_code(char c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'T': return 2;
    case 'G': return 3;
    default:  return -1;
    }
}

// up to 32 bases of s[from .. from+count) packed little-endian, 2 bits each
inline uint64_t pack_bases(const char *s, size_t from, size_t count)
{
    uint64_t k = 0;
    for (size_t i = 0; i < count; i++) k |= (uint64_t)base_code(s[from + i]) << (2 * i);
    return k;
}

// Checks alphabet and prefix-freeness (which includes "no duplicates").
// Returns "" or a description of the problem.
template <class Off>
inline std::string check_pattern_set(const char *bases, const Off *off, uint32_t n, bool allow_empty)
{
    for (uint32_t i = 0; i < n; i++) {
        if (off[i + 1] < off[i]) return "offsets must be non-decreasing";
        if (!allow_empty && off[i + 1] == off[i]) return "empty pattern " + std::to_string(i);
        for (Off p = off[i]; p < off[i + 1]; p++)
            if (base_code(bases[p]) < 0)
                return "pattern " + std::to_string(i) + " has a character outside ACGT";
    }
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    auto less = [&](uint32_t a, uint32_t b) {
        size_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
        int c = memcmp(bases + off[a], bases + off[b], la < lb ? la : lb);
        if (c) return c < 0;
        return la < lb;
    };
    std::sort(order.begin(), order.end(), less);
    for (uint32_t i = 0; i + 1 < n; i++) {
        uint32_t a = order[i], b = order[i + 1];
        size_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
        if (la <= lb && memcmp(bases + off[a], bases + off[b], la) == 0)
            return "pattern " + std::to_string(a) + " is a prefix of pattern " + std::to_string(b) +
                   " (the set must be prefix-free)";
    }
    return "";
}

struct HostTagTable {
    std::vector<TagEntry> entries;
    std::vector<uint64_t> ext;
    TagTable t;        // entries/ext pointers are filled in by the owner (host or device)
};

inline std::string build_tag_table(const char *bases, const uint64_t *off, const int32_t *col,
                                   uint32_t ntags, uint32_t flags, HostTagTable &out)
{
    out.entries.clear();
    out.ext.clear();
    memset(&out.t, 0, sizeof(out.t));
    if (flags & 1u) {                       // TDG_ANY_BASE
        if (ntags != 1 || off[1] != off[0]) return "TDG_ANY_BASE needs exactly one empty tag";
        out.t.any_base = 1;
        out.t.any_col = col[0];
        return "";
    }
    if (ntags == 0) return "empty tag set";
    std::string why = check_pattern_set(bases, off, ntags, false);
    if (!why.empty()) return why;

    // Length classes: tags in a class are hashed on their first K bases, K = the
    // shortest length in the class (at most 32).  Typical tag sets (every tag at
    // least 12 bases) form a single class, i.e. one probe sequence per read.
    static const uint32_t lower[TDG_MAX_CLASSES] = {12, 6, 3, 1};
    std::vector<uint32_t> members[TDG_MAX_CLASSES];
    uint32_t minlen = 0xFFFFFFFFu, maxlen = 0;
    for (uint32_t i = 0; i < ntags; i++) {
        uint32_t L = (uint32_t)(off[i + 1] - off[i]);
        minlen = std::min(minlen, L);
        maxlen = std::max(maxlen, L);
        int c = 0;
        while (L < lower[c]) c++;
        members[c].push_back(i);
    }
    out.t.min_len = minlen;
    out.t.max_len = maxlen;
    uint32_t base = 0;
    for (int c = 0; c < TDG_MAX_CLASSES; c++) {
        if (members[c].empty()) continue;
        uint32_t K = 32;
        for (uint32_t i : members[c]) K = std::min<uint32_t>(K, (uint32_t)(off[i + 1] - off[i]));
        // load <= 1/8 while the table stays small next to the L2 (<= 32 MiB), else <= 1/4,
        // else <= 1/2: short probe sequences, almost always inside the first 64-byte line
        size_t n = members[c].size();
        size_t want = 8 * n * sizeof(TagEntry) <= ((size_t)32 << 20) ? 8 * n
                    : 4 * n * sizeof(TagEntry) <= ((size_t)64 << 20) ? 4 * n : 2 * n;
        uint32_t slots = 16;
        while (slots < want) slots <<= 1;
        TagClass &tc = out.t.cls[out.t.n_classes++];
        tc.K = K;
        tc.base = base;
        tc.mask = slots - 1;
        tc.pad = 0;
        out.entries.resize(base + slots);
        for (uint32_t s = 0; s < slots; s++) {
            TagEntry &e = out.entries[base + s];
            e.k0 = e.k1 = 0;
            e.len = TDG_EMPTY_LEN;
            e.col = -1;
            e.ext = 0;
            e.pad = 0;
        }
        for (uint32_t i : members[c]) {
            const char *s = bases + off[i];
            uint32_t L = (uint32_t)(off[i + 1] - off[i]);
            TagEntry e;
            e.k0 = pack_bases(s, 0, std::min<uint32_t>(L, 32));
            e.k1 = L > 32 ? pack_bases(s, 32, std::min<uint32_t>(L - 32, 32)) : 0;
            e.len = L;
            e.col = col[i];
            e.ext = 0;
            e.pad = 0;
            if (L > TDG_LEN_MASK) return "tag " + std::to_string(i) + " is too long";
            if (L > 64) {
                e.ext = (uint32_t)out.ext.size();
                for (uint32_t p = 64; p < L; p += 32) out.ext.push_back(pack_bases(s, p, std::min<uint32_t>(L - p, 32)));
            }
            const uint32_t home = tag_slot(e.k0 & lowmask(K), tc.mask);
            uint32_t h = home;
            while (out.entries[base + h].len != TDG_EMPTY_LEN) h = (h + 1) & tc.mask;
            out.entries[base + h] = e;
            // every slot pair the sequence crossed before it found room must tell lookups to go on
            for (uint32_t p2 = home; p2 != (h & ~1u); p2 = (p2 + 2) & tc.mask) out.entries[base + p2].len |= TDG_LEN_MORE;
        }
        base += slots;
    }
    // the long form of the fast matcher reads three words per candidate without looking at its length,
    // and index 0 for lanes that have no candidate
    for (int k = 0; k < 3; k++) out.ext.push_back(0);
    return "";
}

// BarTable header followed by the entries, as one byte blob.
inline std::string build_bar_table(const char *bases, const uint32_t *off, const int32_t *row,
                                   const uint32_t *tag_off, uint32_t npat, uint32_t flags,
                                   std::vector<uint8_t> &blob)
{
    BarTable hdr;
    memset(&hdr, 0, sizeof(hdr));
    std::vector<BarEntry> ents;
    if (flags & 1u) {                       // TDG_ANY_BASE
        if (npat != 1 || off[1] != off[0]) return "TDG_ANY_BASE needs exactly one empty pattern";
        hdr.any_base = 1;
        hdr.any_row = row[0];
        hdr.any_tag_off = tag_off[0];
        hdr.max_tag_off = tag_off[0];
    } else {
        if (npat == 0) return "empty barcode pattern set";
        std::string why = check_pattern_set(bases, off, npat, false);
        if (!why.empty()) return why;
        std::vector<std::vector<BarEntry>> buckets(256);
        for (uint32_t i = 0; i < npat; i++) {
            uint32_t L = off[i + 1] - off[i];
            if (L > 32) return "barcode+cutsite pattern " + std::to_string(i) + " is longer than 32 bases";
            if (tag_off[i] > 0xFFFFu) return "tag offset too large";
            BarEntry e;
            e.key = pack_bases(bases + off[i], 0, L);
            e.row = row[i];
            e.len = (uint16_t)L;
            e.tag_off = (uint16_t)tag_off[i];
            hdr.max_len = std::max(hdr.max_len, L);
            hdr.max_tag_off = std::max(hdr.max_tag_off, tag_off[i]);
            if (L >= 4) {
                buckets[e.key & 0xFF].push_back(e);
            } else {
                // a short pattern can start any bucket that extends it
                uint32_t fixed = (uint32_t)e.key, free_bits = 8 - 2 * L;
                for (uint32_t x = 0; x < (1u << free_bits); x++) buckets[fixed | (x << (2 * L))].push_back(e);
            }
        }
        for (int b = 0; b < 256; b++) {
            hdr.bucket[b] = (uint32_t)ents.size();
            ents.insert(ents.end(), buckets[b].begin(), buckets[b].end());
        }
        hdr.bucket[256] = (uint32_t)ents.size();
        hdr.n_entries = (uint32_t)ents.size();
        // Patterns of at most 16 bases leave the upper half of `key` free: it carries the 32-bit
        // compare mask for the kernel's fast matcher (every other reader masks the key to the
        // pattern's length, so the extra bits are invisible there).
        if (hdr.max_len <= 16)
            for (BarEntry &e : ents) e.key |= (uint64_t)(uint32_t)lowmask(e.len) << 32;
    }
    // two zero entries behind the last one: the fast matcher loads a bucket's first two entries
    // without looking at the bucket's size first
    ents.resize(ents.size() + 2, BarEntry{0, 0, 0, 0});
    blob.resize(sizeof(BarTable) + ents.size() * sizeof(BarEntry));
    memcpy(blob.data(), &hdr, sizeof(hdr));
    if (!ents.empty()) memcpy(blob.data() + sizeof(hdr), ents.data(), ents.size() * sizeof(BarEntry));
    return "";
}

// Byte-wise window for the host-side self test.
struct HostFetch {
    const uint8_t *p;
    uint32_t limit;
    TDG_HD void load8(uint32_t off, uint32_t w[8]) const
    {
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

}  // namespace tdg
