// TEST HARNESS, not product code: compiles the product's table builders and per-read
// match (tagdigger_b200/csrc/tdg_tables.h, tdg_match.h -- the very code the CUDA kernel
// inlines) for the host, so that tests can look single reads up in the packed tables
// without a GPU and compare with the oracle.  Built into tests/native/libtable_check.so
// by tests/table_check.py; nothing under tagdigger_b200/ links or loads it.
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../tagdigger_b200/csrc/tdg_tables.h"

struct tck {
    tdg::HostTagTable tags;
    std::vector<uint8_t> bar;
    uint32_t cols = 0;
    std::string err;
    bool have_tags = false, have_bar = false;
};

extern "C" {

tck *tck_create(void) { return new (std::nothrow) tck(); }
void tck_destroy(tck *t) { delete t; }
const char *tck_error(const tck *t) { return t->err.c_str(); }

int tck_set_tags(tck *t, const char *bases, const uint64_t *off, const int32_t *col, uint32_t n, uint32_t flags, uint32_t cols)
{
    t->have_tags = false;
    t->err = tdg::build_tag_table(bases, off, col, n, flags, t->tags);
    if (!t->err.empty()) return -2;
    t->cols = cols;
    t->have_tags = true;
    return 0;
}

int tck_set_bars(tck *t, const char *bases, const uint32_t *off, const int32_t *row, const uint32_t *tag_off, uint32_t n,
                 uint32_t flags)
{
    t->have_bar = false;
    t->err = tdg::build_bar_table(bases, off, row, tag_off, n, flags, t->bar);
    if (!t->err.empty()) return -2;
    t->have_bar = true;
    return 0;
}

// matrix cell (row * cols + col), -1 = barcode but no tag, -2 = no barcode, -3 = tables missing
long long tck_match(tck *t, const char *read, size_t len)
{
    if (!t->have_tags || !t->have_bar) return -3;
    const uint8_t *p = (const uint8_t *)read;
    size_t pos = 0;
    while (pos < len) {
        uint32_t c = p[pos];
        if (tdg::is_lead_space(c)) { pos++; continue; }
        if (c >= 0xC2 && c <= 0xE3 && pos + 2 < len) {
            uint32_t u = tdg::utf8_space(c, p[pos + 1], p[pos + 2]);
            if (u) { pos += u; continue; }
        }
        break;
    }
    tdg::HostFetch f;
    f.p = p + pos;
    f.limit = (uint32_t)(len - pos);
    tdg::TagTable tt = t->tags.t;
    tt.entries = t->tags.entries.data();
    tt.ext = t->tags.ext.data();
    const tdg::BarTable *bar = (const tdg::BarTable *)t->bar.data();
    const tdg::BarEntry *bent = (const tdg::BarEntry *)(t->bar.data() + sizeof(tdg::BarTable));
    tdg::MatchResult r = tdg::match_line(f, bar, bent, tt);
    if (r.row < 0) return -2;
    if (r.col < 0) return -1;
    return (long long)r.row * t->cols + r.col;
}

// how many slot pairs of the tag table carry the "probe sequence continues" flag, and the table size
void tck_table_stats(const tck *t, uint64_t out[3])
{
    uint64_t more = 0, used = 0;
    for (const auto &e : t->tags.entries) {
        if (e.len == TDG_EMPTY_LEN) continue;
        used++;
        if (e.len & TDG_LEN_MORE) more++;
    }
    out[0] = t->tags.entries.size();
    out[1] = used;
    out[2] = more;
}

}  // extern "C"
