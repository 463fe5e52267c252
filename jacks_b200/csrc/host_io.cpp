// Host-side ingest and result formatting (no GPU work): the two places where the Python reference spends
// minutes at the Avana shape once the numerics are on the GPU.
//
//   jacks_table_scan / jacks_table_read   delimiter-separated count table -> float64 matrix of selected columns
//                                         + (offset, length) of every row's first field.  Replaces the per-cell
//                                         eval() of preprocess.py:136-149; cells that are not plain numbers come
//                                         back as NaN with a flag so the caller can eval() just those.
//   jacks_table_write                     rows of '%5e' / '%6e' formatted values with a label prefix, as
//                                         writeJacksWResults / writeFoldChanges produce them (jacks_io.py:91-156).
//   jacks_pseudo_draw                     the pseudo-gene draws of createPseudoNonessGenes (jacks_io.py:256-264) on
//                                         the caller's Mersenne-Twister state: 900,000 random.choice calls at 1e5
//                                         pseudo-genes, 1.2 s in the interpreter.
//
// Quoting follows Python's csv 'excel' dialect as far as count tables need it: a field may be wrapped in double
// quotes, inside which delimiters are literal and "" is an escaped quote.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/jacks_b200.h"
#include "capi_internal.h"

using namespace jacks;

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    int open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(JACKS_E_INVALID, "cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) return fail(JACKS_E_INVALID, "cannot stat %s", path);
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return JACKS_OK; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return fail(JACKS_E_NOMEM, "cannot map %s", path);
        p = (const char*)m;
        return JACKS_OK;
    }
    ~Mapped() {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

// end of the record starting at `b` (exclusive of the line terminator); *next = start of the next record
inline const char* record_end(const char* b, const char* end, const char** next) {
    // the usual record has no quote: its end is the first line terminator (two memchr sweeps instead of a byte loop)
    const char* q = b;
    {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(end - b));
        const char* lim = nl ? nl : end;
        if (!memchr(b, '"', (size_t)(lim - b))) {
            const char* cr = (const char*)memchr(b, '\r', (size_t)(lim - b));
            q = cr ? cr : lim;
            const char* e0 = q;
            if (q < end && *q == '\r') ++q;
            if (q < end && *q == '\n') ++q;
            *next = q;
            return e0;
        }
    }
    bool quoted = false;
    for (; q < end; ++q) {
        if (*q == '"') quoted = !quoted;
        else if (!quoted && (*q == '\n' || *q == '\r')) break;
    }
    const char* e = q;
    if (q < end && *q == '\r') ++q;
    if (q < end && *q == '\n') ++q;
    *next = q;
    return e;
}

}  // namespace

extern "C" int jacks_table_scan(const char* path, int64_t skip_lines, int64_t* n_records, int64_t* data_offset) {
    if (!path || !n_records || !data_offset) return fail(JACKS_E_INVALID, "jacks_table_scan: NULL argument");
    Mapped m;
    int rc = m.open(path);
    if (rc != JACKS_OK) return rc;
    const char* b = m.p;
    const char* end = m.p + m.n;
    const char* next = b;
    for (int64_t i = 0; i < skip_lines && b < end; ++i) { record_end(b, end, &next); b = next; }
    *data_offset = (int64_t)(b - m.p);
    int64_t cnt = 0;
    while (b < end) {
        const char* e = record_end(b, end, &next);
        if (e > b) ++cnt;                       // csv skips completely empty lines
        b = next;
    }
    *n_records = cnt;
    return JACKS_OK;
}

extern "C" int jacks_table_read(const char* path, int64_t data_offset, char delim, int64_t n_records, int32_t n_sel,
                                const int32_t* sel_cols, double* values, uint8_t* not_numeric, int32_t n_str,
                                const int32_t* str_cols, int64_t* label_off, int32_t* label_len) {
    if (!path || (n_sel > 0 && (!sel_cols || !values)) || n_str < 1 || !str_cols || !label_off || !label_len)
        return fail(JACKS_E_INVALID, "jacks_table_read: NULL argument");
    Mapped m;
    int rc = m.open(path);
    if (rc != JACKS_OK) return rc;
    int max_col = 0;
    for (int i = 0; i < n_sel; ++i) { if (sel_cols[i] < 0) return fail(JACKS_E_INVALID, "negative column index"); if (sel_cols[i] > max_col) max_col = sel_cols[i]; }
    std::vector<std::vector<int>> want(max_col + 1);          // column -> output slots (a column may be selected twice)
    for (int i = 0; i < n_sel; ++i) want[sel_cols[i]].push_back(i);
    int last_col = n_sel > 0 ? max_col : 0;                     // nothing beyond this column is asked for
    for (int k = 0; k < n_str; ++k) if (str_cols[k] > last_col) last_col = str_cols[k];
    const char* b = m.p + data_offset;
    const char* end = m.p + m.n;
    const char* next = b;
    // pass 1 (serial, a byte scan): record boundaries; pass 2 (threads): split and convert the fields of each record
    std::vector<std::pair<const char*, const char*>> recs;
    recs.reserve((size_t)n_records);
    while (b < end && (int64_t)recs.size() < n_records) {
        const char* e = record_end(b, end, &next);
        if (e != b) recs.emplace_back(b, e);
        b = next;
    }
    int64_t r = (int64_t)recs.size();
    auto parse_block = [&](int64_t r0, int64_t r1) {
        std::string tmp;
        for (int64_t r = r0; r < r1; ++r) {
            const char* b = recs[(size_t)r].first;
            const char* e = recs[(size_t)r].second;
            for (int i = 0; i < n_sel; ++i) { values[r * n_sel + i] = NAN; if (not_numeric) not_numeric[r * n_sel + i] = 1; }
            for (int k = 0; k < n_str; ++k) { label_off[r * n_str + k] = 0; label_len[r * n_str + k] = -1; }     // -1: field absent
            const char* f = b;
            int col = 0;
            while (f <= e) {
                const char* fe;
                const char* vb; const char* ve;
                if (f < e && *f == '"') {                         // quoted field
                    const char* q = f + 1;
                    while (q < e) { if (*q == '"') { if (q + 1 < e && q[1] == '"') { q += 2; continue; } break; } ++q; }
                    vb = f + 1; ve = q;
                    fe = (q < e) ? q + 1 : e;
                    while (fe < e && *fe != delim) ++fe;
                } else {
                    fe = f;
                    while (fe < e && *fe != delim) ++fe;
                    vb = f; ve = fe;
                }
                for (int k = 0; k < n_str; ++k)
                    if (col == str_cols[k]) { label_off[r * n_str + k] = (int64_t)(vb - m.p); label_len[r * n_str + k] = (int32_t)(ve - vb); }
                if (col <= max_col && !want[col].empty()) {
                    while (vb < ve && (*vb == ' ' || *vb == '\t')) ++vb;
                    while (ve > vb && (ve[-1] == ' ' || ve[-1] == '\t')) --ve;
                    // a run of at most 15 digits (every cell of an ordinary count table): exact in a double, no strtod
                    bool digits = (ve > vb) && (ve - vb <= 15);
                    unsigned long long iv = 0;
                    for (const char* c = vb; digits && c < ve; ++c) {
                        if (*c >= '0' && *c <= '9') iv = iv * 10 + (unsigned)(*c - '0');
                        else digits = false;
                    }
                    if (digits) {
                        for (int slot : want[col]) { values[r * n_sel + slot] = (double)iv; if (not_numeric) not_numeric[r * n_sel + slot] = 0; }
                    } else if (ve > vb) {
                        tmp.assign(vb, ve);
                        char* stop = nullptr;
                        const double v = strtod(tmp.c_str(), &stop);
                        // plain decimal literals only: Python's eval() reads "0x10", "1_000", "nan" differently from strtod
                        bool plain = (stop == tmp.c_str() + tmp.size());
                        for (char ch : tmp) if (!((ch >= '0' && ch <= '9') || ch == '.' || ch == '-' || ch == '+' || ch == 'e' || ch == 'E')) { plain = false; break; }
                        if (plain) for (int slot : want[col]) { values[r * n_sel + slot] = v; if (not_numeric) not_numeric[r * n_sel + slot] = 0; }
                    }
                }
                if (col == last_col) break;                       // (reading guide -> gene pairs off a count table stops at column 1)
                ++col;
                if (fe >= e) break;
                f = fe + 1;
            }
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)std::max(1u, std::min(16u, hw ? hw : 1u));
    if (m.n < (4u << 20) || r < 1024) n_thr = 1;
    if (n_thr == 1) parse_block(0, r);
    else {
        std::vector<std::thread> pool;
        const int64_t per = (r + n_thr - 1) / n_thr;
        for (int t = 0; t < n_thr; ++t) {
            const int64_t r0 = per * t, r1 = std::min<int64_t>(r, r0 + per);
            if (r0 < r1) pool.emplace_back(parse_block, r0, r1);
        }
        for (auto& th : pool) th.join();
    }
    if (r != n_records) return fail(JACKS_E_INVALID, "%s: expected %lld records, found %lld", path, (long long)n_records, (long long)r);
    return JACKS_OK;
}

namespace {

// format rows [r0, r1) into `out`: the text goes straight into one buffer sized for the worst case (24 characters per
// value: sign, 17 characters of "d.dddddde+ddd", separators), no per-value appends
void format_rows(int64_t r0, int64_t r1, int32_t n_cols, const char* labels, const int64_t* label_off, const double* values,
                 const uint8_t* blank, int32_t width, std::string* out) {
    const size_t per_value = (size_t)(width > 24 ? width : 24) + 1;
    const size_t cap = (size_t)(label_off[r1] - label_off[r0]) + (size_t)(r1 - r0) * ((size_t)n_cols * per_value + 1) + 64;
    std::unique_ptr<char[]> buf(new char[cap]);
    char* p = buf.get();
    for (int64_t r = r0; r < r1; ++r) {
        const size_t ll = (size_t)(label_off[r + 1] - label_off[r]);
        memcpy(p, labels + label_off[r], ll);
        p += ll;
        for (int c = 0; c < n_cols; ++c) {
            *p++ = '\t';
            if (blank && blank[r * n_cols + c]) continue;
            const double v = values[r * n_cols + c];
            if (v != v) p += snprintf(p, per_value, "%*s", width, "nan");              // Python prints nan without a sign
            else if (v - v != 0.0) p += snprintf(p, per_value, "%*e", width, v);        // +-inf
            else {
                // '%e' of a finite double: std::to_chars(scientific, 6) is the same correctly rounded decimal with the same
                // two-or-three-digit exponent, at a third of snprintf's cost (checked on 4e6 doubles over the whole exponent
                // range incl. ties at the seventh digit; tests/test_host_logic.py compares the writer with Python's '%e')
                const std::to_chars_result tc = std::to_chars(p, p + per_value, v, std::chars_format::scientific, 6);
                if (tc.ptr - p < width) p += snprintf(p, per_value, "%*e", width, v);   // never for width 5 / 6: the text has >= 12 characters
                else p = tc.ptr;
            }
        }
        *p++ = '\n';
    }
    out->assign(buf.get(), (size_t)(p - buf.get()));
}

}  // namespace

extern "C" int jacks_table_write(const char* path, const char* header, int64_t n_rows, int32_t n_cols, const char* labels,
                                 const int64_t* label_off, const double* values, const uint8_t* blank, int32_t width) {
    if (!path || !header || (n_rows > 0 && (!labels || !label_off || (n_cols > 0 && !values))))
        return fail(JACKS_E_INVALID, "jacks_table_write: NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(JACKS_E_INVALID, "cannot write %s: %s", path, strerror(errno));
    fputs(header, f);
    // number formatting dominates (snprintf "%e" ~ 0.2 us per value): rows are formatted by a pool of threads into
    // per-block strings and written in order
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)(hw ? hw : 1);
    if (n_thr > 32) n_thr = 32;
    const int64_t cells = n_rows * (int64_t)(n_cols > 0 ? n_cols : 1);
    if (cells < 200000) n_thr = 1;
    const int64_t block = 4096;
    const int64_t n_blocks = (n_rows + block - 1) / block;
    bool ok = true;
    // batches of n_thr blocks; the next batch is formatted while the previous one is written (the write of a 400 MB table
    // costs as much as its formatting)
    auto format_batch = [&](int64_t b0, std::vector<std::string>* outs, std::vector<std::thread>* pool) {
        const int nb = (int)((n_blocks - b0 < n_thr) ? (n_blocks - b0) : n_thr);
        outs->assign(nb > 0 ? nb : 0, std::string());
        for (int t = 0; t < nb; ++t) {
            const int64_t r0 = (b0 + t) * block, r1 = (r0 + block < n_rows) ? r0 + block : n_rows;
            if (n_thr == 1) format_rows(r0, r1, n_cols, labels, label_off, values, blank, width, &(*outs)[t]);
            else pool->emplace_back(format_rows, r0, r1, n_cols, labels, label_off, values, blank, width, &(*outs)[t]);
        }
    };
    std::vector<std::string> cur, nxt;
    std::vector<std::thread> pool_cur, pool_nxt;
    format_batch(0, &cur, &pool_cur);
    for (int64_t b0 = 0; b0 < n_blocks && ok; b0 += n_thr) {
        for (auto& th : pool_cur) th.join();
        pool_cur.clear();
        if (b0 + n_thr < n_blocks) format_batch(b0 + n_thr, &nxt, &pool_nxt);
        for (size_t t = 0; t < cur.size(); ++t)
            if (fwrite(cur[t].data(), 1, cur[t].size(), f) != cur[t].size()) ok = false;
        cur.swap(nxt);
        pool_cur.swap(pool_nxt);
    }
    for (auto& th : pool_cur) th.join();
    if (fclose(f) != 0 || !ok) return fail(JACKS_E_INVALID, "error writing %s", path);
    return JACKS_OK;
}

// ---- pseudo-gene draws on Python's random stream ---------------------------------------------------------------------
namespace {

// MT19937 exactly as CPython's _random module runs it (Modules/_randommodule.c: genrand_uint32).
struct Mt19937 {
    uint32_t mt[624];
    int pos;
    uint32_t next() {
        if (pos >= 624) {
            const uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
            int k = 0;
            for (; k < 624 - 397; ++k) { const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO); mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
            for (; k < 623; ++k) { const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO); mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
            const uint32_t y = (mt[623] & UP) | (mt[0] & LO);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // Random._randbelow_with_getrandbits(n), 1 <= n < 2^31: k = n.bit_length(); r = getrandbits(k) until r < n, and
    // getrandbits(k <= 32) is one output shifted right by 32 - k
    int64_t below(int64_t n) {
        int k = 0;
        while ((n >> k) != 0) ++k;
        uint32_t r = next() >> (32 - k);
        while ((int64_t)r >= n) r = next() >> (32 - k);
        return (int64_t)r;
    }
};

}  // namespace

extern "C" int jacks_pseudo_draw(uint32_t* mt_state, const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                                 const int64_t* ctrl_genes, int64_t n_ctrl, int64_t n_pseudo, int64_t* out_offsets,
                                 int32_t* out_rows, int64_t rows_cap, int64_t* rows_needed) {
    if (!mt_state || !out_offsets || !rows_needed || n_pseudo < 0 || n_genes < 0 || n_ctrl < 0 || rows_cap < 0)
        return fail(JACKS_E_INVALID, "jacks_pseudo_draw: bad argument");
    if ((n_genes > 0 && (!gene_offsets || !guide_rows)) || (n_ctrl > 0 && !ctrl_genes) || (rows_cap > 0 && !out_rows))
        return fail(JACKS_E_INVALID, "jacks_pseudo_draw: NULL argument");
    if (mt_state[624] > 624) return fail(JACKS_E_INVALID, "jacks_pseudo_draw: generator position %u outside [0, 624]", mt_state[624]);
    if (n_genes >= (1LL << 31) || n_ctrl >= (1LL << 31)) return fail(JACKS_E_INVALID, "jacks_pseudo_draw: too many genes");
    for (int64_t c = 0; c < n_ctrl; ++c)
        if (ctrl_genes[c] < 0 || ctrl_genes[c] >= n_genes) return fail(JACKS_E_INVALID, "control gene %lld outside the gene index", (long long)ctrl_genes[c]);
    Mt19937 g;
    memcpy(g.mt, mt_state, sizeof g.mt);
    g.pos = (int)mt_state[624];
    int64_t total = 0;
    out_offsets[0] = 0;
    for (int64_t i = 0; i < n_pseudo; ++i) {
        if (n_genes == 0) return fail(JACKS_E_INVALID, "Cannot choose from an empty sequence");          // random.choice([])
        const int64_t pick = g.below(n_genes);
        const int64_t n_guides = gene_offsets[pick + 1] - gene_offsets[pick];
        for (int64_t j = 0; j < n_guides; ++j) {
            if (n_ctrl == 0) return fail(JACKS_E_INVALID, "Cannot choose from an empty sequence");
            const int64_t cg = ctrl_genes[g.below(n_ctrl)];
            const int64_t o = gene_offsets[cg], n = gene_offsets[cg + 1] - o;
            if (n <= 0) return fail(JACKS_E_INVALID, "Cannot choose from an empty sequence");
            if (n >= (1LL << 31)) return fail(JACKS_E_INVALID, "jacks_pseudo_draw: gene with too many guides");
            const int32_t row = guide_rows[o + g.below(n)];
            if (total < rows_cap) out_rows[total] = row;
            ++total;
        }
        out_offsets[i + 1] = total;
    }
    *rows_needed = total;
    if (total > rows_cap) return fail(JACKS_E_SHAPE, "jacks_pseudo_draw: %lld rows drawn, buffer holds %lld", (long long)total, (long long)rows_cap);
    memcpy(mt_state, g.mt, sizeof g.mt);
    mt_state[624] = (uint32_t)g.pos;
    return JACKS_OK;
}
