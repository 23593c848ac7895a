// sdx_io.cpp — host-side format path of libsdx.so (no GPU): the sample -> events TSV files that carry
// feature matrices and null matrices between the reference's generator and driver.
//
//   writer  replaces the row building and file writing of src/generate_null_matrices.py:75-86
//   reader  replaces the parse of src/i_o.py:18-88 for files whose samples and events are known
//           (every null file of a feature matrix), producing packed feature rows directly.
//
// Format (src/i_o.py:54-63): lines starting with '#' are comments (the writer emits "#Samples\tEvents"),
// every other line is "sample<TAB>event<TAB>event..."; a sample without events is a bare name.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sdx.h"

static thread_local std::string t_io_error;

extern "C" const char *sdx_io_last_error(void) { return t_io_error.c_str(); }

// One null: feature rows (F x wpr words over S samples) -> text.  Events of a sample are listed in feature order,
// as zip(permuted_mat[r], genes) does (src/generate_null_matrices.py:78).
static bool write_one(const uint32_t *rows, int F, int S, int wpr, const char *const *samples, const size_t *slen,
                      const char *const *genes, const size_t *glen, const std::string &path, std::vector<uint32_t> &cnt,
                      std::vector<int32_t> &ev, std::string &buf) {
    // counting sort of (sample, feature) pairs by sample, features ascending
    cnt.assign((size_t)S + 1, 0);
    for (int g = 0; g < F; ++g)
        for (int wi = 0; wi < wpr; ++wi) {
            uint32_t x = rows[(size_t)g * wpr + wi];
            while (x) { const int s = wi * 32 + __builtin_ctz(x); x &= x - 1; if (s < S) cnt[s + 1]++; }
        }
    for (int s = 0; s < S; ++s) cnt[s + 1] += cnt[s];
    ev.resize(cnt[S]);
    std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
    for (int g = 0; g < F; ++g)
        for (int wi = 0; wi < wpr; ++wi) {
            uint32_t x = rows[(size_t)g * wpr + wi];
            while (x) { const int s = wi * 32 + __builtin_ctz(x); x &= x - 1; if (s < S) ev[pos[s]++] = g; }
        }
    buf.clear();
    buf.append("#Samples\tEvents\n");
    for (int s = 0; s < S; ++s) {
        buf.append(samples[s], slen[s]);
        for (uint32_t q = cnt[s]; q < cnt[s + 1]; ++q) { buf.push_back('\t'); buf.append(genes[ev[q]], glen[ev[q]]); }
        buf.push_back('\n');
    }
    FILE *fh = fopen(path.c_str(), "wb");
    if (!fh) return false;
    const bool ok = fwrite(buf.data(), 1, buf.size(), fh) == buf.size();
    return fclose(fh) == 0 && ok;
}

extern "C" int sdx_write_nulls_tsv(const uint32_t *feature_rows, int64_t n_nulls, int n_features, int n_samples,
                                   int words_per_row, const char *const *sample_names, const char *const *event_names,
                                   const char *path_prefix, int64_t first_index, int n_threads) {
    if (!feature_rows || n_nulls < 0 || n_features < 1 || n_samples < 1 || !sample_names || !event_names || !path_prefix ||
        words_per_row != (n_samples + 31) / 32) {
        t_io_error = "sdx_write_nulls_tsv: bad argument";
        return SDX_ERR_INVALID;
    }
    std::vector<size_t> slen(n_samples), glen(n_features);
    for (int s = 0; s < n_samples; ++s) slen[s] = strlen(sample_names[s]);
    for (int g = 0; g < n_features; ++g) glen[g] = strlen(event_names[g]);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n_nulls) nt = (int)(n_nulls > 0 ? n_nulls : 1);
    std::atomic<int64_t> next(0);
    std::atomic<int> failed(0);
    const std::string prefix(path_prefix);
    const size_t per = (size_t)n_features * words_per_row;
    auto work = [&]() {
        std::vector<uint32_t> cnt;
        std::vector<int32_t> ev;
        std::string buf;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_nulls) break;
            const std::string path = prefix + std::to_string(first_index + i) + ".tsv";
            if (!write_one(feature_rows + (size_t)i * per, n_features, n_samples, words_per_row, sample_names, slen.data(),
                           event_names, glen.data(), path, cnt, ev, buf))
                failed.store(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (failed.load()) { t_io_error = std::string("sdx_write_nulls_tsv: cannot write under ") + prefix; return SDX_ERR_INVALID; }
    return SDX_OK;
}

struct NameIndex {
    std::unordered_map<std::string, int> map;
    NameIndex(const char *const *names, int n) {
        map.reserve((size_t)n * 2);
        for (int i = 0; i < n; ++i) map.emplace(names[i], i);
    }
    int find(const char *p, size_t len, std::string &tmp) const {
        tmp.assign(p, len);
        auto it = map.find(tmp);
        return it == map.end() ? -1 : it->second;
    }
};

static bool read_one(const std::string &path, const NameIndex &samples, const NameIndex &events, int F, int wpr, uint32_t *rows,
                     int64_t *n_unknown) {
    FILE *fh = fopen(path.c_str(), "rb");
    if (!fh) return false;
    std::string data;
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, fh)) > 0) data.append(chunk, got);
    fclose(fh);
    memset(rows, 0, sizeof(uint32_t) * (size_t)F * wpr);
    std::string tmp;
    int64_t unknown = 0;
    size_t p = 0;
    const size_t n = data.size();
    while (p < n) {
        size_t e = data.find('\n', p);
        if (e == std::string::npos) e = n;
        size_t end = e;
        while (end > p && (data[end - 1] == '\r' || data[end - 1] == ' ' || data[end - 1] == '\t')) --end;     // rstrip, as the reference
        if (end > p && data[p] != '#') {
            size_t q = p;
            size_t tab = data.find('\t', q);
            if (tab == std::string::npos || tab > end) tab = end;
            const int s = samples.find(data.data() + q, tab - q, tmp);
            if (s >= 0) {                                   // sample whitelist (src/i_o.py:57)
                q = tab + 1;
                while (q <= end && tab < end) {
                    tab = data.find('\t', q);
                    if (tab == std::string::npos || tab > end) tab = end;
                    const int g = events.find(data.data() + q, tab - q, tmp);
                    if (g >= 0) rows[(size_t)g * wpr + (s >> 5)] |= 1u << (s & 31);
                    else ++unknown;
                    q = tab + 1;
                }
            }
        }
        p = e + 1;
    }
    if (n_unknown) *n_unknown += unknown;
    return true;
}

extern "C" int sdx_read_nulls_tsv(const char *path_prefix, int64_t first_index, int64_t n_nulls, int n_features, int n_samples,
                                  int words_per_row, const char *const *sample_names, const char *const *event_names,
                                  uint32_t *feature_rows, int64_t *n_unknown_events, int n_threads) {
    if (!path_prefix || n_nulls < 0 || n_features < 1 || n_samples < 1 || !sample_names || !event_names || !feature_rows ||
        words_per_row != (n_samples + 31) / 32) {
        t_io_error = "sdx_read_nulls_tsv: bad argument";
        return SDX_ERR_INVALID;
    }
    const NameIndex samples(sample_names, n_samples), events(event_names, n_features);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n_nulls) nt = (int)(n_nulls > 0 ? n_nulls : 1);
    std::atomic<int64_t> next(0), unknown(0);
    std::atomic<int64_t> failed(-1);
    const std::string prefix(path_prefix);
    const size_t per = (size_t)n_features * words_per_row;
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_nulls) break;
            int64_t u = 0;
            if (!read_one(prefix + std::to_string(first_index + i) + ".tsv", samples, events, n_features, words_per_row,
                          feature_rows + (size_t)i * per, &u))
                failed.store(first_index + i);
            unknown.fetch_add(u);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (n_unknown_events) *n_unknown_events = unknown.load();
    if (failed.load() >= 0) {
        t_io_error = std::string("sdx_read_nulls_tsv: cannot read ") + prefix + std::to_string(failed.load()) + ".tsv";
        return SDX_ERR_INVALID;
    }
    return SDX_OK;
}
