// host_capi.cpp — liburlhost.so: C entry points around the host-side input/output helpers of `score` (fast_io.hpp) so the
// tests can hold them to libc's printf and to the CPU restatement of the reference's reader.  No CUDA.
#include "fast_io.hpp"

using namespace urlhost;

extern "C" {

int urlhost_format_score(float x, char *out /* >= 64 bytes */) { return format_score(x, out); }

struct urlhost_csv { ParsedCsv csv; std::string err; };
static thread_local std::string g_err;

urlhost_csv *urlhost_csv_open(const char *path, char delimiter, int has_header, int threads) {
    auto *c = new urlhost_csv();
    try { c->csv = parse_csv(path, delimiter, has_header != 0, threads); } catch (const std::exception &e) { g_err = e.what(); delete c; return nullptr; }
    return c;
}
const char *urlhost_last_error(void) { return g_err.c_str(); }
void urlhost_csv_free(urlhost_csv *c) { delete c; }
int urlhost_csv_p(urlhost_csv *c) { return c->csv.p; }
int64_t urlhost_csv_n(urlhost_csv *c) { return c->csv.n; }
int urlhost_csv_cardinality(urlhost_csv *c, int column) { return (int)c->csv.values[column].size(); }
const char *urlhost_csv_value(urlhost_csv *c, int column, int index) { return c->csv.values[column][index].c_str(); }
const char *urlhost_csv_header(urlhost_csv *c, int column) { return column < (int)c->csv.header.size() ? c->csv.header[column].c_str() : ""; }
void urlhost_csv_codes(urlhost_csv *c, int column, int32_t *out) { memcpy(out, c->csv.codes[column].data(), (size_t)c->csv.n * sizeof(int32_t)); }

} // extern "C"
