// Host side of the result pipeline: the reference writes every chain's traces with
// np.savetxt (R:454-481) and reads them back with np.loadtxt in show_results (R:794-831).  Once sampling
// takes a tenth of a second those two calls ARE the run time of run_chains(), so the library carries
// byte-compatible replacements: same text, same parse, no interpreter in the loop (callers run one
// chain per thread; ctypes drops the GIL).  No CUDA here.
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptfnn.h"

namespace {

// np.savetxt applies ONE printf conversion to every value; accept exactly that: '%' [flags] [width] ['.' prec] (e|E|f|F|g|G)
bool format_ok(const char *fmt) {
    if (!fmt || fmt[0] != '%') return false;
    const char *p = fmt + 1;
    while (*p && strchr("-+ #0", *p)) ++p;
    while (*p >= '0' && *p <= '9') ++p;
    if (*p == '.') { ++p; while (*p >= '0' && *p <= '9') ++p; }
    return *p && strchr("eEfFgG", *p) && p[1] == '\0' && (p - fmt) < 16;
}

// '%[1].Ne' / '%[1].Nf' (every format the reference uses) go through std::to_chars: the shortest path to the
// same correctly rounded digits printf produces, ~5x faster than glibc's snprintf.  -1: not such a format.
int fast_precision(const char *fmt, char &conv) {
    const char *p = fmt + 1;
    if (*p == '1') ++p;                                          // a minimum width of 1 never pads
    if (*p != '.') return -1;
    ++p;
    if (*p < '0' || *p > '9') return -1;
    int prec = 0;
    while (*p >= '0' && *p <= '9') prec = prec * 10 + (*p++ - '0');
    if ((*p != 'e' && *p != 'f') || p[1] != '\0' || prec > 60) return -1;
    conv = *p;
    return prec;
}

}  // namespace

// np.savetxt(path, X, fmt=fmt): one row per line, values separated by one space, '\n' line ends
// (R:455-481).  X is [rows, cols] with `row_stride` doubles between rows; a 1-D array is rows x 1.
extern "C" int ptfnn_savetxt(const char *path, const double *data, int64_t rows, int64_t cols, int64_t row_stride,
                             const char *fmt) {
    if (!path || (!data && rows * cols > 0) || rows < 0 || cols < 1 || row_stride < cols || !format_ok(fmt)) return PTFNN_E_INVALID;
    FILE *f = fopen(path, "wb");
    if (!f) return PTFNN_E_STATE;
    std::vector<char> buf(1 << 20);
    size_t used = 0;
    bool ok = true;
    char conv = 'e';
    const int prec = fast_precision(fmt, conv);
    for (int64_t r = 0; r < rows && ok; ++r) {
        const double *row = data + r * row_stride;
        for (int64_t c = 0; c < cols; ++c) {
            if (buf.size() - used < 400) {                       // '%f' of 1e308 needs ~320 characters
                ok = fwrite(buf.data(), 1, used, f) == used;
                used = 0;
                if (!ok) break;
            }
            char *at = buf.data() + used, *lim = buf.data() + buf.size();
            if (prec >= 0 && std::isfinite(row[c])) {
                const auto res = std::to_chars(at, lim, row[c], conv == 'e' ? std::chars_format::scientific : std::chars_format::fixed, prec);
                if (res.ec != std::errc()) { ok = false; break; }
                used += (size_t)(res.ptr - at);
            } else {
                const int n = snprintf(at, (size_t)(lim - at), fmt, row[c]);
                if (n < 0 || n >= lim - at) { ok = false; break; }
                used += (size_t)n;
            }
            buf[used++] = c + 1 < cols ? ' ' : '\n';
        }
    }
    if (ok && used) ok = fwrite(buf.data(), 1, used, f) == used;
    if (fclose(f) != 0) ok = false;
    return ok ? PTFNN_OK : PTFNN_E_STATE;
}

// np.loadtxt(path) for the files above: whitespace-separated floats, every line the same number of
// columns, blank lines skipped.  Two calls: out == NULL reports the shape; then `capacity` >= rows*cols.
extern "C" int ptfnn_loadtxt(const char *path, double *out, int64_t capacity, int64_t *rows_out, int64_t *cols_out) {
    if (!path || !rows_out || !cols_out) return PTFNN_E_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) return PTFNN_E_STATE;
    std::string text;
    {
        std::vector<char> buf(1 << 20);
        size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) text.append(buf.data(), n);
    }
    const bool read_ok = !ferror(f);
    fclose(f);
    if (!read_ok) return PTFNN_E_STATE;
    int64_t rows = 0, cols = -1, count = 0;
    const char *p = text.c_str(), *end = p + text.size();
    while (p < end) {
        int64_t in_line = 0;
        while (p < end && *p != '\n') {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
            if (p >= end || *p == '\n') break;
            double v = 0.0;
            const char *q = p;
            const auto res = std::from_chars(p, end, v);         // (correctly rounded, like strtod; several times faster)
            if (res.ec == std::errc()) q = res.ptr;
            else {                                               // '+1.5', hex floats, out-of-range spellings: strtod decides
                char *e = nullptr;
                v = strtod(p, &e);
                q = e;
            }
            if (q == p) return PTFNN_E_INVALID;                  // not a number
            if (out) {
                if (count >= capacity) return PTFNN_E_INVALID;
                out[count] = v;
            }
            ++count; ++in_line;
            p = q;
        }
        if (p < end) ++p;                                        // the '\n'
        if (in_line == 0) continue;
        if (cols < 0) cols = in_line;
        else if (cols != in_line) return PTFNN_E_INVALID;        // ragged file (np.loadtxt raises too)
        ++rows;
    }
    *rows_out = rows;
    *cols_out = cols < 0 ? 0 : cols;
    return PTFNN_OK;
}
