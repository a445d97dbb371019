// Fast MatrixMarket loader for the kept C host: ONE pass over the file instead of the reference's two
// fscanf passes (SSS_mat_read -> mmio_info + mmio_data, /root/reference/amg/SSS_main.c:12-22, mmio_highlevel.h:10-305), parsed
// by all host cores, with the reference loader's semantics entry for entry: coordinate format only; real / integer / pattern
// (value 1.0) / complex (imaginary part dropped); symmetric and hermitian files expanded entry by entry (the mirror of an
// off-diagonal entry directly follows it in row order); entries keep their FILE order inside each row -- no sorting, no merging
// of duplicates (the C/F splitting of the setup depends on that order).  Optional binary cache next to the file.
// Pure host code.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <omp.h>
#include <sys/stat.h>

#include "../../include/amg_b200.h"

namespace {

struct Parsed { std::vector<int> ri, ci; std::vector<double> va; };

inline const char *skip_ws(const char *p, const char *e) { while (p < e && (unsigned char)*p <= ' ') ++p; return p; }
inline bool parse_int(const char *&p, const char *e, int &v) {
    p = skip_ws(p, e);
    if (p >= e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    if (p >= e || !isdigit((unsigned char)*p)) return false;
    long long x = 0;
    while (p < e && isdigit((unsigned char)*p)) { x = x * 10 + (*p - '0'); ++p; }
    v = (int)(neg ? -x : x);
    return true;
}
inline bool parse_double(const char *&p, const char *e, double &v) {      // strtod: correctly rounded, like the reference's fscanf("%lg")
    p = skip_ws(p, e);
    if (p >= e) return false;
    char *end = nullptr;
    v = strtod(p, &end);
    if (end == p) return false;
    p = end;
    return true;
}

// entries of [b, e) (whole lines), appended in order; false on a malformed entry
bool parse_chunk(const char *b, const char *e, int field, int *ri, int *ci, double *va, long long cap, long long &got) {
    const char *p = b;
    got = 0;
    for (;;) {
        p = skip_ws(p, e);
        if (p >= e) return true;
        if (got >= cap) return false;
        int i, j;
        double v = 1.0, im;
        if (!parse_int(p, e, i) || !parse_int(p, e, j)) return false;
        if (field == 0) { if (!parse_double(p, e, v)) return false; }                                         // real
        else if (field == 1) { if (!parse_double(p, e, v) || !parse_double(p, e, im)) return false; }          // complex: real part
        else if (field == 2) { int iv; if (!parse_int(p, e, iv)) return false; v = iv; }                       // integer
        ri[got] = i - 1; ci[got] = j - 1; va[got] = v;
        ++got;
    }
}

bool read_cache(const std::string &cache, const struct stat &st, amgb200_mat *A) {
    FILE *f = fopen(cache.c_str(), "rb");
    if (!f) return false;
    long long hd[6];
    bool ok = fread(hd, sizeof(hd), 1, f) == 1 && hd[0] == 0x58544d3030324241LL && hd[1] == (long long)st.st_size && hd[2] == (long long)st.st_mtime;
    if (ok) {
        A->num_rows = (int)hd[3]; A->num_cols = (int)hd[4]; A->num_nnzs = (int)hd[5];
        A->row_ptr = (int *)malloc(((size_t)hd[3] + 1) * sizeof(int));
        A->col_idx = (int *)malloc((size_t)std::max<long long>(hd[5], 1) * sizeof(int));
        A->val = (double *)malloc((size_t)std::max<long long>(hd[5], 1) * sizeof(double));
        ok = fread(A->row_ptr, sizeof(int), (size_t)hd[3] + 1, f) == (size_t)hd[3] + 1 && fread(A->col_idx, sizeof(int), (size_t)hd[5], f) == (size_t)hd[5] &&
             fread(A->val, sizeof(double), (size_t)hd[5], f) == (size_t)hd[5];
        if (!ok) { free(A->row_ptr); free(A->col_idx); free(A->val); memset(A, 0, sizeof(*A)); }
    }
    fclose(f);
    return ok;
}

void write_cache(const std::string &cache, const struct stat &st, const amgb200_mat *A) {
    FILE *f = fopen(cache.c_str(), "wb");
    if (!f) return;
    const long long hd[6] = {0x58544d3030324241LL, (long long)st.st_size, (long long)st.st_mtime, A->num_rows, A->num_cols, A->num_nnzs};
    fwrite(hd, sizeof(hd), 1, f);
    fwrite(A->row_ptr, sizeof(int), (size_t)A->num_rows + 1, f);
    fwrite(A->col_idx, sizeof(int), (size_t)A->num_nnzs, f);
    fwrite(A->val, sizeof(double), (size_t)A->num_nnzs, f);
    fclose(f);
}

}  // namespace

// Returns 0 on success; -1 cannot open, -2 bad banner / unsupported format, -4 bad size line, -5 malformed entries (the
// reference ignores its loader's return codes, SSS_main.c:16,20; a caller of this function should not).  Arrays are malloc'ed
// (SSS_free / free).  AMGB200_MTX_CACHE=1: keep / use `<file>.amgb200cache` (raw CSR, validated against the file's size and mtime).
extern "C" int amgb200_read_mtx(const char *filename, amgb200_mat *A) {
    memset(A, 0, sizeof(*A));
    struct stat st;
    if (stat(filename, &st) != 0) return -1;
    const bool use_cache = getenv("AMGB200_MTX_CACHE") && atoi(getenv("AMGB200_MTX_CACHE"));
    const std::string cache = std::string(filename) + ".amgb200cache";
    if (use_cache && read_cache(cache, st, A)) return 0;
    FILE *f = fopen(filename, "rb");
    if (!f) return -1;
    std::vector<char> buf((size_t)st.st_size + 1);
    const size_t got_bytes = fread(buf.data(), 1, (size_t)st.st_size, f);
    fclose(f);
    buf[got_bytes] = 0;
    const char *p = buf.data(), *end = buf.data() + got_bytes;
    // banner: %%MatrixMarket matrix coordinate <field> <symmetry>   (mm_read_banner: tokens compared in lower case)
    const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
    if (!eol) return -2;
    std::string banner(p, eol);
    for (char &c : banner) c = (char)tolower((unsigned char)c);
    char t0[64], t1[64], t2[64], t3[64], t4[64];
    if (sscanf(banner.c_str(), "%63s %63s %63s %63s %63s", t0, t1, t2, t3, t4) != 5 || strcmp(t0, "%%matrixmarket") || strcmp(t1, "matrix") || strcmp(t2, "coordinate")) return -2;
    int field;                                    // 0 real, 1 complex, 2 integer, 3 pattern
    if (!strcmp(t3, "real")) field = 0; else if (!strcmp(t3, "complex")) field = 1; else if (!strcmp(t3, "integer")) field = 2; else if (!strcmp(t3, "pattern")) field = 3; else return -2;
    const bool sym = !strcmp(t4, "symmetric") || !strcmp(t4, "hermitian");
    if (!sym && strcmp(t4, "general") && strcmp(t4, "skew-symmetric")) return -2;       // (skew-symmetric files are read as stored, like the reference)
    p = eol + 1;
    while (p < end && *p == '%') { eol = (const char *)memchr(p, '\n', (size_t)(end - p)); p = eol ? eol + 1 : end; }
    int m, n, nz;
    if (!parse_int(p, end, m) || !parse_int(p, end, n) || !parse_int(p, end, nz) || m <= 0 || n <= 0 || nz < 0) return -4;
    // ---- the entries: the rest of the file is cut at line ends into one piece per thread, each piece parsed independently
    std::vector<int> ri((size_t)std::max(nz, 1)), ci((size_t)std::max(nz, 1));
    std::vector<double> va((size_t)std::max(nz, 1));
    const int nt = std::max(1, std::min(omp_get_max_threads(), (int)((end - p) / (1 << 20)) + 1));
    std::vector<const char *> cut((size_t)nt + 1);
    cut[0] = p; cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char *q = p + (end - p) * t / nt;
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
        cut[t] = nl ? nl + 1 : end;
    }
    std::vector<long long> lines((size_t)nt + 1, 0);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; ++t) {               // entries per piece = non-blank lines (one entry per line: checked by the parse below)
        long long c = 0;
        const char *q = cut[t];
        while (q < cut[t + 1]) {
            const char *nl = (const char *)memchr(q, '\n', (size_t)(cut[t + 1] - q));
            const char *le = nl ? nl : cut[t + 1];
            if (skip_ws(q, le) < le) ++c;
            q = le + 1;
        }
        lines[t + 1] = c;
    }
    for (int t = 0; t < nt; ++t) lines[t + 1] += lines[t];
    bool ok = lines[nt] == nz;
    if (ok) {
        int bad = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nt) reduction(+ : bad)
        for (int t = 0; t < nt; ++t) {
            long long got = 0;
            const long long cap = lines[t + 1] - lines[t];
            if (!parse_chunk(cut[t], cut[t + 1], field, ri.data() + lines[t], ci.data() + lines[t], va.data() + lines[t], cap, got) || got != cap) ++bad;
        }
        ok = bad == 0;
    }
    if (!ok) {                                    // entries not one per line (fscanf does not care): one sequential token pass
        long long got = 0;
        if (!parse_chunk(p, end, field, ri.data(), ci.data(), va.data(), nz, got) || got != nz) return -5;
    }
    for (int k = 0; k < nz; ++k) if (ri[k] < 0 || ri[k] >= m || ci[k] < 0 || ci[k] >= n) return -5;
    // ---- CSR in file order (mmio_highlevel.h:240-295): count, exclusive scan, stable placement; symmetric files: each off-diagonal
    // entry is placed, then its mirror
    std::vector<int> cnt((size_t)m + 1, 0);
    for (int k = 0; k < nz; ++k) { cnt[ri[k]]++; if (sym && ri[k] != ci[k]) cnt[ci[k]]++; }
    A->num_rows = m; A->num_cols = n;
    A->row_ptr = (int *)malloc(((size_t)m + 1) * sizeof(int));
    long long tot = 0;
    for (int i = 0; i < m; ++i) { A->row_ptr[i] = (int)tot; tot += cnt[i]; }
    A->row_ptr[m] = (int)tot;
    A->num_nnzs = (int)tot;
    A->col_idx = (int *)malloc((size_t)std::max<long long>(tot, 1) * sizeof(int));
    A->val = (double *)malloc((size_t)std::max<long long>(tot, 1) * sizeof(double));
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int k = 0; k < nz; ++k) {
        int o = A->row_ptr[ri[k]] + cnt[ri[k]]++;
        A->col_idx[o] = ci[k]; A->val[o] = va[k];
        if (sym && ri[k] != ci[k]) {
            o = A->row_ptr[ci[k]] + cnt[ci[k]]++;
            A->col_idx[o] = ri[k]; A->val[o] = va[k];
        }
    }
    if (use_cache) write_cache(cache, st, A);
    return 0;
}
