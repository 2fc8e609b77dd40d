/* at_io.c — memory-mapped, multi-threaded edge-list parser (C ABI: include/at_io.h).  Plain C + pthreads, host only.
 *
 * The file is cut into one chunk per thread at row boundaries; pass 1 counts the surviving rows of every chunk, an
 * exclusive prefix sum turns the counts into output offsets, pass 2 parses again and writes in place: rows keep their
 * order without any inter-thread communication. */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "at_io.h"

static __thread char g_err[512] = "";

static int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

/* pandas' default NA tokens: they make a value missing without making the column text */
static int is_na_token(const char* p, size_t len)
{
    static const char* const na[] = {"nan", "NaN", "NA", "N/A", "n/a", "NULL", "null", "#N/A", "None", "<NA>", "-nan", "-NaN", "#NA",
                                     "1.#IND", "-1.#IND", "1.#QNAN", "-1.#QNAN", "#N/A N/A", "inf", "-inf", "+inf", "Inf", "-Inf",
                                     "+Inf", "INF", "-INF", "+INF", "infinity", "-infinity", "+infinity", "Infinity", "-Infinity",
                                     "+Infinity"};
    for (size_t i = 0; i < sizeof(na) / sizeof(na[0]); ++i)
        if (strlen(na[i]) == len && memcmp(na[i], p, len) == 0) return 1;
    return 0;
}

/* One field [p, e): 1 and *out if numeric, 0 if missing (empty or an NA token), -1 if it is text. */
static int parse_field(const char* p, const char* e, int64_t* out)
{
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    while (e > p && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) --e;
    if (e - p >= 2 && *p == '"' && e[-1] == '"') { ++p; --e; }      /* "12" */
    if (p >= e) return 0;
    /* fast path: [+-]digits, at most 18 digits (no overflow) */
    const char* q = p;
    int neg = 0;
    if (*q == '-' || *q == '+') { neg = *q == '-'; ++q; }
    if (q < e && e - q <= 18) {
        int64_t v = 0;
        const char* r = q;
        while (r < e && *r >= '0' && *r <= '9') v = v * 10 + (*r++ - '0');
        if (r == e) { *out = neg ? -v : v; return 1; }
    }
    /* general path: a floating-point literal (pandas' to_numeric), truncated like astype(int64) */
    char buf[64];
    const size_t len = (size_t)(e - p);
    if (len >= sizeof(buf)) return -1;
    if (is_na_token(p, len)) return 0;
    memcpy(buf, p, len);
    buf[len] = 0;
    for (size_t i = 0; i < len; ++i) {                       /* strtod also accepts hex floats and "nan(...)": pandas does not */
        const char c = buf[i];
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) return -1;
    }
    char* end = NULL;
    errno = 0;
    const double d = strtod(buf, &end);
    if (end != buf + len) return -1;
    if (errno == ERANGE || !isfinite(d) || fabs(d) >= 9.2e18) return 0;
    *out = (int64_t)d;
    return 1;
}

/* One row [p, e) without its '\n'.  Returns 1 (edge in *s, *d), 0 (dropped) or -1 (empty row).  *text |= some field is text. */
static int parse_row(const char* p, const char* e, int64_t* s, int64_t* d, int* text)
{
    const char* t = p;
    while (t < e && (*t == ' ' || *t == '\t' || *t == '\r')) ++t;
    if (t == e) return -1;
    int field = 0, ok = 1;
    const char* f = p;
    for (const char* c = p;; ++c) {
        if (c == e || *c == ',') {
            int64_t v = 0;
            const int kind = parse_field(f, c, &v);
            if (field == 0) { *s = v; ok &= kind == 1; }
            else if (field == 1) { *d = v; ok &= kind == 1; }
            if (kind < 0) *text = 1;
            ++field;
            f = c + 1;
            if (c == e) break;
        }
    }
    return (ok && field >= 2) ? 1 : 0;
}

typedef struct {
    const char* begin;      /* chunk [begin, end), begins at a row start */
    const char* end;
    int64_t count;          /* pass 1 */
    int64_t offset;         /* pass 2: first output slot */
    int64_t* src;
    int64_t* dst;
    int64_t max_id;
    int text;               /* pass 1: some field of the chunk is text */
} chunk_t;

static void* run_chunk(void* arg)
{
    chunk_t* c = (chunk_t*)arg;
    const char* p = c->begin;
    int64_t n = 0, mx = -1;
    int text = 0;
    const char* const end = c->end;
    while (p < end) {
        int64_t s, d;
        /* fast path for the common row "digits,digits\n": one scan, no calls */
        {
            const char* q = p;
            int64_t a = 0, b = 0;
            while (q < end && (unsigned)(*q - '0') < 10u) a = a * 10 + (*q++ - '0');
            if (q > p && q - p <= 18 && q < end && *q == ',') {
                const char* f = ++q;
                while (q < end && (unsigned)(*q - '0') < 10u) b = b * 10 + (*q++ - '0');
                if (q > f && q - f <= 18) {
                    const char* next = NULL;
                    if (q == end) next = end;
                    else if (*q == '\n') next = q + 1;
                    else if (*q == '\r' && q + 1 < end && q[1] == '\n') next = q + 2;
                    if (next) {
                        if (c->src) {
                            c->src[c->offset + n] = a;
                            c->dst[c->offset + n] = b;
                            if (a > mx) mx = a;
                            if (b > mx) mx = b;
                        }
                        ++n;
                        p = next;
                        continue;
                    }
                }
            }
        }
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl : end;
        if (parse_row(p, e, &s, &d, &text) == 1) {
            if (c->src) {
                c->src[c->offset + n] = s;
                c->dst[c->offset + n] = d;
                if (s > mx) mx = s;
                if (d > mx) mx = d;
            }
            ++n;
        }
        p = e + 1;
    }
    c->count = n;
    c->max_id = mx;
    c->text = text;
    return NULL;
}

static int run_file(const char* path, int32_t n_threads, int64_t capacity, int64_t* src, int64_t* dst, int64_t* n_edges, int64_t* max_id)
{
    if (!path || !n_edges) return fail("at_io: %s", "null path / n_edges");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail("at_io: cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return fail("at_io: fstat %s: %s", path, strerror(errno)); }
    const size_t size = (size_t)st.st_size;
    *n_edges = 0;
    if (max_id) *max_id = -1;
    if (size == 0) { close(fd); return 0; }
    const char* data = (const char*)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) return fail("at_io: mmap %s: %s", path, strerror(errno));
    madvise((void*)data, size, MADV_SEQUENTIAL);
    const char* end = data + size;

    /* first non-empty row: it is a header if ANY field of the file is text (see at_io.h), decided after pass 1 */
    const char* body = data;
    const char* first_end = NULL;
    int first_valid = 0;
    while (body < end) {
        const char* nl = (const char*)memchr(body, '\n', (size_t)(end - body));
        const char* e = nl ? nl : end;
        int64_t s, d;
        int text = 0;
        const int r = parse_row(body, e, &s, &d, &text);
        if (r == -1) { body = e + 1; continue; }              /* empty row: keep looking */
        first_end = e;
        first_valid = r == 1;
        break;
    }
    if (body > end) body = end;

    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    int nt = n_threads > 0 ? n_threads : (int)(cores > 0 ? cores : 1);
    const size_t body_size = (size_t)(end - body);
    if ((size_t)nt > body_size / 65536 + 1) nt = (int)(body_size / 65536 + 1);   /* small files: few threads */
    if (nt > 256) nt = 256;
    chunk_t* chunks = (chunk_t*)calloc((size_t)nt, sizeof(chunk_t));
    pthread_t* threads = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    if (!chunks || !threads) { free(chunks); free(threads); munmap((void*)data, size); return fail("at_io: %s", "out of memory"); }
    const char* p = body;
    for (int i = 0; i < nt; ++i) {
        const char* q = (i == nt - 1) ? end : body + body_size * (size_t)(i + 1) / (size_t)nt;
        if (q < p) q = p;
        if (q < end && i != nt - 1) {                          /* advance to the next row start */
            const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
            q = nl ? nl + 1 : end;
        }
        chunks[i].begin = p;
        chunks[i].end = q;
        p = q;
    }
    int rc = 0, header_done = 0;
    for (int pass = 0; pass < (src ? 2 : 1) && rc == 0; ++pass) {
        if (pass == 1) {
            int text = 0;
            for (int i = 0; i < nt; ++i) text |= chunks[i].text;
            if (text && first_end != NULL) {                   /* drop the header row (it always lies in chunk 0) */
                chunks[0].begin = first_end + 1 < chunks[0].end ? first_end + 1 : chunks[0].end;
                chunks[0].count -= first_valid;
            }
            header_done = 1;
            int64_t total = 0;
            for (int i = 0; i < nt; ++i) { chunks[i].offset = total; total += chunks[i].count; chunks[i].src = src; chunks[i].dst = dst; }
            if (total > capacity) { rc = fail("at_io: %lld edges do not fit the capacity %lld", (long long)total, (long long)capacity); break; }
        }
        int started = 0;
        for (int i = 0; i < nt; ++i) {
            if (pthread_create(&threads[i], NULL, run_chunk, &chunks[i]) != 0) { run_chunk(&chunks[i]); threads[i] = 0; }
            else started |= 1;
        }
        for (int i = 0; i < nt; ++i)
            if (threads[i]) pthread_join(threads[i], NULL);
        (void)started;
    }
    if (rc == 0) {
        if (!header_done) {                                    /* count-only call: apply the header rule to the counts */
            int text = 0;
            for (int i = 0; i < nt; ++i) text |= chunks[i].text;
            if (text && first_end != NULL) chunks[0].count -= first_valid;
        }
        int64_t total = 0, mx = -1;
        for (int i = 0; i < nt; ++i) { total += chunks[i].count; if (chunks[i].max_id > mx) mx = chunks[i].max_id; }
        *n_edges = total;
        if (max_id) *max_id = mx;
    }
    free(chunks);
    free(threads);
    munmap((void*)data, size);
    return rc;
}

int at_io_abi_version(void) { return AT_IO_ABI_VERSION; }
const char* at_io_last_error(void) { return g_err; }

int at_io_edge_list_count(const char* path, int32_t n_threads, int64_t* n_edges)
{
    return run_file(path, n_threads, 0, NULL, NULL, n_edges, NULL);
}

int at_io_edge_list_parse(const char* path, int32_t n_threads, int64_t capacity, int64_t* src, int64_t* dst, int64_t* n_edges,
                          int64_t* max_id)
{
    if (!src || !dst) return fail("at_io: %s", "null output arrays");
    if (capacity < 0) return fail("at_io: %s", "negative capacity");
    return run_file(path, n_threads, capacity, src, dst, n_edges, max_id);
}
