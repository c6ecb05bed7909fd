/* mtx_io.c -- see mtx_io.h */
#define _POSIX_C_SOURCE 200809L
#include "mtx_io.h"
#include <ctype.h>
#include <err.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>

static double now(void)
{
        struct timeval tv;
        gettimeofday(&tv, NULL);
        return tv.tv_sec + tv.tv_usec * 1e-6;
}

static char *slurp(const char *filename, size_t *len)
{
        FILE *f = fopen(filename, "rb");
        if (!f)
                err(1, "impossible d'ouvrir %s", filename);
        struct stat sb;
        if (fstat(fileno(f), &sb) != 0)
                err(1, "stat %s", filename);
        size_t size = (size_t)sb.st_size;
        char *buf = malloc(size + 1);
        if (!buf)
                err(1, "Cannot allocate %zu bytes for %s", size, filename);
        size_t got = 0;
        while (got < size) {
                size_t r = fread(buf + got, 1, size - got, f);
                if (r == 0)
                        break;
                got += r;
        }
        fclose(f);
        buf[got] = 0;
        *len = got;
        return buf;
}

/* next line starting at *pos; returns pointer and sets *linelen; advances *pos */
static const char *next_line(const char *buf, size_t len, size_t *pos, size_t *linelen)
{
        if (*pos >= len)
                return NULL;
        const char *s = buf + *pos;
        const char *e = memchr(s, '\n', len - *pos);
        size_t l = e ? (size_t)(e - s) : len - *pos;
        *pos += l + (e ? 1 : 0);
        *linelen = l;
        return s;
}

/* scan one optionally signed decimal integer, skipping leading white space */
static inline int scan_int(const char **pp, const char *end, long long *out)
{
        const char *p = *pp;
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r'))
                p++;
        if (p >= end)
                return 0;
        int neg = 0;
        if (*p == '-' || *p == '+') {
                neg = (*p == '-');
                p++;
        }
        if (p >= end || *p < '0' || *p > '9')
                return 0;
        long long v = 0;
        while (p < end && *p >= '0' && *p <= '9')
                v = v * 10 + (*p++ - '0');
        *out = neg ? -v : v;
        *pp = p;
        return 1;
}

static int ieq(const char *a, const char *b)
{
        for (; *a && *b; a++, b++)
                if (tolower((unsigned char)*a) != tolower((unsigned char)*b))
                        return 0;
        return *a == *b;
}

void mtx_load(struct coo_matrix *M, const char *filename, uint64_t prime)
{
        printf("Loading matrix from %s\n", filename);
        fflush(stdout);
        double t0 = now();
        size_t len, pos = 0, ll;
        char *buf = slurp(filename, &len);

        /* banner: %%MatrixMarket object format field symmetry */
        const char *line = next_line(buf, len, &pos, &ll);
        char banner[16] = "", object[64] = "", format[64] = "", field[64] = "", symm[64] = "";
        char first[512];
        if (!line)
                errx(1, "Could not process Matrix Market banner.\n");
        snprintf(first, sizeof(first), "%.*s", (int)(ll < 500 ? ll : 500), line);
        if (sscanf(first, "%15s %63s %63s %63s %63s", banner, object, format, field, symm) != 5 ||
            strcmp(banner, "%%MatrixMarket") != 0)
                errx(1, "Could not process Matrix Market banner.\n");
        if (!ieq(object, "matrix") || !ieq(format, "coordinate"))
                errx(1, "Matrix Market type: [%s %s %s %s] not supported (only sparse matrices are OK)",
                     object, format, field, symm);
        if (!ieq(symm, "general") || !ieq(field, "integer"))
                errx(1, "Matrix type [%s %s %s %s] not supported (only integer general are OK)",
                     object, format, field, symm);

        /* skip comments and blank lines up to the size line */
        long long nr = 0, nc = 0, nz = 0;
        for (;;) {
                line = next_line(buf, len, &pos, &ll);
                if (!line)
                        errx(1, "Cannot read matrix size");
                if (ll == 0 || line[0] == '%')
                        continue;
                const char *p = line, *e = line + ll;
                if (!scan_int(&p, e, &nr) || !scan_int(&p, e, &nc) || !scan_int(&p, e, &nz))
                        continue;        /* mmio keeps scanning until three numbers parse */
                break;
        }
        if (nr <= 0 || nc <= 0 || nz < 0 || nr > 0x7fffffff || nc > 0x7fffffff)
                errx(1, "Cannot read matrix size");
        fprintf(stderr, "  - [matrix coordinate integer general] %lld x %lld with %lld nz\n", nr, nc, nz);
        fprintf(stderr, "  - Allocating %.1f MByte\n", 1e-6 * (12.0 * nz));

        size_t cnt = nz > 0 ? (size_t)nz : 1;
        int *Mi = malloc(cnt * sizeof(*Mi));
        int *Mj = malloc(cnt * sizeof(*Mj));
        uint32_t *Mx = malloc(cnt * sizeof(*Mx));
        if (!Mi || !Mj || !Mx)
                err(1, "Cannot allocate sparse matrix");

        const char *p = buf + pos, *end = buf + len;
        for (long long u = 0; u < nz; u++) {
                long long a, b, c;
                if (!scan_int(&p, end, &a) || !scan_int(&p, end, &b) || !scan_int(&p, end, &c))
                        errx(1, "parse error entry %lld\n", u);
                Mi[u] = (int)(a - 1);            /* MatrixMarket is 1-based */
                Mj[u] = (int)(b - 1);
                /* the reference reads "%d" into a u32 and then reduces: -k becomes 2^32-k (F9) */
                Mx[u] = (uint32_t)((uint64_t)(uint32_t)(int)c % prime);
        }
        double dt = now() - t0;
        printf("  - Read %s: %.1f MB in %.2fs (%.1f MB/s)\n", filename, len / 1048576., dt,
               len / 1048576. / (dt > 0 ? dt : 1e-9));
        free(buf);
        M->nrows = (int)nr; M->ncols = (int)nc; M->nnz = (long)nz;
        M->i = Mi; M->j = Mj; M->x = Mx;
}

void mtx_free(struct coo_matrix *M)
{
        free(M->i); free(M->j); free(M->x);
        M->i = M->j = NULL; M->x = NULL;
}

/* fast "%d\n" formatter for values below 2^31 */
static inline char *put_u32_line(char *w, uint32_t v)
{
        char tmp[12];
        int k = 0;
        int32_t s = (int32_t)v;                 /* the reference prints with %d */
        uint32_t a = s < 0 ? (uint32_t)(-(int64_t)s) : (uint32_t)s;
        if (s < 0)
                *w++ = '-';
        do {
                tmp[k++] = (char)('0' + a % 10);
                a /= 10;
        } while (a);
        while (k)
                *w++ = tmp[--k];
        *w++ = '\n';
        return w;
}

static void write_lines(FILE *f, const char *filename, long count, long stride, const uint32_t *v)
{
        enum { CHUNK = 1 << 16 };
        char *buf = malloc((size_t)CHUNK * 12);
        if (!buf)
                err(1, "cannot allocate output buffer");
        for (long base = 0; base < count; base += CHUNK) {
                long m = count - base < CHUNK ? count - base : CHUNK;
                char *w = buf;
                for (long t = 0; t < m; t++)
                        w = put_u32_line(w, v[(base + t) * stride]);
                if (fwrite(buf, 1, (size_t)(w - buf), f) != (size_t)(w - buf))
                        err(1, "cannot write %s", filename);
        }
        free(buf);
}

void kernel_block_save(const char *filename, int nrows, int n, const uint32_t *v)
{
        printf("Saving result in %s\n", filename);
        FILE *f = fopen(filename, "w");
        if (!f)
                err(1, "cannot open %s", filename);
        fprintf(f, "%%%%MatrixMarket matrix array integer general\n");
        fprintf(f, "%%block of left-kernel vector computed by lanczos_modp\n");
        fprintf(f, "%d %d\n", nrows, n);
        for (int col = 0; col < n; col++)           /* column-major, one value per line */
                write_lines(f, filename, nrows, n, v + col);
        if (fclose(f) != 0)
                err(1, "cannot write %s", filename);
}

void vector_save(const char *filename, long count, const uint32_t *v)
{
        char tmpname[4096];
        snprintf(tmpname, sizeof(tmpname), "%s.tmp", filename);
        FILE *f = fopen(tmpname, "w");
        if (!f)
                err(1, "cannot open %s", tmpname);
        printf("\t\t>> Making a snapshot of a vector in %s\n", filename);
        write_lines(f, tmpname, count, 1, v);
        if (fclose(f) != 0)
                err(1, "cannot write %s", tmpname);
        if (rename(tmpname, filename) != 0)
                err(1, "cannot rename %s to %s", tmpname, filename);
}

void vector_load(const char *filename, long count, uint32_t *v)
{
        size_t len;
        char *buf = slurp(filename, &len);
        const char *p = buf, *end = buf + len;
        long got = 0;
        long long val;
        while (scan_int(&p, end, &val)) {
                if (got >= count)
                        errx(1, "%s holds more than %ld values (matrix, --n or side differ from the checkpoint?)",
                             filename, count);
                v[got++] = (uint32_t)(int)val;
        }
        for (; got < count; got++)
                v[got] = 0;
        free(buf);
}
