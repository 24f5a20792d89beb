/*
 * CPU ORACLE (C restatement) -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of the reference's two hot stages, for CPU baseline timing
 * (bench.py cpu_baseline / --impl reference) and for oracle runs too large for
 * the Python restatement (oracle/phylo_oracle.py).  The shipped package never
 * links or loads this file.  Pinning: tests/test_oracle_c.py checks every function
 * here against oracle/phylo_oracle.py, which is itself pinned to the reference's
 * own outputs (tests/golden/).  KT: parity unpinned against real Biopython, see
 * the header of phylo_oracle.py.
 *
 * file:line citations point into /root/reference/phylopackage.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* minimal parallel-for over [0, n) with dynamic chunking (OpenMP is not in this image) */
typedef void (*row_fn)(int64_t i, void* ctx);
typedef struct { row_fn fn; void* ctx; int64_t n; int64_t next; int64_t chunk; pthread_mutex_t mu; } pf_t;
static void* pf_worker(void* arg) {
    pf_t* p = (pf_t*)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        int64_t lo = p->next;
        p->next += p->chunk;
        pthread_mutex_unlock(&p->mu);
        if (lo >= p->n) break;
        int64_t hi = lo + p->chunk < p->n ? lo + p->chunk : p->n;
        for (int64_t i = lo; i < hi; ++i) p->fn(i, p->ctx);
    }
    return 0;
}
static void parallel_for(int64_t n, int64_t chunk, int threads, row_fn fn, void* ctx) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pf_t p; p.fn = fn; p.ctx = ctx; p.n = n; p.next = 0; p.chunk = chunk > 0 ? chunk : 1;
    pthread_mutex_init(&p.mu, 0);
    pthread_t th[256];
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], 0, pf_worker, &p);
    pf_worker(&p);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], 0);
    pthread_mutex_destroy(&p.mu);
}

/* C,G,A,T -> 0..3 (bin/phyloligo.py:653 word order); anything else 255.
 * Lower case counts: the reference upper-cases first (bin/phyloligo.py:683). */
static int code_of(unsigned char c) {
    switch (c) {
        case 'C': case 'c': return 0;
        case 'G': case 'g': return 1;
        case 'A': case 'a': return 2;
        case 'T': case 't': return 3;
        default: return 255;
    }
}

/* select_strand (bin/phyloligo.py:124-149) + cut_sequence_and_count_pattern (:601-631).
 * strand: 0 plus, 1 minus, 2 both (= seq + revcomp(seq), no separator).
 * counts has 4^k entries and is overwritten.  Returns the number of words. */
uint64_t oracle_count(const unsigned char* seq, int64_t len, const char* pattern, int strand, uint64_t* counts) {
    const int width = (int)strlen(pattern);
    int ones[64], k = 0;
    for (int i = 0; i < width; ++i)
        if (pattern[i] == '1') ones[k++] = i;
    const int64_t dim = (int64_t)1 << (2 * k);
    memset(counts, 0, (size_t)dim * sizeof(uint64_t));
    const int64_t flen = strand == 2 ? 2 * len : len;
    unsigned char* s = (unsigned char*)malloc((size_t)(flen > 0 ? flen : 1));
    int64_t p = 0;
    if (strand == 0 || strand == 2)
        for (int64_t i = 0; i < len; ++i) s[p++] = (unsigned char)code_of(seq[i]);
    if (strand == 1 || strand == 2)
        for (int64_t i = len - 1; i >= 0; --i) {
            int c = code_of(seq[i]);
            s[p++] = (unsigned char)(c == 255 ? 255 : (c ^ 1)); /* complement: C<->G, A<->T */
        }
    uint64_t total = 0;
    int64_t run = 0; /* length of the current ACGT run ending at i */
    for (int64_t i = 0; i < flen; ++i) {
        run = (s[i] == 255) ? 0 : run + 1;
        if (run >= width) { /* window [i-width+1, i] lies inside one run */
            const unsigned char* w = s + (i - width + 1);
            int64_t word = 0;
            for (int j = 0; j < k; ++j) word = word * 4 + w[ones[j]];
            counts[word]++;
            total++;
        }
    }
    free(s);
    return total;
}

/* count2freq (bin/phyloligo.py:633-661): count/total as float64, zeros when total == 0 */
void oracle_freq(const uint64_t* counts, uint64_t total, int64_t dim, double* freq) {
    for (int64_t i = 0; i < dim; ++i) freq[i] = total ? (double)counts[i] / (double)total : 0.0;
}

/* Profile a batch of sequences given as one buffer + offsets; `threads` workers over records. */
typedef struct { const unsigned char* text; const int64_t* begin; const int64_t* end; const char* pattern;
                 int strand; int64_t dim; double* freq; } prof_ctx;
static void prof_row(int64_t r, void* vctx) {
    prof_ctx* c = (prof_ctx*)vctx;
    uint64_t* counts = (uint64_t*)malloc((size_t)c->dim * sizeof(uint64_t));
    uint64_t total = oracle_count(c->text + c->begin[r], c->end[r] - c->begin[r], c->pattern, c->strand, counts);
    oracle_freq(counts, total, c->dim, c->freq + r * c->dim);
    free(counts);
}
void oracle_profile_batch(const unsigned char* text, const int64_t* begin, const int64_t* end, int64_t n,
                          const char* pattern, int strand, double* freq /* n x dim */, int threads) {
    int k = 0;
    for (const char* q = pattern; *q; ++q) k += (*q == '1');
    prof_ctx c = {text, begin, end, pattern, strand, (int64_t)1 << (2 * k), freq};
    parallel_for(n, 8, threads, prof_row, &c);
}

/* Eucl, core/phylodist.py:36-41 */
double oracle_eucl(const double* a, const double* b, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double t = (a[i] - b[i]) * (a[i] - b[i]);
        if (isnan(t) || isinf(t)) t = 0.0;
        s += t;
    }
    return sqrt(s);
}

/* KL 1-D, core/phylodist.py:18-24: sum a ln(a/b), NaN/Inf terms -> 0 */
static double kl(const double* a, const double* h, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double t = a[i] * log(a[i] / h[i]);
        if (isnan(t) || isinf(t)) t = 0.0;
        s += t;
    }
    return s;
}

/* JSD 1-D, core/phylodist.py:43-48 */
double oracle_jsd(const double* a, const double* b, int64_t d) {
    double* h = (double*)malloc((size_t)d * sizeof(double));
    for (int64_t i = 0; i < d; ++i) h[i] = 0.5 * (a[i] + b[i]);
    double r = 0.5 * (kl(a, h, d) + kl(b, h, d));
    free(h);
    return r;
}

/* Bray-Curtis as scipy computes it (core/phylodist.py:79, bin/phyloligo.py:381) */
double oracle_bc(const double* a, const double* b, int64_t d) {
    double num = 0.0, den = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        num += fabs(a[i] - b[i]);
        den += fabs(a[i] + b[i]);
    }
    return num / den;
}

/* KT = 1 - kendall distance of the C Clustering Library (core/phylodist.py:71-74) */
double oracle_kt(const double* a, const double* b, int64_t d) {
    int64_t con = 0, dis = 0, exx = 0, exy = 0;
    int flag = 0;
    for (int64_t i = 0; i < d; ++i)
        for (int64_t j = 0; j < i; ++j) {
            double x1 = a[i], x2 = a[j], y1 = b[i], y2 = b[j];
            if (x1 < x2 && y1 < y2) con++;
            if (x1 > x2 && y1 > y2) con++;
            if (x1 < x2 && y1 > y2) dis++;
            if (x1 > x2 && y1 < y2) dis++;
            if (x1 == x2 && y1 != y2) exx++;
            if (x1 != x2 && y1 == y2) exy++;
            flag = 1;
        }
    if (!flag) return 1.0 - 0.0;
    double denomx = (double)(con + dis + exx), denomy = (double)(con + dis + exy);
    if (denomx == 0 || denomy == 0) return 1.0 - 1.0;
    double tau = (double)(con - dis) / sqrt(denomx * denomy);
    return 1.0 - (1.0 - tau);
}

static void ranks(const double* a, int64_t d, double* r) {
    for (int64_t i = 0; i < d; ++i) {
        int64_t less = 0, eq = 0;
        for (int64_t j = 0; j < d; ++j) {
            less += a[j] < a[i];
            eq += a[j] == a[i];
        }
        r[i] = (double)less + 0.5 * (double)(eq + 1);
    }
}

/* SC = 1 - Spearman rho (intended meaning of core/phylodist.py:82-85) */
double oracle_sc(const double* a, const double* b, int64_t d) {
    double* ra = (double*)malloc((size_t)d * sizeof(double));
    double* rb = (double*)malloc((size_t)d * sizeof(double));
    ranks(a, d, ra);
    ranks(b, d, rb);
    const double mean = 0.5 * (double)(d + 1);
    double sab = 0, saa = 0, sbb = 0;
    for (int64_t i = 0; i < d; ++i) {
        double x = ra[i] - mean, y = rb[i] - mean;
        sab += x * y; saa += x * x; sbb += y * y;
    }
    free(ra); free(rb);
    double den = sqrt(saa * sbb);
    if (den == 0.0) return NAN;
    return 1.0 - sab / den;
}

/* metric: 0 Eucl, 1 JSD, 2 KT, 3 BC, 4 SC.  Rows [r0, r1) of the full matrix, all columns. */
typedef struct { int metric; const double* X; int64_t n, d, r0; double* out; } pw_ctx;
static void pw_row(int64_t ii, void* vctx) {
    pw_ctx* c = (pw_ctx*)vctx;
    const int64_t i = c->r0 + ii;
    for (int64_t j = 0; j < c->n; ++j) {
        const double* a = c->X + i * c->d;
        const double* b = c->X + j * c->d;
        double v;
        switch (c->metric) {
            case 0: v = oracle_eucl(a, b, c->d); break;
            case 1: v = oracle_jsd(a, b, c->d); break;
            case 2: v = oracle_kt(a, b, c->d); break;
            case 3: v = oracle_bc(a, b, c->d); break;
            default: v = oracle_sc(a, b, c->d); break;
        }
        c->out[ii * c->n + j] = v;
    }
}
void oracle_pairwise_rows(int metric, const double* X, int64_t n, int64_t d, int64_t r0, int64_t r1, double* out,
                          int threads) {
    pw_ctx c = {metric, X, n, d, r0, out};
    parallel_for(r1 - r0, 1, threads, pw_row, &c);
}
