/*
 * qk_synth.c -- seeded synthetic inputs for the count path (bench + tests).
 *
 *   qk_synth ref   : random reference FASTA (contigs, segmental duplications, an N block)
 *   qk_synth ctrl  : a control-region BED over that reference (for `search -c`)
 *   qk_synth reads : short or HiFi-like reads sampled from a reference, FASTA or FASTQ
 *   qk_synth dict  : a QM11 dictionary (.qm [+ .qgc]) of the unique canonical k-mers of a
 *                    reference, in reference order -- the same key set and chain order
 *                    `quicKmer2 search -e 0` produces (Q.c:824-923, 1217-1299), built with
 *                    a different algorithm (parallel occurrence table, then a sequential
 *                    reference-order insert).  Slot placement follows the reference's
 *                    probe rule (Q.c:66-99) so the reference's own `count` can read it.
 *
 * The workloads follow SURVEY.md 8(d): iid uniform ACGT, reads = reference substrings at
 * uniform starts, odd-indexed reads reverse-complemented, per-base substitution errors.
 * Everything is a pure function of the seed (splitmix64 / xoshiro256**).
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ rng ----------- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed)
{
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static inline uint64_t rng_below(rng_t *r, uint64_t n)
{
    return (uint64_t)(((unsigned __int128)rng_next(r) * n) >> 64);
}
static inline double rng_unit(rng_t *r) { return (rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
/* distance to the next event of a Bernoulli(p) process, >= 1 */
static uint64_t rng_gap(rng_t *r, double p)
{
    if (p <= 0) return UINT64_MAX;
    double u = rng_unit(r);
    if (u <= 0) u = 1e-300;
    return (uint64_t)(log(u) / log1p(-p)) + 1;
}

/* ------------------------------------------------------------------ args ---------- */
static const char *arg_str(int argc, char **argv, const char *name, const char *dflt)
{
    for (int i = 2; i + 1 < argc; ++i)
        if (!strcmp(argv[i], name)) return argv[i + 1];
    return dflt;
}
static int arg_flag(int argc, char **argv, const char *name)
{
    for (int i = 2; i < argc; ++i)
        if (!strcmp(argv[i], name)) return 1;
    return 0;
}
static uint64_t arg_u64(int argc, char **argv, const char *name, uint64_t dflt)
{
    const char *s = arg_str(argc, argv, name, NULL);
    if (!s) return dflt;
    char *end;
    double v = strtod(s, &end);
    if (*end == 'K' || *end == 'k') v *= 1e3;
    else if (*end == 'M' || *end == 'm') v *= 1e6;
    else if (*end == 'G' || *end == 'g') v *= 1e9;
    return (uint64_t)(v + 0.5);
}

/* ------------------------------------------------------------------ fasta in ------ */
typedef struct {
    uint8_t *seq;       /* all contigs concatenated, newlines removed          */
    uint64_t n;         /* total bases                                         */
    uint64_t *start;    /* n_contigs + 1 offsets into seq                      */
    char **name;
    uint32_t n_contigs;
} genome_t;

static int genome_load(const char *path, genome_t *g)
{
    memset(g, 0, sizeof *g);
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "qk_synth: cannot open %s\n", path); return 1; }
    fseeko(f, 0, SEEK_END);
    off_t sz = ftello(f);
    fseeko(f, 0, SEEK_SET);
    uint8_t *raw = malloc((size_t)sz + 1);
    if (fread(raw, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return 1; }
    fclose(f);
    g->seq = malloc((size_t)sz + 1);
    uint32_t cap = 64;
    g->start = malloc((cap + 1) * sizeof(uint64_t));
    g->name = malloc(cap * sizeof(char *));
    size_t i = 0;
    while (i < (size_t)sz) {
        size_t e = i;
        while (e < (size_t)sz && raw[e] != '\n') ++e;
        if (raw[i] == '>') {
            if (g->n_contigs == cap) {
                cap *= 2;
                g->start = realloc(g->start, (cap + 1) * sizeof(uint64_t));
                g->name = realloc(g->name, cap * sizeof(char *));
            }
            g->start[g->n_contigs] = g->n;
            g->name[g->n_contigs] = strndup((char *)raw + i + 1, e - i - 1);
            g->n_contigs++;
        } else {
            memcpy(g->seq + g->n, raw + i, e - i);
            g->n += e - i;
        }
        i = e + 1;
    }
    g->start[g->n_contigs] = g->n;
    free(raw);
    return 0;
}

/* ------------------------------------------------------------------ ref ----------- */
static int cmd_ref(int argc, char **argv)
{
    const char *out = arg_str(argc, argv, "--out", NULL);
    uint64_t n = arg_u64(argc, argv, "--bases", 1000000);
    uint32_t contigs = (uint32_t)arg_u64(argc, argv, "--contigs", 1);
    uint64_t seed = arg_u64(argc, argv, "--seed", 1);
    uint64_t segdups = arg_u64(argc, argv, "--segdups", 0);
    uint64_t seglen = arg_u64(argc, argv, "--segdup-len", 20000);
    uint64_t div_ppm = arg_u64(argc, argv, "--divergence-ppm", 10000);
    uint64_t nblock = arg_u64(argc, argv, "--nblock", 0);
    uint32_t width = (uint32_t)arg_u64(argc, argv, "--line", 60);
    if (!out) { fprintf(stderr, "qk_synth ref: --out required\n"); return 1; }
    rng_t r;
    rng_seed(&r, seed);
    uint8_t *s = malloc(n);
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t i = 0; i < n; i += 32) {
        uint64_t bits = rng_next(&r);
        for (int j = 0; j < 32 && i + j < n; ++j, bits >>= 2) s[i + j] = acgt[bits & 3];
    }
    /* segmental duplications: copy [src, src+len) over [dst, dst+len) with divergence */
    for (uint64_t d = 0; d < segdups && n > 2 * seglen; ++d) {
        uint64_t src = rng_below(&r, n - seglen), dst = rng_below(&r, n - seglen);
        if (src + seglen > dst && dst + seglen > src) continue; /* overlapping: skip */
        memcpy(s + dst, s + src, seglen);
        double p = div_ppm * 1e-6;
        for (uint64_t pos = rng_gap(&r, p) - 1; pos < seglen; pos += rng_gap(&r, p))
            s[dst + pos] = acgt[(((s[dst + pos] >> 1) & 3) + 1 + rng_below(&r, 3)) & 3];
    }
    if (nblock && nblock < n / 2) {
        uint64_t at = n / 3;
        memset(s + at, 'N', nblock);
    }
    FILE *f = fopen(out, "wb");
    if (!f) return 1;
    uint64_t per = n / contigs;
    for (uint32_t c = 0; c < contigs; ++c) {
        uint64_t a = c * per, b = (c + 1 == contigs) ? n : a + per;
        fprintf(f, ">chr%u\n", c + 1);
        for (uint64_t i = a; i < b; i += width) {
            uint64_t w = b - i < width ? b - i : width;
            fwrite(s + i, 1, w, f);
            fputc('\n', f);
        }
    }
    fclose(f);
    free(s);
    return 0;
}

/* ------------------------------------------------------------------ ctrl ---------- */
static int cmd_ctrl(int argc, char **argv)
{
    const char *ref = arg_str(argc, argv, "--ref", NULL), *out = arg_str(argc, argv, "--out", NULL);
    uint64_t block = arg_u64(argc, argv, "--block", 100000);
    genome_t g;
    if (!ref || !out || genome_load(ref, &g)) return 1;
    FILE *f = fopen(out, "w");
    if (!f) return 1;
    for (uint32_t c = 0; c < g.n_contigs; ++c) {
        uint64_t len = g.start[c + 1] - g.start[c];
        for (uint64_t a = block; a + block <= len; a += 2 * block)
            fprintf(f, "%s\t%llu\t%llu\n", g.name[c], (unsigned long long)a, (unsigned long long)(a + block));
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ reads --------- */
static inline uint8_t comp(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;
    }
}

static int cmd_reads(int argc, char **argv)
{
    const char *ref = arg_str(argc, argv, "--ref", NULL), *out = arg_str(argc, argv, "--out", NULL);
    uint64_t n_reads = arg_u64(argc, argv, "--n", 1000);
    uint64_t len = arg_u64(argc, argv, "--len", 150);
    uint64_t seed = arg_u64(argc, argv, "--seed", 42);
    uint64_t err_ppm = arg_u64(argc, argv, "--err-ppm", 2000);
    int fastq = arg_flag(argc, argv, "--fastq");
    int hifi = arg_flag(argc, argv, "--hifi");          /* log-normal lengths, median --len */
    int randqual = arg_flag(argc, argv, "--rand-qual"); /* quality bytes over '!'..'J' incl '@','>' */
    int crlf = arg_flag(argc, argv, "--crlf");
    uint64_t lower_ppm = arg_u64(argc, argv, "--lower-ppm", 0); /* reads emitted in lower case */
    uint64_t max_len = arg_u64(argc, argv, "--max-len", 99998);
    uint64_t min_len = arg_u64(argc, argv, "--min-len", 1000);
    double sigma = arg_u64(argc, argv, "--sigma-milli", 500) * 1e-3;
    genome_t g;
    if (!ref || !out || genome_load(ref, &g)) return 1;
    FILE *f = fopen(out, "wb");
    if (!f) return 1;
    setvbuf(f, NULL, _IOFBF, 8 << 20);
    rng_t r;
    rng_seed(&r, seed);
    uint8_t *buf = malloc(max_len + 2), *qual = malloc(max_len + 2);
    memset(qual, 'I', max_len + 1);
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    double p = err_ppm * 1e-6;
    for (uint64_t i = 0; i < n_reads; ++i) {
        uint64_t L = len;
        if (hifi) {
            double u1 = rng_unit(&r), u2 = rng_unit(&r);
            if (u1 <= 0) u1 = 1e-300;
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            double v = (double)len * exp(sigma * z);
            L = v < (double)min_len ? min_len : v > (double)max_len ? max_len : (uint64_t)v;
        }
        /* contig chosen in proportion to its length, then a uniform start that fits */
        uint64_t at = rng_below(&r, g.n);
        uint32_t lo = 0, hi = g.n_contigs;
        while (hi - lo > 1) { uint32_t m = (lo + hi) / 2; if (g.start[m] <= at) lo = m; else hi = m; }
        uint64_t clen = g.start[lo + 1] - g.start[lo];
        if (clen < L) L = clen;
        at = g.start[lo] + rng_below(&r, clen - L + 1);
        if (i & 1) for (uint64_t j = 0; j < L; ++j) buf[j] = comp(g.seq[at + L - 1 - j]);
        else memcpy(buf, g.seq + at, L);
        for (uint64_t pos = rng_gap(&r, p) - 1; pos < L; pos += rng_gap(&r, p))
            if (buf[pos] != 'N') buf[pos] = acgt[(((buf[pos] >> 1) & 3) + 1 + rng_below(&r, 3)) & 3];
        if (lower_ppm && rng_below(&r, 1000000) < lower_ppm)
            for (uint64_t j = 0; j < L; ++j) buf[j] |= 0x20;
        if (randqual) for (uint64_t j = 0; j < L; ++j) qual[j] = (uint8_t)('!' + rng_below(&r, 42));
        const char *eol = crlf ? "\r\n" : "\n";
        fprintf(f, "%cr%llu%s", fastq ? '@' : '>', (unsigned long long)i, eol);
        fwrite(buf, 1, L, f);
        fputs(eol, f);
        if (fastq) {
            fprintf(f, "+%s", eol);
            fwrite(qual, 1, L, f);
            fputs(eol, f);
        }
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ dict ---------- */
/* canonical key stream of a contig, as `search` computes it (Q.c:845-864, 1003-1018) */
typedef struct { uint64_t fwd, rc, mask; uint32_t charge, k; } roll_t;
static inline void roll_reset(roll_t *s) { s->fwd = s->rc = 0; s->charge = 0; }
static inline int roll_push(roll_t *s, uint8_t c, uint64_t *key)
{
    if (c == 'N') { roll_reset(s); return 0; }
    uint64_t code = (c >> 1) & 3;
    s->fwd = (s->fwd << 2) | code;
    s->rc = (s->rc | (((code - 2) & 3) << 60)) >> 2;
    uint64_t km = s->fwd & s->mask;
    if (km > s->rc) km = s->rc;
    if (s->charge < s->k) s->charge++;
    *key = km;
    return km != 0 && s->charge == s->k;
}

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
    return x;
}
static inline uint64_t djb(uint64_t key)
{
    uint64_t h = 5381;
    for (int b = 0; b < 8; ++b, key >>= 8) h = h * 33 + (key & 0xFF);
    return h;
}

typedef struct {
    const genome_t *g; uint32_t k; uint64_t mask;
    uint64_t *tkeys; uint8_t *tocc; uint64_t tmask;
    uint32_t tid, nthreads;
} occ_job;

/* occurrence table: parallel insert over contig slices (each thread re-primes its window) */
static void *occ_worker(void *arg)
{
    occ_job *j = arg;
    const genome_t *g = j->g;
    for (uint32_t c = 0; c < g->n_contigs; ++c) {
        uint64_t a = g->start[c], b = g->start[c + 1], len = b - a;
        uint64_t lo = a + len * j->tid / j->nthreads, hi = a + len * (j->tid + 1) / j->nthreads;
        uint64_t prime = lo - a < 64 ? a : lo - 64; /* 64 >= 32 bases of history */
        roll_t s = {0, 0, j->mask, 0, j->k};
        /* the run length since the last N matters only up to k, so 64 bytes of warm-up suffice
         * unless the slice starts inside the first 64 bases (then we start at the contig start) */
        for (uint64_t i = prime; i < hi; ++i) {
            uint64_t key;
            int emit = roll_push(&s, g->seq[i], &key);
            if (i < lo || !emit) continue;
            uint64_t h = mix64(key) & j->tmask;
            for (;;) {
                uint64_t cur = __atomic_load_n(&j->tkeys[h], __ATOMIC_RELAXED);
                if (cur == 0) {
                    uint64_t exp = 0;
                    if (__atomic_compare_exchange_n(&j->tkeys[h], &exp, key, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                        cur = key;
                    else cur = exp;
                }
                if (cur == key) {
                    uint8_t o = __atomic_load_n(&j->tocc[h], __ATOMIC_RELAXED);
                    while (o < 2 && !__atomic_compare_exchange_n(&j->tocc[h], &o, (uint8_t)(o + 1), 0,
                                                               __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    break;
                }
                h = (h + 1) & j->tmask;
            }
        }
    }
    return NULL;
}

static int cmd_dict(int argc, char **argv)
{
    const char *ref = arg_str(argc, argv, "--ref", NULL);
    const char *ctrl = arg_str(argc, argv, "--ctrl-block", NULL); /* control = alternating blocks */
    uint32_t k = (uint32_t)arg_u64(argc, argv, "--k", 30);
    uint64_t slots = arg_u64(argc, argv, "--slots", 0);
    uint32_t nthreads = (uint32_t)arg_u64(argc, argv, "--threads", 8);
    const char *out = arg_str(argc, argv, "--out", ref); /* prefix: writes <out>.qm [.qgc] */
    genome_t g;
    if (!ref || genome_load(ref, &g)) return 1;
    uint64_t mask = ((uint64_t)1 << ((2 * k) & 63)) - 1;

    uint64_t tsize = 1;
    while (tsize < 2 * g.n + 16) tsize <<= 1;
    uint64_t *tkeys = calloc(tsize, 8);
    uint8_t *tocc = calloc(tsize, 1);
    if (!tkeys || !tocc) { fprintf(stderr, "qk_synth dict: out of memory\n"); return 1; }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64];
    occ_job jobs[64];
    for (uint32_t t = 0; t < nthreads; ++t) {
        jobs[t] = (occ_job){&g, k, mask, tkeys, tocc, tsize - 1, t, nthreads};
        pthread_create(&th[t], NULL, occ_worker, &jobs[t]);
    }
    for (uint32_t t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);

    /* pass 2: reference order; keep positions whose key occurs exactly once */
    uint64_t cap = g.n + 1, n_uniq = 0;
    uint64_t *okeys = malloc(cap * 8);
    uint16_t *ogc = malloc(cap * 2);
    uint64_t ctrl_block = ctrl ? arg_u64(argc, argv, "--ctrl-block", 100000) : 0;
    for (uint32_t c = 0; c < g.n_contigs; ++c) {
        uint64_t a = g.start[c], b = g.start[c + 1];
        roll_t s = {0, 0, mask, 0, k};
        /* GC bin of the 400-base window centred on the k-mer: 0..400 (ours, not Q.c:1025) */
        uint32_t gc = 0;
        const int64_t half_lead = (400 - (int64_t)k) / 2, half_trail = (400 + (int64_t)k) / 2;
        for (int64_t i = 0; i < half_lead && a + (uint64_t)i < b; ++i) gc += (g.seq[a + i] == 'G' || g.seq[a + i] == 'C');
        for (uint64_t i = a; i < b; ++i) {
            int64_t lead = (int64_t)(i - a) + half_lead, trail = (int64_t)(i - a) - half_trail;
            if (a + (uint64_t)lead < b) gc += (g.seq[a + lead] == 'G' || g.seq[a + lead] == 'C');
            if (trail >= 0) gc -= (g.seq[a + trail] == 'G' || g.seq[a + trail] == 'C');
            uint64_t key;
            if (!roll_push(&s, g.seq[i], &key)) continue;
            uint64_t h = mix64(key) & (tsize - 1);
            while (tkeys[h] != key) h = (h + 1) & (tsize - 1);
            if (tocc[h] != 1) continue;
            uint16_t v = (uint16_t)(gc > 400 ? 400 : gc);
            if (ctrl_block && (((i - a) / ctrl_block) & 1)) v |= 0x8000;
            okeys[n_uniq] = key;
            ogc[n_uniq] = v;
            ++n_uniq;
        }
    }
    free(tkeys);
    free(tocc);
    if (!slots) { slots = 1; while (slots < 2 * n_uniq + 2) slots <<= 1; }
    else { uint64_t p2 = 1; while (p2 < slots) p2 <<= 1; slots = p2; }
    if (slots > ((uint64_t)1 << 32) || n_uniq == 0 || n_uniq > slots * 8 / 10) {
        fprintf(stderr, "qk_synth dict: %llu unique k-mers do not fit %llu slots\n",
                (unsigned long long)n_uniq, (unsigned long long)slots);
        return 1;
    }
    /* reference-compatible placement (Q.c:90-99): home = djb & (H-1), walk toward the middle */
    uint64_t *keys = calloc(slots, 8);
    uint32_t *next = calloc(slots, 4);
    uint32_t first = 0, prev = 0;
    for (uint64_t i = 0; i < n_uniq; ++i) {
        uint64_t sidx = djb(okeys[i]) & (slots - 1);
        int64_t step = (sidx & (slots >> 1)) ? -1 : 1;
        while (keys[sidx] != 0) sidx = (uint64_t)((int64_t)sidx + step);
        keys[sidx] = okeys[i];
        if (i == 0) first = (uint32_t)sidx; else next[prev] = (uint32_t)sidx;
        prev = (uint32_t)sidx;
    }
    next[prev] = first;

    char path[4096];
    snprintf(path, sizeof path, "%s.qm", out);
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    uint8_t hdr[24] = {'Q', 'M', '1', '1', (uint8_t)k, 0, 100, 100};
    uint64_t first64 = first;
    memcpy(hdr + 8, &slots, 8);
    memcpy(hdr + 16, &first64, 8);
    fwrite(hdr, 1, 24, f);
    fwrite(keys, 8, slots, f);
    fwrite(next, 4, slots, f);
    fclose(f);
    if (ctrl) {
        snprintf(path, sizeof path, "%s.qgc", out);
        f = fopen(path, "wb");
        if (!f) return 1;
        fwrite(ogc, 2, n_uniq, f);
        fclose(f);
    }
    printf("{\"k\": %u, \"slots\": %llu, \"unique_kmers\": %llu, \"first\": %u}\n", k,
           (unsigned long long)slots, (unsigned long long)n_uniq, first);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2) {
        if (!strcmp(argv[1], "ref")) return cmd_ref(argc, argv);
        if (!strcmp(argv[1], "ctrl")) return cmd_ctrl(argc, argv);
        if (!strcmp(argv[1], "reads")) return cmd_reads(argc, argv);
        if (!strcmp(argv[1], "dict")) return cmd_dict(argc, argv);
    }
    fprintf(stderr, "usage: qk_synth {ref|ctrl|reads|dict} --opt value ...\n");
    return 1;
}
