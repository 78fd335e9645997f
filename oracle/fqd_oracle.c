/*
 * fqd_oracle.c - CPU restatement of fastq-dupaway's deduplication hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared against; nothing in the
 * product (fastq-dupaway_b200/, include/) may link, import or execute it.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against
 *   (1) every fixture pair of the reference's own test-suite (test/inputs <-> test/expected, copied as data
 *       into tests/golden/ref_fixtures), and
 *   (2) the UNMODIFIED reference sources compiled into oracle/_ref/fastq-dupaway[-stable] (oracle/Makefile)
 *       on seeded random FASTA/FASTQ inputs, SE and PE, every mode.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference root).
 * The code is a literal, sequential restatement: one block holding the whole input (the reference's
 * multi-block protocol yields the same record stream), base-5 packed keys and an exact
 * set for --fast, a STABLE sort + the stateful comparator scan for sequence mode (the
 * reference's std::sort is unstable; the stable order is what fastq-dupaway-stable produces), and the
 * single-block merge-join with its end-of-stream quirk for --unordered.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { FQDO_FASTQ = 0, FQDO_FASTA = 1 };
enum { FQDO_TIGHT = 1, FQDO_LOOSE = 2, FQDO_HAMMING = 3 };
enum {
    FQDO_OK = 0,
    FQDO_ERR_EMPTY = 1,      /* "Not enough memory to read a single object!"  src/bufferedinput.hpp:82-85 */
    FQDO_ERR_BAD_START = 2,  /* "Fastq/Fasta record should start with @/> symbol!" src/fastqview.cpp:121-126 */
    FQDO_ERR_LEN_MISMATCH = 3, /* seq/qual length mismatch  src/fastqview.cpp:117,128-138 */
    FQDO_ERR_BAD_BASE = 4,   /* "Supported sequence character set" src/seq_utils.cpp:17-19 */
    FQDO_ERR_NOMEM = 5
};

typedef struct {
    uint64_t total;      /* records / pairs processed (the -v line)              */
    uint64_t dups;       /* duplicates removed                                   */
    uint64_t unmatched;  /* --unordered: non-matching entries skipped            */
    int32_t  err;        /* FQDO_* code                                          */
    int32_t  err_char;   /* offending byte for BAD_START / BAD_BASE              */
    uint64_t err_record; /* index of the record that triggered the error         */
} fqdo_stats;

/* One parsed record: byte offsets into the input buffer.  Mirrors FastqView / FastaView members
 * (src/fastqview.hpp:44-46, src/fastaview.hpp:43-45): lengths INCLUDE the trailing '\n'. */
typedef struct {
    int64_t start, idlen, seqlen, f3len, quallen;
    int64_t tag_off, tag_len;   /* *ViewWithId only */
} rec_t;

static const char* find_nl(const char* p, const char* stop) {
    const char* q = (const char*)memchr(p, '\n', (size_t)(stop - p));
    return q ? q : stop;
}

/* FastqView::read_new (src/fastqview.cpp:89-119) and FastaView::read_new (src/fastaview.cpp:75-93).
 * Returns bytes consumed, -1 if the buffer ends before the record is complete, -2 / -3 on the two
 * validation errors (the reference throws there). */
static int64_t read_new(const char* base, const char* start, const char* stop, int format, rec_t* r, int* bad_char) {
    if (start >= stop) return -1;
    char lead = (format == FQDO_FASTQ) ? '@' : '>';
    if (*start != lead) { *bad_char = (unsigned char)*start; return -2; }
    const char* ptr = find_nl(start, stop);
    if (ptr == stop) return -1;
    r->start = start - base;
    r->idlen = ptr - start + 1;
    int64_t so_far = r->idlen;
    ++ptr;
    ptr = find_nl(ptr, stop);
    if (ptr == stop) return -1;
    r->seqlen = ptr - (start + so_far) + 1;
    so_far += r->seqlen;
    r->f3len = 0; r->quallen = 0;
    if (format == FQDO_FASTQ) {
        ++ptr;
        ptr = find_nl(ptr, stop);
        if (ptr == stop) return -1;
        r->f3len = ptr - (start + so_far) + 1;
        so_far += r->f3len;
        ++ptr;
        ptr = find_nl(ptr, stop);
        if (ptr == stop) return -1;
        r->quallen = ptr - (start + so_far) + 1;
        if (r->quallen != r->seqlen) return -3;
        so_far += r->quallen;
    }
    return so_far;
}

/* FastqViewWithId::read_new tag extraction (src/fastqview.cpp:190-204; src/fastaview.cpp:153-167). */
static void extract_tag(const char* base, rec_t* r) {
    const char* id = base + r->start;
    const char* end = id + r->idlen;
    const char* dot = (const char*)memchr(id, '.', (size_t)r->idlen);
    const char* tag = dot ? dot + 1 : id + 1;
    const char* sp = (tag < end) ? (const char*)memchr(tag, ' ', (size_t)(end - tag)) : NULL;
    if (!sp) sp = end;
    r->tag_off = tag - base;
    r->tag_len = sp - tag;
}

/* strncmp semantics (stops at NUL) + shorter-first, i.e. FastqView::cmp (src/fastqview.cpp:56-67) and
 * FastqViewWithId::cmp (src/fastqview.cpp:168-178). */
static int cmp_bytes(const char* a, int64_t la, const char* b, int64_t lb) {
    int64_t m = la < lb ? la : lb;
    int res = (m > 0) ? strncmp(a, b, (size_t)m) : 0;
    if (res == 0 && la < lb) return -1;
    if (res == 0 && la > lb) return 1;
    return res;
}

/* ---------------------------------------------------------------------------------------------------
 * BufferedInput<T> on one block holding the whole input (src/bufferedinput.hpp:57-103).
 * `cur` is pre-parsed; next() hands it out and parses the following record; block_end flips when that
 * parse fails.  Validation errors surface from refresh()/next() exactly where the reference throws. */
typedef struct {
    const char* buf; int64_t n; int format; int with_id;
    rec_t cur; int have_cur; int64_t pos; int block_end;
    int err; int err_char; uint64_t n_parsed;
} binput;

static int bi_parse(binput* b) {
    rec_t r; memset(&r, 0, sizeof r);
    int ch = 0;
    int64_t got = read_new(b->buf, b->buf + b->pos, b->buf + b->n, b->format, &r, &ch);
    if (got == -2) { b->err = FQDO_ERR_BAD_START; b->err_char = ch; return -2; }
    if (got == -3) { b->err = FQDO_ERR_LEN_MISMATCH; return -3; }
    if (got < 0) { b->have_cur = 0; return -1; }
    if (b->with_id) extract_tag(b->buf, &r);
    b->cur = r; b->have_cur = 1; b->pos += got; b->n_parsed++;
    return 0;
}
/* set_file + first refresh (src/bufferedinput.hpp:38-42,57-88) */
static int bi_open(binput* b, const char* buf, int64_t n, int format, int with_id) {
    memset(b, 0, sizeof *b);
    b->buf = buf; b->n = n; b->format = format; b->with_id = with_id;
    int rc = bi_parse(b);
    if (rc == -1) { b->err = FQDO_ERR_EMPTY; return -1; }
    return rc;
}
/* next() (src/bufferedinput.hpp:90-103): returns 0 and fills *out, or the (negative) validation error. */
static int bi_next(binput* b, rec_t* out) {
    *out = b->cur;
    int rc = bi_parse(b);
    if (rc == -1) { b->block_end = 1; return 0; }
    return rc;
}

/* ---------------------------------------------------------------------------------------------------
 * SeqUtils::_char2number / pattern2number / seq2hash (src/seq_utils.cpp:3-49, CHUNKSIZE src/seq_utils.hpp:9)
 * Returns the number of 64-bit words written, or -1 on an unsupported character (stored in *bad). */
#define FQDO_CHUNK 17
static int char2number(unsigned char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return -1; }
}
int64_t fqdo_seq2hash(const char* seq, int64_t len, uint64_t* out, int32_t* bad) {
    int64_t nchunks = len / FQDO_CHUNK;
    if (nchunks * FQDO_CHUNK < len) ++nchunks;
    for (int64_t i = 0; i < nchunks; ++i) {
        int64_t l = len - i * FQDO_CHUNK; if (l > FQDO_CHUNK) l = FQDO_CHUNK;
        uint64_t v = 0;
        for (int64_t k = 0; k < l; ++k) {
            int d = char2number((unsigned char)seq[i * FQDO_CHUNK + k]);
            if (d < 0) { if (bad) *bad = (unsigned char)seq[i * FQDO_CHUNK + k]; return -1; }
            v = 5 * v + (uint64_t)d;
        }
        out[i] = v;
    }
    return nchunks;
}

/* ---------------------------------------------------------------------------------------------------
 * Exact set of setRecord / setRecordPair keys (src/hash_dup_remover.hpp:19-71, src/hash_dup_remover.cpp:4-33).
 * A key is the word sequence [len1, words1..., (len2, words2...)] - equality on it is exactly operator==.
 * The bucket hash is irrelevant to the output; FNV-1a over the words is used. */
typedef struct { uint64_t* pool; size_t pool_len, pool_cap; uint64_t* slots; size_t nslots, count; } keyset;

static int ks_init(keyset* s) {
    memset(s, 0, sizeof *s);
    s->nslots = 1u << 16;
    s->slots = (uint64_t*)calloc(s->nslots, sizeof(uint64_t));
    s->pool_cap = 1u << 16;
    s->pool = (uint64_t*)malloc(s->pool_cap * sizeof(uint64_t));
    return (s->slots && s->pool) ? 0 : -1;
}
static void ks_free(keyset* s) { free(s->slots); free(s->pool); }
static uint64_t ks_hash(const uint64_t* k, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= k[i]; h *= 1099511628211ull; h ^= h >> 29; }
    return h;
}
/* pool entry layout: [nwords, words...]; slot value = pool offset + 1 */
static int ks_grow(keyset* s) {
    size_t nn = s->nslots * 2;
    uint64_t* ns = (uint64_t*)calloc(nn, sizeof(uint64_t));
    if (!ns) return -1;
    for (size_t i = 0; i < s->nslots; ++i) {
        uint64_t v = s->slots[i];
        if (!v) continue;
        const uint64_t* e = s->pool + (v - 1);
        size_t p = ks_hash(e + 1, (size_t)e[0]) & (nn - 1);
        while (ns[p]) p = (p + 1) & (nn - 1);
        ns[p] = v;
    }
    free(s->slots); s->slots = ns; s->nslots = nn;
    return 0;
}
/* returns 1 if inserted (key was new), 0 if already present, -1 on OOM */
static int ks_insert(keyset* s, const uint64_t* k, size_t n) {
    if ((s->count + 1) * 2 > s->nslots && ks_grow(s)) return -1;
    size_t p = ks_hash(k, n) & (s->nslots - 1);
    while (s->slots[p]) {
        const uint64_t* e = s->pool + (s->slots[p] - 1);
        if (e[0] == n && memcmp(e + 1, k, n * sizeof(uint64_t)) == 0) return 0;
        p = (p + 1) & (s->nslots - 1);
    }
    if (s->pool_len + n + 1 > s->pool_cap) {
        size_t nc = s->pool_cap * 2; while (nc < s->pool_len + n + 1) nc *= 2;
        uint64_t* np = (uint64_t*)realloc(s->pool, nc * sizeof(uint64_t));
        if (!np) return -1;
        s->pool = np; s->pool_cap = nc;
    }
    s->pool[s->pool_len] = n;
    memcpy(s->pool + s->pool_len + 1, k, n * sizeof(uint64_t));
    s->slots[p] = s->pool_len + 1;
    s->pool_len += n + 1;
    s->count++;
    return 1;
}

/* scratch key builder: setRecord(seq, seq_len-1) / setRecordPair (src/hash_dup_remover.cpp:4-8,16-24) */
typedef struct { uint64_t* w; size_t cap; } keybuf;
static int kb_reserve(keybuf* k, size_t n) {
    if (n <= k->cap) return 0;
    uint64_t* nw = (uint64_t*)realloc(k->w, n * sizeof(uint64_t));
    if (!nw) return -1;
    k->w = nw; k->cap = n; return 0;
}
static int64_t build_key(keybuf* kb, size_t at, const char* buf, const rec_t* r, int32_t* bad) {
    int64_t len = r->seqlen - 1;
    if (kb_reserve(kb, at + 2 + (size_t)(len / FQDO_CHUNK + 1))) return -2;
    kb->w[at] = (uint64_t)len;
    int64_t nw = fqdo_seq2hash(buf + r->start + r->idlen, len, kb->w + at + 1, bad);
    if (nw < 0) return -1;
    return nw + 1;
}

/* ---------------------------------------------------------------------------------------------------
 * HashDupRemover<T>::impl_filterSE (src/hash_dup_remover.hpp:105-148).
 * out_idx receives the indices (0-based, input order) of the records that are WRITTEN. */
int fqdo_fast_se(const char* buf, int64_t n, int format, uint64_t* out_idx, uint64_t* n_out, fqdo_stats* st) {
    memset(st, 0, sizeof *st); *n_out = 0;
    binput b; keyset ks; keybuf kb = {0, 0}; rec_t r; int32_t bad = 0;
    if (bi_open(&b, buf, n, format, 0)) { st->err = b.err; st->err_char = b.err_char; st->err_record = 0; return st->err; }
    if (ks_init(&ks)) { st->err = FQDO_ERR_NOMEM; return st->err; }
    uint64_t idx = 0;
    int first = 1;
    /* first record is written and inserted unconditionally (:118-124); then the block loop (:126-144).
     * With a single block the two are the same "fetch, test, write" step - except that the first record is WRITTEN
     * (:122) before it is keyed (:123), so a base outside {A,C,G,T,N} in record 0 leaves record 0 in the output. */
    while (first || !b.block_end) {
        first = 0;
        if (bi_next(&b, &r)) { st->err = b.err; st->err_char = b.err_char; st->err_record = b.n_parsed; break; }
        int64_t kw = build_key(&kb, 0, buf, &r, &bad);
        if (kw == -1) {
            if (idx == 0) out_idx[(*n_out)++] = 0;
            st->err = FQDO_ERR_BAD_BASE; st->err_char = bad; st->err_record = idx; break;
        }
        if (kw < 0) { st->err = FQDO_ERR_NOMEM; break; }
        st->total++;
        int ins = ks_insert(&ks, kb.w, (size_t)kw);
        if (ins < 0) { st->err = FQDO_ERR_NOMEM; break; }
        if (ins) out_idx[(*n_out)++] = idx; else st->dups++;
        idx++;
    }
    ks_free(&ks); free(kb.w);
    return st->err;
}

/* HashDupRemover<T>::impl_filterPE (src/hash_dup_remover.hpp:194-255): lock-stepped, stops at the shorter file. */
int fqdo_fast_pe(const char* buf1, int64_t n1, const char* buf2, int64_t n2, int format,
                 uint64_t* out_idx, uint64_t* n_out, fqdo_stats* st) {
    memset(st, 0, sizeof *st); *n_out = 0;
    binput b1, b2; keyset ks; keybuf kb = {0, 0}; rec_t l, r; int32_t bad = 0;
    if (bi_open(&b1, buf1, n1, format, 1)) { st->err = b1.err; st->err_char = b1.err_char; return st->err; }
    if (bi_open(&b2, buf2, n2, format, 1)) { st->err = b2.err; st->err_char = b2.err_char; return st->err; }
    if (ks_init(&ks)) { st->err = FQDO_ERR_NOMEM; return st->err; }
    uint64_t idx = 0; int first = 1;
    while (first || (!b1.block_end && !b2.block_end)) {
        first = 0;
        if (bi_next(&b1, &l)) { st->err = b1.err; st->err_char = b1.err_char; st->err_record = b1.n_parsed; break; }
        if (bi_next(&b2, &r)) { st->err = b2.err; st->err_char = b2.err_char; st->err_record = b2.n_parsed; break; }
        /* the first pair is written to both files (:220-221) before it is keyed (:222-227) */
        int64_t k1 = build_key(&kb, 0, buf1, &l, &bad);
        if (k1 == -1) { if (idx == 0) out_idx[(*n_out)++] = 0; st->err = FQDO_ERR_BAD_BASE; st->err_char = bad; st->err_record = idx; break; }
        int64_t k2 = (k1 < 0) ? k1 : build_key(&kb, (size_t)k1, buf2, &r, &bad);
        if (k2 == -1) { if (idx == 0) out_idx[(*n_out)++] = 0; st->err = FQDO_ERR_BAD_BASE; st->err_char = bad; st->err_record = idx; break; }
        if (k1 < 0 || k2 < 0) { st->err = FQDO_ERR_NOMEM; break; }
        st->total++;
        int ins = ks_insert(&ks, kb.w, (size_t)(k1 + k2));
        if (ins < 0) { st->err = FQDO_ERR_NOMEM; break; }
        if (ins) out_idx[(*n_out)++] = idx; else st->dups++;
        idx++;
    }
    ks_free(&ks); free(kb.w);
    return st->err;
}

/* ---------------------------------------------------------------------------------------------------
 * Stable merge sort of record indices.  The reference uses std::sort (unstable, src/external_sort.hpp:105,
 * src/paired_external_sort.hpp:135); "stable on input index" is the documented tie-break. */
typedef struct {
    const char* buf1; const rec_t* r1; const char* buf2; const rec_t* r2; int by_tag;
} sortctx;

static int rec_cmp(const sortctx* c, uint64_t a, uint64_t b) {
    if (c->by_tag)   /* FastqViewWithId::cmp src/fastqview.cpp:168-178 */
        return cmp_bytes(c->buf1 + c->r1[a].tag_off, c->r1[a].tag_len, c->buf1 + c->r1[b].tag_off, c->r1[b].tag_len);
    /* FastqView::cmp src/fastqview.cpp:56-67 on sequence INCLUDING '\n'; RecordPair::operator< src/paired_external_sort.hpp:20-26 */
    int v = cmp_bytes(c->buf1 + c->r1[a].start + c->r1[a].idlen, c->r1[a].seqlen,
                      c->buf1 + c->r1[b].start + c->r1[b].idlen, c->r1[b].seqlen);
    if (v || !c->r2) return v;
    return cmp_bytes(c->buf2 + c->r2[a].start + c->r2[a].idlen, c->r2[a].seqlen,
                     c->buf2 + c->r2[b].start + c->r2[b].idlen, c->r2[b].seqlen);
}
static void msort(const sortctx* c, uint64_t* a, uint64_t* tmp, size_t n) {
    if (n < 2) return;
    size_t h = n / 2;
    msort(c, a, tmp, h); msort(c, a + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (rec_cmp(c, a[j], a[i]) < 0) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(uint64_t));
}

/* Parse a whole file into a record table the way the sorters' read loops do
 * (src/external_sort.hpp:98-103, src/paired_external_sort.hpp:128-133). */
static int parse_all(const char* buf, int64_t n, int format, int with_id, rec_t** out, uint64_t* cnt, fqdo_stats* st) {
    binput b; *out = NULL; *cnt = 0;
    if (bi_open(&b, buf, n, format, with_id)) { st->err = b.err; st->err_char = b.err_char; return st->err; }
    size_t cap = 1024; rec_t* v = (rec_t*)malloc(cap * sizeof(rec_t));
    if (!v) { st->err = FQDO_ERR_NOMEM; return st->err; }
    while (!b.block_end) {
        if (*cnt == cap) { cap *= 2; rec_t* nv = (rec_t*)realloc(v, cap * sizeof(rec_t)); if (!nv) { free(v); st->err = FQDO_ERR_NOMEM; return st->err; } v = nv; }
        if (bi_next(&b, &v[*cnt])) { st->err = b.err; st->err_char = b.err_char; st->err_record = b.n_parsed; *out = v; return st->err; }
        (*cnt)++;
    }
    *out = v;
    return 0;
}

/* SeqUtils::hammingDistance (src/seq_utils.cpp:65-72) */
static uint64_t hamming(const char* a, const char* b, int64_t len) {
    uint64_t res = 0;
    for (int64_t i = 0; i < len; ++i) if (a[i] != b[i]) ++res;
    return res;
}
/* {Tight,Loose,Hamming}Comparator::compare single-mate forms (src/comparator.cpp:45-49,60-63,78-82);
 * h = stored head ("m_buf"), c = current record; lengths include '\n'. */
static int cmp1(int mode, uint32_t dist, const char* h, int64_t hl, const char* c, int64_t cl) {
    switch (mode) {
    case FQDO_TIGHT:   if (cl != hl) return 0; return strncmp(c, h, (size_t)cl) == 0;
    case FQDO_LOOSE: { int64_t m = (cl - 1 < hl - 1) ? cl - 1 : hl - 1; return m <= 0 ? 1 : strncmp(c, h, (size_t)m) == 0; }
    case FQDO_HAMMING: if (cl != hl) return 0; return hamming(h, c, cl) <= dist;
    }
    return 0;
}

/* SeqDupRemover<T>::filterSE/PE: sort (stable) then impl_filterSE/PE scan
 * (src/seq_dup_remover.hpp:40-109,111-218; comparators src/comparator.cpp:45-91).
 * out_idx = indices of WRITTEN records in EMISSION (sorted) order.  If cluster_of != NULL it receives, for
 * every record in sorted order position p, the input index of its cluster head (for --write-clusters) and
 * order_out the sorted permutation. */
int fqdo_seq(const char* buf1, int64_t n1, const char* buf2, int64_t n2, int format, int mode, uint32_t dist,
             uint64_t* out_idx, uint64_t* n_out, uint64_t* order_out, uint64_t* head_out, fqdo_stats* st) {
    memset(st, 0, sizeof *st); *n_out = 0;
    int paired = buf2 != NULL;
    rec_t *r1 = NULL, *r2 = NULL; uint64_t c1 = 0, c2 = 0;
    if (parse_all(buf1, n1, format, 0, &r1, &c1, st)) { free(r1); return st->err; }
    if (paired && parse_all(buf2, n2, format, 0, &r2, &c2, st)) { free(r1); free(r2); return st->err; }
    uint64_t n = paired ? (c1 < c2 ? c1 : c2) : c1;
    uint64_t* ord = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    uint64_t* tmp = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    if (!ord || !tmp) { free(r1); free(r2); free(ord); free(tmp); st->err = FQDO_ERR_NOMEM; return st->err; }
    for (uint64_t i = 0; i < n; ++i) ord[i] = i;
    sortctx c = { buf1, r1, paired ? buf2 : NULL, paired ? r2 : NULL, 0 };
    msort(&c, ord, tmp, (size_t)n);
    /* comparator scan: head = last record for which compare() returned false (set_seq), plus the loose
     * "keep the longest as reference" rule (src/seq_dup_remover.hpp:93-98,194-202). */
    uint64_t head = 0, head_written = 0;
    for (uint64_t p = 0; p < n; ++p) {
        uint64_t i = ord[p];
        int dup = 0;
        if (p > 0) {
            const char* hs1 = buf1 + r1[head].start + r1[head].idlen; int64_t hl1 = r1[head].seqlen;
            const char* cs1 = buf1 + r1[i].start + r1[i].idlen;       int64_t cl1 = r1[i].seqlen;
            dup = cmp1(mode, dist, hs1, hl1, cs1, cl1);
            if (dup && paired) {
                const char* hs2 = buf2 + r2[head].start + r2[head].idlen; int64_t hl2 = r2[head].seqlen;
                const char* cs2 = buf2 + r2[i].start + r2[i].idlen;       int64_t cl2 = r2[i].seqlen;
                dup = cmp1(mode, dist, hs2, hl2, cs2, cl2);
                if (dup && mode == FQDO_LOOSE)   /* same-sidedness src/comparator.cpp:72-73 */
                    dup = ((hl1 <= cl1) && (hl2 <= cl2)) || ((hl1 > cl1) && (hl2 > cl2));
            }
        }
        st->total++;
        if (!dup) {
            head = i; head_written = i;
            out_idx[(*n_out)++] = i;
        } else {
            st->dups++;
            if (mode == FQDO_LOOSE) {
                int longer = r1[head].seqlen <= r1[i].seqlen;
                if (paired) longer = longer && (r2[head].seqlen <= r2[i].seqlen);
                if (longer) head = i;
            }
        }
        if (order_out) order_out[p] = i;
        if (head_out) head_out[p] = head_written;
    }
    free(r1); free(r2); free(ord); free(tmp);
    return 0;
}

/* HashDupRemover<T>::filterPE(unordered) + impl_filterPE_unordered (src/hash_dup_remover.hpp:150-192,257-347):
 * both files sorted by ID tag (stable here), then the merge-join with the end-of-stream rule of
 * SURVEY.md F5 (single-block semantics, section 3.3).  out_idx1/out_idx2 = input indices of the written
 * R1 / R2 records, in emission order. */
int fqdo_fast_pe_unordered(const char* buf1, int64_t n1, const char* buf2, int64_t n2, int format,
                           uint64_t* out_idx1, uint64_t* out_idx2, uint64_t* n_out, fqdo_stats* st) {
    memset(st, 0, sizeof *st); *n_out = 0;
    rec_t *r1 = NULL, *r2 = NULL; uint64_t c1 = 0, c2 = 0;
    if (parse_all(buf1, n1, format, 1, &r1, &c1, st)) { free(r1); return st->err; }
    if (parse_all(buf2, n2, format, 1, &r2, &c2, st)) { free(r1); free(r2); return st->err; }
    uint64_t mx = c1 > c2 ? c1 : c2;
    uint64_t* o1 = (uint64_t*)malloc(c1 * sizeof(uint64_t));
    uint64_t* o2 = (uint64_t*)malloc(c2 * sizeof(uint64_t));
    uint64_t* tmp = (uint64_t*)malloc(mx * sizeof(uint64_t));
    keyset ks; keybuf kb = {0, 0}; int32_t bad = 0;
    if (!o1 || !o2 || !tmp || ks_init(&ks)) { st->err = FQDO_ERR_NOMEM; free(r1); free(r2); free(o1); free(o2); free(tmp); return st->err; }
    for (uint64_t i = 0; i < c1; ++i) o1[i] = i;
    for (uint64_t i = 0; i < c2; ++i) o2[i] = i;
    sortctx s1 = { buf1, r1, NULL, NULL, 1 }, s2 = { buf2, r2, NULL, NULL, 1 };
    msort(&s1, o1, tmp, (size_t)c1);
    msort(&s2, o2, tmp, (size_t)c2);
    uint64_t i = 0, j = 0;
    int last = 0;
    for (;;) {
        /* the loop condition of :279-281 in single-block form: run while neither side has fetched its last
         * record, then exactly one more comparison without advancing (:317-340). */
        if (!(i + 1 < c1 && j + 1 < c2)) last = 1;
        const rec_t* a = &r1[o1[i]]; const rec_t* b = &r2[o2[j]];
        int c = cmp_bytes(buf1 + a->tag_off, a->tag_len, buf2 + b->tag_off, b->tag_len);
        if (c < 0) { st->unmatched++; i++; }
        else if (c > 0) { st->unmatched++; j++; }
        else {
            int64_t k1 = build_key(&kb, 0, buf1, a, &bad);
            if (k1 == -1) { st->err = FQDO_ERR_BAD_BASE; st->err_char = bad; break; }
            int64_t k2 = (k1 < 0) ? k1 : build_key(&kb, (size_t)k1, buf2, b, &bad);
            if (k2 == -1) { st->err = FQDO_ERR_BAD_BASE; st->err_char = bad; break; }
            if (k1 < 0 || k2 < 0) { st->err = FQDO_ERR_NOMEM; break; }
            st->total++;
            int ins = ks_insert(&ks, kb.w, (size_t)(k1 + k2));
            if (ins < 0) { st->err = FQDO_ERR_NOMEM; break; }
            if (ins) { out_idx1[*n_out] = o1[i]; out_idx2[*n_out] = o2[j]; (*n_out)++; } else st->dups++;
            i++; j++;
        }
        if (last) break;
    }
    ks_free(&ks); free(kb.w); free(r1); free(r2); free(o1); free(o2); free(tmp);
    return st->err;
}

/* Record table export for the harness: offsets/lengths of each record as the reference's views see them.
 * fields per record: start, idlen, seqlen, f3len, quallen, tag_off, tag_len  (7 x int64). */
int fqdo_split(const char* buf, int64_t n, int format, int with_id, int64_t* table, uint64_t cap, uint64_t* cnt, fqdo_stats* st) {
    memset(st, 0, sizeof *st);
    rec_t* r = NULL; *cnt = 0;
    parse_all(buf, n, format, with_id, &r, cnt, st);   /* on a validation error: the records before it */
    for (uint64_t i = 0; i < *cnt && i < cap; ++i) {
        int64_t* t = table + 7 * i;
        t[0] = r[i].start; t[1] = r[i].idlen; t[2] = r[i].seqlen; t[3] = r[i].f3len; t[4] = r[i].quallen;
        t[5] = r[i].tag_off; t[6] = r[i].tag_len;
    }
    free(r);
    return st->err;
}
