// fake_fqd.cpp - TEST DOUBLE of the C ABI (include/fqd.h) for CPU tests of the drop-in binary's HOST logic
// (tests/test_cli_host_logic.py).  It is not a CPU path of the product: nothing outside tests/ builds, links or loads
// it, it announces itself on stderr, and it implements only what the ordered --fast driver calls, in the plainest way
// (newline counting + std::unordered_set of the sequence text).  The engine's real semantics - and everything about
// speed - are tested on the GPU against the oracle; this file exists so that the code AROUND the engine (readers,
// rings of blocks, tail carry, paired lock-step, restarts after capacity / row-width estimates, asynchronous writers,
// gzip in and out, the -v lines) runs in the CPU test-suite through the very same binary.
//   FAKE_FQD_SHRINK=k   pretend the key store holds cfg.max_records / k records (forces the restart path)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/fqd.h"

struct fqd_handle {
    fqd_config cfg;
    fqd_stats_t st;
    std::unordered_set<std::string> seen;
    std::vector<uint32_t> rec_start[2];
    std::vector<uint8_t> dup;
    uint64_t n_records = 0, capacity = 0;
    std::string err;
};

static std::string g_err;

extern "C" {

int fqd_abi_version(void) { return FQD_ABI_VERSION; }

int fqd_create(const fqd_config* cfg, fqd_handle** out) {
    if (!cfg || !out || cfg->abi_version != FQD_ABI_VERSION) { g_err = "bad config"; return FQD_ERR_INVALID; }
    static bool announced = false;
    if (!announced) { std::fprintf(stderr, "[fake_fqd] TEST DOUBLE of libfqd_cuda.so - host-logic tests only\n"); announced = true; }
    fqd_handle* h = new fqd_handle();
    h->cfg = *cfg;
    std::memset(&h->st, 0, sizeof h->st);
    const char* e = std::getenv("FAKE_FQD_SHRINK");
    const long long k = e ? std::atoll(e) : 1;
    h->capacity = cfg->max_records / (uint64_t)(k > 0 ? k : 1);
    *out = h;
    return FQD_OK;
}
void fqd_destroy(fqd_handle* h) { delete h; }
const char* fqd_last_error(const fqd_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }
int fqd_host_alloc(void** p, size_t bytes) { *p = std::malloc(bytes); return *p ? FQD_OK : FQD_ERR_CUDA; }
int fqd_host_free(void* p) { std::free(p); return FQD_OK; }
int fqd_stats(fqd_handle* h, fqd_stats_t* out) { if (!h || !out) return FQD_ERR_INVALID; *out = h->st; return FQD_OK; }

// complete records from the start of [p, p + n): offsets into rec (n_rec + 1 entries)
static void split(const char* p, size_t n, int lpr, std::vector<uint32_t>& rec) {
    rec.clear(); rec.push_back(0);
    const char* cur = p; const char* end = p + n;
    for (;;) {
        const char* q = cur; int lines = 0;
        while (lines < lpr && q < end) { const char* nl = (const char*)std::memchr(q, '\n', end - q); if (!nl) break; q = nl + 1; ++lines; }
        if (lines < lpr) break;
        cur = q;
        rec.push_back((uint32_t)(cur - p));
    }
}

static std::string seq_of(const char* rec, const char* rec_end) {
    const char* l1 = (const char*)std::memchr(rec, '\n', rec_end - rec) + 1;
    const char* l1e = (const char*)std::memchr(l1, '\n', rec_end - l1);
    return std::string(l1, l1e);
}

// The engine's contract for data errors (include/fqd.h, csrc/fqd_api.cu::fold_chunk), restated: the first error in the
// order in which the reference would meet it.  Fetching pair k pre-parses record k+1 of each mate (left first), so a
// malformed record e (bad first byte, or FASTQ sequence / quality lengths differ) aborts before pair e-1 is processed;
// a base outside {A,C,G,T,N} in pair j aborts while pair j is keyed (left mate first).
int fqd_push(fqd_handle* h, const char* r1, size_t n1, const char* r2, size_t n2, fqd_chunk_result* res) {
    if (!h || h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) { if (h) h->err = "the fake knows ordered --fast only"; return FQD_ERR_INVALID; }
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) { h->err = "chunk larger than max_chunk_bytes"; return FQD_ERR_INVALID; }
    const int mates = h->cfg.paired ? 2 : 1;
    const bool fasta = h->cfg.format == FQD_FORMAT_FASTA;
    const int lpr = fasta ? 2 : 4;
    const char lead = fasta ? '>' : '@';
    const char* buf[2] = {r1, r2}; const size_t len[2] = {n1, n2};
    for (int m = 0; m < mates; ++m) split(buf[m], len[m], lpr, h->rec_start[m]);
    size_t pairs = h->rec_start[0].size() - 1;
    if (mates == 2) pairs = std::min(pairs, h->rec_start[1].size() - 1);
    for (int m = 0; m < mates; ++m) h->rec_start[m].resize(pairs + 1);
    h->dup.assign(pairs, 0);
    size_t n_ok = pairs;
    const uint64_t first = h->n_records;
    if (!h->st.err) {
        bool have = false; long long best_t = 0; int code = 0, ch = 0, mate = 0; uint64_t rec = 0; size_t ok = 0;
        auto consider = [&](long long t, int c, int chr, int m, uint64_t r, size_t nk) {
            if (!have || t < best_t) { have = true; best_t = t; code = c; ch = chr; mate = m; rec = r; ok = nk; }
        };
        for (int m = 0; m < mates; ++m) {
            for (size_t e = 0; e <= pairs; ++e) {            // record `pairs` is the (possibly incomplete) one that follows
                const size_t off = h->rec_start[m][e];
                if (off >= len[m]) break;
                const long long t = e == 0 ? (first == 0 ? -8 + m : -4 + m) : (long long)(e - 1) * 4 + m;
                if (buf[m][off] != lead) { consider(t, FQD_ERR_BAD_START, (unsigned char)buf[m][off], m, first + e, e == 0 ? 0 : e - 1); break; }
                if (e == pairs) break;
                const char* rb = buf[m] + off; const char* re = buf[m] + h->rec_start[m][e + 1];
                const std::string s = seq_of(rb, re);
                if (!fasta) {
                    const char* l3 = rb; for (int k = 0; k < 3; ++k) l3 = (const char*)std::memchr(l3, '\n', re - l3) + 1;
                    const size_t qlen = (size_t)(re - l3) - 1;
                    if (qlen != s.size()) { consider(t, FQD_ERR_LEN_MISMATCH, 0, m, first + e, e == 0 ? 0 : e - 1); break; }
                }
                if (s.size() > h->cfg.max_seq_len) { consider(1ll << 60, FQD_ERR_SEQ_TOO_LONG, 0, m, 0, 0); break; }
                size_t bad = s.find_first_not_of("ACGTN");
                if (bad != std::string::npos) { consider((long long)e * 4 + 2 + m, FQD_ERR_BAD_BASE, (unsigned char)s[bad], m, first + e, e); break; }
            }
        }
        if (have) { h->st.err = code; h->st.err_char = ch; h->st.err_mate = mate; h->st.err_record = rec; n_ok = ok; }
    } else {
        n_ok = 0;
    }
    uint64_t dups = 0;
    for (size_t i = 0; i < n_ok; ++i) {
        if (first + i >= h->capacity) { if (!h->st.err) h->st.err = FQD_ERR_CAPACITY; n_ok = i; break; }
        std::string key;
        for (int m = 0; m < mates; ++m) { key += seq_of(buf[m] + h->rec_start[m][i], buf[m] + h->rec_start[m][i + 1]); key += '\n'; }
        if (!h->seen.insert(key).second) { h->dup[i] = 1; ++dups; }
    }
    h->st.total += n_ok; h->st.dups += dups;
    if (res) {
        std::memset(res, 0, sizeof *res);
        res->n_records = n_ok; res->first_record = first; res->n_survivors = n_ok - dups;
        res->dup = h->dup.data();
        for (int m = 0; m < mates; ++m) { res->rec_start[m] = h->rec_start[m].data(); res->consumed[m] = len[m] ? h->rec_start[m][pairs] : 0; }
    }
    h->n_records += pairs;
    return FQD_OK;
}

// whole-input modes: not in the fake
int fqd_append(fqd_handle* h, int, const char*, size_t) { if (h) h->err = "the fake knows ordered --fast only"; return FQD_ERR_INVALID; }
int fqd_finish(fqd_handle* h) { if (h) h->err = "the fake knows ordered --fast only"; return FQD_ERR_INVALID; }
int fqd_emit(fqd_handle*, int, void*, size_t, size_t*, int*) { return FQD_ERR_INVALID; }
int fqd_emit_clusters(fqd_handle*, int, void*, size_t, size_t*, int*) { return FQD_ERR_INVALID; }

}  // extern "C"
