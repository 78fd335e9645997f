// fake_fqd.cpp - TEST DOUBLE of the C ABI (include/fqd.h) for CPU tests of the drop-in binary's HOST logic
// (tests/test_cli_host_logic.py).  It is not a CPU path of the product: nothing outside tests/ builds, links or loads
// it, it announces itself on stderr, and it implements what the drivers call in the plainest way: ordered --fast by
// newline counting + a std::unordered_set of the sequence text, the whole-input modes (sequence-based, --fast
// --unordered, cluster files) by handing the collected input to the oracle's C functions (oracle/fqd_oracle.c, linked
// in - checker code serving a test, never the product).  The engine's real semantics - and everything about
// speed - are tested on the GPU against the oracle; this file exists so that the code AROUND the engine (readers,
// rings of blocks, tail carry, paired lock-step, restarts after capacity / row-width estimates, asynchronous writers,
// gzip in and out, the -v lines) runs in the CPU test-suite through the very same binary.
//   FAKE_FQD_SHRINK=k   pretend the key store holds cfg.max_records / k records (forces the restart path)
//   FAKE_FQD_REQUIRE_DISCARD=1   fqd_emit refuses: the test expects the host to take the discarded-input path
//   FAKE_FQD_DEVICE_BYTES=n   what fqd_device_memory reports as free (default 1 TiB): makes the host choose between the
//                       resident and the discarded-input path of the whole-input modes
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/fqd.h"

// the oracle's C interface (oracle/fqd_oracle.c)
extern "C" {
typedef struct { uint64_t total, dups, unmatched; int32_t err, err_char; uint64_t err_record; } fqdo_stats;
int fqdo_seq(const char* buf1, int64_t n1, const char* buf2, int64_t n2, int format, int mode, uint32_t dist,
             uint64_t* out_idx, uint64_t* n_out, uint64_t* order_out, uint64_t* head_out, fqdo_stats* st);
int fqdo_fast_pe_unordered(const char* buf1, int64_t n1, const char* buf2, int64_t n2, int format,
                           uint64_t* out_idx1, uint64_t* out_idx2, uint64_t* n_out, fqdo_stats* st);
int fqdo_split(const char* buf, int64_t n, int format, int with_id, int64_t* table, uint64_t cap, uint64_t* cnt, fqdo_stats* st);
}

struct fqd_handle {
    fqd_config cfg;
    fqd_stats_t st;
    // whole-input modes
    std::string in[2];
    bool finished = false;
    std::string out[2], clusters[2];     // what fqd_emit / fqd_emit_clusters hand out, whole records / lines at a time
    std::vector<size_t> out_cut[2], cl_cut[2];   // record / line boundaries inside them
    size_t out_pos[2] = {0, 0}, cl_pos[2] = {0, 0};
    bool discard = false;                        // fqd_discard_input: fqd_emit* refuse, the lists below are what the host gets
    std::vector<uint64_t> em_off[2], cl_off[2];
    std::vector<uint32_t> em_len[2], cl_len[2];
    std::vector<uint8_t> cl_head[2];
    size_t appended = 0;
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
    // like the real engine since round 2, the double grows its tables as it fills (max_records is a first size, not a
    // limit); FAKE_FQD_SHRINK makes it refuse instead, which is what the real engine does when the device is full
    h->capacity = e ? cfg->max_records / (uint64_t)(k > 0 ? k : 1) : ~0ull;
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

// ---- whole-input modes: collect, then let the oracle decide ---------------------------------------------------------
int fqd_append(fqd_handle* h, int mate, const char* buf, size_t n) {
    if (!h || mate < 0 || mate > 1 || h->finished) return FQD_ERR_INVALID;
    if (h->cfg.mode == FQD_MODE_FAST && !h->cfg.unordered) { h->err = "fqd_append is for the whole-input modes"; return FQD_ERR_INVALID; }
    h->in[mate].append(buf, n);
    h->appended += n;
    return FQD_OK;
}

static int map_err(int e) { static const int m[6] = {0, FQD_ERR_EMPTY, FQD_ERR_BAD_START, FQD_ERR_LEN_MISMATCH, FQD_ERR_BAD_BASE, FQD_ERR_CUDA}; return e >= 0 && e < 6 ? m[e] : FQD_ERR_CUDA; }

int fqd_finish(fqd_handle* h) {
    if (!h || h->finished) return FQD_ERR_INVALID;
    h->finished = true;
    const int mates = h->cfg.paired ? 2 : 1;
    const int fmt = h->cfg.format;
    const bool unordered = h->cfg.unordered != 0;
    // record tables (as far as the input parses)
    std::vector<int64_t> table[2]; uint64_t cnt[2] = {0, 0};
    int parse_err_mate = -1; uint64_t parse_err_record = 0;      // the oracle's statistics do not say which file
    for (int m = 0; m < mates; ++m) {
        const uint64_t cap = (uint64_t)std::count(h->in[m].begin(), h->in[m].end(), '\n') / 2 + 2;
        table[m].assign(cap * 7, 0);
        fqdo_stats ps;
        fqdo_split(h->in[m].data(), (int64_t)h->in[m].size(), fmt, unordered ? 1 : 0, table[m].data(), cap, &cnt[m], &ps);
        if (ps.err && parse_err_mate < 0) { parse_err_mate = m; parse_err_record = ps.err_record; }
    }
    // the limits a real engine is created with: the host starts over with larger ones
    for (int m = 0; m < mates; ++m) {
        if (cnt[m] > h->capacity) { h->st.err = FQD_ERR_CAPACITY; return FQD_OK; }
        for (uint64_t i = 0; i < cnt[m]; ++i) {
            const int64_t* t = &table[m][7 * i];
            if ((uint64_t)(t[2] - 1) > h->cfg.max_seq_len) { h->st.err = FQD_ERR_SEQ_TOO_LONG; return FQD_OK; }
            if (unordered && (uint64_t)t[6] > (h->cfg.max_tag_len ? h->cfg.max_tag_len : 32u)) { h->st.err = FQD_ERR_TAG_TOO_LONG; return FQD_OK; }
            if (!unordered && !h->cfg.byte_keys) {
                const char* q = h->in[m].data() + t[0] + t[1];
                for (int64_t k = 0; k + 1 < t[2]; ++k)
                    if (!std::strchr("ACGTN", q[k]) || q[k] == 0) { h->st.err = FQD_ERR_UNSUPPORTED_BYTE; h->st.err_char = (unsigned char)q[k]; return FQD_OK; }
            }
        }
    }
    const uint64_t cap = std::max(cnt[0], cnt[1]) + 2;
    std::vector<uint64_t> o1(cap), o2(cap), order(cap), head(cap);
    uint64_t n_out = 0; fqdo_stats st;
    if (unordered) {
        fqdo_fast_pe_unordered(h->in[0].data(), (int64_t)h->in[0].size(), h->in[1].data(), (int64_t)h->in[1].size(), fmt,
                               o1.data(), o2.data(), &n_out, &st);
    } else {
        fqdo_seq(h->in[0].data(), (int64_t)h->in[0].size(), mates == 2 ? h->in[1].data() : nullptr,
                 mates == 2 ? (int64_t)h->in[1].size() : 0, fmt, h->cfg.mode, h->cfg.hamming_dist,
                 o1.data(), &n_out, order.data(), head.data(), &st);
        o2 = o1;
    }
    h->st.total = st.total; h->st.dups = st.dups; h->st.unmatched = st.unmatched;
    h->st.err = map_err(st.err); h->st.err_char = st.err_char; h->st.err_record = st.err_record;
    if ((h->st.err == FQD_ERR_BAD_START || h->st.err == FQD_ERR_LEN_MISMATCH) && parse_err_mate >= 0) {
        h->st.err_mate = parse_err_mate; h->st.err_record = parse_err_record;      // index of the malformed record in its file
    }
    auto span = [&](int m, uint64_t i, size_t& off, size_t& len) {
        const int64_t* t = &table[m][7 * i];
        off = (size_t)t[0]; len = (size_t)(t[1] + t[2] + t[3] + t[4]);
    };
    for (int m = 0; m < mates; ++m) {
        const std::vector<uint64_t>& idx = m == 0 ? o1 : o2;
        h->out_cut[m].assign(1, 0);
        for (uint64_t k = 0; k < n_out; ++k) {
            size_t off, len; span(m, idx[k], off, len);
            h->out[m].append(h->in[m], off, len);
            h->out_cut[m].push_back(h->out[m].size());
            h->em_off[m].push_back(off); h->em_len[m].push_back((uint32_t)len);
        }
        h->cl_cut[m].assign(1, 0);
        if (!unordered && !h->st.err) {
            std::unordered_set<uint64_t> written(o1.begin(), o1.begin() + (ptrdiff_t)n_out);
            for (uint64_t p = 0; p < st.total; ++p) {
                const int64_t* t = &table[m][7 * order[p]];
                if (!written.count(order[p])) h->clusters[m] += "--";
                h->clusters[m].append(h->in[m], (size_t)t[0], (size_t)t[1]);
                h->cl_cut[m].push_back(h->clusters[m].size());
                h->cl_off[m].push_back((uint64_t)t[0]); h->cl_len[m].push_back((uint32_t)(t[1] + t[2] + t[3] + t[4]));
                h->cl_head[m].push_back(written.count(order[p]) ? 1 : 0);
            }
        }
    }
    return FQD_OK;
}

static int stream_out(const std::string& data, const std::vector<size_t>& cut, size_t& pos, void* dst, size_t cap, size_t* n_bytes, int* done) {
    // whole units (records / lines), as many as fit
    size_t a = pos, b = pos;
    while (b + 1 < cut.size() && cut[b + 1] - cut[a] <= cap) ++b;
    if (b == a && a + 1 < cut.size()) return FQD_ERR_INVALID;           // cap smaller than one unit
    *n_bytes = cut.empty() ? 0 : cut[b] - cut[a];
    if (*n_bytes) std::memcpy(dst, data.data() + cut[a], *n_bytes);
    pos = b;
    *done = b + 1 >= cut.size();
    return FQD_OK;
}
int fqd_emit(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done) {
    if (!h || !h->finished || mate < 0 || mate > 1) return FQD_ERR_INVALID;
    if (h->discard) { h->err = "[fake_fqd] fqd_emit after fqd_discard_input"; return FQD_ERR_INVALID; }
    if (std::getenv("FAKE_FQD_REQUIRE_DISCARD")) { h->err = "[fake_fqd] the test expected the discarded-input path"; return FQD_ERR_INVALID; }
    return stream_out(h->out[mate], h->out_cut[mate], h->out_pos[mate], dst, cap, n_bytes, done);
}
int fqd_emit_clusters(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done) {
    if (!h || !h->finished || mate < 0 || mate > 1) return FQD_ERR_INVALID;
    if (h->discard) { h->err = "[fake_fqd] fqd_emit_clusters after fqd_discard_input"; return FQD_ERR_INVALID; }
    return stream_out(h->clusters[mate], h->cl_cut[mate], h->cl_pos[mate], dst, cap, n_bytes, done);
}

// ---- the discarded-input path (include/fqd.h): lists instead of bytes -------------------------------------------------
int fqd_discard_input(fqd_handle* h, int on) {
    if (!h || h->appended) { if (h) h->err = "fqd_discard_input after the first fqd_append"; return FQD_ERR_INVALID; }
    h->discard = on != 0;
    return FQD_OK;
}
int fqd_emission_count(fqd_handle* h, uint64_t* n_written, uint64_t* n_sorted) {
    if (!h || !h->finished) return FQD_ERR_INVALID;
    if (n_written) *n_written = h->em_off[0].size();
    if (n_sorted) *n_sorted = h->cl_off[0].size();
    return FQD_OK;
}
int fqd_emission_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len) {
    if (!h || !h->finished || mate < 0 || mate > 1) return FQD_ERR_INVALID;
    const uint64_t n = h->em_off[mate].size();
    if (first > n || count > n - first) { h->err = "fqd_emission_read: window beyond the emission list"; return FQD_ERR_INVALID; }
    for (uint64_t i = 0; i < count; ++i) { off[i] = h->em_off[mate][first + i]; len[i] = h->em_len[mate][first + i]; }
    return FQD_OK;
}
int fqd_cluster_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len, uint8_t* head) {
    if (!h || !h->finished || mate < 0 || mate > 1) return FQD_ERR_INVALID;
    const uint64_t n = h->cl_off[mate].size();
    if (first > n || count > n - first) { h->err = "fqd_cluster_read: window beyond the records processed"; return FQD_ERR_INVALID; }
    for (uint64_t i = 0; i < count; ++i) { off[i] = h->cl_off[mate][first + i]; len[i] = h->cl_len[mate][first + i]; head[i] = h->cl_head[mate][first + i]; }
    return FQD_OK;
}
int fqd_device_memory(int, size_t* free_bytes, size_t* total_bytes) {
    const char* e = std::getenv("FAKE_FQD_DEVICE_BYTES");
    const size_t v = e ? (size_t)std::atoll(e) : (size_t)1 << 40;
    if (free_bytes) *free_bytes = v;
    if (total_bytes) *total_bytes = v;
    return FQD_OK;
}

// Several engines behind one binary (FQD_DEVICES) need real devices: the double refuses, loudly.
static int no_shards(fqd_handle* h) { if (h) h->err = "[fake_fqd] the test double has no sharded path"; return FQD_ERR_INVALID; }
int fqd_shard2_init(fqd_handle* h, uint32_t, uint32_t, uint32_t) { return no_shards(h); }
int fqd_shard2_link(fqd_handle* h, uint32_t, fqd_handle*) { return no_shards(h); }
int fqd_shard2_push_host(fqd_handle* h, uint64_t, const char*, size_t, const char*, size_t, uint64_t*, uint64_t*) { return no_shards(h); }
int fqd_shard2_insert(fqd_handle* h, uint64_t) { return no_shards(h); }
int fqd_shard2_apply(fqd_handle* h, uint64_t) { return no_shards(h); }
int fqd_shard2_result(fqd_handle* h, uint64_t, size_t, size_t, fqd_chunk_result*) { return no_shards(h); }
}  // extern "C"
