// dup_remover.cpp - the two drivers of the drop-in binary on top of the C ABI (include/fqd.h).
#include "dup_remover.hpp"

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <iostream>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../../include/fqd.h"
#include "io.hpp"
#include "replay.hpp"

namespace fqdhost {
namespace {

struct EngineDeleter { void operator()(fqd_handle* h) const { fqd_destroy(h); } };
using EnginePtr = std::unique_ptr<fqd_handle, EngineDeleter>;

// Everything a --fast run holds, in construction order (destroyed in reverse: engine, writers - which finish what
// is queued -, readers and their pinned blocks, output files).
struct OrderedJob {
    std::vector<std::unique_ptr<OutputFile>> outs;
    MateStream ms[2];
    std::vector<std::unique_ptr<AsyncWriter>> writers;
    EnginePtr eng;
};

// After a successful run the outputs are closed and nothing else needs an orderly end: unpinning gigabytes of
// staging memory and freeing the device allocations one by one costs 0.2 - 2 s that process exit does for free
// (main() leaves through _exit).  FQD_ORDERLY_EXIT=1 keeps the destructors (leak checkers).
template <class T> void leave_to_the_os(std::unique_ptr<T>& job) {
    static const bool orderly = std::getenv("FQD_ORDERLY_EXIT") != nullptr;
    if (!orderly) (void)job.release();
}

struct WholeJob {
    std::vector<std::unique_ptr<BlockReader>> readers;
    std::vector<std::shared_ptr<InputReplay>> replay;      // discarded-input path only
    EnginePtr eng;
};

struct Restart : std::exception { int what_code; explicit Restart(int c) : what_code(c) {} };

[[noreturn]] void throw_engine_error(fqd_handle* h, int rc) {
    std::string msg = "CUDA engine failure";
    const char* e = fqd_last_error(h);
    if (e && *e) msg += std::string(": ") + e;
    msg += " (status " + std::to_string(rc) + ")";
    throw std::runtime_error(msg);
}

// The reference's messages for data errors (src/seq_utils.cpp:17-19, src/fastqview.cpp:121-138,
// src/fastaview.cpp:95-100, src/bufferedinput.hpp:82-85).
[[noreturn]] void throw_data_error(const fqd_stats_t& st, bool fasta, const char* rec = nullptr, size_t rec_len = 0) {
    switch (st.err) {
    case FQD_ERR_BAD_BASE:
        std::cerr << "Error: unknown character in DNA sequence: " << (char)st.err_char << '\n';
        throw std::runtime_error("Supported sequence character set: {A, N, C, G, T}!");
    case FQD_ERR_BAD_START:
        std::cerr << "Invalid record start character: " << (char)st.err_char << std::endl;
        throw std::runtime_error(fasta ? "Fasta record should start with > symbol!" : "Fastq record should start with @ symbol!");
    case FQD_ERR_LEN_MISMATCH: {
        if (rec) {   // "Found sequence <seq> of length <n> and quality string <qual> of length <m>"
            const char* end = rec + rec_len;
            const char* l[5] = {rec, nullptr, nullptr, nullptr, nullptr};
            for (int i = 1; i < 5; ++i) {
                const char* nl = l[i - 1] ? (const char*)memchr(l[i - 1], '\n', end - l[i - 1]) : nullptr;
                l[i] = nl ? nl + 1 : nullptr;
            }
            if (l[4]) {
                std::cerr << "Found sequence ";
                std::cerr.write(l[1], l[2] - l[1] - 1);
                std::cerr << " of length " << (l[2] - l[1]) << " and quality string ";
                std::cerr.write(l[3], l[4] - l[3] - 1);
                std::cerr << " of length " << (l[4] - l[3]) << std::endl;
            }
        }
        throw std::runtime_error("Sequence and Quality fields of Fastq record should have the same length!");
    }
    case FQD_ERR_EMPTY:
        throw std::runtime_error("Not enough memory to read a single object!");
    default:
        throw std::runtime_error("Data error " + std::to_string(st.err));
    }
}

// FQD_BLOCK_BYTES: staging block size for tests (many small blocks exercise the tail carry and the recycling of
// blocks behind the asynchronous writers on small inputs)
size_t test_block_override(size_t block) {
    const char* e = std::getenv("FQD_BLOCK_BYTES");
    if (!e) return block;
    long long v = std::atoll(e);
    return v >= 4096 ? ((size_t)v & ~(size_t)4095) : block;
}

size_t file_size_or_zero(const std::string& name) {
    struct stat sb;
    return stat(name.c_str(), &sb) == 0 ? (size_t)sb.st_size : 0;
}

// A pipe, a FIFO, /dev/stdin, `<(zcat x.gz)`: can be read once, has no size.
bool rereadable(const std::string& name) {
    struct stat sb;
    return stat(name.c_str(), &sb) == 0 && S_ISREG(sb.st_mode);
}

// Restarting a job means reading the input again (wider key rows; tables the engine could not grow in place).
void check_restart_possible(const std::string* in, int mates, const char* why) {
    for (int m = 0; m < mates; ++m)
        if (!rereadable(in[m]))
            throw std::runtime_error(std::string("input ") + in[m] + " is not a regular file and cannot be read a second time (" + why +
                                     "); write it to a file first");
}

void check_outputs(const std::vector<std::unique_ptr<OutputFile>>& outs, const std::string* names) {
    for (size_t m = 0; m < outs.size(); ++m)
        if (int e = outs[m]->error())
            throw std::runtime_error("writing " + names[m] + " failed: " + std::strerror(e));
}

// How much larger than the file is its content?  Plain: 1.  ".gz": the ratio measured on what the reader has
// inflated by now (the first block is there, i.e. tens of megabytes) plus 10 %, else a generous 8 (a wrong guess
// costs device memory or a restart with doubled tables, never the result).
double gz_expansion(const std::string& name, const BlockReader& reader) {
    if (!has_gz_ext(name)) return 1.0;
    const double h = reader.expansion_hint();
    return h > 0 ? std::max(1.0, h * 1.1) : 8.0;
}

// Record `index` (0-based) of a file, read again from disk: the whole-input modes stream their input to the device and
// keep nothing, but the reference's message for a sequence / quality length mismatch quotes both strings
// (src/fastqview.cpp:128-138).  Error path only.
std::string fetch_record(const std::string& name, uint64_t index, int lpr) {
    InputFile f(name);
    std::vector<char> buf(1u << 20);
    std::string rec;
    uint64_t lines_to_skip = index * (uint64_t)lpr;
    int lines_wanted = lpr;
    while (!f.eof() && lines_wanted > 0) {
        const size_t n = f.read(buf.data(), buf.size());
        const char* p = buf.data(); const char* end = p + n;
        while (p < end && lines_wanted > 0) {
            const char* nl = (const char*)memchr(p, '\n', end - p);
            const char* stop = nl ? nl + 1 : end;
            if (lines_to_skip == 0) rec.append(p, stop);
            if (nl) { if (lines_to_skip) --lines_to_skip; else --lines_wanted; }
            p = stop;
        }
    }
    return rec;
}

// Longest sequence line and mean record size of a sample (first bytes of the first block).
void sample_geometry(const char* p, size_t n, int lpr, size_t& max_seq, double& avg_rec) {
    max_seq = 0; avg_rec = 0;
    size_t line = 0, recs = 0, last_rec_end = 0;
    const char* cur = p; const char* end = p + n;
    while (cur < end) {
        const char* nl = (const char*)memchr(cur, '\n', end - cur);
        if (!nl) break;
        if ((int)(line % lpr) == 1) max_seq = std::max<size_t>(max_seq, nl - cur);
        ++line;
        if ((int)(line % lpr) == 0) { ++recs; last_rec_end = nl + 1 - p; }
        cur = nl + 1;
    }
    if (recs) avg_rec = (double)last_rec_end / recs; else { avg_rec = 64; max_seq = std::max<size_t>(max_seq, n); }
}

// the -v lines of the whole-input modes (src/seq_dup_remover.hpp:107-108,216-217; src/hash_dup_remover.hpp:344-346)
void print_whole_summary(const fqd_stats_t& st, bool unordered, int mates, bool verbose) {
    if (!verbose) return;
    if (unordered) {
        std::cout << st.total << " valid read pairs processed, out of which " << st.dups << " duplicates were removed.\n";
        std::cout << st.unmatched << " Non-matching entries from both files were skipped.\n";
    } else if (mates == 1) {
        std::cout << st.total << " reads processed, out of which " << st.dups << " duplicates were removed.\n";
    } else {
        std::cout << st.total << " read pairs processed, out of which " << st.dups << " duplicates were removed.\n";
    }
}

// FQD_WHOLE_INPUT=resident | discard: where the raw bytes of a whole-input job live while it is sorted (default: decided
// from the input size and the device's free memory, see run_whole_input)
int whole_input_policy() {
    const char* e = std::getenv("FQD_WHOLE_INPUT");
    if (!e) return 0;
    if (!std::strcmp(e, "resident")) return 1;
    if (!std::strcmp(e, "discard")) return 2;
    return 0;
}

// Device bytes one record (pair) occupies without its raw text: key row(s), (offset, length, sequence length) per
// mate, tags / hashes of --unordered, and the sort's ping-pong keys and index lists.
size_t resident_bytes_per_record(const fqd_config& cfg, int mates) {
    const size_t row = cfg.byte_keys ? ((size_t)cfg.max_seq_len + 1 + 7) / 8 * 8 : (((size_t)cfg.max_seq_len * 3 + 63) / 64 + 1) * 8;
    size_t b = (size_t)mates * (row + 16) + 64;
    if (cfg.unordered) b += (size_t)mates * (((cfg.max_tag_len ? cfg.max_tag_len : 32) + 7) / 8 * 8 + 12) + 2 * row + 16;
    return b;
}

// FQD_LIST_WINDOW: records per list window (test knob: small jobs then cross many windows)
uint64_t list_window() {
    const char* e = std::getenv("FQD_LIST_WINDOW");
    const long long v = e ? std::atoll(e) : 0;
    return v > 0 ? (uint64_t)v : (uint64_t)1 << 20;
}

// The written records of one mate, fetched from the host's copy of the input in emission order: the lists come from
// the device window by window, a window's records are gathered and written by the output file's workers (write_runs:
// several at once, each at the file offset its bytes belong to) while the next window's lists arrive.
void gather_from_replay(fqd_handle* eng, int m, uint64_t n_written, const InputReplay& src, OutputFile& out) {
    const uint64_t WIN = list_window();
    std::vector<uint64_t> off((size_t)std::min<uint64_t>(WIN, std::max<uint64_t>(n_written, 1)));
    std::vector<uint32_t> len(off.size());
    std::mutex mu; std::condition_variable cv; int in_flight = 0;
    AsyncWriter writer(out);
    for (uint64_t k = 0; k < n_written; k += WIN) {
        const uint64_t c = std::min(WIN, n_written - k);
        int rc = fqd_emission_read(eng, m, k, c, off.data(), len.data());
        if (rc) { writer.drain(); throw_engine_error(eng, rc); }
        std::vector<Run> runs;
        runs.reserve((size_t)c);
        size_t total = 0;
        for (uint64_t i = 0; i < c; ++i) {
            if (off[i] > src.size() || len[i] > src.size() - off[i]) { writer.drain(); throw std::runtime_error("the engine returned a record outside the input"); }
            if (!runs.empty() && runs.back().off + runs.back().len == off[i]) runs.back().len += len[i];      // neighbours in the input too
            else runs.push_back(Run{(size_t)off[i], len[i], total});
            total += len[i];
        }
        { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return in_flight < 3; }); ++in_flight; }      // ~24 MB of runs each
        writer.write_runs(src.data(), std::move(runs), total);
        writer.then([&mu, &cv, &in_flight] { { std::lock_guard<std::mutex> g(mu); --in_flight; } cv.notify_all(); });
    }
    writer.drain();
}

// <out>.clusters from the host's copy of the input (src/seq_dup_remover.hpp:60-62,75-76,89-101; src/file_utils.cpp:98-112):
// one line per record in sorted order, the ID line of a written record, "--" + the ID line of a removed one.
void clusters_from_replay(fqd_handle* eng, int m, uint64_t n_sorted, const InputReplay& src, const std::string& out_name) {
    std::ofstream cf(out_name + ".clusters", std::ios::binary);
    const uint64_t WIN = list_window();
    std::vector<uint64_t> off((size_t)std::min<uint64_t>(WIN, std::max<uint64_t>(n_sorted, 1)));
    std::vector<uint32_t> len(off.size());
    std::vector<uint8_t> head(off.size());
    std::string text;
    for (uint64_t k = 0; k < n_sorted; k += WIN) {
        const uint64_t c = std::min(WIN, n_sorted - k);
        int rc = fqd_cluster_read(eng, m, k, c, off.data(), len.data(), head.data());
        if (rc) throw_engine_error(eng, rc);
        text.clear();
        for (uint64_t i = 0; i < c; ++i) {
            if (off[i] > src.size() || len[i] > src.size() - off[i]) throw std::runtime_error("the engine returned a record outside the input");
            const char* p = src.data() + off[i];
            const char* nl = (const char*)memchr(p, '\n', len[i]);
            if (!head[i]) text += "--";
            text.append(p, nl ? (size_t)(nl - p) + 1 : (size_t)len[i]);
        }
        cf.write(text.data(), (std::streamsize)text.size());
    }
}

}  // namespace

// -------------------------------------------------------------------------------------------------------------
// HashDupRemover: --fast, input order preserved (src/hash_dup_remover.hpp:105-148,194-255)
void HashDupRemover::filterSE(const std::string& infile, const std::string& outfile) {
    const std::string in[1] = {infile}, out[1] = {outfile};
    run_ordered(in, out, 1);
}

void HashDupRemover::filterPE(const std::string& infile1, const std::string& infile2,
                              const std::string& outfile1, const std::string& outfile2, bool unordered) {
    const std::string in[2] = {infile1, infile2}, out[2] = {outfile1, outfile2};
    if (unordered) run_whole_input(FQD_MODE_FAST, m_fasta, true, 0, 2, in, out, false, m_verbose, m_memlimit, m_device);
    else run_ordered(in, out, 2);
}

// FQD_DEVICES="0,1,2,3" (CUDA ordinals; the same ordinal may be listed twice - two shards on one GPU, for tests)
static std::vector<int> devices_from_env() {
    std::vector<int> d;
    const char* e = std::getenv("FQD_DEVICES");
    if (!e) return d;
    for (const char* p = e; *p;) {
        char* end = nullptr;
        const long v = std::strtol(p, &end, 10);
        if (end == p) break;
        d.push_back((int)v);
        p = *end == ',' ? end + 1 : end;
    }
    return d;
}

void HashDupRemover::run_ordered(const std::string* in, const std::string* out, int mates) {
    {
        const std::vector<int> devs = devices_from_env();
        if (devs.size() > 1) { run_ordered_multi(in, out, mates, devs); return; }
    }
    const int lpr = m_fasta ? 2 : 4;
    const char lead = m_fasta ? '>' : '@';
    // pinned staging: 3 blocks (head room + data) per input file inside the -m budget
    size_t block = (size_t)m_memlimit / (size_t)(mates * 3 * 2);
    block = std::min<size_t>(std::max<size_t>(block, 4u << 20), 64u << 20) & ~(size_t)4095;
    block = test_block_override(block);
    double growth = 1.0;
    unsigned seq_growth = 0;

    for (int attempt = 0; attempt < 8; ++attempt) {
        // outputs are created first, like the reference (an unreadable input leaves empty outputs behind)
        auto job = std::make_unique<OrderedJob>();
        auto& outs = job->outs; auto& ms = job->ms; auto& writers = job->writers; auto& eng = job->eng;
        for (int m = 0; m < mates; ++m) outs.emplace_back(new OutputFile(out[m]));
        trace("outputs created");
        for (int m = 0; m < mates; ++m) ms[m].reader.reset(new BlockReader(in[m], block));
        trace("readers started (pinned blocks allocated)");
        // one writer thread per output: the survivors of chunk c are written while chunk c+1 is on the device; an
        // input block goes back to its reader only after the writes that read from it (declared after `ms`, so the
        // writers finish before the blocks are freed, whatever the way out of this scope)
        for (int m = 0; m < mates; ++m) {
            writers.emplace_back(new AsyncWriter(*outs[m]));
            AsyncWriter* w = writers[m].get();
            BlockReader* r = ms[m].reader.get();
            ms[m].release = [w, r](Block* b) { w->then([r, b] { r->release(b); }); };
        }
        auto close_outputs = [&] { for (auto& w : writers) w->drain(); for (auto& o : outs) o->close(); check_outputs(outs, out); };
        std::vector<Run> runs;
        struct { bool valid = false; std::vector<char> bytes[2]; } held;      // last written pair of the previous chunk
        for (int m = 0; m < mates; ++m)
            if (!ms[m].refill()) throw std::runtime_error("Not enough memory to read a single object!");   // empty file
        for (int m = 0; m < mates; ++m)       // the very first record is validated by set_file()'s refresh (src/bufferedinput.hpp:76-86)
            if (ms[m].len && ms[m].ptr[0] != lead) {
                fqd_stats_t st; memset(&st, 0, sizeof st); st.err = FQD_ERR_BAD_START; st.err_char = (unsigned char)ms[m].ptr[0];
                throw_data_error(st, m_fasta);
            }

        size_t max_seq = 0; double avg_rec = 0; uint64_t est_records = 0;
        for (int m = 0; m < mates; ++m) {
            size_t ms_ = 0; double ar = 0;
            sample_geometry(ms[m].ptr, std::min<size_t>(ms[m].len, 8u << 20), lpr, ms_, ar);
            max_seq = std::max(max_seq, ms_);
            size_t fsz = file_size_or_zero(in[m]);
            double expand = gz_expansion(in[m], *ms[m].reader);
            uint64_t est = (uint64_t)((double)fsz * expand / std::max(ar, 8.0) * 1.02) + (1u << 16);
            est_records = m == 0 ? est : std::min(est_records, est);
            avg_rec = std::max(avg_rec, ar);
        }
        fqd_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = FQD_ABI_VERSION; cfg.device = m_device; cfg.mode = FQD_MODE_FAST;
        cfg.format = m_fasta ? FQD_FORMAT_FASTA : FQD_FORMAT_FASTQ; cfg.paired = mates == 2;
        cfg.max_seq_len = (uint32_t)((std::max<size_t>(max_seq, 20) + 19) / 20 * 20) << seq_growth;
        cfg.max_records = (uint64_t)((double)est_records * growth);
        cfg.max_chunk_bytes = 2 * block + 4096;
        cfg.max_chunk_records = 0;
        cfg.max_tag_len = 0;
        fqd_handle* hraw = nullptr;
        trace("first block read");
        int rc = fqd_create(&cfg, &hraw);
        if (rc) throw_engine_error(nullptr, rc);
        eng.reset(hraw);
        trace("engine created");

        bool restart = false;
        uint64_t total = 0, dups = 0;
        for (;;) {
            // make sure every mate has something new to parse; a mate whose tail is already large waits
            for (int m = 0; m < mates; ++m)
                if (ms[m].len < block / 2) ms[m].refill();
            fqd_chunk_result res;
            rc = fqd_push(eng.get(), ms[0].ptr, ms[0].len, mates == 2 ? ms[1].ptr : nullptr, mates == 2 ? ms[1].len : 0, &res);
            if (rc) throw_engine_error(eng.get(), rc);
            fqd_stats_t st;
            fqd_stats(eng.get(), &st);
            if (st.err == FQD_ERR_SEQ_TOO_LONG) { check_restart_possible(in, mates, "a later sequence is longer than the key rows sized from the first block"); ++seq_growth; restart = true; break; }
            if (st.err == FQD_ERR_CAPACITY) { check_restart_possible(in, mates, "the key store could not be grown in place"); growth *= 2.0; restart = true; break; }
            size_t n = (size_t)res.n_records;
            uint64_t chunk_dups = n - res.n_survivors;
            // A record that does not start with '@'/'>' aborts the run while the record BEFORE it is fetched
            // (src/bufferedinput.hpp:90-103 pre-parses it; src/fastqview.cpp:91-92 checks the first byte before
            // anything else), so that last record is not written.  Inside the chunk the engine reports it; at the
            // chunk's end the byte comes from the carried tail or from the next block.
            int tail_err_mate = -1, tail_char = 0;
            if (st.err == 0 && n > 0) {
                for (int m = 0; m < mates && tail_err_mate < 0; ++m) {
                    int nb = res.consumed[m] < ms[m].len ? (unsigned char)ms[m].ptr[res.consumed[m]] : ms[m].peek_next_byte();
                    if (nb >= 0 && nb != lead) { tail_err_mate = m; tail_char = nb; }
                }
            }
            if (tail_err_mate >= 0) {
                const uint8_t last_dup = res.dup[n - 1];
                n -= 1;
                chunk_dups -= last_dup;
            }
            // The lazy pre-parse reaches across chunks as well: the LAST pair of a chunk is written only once the
            // record that follows it has parsed - which may need the next block (a sequence / quality length mismatch
            // shows when the record is complete).  So that pair is held back (a copy of its bytes) and written in
            // front of the next chunk's survivors, or dropped if that chunk opens with a malformed record.
            const bool stops_here = st.err != 0 || tail_err_mate >= 0;
            if (held.valid && (n > 0 || stops_here)) {
                const bool next_is_malformed = (st.err == FQD_ERR_BAD_START || st.err == FQD_ERR_LEN_MISMATCH) &&
                                               st.err_record == res.first_record;
                if (!next_is_malformed)
                    for (int m = 0; m < mates; ++m) writers[m]->write_owned(std::move(held.bytes[m]));
                held.valid = false;
            }
            size_t n_now = n;
            if (n > 0 && !stops_here) {
                n_now = n - 1;
                if (!res.dup[n - 1]) {
                    held.valid = true;
                    for (int m = 0; m < mates; ++m) {
                        const char* b = ms[m].ptr + res.rec_start[m][n - 1];
                        held.bytes[m].assign(b, (const char*)(ms[m].ptr + res.rec_start[m][n]));
                    }
                }
            }
            for (int m = 0; m < mates; ++m) {
                const size_t bytes = survivor_runs(res.rec_start[m], res.dup, n_now, runs);
                writers[m]->write_runs(ms[m].ptr, std::move(runs), bytes);
                runs = std::vector<Run>();
            }
            total += n; dups += chunk_dups;
            if (tail_err_mate >= 0) {
                st.err = FQD_ERR_BAD_START; st.err_char = tail_char;
                close_outputs();
                throw_data_error(st, m_fasta);
            }
            if (st.err) {
                if (st.err == FQD_ERR_BAD_BASE && st.err_record == 0) {
                    // the reference writes the very first record (pair) BEFORE it keys it
                    // (src/hash_dup_remover.hpp:118-124,216-228): a bad base there still leaves it in the output
                    for (int m = 0; m < mates; ++m) {
                        const char* b = ms[m].ptr + res.rec_start[m][0];
                        writers[m]->write_owned(std::vector<char>(b, (const char*)(ms[m].ptr + res.rec_start[m][1])));
                    }
                }
                close_outputs();
                const int em = st.err_mate;
                size_t e = (size_t)(st.err_record - res.first_record);
                const char* rec = nullptr; size_t rl = 0;
                if (st.err == FQD_ERR_LEN_MISMATCH) { rec = ms[em].ptr + res.rec_start[em][e]; rl = ms[em].len - res.rec_start[em][e]; }
                throw_data_error(st, m_fasta, rec, rl);
            }
            for (int m = 0; m < mates; ++m) { ms[m].ptr += res.consumed[m]; ms[m].len -= res.consumed[m]; }
            if (res.n_records == 0) {
                // nothing complete in what we have: read more, or stop at the end of a file
                bool progressed = false;
                for (int m = 0; m < mates; ++m) {
                    bool has_rec = false;   // does this mate still hold a complete record?  (cheap upper bound: lpr newlines)
                    size_t cnt = 0; const char* c = ms[m].ptr; const char* e2 = c + ms[m].len;
                    while (cnt < (size_t)lpr && c < e2) { const char* nl = (const char*)memchr(c, '\n', e2 - c); if (!nl) break; ++cnt; c = nl + 1; }
                    has_rec = cnt == (size_t)lpr;
                    if (!has_rec) progressed |= ms[m].refill();
                }
                if (!progressed) break;      // a file is exhausted: stop at the shorter one (src/hash_dup_remover.hpp:228-230)
            }
        }
        trace("last chunk processed");
        if (restart) continue;
        if (held.valid)
            for (int m = 0; m < mates; ++m) writers[m]->write_owned(std::move(held.bytes[m]));
        if (total == 0) {
            fqd_stats_t st; memset(&st, 0, sizeof st); st.err = FQD_ERR_EMPTY;
            throw_data_error(st, m_fasta);
        }
        close_outputs();
        trace("outputs closed");
        if (m_verbose) {
            if (mates == 1) std::cout << total << " reads processed, out of which " << dups << " duplicates were removed.\n";
            else std::cout << total << " read pairs processed, out of which " << dups << " duplicates were removed.\n";
        }
        leave_to_the_os(job);
        return;
    }
    throw std::runtime_error("input exceeds the device capacity of the fast-mode key store");
}


// -------------------------------------------------------------------------------------------------------------
// --fast over several GPUs of the box (FQD_DEVICES): the chunks of the input are dealt round-robin to one engine per
// GPU, every engine owns one hash range of the key space (csrc/shard2.cuh: rows travel over NVLink into the owner's
// key store, flags come back the same way), and the survivors are written in input order exactly as in run_ordered.
// What this buys is CAPACITY - the key set of an input that one GPU cannot hold (BASELINE configs[4]: 1 G pairs = 128 GB
// of key rows + table) is spread over N x 180 GB; the files themselves are read and written by the same host threads
// at the same speed.  One host thread enqueues everything in program order, so the engines need no barrier.
void HashDupRemover::run_ordered_multi(const std::string* in, const std::string* out, int mates, const std::vector<int>& devices) {
    const int lpr = m_fasta ? 2 : 4;
    const char lead = m_fasta ? '>' : '@';
    const uint32_t N = (uint32_t)devices.size();
    size_t block = (size_t)m_memlimit / (size_t)(mates * (N + 3) * 2);
    block = std::min<size_t>(std::max<size_t>(block, 4u << 20), 64u << 20) & ~(size_t)4095;
    block = test_block_override(block);
    double growth = 1.0;
    unsigned seq_growth = 0;

    for (int attempt = 0; attempt < 8; ++attempt) {
        std::vector<std::unique_ptr<OutputFile>> outs;
        MateStream ms[2];
        std::vector<std::unique_ptr<AsyncWriter>> writers;
        std::vector<EnginePtr> eng;
        for (int m = 0; m < mates; ++m) outs.emplace_back(new OutputFile(out[m]));
        for (int m = 0; m < mates; ++m) ms[m].reader.reset(new BlockReader(in[m], block, (int)N + 3));
        // blocks go back to their reader only behind the writes of the round that still reads from them
        std::vector<Block*> deferred[2];
        for (int m = 0; m < mates; ++m) {
            writers.emplace_back(new AsyncWriter(*outs[m]));
            std::vector<Block*>* d = &deferred[m];
            ms[m].release = [d](Block* b) { d->push_back(b); };
        }
        auto flush_deferred = [&] {
            for (int m = 0; m < mates; ++m) {
                AsyncWriter* w = writers[m].get(); BlockReader* r = ms[m].reader.get();
                for (Block* b : deferred[m]) w->then([r, b] { r->release(b); });
                deferred[m].clear();
            }
        };
        auto close_outputs = [&] { for (auto& w : writers) w->drain(); for (auto& o : outs) o->close(); check_outputs(outs, out); };
        for (int m = 0; m < mates; ++m)
            if (!ms[m].refill()) throw std::runtime_error("Not enough memory to read a single object!");
        for (int m = 0; m < mates; ++m)
            if (ms[m].len && ms[m].ptr[0] != lead) {
                fqd_stats_t st; memset(&st, 0, sizeof st); st.err = FQD_ERR_BAD_START; st.err_char = (unsigned char)ms[m].ptr[0];
                throw_data_error(st, m_fasta);
            }
        size_t max_seq = 0; double avg_rec = 1e9; uint64_t est_records = 0;
        for (int m = 0; m < mates; ++m) {
            size_t ms_ = 0; double ar = 0;
            sample_geometry(ms[m].ptr, std::min<size_t>(ms[m].len, 8u << 20), lpr, ms_, ar);
            max_seq = std::max(max_seq, ms_);
            double expand = gz_expansion(in[m], *ms[m].reader);
            uint64_t est = (uint64_t)((double)file_size_or_zero(in[m]) * expand / std::max(ar, 8.0) * 1.02) + (1u << 16);
            est_records = m == 0 ? est : std::min(est_records, est);
            avg_rec = std::min(avg_rec, std::max(ar, 8.0));
        }
        // a chunk is at most 2 blocks of the mate with the shortest records; an owner gets 1/N of it from every source
        const uint64_t chunk_records = (uint64_t)((double)(2 * block + 4096) / avg_rec * 1.1 * growth) + 1024;
        uint64_t region = chunk_records / N + (uint64_t)(8.0 * std::sqrt((double)chunk_records / N)) + chunk_records / (N * 32) + 1024;
        region = (region + 15) / 16 * 16;
        const uint64_t n_chunks_est = (uint64_t)((double)est_records * growth / std::max<double>(1.0, (double)block / avg_rec)) / N + 4;   // rounds
        fqd_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = FQD_ABI_VERSION; cfg.mode = FQD_MODE_FAST;
        cfg.format = m_fasta ? FQD_FORMAT_FASTA : FQD_FORMAT_FASTQ; cfg.paired = mates == 2;
        cfg.max_seq_len = (uint32_t)((std::max<size_t>(max_seq, 20) + 19) / 20 * 20) << seq_growth;
        cfg.max_records = n_chunks_est * N * region + 1024;       // rounds x (one region per source)
        cfg.max_chunk_bytes = 2 * block + 4096;
        cfg.max_chunk_records = chunk_records;
        for (uint32_t r = 0; r < N; ++r) {
            cfg.device = devices[r];
            fqd_handle* hraw = nullptr;
            int rc = fqd_create(&cfg, &hraw);
            if (rc) throw_engine_error(nullptr, rc);
            eng.emplace_back(hraw);
            rc = fqd_shard2_init(hraw, N, r, (uint32_t)region);
            if (rc) throw_engine_error(hraw, rc);
        }
        for (uint32_t a = 0; a < N; ++a)
            for (uint32_t b = 0; b < N; ++b)
                if (a != b) { int rc = fqd_shard2_link(eng[a].get(), b, eng[b].get()); if (rc) throw_engine_error(eng[a].get(), rc); }
        trace("engines created and linked");

        struct InFlight { const char* ptr[2]; size_t len[2]; uint64_t n; uint64_t consumed[2]; int next_byte[2]; bool pushed; };
        std::vector<InFlight> fl(N);
        std::vector<Run> runs;
        struct { bool valid = false; std::vector<char> bytes[2]; } held;
        bool restart = false, input_done = false;
        uint64_t total = 0, dups = 0, first_record = 0;
        for (uint64_t round = 0; !input_done && !restart; ++round) {
            // ---- deal one chunk to every engine, in input order (an engine without input gets an empty chunk)
            uint32_t n_pushed = 0;
            for (uint32_t r = 0; r < N; ++r) {
                InFlight& f = fl[r];
                f.pushed = false; f.n = 0;
                if (!input_done) {
                    for (int m = 0; m < mates; ++m)
                        if (ms[m].len < block / 2) ms[m].refill();
                }
                for (int m = 0; m < 2; ++m) { f.ptr[m] = m < mates ? ms[m].ptr : nullptr; f.len[m] = (m < mates && !input_done) ? ms[m].len : 0; f.consumed[m] = 0; f.next_byte[m] = -1; }
                int rc = fqd_shard2_push_host(eng[r].get(), round, f.ptr[0], f.len[0], mates == 2 ? f.ptr[1] : nullptr, mates == 2 ? f.len[1] : 0, &f.n, f.consumed);
                if (rc == FQD_ERR_CAPACITY) { check_restart_possible(in, mates, "the key stores could not hold another round of chunks"); growth *= 2.0; restart = true; break; }
                if (rc) throw_engine_error(eng[r].get(), rc);
                f.pushed = true; ++n_pushed;
                if (input_done) continue;
                for (int m = 0; m < mates; ++m)
                    f.next_byte[m] = f.consumed[m] < f.len[m] ? (unsigned char)f.ptr[m][f.consumed[m]] : ms[m].peek_next_byte();
                for (int m = 0; m < mates; ++m) { ms[m].ptr += f.consumed[m]; ms[m].len -= f.consumed[m]; }
                if (f.n == 0) {
                    // nothing complete in what we have: read more, or stop at the end of a file
                    bool progressed = false;
                    for (int m = 0; m < mates; ++m) {
                        size_t cnt = 0; const char* c = ms[m].ptr; const char* e2 = c + ms[m].len;
                        while (cnt < (size_t)lpr && c < e2) { const char* nl = (const char*)memchr(c, '\n', e2 - c); if (!nl) break; ++cnt; c = nl + 1; }
                        if (cnt != (size_t)lpr) progressed |= ms[m].refill();
                    }
                    if (!progressed) input_done = true;      // a file is exhausted: stop at the shorter one
                }
            }
            if (restart) break;
            for (uint32_t r = 0; r < N; ++r) { int rc = fqd_shard2_insert(eng[r].get(), round); if (rc) throw_engine_error(eng[r].get(), rc); }
            for (uint32_t r = 0; r < N; ++r) { int rc = fqd_shard2_apply(eng[r].get(), round); if (rc) throw_engine_error(eng[r].get(), rc); }
            // ---- the round's results, in input order: the same rules as run_ordered
            for (uint32_t r = 0; r < N && !restart; ++r) {
                InFlight& f = fl[r];
                fqd_chunk_result res;
                int rc = fqd_shard2_result(eng[r].get(), first_record, f.len[0], f.len[1], &res);
                if (rc) throw_engine_error(eng[r].get(), rc);
                fqd_stats_t st;
                fqd_stats(eng[r].get(), &st);
                if (st.err == FQD_ERR_SEQ_TOO_LONG) { check_restart_possible(in, mates, "a later sequence is longer than the key rows sized from the first block"); ++seq_growth; restart = true; break; }
                if (st.err == FQD_ERR_CAPACITY) { check_restart_possible(in, mates, "a key-store region overflowed"); growth *= 2.0; restart = true; break; }
                size_t n = (size_t)res.n_records;
                uint64_t chunk_dups = n - res.n_survivors;
                int tail_err_mate = -1, tail_char = 0;
                if (st.err == 0 && n > 0)
                    for (int m = 0; m < mates && tail_err_mate < 0; ++m)
                        if (f.next_byte[m] >= 0 && f.next_byte[m] != lead) { tail_err_mate = m; tail_char = f.next_byte[m]; }
                if (tail_err_mate >= 0) { const uint8_t last_dup = res.dup[n - 1]; n -= 1; chunk_dups -= last_dup; }
                const bool stops_here = st.err != 0 || tail_err_mate >= 0;
                if (held.valid && (n > 0 || stops_here)) {
                    const bool next_is_malformed = (st.err == FQD_ERR_BAD_START || st.err == FQD_ERR_LEN_MISMATCH) && st.err_record == res.first_record;
                    if (!next_is_malformed)
                        for (int m = 0; m < mates; ++m) writers[m]->write_owned(std::move(held.bytes[m]));
                    held.valid = false;
                }
                size_t n_now = n;
                if (n > 0 && !stops_here) {
                    n_now = n - 1;
                    if (!res.dup[n - 1]) {
                        held.valid = true;
                        for (int m = 0; m < mates; ++m) held.bytes[m].assign(f.ptr[m] + res.rec_start[m][n - 1], f.ptr[m] + res.rec_start[m][n]);
                    }
                }
                for (int m = 0; m < mates; ++m) {
                    const size_t bytes = survivor_runs(res.rec_start[m], res.dup, n_now, runs);
                    writers[m]->write_runs(f.ptr[m], std::move(runs), bytes);
                    runs = std::vector<Run>();
                }
                total += n; dups += chunk_dups; first_record += res.n_records;
                if (tail_err_mate >= 0) {
                    st.err = FQD_ERR_BAD_START; st.err_char = tail_char;
                    close_outputs();
                    throw_data_error(st, m_fasta);
                }
                if (st.err) {
                    if (st.err == FQD_ERR_BAD_BASE && st.err_record == 0)
                        for (int m = 0; m < mates; ++m) writers[m]->write_owned(std::vector<char>(f.ptr[m] + res.rec_start[m][0], f.ptr[m] + res.rec_start[m][1]));
                    close_outputs();
                    const int em = st.err_mate;
                    const size_t e = (size_t)(st.err_record - res.first_record);
                    const char* rec = nullptr; size_t rl = 0;
                    if (st.err == FQD_ERR_LEN_MISMATCH) { rec = f.ptr[em] + res.rec_start[em][e]; rl = f.len[em] - res.rec_start[em][e]; }
                    throw_data_error(st, m_fasta, rec, rl);
                }
            }
            flush_deferred();
        }
        if (restart) { for (auto& w : writers) w->drain(); continue; }
        if (held.valid)
            for (int m = 0; m < mates; ++m) writers[m]->write_owned(std::move(held.bytes[m]));
        if (total == 0) {
            fqd_stats_t st; memset(&st, 0, sizeof st); st.err = FQD_ERR_EMPTY;
            throw_data_error(st, m_fasta);
        }
        flush_deferred();
        close_outputs();
        if (m_verbose) {
            if (mates == 1) std::cout << total << " reads processed, out of which " << dups << " duplicates were removed.\n";
            else std::cout << total << " read pairs processed, out of which " << dups << " duplicates were removed.\n";
        }
        return;
    }
    throw std::runtime_error("input exceeds the device capacity of the fast-mode key stores");
}

// -------------------------------------------------------------------------------------------------------------
// SeqDupRemover: sort + comparator scan, output in sorted order (src/seq_dup_remover.hpp:40-109,111-218)
void SeqDupRemover::filterSE(const std::string& infile, const std::string& outfile) {
    const std::string in[1] = {infile}, out[1] = {outfile};
    int mode = m_ctype == CT_LOOSE ? FQD_MODE_SEQ_LOOSE : m_ctype == CT_HAMMING ? FQD_MODE_SEQ_HAMMING : FQD_MODE_SEQ_TIGHT;
    run_whole_input(mode, m_fasta, false, m_dist, 1, in, out, m_write_clusters, m_verbose, m_memlimit, m_device);
}

void SeqDupRemover::filterPE(const std::string& infile1, const std::string& infile2,
                             const std::string& outfile1, const std::string& outfile2) {
    const std::string in[2] = {infile1, infile2}, out[2] = {outfile1, outfile2};
    int mode = m_ctype == CT_LOOSE ? FQD_MODE_SEQ_LOOSE : m_ctype == CT_HAMMING ? FQD_MODE_SEQ_HAMMING : FQD_MODE_SEQ_TIGHT;
    run_whole_input(mode, m_fasta, false, m_dist, 2, in, out, m_write_clusters, m_verbose, m_memlimit, m_device);
}

void run_whole_input(int mode, bool fasta, bool unordered, unsigned dist, int mates, const std::string* in,
                     const std::string* out, bool write_clusters, bool verbose, ssize_t memlimit, int device) {
    const int lpr = fasta ? 2 : 4;
    size_t block = (size_t)memlimit / (size_t)(mates * 3 * 2);
    block = std::min<size_t>(std::max<size_t>(block, 4u << 20), 64u << 20) & ~(size_t)4095;
    block = test_block_override(block);
    double growth = 1.0;
    unsigned seq_growth = 0, tag_growth = 0;
    bool byte_keys = false;         // sequence-based modes order ANY byte (src/fastqview.cpp:56-67): raw-byte key rows
    // Where the raw bytes live while the job sorts: on the device (it gathers the output itself), or - inputs larger than
    // device memory - nowhere on the device: only key rows + record tables stay, the host fetches the written records
    // from its own copy of the input (replay.hpp).  The reference's counterpart is the external sort's disk chunks.
    const int policy = whole_input_policy();
    bool discard = policy == 2;
    // A pipe can be read once, but its complete spool can be read again: `in` is what the user named, `eff` what this
    // attempt reads (the original, or /proc/self/fd/<spool> after the first pass), `kept` holds the spools alive.
    const std::vector<std::string> original(in, in + mates);
    std::vector<std::string> eff(original);
    std::vector<std::shared_ptr<InputReplay>> kept(mates);
    in = eff.data();
    auto can_restart = [&](const char* why) {
        for (int m = 0; m < mates; ++m)
            if (!rereadable(eff[m]))
                throw std::runtime_error(std::string("input ") + original[m] + " is not a regular file and cannot be read a second time (" + why +
                                         "); write it to a file first");
    };

    for (int attempt = 0; attempt < 10; ++attempt) {
        auto job = std::make_unique<WholeJob>();
        auto& readers = job->readers; auto& eng = job->eng; auto& replay = job->replay;
        for (int m = 0; m < mates; ++m) readers.emplace_back(new BlockReader(in[m], block));
        // geometry of the first block of every file sizes the key rows and the record tables
        std::vector<Block*> first(mates, nullptr);
        size_t max_seq = 0; uint64_t est_records = 0;
        double raw_bytes = 0; bool sizes_known = true;
        for (int m = 0; m < mates; ++m) {
            first[m] = readers[m]->next();
            if (!first[m] || first[m]->len == 0) throw std::runtime_error("Not enough memory to read a single object!");
            size_t ms_ = 0; double ar = 0;
            sample_geometry(first[m]->data(), std::min<size_t>(first[m]->len, 8u << 20), lpr, ms_, ar);
            max_seq = std::max(max_seq, ms_);
            double expand = gz_expansion(in[m], *readers[m]);
            uint64_t est = (uint64_t)((double)file_size_or_zero(in[m]) * expand / std::max(ar, 8.0) * 1.02) + (1u << 16);
            est_records = std::max(est_records, est);
            raw_bytes += (double)file_size_or_zero(in[m]) * expand;
            if (!rereadable(in[m])) sizes_known = false;
        }
        fqd_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = FQD_ABI_VERSION; cfg.device = device; cfg.mode = mode;
        cfg.format = fasta ? FQD_FORMAT_FASTA : FQD_FORMAT_FASTQ; cfg.paired = mates == 2; cfg.unordered = unordered;
        cfg.hamming_dist = dist;
        cfg.max_seq_len = (uint32_t)((std::max<size_t>(max_seq, 20) + 19) / 20 * 20) << seq_growth;
        cfg.max_records = (uint64_t)((double)est_records * growth);
        cfg.max_chunk_bytes = 1ull << 30;              // device segments of 1 GiB
        cfg.max_tag_len = 32u << tag_growth;
        cfg.byte_keys = byte_keys ? 1u : 0u;
        // a pipe has no size to decide by and cannot be read twice: spool it (the reference writes its whole input to the
        // temporary directory as sorted chunks, src/external_sort.hpp:104-113) - any size works, and so do restarts
        if (policy == 0 && !sizes_known) discard = true;
        if (policy == 0 && !discard && sizes_known) {
            // resident needs the raw bytes + what discard needs anyway + staging for the output gather
            size_t free_b = 0;
            if (fqd_device_memory(device, &free_b, nullptr) == FQD_OK) {
                const double need = raw_bytes * 1.03 + (double)cfg.max_records * (double)resident_bytes_per_record(cfg, mates) + 3.0 * (double)(1ull << 30);
                if (need > 0.92 * (double)free_b) discard = true;
            }
            // Plain files that the page cache can hold are better off on the discarded-input path even when they would fit:
            // nothing is spooled (the files are mapped), the device recycles two segments instead of allocating one per
            // GiB of input (ingest 26 vs 7 GB/s at 25 M pairs), and the host's gather writes as fast as fqd_emit + D2H does.
            bool all_plain = true;
            for (int m = 0; m < mates; ++m) all_plain = all_plain && !has_gz_ext(in[m]);
            const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGESIZE);
            if (all_plain && pages > 0 && psz > 0 && raw_bytes < 0.4 * (double)pages * (double)psz) discard = true;
        }
        fqd_handle* hraw = nullptr;
        int rc = fqd_create(&cfg, &hraw);
        if (rc) throw_engine_error(nullptr, rc);
        eng.reset(hraw);
        if (discard) {
            rc = fqd_discard_input(eng.get(), 1);
            if (rc) throw_engine_error(eng.get(), rc);
            for (int m = 0; m < mates; ++m)
                replay.emplace_back(new InputReplay(in[m], rereadable(in[m]) && !has_gz_ext(in[m]), spool_dir_for(out[m])));
        }
        trace(discard ? "engine created (raw input not kept on the device)" : "engine created");

        // ship every block to the device (the sort/join needs the whole input; nothing is kept on the host)
        bool out_of_device_memory = false;
        for (int m = 0; m < mates && !out_of_device_memory; ++m) {
            // a spooled input is written behind the caller's back, like an output: the block goes to the device, then to the
            // spool's ordered writer (several workers pwrite it), and returns to its reader when that is done
            std::unique_ptr<OutputFile> spool_file;
            std::unique_ptr<AsyncWriter> spool;
            if (discard && replay[m]->spooling()) {
                spool_file.reset(new OutputFile(replay[m]->proc_path()));
                spool.reset(new AsyncWriter(*spool_file));
            }
            Block* b = first[m];
            while (b) {
                rc = fqd_append(eng.get(), m, b->data(), b->len);
                if (rc == FQD_ERR_CUDA && !discard && policy == 0 && std::strstr(fqd_last_error(eng.get()), "out of memory")) { out_of_device_memory = true; break; }
                if (rc) throw_engine_error(eng.get(), rc);
                if (spool) {
                    BlockReader* rd = readers[m].get();
                    spool->write_runs(b->data(), std::vector<Run>{Run{0, b->len, 0}}, b->len);
                    spool->then([rd, b] { rd->release(b); });
                } else {
                    readers[m]->release(b);
                }
                b = readers[m]->next();
            }
            if (spool) {
                spool->drain();
                spool.reset();
                spool_file->close();
                if (int e = spool_file->error())
                    throw std::runtime_error(std::string("writing the input spool failed: ") + std::strerror(e) +
                                             " (set FQD_SPOOL_DIR to a directory with room for the uncompressed input)");
            }
        }
        if (out_of_device_memory) {       // the size estimate was too kind (a .gz that expands more than its head suggested)
            can_restart("the input does not fit into device memory");
            discard = true;
            continue;
        }
        if (discard)                       // every byte of a pipe is in its spool now: later attempts and error messages read that
            for (int m = 0; m < mates; ++m)
                if (replay[m]->spooling() && !rereadable(eff[m])) { kept[m] = replay[m]; eff[m] = kept[m]->proc_path(); }
        trace("input on the device");
        rc = fqd_finish(eng.get());
        if (rc == FQD_ERR_CUDA && !discard && policy == 0 && std::strstr(fqd_last_error(eng.get()), "out of memory") && rereadable(in[0]) && (mates == 1 || rereadable(in[1]))) {
            discard = true;
            continue;
        }
        if (rc) throw_engine_error(eng.get(), rc);
        trace("sorted / joined / scanned");
        if (std::getenv("FQD_TRACE")) {      // the pool keeps what it ever held: this is the job's high-water mark
            size_t free_b = 0, total_b = 0;
            if (fqd_device_memory(device, &free_b, &total_b) == FQD_OK)
                std::fprintf(stderr, "[host-trace] device memory high-water mark %.2f GiB of %.2f GiB\n", (double)(total_b - free_b) / (1ull << 30), (double)total_b / (1ull << 30));
        }
        fqd_stats_t st;
        fqd_stats(eng.get(), &st);
        if (st.err == FQD_ERR_SEQ_TOO_LONG) { can_restart("a later sequence is longer than the key rows sized from the first block"); ++seq_growth; continue; }
        if (st.err == FQD_ERR_CAPACITY) { can_restart("the record tables could not be grown in place"); growth *= 2.0; continue; }
        if (st.err == FQD_ERR_TAG_TOO_LONG) { can_restart("an ID tag is longer than the tag rows"); ++tag_growth; continue; }
        if (st.err == FQD_ERR_UNSUPPORTED_BYTE && !byte_keys) { can_restart("a sequence holds a byte outside {A,C,G,T,N}: raw-byte key rows are needed"); byte_keys = true; continue; }
        // parse errors surface while the inputs are being sorted, before any output file exists
        // (src/seq_dup_remover.hpp:44-50, src/hash_dup_remover.hpp:160-174)
        if (st.err == FQD_ERR_LEN_MISMATCH) {
            const std::string rec = fetch_record(in[st.err_mate == 1 && mates == 2 ? 1 : 0], st.err_record, lpr);
            throw_data_error(st, fasta, rec.data(), rec.size());
        }
        if (st.err && st.err != FQD_ERR_BAD_BASE) throw_data_error(st, fasta);

        std::vector<std::unique_ptr<OutputFile>> outs;
        for (int m = 0; m < mates; ++m) outs.emplace_back(new OutputFile(out[m]));
        if (discard) {
            uint64_t n_written = 0, n_sorted = 0;
            rc = fqd_emission_count(eng.get(), &n_written, &n_sorted);
            if (rc) throw_engine_error(eng.get(), rc);
            for (int m = 0; m < mates; ++m) {
                replay[m]->seal();
                replay[m]->prefault(io_threads());
                gather_from_replay(eng.get(), m, n_written, *replay[m], *outs[m]);
            }
            for (auto& o : outs) o->close();
            check_outputs(outs, out);
            if (write_clusters)
                for (int m = 0; m < mates; ++m) clusters_from_replay(eng.get(), m, n_sorted, *replay[m], out[m]);
            trace("outputs closed (gathered on the host)");
            if (st.err == FQD_ERR_BAD_BASE) throw_data_error(st, fasta);
            print_whole_summary(st, unordered, mates, verbose);
            leave_to_the_os(job);
            return;
        }
        // two pinned staging buffers: the device gathers + copies the next piece while a writer thread writes the last
        const size_t cap = 64u << 20;
        void* stage_buf[2] = {nullptr, nullptr};
        for (auto& p : stage_buf)
            if (fqd_host_alloc(&p, cap) != FQD_OK) throw std::runtime_error("pinned host allocation failed (no usable CUDA device? this build has no CPU path)");
        struct Free { void** p; ~Free() { fqd_host_free(p[0]); fqd_host_free(p[1]); } } guard{stage_buf};
        void* stage = stage_buf[0];
        for (int m = 0; m < mates; ++m) {
            std::mutex mu; std::condition_variable cv; bool busy[2] = {false, false};
            AsyncWriter writer(*outs[m]);
            int slot = 0;
            for (;;) {
                { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return !busy[slot]; }); }
                size_t nb = 0; int done = 0;
                rc = fqd_emit(eng.get(), m, stage_buf[slot], cap, &nb, &done);
                if (rc) { writer.drain(); throw_engine_error(eng.get(), rc); }
                if (nb) {
                    { std::lock_guard<std::mutex> g(mu); busy[slot] = true; }
                    writer.write_runs((const char*)stage_buf[slot], std::vector<Run>{Run{0, nb, 0}}, nb);
                    writer.then([&mu, &cv, &busy, slot] { { std::lock_guard<std::mutex> g(mu); busy[slot] = false; } cv.notify_all(); });
                    slot ^= 1;
                }
                if (done) break;
            }
            writer.drain();
        }
        for (auto& o : outs) o->close();
        check_outputs(outs, out);
        if (write_clusters) {
            // ClusterFile (src/file_utils.cpp:98-112): plain text <outfile>.clusters next to every output file
            for (int m = 0; m < mates; ++m) {
                std::ofstream cf(out[m] + ".clusters", std::ios::binary);
                for (;;) {
                    size_t nb = 0; int done = 0;
                    rc = fqd_emit_clusters(eng.get(), m, stage, cap, &nb, &done);
                    if (rc) throw_engine_error(eng.get(), rc);
                    cf.write((const char*)stage, (std::streamsize)nb);
                    if (done) break;
                }
            }
        }
        trace("outputs closed");
        if (st.err == FQD_ERR_BAD_BASE) throw_data_error(st, fasta);     // --unordered: raised while pairs are keyed
        print_whole_summary(st, unordered, mates, verbose);
        leave_to_the_os(job);
        return;
    }
    throw std::runtime_error("input exceeds the device capacity");
}

}  // namespace fqdhost
