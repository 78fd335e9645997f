// seq_stub.cpp - the sequence-mode binding of INTEGRATION.md section 2 as a standalone program: a plain C++ caller with
// nothing but include/fqd.h - no host layer of this repository - runs `--compare-seq <mode>` over one plain FASTQ / FASTA
// file on the discarded-input path and writes the survivors (and, optionally, the cluster file) from its own mapping of
// the input.  tests/test_integration_stub.py builds it against the test double (CPU) and against libfqd_cuda.so (GPU) and
// compares the bytes with the oracle.
//   seq_stub <in> <out> tight|loose|tail-hamming <distance> fastq|fasta [clusters]
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fqd.h"

static int die(fqd_handle* h, const char* what) { std::fprintf(stderr, "%s: %s\n", what, fqd_last_error(h)); return 2; }

int main(int argc, char** argv) {
    if (argc < 6) return 64;
    const std::string mode = argv[3];
    fqd_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = FQD_ABI_VERSION;
    cfg.mode = mode == "loose" ? FQD_MODE_SEQ_LOOSE : mode == "tail-hamming" ? FQD_MODE_SEQ_HAMMING : FQD_MODE_SEQ_TIGHT;
    cfg.hamming_dist = (uint32_t)std::atoi(argv[4]);
    cfg.format = std::string(argv[5]) == "fasta" ? FQD_FORMAT_FASTA : FQD_FORMAT_FASTQ;
    cfg.max_seq_len = 300; cfg.max_records = 1024 /* grows in place */; cfg.max_chunk_bytes = 1u << 20;
    fqd_handle* h = nullptr;
    if (fqd_create(&cfg, &h)) return die(nullptr, "fqd_create");
    if (fqd_discard_input(h, 1)) return die(h, "fqd_discard_input");
    const int fd = open(argv[1], O_RDONLY);
    struct stat sb;
    if (fd < 0 || fstat(fd, &sb) != 0) return 66;
    const size_t size = (size_t)sb.st_size;
    const char* file = size ? (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    if (file == MAP_FAILED) return 66;
    const size_t piece = 300000;          // small pieces: several segments are parsed and freed on the way
    for (size_t o = 0; o < size; o += piece)
        if (fqd_append(h, 0, file + o, std::min(piece, size - o))) return die(h, "fqd_append");
    if (fqd_finish(h)) return die(h, "fqd_finish");
    fqd_stats_t st;
    fqd_stats(h, &st);
    if (st.err) { std::fprintf(stderr, "data error %d at record %llu\n", st.err, (unsigned long long)st.err_record); return 1; }
    uint64_t n_written = 0, n_sorted = 0;
    if (fqd_emission_count(h, &n_written, &n_sorted)) return die(h, "fqd_emission_count");
    std::vector<uint64_t> off(1000);
    std::vector<uint32_t> len(off.size());
    std::vector<uint8_t> head(off.size());
    FILE* out = std::fopen(argv[2], "wb");
    for (uint64_t k = 0; k < n_written; k += off.size()) {
        const uint64_t c = std::min<uint64_t>(off.size(), n_written - k);
        if (fqd_emission_read(h, 0, k, c, off.data(), len.data())) return die(h, "fqd_emission_read");
        for (uint64_t i = 0; i < c; ++i) std::fwrite(file + off[i], 1, len[i], out);
    }
    std::fclose(out);
    if (argc > 6) {
        FILE* cl = std::fopen((std::string(argv[2]) + ".clusters").c_str(), "wb");
        for (uint64_t k = 0; k < n_sorted; k += off.size()) {
            const uint64_t c = std::min<uint64_t>(off.size(), n_sorted - k);
            if (fqd_cluster_read(h, 0, k, c, off.data(), len.data(), head.data())) return die(h, "fqd_cluster_read");
            for (uint64_t i = 0; i < c; ++i) {
                const char* p = file + off[i];
                const char* nl = (const char*)std::memchr(p, '\n', len[i]);
                if (!head[i]) std::fwrite("--", 1, 2, cl);
                std::fwrite(p, 1, nl ? (size_t)(nl - p) + 1 : (size_t)len[i], cl);
            }
        }
        std::fclose(cl);
    }
    std::printf("%llu reads processed, out of which %llu duplicates were removed.\n", (unsigned long long)st.total, (unsigned long long)st.dups);
    fqd_destroy(h);
    return 0;
}
