// dup_remover.hpp - host-side mirror of the reference's two drivers, same constructor arguments and methods:
//   HashDupRemover<T>{memlimit, tempdir, verbose}.filterSE / filterPE(.., unordered)   src/hash_dup_remover.hpp:73-94
//   SeqDupRemover<T>{memlimit, comparator, tempdir, write_clusters, verbose}.filterSE / filterPE   src/seq_dup_remover.hpp:12-38
// The record type T (FastqView / FastaView) becomes a runtime `fasta` flag, the comparator object becomes
// (ComparatorType, distance), and there is no temp directory: nothing is spilled to disk - sorting, joining and
// the duplicate set live in HBM behind the C ABI (include/fqd.h).  Errors are std::runtime_error with the
// reference's messages, caught in main() exactly like the reference (src/main.cpp:250-259).
#pragma once
#include <sys/types.h>
#include <string>
#include <vector>

#include "options.hpp"

namespace fqdhost {

class HashDupRemover {
public:
    HashDupRemover(ssize_t memlimit, bool fasta, bool verbose, int device)
        : m_memlimit(memlimit), m_fasta(fasta), m_verbose(verbose), m_device(device) {}
    void filterSE(const std::string& infile, const std::string& outfile);
    void filterPE(const std::string& infile1, const std::string& infile2,
                  const std::string& outfile1, const std::string& outfile2, bool unordered);
private:
    void run_ordered(const std::string* in, const std::string* out, int mates);
    // FQD_DEVICES=0,1,...: the duplicate set is sharded by hash range over several GPUs (one process, one thread)
    void run_ordered_multi(const std::string* in, const std::string* out, int mates, const std::vector<int>& devices);
    ssize_t m_memlimit;
    bool m_fasta, m_verbose;
    int m_device;
};

class SeqDupRemover {
public:
    SeqDupRemover(ssize_t memlimit, ComparatorType ctype, unsigned hammdist, bool fasta, bool write_clusters, bool verbose, int device)
        : m_memlimit(memlimit), m_ctype(ctype), m_dist(hammdist), m_fasta(fasta), m_write_clusters(write_clusters),
          m_verbose(verbose), m_device(device) {}
    void filterSE(const std::string& infile, const std::string& outfile);
    void filterPE(const std::string& infile1, const std::string& infile2,
                  const std::string& outfile1, const std::string& outfile2);
private:
    ssize_t m_memlimit;
    ComparatorType m_ctype;
    unsigned m_dist;
    bool m_fasta, m_write_clusters, m_verbose;
    int m_device;
};

// shared by both drivers: whole-input modes (sequence-based, --fast --unordered)
void run_whole_input(int mode, bool fasta, bool unordered, unsigned dist, int mates, const std::string* in,
                     const std::string* out, bool write_clusters, bool verbose, ssize_t memlimit, int device);

}  // namespace fqdhost
