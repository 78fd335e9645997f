// options.hpp - command line of the drop-in binary: same flags, validation rules, messages and exit codes as
// the reference's parse_args() (src/main.cpp:28-179).  Boost.Program_options is replaced by a small parser
// with the same surface: long/short names, "--opt value" and "--opt=value", unambiguous long-name prefixes.
#pragma once
#include <sys/types.h>
#include <string>

namespace fqdhost {

const char* const VERSION = "fastq-dupaway V1.5.0-b200";      // src/constants.hpp:10 (tests check the "fastq-dupaway V" prefix)

enum ComparatorType { CT_NONE, CT_TIGHT, CT_LOOSE, CT_HAMMING };   // src/comparator.hpp:10-16

struct Options {                      // src/main.cpp:28-38
    bool fasta = false, paired = false, hash = false;
    ssize_t memLimit = 2048L * 1024L * 1024L;
    std::string input_1, input_2, output_1, output_2;
    ComparatorType ctype = CT_TIGHT;
    unsigned hammdist = 2;
    bool unordered = false, verbose = false, write_clusters = false;
    // extensions (environment, not flags, so the flag surface stays the reference's)
    int device = 0;
};

// false => the caller exits with status 1 (help requested or invalid arguments), like the reference.
bool parse_args(int argc, char** argv, Options& opts);

}  // namespace fqdhost
