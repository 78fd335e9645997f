// main.cpp - the drop-in `fastq-dupaway` binary: argument handling and dispatch as in the reference
// (src/main.cpp:181-262), with the two drivers backed by the B200 engine (libfqd_cuda.so).
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "dup_remover.hpp"
#include "io.hpp"
#include "options.hpp"

using namespace fqdhost;

// FQD_BACKTRACE=1: raw return addresses on a crash (resolve with addr2line against the binary / libfqd_cuda.so)
static void crash_handler(int sig) {
    void* frames[48];
    const int n = backtrace(frames, 48);
    const char msg[] = "fastq-dupaway: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char** argv) {
    if (std::getenv("FQD_BACKTRACE")) { signal(SIGSEGV, crash_handler); signal(SIGBUS, crash_handler); signal(SIGABRT, crash_handler); }
    Options opts;
    if (!parse_args(argc, argv, opts)) return 1;
    trace("start");
    try {
        if (!opts.hash) {
            SeqDupRemover remover(opts.memLimit, opts.ctype, opts.hammdist, opts.fasta, opts.write_clusters, opts.verbose, opts.device);
            if (opts.paired) remover.filterPE(opts.input_1, opts.input_2, opts.output_1, opts.output_2);
            else remover.filterSE(opts.input_1, opts.output_1);
        } else {
            HashDupRemover remover(opts.memLimit, opts.fasta, opts.verbose, opts.device);
            if (opts.paired) remover.filterPE(opts.input_1, opts.input_2, opts.output_1, opts.output_2, opts.unordered);
            else remover.filterSE(opts.input_1, opts.output_1);
        }
    } catch (const std::exception& exc) {
        std::cerr << "An error occured during fastq-dupaway execution:\n";
        std::cerr << exc.what() << '\n';
        return 1;
    } catch (...) {
        std::cerr << "Unknown error occured during fastq-dupaway execution!\n";
        return 1;
    }
    trace("done");
    // outputs are closed; the rest (CUDA context, pinned staging, worker threads) is left to process exit
    std::cout.flush();
    std::fflush(nullptr);
    if (!std::getenv("FQD_ORDERLY_EXIT")) _exit(0);
    return 0;
}
