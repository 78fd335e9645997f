#include "options.hpp"

#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace fqdhost {
namespace {

struct OptDef { const char* long_name; char short_name; bool takes_value; const char* text; };

const OptDef DEFS[] = {
    {"help", 'h', false, "Produce help message and exit"},
    {"verbose", 'v', false, "Report run summary after program execution."},
    {"input-1", 'i', true, "First input file (required)"},
    {"input-2", 'u', true, "Second input file (optional, enables paired-end mode)"},
    {"output-1", 'o', true, "First output file (required)"},
    {"output-2", 'p', true, "Second output file (optional, required for paired-end mode)"},
    {"mem-limit", 'm', true,
     "Memory limit in megabytes (default 2048 = 2Gb).\nSupported value range is [500 <-> 10240 (10 Gb)]\n"
     "On the B200 build this bounds the pinned host staging buffers; device memory is sized from the GPU."},
    {"format", 0, true, "input file format: fastq (default) or fasta."},
    {"compare-seq", 0, true,
     "Sequence comparison mode for deduplication step.\nSupported options:\n"
     "\ttight (default): compare sequences directly, sequences of different lengths are considered different.\n"
     "\tloose: compare sequences directly, sequences of different lengths are considered duplicates if shorter"
     " sequence exactly matches with prefix of longer sequence.\n"
     "\ttail-hamming: considers a pair of sequences as duplicates if those differ by no more than a set number of"
     " mismatches (default 2). Sequences of different lengths will not be compared."},
    {"distance", 0, true, "A threshold value for 'tail-hamming' distance calculation. Should be a non-negative integer. Default value is 2."},
    {"write-clusters", 0, false,
     "Write ids of identified duplicate clusters to <output-file>.clusters (2 files in paired mode).\n"
     "This option is only supported by the sequence-based modes."},
    {"fast", 0, false,
     "Use hash-based approach instead of sequence-based.\nOnly complete duplicates will be filtered out."},
    {"unordered", 0, false,
     "This option is supported only by 'fast' mode for paired inputs.\n"
     "If enabled, both input files will be sorted by read IDs before deduplication."},
};
const int NDEFS = sizeof(DEFS) / sizeof(DEFS[0]);

void print_help(std::ostream& os) {
    os << VERSION << "\n";
    os << "Supported options:\n";
    for (int i = 0; i < NDEFS; ++i) {
        std::string head = "  ";
        if (DEFS[i].short_name) { head += "-"; head += DEFS[i].short_name; head += " [ --"; head += DEFS[i].long_name; head += " ]"; }
        else { head += "--"; head += DEFS[i].long_name; }
        if (DEFS[i].takes_value) head += " arg";
        os << head << "\n        ";
        for (const char* c = DEFS[i].text; *c; ++c) { os << *c; if (*c == '\n') os << "        "; }
        os << "\n";
    }
    os << "\n";
}

const OptDef* find_long(const std::string& n) {
    const OptDef* hit = nullptr; int nhit = 0;
    for (int i = 0; i < NDEFS; ++i) {
        std::string ln = DEFS[i].long_name;
        if (ln == n) return &DEFS[i];
        if (!n.empty() && ln.compare(0, n.size(), n) == 0) { hit = &DEFS[i]; ++nhit; }
    }
    if (nhit == 1) return hit;
    if (nhit > 1) throw std::runtime_error("option '--" + n + "' is ambiguous");
    throw std::runtime_error("unrecognised option '--" + n + "'");
}
const OptDef* find_short(char c) {
    for (int i = 0; i < NDEFS; ++i) if (DEFS[i].short_name == c) return &DEFS[i];
    throw std::runtime_error(std::string("unrecognised option '-") + c + "'");
}

long long parse_int(const std::string& name, const std::string& text, bool is_unsigned) {
    std::istringstream is(text);
    long long v = 0;
    if (is_unsigned && !text.empty() && text[0] == '-')
        throw std::runtime_error("the argument ('" + text + "') for option '--" + name + "' is invalid");
    is >> v;
    if (is.fail() || !is.eof())
        throw std::runtime_error("the argument ('" + text + "') for option '--" + name + "' is invalid");
    return v;
}

}  // namespace

bool parse_args(int argc, char** argv, Options& opts) {
    try {
        std::map<std::string, std::string> vm;
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            const OptDef* o = nullptr;
            std::string val; bool has_val = false;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                std::string n = a.substr(2);
                size_t eq = n.find('=');
                if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); has_val = true; }
                o = find_long(n);
            } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
                o = find_short(a[1]);
                if (a.size() > 2) { val = a.substr(2); has_val = true; }
            } else {
                throw std::runtime_error("too many positional options have been specified on the command line");
            }
            if (o->takes_value && !has_val) {
                if (i + 1 >= argc) throw std::runtime_error(std::string("the required argument for option '--") + o->long_name + "' is missing");
                val = argv[++i];
            }
            if (vm.count(o->long_name)) throw std::runtime_error(std::string("option '--") + o->long_name + "' cannot be specified more than once");
            vm[o->long_name] = val;
        }
        if (vm.count("help")) { print_help(std::cerr); return false; }        // src/main.cpp:85-90 (stderr, exit 1)
        if (!vm.count("input-1")) throw std::runtime_error("the option '--input-1' is required but missing");
        if (!vm.count("output-1")) throw std::runtime_error("the option '--output-1' is required but missing");
        opts.input_1 = vm["input-1"]; opts.output_1 = vm["output-1"];
        if (vm.count("input-2")) opts.input_2 = vm["input-2"];
        if (vm.count("output-2")) opts.output_2 = vm["output-2"];
        opts.verbose = vm.count("verbose") > 0;
        opts.write_clusters = vm.count("write-clusters") > 0;
        opts.unordered = vm.count("unordered") > 0;
        const bool hash_opt = vm.count("fast") > 0;
        if (vm.count("distance")) opts.hammdist = (unsigned)parse_int("distance", vm["distance"], true);
        long long mem = 0;
        if (vm.count("mem-limit")) mem = parse_int("mem-limit", vm["mem-limit"], false);

        // the rules of src/main.cpp:93-164, in the same order
        if ((vm.count("input-2") > 0) != (vm.count("output-2") > 0))
            throw std::runtime_error("Both input-2 and output-2 arguments are required for paired-end mode!");
        if (vm.count("input-2")) opts.paired = true;
        if (vm.count("input-2")) {
            if (opts.input_1 == opts.input_2) throw std::runtime_error("Paired input files should not be the same file!");
            if (opts.output_1 == opts.output_2) throw std::runtime_error("Paired output files should not be the same file!");
        }
        if (vm.count("format")) {
            const std::string& v = vm["format"];
            if (v == "fastq") ;
            else if (v == "fasta") opts.fasta = true;
            else throw std::runtime_error("Only \"fastq\" or \"fasta\" file formats are supported!");
        }
        if (vm.count("compare-seq")) {
            const std::string& v = vm["compare-seq"];
            if (v == "tight") ;
            else if (v == "loose") opts.ctype = CT_LOOSE;
            else if (v == "tail-hamming") opts.ctype = CT_HAMMING;
            else throw std::runtime_error("Unsupported compare-seq type provided!");
        }
        if (vm.count("mem-limit")) {
            if (mem >= 500 && mem <= 10240) opts.memLimit = (ssize_t)mem * 1024L * 1024L;
            else throw std::runtime_error("Value of unsupported range provided for --mem-limit option!");
        }
        if (hash_opt) {
            opts.hash = true;
            opts.ctype = CT_NONE;
            if (vm.count("compare-seq") || vm.count("distance") || opts.write_clusters)
                throw std::runtime_error("--fast mode was enabled, but argument(s) for sequence-based mode were provided!");
        }
        if (opts.unordered) {
            if (!hash_opt) throw std::runtime_error("--unordered argument can only be used with --fast mode!");
            if (!vm.count("input-2")) throw std::runtime_error("--unordered argument can only be used with paired inputs!");
        }
        if (const char* d = getenv("FQD_DEVICE")) opts.device = atoi(d);
    } catch (const std::exception& e) {
        std::cerr << "An error occured during arguments parsing:\n";
        std::cerr << e.what() << '\n';
        return false;
    } catch (...) {
        std::cerr << "Unknown error occured during arguments parsing!\n";
        return false;
    }
    return true;
}

}  // namespace fqdhost
