// seqmode.cuh - whole-input modes (sequence-based sort + comparator scan, --fast --unordered tag join).
#pragma once
#include <string>
#include "../../include/fqd.h"
#include "common.cuh"

namespace fqd {
struct SeqState { int dummy; };
static int seq_create(SeqState** out, const fqd_config*, cudaStream_t, int, u32, std::string* err) {
    *out = nullptr; *err = "sequence / unordered modes are not built yet"; return FQD_ERR_INVALID;
}
static void seq_destroy(SeqState*) {}
static int seq_reset(SeqState*, std::string*) { return FQD_ERR_INVALID; }
static int seq_append(SeqState*, int, const void*, size_t, bool, std::string*) { return FQD_ERR_INVALID; }
static int seq_finish(SeqState*, std::string*) { return FQD_ERR_INVALID; }
static int seq_emission(SeqState*, fqd_emission_t*, std::string*) { return FQD_ERR_INVALID; }
static void seq_stats(SeqState*, fqd_stats_t*) {}
static double seq_device_ms(SeqState*) { return 0.0; }
static u64 seq_launches(SeqState*) { return 0; }
}  // namespace fqd
