// pb200_result: flat copy of the Aligner's final MUM/LCB lists behind the C ABI (include/parsnp_b200.h).
#pragma once
#include <vector>
#include <string>
#include <cstdint>
#include "aligner.h"
#include "../../../include/parsnp_b200.h"

struct pb200_result {
    int n = 0;
    pb200::pod_vector<int64_t> m_length, m_slength, m_start, m_end;
    pb200::pod_vector<uint8_t> m_fwd;
    std::vector<int32_t> c_type;
    std::vector<int64_t> c_nmums, c_length, c_start, c_end;
    std::vector<int64_t> c_mum_off, c_mum_idx;      // MUM indices (into the MUM list) of every cluster
    std::vector<int32_t> u_genome;                  // unaligned regions (PB200_FLAG_UNALIGNED)
    std::vector<int64_t> u_start, u_end;
    std::vector<int64_t> trace;
    std::vector<double> stats;
};

namespace pb200 {
extern thread_local std::string g_last_error;
pb200_result* make_result(const Aligner& a, bool unaligned = false);
AlignParams to_align_params(const pb200_params* p);
int default_host_threads();
void install_backtrace_handler();   // PB200_BACKTRACE=1      // PB200_HOST_THREADS, else min(32, cores / local ranks)
}
