// XMFA writer of parsnp_b200_core: restates Aligner::writeOutput (src/parsnp.cpp:505-1079) on the flat MUM/LCB lists of
// the C ABI.  MUM columns lower case, inter-MUM regions aligned with the reference's vendored libMUSCLE (upper case),
// LCB records `> i:start+1-end +/- cluster<z+1> s<contig>:p<pos>` (src/parsnp.cpp:980-1054).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace pb200 {

struct XmfaInput {
    int n = 0;
    std::vector<const std::string*> genomes;             // ingested texts
    std::vector<std::string> fasta_names, headers;       // file basenames, first FASTA lines
    std::vector<int64_t> genome_sizes;                   // text size minus contig padding
    std::vector<std::map<int, std::string>> pos2hdr;     // contig start -> "s<k>"
    int c = 21, doalign = 2, cores = 1;
    bool recombfilter = false;                           // ini [LCB] recombfilter: every LCB also goes to <outdir>/blocks/b<k>/seq.fna
    std::string outdir;                                  // (only used for blocks/)
    // clusters in this->clusters order
    std::vector<int32_t> ctype;
    std::vector<int64_t> cstart, cend;                   // [K*n]
    std::vector<int64_t> cmum_off, cmum_idx;             // per cluster MUM indices
    // MUM list
    std::vector<int64_t> mlen, mstart, mend;             // [M], [M*n], [M*n]
    std::vector<uint8_t> mfwd;                           // [M*n]
};

// aligns with MUSCLE (reference settings: DNA, 1 iteration, stable, ClustalW weights - src/MuscleInterface.cpp:38-49)
bool muscle_align(const std::vector<std::string>& seqs, std::vector<std::string>& out);
bool muscle_available();

// returns false when libMUSCLE is not linked
bool write_xmfa(const XmfaInput& in, const std::string& path);

// parsnp.unalign (Aligner::setUnalignableRegions, src/parsnp.cpp:2310-2382) from the records of pb200_result_unaligned
bool write_unaligned(const XmfaInput& in, const std::vector<int32_t>& genome, const std::vector<int64_t>& start,
                     const std::vector<int64_t>& end, const std::string& path);

// what the statistics block of parsnpAligner.log needs besides the MUM / LCB lists
struct LogInput {
    std::vector<std::string> files;                      // the paths as the ini gives them
    std::vector<int64_t> a, c, g, t;                     // base counts per genome (ingest)
    int d = 300, q = 30, filter = 1;
    float anchor_size = 0;                               // Calculator(anchors, shortest genome)
    long anchors_found = 0, mums_filtered = 0, clusters_filtered = 0;
    std::vector<int64_t> cnm;                            // MUMs per cluster
    double t_anchor = 0, t_coarsen = 0, t_lcb = 0, t_total = 0;
};

// statistics block of parsnpAligner.log (src/parsnp.cpp:1082-1190), line for line; uses n, fasta_names, genome_sizes, c,
// ctype, cmum_off/cmum_idx, mlen, mstart, mfwd of `in`
bool write_log(const XmfaInput& in, const LogInput& li, const std::string& path);

}  // namespace pb200
