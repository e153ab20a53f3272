#include "xmfa.h"
#include <atomic>
#include <sstream>
#include <sys/stat.h>
#include <sys/types.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <fstream>
#include <iostream>

namespace pb200 {

namespace {
std::string reversec(const std::string& s) {           // Aligner::reversec (src/parsnp.cpp:1294-1393) on the ingest alphabet
    std::string g;
    g.reserve(s.size());
    for (char ch : s) {
        switch (toupper((unsigned char)ch)) {
            case 'A': g.push_back('T'); break;
            case 'G': g.push_back('C'); break;
            case 'C': g.push_back('G'); break;
            case 'T': g.push_back('A'); break;
            default: g.push_back('N'); break;
        }
    }
    std::reverse(g.begin(), g.end());
    return g;
}
std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }
std::string upper(std::string s) { for (auto& c : s) c = (char)toupper((unsigned char)c); return s; }
std::string sub(const std::string& g, int64_t pos, int64_t count) {     // std::string::substr semantics (count clamps; negative = npos)
    if (pos < 0 || (size_t)pos > g.size()) return std::string();
    return g.substr((size_t)pos, count < 0 ? std::string::npos : (size_t)count);
}
}  // namespace

bool write_xmfa(const XmfaInput& in, const std::string& path) {
    const int n = in.n;
    const int64_t K = (int64_t)in.ctype.size();
    std::vector<std::vector<std::string>> aln((size_t)K, std::vector<std::string>((size_t)n));
    auto MS = [&](int64_t m, int i) { return in.mstart[(size_t)m * n + i]; };
    auto ME = [&](int64_t m, int i) { return in.mend[(size_t)m * n + i]; };
    auto MF = [&](int64_t m, int i) { return in.mfwd[(size_t)m * n + i] != 0; };

    // ---- alignment assembly per LCB (src/parsnp.cpp:646-919); MUSCLE runs on the inter-MUM regions
    std::atomic<bool> muscle_failed(false);          // (set from the OpenMP threads of the loop below)
#pragma omp parallel for schedule(dynamic) num_threads(in.cores > 0 ? in.cores : 1)
    for (int64_t z = 0; z < K; z++) {
        const int64_t m0 = in.cmum_off[z], m1 = in.cmum_off[z + 1];
        if (!(in.ctype[z] == 1 && m1 > m0 && in.doalign != 0)) continue;
        const int64_t first = in.cmum_idx[m0];
        std::vector<std::string>& T = aln[z];
        for (int64_t t = m0; t < m1; t++) {
            const int64_t dt = in.cmum_idx[t];
            const bool has_next = t + 1 < m1;
            const int64_t nx = has_next ? in.cmum_idx[t + 1] : -1;
            std::vector<std::string> reg((size_t)n);
            if (m1 - m0 == 1) {
                for (int i = 0; i < n; i++) {
                    std::string s = sub(*in.genomes[i], MS(dt, i), in.mlen[dt]);
                    T[i].append(lower(MF(first, i) ? s : reversec(s)));
                }
            } else if (has_next) {
                for (int i = 0; i < n; i++) {
                    if (!MF(first, i)) {
                        if (i == 0) std::cout << "Error!! MUM in - orientation for ref genome**" << std::endl;
                        T[i].append(lower(reversec(sub(*in.genomes[i], MS(dt, i), in.mlen[dt]))));
                        if (MS(dt, i) - ME(nx, i) >= 1) {
                            std::string s1 = reversec(sub(*in.genomes[i], ME(nx, i), MS(dt, i) - ME(nx, i)));
                            reg[i] = s1.size() >= 1 ? upper(s1) : std::string();
                        }
                    } else {
                        T[i].append(lower(sub(*in.genomes[i], MS(dt, i), in.mlen[dt])));
                        std::string s1 = sub(*in.genomes[i], ME(dt, i), MS(nx, i) - ME(dt, i));
                        reg[i] = s1.size() >= 1 ? upper(s1) : std::string();
                    }
                }
            }
            if (has_next && in.doalign) {
                size_t maxl = 0, minl = 1000000;
                for (int k = 0; k < n; k++) { maxl = std::max(maxl, reg[k].size()); minl = std::min(minl, reg[k].size()); }
                if (maxl > 1 && minl > 0) {
                    if (n - 1 <= 0) {
                        for (int j = 0; j < n; j++) T[j].append(reg[j]);       // total_skipped (0) >= nnum-1 only when nnum == 1
                    } else {
                        std::vector<std::string> res;
                        if (!muscle_align(reg, res)) muscle_failed.store(true, std::memory_order_relaxed);
                        for (size_t i = 0; i < res.size() && i < (size_t)n; i++) T[i].append(res[i]);
                    }
                } else if (maxl > 0) {
                    for (int p = 0; p < n; p++) {
                        if (reg[p].size() > 0) T[p].append(upper(reg[p]));
                        if (reg[p].size() < maxl) T[p].append(maxl - reg[p].size(), '-');
                    }
                }
            }
            if (!has_next && m1 - m0 > 1) {
                for (int i = 0; i < n; i++) {
                    std::string s = sub(*in.genomes[i], MS(dt, i), in.mlen[dt]);
                    T[i].append(lower(MF(first, i) ? s : reversec(s)));
                }
            }
        }
    }
    if (muscle_failed.load()) return false;

    // ---- recombfilter: blocks/ and one directory per LCB that has MUMs (src/parsnp.cpp:538-543, 605-644)
    const std::string lcbprefix = in.outdir + "/blocks/b";
    if (in.recombfilter) {
        ::mkdir((in.outdir + "/blocks/").c_str(), 0777);
        for (int64_t z = 0; z < K; z++) {
            if (!(in.ctype[z] == 1 && in.cmum_off[z + 1] > in.cmum_off[z] && in.doalign != 0)) continue;
            ::mkdir((lcbprefix + std::to_string(z + 1)).c_str(), 0777);
        }
    }
    // ---- XMFA (src/parsnp.cpp:586-598, 921-1071)
    std::ofstream x(path.c_str());
    long total_clusters = 0;
    for (int64_t z = 0; z < K; z++) if (in.ctype[z] == 1) total_clusters++;
    x << "#FormatVersion Mauve" << std::endl;
    x << "#SequenceCount " << n << std::endl;
    for (int z = 0; z < n; z++) {
        x << "##SequenceIndex " << z + 1 << std::endl;
        x << "##SequenceFile " << in.fasta_names[z] << std::endl;
        x << "##SequenceHeader " << in.headers[z] << std::endl;
        x << "##SequenceLength " << in.genome_sizes[z] << "bp" << std::endl;
    }
    x << "#IntervalCount " << total_clusters << std::endl;
    int prev_end = 0;
    for (int64_t z = 0; z < K; z++) {
        const int64_t m0 = in.cmum_off[z], m1 = in.cmum_off[z + 1];
        std::vector<std::string>& T = aln[z];
        if (!(in.ctype[z] == 1 && m1 > m0 && in.doalign != 0 && T[0].size() > (size_t)in.c)) continue;
        std::vector<int64_t> cs(in.cstart.begin() + z * n, in.cstart.begin() + (z + 1) * n);
        std::vector<int64_t> ce(in.cend.begin() + z * n, in.cend.begin() + (z + 1) * n);
        const int64_t first = in.cmum_idx[m0], last = in.cmum_idx[m1 - 1];
        // overlap trim with the previous LCB, reproduced literally (src/parsnp.cpp:927-951).  The scan position of the first
        // loop never advances, so cols_to_trim == overlap.  The second loop counts the non-gap columns of row 0 over the length
        // of row i - but row 0 has lost its head by the time rows 1.. are counted, and the reference keeps indexing it up to the
        // old length: behind the new terminator (which counts as a base) std::string still holds the old tail, unmoved.
        const int lcb_start = (int)cs[0] + 1, lcb_end = (int)ce[0];
        int overlap = std::max(0, prev_end - lcb_start);
        if (overlap > 0 && !T[0].empty() && T[0][0] != '-') {
            const int cols_to_trim = overlap;
            const std::string row0 = T[0];
            const size_t kept = row0.size() - std::min(row0.size(), (size_t)cols_to_trim);
            for (int i = 0; i < n; i++) {
                int cnt = 0;
                for (size_t pos = 0; pos < T[i].size(); pos++) {
                    char ch;
                    if (i == 0) ch = pos < row0.size() ? row0[pos] : '\0';
                    else if (pos < kept) ch = row0[pos + (row0.size() - kept)];
                    else ch = (pos == kept || pos >= row0.size()) ? '\0' : row0[pos];
                    if (ch != '-') cnt++;
                }
                cs[i] += cnt;
                T[i].erase(0, (size_t)cols_to_trim);
            }
        }
        prev_end = lcb_end;
        if (!(T[0].size() > (size_t)in.c)) continue;
        char b[16];
        snprintf(b, sizeof b, "%d", (int)z + 1);
        std::ofstream clcb;                                  // the LCB's own copy (records only, no "=" line; src/parsnp.cpp:958-963)
        if (in.recombfilter) clcb.open((lcbprefix + b + "/seq.fna").c_str());
        for (int i = 0; i < n; i++) {
            const std::string& s1s = T[i];
            const bool fwd = MF(first, i);
            std::ostringstream rec;
            if (fwd) rec << "> " << i + 1 << ":" << cs[i] + 1 << "-" << ce[i] << " ";
            else rec << "> " << i + 1 << ":" << MS(last, i) + 1 << "-" << ME(first, i) << " ";
            bool hit1 = false, hit2 = false;
            std::string hdr1, lasthdr1;
            int seqstart = 0, laststart = 0;
            for (auto it = in.pos2hdr[i].begin(); it != in.pos2hdr[i].end(); ++it) {
                if (hit1 && cs[i] < it->first) { hit2 = true; hdr1 = lasthdr1; seqstart = laststart; break; }
                else if (cs[i] >= it->first && !hit2) { hit1 = true; laststart = it->first; lasthdr1 = it->second; continue; }
                else if (hit1 & hit2) { hdr1 = lasthdr1; seqstart = laststart; break; }
            }
            if (hit1 && !hit2) { hdr1 = lasthdr1; seqstart = laststart; }
            int offset = 0;
            if (hdr1 == "") { hdr1 = "s1"; offset = -1; }
            else if (hdr1 != "s1") offset = -1;
            if (!fwd) rec << "- cluster" << b << " " << hdr1 << ":p" << (cs[i] - seqstart) + 1 + in.mlen[first] + offset << "\n";
            else rec << "+ cluster" << b << " " << hdr1 << ":p" << (cs[i] - seqstart) + 1 + offset << "\n";
            const size_t width = 80;
            size_t k = 0;
            for (; k + width < s1s.size(); k += width) rec << s1s.substr(k, width) << "\n";
            rec << s1s.substr(k, s1s.size() - k) << "\n";
            x << rec.str();
            if (in.recombfilter) clcb << rec.str();
        }
        x << "=" << std::endl;
    }
    return true;
}

bool write_unaligned(const XmfaInput& in, const std::vector<int32_t>& genome, const std::vector<int64_t>& start,
                     const std::vector<int64_t>& end, const std::string& path) {
    std::ofstream u(path.c_str());
    if (!u) return false;
    for (size_t r = 0; r < genome.size(); r++) {
        const int k = genome[r];
        const int64_t startpos = start[r], endpos = end[r];
        u << ">" << k + 1 << ":" << startpos << "-" << endpos << " + " << in.fasta_names[(size_t)k] << "\n";
        const std::string& g = *in.genomes[(size_t)k];
        const std::string s1 = (startpos >= 0 && (size_t)startpos <= g.size()) ? g.substr((size_t)startpos, (size_t)std::max<int64_t>(0, endpos - startpos))
                                                                                : std::string();
        size_t pos = 0;
        while (pos + 80 < s1.size()) { u << s1.substr(pos, 80) << "\n"; pos += 80; }
        if (pos + 1 < s1.size()) u << s1.substr(pos, s1.size()) << "\n";       // (a last line of exactly one base is dropped, src/parsnp.cpp:2366)
        if (s1.size() == 0) u << "-" << "\n";
        u << "=" << "\n";
    }
    return true;
}

bool write_log(const XmfaInput& in, const LogInput& li, const std::string& path) {
    using namespace std;
    const int n = in.n;
    const int64_t K = (int64_t)in.ctype.size(), M = (int64_t)in.mlen.size();
    ofstream log(path.c_str());
    if (!log) return false;
    log << "Number of sequences analyzed:" << setiosflags(ios::fixed) << setprecision(1) << setw(10) << n << endl << endl;
    for (int i = 0; i < n; i++) {
        log << "Sequence " << i + 1 << " : " << li.files[i] << endl;
        log << in.fasta_names[i] << endl;
        log << "Length:" << setw(10) << (long)in.genome_sizes[i] << " bps" << endl;
        log << " GC:" << setw(10) << setiosflags(ios::fixed) << setprecision(1) << (float(li.g[i]) + float(li.c[i])) << endl;
        log << " AT:" << setw(10) << setiosflags(ios::fixed) << setprecision(1) << (float(li.a[i]) + float(li.t[i])) << endl;
    }
    log << setw(2) << setiosflags(ios::left) << "d value:   " << setw(2) << li.d << endl;
    log << setw(2) << "q value:   " << setw(2) << li.q << endl << endl;
    log << setw(2) << "Mum anchor size:   " << setw(2) << li.anchor_size << endl;
    log << setw(2) << "Number of MUM anchors found:   " << setw(2) << li.anchors_found << endl;
    if (M + li.mums_filtered >= li.anchors_found) log << setw(2) << "Number of MUMs found:   " << setw(2) << (M + li.mums_filtered) - li.anchors_found << endl;
    else log << setw(2) << "Number of MUMs found:   " << setw(2) << 0 << endl;
    log << setw(2) << "Total MUMs found((Anchors+MUMs)-filtered):   " << setw(2) << M << endl << endl;
    log << setw(2) << "Random MUM length:   " << setw(2) << li.filter << endl;
    log << setw(2) << "Minimum Cluster length:   " << setw(2) << in.c << endl;
    log << setw(2) << "Number of MUMs filtered:   " << setw(2) << li.mums_filtered << endl;
    log << setw(2) << "Number of Clusters filtered:   " << setw(2) << li.clusters_filtered << endl << endl;
    long ccount = 0;
    for (int64_t i = 0; i < K; i++) if (in.ctype[i] && li.cnm[i] > 0) ccount++;
    log << setw(2) << "Number of clusters created:   " << setw(2) << ccount << endl;
    if (K == 0) log << setw(2) << "Number of clusters created:   " << setw(2) << "NONE" << endl;
    if (ccount) log << setw(2) << "Average number of MUMs per cluster:   " << setw(2) << M / ccount << endl;     // (the reference divides by zero here)
    // LCB coverage per sequence (src/parsnp.cpp:1141-1160): |last MUM end - first MUM start| of every LCB, per strand
    vector<long> coverage(n, 0);
    long avg = 0, totcoverage = 0, totsize = 0;
    for (int i = 0; i < n; i++)
        for (int64_t c = 0; c < K; c++) {
            if (!in.ctype[c] || in.cmum_off[c + 1] <= in.cmum_off[c]) continue;
            const int64_t front = in.cmum_idx[in.cmum_off[c]], back = in.cmum_idx[in.cmum_off[c + 1] - 1];
            long span;
            if (in.mfwd[front * n + i]) span = labs((long)((in.mstart[back * n + i] + in.mlen[back]) - in.mstart[front * n + i]));
            else span = labs((long)((in.mstart[front * n + i] + in.mlen[front]) - in.mstart[back * n + i]));
            coverage[i] += span;
            if (i == 0) avg += span;
        }
    if (ccount) log << setw(2) << "Average cluster length:   " << avg / ccount << " bps" << endl;
    for (int i = 0; i < n; i++) {
        float percent = (float)coverage[i] / ((float)(li.g[i] + li.c[i]) + (float)(li.a[i] + li.t[i]));
        log << setw(2) << "Cluster coverage in sequence " << i + 1 << ":   " << setiosflags(ios::fixed) << setprecision(1) << 100.00 * percent << "%" << endl;
        totcoverage += coverage[i];
        totsize += (long)in.genome_sizes[i];
    }
    float percent = (float)totcoverage / (float)totsize;
    log << setw(2) << "Total coverage among all sequences:   " << setiosflags(ios::fixed) << setprecision(1) << 100.00 * percent << "%" << endl << endl;
    log << setw(2) << " MUM anchor search elapsed time:   " << li.t_anchor << "s " << endl;
    log << setw(2) << " MUM coarsening elapsed time:   " << li.t_coarsen << "s " << endl;
    if (li.filter) log << setw(2) << " MUM filtering elapsed time:   " << 0.0 << "s " << endl;
    log << setw(2) << " MUM clustering elapsed time:   " << li.t_lcb << "s " << endl;
    log << setw(2) << " Inter-clustering elapsed time:   " << 0.0 << "s " << endl;
    log << setw(2) << " Total running time:   " << li.t_total << "s " << endl;
    return true;
}

}  // namespace pb200
