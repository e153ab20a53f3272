// TEST-ONLY tool: writes the XMFA from an oracle MUM/LCB dump (hook H1 of oracle/build_ref.py) through the product's XMFA
// writer (parsnp_b200/csrc/main/xmfa.cpp), so the writer can be checked against the reference's XMFA md5 without a GPU.
//   xmfa_from_dump <ini> <dump.txt> <out.xmfa> [<outdir for blocks/, parsnp.unalign, parsnpAligner.log>
//                  [<unaligned records: "genome start end" lines, or -> [<log counters: "anchors_found mums_filtered clusters_filtered">]]]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include "../parsnp_b200/csrc/host/ingest.h"
#include "../parsnp_b200/csrc/host/minsize.h"
#include "../parsnp_b200/csrc/main/xmfa.h"
using namespace std;
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    pb200::IniFile ini;
    ini.read(argv[1]);
    const int d = ini.get_i("LCB", "d");
    const int qfiles = (int)ini.num_values("Query") / 2;
    vector<pb200::IngestedGenome> G((size_t)qfiles + 1);
    pb200::XmfaInput xi;
    xi.n = qfiles + 1;
    for (int i = 0; i <= qfiles; i++) {
        string path; bool rev;
        char b[64];
        if (i == 0) { path = ini.get("Reference", "file"); rev = ini.get_b("Reference", "reverse"); }
        else { snprintf(b, sizeof b, "file%d", i); path = ini.get("Query", b); snprintf(b, sizeof b, "reverse%d", i); rev = ini.get_b("Query", b); }
        if (!pb200::ingest_fasta(path, i == 0, d, rev, G[i])) return 3;
        size_t loc = path.rfind('/');
        xi.fasta_names.push_back(loc == string::npos ? path : path.substr(loc + 1));
    }
    for (auto& g : G) {
        xi.genomes.push_back(&g.text); xi.headers.push_back(g.header); xi.genome_sizes.push_back((int64_t)g.text.size() - g.padding);
        map<int, string> p2h; p2h[1] = "s1";
        for (size_t k = 0; k + 1 < g.contig_ends.size(); k++) p2h[(int)g.contig_ends[k]] = "s" + to_string(k + 2);
        xi.pos2hdr.push_back(p2h);
    }
    xi.c = ini.get_i("LCB", "c"); xi.doalign = ini.get_i("LCB", "doalign"); xi.cores = 1;
    ifstream f(argv[2]);
    string line;
    vector<vector<int64_t>> members;
    vector<int64_t> cnm;
    bool located_ok = true, explicit_lists = false;
    while (getline(f, line)) {
        istringstream is(line);
        string tag; is >> tag;
        if (tag == "M") {
            long len, sl; is >> len >> sl; xi.mlen.push_back(len);
            string tok;
            while (is >> tok) { long a, b; int fw; sscanf(tok.c_str(), "%ld:%ld:%d", &a, &b, &fw); xi.mstart.push_back(a); xi.mend.push_back(b); xi.mfwd.push_back((uint8_t)fw); }
        } else if (tag == "C") {
            int type; long cnt, len; is >> type >> cnt >> len; xi.ctype.push_back(type); cnm.push_back(cnt);
            string tok;
            while (is >> tok) { long a, b; sscanf(tok.c_str(), "%ld:%ld", &a, &b); xi.cstart.push_back(a); xi.cend.push_back(b); }
            // An LCB owns a consecutive run of the MUM list (sorted by reference start); MUMs of clusters the filter removed
            // stay in the list between the runs, so the run is located by the LCB's reference start and checked by its end - unless an "I" line
            // (the product's own cluster -> MUM index list, tests/refcmp.py write_dump) follows and replaces it.
            members.emplace_back();
            if (type == 1) {
                const size_t N = (size_t)xi.n;
                const long c0 = xi.cstart[xi.cstart.size() - N], c1 = xi.cend[xi.cend.size() - N];
                size_t m = 0;
                while (m < xi.mlen.size() && xi.mstart[m * N] != c0) m++;
                for (long k = 0; k < cnt && m + k < xi.mlen.size(); k++) members.back().push_back((int64_t)(m + k));
                if ((long)members.back().size() != cnt || xi.mend[(m + cnt - 1) * N] != c1) located_ok = false;
            }
        } else if (tag == "I" && !members.empty()) {
            members.back().clear();
            long m;
            while (is >> m) members.back().push_back(m);
            explicit_lists = true;
        }
    }
    if (!located_ok && !explicit_lists) { fprintf(stderr, "xmfa_from_dump: an LCB interval does not hold the number of MUMs the dump states\n"); return 6; }
    xi.cmum_off.push_back(0);
    for (auto& v : members) { xi.cmum_idx.insert(xi.cmum_idx.end(), v.begin(), v.end()); xi.cmum_off.push_back((int64_t)xi.cmum_idx.size()); }
    if (argc > 4) { xi.outdir = argv[4]; xi.recombfilter = ini.get_b("LCB", "recombfilter"); }
    if (!pb200::write_xmfa(xi, argv[3])) return 4;
    if (argc > 6) {                             // parsnpAligner.log through the product's writer (csrc/main/xmfa.cpp write_log)
        pb200::LogInput li;
        ifstream sf(argv[6]);
        sf >> li.anchors_found >> li.mums_filtered >> li.clusters_filtered;
        int64_t slength = 500000000;
        for (int i = 0; i <= qfiles; i++) {
            char b[64];
            if (i == 0) li.files.push_back(ini.get("Reference", "file"));
            else { snprintf(b, sizeof b, "file%d", i); li.files.push_back(ini.get("Query", b)); }
            li.a.push_back(G[i].a); li.c.push_back(G[i].c); li.g.push_back(G[i].g); li.t.push_back(G[i].t);
            slength = min<int64_t>(slength, (int64_t)G[i].text.size());
        }
        li.d = d; li.q = ini.get_i("LCB", "q"); li.filter = ini.get_i("MUM", "filter");
        li.anchor_size = (float)pb200::MinSizeExpr(ini.get("MUM", "anchors"))(slength);
        li.cnm = cnm;
        if (!pb200::write_log(xi, li, string(argv[4]) + "/parsnpAligner.log")) return 7;
    }
    if (argc > 5 && string(argv[5]) != "-") {
        ifstream uf(argv[5]);
        vector<int32_t> ug; vector<int64_t> us, ue;
        long a, b, c;
        while (uf >> a >> b >> c) { ug.push_back((int32_t)a); us.push_back(b); ue.push_back(c); }
        if (!pb200::write_unaligned(xi, ug, us, ue, string(argv[4]) + "/parsnp.unalign")) return 5;
    }
    return 0;
}
