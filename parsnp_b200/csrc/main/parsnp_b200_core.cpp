// parsnp_b200_core <ini> - the process boundary of parsnp_core (src/parsnp.cpp:2792-3299) on top of libparsnp_b200.so.
//
// Same argv (-h, -v, ini path), same ini keys and FASTA ingest rules, same exit codes (0 also for "NO MUMS FOUND", 1 for a
// missing ini / reference).  calcmumi=1 -> <outdir>/all.mumi;  otherwise the MUM + LCB path -> <outdir>/parsnpAligner.xmfa
// (inter-MUM regions aligned with the reference's vendored libMUSCLE, linked unchanged), <outdir>/parsnpAligner.log
// (MUMS FOUND / NO MUMS FOUND, then the statistics block the Python driver parses, parsnp:1530-1536) and
// <outdir>/parsnpAligner.mums (MUM and LCB coordinates); recombfilter=1 -> <outdir>/blocks/b<k>/seq.fna (src/parsnp.cpp:538-543,
// 958-963); unaligned=1 -> <outdir>/parsnp.unalign (Aligner::setUnalignableRegions, src/parsnp.cpp:2310-2382).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <sys/file.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/types.h>
#include "../../../include/parsnp_b200.h"
#include "../host/ingest.h"
#include "xmfa.h"

using namespace std;

static string basename_of(const string& p) {       // src/parsnp.cpp:2925-2931
    size_t loc = p.rfind('/');
    return loc == string::npos ? p : p.substr(loc + 1);
}

// one advisory lock file per device ordinal; returns the chosen ordinal (the descriptor stays open = locked until exit)
static int choose_device() {
    if (const char* e = getenv("PB200_DEVICE")) return atoi(e);
    const int count = pb200_device_count();
    if (count <= 1) return 0;
    const char* dir = getenv("PB200_LOCK_DIR");
    const string base = string(dir && *dir ? dir : "/tmp") + "/parsnp_b200.gpu";
    const int first = (int)(getpid() % count);
    for (int t = 0; t < count; t++) {
        const int k = (first + t) % count;                  // start at pid % count: concurrent starters probe different files first
        const string path = base + to_string(k) + ".lock";
        int fd = open(path.c_str(), O_CREAT | O_RDWR, 0666);
        if (fd < 0) continue;
        if (flock(fd, LOCK_EX | LOCK_NB) == 0) return k;    // (never closed: released by the kernel at exit)
        close(fd);
    }
    return first;                                           // more processes than devices: share round-robin
}

int main(int argc, char** argv) {
    bool help = false, version = false;
    for (int i = 0; i < argc; i++)
        if (argv[i][0] == '-') { if (argv[i][1] == 'h') help = true; else if (argv[i][1] == 'v') version = true; }
    if (help) {
        cout << "parsnp options:" << endl << "   -h <display this message>" << endl << "   -v <display the version>" << endl
             << "   <parameter file with options>" << endl;
        return 0;
    }
    if (version) { cout << "Parsnp v1.0.1 (" << pb200_version() << ")" << endl; return 0; }
    if (argc < 2) { cout << "ERROR: No parameter file specified!" << endl; return 1; }
    pb200::IniFile ini;
    ini.read(argv[1]);
    pb200_params prm;
    pb200_params_default(&prm);
    prm.c = ini.get_i("LCB", "c");
    prm.d = ini.get_i("LCB", "d");
    prm.diagdiff = (float)ini.get_f("LCB", "diagdiff");
    prm.q = ini.get_i("LCB", "q");
    prm.p = ini.get_i("LCB", "p");
    const string anchors = ini.get("MUM", "anchors"), mums = ini.get("MUM", "mums");
    prm.anchors = anchors.c_str();
    prm.mums = mums.c_str();
    prm.anchors_only = ini.get_b("MUM", "anchorsonly");
    prm.filter = ini.get_i("MUM", "filter");
    const bool calc_mumi = ini.get_b("MUM", "calcmumi");
    const bool recomb_filter = ini.get_b("LCB", "recombfilter"), do_unalign = ini.get_b("LCB", "unaligned");
    if (do_unalign) prm.flags |= PB200_FLAG_UNALIGNED;
    const string outdir = ini.get("Output", "outdir", "output");
    const bool reverse_ref = ini.get_b("Reference", "reverse");
    const int qfiles = (int)ini.num_values("Query") / 2;

    vector<pb200::IngestedGenome> G((size_t)qfiles + 1);
    vector<string> files, names;
    for (int i = 0; i <= qfiles; i++) {
        string path;
        bool rev;
        if (i == 0) { path = ini.get("Reference", "file"); rev = reverse_ref; }
        else {
            char b[64];
            snprintf(b, sizeof b, "file%d", i); path = ini.get("Query", b);
            snprintf(b, sizeof b, "reverse%d", i); rev = ini.get_b("Query", b);
        }
        if (!pb200::ingest_fasta(path, i == 0, prm.d, rev, G[i])) {
            if (i == 0) cout << " Cannot open reference file ! " << endl;
            else cout << " Cannot open query file: " << path << endl;
            return 1;
        }
        files.push_back(path);
        names.push_back(basename_of(path));
        const double sz = (double)G[i].text.size();
        cout << names.back() << ",Len:" << G[i].text.size() << ",GC:" << ((float(G[i].g) + float(G[i].c)) / float(sz - G[i].n)) * 100 << endl;
    }
    const int n = qfiles + 1;
    vector<const uint8_t*> seqs;
    vector<int64_t> lens;
    for (auto& g : G) { seqs.push_back((const uint8_t*)g.text.data()); lens.push_back((int64_t)g.text.size()); }
    // GPU choice is implicit: the Python driver launches up to min(threads, partitions) concurrent copies of this binary with
    // identical arguments apart from the ini (parsnp:1572-1597), so the processes spread themselves over the visible devices
    // (CUDA_VISIBLE_DEVICES is honoured by the runtime): PB200_DEVICE if set, else the first device whose advisory lock file
    // is free, else pid % devices.  The lock (flock on /tmp/parsnp_b200.gpu<k>.lock) is held until the process exits.
    int device = choose_device();

    pb200_genomes* dev = nullptr;
    int rc = pb200_genomes_create(device, n, seqs.data(), lens.data(), &dev);
    if (rc != 0) { cerr << "parsnp_b200_core: " << pb200_last_error() << endl; return 1; }

    if (calc_mumi) {                                        // src/parsnp.cpp:1964-2080
        cerr << "Calculating mumi distances.." << endl;
        vector<double> v((size_t)max(qfiles, 1));
        rc = pb200_mumi(dev, &prm, v.data());
        if (rc != 0) { cerr << "parsnp_b200_core: " << pb200_last_error() << endl; return 1; }
        FILE* f = fopen((outdir + "/all.mumi").c_str(), "w");
        if (!f) { cerr << "parsnp_b200_core: cannot write " << outdir << "/all.mumi" << endl; return 1; }
        for (int i = 0; i < qfiles; i++) fprintf(f, "%d:%f\n", i + 1, v[i]);
        fclose(f);
        pb200_genomes_free(dev);
        return 0;
    }

    {   // src/parsnp.cpp:3166-3168, 521-537: an output directory that does not exist is created (one level, like `mkdir <outdir>`)
        struct stat sb;
        if (stat(outdir.c_str(), &sb) != 0 && mkdir(outdir.c_str(), 0777) != 0) {
            cerr << "ParSNP:: error creating output directory, exiting.." << endl;
            return 1;
        }
    }
    cerr << "Searching for initial MUM anchors..." << endl;
    pb200_result* res = nullptr;
    rc = pb200_align_resident(dev, &prm, &res);
    const string logpath = outdir + "/parsnpAligner.log";
    if (rc == PB200_ERR_NO_MUMS) {                          // src/parsnp.cpp:3223-3229
        ofstream mfile(logpath.c_str());
        mfile << "NO MUMS FOUND" << endl;
        pb200_result_free(res);                             // (*out is valid on PB200_ERR_NO_MUMS, see parsnp_b200.h)
        pb200_genomes_free(dev);
        return 0;
    }
    if (rc != 0) { cerr << "parsnp_b200_core: " << pb200_last_error() << endl; return 1; }
    const int64_t M = pb200_result_num_mums(res), K = pb200_result_num_clusters(res);
    vector<int64_t> mlen(M), msl(M), mst((size_t)M * n), men((size_t)M * n);
    vector<uint8_t> mfw((size_t)M * n);
    pb200_result_mums(res, mlen.data(), msl.data(), mst.data(), men.data(), mfw.data());
    vector<int32_t> ctype(K);
    vector<int64_t> cnm(K), clen(K), cst((size_t)K * n), cen((size_t)K * n);
    pb200_result_clusters(res, ctype.data(), cnm.data(), clen.data(), cst.data(), cen.data());
    double stats[32]; int ns = pb200_result_stats(res, stats, 32);
    const long anchors_found = ns > 0 ? (long)stats[0] : 0;

    // MUM / LCB coordinates (same record layout as the oracle dump hook, oracle/build_ref.py H1)
    {
        FILE* f = fopen((outdir + "/parsnpAligner.mums").c_str(), "w");
        if (f) {
            fprintf(f, "N %d\n", n);
            for (int64_t i = 0; i < M; i++) {
                fprintf(f, "M %ld %ld", (long)mlen[i], (long)msl[i]);
                for (int k = 0; k < n; k++) fprintf(f, " %ld:%ld:%d", (long)mst[i * n + k], (long)men[i * n + k], (int)mfw[i * n + k]);
                fprintf(f, "\n");
            }
            for (int64_t i = 0; i < K; i++) {
                fprintf(f, "C %d %ld %ld", ctype[i], (long)cnm[i], (long)clen[i]);
                for (int k = 0; k < n; k++) fprintf(f, " %ld:%ld", (long)cst[i * n + k], (long)cen[i * n + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    // XMFA (Aligner::writeOutput, src/parsnp.cpp:505-1079)
    {
        pb200::XmfaInput xi;
        xi.n = n;
        for (auto& g : G) {
            xi.genomes.push_back(&g.text);
            xi.headers.push_back(g.header);
            xi.genome_sizes.push_back((int64_t)g.text.size() - g.padding);
            std::map<int, std::string> p2h;
            p2h[1] = "s1";                                  // src/parsnp.cpp:2921,2947
            for (size_t k = 0; k + 1 < g.contig_ends.size(); k++) p2h[(int)g.contig_ends[k]] = "s" + to_string(k + 2);   // 3120-3122
            xi.pos2hdr.push_back(p2h);
        }
        xi.fasta_names = names;
        xi.c = prm.c;
        xi.doalign = ini.get_i("LCB", "doalign");
        xi.cores = max(1, ini.get_i("LCB", "cores"));
        if (xi.cores > 64) xi.cores = 64;                   // libMUSCLE thread-local storage holds 64 slots (threadstorage.h)
        xi.recombfilter = recomb_filter;
        xi.outdir = outdir;
        xi.ctype = ctype; xi.cstart = cst; xi.cend = cen;
        xi.cmum_off.resize(K + 1);
        const int tot = pb200_result_cluster_mums(res, nullptr, nullptr);
        xi.cmum_idx.resize((size_t)max(tot, 1));
        pb200_result_cluster_mums(res, xi.cmum_off.data(), xi.cmum_idx.data());
        xi.mlen = mlen; xi.mstart = mst; xi.mend = men; xi.mfwd = mfw;
        { ofstream allmums("allmums.out"); }              // the reference creates this (empty) file in the CWD: src/parsnp.cpp:600-601
        if (!pb200::muscle_available()) {
            // the XMFA is the main output of this process: without the inter-MUM aligner there is none, and the Python driver
            // treats exit code 0 as success (parsnp:1146) - so this build configuration fails loudly
            cerr << "parsnp_b200_core: built without libMUSCLE - cannot write parsnpAligner.xmfa (rebuild with PB200_MUSCLE_SRC=<parsnp>/muscle)" << endl;
            return 1;
        }
        else if (!pb200::write_xmfa(xi, outdir + "/parsnpAligner.xmfa")) { cerr << "parsnp_b200_core: XMFA writer failed" << endl; return 1; }
        if (do_unalign) {                                   // src/parsnp.cpp:3283-3287
            const int64_t U = pb200_result_unaligned(res, nullptr, nullptr, nullptr);
            vector<int32_t> ug((size_t)U); vector<int64_t> us((size_t)U), ue((size_t)U);
            pb200_result_unaligned(res, ug.data(), us.data(), ue.data());
            if (!pb200::write_unaligned(xi, ug, us, ue, outdir + "/parsnp.unalign")) { cerr << "parsnp_b200_core: cannot write parsnp.unalign" << endl; return 1; }
        }
    }
    // statistics block of parsnpAligner.log (src/parsnp.cpp:1082-1190), line for line; the elapsed-time lines carry this run's
    // own timings (the reference's have a resolution of one second)
    {
        vector<double> sv(64, 0.0);
        const int nsv = pb200_result_stats(res, sv.data(), 64);
        auto stat = [&](const char* name) -> double {
            const string names = pb200_stats_names();
            size_t pos = 0; int idx = 0;
            while (pos <= names.size()) {
                size_t e = names.find(',', pos);
                if (e == string::npos) e = names.size();
                if (names.compare(pos, e - pos, name) == 0) return idx < nsv ? sv[idx] : 0.0;
                pos = e + 1; idx++;
            }
            return 0.0;
        };
        pb200::LogInput li;
        li.files = files;
        for (auto& g : G) { li.a.push_back(g.a); li.c.push_back(g.c); li.g.push_back(g.g); li.t.push_back(g.t); }
        li.d = prm.d; li.q = prm.q; li.filter = prm.filter;
        int64_t slength = 500000000;
        for (auto& g : G) slength = min<int64_t>(slength, (int64_t)g.text.size());
        li.anchor_size = (float)pb200_minsize(prm.anchors, slength);
        li.anchors_found = anchors_found;
        li.mums_filtered = (long)stat("mums_filtered"); li.clusters_filtered = (long)stat("clusters_filtered");
        li.cnm = cnm;
        li.t_anchor = stat("t_anchor_search") + stat("t_anchor_host"); li.t_coarsen = stat("t_replay"); li.t_lcb = stat("t_lcb");
        li.t_total = stat("t_total");
        pb200::XmfaInput xi;                                // the lists the log reads: names, sizes, clusters, MUMs
        xi.n = n;
        xi.fasta_names = names;
        for (auto& g : G) xi.genome_sizes.push_back((int64_t)g.text.size() - g.padding);
        xi.c = prm.c;
        xi.ctype = ctype;
        xi.cmum_off.resize(K + 1);
        const int tot = pb200_result_cluster_mums(res, nullptr, nullptr);
        xi.cmum_idx.resize((size_t)max(tot, 1));
        pb200_result_cluster_mums(res, xi.cmum_off.data(), xi.cmum_idx.data());
        xi.mlen = mlen; xi.mstart = mst; xi.mfwd = mfw;
        if (!pb200::write_log(xi, li, logpath)) { cerr << "parsnp_b200_core: cannot write " << logpath << endl; return 1; }
    }
    pb200_result_free(res);
    pb200_genomes_free(dev);
    cerr << "Parsnp: Finished core genome alignment" << endl;
    return 0;
}
