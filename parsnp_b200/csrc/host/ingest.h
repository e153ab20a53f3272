// ini parsing and FASTA ingest of the parsnp_core process boundary (B1), restated from the reference's main()
// (src/parsnp.cpp:2866-3160) and the semantics of its CIniFile (src/ext/iniFile.cpp:35-97, 274-287).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace pb200 {

// Sections and value names are case-insensitive; "name=value" lines; ';' / '#' start comments (first of ";#[=" decides).
class IniFile {
public:
    bool read(const std::string& path);
    std::string get(const std::string& section, const std::string& name, const std::string& def = "") const;
    int get_i(const std::string& section, const std::string& name, int def = 0) const;        // atoi, like GetValueI
    double get_f(const std::string& section, const std::string& name, double def = 0.0) const; // atof, like GetValueF
    bool get_b(const std::string& section, const std::string& name, bool def = false) const { return get_i(section, name, def ? 1 : 0) != 0; }
    unsigned num_values(const std::string& section) const;
private:
    std::vector<std::string> sections_;
    std::vector<std::vector<std::pair<std::string, std::string>>> values_;
    int find_section(const std::string& s) const;
};

struct IngestedGenome {
    std::string text;          // A,C,G,T,N only
    std::string header;        // first line of the file (whatever it contains)
    int64_t padding = 0;       // N's inserted between contigs (queries only)
    int64_t a = 0, c = 0, g = 0, t = 0, n = 0;
    std::vector<int64_t> contig_ends;
};
// the character loop of src/parsnp.cpp:2999-3133: first line = header; ACGT kept (case folded, complemented when
// `reverse`); IUPAC codes and '-' -> N; U -> T; any other character skipped; every further '>' line ends a contig and,
// for queries, appends d+10 N's; `reverse` finally reverses the text (src/parsnp.cpp:3137-3138).
bool ingest_fasta(const std::string& path, bool is_reference, int d, bool reverse, IngestedGenome& out);

}  // namespace pb200
