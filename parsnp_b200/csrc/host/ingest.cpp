#include "ingest.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>

namespace pb200 {

namespace {
std::string lower(std::string s) { for (auto& ch : s) ch = (char)tolower((unsigned char)ch); return s; }
}

int IniFile::find_section(const std::string& s) const {
    const std::string k = lower(s);
    for (size_t i = 0; i < sections_.size(); ++i) if (lower(sections_[i]) == k) return (int)i;
    return -1;
}
bool IniFile::read(const std::string& path) {
    std::ifstream f(path.c_str());
    if (!f) return false;
    std::string line, section;
    int cur = -1;
    while (std::getline(f, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (line.empty()) continue;
        if (!isprint((unsigned char)line[0])) return false;
        size_t p = line.find_first_of(";#[=");
        if (p == std::string::npos) continue;
        if (line[p] == '[') {
            size_t r = line.find_last_of(']');
            if (r != std::string::npos && r > p) {
                section = line.substr(p + 1, r - p - 1);
                cur = find_section(section);
                if (cur < 0) { sections_.push_back(section); values_.emplace_back(); cur = (int)sections_.size() - 1; }
            }
        } else if (line[p] == '=') {
            if (cur < 0) { sections_.push_back(""); values_.emplace_back(); cur = (int)sections_.size() - 1; }
            const std::string name = line.substr(0, p), value = line.substr(p + 1);
            bool found = false;
            for (auto& kv : values_[cur]) if (lower(kv.first) == lower(name)) { kv.second = value; found = true; break; }
            if (!found) values_[cur].emplace_back(name, value);
        }
    }
    return !sections_.empty();
}
std::string IniFile::get(const std::string& section, const std::string& name, const std::string& def) const {
    int s = find_section(section);
    if (s < 0) return def;
    const std::string k = lower(name);
    for (auto& kv : values_[s]) if (lower(kv.first) == k) return kv.second;
    return def;
}
int IniFile::get_i(const std::string& section, const std::string& name, int def) const {
    char b[64]; snprintf(b, sizeof b, "%d", def);
    return atoi(get(section, name, b).c_str());
}
double IniFile::get_f(const std::string& section, const std::string& name, double def) const {
    char b[64]; snprintf(b, sizeof b, "%f", def);
    return atof(get(section, name, b).c_str());
}
unsigned IniFile::num_values(const std::string& section) const {
    int s = find_section(section);
    return s < 0 ? 0u : (unsigned)values_[s].size();
}

bool ingest_fasta(const std::string& path, bool is_reference, int d, bool reverse, IngestedGenome& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = data.find('\n');
    out.header = data.substr(0, pos == std::string::npos ? data.size() : pos);
    if (out.header.size() > 2499) out.header.resize(2499);        // getline(header, 2500)
    pos = (pos == std::string::npos) ? data.size() : pos + 1;
    std::string& g = out.text;
    g.clear();
    g.reserve(data.size());
    while (pos < data.size()) {
        const char raw = data[pos++];
        switch (toupper((unsigned char)raw)) {
            case 'A': out.a++; g.push_back(reverse ? 'T' : 'A'); break;
            case 'G': out.g++; g.push_back(reverse ? 'C' : 'G'); break;
            case 'C': out.c++; g.push_back(reverse ? 'G' : 'C'); break;
            case 'T': out.t++; g.push_back(reverse ? 'A' : 'T'); break;
            case 'U': out.t++; g.push_back('T'); break;
            case 'X': case 'Y': case 'S': case 'W': case 'K': case 'H': case 'R': case 'M': case 'V': case 'D': case 'B': case '-': case 'N':
                out.n++; g.push_back('N'); break;
            case '>': {
                size_t e = data.find('\n', pos);
                pos = (e == std::string::npos) ? data.size() : e + 1;
                if (!is_reference) { g.append((size_t)d + 10, 'N'); out.n += d + 10; out.padding += d + 10; }
                out.contig_ends.push_back(out.n + out.c + out.t + out.a + out.g);
                break;
            }
            default: break;                                          // '\n', blanks and every unlisted character are skipped
        }
    }
    if (reverse) std::reverse(g.begin(), g.end());
    out.contig_ends.push_back(out.n + out.c + out.t + out.a + out.g);
    return true;
}

}  // namespace pb200
