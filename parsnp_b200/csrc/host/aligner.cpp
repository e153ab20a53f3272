#include "aligner.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <unordered_set>
#include <atomic>
#include <thread>

namespace pb200 {

namespace {
inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// Runs fn(chunk) for chunk = 0..nchunks-1 on up to `nthreads` threads (dynamic assignment).  Plain std::thread per phase:
// a handful of phases per alignment, and no spinning worker pool that could starve a co-scheduled process.
template <class F>
void parallel_chunks(int nthreads, long nchunks, F&& fn) {
    if (nthreads <= 1 || nchunks <= 1) { for (long c = 0; c < nchunks; ++c) fn(c); return; }
    std::atomic<long> next(0);
    auto worker = [&]() { for (long c; (c = next.fetch_add(1)) < nchunks;) fn(c); };
    std::vector<std::thread> pool;
    const int extra = (int)std::min<long>(nthreads, nchunks) - 1;
    for (int t = 0; t < extra; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
}
inline uint8_t comp_base(uint8_t c) {          // Aligner::reversec (src/parsnp.cpp:1294-1393) on the ingest alphabet
    switch (c) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'C': return 'G';
        case 'G': return 'C';
        default: return 'N';
    }
}
}  // namespace

// ------------------------------------------------------------------ BitRow
void BitRow::init(int64_t nbits) {
    nbits_ = nbits;
    w_.assign((size_t)((nbits + 63) >> 6) + 1, 0ull);
}
void BitRow::set_range(int64_t a, int64_t b) {
    if (a >= b) return;
    int64_t wa = a >> 6, wb = (b - 1) >> 6;
    uint64_t ma = ~0ull << (a & 63), mb = ~0ull >> (63 - ((b - 1) & 63));
    if (wa == wb) { w_[wa] |= (ma & mb); return; }
    w_[wa] |= ma;
    for (int64_t i = wa + 1; i < wb; ++i) w_[i] = ~0ull;
    w_[wb] |= mb;
}
void BitRow::set_range_atomic(int64_t a, int64_t b) {
    if (a >= b) return;
    int64_t wa = a >> 6, wb = (b - 1) >> 6;
    uint64_t ma = ~0ull << (a & 63), mb = ~0ull >> (63 - ((b - 1) & 63));
    if (wa == wb) { __atomic_fetch_or(&w_[wa], ma & mb, __ATOMIC_RELAXED); return; }
    __atomic_fetch_or(&w_[wa], ma, __ATOMIC_RELAXED);
    for (int64_t i = wa + 1; i < wb; ++i) __atomic_store_n(&w_[i], ~0ull, __ATOMIC_RELAXED);
    __atomic_fetch_or(&w_[wb], mb, __ATOMIC_RELAXED);
}
void BitRow::clear_range(int64_t a, int64_t b) {
    if (a >= b) return;
    int64_t wa = a >> 6, wb = (b - 1) >> 6;
    uint64_t ma = ~0ull << (a & 63), mb = ~0ull >> (63 - ((b - 1) & 63));
    if (wa == wb) { w_[wa] &= ~(ma & mb); return; }
    w_[wa] &= ~ma;
    for (int64_t i = wa + 1; i < wb; ++i) w_[i] = 0ull;
    w_[wb] &= ~mb;
}
int64_t BitRow::run_up(int64_t a, int64_t b) const {
    int64_t i = a;
    while (i < b) {
        uint64_t inv = ~(w_[i >> 6] >> (i & 63));          // first zero bit at or after i
        int avail = 64 - (int)(i & 63);
        int z = inv ? __builtin_ctzll(inv) : 64;
        if (z < avail) { i += z; break; }
        i += avail;
    }
    if (i > b) i = b;
    return i - a;
}
int64_t BitRow::run_down(int64_t a, int64_t b) const {
    int64_t i = b - 1;                                       // examine i, i-1, ...
    while (i >= a) {
        int pos = (int)(i & 63);
        uint64_t inv = ~(w_[i >> 6] << (63 - pos));          // bit 63 corresponds to i
        int z = inv ? __builtin_clzll(inv) : 64;
        int avail = pos + 1;
        if (z < avail) { i -= z; break; }
        i -= avail;
    }
    if (i < a - 1) i = a - 1;
    return (b - 1) - i;
}
int64_t BitRow::prev_set(int64_t i) const {
    if (i < 0) return -1;
    int64_t wi = i >> 6;
    uint64_t cur = w_[wi] & (~0ull >> (63 - (i & 63)));
    for (;;) {
        if (cur) return (wi << 6) + 63 - __builtin_clzll(cur);
        if (wi == 0) return -1;
        cur = w_[--wi];
    }
}
int64_t BitRow::next_set(int64_t i, int64_t limit) const {
    if (i >= limit) return limit;
    int64_t wi = i >> 6, wl = (limit - 1) >> 6;
    uint64_t cur = w_[wi] & (~0ull << (i & 63));
    for (;;) {
        if (cur) { int64_t r = (wi << 6) + __builtin_ctzll(cur); return r < limit ? r : limit; }
        if (wi >= wl) return limit;
        cur = w_[++wi];
    }
}

// ------------------------------------------------------------------ Aligner basics
Aligner::Aligner(int n, const uint8_t* const* seq, const int64_t* len, const AlignParams& prm, SearchBackend* be)
    : n_(n), prm_(prm), be_(be), anchor_expr_(prm.anchors), mum_expr_(prm.mums) {
    if (!be) throw std::runtime_error("parsnp_b200: no search backend (the CUDA engine is required)");
    seq_.assign(seq, seq + n);
    len_.assign(len, len + n);
    rp_.n = n;
    truth_.layout.resize(n);
    for (int i = 0; i < n; ++i) {                      // src/parsnp.cpp:3181-3186
        truth_.layout[i].init(len_[i] + 1);
        truth_.layout[i].set_range(len_[i], len_[i] + 1);
    }
    be_->set_genomes(n, seq, len);
}

int RegionPool::add(const int64_t* start, const int64_t* end) {
    int id = (int)slen.size();
    coord.insert(coord.end(), start, start + n);
    coord.insert(coord.end(), end, end + n);
    int64_t sl = 500000000;                             // TRegion ctor, src/LCR.cpp:16-37
    for (int i = 0; i < n; ++i) sl = std::min(sl, end[i] - start[i]);
    slen.push_back(sl);
    return id;
}
bool Aligner::region_equal(int a, int b) const {        // operator==, src/LCR.cpp:48-58
    return std::memcmp(rstart(a), rstart(b), sizeof(int64_t) * 2 * n_) == 0;
}
uint64_t Aligner::coords_hash(const int64_t* p, int count) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < count; ++i) {
        h ^= (uint64_t)p[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 29;
    }
    return h;
}
void Aligner::CoordIndex::insert(uint64_t hash, int value) {
    if ((count + 1) * 2 > h.size()) {                       // grow / rehash at 50 % load
        std::vector<uint64_t> oh;
        std::vector<int> ov;
        oh.swap(h);
        ov.swap(v);
        const size_t nsz = oh.empty() ? 1024 : oh.size() * 2;
        h.assign(nsz, 0);
        v.assign(nsz, -1);
        count = 0;
        for (size_t i = 0; i < oh.size(); ++i) if (ov[i] >= 0) insert(oh[i], ov[i]);
    }
    const size_t mask = h.size() - 1;
    size_t i = (size_t)hash & mask;
    while (v[i] >= 0) i = (i + 1) & mask;
    h[i] = hash;
    v[i] = value;
    ++count;
}
int Aligner::cache_lookup_coords(const int64_t* coords) const {
    const size_t bytes = sizeof(int64_t) * 2 * n_;
    return cache_map_.find(coords_hash(coords, 2 * n_),
                           [&](int e) { return std::memcmp(rstart(cache_entries_[e].region), coords, bytes) == 0; });
}

int Aligner::minsize_cached(bool anchors, int64_t slength) {
    std::unordered_map<int64_t, int>& c = minsize_cache_[anchors ? 1 : 0];
    auto it = c.find(slength);
    if (it != c.end()) return it->second;
    int v = anchors ? anchor_expr_(slength) : mum_expr_(slength);
    c.emplace(slength, v);
    return v;
}

// ------------------------------------------------------------------ batched search (setMums1 up to the emission loop)
void Aligner::search_regions(const std::vector<int>& regs, bool anchors) {
    if (regs.empty()) return;
    std::vector<WindowTask> tasks;
    std::vector<int64_t> coords;
    std::vector<int> first_task(regs.size() + 1, 0);
    const int nq = n_ - 1;
    for (size_t ri = 0; ri < regs.size(); ++ri) {
        const int r = regs[ri];
        const int64_t* rs = rstart(r);
        const int64_t* re = rend(r);
        first_task[ri] = (int)tasks.size();
        const int minsize = minsize_cached(anchors, rp_.slen[r]);
        const int64_t coff = (int64_t)coords.size();
        for (int j = 1; j < n_; ++j) coords.push_back(rs[j]);
        for (int j = 1; j < n_; ++j) coords.push_back(re[j] - rs[j]);
        // reference window loop, src/parsnp.cpp:1519-1547 (ssize == size_t arithmetic)
        const int64_t L0 = re[0] - rs[0];
        uint64_t p = ((int64_t)prm_.p > L0) ? (uint64_t)L0 : (uint64_t)prm_.p;
        uint64_t partpos = 0;
        while (partpos < (uint64_t)L0) {
            if (partpos + p > (uint64_t)L0) {
                p = (uint64_t)L0 - partpos;
                if (p < 50) { p = 50 + p; partpos = partpos - 50; }
            }
            WindowTask t;
            t.ref_start = rs[0] + (int64_t)partpos;
            t.ref_len = (int64_t)p;
            t.coord_off = coff;
            t.minsize = minsize;
            t.pad = 0;
            if (t.ref_start < 0 || t.ref_start + t.ref_len > len_[0] || t.ref_len <= 0)
                throw std::runtime_error("parsnp_b200: reference window outside genome (p < 50?)");
            tasks.push_back(t);
            partpos += p;
        }
    }
    first_task[regs.size()] = (int)tasks.size();
    CandBatch cb;
    cb.nq = nq;
    if (!tasks.empty()) be_->search(tasks.data(), (int)tasks.size(), coords.data(), cb);
    else cb.off.assign(1, 0);
    stats_.windows_searched += (int64_t)tasks.size();
    stats_.regions_searched += (int64_t)regs.size();
    // append to the cache stores
    const int64_t base = (int64_t)ck_.size();
    ck_.insert(ck_.end(), cb.k.begin(), cb.k.end());
    clon_.insert(clon_.end(), cb.lon.begin(), cb.lon.end());
    csp_.insert(csp_.end(), cb.sp.begin(), cb.sp.end());
    cfwd_.insert(cfwd_.end(), cb.fwd.begin(), cb.fwd.end());
    stats_.candidates += (int64_t)cb.k.size();
    for (size_t ri = 0; ri < regs.size(); ++ri) {
        CacheEntry e;
        e.region = regs[ri];
        e.first_win = (int64_t)wins_.size();
        e.nwin = first_task[ri + 1] - first_task[ri];
        for (int t = first_task[ri]; t < first_task[ri + 1]; ++t) {
            WinRec w;
            w.ref_start = tasks[t].ref_start;
            w.ref_len = tasks[t].ref_len;
            w.cand_off = base + cb.off[t];
            w.ncand = (int32_t)(cb.off[t + 1] - cb.off[t]);
            w.minsize = tasks[t].minsize;
            wins_.push_back(w);
        }
        cache_map_.insert(coords_hash(rstart(regs[ri]), 2 * n_), (int)cache_entries_.size());
        cache_entries_.push_back(e);
    }
}

// ------------------------------------------------------------------ setMums1 loop D (src/parsnp.cpp:1713-1842)
void Aligner::accept_candidates(const int64_t* rs, const int64_t* re, int64_t rsl, int cache_idx, std::vector<BitRow>& layout, MumPool& mp,
                                std::vector<int>& found, bool atomic, bool trace) {
    const CacheEntry& ce = cache_entries_[cache_idx];
    const int nq = n_ - 1;
    int64_t st_buf[64];
    uint8_t fw_buf[64];
    std::vector<int64_t> st_vec;
    std::vector<uint8_t> fw_vec;
    int64_t* st = st_buf;
    uint8_t* fw = fw_buf;
    if (n_ > 64) { st_vec.resize(n_); fw_vec.resize(n_); st = st_vec.data(); fw = fw_vec.data(); }
    for (int wi = 0; wi < ce.nwin; ++wi) {
        const WinRec& win = wins_[ce.first_win + wi];
        if (trace) trace_.emplace_back(win.ref_start, win.ref_len);
        for (int32_t c = 0; c < win.ncand; ++c) {
            const int64_t ci = win.cand_off + c;
            const int64_t LON = clon_[ci];
            bool bad = false;
            // Mum.DSP is 1-based (src/parsnp.cpp:1671,1681); range pre-check in unsigned arithmetic (1723)
            uint64_t dsp0 = (uint64_t)((int64_t)ck_[ci] + 1 + win.ref_start);
            if ((uint64_t)(dsp0 - (uint64_t)rs[0]) > (uint64_t)(uint32_t)(re[0] - rs[0])) bad = true;
            st[0] = (int64_t)dsp0 - 1;
            fw[0] = 1;
            for (int j = 1; j < n_; ++j) {
                uint64_t dsp = (uint64_t)((int64_t)csp_[ci * nq + (j - 1)] + 1 + rs[j]);
                if ((uint64_t)(dsp - (uint64_t)rs[j]) > (uint64_t)(uint32_t)(re[j] - rs[j])) bad = true;
                st[j] = (int64_t)dsp - 1;
                fw[j] = cfwd_[ci * nq + (j - 1)];
            }
            if (bad) continue;
            // TMum ctor (src/TMum.cpp:13-72): reverse-strand start uses the WHOLE genome length; `ok` = last genome
            bool ok = true, any_fail = false;
            for (int j = 0; j < n_; ++j)
                if (!fw[j]) st[j] = len_[j] - (st[j] + LON);
            for (int j = 0; j < n_; ++j) {
                if (st[j] + LON > len_[j] || st[j] < 0) { ok = false; any_fail = true; }
                else ok = true;
            }
            if (any_fail) ok = false;      // (a middle-genome failure makes the reference throw; unreachable)
            if (!ok || LON < 5) continue;
            // trim (src/parsnp.cpp:1399-1477): every trim shifts ALL genomes, strand ignored
            int64_t length = LON;
            for (int j = 0; j < n_; ++j) {
                int64_t t1 = layout[j].run_up(st[j], st[j] + length);
                if (t1) { for (int i = 0; i < n_; ++i) st[i] += t1; length -= t1; }
                int64_t t2 = layout[j].run_down(st[j], st[j] + length);
                length -= t2;
                if (length <= 0) break;          // nothing left: the remaining genomes' loops would not execute (src/parsnp.cpp:1409,1443)
            }
            if (length < 2 || n_ <= 1) continue;
            // reverse-strand genomes are verified against the reference substring (src/parsnp.cpp:1800-1825)
            bool badmum = false;
            for (int k = 0; k < n_ && !badmum; ++k) {
                if (fw[k]) continue;
                const uint8_t* g0 = seq_[0] + st[0];
                const uint8_t* gk = seq_[k] + st[k];
                for (int64_t t = 0; t < length; ++t)
                    if (comp_base(gk[length - 1 - t]) != g0[t]) { badmum = true; break; }
            }
            if (badmum) continue;
            for (int k = 0; k < n_; ++k) {
                if (atomic) layout[k].set_range_atomic(st[k], st[k] + length);
                else layout[k].set_range(st[k], st[k] + length);
            }
            MumRec m;
            m.length = length;
            m.slength = rsl;
            m.off = (int64_t)mp.start.size();
            m.alive = true;
            mp.start.insert(mp.start.end(), st, st + n_);
            mp.fwd.insert(mp.fwd.end(), fw, fw + n_);
            found.push_back((int)mp.mums.size());
            mp.mums.push_back(m);
        }
    }
}

// determineRegion (src/parsnp.cpp:1199-1290) into tmp coordinate buffers; returns slength
static int64_t det_region(const std::vector<BitRow>& layout, const std::vector<int64_t>& len, int n,
                          const int64_t* mstart, int64_t mlen, bool left, int64_t* S, int64_t* E) {
    int64_t sl = 500000000;
    for (int i = 0; i < n; ++i) {
        if (left) {
            int64_t cp = layout[i].prev_set(mstart[i] - 1);
            if (cp < 0) cp = 0;
            S[i] = cp + 1;
            E[i] = mstart[i] - 1;
        } else {
            int64_t en = mstart[i] + mlen;
            int64_t cp = en + 1;
            if (cp < len[i]) cp = layout[i].next_set(cp, len[i]);
            S[i] = en + 1;
            E[i] = cp - 1;
        }
        sl = std::min(sl, E[i] - S[i]);
    }
    return sl;
}

// ------------------------------------------------------------------ anchors (src/parsnp.cpp:2121-2174)
void Aligner::set_initial_clusters() {
    double t0 = now_s();
    std::vector<int64_t> S(n_, 0), E(len_);
    int whole = rp_.add(S.data(), E.data());
    search_regions(std::vector<int>(1, whole), true);
    double t1 = now_s();
    stats_.t_anchor_search = t1 - t0;
    std::vector<int> found;
    accept_candidates(rstart(whole), rend(whole), rp_.slen[whole], cache_lookup(whole), truth_.layout, mp_, found, false, trace_on_);
    all_mums_ = found;
    stats_.anchors = (int64_t)found.size();
    // determineRegion of every anchor: mumlayout is final here (all anchors placed), so the scans are independent and run
    // in parallel over blocks of anchors; the push rules (src/parsnp.cpp:2153-2172) are then applied in order
    const size_t B = 16384;
    std::vector<int64_t> buf(4 * std::min(B, found.size() + 1) * (size_t)n_);
    std::vector<int64_t> sl(2 * B);
    std::vector<int64_t> prevS(n_), prevE(n_);
    bool have_r = false;
    for (size_t b0 = 0; b0 < found.size(); b0 += B) {
        const size_t bn = std::min(B, found.size() - b0);
        const long per = 128;
        parallel_chunks(bn > 512 ? threads_ : 1, ((long)bn + per - 1) / per, [&](long c) {
            for (long x = c * per; x < std::min<long>((long)bn, (c + 1) * per); ++x) {
                const MumRec& m = mums_[found[b0 + x]];
                const int64_t* ms = &mum_start_[m.off];
                int64_t* p = &buf[(size_t)x * 4 * n_];
                sl[2 * x] = det_region(truth_.layout, len_, n_, ms, m.length, true, p, p + n_);
                sl[2 * x + 1] = det_region(truth_.layout, len_, n_, ms, m.length, false, p + 2 * n_, p + 3 * n_);
            }
        });
        for (size_t x = 0; x < bn; ++x) {
            const size_t i = b0 + x;
            const int64_t* lS = &buf[x * 4 * n_]; const int64_t* lE = lS + n_; const int64_t* rS = lE + n_; const int64_t* rE = rS + n_;
            bool l_eq_r = have_r && std::memcmp(lS, prevS.data(), sizeof(int64_t) * n_) == 0 && std::memcmp(lE, prevE.data(), sizeof(int64_t) * n_) == 0;
            if (sl[2 * x] > prm_.q && (i == 0 || !l_eq_r)) initial_regions_.push_back(rp_.add(lS, lE));
            have_r = true;
            bool r_eq_l = std::memcmp(lS, rS, sizeof(int64_t) * n_) == 0 && std::memcmp(lE, rE, sizeof(int64_t) * n_) == 0;
            if (sl[2 * x + 1] > prm_.q && !r_eq_l) initial_regions_.push_back(rp_.add(rS, rE));
            std::memcpy(prevS.data(), rS, sizeof(int64_t) * n_);
            std::memcpy(prevE.data(), rE, sizeof(int64_t) * n_);
        }
    }
    stats_.t_anchor_host = now_s() - t1;
}

// ------------------------------------------------------------------ speculative level-synchronous discovery
// One level over frontier[a,b) (sorted by start[0]): accept on the scratch layout, collect the children's coordinates.
// Only a predictor of which regions the exact replay will ask for - races between threads merely cost cache misses.
void Aligner::speculate_range(const std::vector<int>& frontier, size_t a, size_t b, std::vector<BitRow>& layout, MumPool& mp, RegionPool& out,
                              bool atomic) {
    std::vector<int> found;
    std::vector<int64_t> lS(n_), lE(n_), rS(n_), rE(n_);
    int prev = -1;
    for (size_t x = a; x < b; ++x) {
        const int r = frontier[x];
        if (prev >= 0 && region_equal(prev, r)) continue;
        prev = r;
        found.clear();
        const int ci = cache_lookup(r);
        if (ci < 0) continue;
        accept_candidates(rstart(r), rend(r), rp_.slen[r], ci, layout, mp, found, atomic, false);
        int64_t lsl = 0;
        for (size_t i = 0; i < found.size(); ++i) {
            const MumRec& m = mp.mums[found[i]];
            const int64_t* ms = &mp.start[m.off];
            if (i == 0) lsl = det_region(layout, len_, n_, ms, m.length, true, lS.data(), lE.data());
            int64_t rsl = det_region(layout, len_, n_, ms, m.length, false, rS.data(), rE.data());
            if (lsl > prm_.q) out.add(lS.data(), lE.data());
            if (rsl > prm_.q) out.add(rS.data(), rE.data());
            if (i + 1 < found.size()) {
                const MumRec& m2 = mp.mums[found[i + 1]];
                lsl = det_region(layout, len_, n_, &mp.start[m2.off], m2.length, true, lS.data(), lE.data());
            }
        }
    }
}

void Aligner::speculate(const std::vector<int>& initial, const World& truth) {
    World spec = truth;                                    // scratch copy of mumlayout
    std::vector<int> frontier = initial, need;
    while (!frontier.empty()) {
        double t0 = now_s();
        need.clear();
        {
            std::unordered_multimap<uint64_t, int> seen;
            for (int r : frontier) {
                if (cache_lookup(r) >= 0) continue;
                uint64_t h = coords_hash(rstart(r), 2 * n_);
                bool dup = false;
                auto range = seen.equal_range(h);
                for (auto it = range.first; it != range.second; ++it) if (region_equal(it->second, r)) { dup = true; break; }
                if (dup) continue;
                seen.emplace(h, r);
                need.push_back(r);
            }
        }
        search_regions(need, false);
        stats_.spec_regions += (int64_t)need.size();
        stats_.spec_levels++;
        double t1 = now_s();
        stats_.t_spec_search += t1 - t0;
        std::stable_sort(frontier.begin(), frontier.end(), [&](int a, int b) { return rstart(a)[0] < rstart(b)[0]; });
        // chunks of the frontier are processed concurrently on the shared scratch layout
        const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads_, frontier.size() / 256 + 1));
        const size_t nchunks = T > 1 ? (size_t)T * 4 : 1;
        std::vector<RegionPool> outs(nchunks);
        for (auto& o : outs) o.n = n_;
        parallel_chunks(T, (long)nchunks, [&](long c) {
            MumPool mp;
            const size_t a = frontier.size() * (size_t)c / nchunks, b = frontier.size() * (size_t)(c + 1) / nchunks;
            speculate_range(frontier, a, b, spec.layout, mp, outs[c], T > 1);
        });
        std::vector<int> next;
        for (auto& o : outs)
            for (int i = 0; i < o.size(); ++i) next.push_back(rp_.add(o.start(i), o.end(i)));
        frontier.swap(next);
        stats_.t_spec_host += now_s() - t1;
    }
}

// ------------------------------------------------------------------ doWork (src/parsnp.cpp:173-317), exact order
namespace {
struct QE { int64_t s0; int id; };
inline bool operator<(const QE& a, const QE& b) { return a.s0 < b.s0; }   // operator<, src/LCR.cpp:42
}

void Aligner::process_queue_exact(const std::vector<int>& initial, RegionPool& rp, std::vector<BitRow>& layout, MumPool& mp,
                                  std::vector<int>& out_mums) {
    // exact emulation of `vector<TRegion> regions`: slow mode keeps the vector itself; fast mode is valid while
    // all start[0] keys are distinct (then every correct sort yields the same sequence).
    auto req = [&](int a, int b) { return std::memcmp(rp.start(a), rp.start(b), sizeof(int64_t) * 2 * n_) == 0; };
    std::vector<QE> vec;
    for (int r : initial) vec.push_back(QE{rp.start(r)[0], r});
    std::map<int64_t, int> fast;
    bool fast_mode = false;
    std::vector<int> found, children;
    std::vector<int64_t> lS(n_), lE(n_), rS(n_), rE(n_);
    while (fast_mode ? !fast.empty() : !vec.empty()) {
        int cur;
        if (fast_mode) { cur = fast.begin()->second; fast.erase(fast.begin()); }
        else { cur = vec.front().id; vec.erase(vec.begin()); }
        int ci = cache_lookup_coords(rp.start(cur));
        if (ci < 0) {
            double ts = now_s();
            search_regions(std::vector<int>(1, cur), false);          // a region the speculation did not predict
            stats_.t_replay_search += now_s() - ts;
            stats_.replay_misses++;
            ci = cache_lookup_coords(rp.start(cur));
        }
        found.clear();
        accept_candidates(rp.start(cur), rp.end(cur), rp.slen[cur], ci, layout, mp, found, false, trace_on_);
        children.clear();
        int64_t lsl = 0;
        for (size_t i = 0; i < found.size(); ++i) {
            const MumRec& m = mp.mums[found[i]];
            const int64_t* ms = &mp.start[m.off];
            if (i == 0) lsl = det_region(layout, len_, n_, ms, m.length, true, lS.data(), lE.data());
            int64_t rsl = det_region(layout, len_, n_, ms, m.length, false, rS.data(), rE.data());
            if (lsl > prm_.q) children.push_back(rp.add(lS.data(), lE.data()));
            if (rsl > prm_.q) children.push_back(rp.add(rS.data(), rE.data()));
            if (i + 1 < found.size()) {
                const MumRec& m2 = mp.mums[found[i + 1]];
                lsl = det_region(layout, len_, n_, &mp.start[m2.off], m2.length, true, lS.data(), lE.data());
            }
            out_mums.push_back(found[i]);
        }
        // sort + drop adjacent duplicates (src/parsnp.cpp:291-306)
        if (fast_mode) {
            bool distinct_tie = false;
            for (size_t a = 0; a < children.size() && !distinct_tie; ++a) {
                auto it = fast.find(rp.start(children[a])[0]);
                if (it != fast.end() && !req(it->second, children[a])) distinct_tie = true;
                for (size_t b = 0; b < a && !distinct_tie; ++b)
                    if (rp.start(children[a])[0] == rp.start(children[b])[0] && !req(children[a], children[b])) distinct_tie = true;
            }
            if (!distinct_tie) {
                for (int ch : children) fast.emplace(rp.start(ch)[0], ch);   // identical duplicates collapse
                continue;
            }
            vec.clear();
            for (auto& kv : fast) vec.push_back(QE{kv.first, kv.second});
            fast.clear();
            fast_mode = false;
        }
        stats_.slow_queue_iters++;
        for (int ch : children) vec.push_back(QE{rp.start(ch)[0], ch});
        if (!vec.empty()) std::sort(vec.begin(), vec.end());
        {
            size_t rsize = vec.size();
            if (rsize) {
                for (size_t m = 0; m + 1 < rsize;) {
                    if (req(vec[m].id, vec[m + 1].id)) { vec.erase(vec.begin() + m); rsize -= 1; }
                    else ++m;
                }
            }
        }
        bool strict = true;
        for (size_t m = 0; m + 1 < vec.size(); ++m) if (!(vec[m].s0 < vec[m + 1].s0)) { strict = false; break; }
        if (strict) {
            fast.clear();
            for (auto& e : vec) fast.emplace_hint(fast.end(), e.s0, e.id);
            vec.clear();
            fast_mode = true;
        }
    }
}

void Aligner::do_work_exact() {
    double t0 = now_s();
    std::vector<int> out;
    process_queue_exact(initial_regions_, rp_, truth_.layout, mp_, out);
    all_mums_.insert(all_mums_.end(), out.begin(), out.end());
    stats_.t_replay += now_s() - t0;
}

// sort(this->mums) by start[0] (operator<, src/TMum.cpp:151); starts are distinct (accepted MUMs are disjoint on the
// reference), so the order is unique.  Sorts compact (start0,id) pairs and skips the work when already sorted.
void Aligner::sort_final_mums() {
    const size_t M = final_mums_.size();
    bool sorted = true;
    int64_t prev = INT64_MIN;
    for (size_t i = 0; i < M; ++i) {
        int64_t s0 = mum_start_[mums_[final_mums_[i]].off];
        if (s0 < prev) { sorted = false; break; }
        prev = s0;
    }
    if (sorted) return;
    // LSD byte radix sort of (start0, id) pairs: keys are distinct, so the result is the unique ascending order
    std::vector<std::pair<int64_t, int>> kv(M), tmp(M);
    int64_t maxkey = 0;
    for (size_t i = 0; i < M; ++i) {
        kv[i] = std::make_pair(mum_start_[mums_[final_mums_[i]].off], final_mums_[i]);
        maxkey = std::max(maxkey, kv[i].first);
    }
    for (int shift = 0; shift < 64 && (maxkey >> shift) != 0; shift += 8) {
        size_t cnt[257] = {0};
        for (size_t i = 0; i < M; ++i) cnt[((uint64_t)kv[i].first >> shift & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < M; ++i) tmp[cnt[(uint64_t)kv[i].first >> shift & 0xff]++] = kv[i];
        kv.swap(tmp);
    }
    for (size_t i = 0; i < M; ++i) final_mums_[i] = kv[i].second;
}

// ------------------------------------------------------------------ filterRandom1 (src/parsnp.cpp:327-425)
void Aligner::filter_random1() {
    // with the ini's filter=1 (`rvalue` = 1) no MUM has length <= 1, so only the sort has an effect;
    // larger values are restated literally below.
    sort_final_mums();
    const int rvalue = prm_.random;
    size_t numums = final_mums_.size();
    if (numums == 0) return;
    for (size_t ms = 0; ms + 1 < numums; ++ms) {
        const MumRec& mt = mums_[final_mums_[ms]];
        if (mt.length > rvalue) continue;
        const MumRec& nt = mums_[final_mums_[ms + 1]];
        const int64_t* mts = &mum_start_[mt.off];
        const int64_t* nts = &mum_start_[nt.off];
        bool adjacent = true;
        for (int k = 0; k < n_ && adjacent; ++k) {
            int64_t mte = mts[k] + mt.length;
            int64_t gap = std::llabs(nts[k]) - std::llabs(mte);
            if (gap < 0 || gap > 5000) { adjacent = false; break; }
            for (int64_t m = mte + 1; m < nts[k]; ++m) if (truth_.layout[k].get(m)) { adjacent = false; break; }
            if (ms != 0) {
                const MumRec& pm = mums_[final_mums_[ms - 1]];
                int64_t pe = mum_start_[pm.off + k] + pm.length;
                int64_t g2 = std::llabs(mts[k]) - std::llabs(pe);
                if (g2 < 0 || g2 > 5000) { adjacent = false; break; }
                for (int64_t m = pe + 1; m < mts[k]; ++m) if (truth_.layout[k].get(m)) { adjacent = false; break; }
            }
        }
        if (!adjacent) {
            for (int k = 0; k < n_; ++k) truth_.layout[k].clear_range(mts[k], mts[k] + mt.length);
            mums_[final_mums_[ms]].alive = false;
            final_mums_.erase(final_mums_.begin() + ms);
            ms -= 1;               // size_t wrap + ++ms == stay, like the reference's ulong msize
            numums -= 1;
        }
    }
}

// ------------------------------------------------------------------ setFinalClusters (src/parsnp.cpp:2563-2719)
void Aligner::set_final_clusters(std::vector<ClusterRec>& out) {
    out.clear();
    sort_final_mums();
    const int64_t M = (int64_t)final_mums_.size();
    if (M == 0) return;
    // contiguous copies in sorted order (the pools are in discovery order): the chaining below streams through them
    std::vector<int64_t> ss((size_t)M * n_), sl((size_t)M);
    std::vector<uint8_t> sf((size_t)M * n_);
    const long per = 4096;
    parallel_chunks(M > 32768 ? threads_ : 1, ((long)M + per - 1) / per, [&](long c) {
        for (long i = c * per; i < std::min<long>((long)M, (c + 1) * per); ++i) {
            const MumRec& m = mums_[final_mums_[i]];
            sl[i] = m.length;
            std::memcpy(&ss[(size_t)i * n_], &mum_start_[m.off], sizeof(int64_t) * n_);
            std::memcpy(&sf[(size_t)i * n_], &mum_fwd_[m.off], (size_t)n_);
        }
    });
    auto S = [&](int64_t i, int k) { return ss[(size_t)i * n_ + k]; };
    auto Len = [&](int64_t i) { return sl[i]; };
    auto F = [&](int64_t i, int k) { return (int)sf[(size_t)i * n_ + k]; };
    auto new_cluster = [&](int64_t i) {
        ClusterRec c;
        c.type = 1;
        c.length = Len(i);
        c.mums.push_back((int)i);
        c.start.resize(n_);
        c.end.resize(n_);
        for (int k = 0; k < n_; ++k) { c.start[k] = S(i, k); c.end[k] = S(i, k) + Len(i); }
        return c;
    };
    auto add_mum = [&](ClusterRec& c, int64_t i) {          // Cluster::addMum, src/LCB.cpp:31-37
        for (int k = 0; k < n_; ++k) c.end[k] = S(i, k) + Len(i);
        c.length += Len(i);
        c.mums.push_back((int)i);
    };
    ClusterRec cluster = new_cluster(0);
    bool addmum = true;
    const float dd = prm_.diag_diff;
    for (int64_t nt = 1; nt < M; ++nt) {
        if (Len(nt) < prm_.random) { addmum = true; continue; }
        if (!addmum) cluster = new_cluster(nt - 1);
        addmum = true;
        float max_length_region = 0;
        float min_length_region = (float)(prm_.d + 10);
        const int64_t back = cluster.mums.back(), front = cluster.mums.front();
        for (int k = 0; k < n_; ++k) {
            const int f = F(nt, k);
            const int64_t gap = S(nt, k) - cluster.end[k];
            const int64_t rgap = S(back, k) - (S(nt, k) + Len(nt));
            if (f && (float)gap > max_length_region) max_length_region = (float)gap;
            else if (!f && (float)rgap > max_length_region) max_length_region = (float)gap;   // sic (src/parsnp.cpp:2608-2611)
            if (f && (float)gap < min_length_region) min_length_region = (float)gap;
            else if (!f && (float)rgap < min_length_region) min_length_region = (float)rgap;
            if ((f != F(back, k)) || (f != F(front, k))) addmum = false;
            else if (f && gap < 0) addmum = false;
            else if (!f && gap >= 0) addmum = false;
            else if (f && gap > prm_.d) addmum = false;
            else if (!f && rgap > prm_.d) addmum = false;
            if (!addmum) break;
        }
        if (addmum) {
            if (min_length_region == 0) min_length_region = 1;
            if (max_length_region == 0) max_length_region = 1;
            if (dd > 1.0) {
                if (max_length_region - min_length_region < dd) add_mum(cluster, nt);
                // else: the MUM is silently skipped and the cluster stays open (src/parsnp.cpp:2684-2691)
            } else if (min_length_region / max_length_region >= 1.0 - dd) {
                add_mum(cluster, nt);
            } else {
                addmum = false;
                out.push_back(cluster);
            }
        } else {
            out.push_back(cluster);
        }
    }
    if (!addmum) cluster = new_cluster(M - 1);
    out.push_back(cluster);
}

// ------------------------------------------------------------------ filterRandomClustersSimple1 (src/parsnp.cpp:433-497)
void Aligner::filter_clusters_simple(std::vector<ClusterRec>& cl) {
    std::sort(cl.begin(), cl.end(), [](const ClusterRec& a, const ClusterRec& b) { return a.start[0] < b.start[0]; });
    size_t num = cl.size();
    if (num == 0) return;
    for (size_t cs = 0; cs + 1 < num;) {
        if (cl[cs].length <= prm_.c) {
            for (int mi : cl[cs].mums) {
                MumRec& m = mums_[final_mums_[mi]];
                for (int k = 0; k < n_; ++k) truth_.layout[k].clear_range(mum_start_[m.off + k], mum_start_[m.off + k] + m.length);
                m.alive = false;
            }
            cl.erase(cl.begin() + cs);
            num -= 1;
        } else {
            ++cs;
        }
    }
    std::vector<int> keep;
    for (int id : final_mums_) if (mums_[id].alive) keep.push_back(id);
    final_mums_.swap(keep);
}

// ------------------------------------------------------------------ setInterClusterRegions (src/parsnp.cpp:2389-2460)
void Aligner::set_inter_cluster_regions(std::vector<ClusterRec>& cl) {
    std::sort(cl.begin(), cl.end(), [](const ClusterRec& a, const ClusterRec& b) { return a.start[0] < b.start[0]; });
    std::vector<ClusterRec> inter;
    for (size_t ct = 0; ct + 1 < cl.size(); ++ct) {
        const ClusterRec& c = cl[ct];
        const ClusterRec& nx = cl[ct + 1];
        bool add = true;
        ClusterRec a;
        a.type = 0;
        a.length = 2;
        for (int g = 0; g < n_; ++g) {
            if (nx.start[g] - c.end[g] <= 0) { add = false; break; }
            const int64_t stop = len_[g];
            a.start.push_back(c.end[g]);
            int64_t m = truth_.layout[g].next_set(c.end[g] + 1, stop + 1);   // sentinel bit at `stop`
            if (m > stop) m = stop;
            a.end.push_back(m);                  // emum: start = m-1, end = m (src/parsnp.cpp:2424,2439-2442)
        }
        if (!add) continue;
        for (int g = 0; g < n_; ++g) if (a.end[g] - a.start[g] < 5) { add = false; break; }
        if (add) inter.push_back(a);
    }
    cl.insert(cl.begin(), inter.begin(), inter.end());
}

// ------------------------------------------------------------------ main sequence (src/parsnp.cpp:3187-3273)
bool Aligner::run() {
    double t0 = now_s();
    set_initial_clusters();
    if (!prm_.anchors_only) {
        if (speculate_) speculate(initial_regions_, truth_);
        stats_.host_threads = threads_;
        do_work_exact();
    }
    if (all_mums_.empty()) { stats_.t_total = now_s() - t0; return false; }
    double t1 = now_s();
    final_mums_ = all_mums_;
    if (prm_.random) filter_random1();
    set_final_clusters(clusters_);
    filter_clusters_simple(clusters_);
    set_final_clusters(clusters_);
    set_inter_cluster_regions(clusters_);
    stats_.t_lcb = now_s() - t1;
    stats_.t_total = now_s() - t0;
    return true;
}

}  // namespace pb200
