#include "aligner.h"
#include "parallel.h"
#include "accept_impl.h"
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
#include <map>
#include <memory>
#include <mutex>

namespace pb200 {
int default_host_threads();                                  // result.cpp

namespace {
inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

// ------------------------------------------------------------------ BitRow
void BitRow::init(int64_t nbits) {
    nbits_ = nbits;
    w_.assign((size_t)((nbits + 63) >> 6) + 1, 0ull);
}
void BitRow::set_range_slow(int64_t a, int64_t b) {
    int64_t wa = a >> 6, wb = (b - 1) >> 6;
    uint64_t ma = ~0ull << (a & 63), mb = ~0ull >> (63 - ((b - 1) & 63));
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
void BitRow::set_range_owned(int64_t a, int64_t b, int64_t wlo, int64_t whi) {
    if (a >= b) return;
    const int64_t wa = a >> 6, wb = (b - 1) >> 6;
    for (int64_t wi = wa; wi <= wb; ++wi) {
        uint64_t m = ~0ull;
        if (wi == wa) m &= ~0ull << (a & 63);
        if (wi == wb) m &= ~0ull >> (63 - ((b - 1) & 63));
        if (wi > wlo && wi < whi) __atomic_store_n(&w_[(size_t)wi], __atomic_load_n(&w_[(size_t)wi], __ATOMIC_RELAXED) | m, __ATOMIC_RELAXED);
        else __atomic_fetch_or(&w_[(size_t)wi], m, __ATOMIC_RELAXED);
    }
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
int64_t BitRow::run_up_slow(int64_t a, int64_t b) const {
    int64_t i = a;
    while (i < b) {
        uint64_t inv = ~(word(i >> 6) >> (i & 63));          // first zero bit at or after i
        int avail = 64 - (int)(i & 63);
        int z = inv ? __builtin_ctzll(inv) : 64;
        if (z < avail) { i += z; break; }
        i += avail;
    }
    if (i > b) i = b;
    return i - a;
}
int64_t BitRow::run_down_slow(int64_t a, int64_t b) const {
    int64_t i = b - 1;                                       // examine i, i-1, ...
    while (i >= a) {
        int pos = (int)(i & 63);
        uint64_t inv = ~(word(i >> 6) << (63 - pos));          // bit 63 corresponds to i
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
    uint64_t cur = word(wi) & (~0ull >> (63 - (i & 63)));
    for (;;) {
        if (cur) return (wi << 6) + 63 - __builtin_clzll(cur);
        if (wi == 0) return -1;
        cur = word(--wi);
    }
}
int64_t BitRow::prev_set_from(int64_t i, int64_t lo) const {
    if (i < lo || i < 0) return -1;
    if (lo < 0) lo = 0;
    int64_t wi = i >> 6;
    const int64_t wl = lo >> 6;
    uint64_t cur = word(wi) & (~0ull >> (63 - (i & 63)));
    for (;;) {
        if (cur) { const int64_t r = (wi << 6) + 63 - __builtin_clzll(cur); return r >= lo ? r : -1; }
        if (wi <= wl) return -1;
        cur = word(--wi);
    }
}
void BitRow::copy_range_from(const BitRow& src, int64_t a, int64_t b) {
    if (a >= b) return;
    const int64_t wa = a >> 6, wb = (b - 1) >> 6;
    for (int64_t wi = wa; wi <= wb; ++wi) {
        uint64_t m = ~0ull;
        if (wi == wa) m &= ~0ull << (a & 63);
        if (wi == wb) m &= ~0ull >> (63 - ((b - 1) & 63));
        const uint64_t v = src.word(wi) & m;
        __atomic_fetch_and(&w_[(size_t)wi], ~m | v, __ATOMIC_RELAXED);
        __atomic_fetch_or(&w_[(size_t)wi], v, __ATOMIC_RELAXED);
    }
}
int64_t BitRow::next_set(int64_t i, int64_t limit) const {
    if (i >= limit) return limit;
    int64_t wi = i >> 6, wl = (limit - 1) >> 6;
    uint64_t cur = word(wi) & (~0ull << (i & 63));
    for (;;) {
        if (cur) { int64_t r = (wi << 6) + __builtin_ctzll(cur); return r < limit ? r : limit; }
        if (wi >= wl) return limit;
        cur = word(++wi);
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
    parallel_chunks(n >= 16 ? default_host_threads() : 1, n, [&](long i) {     // src/parsnp.cpp:3181-3186 (125 MB of rows for 200 queries of 5 Mbp)
        truth_.layout[(size_t)i].init(len_[(size_t)i] + 1);
        truth_.layout[(size_t)i].set_range(len_[(size_t)i], len_[(size_t)i] + 1);
    });
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
uint64_t Aligner::coords_hash(const int64_t* p, int count) { return region_coords_hash(p, count); }
void Aligner::CoordIndex::insert(uint64_t hash, int value) {
    if ((count + 1) * 2 > s.size()) {                       // grow / rehash at 50 % load
        std::vector<Slot> os;
        os.swap(s);
        const size_t nsz = os.empty() ? 1024 : os.size() * 2;
        s.assign(nsz, Slot{0, -1, 0});
        count = 0;
        for (size_t i = 0; i < os.size(); ++i) if (os[i].v >= 0) insert(os[i].h, os[i].v);
    }
    const size_t mask = s.size() - 1;
    size_t i = (size_t)hash & mask;
    while (s[i].v >= 0) i = (i + 1) & mask;
    s[i].h = hash;
    s[i].v = value;
    ++count;
}
void Aligner::CoordIndex::reserve(size_t entries) {
    size_t need = 1024;
    while (need < (count + entries) * 2 + 2) need *= 2;
    if (need <= s.size()) return;
    std::vector<Slot> os;
    os.swap(s);
    s.assign(need, Slot{0, -1, 0});
    count = 0;
    for (size_t i = 0; i < os.size(); ++i) if (os[i].v >= 0) insert(os[i].h, os[i].v);
}
void Aligner::CoordIndex::build_parallel(const std::vector<uint64_t>& hashes, const std::vector<uint8_t>& valid, int threads) {
    const size_t N = hashes.size();
    size_t nvalid = 0;
    for (size_t r = 0; r < N; ++r) nvalid += valid[r] != 0;       // (the table is sized by what goes in: the replay indexes a tenth of the regions)
    size_t need = 1024;
    while (need < nvalid * 2 + 2) need *= 2;
    s.resize(need);
    const long per = 8192;
    parallel_chunks(need > 65536 ? threads : 1, ((long)need + per - 1) / per, [&](long c) {
        for (size_t i = (size_t)c * per; i < std::min(need, (size_t)(c + 1) * per); ++i) s[i] = Slot{0, -1, 0};
    });
    const size_t mask = need - 1;
    std::vector<size_t> cnt((size_t)(((long)N + per - 1) / per) + 1, 0);
    parallel_chunks(N > 16384 ? threads : 1, ((long)N + per - 1) / per, [&](long c) {
        size_t k = 0;
        for (size_t r = (size_t)c * per; r < std::min(N, (size_t)(c + 1) * per); ++r) {
            if (!valid[r]) continue;
            size_t i = (size_t)hashes[r] & mask;
            for (;;) {
                int32_t expect = -1;
                if (__atomic_compare_exchange_n(&s[i].v, &expect, (int32_t)r, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
                i = (i + 1) & mask;
            }
            s[i].h = hashes[r];                  // (read only after the pass has ended)
            ++k;
        }
        cnt[(size_t)c] = k;
    });
    count = 0;
    for (size_t k : cnt) count += k;
}
int Aligner::CandCache::lookup(const int64_t* coords) const {
    return lookup(coords, Aligner::coords_hash(coords, 2 * rp.n));
}
int Aligner::CandCache::lookup(const int64_t* coords, uint64_t hash) const {
    const int n = rp.n;
    const size_t bytes = sizeof(int64_t) * 2 * n;
    const int hit = map.find(hash, [&](int e) { return std::memcmp(rp.start(entries[e].region), coords, bytes) == 0; });
    if (hit >= 0 || !sorted_valid) return hit;
    // entries in ascending start[0] order that the index may not hold (the engine's discovery: regions of gaps that were
    // expected to need no lookup, build_region_index): a binary search instead
    const size_t N = entries.size(), stride = 2 * (size_t)n;
    const int64_t* base = rp.start(0);
    size_t a = 0, b = N;
    while (a < b) { const size_t mid = (a + b) >> 1; if (base[mid * stride] < coords[0]) a = mid + 1; else b = mid; }
    for (; a < N && base[a * stride] == coords[0]; ++a)
        if ((*sorted_valid)[a] && std::memcmp(base + a * stride, coords, bytes) == 0) return (int)a;
    return -1;
}

int Aligner::minsize_cached(CandCache& C, bool anchors, int64_t slength) {
    std::unordered_map<int64_t, int>& c = C.minsize[anchors ? 1 : 0];
    auto it = c.find(slength);
    if (it != c.end()) return it->second;
    int v = anchors ? anchor_expr_(slength) : mum_expr_(slength);
    c.emplace(slength, v);
    return v;
}

// ------------------------------------------------------------------ batched search (setMums1 up to the emission loop)
void Aligner::make_window_tasks(CandCache& C, const RegionPool& src, const std::vector<int>& regs, bool anchors, std::vector<WindowTask>& tasks,
                                std::vector<int64_t>& coords, std::vector<int>& first_task) {
    C.rp.n = n_;
    first_task.assign(regs.size() + 1, 0);
    const int nq = n_ - 1;
    tasks.reserve(regs.size());
    coords.reserve(regs.size() * 2 * (size_t)nq);
    for (size_t ri = 0; ri < regs.size(); ++ri) {
        const int r = regs[ri];
        const int64_t* rs = src.start(r);
        const int64_t* re = src.end(r);
        first_task[ri] = (int)tasks.size();
        const int minsize = minsize_cached(C, anchors, src.slen[r]);
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
}

void Aligner::install_search_result(CandCache& C, const RegionPool& src, const std::vector<int>& regs, const std::vector<WindowTask>& tasks,
                                    const std::vector<int>& first_task, CandBatch& cb) {
    const int nq = n_ - 1;
    if (!cb.cnt.empty()) {
        // the engine delivers the windows' candidate blocks in kernel completion order; gather them into window order so that
        // the accept passes (ascending reference order) stream through memory
        const int nt = (int)tasks.size();
        std::vector<int64_t> noff((size_t)nt + 1, 0);
        for (int t = 0; t < nt; ++t) noff[t + 1] = noff[t] + cb.cnt[t];
        const int64_t tot = noff[nt];
        pod_vector<int32_t> nk((size_t)tot), nl((size_t)tot), ns((size_t)tot * nq);
        pod_vector<uint8_t> nf((size_t)tot * nq);
        const long per = 1024;
        parallel_chunks(tot > 65536 ? threads_ : 1, ((long)nt + per - 1) / per, [&](long c) {
            for (long t = c * per; t < std::min<long>(nt, (c + 1) * per); ++t) {
                const size_t cn = (size_t)cb.cnt[t], a = (size_t)cb.off[t], b = (size_t)noff[t];
                if (!cn) continue;
                std::memcpy(nk.data() + b, cb.k.data() + a, cn * 4);
                std::memcpy(nl.data() + b, cb.lon.data() + a, cn * 4);
                if (nq) {
                    std::memcpy(ns.data() + b * nq, cb.sp.data() + a * nq, cn * nq * 4);
                    std::memcpy(nf.data() + b * nq, cb.fwd.data() + a * nq, cn * nq);
                }
            }
        });
        cb.k.swap(nk); cb.lon.swap(nl); cb.sp.swap(ns); cb.fwd.swap(nf);
        cb.off.swap(noff);
        cb.cnt.clear();
    }
    const int32_t chunk = (int32_t)C.chunks.size() - 1;
    C.entries.reserve(C.entries.size() + regs.size());
    C.map.reserve(regs.size());
    C.wins.reserve(C.wins.size() + tasks.size());
    for (size_t ri = 0; ri < regs.size(); ++ri) {
        CacheEntry e;
        e.region = C.rp.add(src.start(regs[ri]), src.end(regs[ri]));
        e.first_win = (int64_t)C.wins.size();
        e.nwin = first_task[ri + 1] - first_task[ri];
        for (int t = first_task[ri]; t < first_task[ri + 1]; ++t) {
            WinRec w;
            w.ref_start = tasks[t].ref_start;
            w.ref_len = tasks[t].ref_len;
            w.cand_off = cb.off[t];
            w.ncand = cb.count(t);
            w.chunk = chunk;
            C.wins.push_back(w);
        }
        C.map.insert(coords_hash(src.start(regs[ri]), 2 * n_), (int)C.entries.size());
        C.entries.push_back(e);
    }
}

void Aligner::search_regions(CandCache& C, const RegionPool& src, const std::vector<int>& regs, bool anchors) {
    if (regs.empty()) return;
    const double tp0 = now_s();
    std::vector<WindowTask> tasks;
    std::vector<int64_t> coords;
    std::vector<int> first_task;
    make_window_tasks(C, src, regs, anchors, tasks, coords, first_task);
    const double tp1 = now_s();
    C.chunks.emplace_back();
    CandBatch& cb = C.chunks.back();
    cb.nq = n_ - 1;
    if (!tasks.empty()) {
        std::lock_guard<std::mutex> lk(backend_mu_);       // one search at a time: the engine owns one stream and one set of buffers
        be_->search(tasks.data(), (int)tasks.size(), coords.data(), cb);
    } else cb.off.assign(1, 0);
    const double tp2 = now_s();
    install_search_result(C, src, regs, tasks, first_task, cb);
    const double tp3 = now_s();
    {
        std::lock_guard<std::mutex> lk(backend_mu_);       // (the statistics are shared between the two producer threads)
        stats_.windows_searched += (int64_t)tasks.size();
        stats_.regions_searched += (int64_t)regs.size();
        stats_.candidates += (int64_t)cb.k.size();
        stats_.t_search_prep += tp1 - tp0;
        stats_.t_search_backend += tp2 - tp1;
        stats_.t_search_cache += tp3 - tp2;
    }
}

// ------------------------------------------------------------------ setMums1 loop D (src/parsnp.cpp:1713-1842): accept_impl.h
void Aligner::accept_candidates(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& C, int cache_idx,
                                std::vector<BitRow>& layout, MumPool& mp, std::vector<int>& found, bool atomic, bool trace) {
    DirectAccess acc{layout, atomic};
    accept_candidates_t(rs, re, rsl, C, cache_idx, acc, mp, found, trace);
}

// The same loop for a big candidate list on an EMPTY layout region (the anchors): candidates whose intervals overlap no other
// candidate's interval in any genome are untouched by the trim loop whatever the order (the only bits in their intervals
// would be their own), so they are validated and placed in parallel; the overlapping ones (few) go through the literal
// sequential loop above, in candidate order, among themselves.  Result identical to accept_candidates().
void Aligner::accept_candidates_parallel(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& CC, int cache_idx,
                                         std::vector<BitRow>& layout, MumPool& mp, std::vector<int>& found, bool trace) {
    const CacheEntry& ce = CC.entries[cache_idx];
    const int nq = n_ - 1;
    const size_t N = (size_t)n_;
    std::vector<int64_t> wbase((size_t)ce.nwin + 1, 0);
    for (int wi = 0; wi < ce.nwin; ++wi) {
        const WinRec& win = CC.wins[ce.first_win + wi];
        wbase[wi + 1] = wbase[wi] + win.ncand;
        if (trace) trace_.emplace_back(win.ref_start, win.ref_len);
    }
    const size_t C = (size_t)wbase[ce.nwin];
    if (C == 0) return;
    pod_vector<int64_t> ST(C * N), STT(C * N), LON(C);          // starts: candidate-major and genome-major
    pod_vector<uint8_t> FW(C * N), state(C);                     // state: 0 = skipped, 1 = valid, 3 = valid + overlapping
    const long per = 256;
    const long nblk = ((long)C + per - 1) / per;
    // ---- pass 1: coordinates and the pre-checks of src/parsnp.cpp:1723 / TMum ctor, per candidate
    parallel_chunks(threads_, nblk, [&](long b) {
        const size_t c0 = (size_t)b * per, c1 = std::min(C, c0 + (size_t)per);
        int wi = (int)(std::upper_bound(wbase.begin(), wbase.end(), (int64_t)c0) - wbase.begin()) - 1;
        for (size_t c = c0; c < c1; ++c) {
            while ((int64_t)c >= wbase[wi + 1]) ++wi;
            const WinRec& win = CC.wins[ce.first_win + wi];
            const CandBatch& cb = CC.chunks[win.chunk];
            const int64_t ci = win.cand_off + ((int64_t)c - wbase[wi]);
            const int64_t lon = cb.LON()[ci];
            int64_t* st = &ST[c * N];
            uint8_t* fw = &FW[c * N];
            bool bad = false;
            const uint64_t dsp0 = (uint64_t)((int64_t)cb.K()[ci] + 1 + win.ref_start);
            if ((uint64_t)(dsp0 - (uint64_t)rs[0]) > (uint64_t)(uint32_t)(re[0] - rs[0])) bad = true;
            st[0] = (int64_t)dsp0 - 1;
            fw[0] = 1;
            bool any_fail = st[0] + lon > len_[0] || st[0] < 0;
            const int32_t* spj = cb.SP() + ci * nq;
            const uint8_t* fwj = cb.FWD() + ci * nq;
            for (int j = 1; j < n_; ++j) {
                const uint64_t dsp = (uint64_t)((int64_t)spj[j - 1] + 1 + rs[j]);
                bad |= (uint64_t)(dsp - (uint64_t)rs[j]) > (uint64_t)(uint32_t)(re[j] - rs[j]);
                int64_t s = (int64_t)dsp - 1;
                const uint8_t f = fwj[j - 1];
                if (!f) s = len_[j] - (s + lon);
                any_fail |= (s + lon > len_[j]) | (s < 0);
                st[j] = s;
                fw[j] = f;
            }
            LON[c] = lon;
            state[c] = (bad || any_fail || lon < 5) ? 0 : 1;
        }
        for (int j = 0; j < n_; ++j) {                      // genome-major copy of the block (contiguous runs)
            int64_t* d = &STT[(size_t)j * C];
            for (size_t c = c0; c < c1; ++c) d[c] = ST[c * N + j];
        }
    });
    // ---- pass 2: per genome, mark the valid candidates whose interval overlaps another valid candidate's
    parallel_chunks(threads_, (long)n_, [&](long j) {
        const int64_t* sj = &STT[(size_t)j * C];
        auto mark = [&](size_t c) { __atomic_store_n(&state[c], (uint8_t)3, __ATOMIC_RELAXED); };
        bool sorted = true;
        int64_t prev = INT64_MIN;
        for (size_t c = 0; c < C && sorted; ++c) {
            if (!(__atomic_load_n(&state[c], __ATOMIC_RELAXED) & 1)) continue;
            if (sj[c] < prev) sorted = false;
            prev = sj[c];
        }
        int64_t max_end = INT64_MIN;
        size_t arg = 0;
        auto visit = [&](size_t c) {
            const int64_t s = sj[c], e = s + LON[c];
            if (s < max_end) { mark(c); mark(arg); }
            if (e > max_end) { max_end = e; arg = c; }
        };
        if (sorted) {
            for (size_t c = 0; c < C; ++c) if (__atomic_load_n(&state[c], __ATOMIC_RELAXED) & 1) visit(c);
        } else {
            std::vector<std::pair<int64_t, uint32_t>> iv;
            iv.reserve(C);
            for (size_t c = 0; c < C; ++c) if (__atomic_load_n(&state[c], __ATOMIC_RELAXED) & 1) iv.emplace_back(sj[c], (uint32_t)c);
            std::sort(iv.begin(), iv.end());
            for (const auto& x : iv) visit(x.second);
        }
    });
    // ---- pass 3a: the overlapping candidates, literally as in accept_candidates(), in candidate order
    pod_vector<int64_t> LEN(C);                                 // accepted length, 0 = not accepted
    for (size_t c = 0; c < C; ++c) {
        LEN[c] = 0;
        if (state[c] != 3) continue;
        int64_t* st = &ST[c * N];
        const uint8_t* fw = &FW[c * N];
        int64_t length = LON[c];
        for (int j = 0; j < n_; ++j) {
            int64_t t1 = layout[j].run_up(st[j], st[j] + length);
            if (t1) { for (int i = 0; i < n_; ++i) st[i] += t1; length -= t1; }
            int64_t t2 = layout[j].run_down(st[j], st[j] + length);
            length -= t2;
            if (length <= 0) break;
        }
        if (length < 2 || n_ <= 1) continue;
        bool badmum = false;
        for (int k = 0; k < n_ && !badmum; ++k) {
            if (fw[k]) continue;
            const uint8_t* g0 = seq_[0] + st[0];
            const uint8_t* gk = seq_[k] + st[k];
            for (int64_t t = 0; t < length; ++t)
                if (comp_base(gk[length - 1 - t]) != g0[t]) { badmum = true; break; }
        }
        if (badmum) continue;
        for (int k = 0; k < n_; ++k) layout[k].set_range(st[k], st[k] + length);
        LEN[c] = length;
    }
    // ---- pass 3b: everybody else, in parallel (nothing to trim)
    parallel_chunks(threads_, nblk, [&](long b) {
        const size_t c0 = (size_t)b * per, c1 = std::min(C, c0 + (size_t)per);
        for (size_t c = c0; c < c1; ++c) {
            if (state[c] != 1) continue;
            const int64_t length = LON[c];
            if (length < 2 || n_ <= 1) continue;
            const int64_t* st = &ST[c * N];
            const uint8_t* fw = &FW[c * N];
            bool badmum = false;
            for (int k = 0; k < n_ && !badmum; ++k) {
                if (fw[k]) continue;
                const uint8_t* g0 = seq_[0] + st[0];
                const uint8_t* gk = seq_[k] + st[k];
                for (int64_t t = 0; t < length; ++t)
                    if (comp_base(gk[length - 1 - t]) != g0[t]) { badmum = true; break; }
            }
            if (badmum) continue;
            for (int k = 0; k < n_; ++k) layout[k].set_range_atomic(st[k], st[k] + length);
            LEN[c] = length;
        }
    });
    if (getenv("PB200_PROFILE_HOST")) {
        size_t nv = 0, no = 0;
        for (size_t c = 0; c < C; ++c) { nv += state[c] & 1; no += state[c] == 3; }
        fprintf(stderr, "[pb200 anchors] parallel accept: %zu candidates, %zu valid, %zu overlapping (sequential)\n", C, nv, no);
    }
    // ---- accepted MUMs into the pool, in candidate order
    std::vector<size_t> blk_cnt((size_t)nblk + 1, 0);
    for (long b = 0; b < nblk; ++b) {
        size_t k = 0;
        for (size_t c = (size_t)b * per; c < std::min(C, (size_t)(b + 1) * per); ++c) k += LEN[c] > 0;
        blk_cnt[b + 1] = blk_cnt[b] + k;
    }
    const size_t A = blk_cnt[nblk];
    const size_t m0 = mp.mums.size(), s0 = mp.start.size();
    mp.mums.resize(m0 + A);
    mp.start.resize(s0 + A * N);
    mp.fwd.resize(s0 + A * N);
    const size_t f0 = found.size();
    found.resize(f0 + A);
    parallel_chunks(threads_, nblk, [&](long b) {
        size_t a = blk_cnt[b];
        for (size_t c = (size_t)b * per; c < std::min(C, (size_t)(b + 1) * per); ++c) {
            if (LEN[c] <= 0) continue;
            MumRec m;
            m.length = LEN[c];
            m.slength = rsl;
            m.off = (int64_t)(s0 + a * N);
            m.alive = true;
            std::memcpy(&mp.start[s0 + a * N], &ST[c * N], sizeof(int64_t) * N);
            std::memcpy(&mp.fwd[s0 + a * N], &FW[c * N], N);
            mp.mums[m0 + a] = m;
            found[f0 + a] = (int)(m0 + a);
            ++a;
        }
    });
}

// determineRegion (src/parsnp.cpp:1199-1290) on one bitmap: accept_impl.h
static int64_t det_region(const std::vector<BitRow>& layout, const std::vector<int64_t>& len, int n,
                          const int64_t* mstart, int64_t mlen, bool left, int64_t* S, int64_t* E) {
    DirectAccess acc{const_cast<std::vector<BitRow>&>(layout), false};
    return det_region_impl(acc, len, n, mstart, mlen, left, S, E);
}

namespace {
std::shared_ptr<const std::vector<int32_t>> minsize_table(const std::string& expr, int N, int threads);
}

// the anchor stage as the engine delivered it: anchors into the MUM pool, mumlayout, the regions between the anchors
void Aligner::install_device_anchors(const AnchorResult& res, int whole) {
    const size_t NA = res.nanchors, NR = res.nregions, N = (size_t)n_;
    const int64_t rsl = rp_.slen[(size_t)whole];
    {   // room for the recursion's MUMs too (about 3x the anchors on divergent genomes)
        const size_t est = NA * 4 + 1024;
        mp_.mums.reserve(est);
        mp_.start.reserve(est * N);
        mp_.fwd.reserve(est * N);
    }
    mp_.mums.resize(NA);
    mp_.start.resize(NA * N);
    mp_.fwd.resize(NA * N);
    all_mums_.resize(NA);
    const long per = 2048;
    parallel_chunks(NA > 8192 ? threads_ : 1, ((long)NA + per - 1) / per, [&](long c) {
        for (size_t x = (size_t)c * per; x < std::min(NA, (size_t)(c + 1) * per); ++x) {
            MumRec m;
            m.length = res.a_lon[x];
            m.slength = rsl;
            m.off = (int64_t)(x * N);
            m.alive = true;
            mp_.mums[x] = m;
            for (size_t g = 0; g < N; ++g) { mp_.start[x * N + g] = res.a_start[x * N + g]; mp_.fwd[x * N + g] = res.a_fwd[x * N + g]; }
            all_mums_[x] = (int)x;
        }
    });
    stats_.anchors = (int64_t)NA;
    if (NA == 0) return;                         // (NO MUMS FOUND: the layout stays as constructed)
    parallel_chunks(threads_, (long)n_, [&](long g) {
        BitRow& row = truth_.layout[(size_t)g];
        std::memcpy(row.words_mut(), res.layout + res.layout_off[(size_t)g], (size_t)row.nwords() * sizeof(uint64_t));
    });
    const int base = rp_.size();
    rp_.coord.resize(rp_.coord.size() + NR * 2 * N);
    rp_.slen.resize(rp_.slen.size() + NR);
    initial_regions_.resize(NR);
    initial_lo_.clear();
    if (res.r_lo) initial_lo_.resize(NR * N);
    parallel_chunks(NR > 8192 ? threads_ : 1, ((long)NR + per - 1) / per, [&](long c) {
        for (size_t r = (size_t)c * per; r < std::min(NR, (size_t)(c + 1) * per); ++r) {
            if (res.r_lo) std::memcpy(&initial_lo_[r * N], res.r_lo + r * N, N * sizeof(int32_t));
            const int32_t* s = res.r_coords + r * 2 * N;
            int64_t* d = &rp_.coord[((size_t)base + r) * 2 * N];
            int64_t sl = 500000000;
            for (size_t g = 0; g < N; ++g) { d[g] = s[g]; d[N + g] = (int64_t)s[g] + s[N + g]; sl = std::min<int64_t>(sl, s[N + g]); }
            rp_.slen[(size_t)base + r] = sl;
            initial_regions_[r] = base + (int)r;
        }
    });
    anchors_on_device_ = true;
}

// ------------------------------------------------------------------ anchors (src/parsnp.cpp:2121-2174)
void Aligner::set_initial_clusters() {
    double t0 = now_s();
    std::vector<int64_t> S(n_, 0), E(len_);
    int whole = rp_.add(S.data(), E.data());
    // The engine can take the whole anchor stage (search + accept on the empty layout + the regions between the anchors:
    // cuda/anchors.cuh) and start following the recursion from there.  It declines (status 0: tiny genomes, a backend without
    // it) or hands the candidates back (status 2: overlapping or non-collinear candidates - the accept below is needed).
    int status = 0;
    if (speculate_ && pipeline_ && !trace_on_ && n_ > 1) {
        std::shared_ptr<const std::vector<int32_t>> tab = minsize_table(prm_.mums, 4160, threads_);
        if (tab) {
            std::vector<WindowTask> tasks;
            std::vector<int64_t> coords;
            std::vector<int> first_task;
            const std::vector<int> regs(1, whole);
            make_window_tasks(main_cache_, rp_, regs, true, tasks, coords, first_task);
            std::vector<int64_t> nwords((size_t)n_);
            for (int g = 0; g < n_; ++g) nwords[(size_t)g] = truth_.layout[(size_t)g].nwords();
            AnchorRequest rq;
            rq.n = n_; rq.tasks = tasks.data(); rq.ntasks = (int)tasks.size(); rq.coords = coords.data(); rq.layout_words = nwords.data();
            rq.q = prm_.q; rq.p = prm_.p; rq.minsize_tab = tab->data(); rq.minsize_n = (int)tab->size();
            rq.follow_recursion = !prm_.anchors_only;
            AnchorResult res;
            {
                std::lock_guard<std::mutex> lk(backend_mu_);
                status = be_->anchor_stage(rq, res);
            }
            if (status != 0) {
                stats_.windows_searched += (int64_t)tasks.size();
                stats_.regions_searched += 1;
                stats_.candidates += (int64_t)(status == 1 ? res.ncand : res.cands.k.size());
            }
            if (status == 1) {
                stats_.t_anchor_search = now_s() - t0;
                install_device_anchors(res, whole);
                stats_.t_anchor_host = now_s() - t0 - stats_.t_anchor_search;
                return;
            }
            if (status == 2) {
                main_cache_.chunks.emplace_back(std::move(res.cands));
                install_search_result(main_cache_, rp_, regs, tasks, first_task, main_cache_.chunks.back());
            }
        }
    }
    if (status == 0) search_regions(main_cache_, rp_, std::vector<int>(1, whole), true);
    double t1 = now_s();
    stats_.t_anchor_search = t1 - t0;
    std::vector<int> found;
    {   // room for the anchors plus the recursion's MUMs (about 3x the anchors on divergent genomes): no regrowth copies below
        const size_t est = (size_t)stats_.candidates * 4 + 1024;
        mp_.mums.reserve(est);
        mp_.start.reserve(est * (size_t)n_);
        mp_.fwd.reserve(est * (size_t)n_);
        found.reserve((size_t)stats_.candidates);
    }
    const char* pmin = getenv("PB200_PAR_ANCHORS_MIN");          // tests: 1 forces the parallel accept, a huge value the serial one
    if (threads_ > 1 && stats_.candidates >= (pmin ? atoll(pmin) : 8192))
        accept_candidates_parallel(rstart(whole), rend(whole), rp_.slen[whole], main_cache_, main_cache_.lookup(rstart(whole)), truth_.layout, mp_,
                                   found, trace_on_);
    else
        accept_candidates(rstart(whole), rend(whole), rp_.slen[whole], main_cache_, main_cache_.lookup(rstart(whole)), truth_.layout, mp_, found,
                          false, trace_on_);
    all_mums_ = found;
    stats_.anchors = (int64_t)found.size();
    const double ta1 = now_s();
    double t_det = 0;
    // determineRegion of every anchor: mumlayout is final here (all anchors placed), so the scans are independent and run
    // in parallel over blocks of anchors; the push rules (src/parsnp.cpp:2153-2172) are then applied in order
    // The push rules compare an anchor's left region with the previous anchor's right region and with its own right region:
    // pairwise tests on the scan results, evaluated in parallel; the pool is then grown once and filled in parallel.
    const size_t B = 262144;
    const size_t stride = 4 * (size_t)n_;
    pod_vector<int64_t> buf(stride * std::min(B, found.size() + 1));
    pod_vector<int64_t> sl(2 * std::min(B, found.size() + 1));
    pod_vector<int32_t> slot(2 * std::min(B, found.size() + 1) + 1);
    const double ta2 = now_s();
    std::vector<int64_t> prevS(n_), prevE(n_);
    const size_t cbytes = sizeof(int64_t) * (size_t)n_;
    for (size_t b0 = 0; b0 < found.size(); b0 += B) {
        const size_t bn = std::min(B, found.size() - b0);
        const long per = 256;
        const int T = bn > 512 ? threads_ : 1;
        const double td0 = now_s();
        parallel_chunks(T, ((long)bn + per - 1) / per, [&](long c) {
            for (long x = c * per; x < std::min<long>((long)bn, (c + 1) * per); ++x) {
                const MumRec& m = mums_[found[b0 + x]];
                const int64_t* ms = &mum_start_[m.off];
                int64_t* p = &buf[(size_t)x * stride];
                sl[2 * x] = det_region(truth_.layout, len_, n_, ms, m.length, true, p, p + n_);
                sl[2 * x + 1] = det_region(truth_.layout, len_, n_, ms, m.length, false, p + 2 * n_, p + 3 * n_);
            }
        });
        t_det += now_s() - td0;
        parallel_chunks(T, ((long)bn + per - 1) / per, [&](long c) {
            for (long x = c * per; x < std::min<long>((long)bn, (c + 1) * per); ++x) {
                const int64_t* lS = &buf[(size_t)x * stride];
                const int64_t* rS = lS + 2 * n_;
                const int64_t* pS = x ? lS - 2 * n_ : prevS.data();        // previous anchor's right region (start, end contiguous)
                const int64_t* pE = x ? lS - n_ : prevE.data();
                const bool first = (b0 + (size_t)x) == 0;
                const bool l_eq_r = !first && std::memcmp(lS, pS, cbytes) == 0 && std::memcmp(lS + n_, pE, cbytes) == 0;
                const bool r_eq_l = std::memcmp(lS, rS, 2 * cbytes) == 0;
                slot[2 * x] = (sl[2 * x] > prm_.q && (first || !l_eq_r)) ? 1 : 0;
                slot[2 * x + 1] = (sl[2 * x + 1] > prm_.q && !r_eq_l) ? 1 : 0;
            }
        });
        int32_t total = 0;
        for (size_t x = 0; x < 2 * bn; ++x) { const int32_t f = slot[x]; slot[x] = total; total += f; }
        slot[2 * bn] = total;
        const int base = rp_.size();
        rp_.coord.resize(rp_.coord.size() + (size_t)total * 2 * n_);
        rp_.slen.resize(rp_.slen.size() + (size_t)total);
        const size_t ir0 = initial_regions_.size();
        initial_regions_.resize(ir0 + (size_t)total);
        parallel_chunks(T, ((long)bn + per - 1) / per, [&](long c) {
            for (long x = c * per; x < std::min<long>((long)bn, (c + 1) * per); ++x) {
                for (int side = 0; side < 2; ++side) {
                    const int32_t at = slot[2 * x + side];
                    if (slot[2 * x + side + 1] == at) continue;
                    const int id = base + at;
                    std::memcpy(&rp_.coord[(size_t)id * 2 * n_], &buf[(size_t)x * stride + (size_t)side * 2 * n_], 2 * cbytes);
                    rp_.slen[id] = sl[2 * x + side];
                    initial_regions_[ir0 + (size_t)at] = id;
                }
            }
        });
        std::memcpy(prevS.data(), &buf[(bn - 1) * stride + 2 * n_], cbytes);
        std::memcpy(prevE.data(), &buf[(bn - 1) * stride + 3 * n_], cbytes);
    }
    stats_.t_anchor_host = now_s() - t1;
    if (getenv("PB200_PROFILE_HOST"))
        fprintf(stderr, "[pb200 anchors ms] accept %.2f alloc %.2f det_region %.2f push %.2f\n", (ta1 - t1) * 1e3, (ta2 - ta1) * 1e3, t_det * 1e3,
                (now_s() - ta2 - t_det) * 1e3);
}

// ------------------------------------------------------------------ speculative level-synchronous discovery
// One level over frontier[a,b) (sorted by start[0]): accept on the scratch layout, collect the children's coordinates.
// Only a predictor of which regions the exact replay will ask for - races between threads merely cost cache misses.
void Aligner::speculate_range(const CandCache& C, const RegionPool& F, const std::vector<int>& frontier, size_t a, size_t b,
                              std::vector<BitRow>& layout, MumPool& mp, RegionPool& out, bool atomic) {
    std::vector<int> found;
    std::vector<int64_t> lS(n_), lE(n_), rS(n_), rE(n_);
    const size_t cbytes = sizeof(int64_t) * 2 * n_;
    int prev = -1;
    for (size_t x = a; x < b; ++x) {
        const int r = frontier[x];
        if (prev >= 0 && std::memcmp(F.start(prev), F.start(r), cbytes) == 0) continue;
        prev = r;
        found.clear();
        const int ci = C.lookup(F.start(r));
        if (ci < 0) continue;
        accept_candidates(F.start(r), F.end(r), F.slen[r], C, ci, layout, mp, found, atomic, false);
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

void Aligner::speculate_slice(CandCache& C, const RegionPool& src, const std::vector<int>& initial, World& spec) {
    RegionPool F;                                          // the slice's frontier regions, level after level
    F.n = n_;
    std::vector<int> frontier, need;
    frontier.reserve(initial.size());
    for (int r : initial) frontier.push_back(F.add(src.start(r), src.end(r)));
    const size_t cbytes = sizeof(int64_t) * 2 * n_;
    while (!frontier.empty()) {
        double t0 = now_s();
        need.clear();
        {
            CoordIndex seen;
            seen.reserve(frontier.size());
            for (int r : frontier) {
                const uint64_t h = coords_hash(F.start(r), 2 * n_);
                if (C.lookup(F.start(r), h) >= 0) continue;
                if (seen.find(h, [&](int o) { return std::memcmp(F.start(o), F.start(r), cbytes) == 0; }) >= 0) continue;
                seen.insert(h, r);
                need.push_back(r);
            }
        }
        search_regions(C, F, need, false);
        stats_.spec_regions += (int64_t)need.size();
        stats_.spec_levels++;
        double t1 = now_s();
        stats_.t_spec_search += t1 - t0;
        std::stable_sort(frontier.begin(), frontier.end(), [&](int a, int b) { return F.start(a)[0] < F.start(b)[0]; });
        // chunks of the frontier are processed concurrently on the shared scratch layout
        // (pipelined: one core of this rank's share belongs to the replay thread)
        const size_t share = (size_t)std::max(1, pipeline_ ? threads_ - 1 : threads_);
        const int T = pipeline_ ? (int)std::max<size_t>(1, std::min<size_t>(share, frontier.size() / 256 + 1)) : 1;   // lock step: aligner.h
        const size_t nchunks = T > 1 ? (size_t)T * 4 : 1;
        std::vector<RegionPool> outs(nchunks);
        for (auto& o : outs) o.n = n_;
        parallel_chunks(T, (long)nchunks, [&](long c) {
            MumPool mp;
            const size_t a = frontier.size() * (size_t)c / nchunks, b = frontier.size() * (size_t)(c + 1) / nchunks;
            speculate_range(C, F, frontier, a, b, spec.layout, mp, outs[c], T > 1);
        });
        std::vector<int> next;
        for (auto& o : outs)
            for (int i = 0; i < o.size(); ++i) next.push_back(F.add(o.start(i), o.end(i)));
        frontier.swap(next);
        stats_.t_spec_host += now_s() - t1;
    }
}

// the speculation thread: slice after slice, each published as soon as it is complete
void Aligner::speculation_thread_main() {
    try {
        for (size_t k = 0; k < slice_cache_.size(); ++k) {
            speculate_slice(*slice_cache_[k], frozen_rp_, slice_regions_[k], spec_world_);
            {
                std::lock_guard<std::mutex> lk(slice_mu_);
                slices_ready_ = (int)k + 1;
            }
            slice_cv_.notify_all();
        }
    } catch (...) {
        std::lock_guard<std::mutex> lk(slice_mu_);
        spec_error_ = std::current_exception();
        slices_ready_ = (int)slice_cache_.size();          // nobody waits for ever; the replay searches on demand and run() rethrows
        slice_cv_.notify_all();
    }
}

const Aligner::CandCache* Aligner::wait_slice(int slice) {
    if (slice < 0 || slice >= (int)slice_cache_.size()) return nullptr;
    std::unique_lock<std::mutex> lk(slice_mu_);
    if (slices_ready_ <= slice) {
        const double t0 = now_s();
        slice_cv_.wait(lk, [&] { return slices_ready_ > slice; });
        stats_.t_replay_wait += now_s() - t0;
    }
    return slice_cache_[slice].get();
}

// ------------------------------------------------------------------ doWork (src/parsnp.cpp:173-317), exact order
namespace {
struct QE { int64_t s0; int id; int slice; uint64_t hash; };      // slice: the speculation slice the region descends from; hash of the coordinates
inline bool operator<(const QE& a, const QE& b) { return a.s0 < b.s0; }   // operator<, src/LCR.cpp:42
}

void Aligner::process_queue_exact(const std::vector<int>& initial, RegionPool& rp, std::vector<BitRow>& layout, MumPool& mp,
                                  std::vector<int>& out_mums, const std::vector<int>* slice_ids) {
    // exact emulation of `vector<TRegion> regions`: slow mode keeps the vector itself; fast mode is valid while
    // all start[0] keys are distinct (then every correct sort yields the same sequence).
    auto req = [&](int a, int b) { return std::memcmp(rp.start(a), rp.start(b), sizeof(int64_t) * 2 * n_) == 0; };
    std::vector<QE> vec;
    for (size_t i = 0; i < initial.size(); ++i)
        vec.push_back(QE{rp.start(initial[i])[0], initial[i], slice_ids ? (*slice_ids)[i] : (i < slice_of_initial_.size() ? slice_of_initial_[i] : -1),
                         coords_hash(rp.start(initial[i]), 2 * n_)});
    int ready_upto = 0;                       // speculation slices [0, ready_upto) are known to be published
    // fast mode: the queue as a vector sorted by DESCENDING start[0] (front of the reference's vector = back of this one).
    // Children of the region just taken lie inside it, i.e. next to the back, so insertion is a short walk + a short move.
    std::vector<QE> fast;
    auto fast_pos = [&](int64_t key) {          // index of the first element (from the front) with s0 <= key
        size_t pos = fast.size();
        int steps = 0;
        while (pos > 0 && fast[pos - 1].s0 < key) {
            --pos;
            if (++steps == 16) {                 // far from the back: binary search over the descending prefix
                pos = (size_t)(std::lower_bound(fast.begin(), fast.begin() + (long)pos, key,
                                                [](const QE& e, int64_t k) { return e.s0 > k; }) - fast.begin());
                return pos;
            }
        }
        // here fast[pos-1].s0 >= key (or pos == 0)
        if (pos > 0 && fast[pos - 1].s0 == key) return pos - 1;
        return pos;
    };
    bool fast_mode = false;
    std::vector<int> found, children;
    std::vector<int64_t> lS(n_), lE(n_), rS(n_), rE(n_);
    static const bool prof = getenv("PB200_PROFILE_HOST") != nullptr;
    uint64_t pc[5] = {0, 0, 0, 0, 0}, pt = 0;
#if defined(__x86_64__)
#define PB_TICKS() __builtin_ia32_rdtsc()
#else
#define PB_TICKS() ((uint64_t)std::chrono::steady_clock::now().time_since_epoch().count())
#endif
#define PROF_MARK(i) do { if (prof) { uint64_t x_ = PB_TICKS(); pc[i] += x_ - pt; pt = x_; } } while (0)
    if (prof) pt = PB_TICKS();
    while (fast_mode ? !fast.empty() : !vec.empty()) {
        int cur, cur_slice;
        uint64_t cur_hash;
        if (fast_mode) { cur = fast.back().id; cur_slice = fast.back().slice; cur_hash = fast.back().hash; fast.pop_back(); }
        else { cur = vec.front().id; cur_slice = vec.front().slice; cur_hash = vec.front().hash; vec.erase(vec.begin()); }
        PROF_MARK(0);
        // candidates: the slice's speculation (wait for it if it is still in flight), else the main cache, else search now
        const CandCache* C = nullptr;
        if (cur_slice >= 0 && cur_slice < (int)slice_cache_.size()) {
            if (cur_slice >= ready_upto) { wait_slice(cur_slice); ready_upto = cur_slice + 1; }
            C = slice_cache_[cur_slice].get();
        }
        int ci = C ? C->lookup(rp.start(cur), cur_hash) : -1;
        if (ci < 0) { C = &main_cache_; ci = main_cache_.lookup(rp.start(cur), cur_hash); }
        PROF_MARK(1);
        if (ci < 0) {
            double ts = now_s();
            search_regions(main_cache_, rp, std::vector<int>(1, cur), false);          // a region the speculation did not predict
            stats_.t_replay_search += now_s() - ts;
            stats_.replay_misses++;
            ci = main_cache_.lookup(rp.start(cur));
        }
        found.clear();
        accept_candidates(rp.start(cur), rp.end(cur), rp.slen[cur], *C, ci, layout, mp, found, false, trace_on_);
        PROF_MARK(2);
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
        PROF_MARK(3);
        // sort + drop adjacent duplicates (src/parsnp.cpp:291-306)
        if (fast_mode) {
            bool distinct_tie = false;
            for (size_t a = 0; a < children.size() && !distinct_tie; ++a) {
                const int64_t key = rp.start(children[a])[0];
                const size_t pos = fast_pos(key);
                if (pos < fast.size() && fast[pos].s0 == key && !req(fast[pos].id, children[a])) distinct_tie = true;
                for (size_t b = 0; b < a && !distinct_tie; ++b)
                    if (key == rp.start(children[b])[0] && !req(children[a], children[b])) distinct_tie = true;
            }
            if (!distinct_tie) {
                for (int ch : children) {                                    // identical duplicates collapse (the first one stays)
                    const int64_t key = rp.start(ch)[0];
                    const size_t pos = fast_pos(key);
                    if (pos < fast.size() && fast[pos].s0 == key) continue;
                    const uint64_t hh = coords_hash(rp.start(ch), 2 * n_);
                    if (cur_slice >= 0 && cur_slice < (int)slice_cache_.size()) slice_cache_[cur_slice]->map.prefetch(hh);   // (looked up a few pops from now)
                    fast.insert(fast.begin() + (long)pos, QE{key, ch, cur_slice, hh});
                }
                PROF_MARK(4);
                continue;
            }
            vec.assign(fast.rbegin(), fast.rend());
            fast.clear();
            fast_mode = false;
        }
        stats_.slow_queue_iters++;
        for (int ch : children) vec.push_back(QE{rp.start(ch)[0], ch, cur_slice, coords_hash(rp.start(ch), 2 * n_)});
        if (!vec.empty()) {
            // The queue is nearly sorted (anchor order, then children next to their parent).  With distinct start[0] keys every
            // correct sort gives the same sequence, so try a bounded insertion sort on a copy; ties (or too much disorder) fall
            // back to the literal std::sort call of the reference on the untouched vector.
            std::vector<QE> tmp(vec);
            size_t budget = 8 * tmp.size() + 64;
            bool done = true;
            for (size_t i = 1; i < tmp.size() && done; ++i) {
                if (!(tmp[i] < tmp[i - 1])) continue;
                const QE x = tmp[i];
                size_t j = i;
                while (j > 0 && x < tmp[j - 1]) {
                    tmp[j] = tmp[j - 1];
                    --j;
                    if (--budget == 0) { done = false; break; }
                }
                tmp[j] = x;
            }
            bool distinct = done;
            for (size_t m = 0; distinct && m + 1 < tmp.size(); ++m) if (!(tmp[m].s0 < tmp[m + 1].s0)) distinct = false;
            if (distinct) vec.swap(tmp);
            else std::sort(vec.begin(), vec.end());
        }
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
            fast.assign(vec.rbegin(), vec.rend());
            vec.clear();
            fast_mode = true;
        }
        PROF_MARK(4);
    }
    if (prof) fprintf(stderr, "[pb200 replay cycles] pop %llu lookup %llu accept %llu det_region %llu queue %llu\n", (unsigned long long)pc[0],
                      (unsigned long long)pc[1], (unsigned long long)pc[2], (unsigned long long)pc[3], (unsigned long long)pc[4]);
#undef PROF_MARK
}

// the deferred index of the discovery's regions (discover_slice): over all searched regions, or without those of `skip`
void Aligner::build_region_index(const std::vector<uint8_t>* skip) {
    if (!disc_index_deferred_ || slice_cache_.empty()) return;
    if (skip) {
        const long per = 8192, N = (long)disc_valid_.size();
        std::vector<uint8_t> v(disc_valid_);
        parallel_chunks(N > 32768 ? threads_ : 1, (N + per - 1) / per, [&](long c) {
            for (long r = c * per; r < std::min(N, (c + 1) * per); ++r) if ((*skip)[(size_t)r]) v[(size_t)r] = 0;
        });
        slice_cache_[0]->map.build_parallel(disc_hashes_, v, threads_);
        return;                                 // (stays "deferred": a fallback to the sequential loop asks for the full index)
    }
    slice_cache_[0]->map.build_parallel(disc_hashes_, disc_valid_, threads_);
    disc_index_deferred_ = false;
}

void Aligner::do_work_exact() {
    double t0 = now_s();
    if (!do_work_parallel()) {                 // replay.cpp: independent gaps on several threads when the anchors allow it
        build_region_index(nullptr);
        std::vector<int> out;
        process_queue_exact(initial_regions_, rp_, truth_.layout, mp_, out);
        all_mums_.insert(all_mums_.end(), out.begin(), out.end());
    }
    stats_.t_replay += now_s() - t0;
}

// sort(this->mums) by start[0] (operator<, src/TMum.cpp:151).  With distinct starts (the rule: accepted MUMs are disjoint on
// the reference) the order is unique: compact (start0,id) pairs are merged / radix-sorted and the work is skipped when already
// sorted.  With ties the reference's own std::sort call is replayed literally (see below).
void Aligner::sort_final_mums() {
    const size_t M = final_mums_.size();
    if (final_sorted_) return;                  // (removals keep the order; only `final_mums_ = all_mums_` resets the flag)
    const double ts0 = now_s();
    struct Report { double t0; size_t* d; ~Report() { if (getenv("PB200_PROFILE_HOST")) fprintf(stderr, "[pb200 sort_final_mums ms] %.2f descents %zu\n", (now_s() - t0) * 1e3, *d); } };
    size_t descents = 0;
    Report rep{ts0, &descents};
    if (sorted_hint_.size() == M && M > 0) {
        // the parallel replay delivered the ascending order with its MUMs (every task's MUMs sorted by its worker, tasks in
        // reference order, merged with the anchors): verify it in parallel; ties still go through the literal std::sort below
        const long per = 8192;
        const long nch = ((long)M + per - 1) / per;
        std::vector<uint8_t> ch_bad((size_t)nch, 0), ch_tie((size_t)nch, 0);
        std::vector<int64_t> ch_minlen((size_t)nch, INT64_MAX);
        std::vector<std::vector<int64_t>> ch_tkeys((size_t)nch);       // the keys that occur more than once (ascending)
        parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
            const size_t i0 = (size_t)c * per, i1 = std::min(M, (size_t)(c + 1) * per);
            int64_t prev = i0 ? mum_start_[mums_[sorted_hint_[i0 - 1]].off] : INT64_MIN, ml = INT64_MAX;
            for (size_t i = i0; i < i1; ++i) {
                const MumRec& m = mums_[sorted_hint_[i]];
                const int64_t s0 = mum_start_[m.off];
                if (s0 < prev) ch_bad[c] = 1;
                if (s0 == prev) { ch_tie[c] = 1; ch_tkeys[(size_t)c].push_back(s0); }
                prev = s0;
                ml = std::min(ml, m.length);
            }
            ch_minlen[c] = ml;
        });
        bool bad = false, tie = false;
        int64_t ml = INT64_MAX;
        for (long c = 0; c < nch; ++c) { bad |= ch_bad[c] != 0; tie |= ch_tie[c] != 0; ml = std::min(ml, ch_minlen[c]); }
        if (!bad && !tie) {
            final_mums_ = sorted_hint_;
            sorted_hint_.clear();
            final_min_length_ = ml;
            final_sorted_ = true;
            return;
        }
        if (bad) sorted_hint_.clear();          // (an unexpected order: the general path)
        if (!bad) {
            // ascending with ties: only the literal call on the initial order can say how the reference orders them
            std::vector<std::pair<int64_t, int>> kv0(M), kvh(M);        // the records in the initial order / in the hint's ascending order
            const std::vector<int> hint_ids(std::move(sorted_hint_));
            sorted_hint_.clear();
            parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
                for (size_t i = (size_t)c * per; i < std::min(M, (size_t)(c + 1) * per); ++i) {
                    kv0[i] = std::make_pair(mum_start_[mums_[final_mums_[i]].off], final_mums_[i]);
                    kvh[i] = std::make_pair(mum_start_[mums_[hint_ids[i]].off], hint_ids[i]);
                }
            });
            std::vector<int64_t> tkeys;
            for (const auto& t : ch_tkeys) tkeys.insert(tkeys.end(), t.begin(), t.end());
            tkeys.erase(std::unique(tkeys.begin(), tkeys.end()), tkeys.end());
            const double tl0 = now_s();
            literal_std_sort_by_first(kv0.data(), M, threads_, tkeys.data(), tkeys.size(), kvh.data());
            if (getenv("PB200_PROFILE_HOST")) fprintf(stderr, "[pb200 literal sort ms] %.2f (%zu tied keys)\n", (now_s() - tl0) * 1e3, tkeys.size());
            parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
                for (size_t i = (size_t)c * per; i < std::min(M, (size_t)(c + 1) * per); ++i) final_mums_[i] = kv0[i].second;
            });
            final_min_length_ = ml;
            final_sorted_ = false;              // the next call sorts again, like the reference
            return;
        }
    }
    std::vector<std::pair<int64_t, int>> kv(M);
    // gather the keys (random reads into the MUM pools) in parallel; per chunk: descents inside, first/last key, min length
    const long per = 8192;
    const long nch = ((long)M + per - 1) / per;
    std::vector<size_t> ch_desc((size_t)nch, 0), ch_split((size_t)nch, 0);
    std::vector<int64_t> ch_max((size_t)nch, 0), ch_minlen((size_t)nch, INT64_MAX);
    parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
        size_t d = 0, sp = 0;
        int64_t mx = 0, ml = INT64_MAX;
        const size_t i0 = (size_t)c * per, i1 = std::min(M, (size_t)(c + 1) * per);
        for (size_t i = i0; i < i1; ++i) {
            const MumRec& m = mums_[final_mums_[i]];
            const int64_t s0 = mum_start_[m.off];
            kv[i] = std::make_pair(s0, final_mums_[i]);
            if (i > i0 && s0 < kv[i - 1].first) { ++d; sp = i; }
            mx = std::max(mx, s0);
            ml = std::min(ml, m.length);
        }
        ch_desc[c] = d; ch_split[c] = sp; ch_max[c] = mx; ch_minlen[c] = ml;
    });
    size_t split = 0;
    int64_t maxkey = 0;
    final_min_length_ = INT64_MAX;
    for (long c = 0; c < nch; ++c) {
        if (ch_desc[c]) { descents += ch_desc[c]; split = ch_split[c]; }
        if (c > 0 && kv[(size_t)c * per].first < kv[(size_t)c * per - 1].first) { ++descents; split = (size_t)c * per; }
        maxkey = std::max(maxkey, ch_max[c]);
        final_min_length_ = std::min(final_min_length_, ch_minlen[c]);
    }
    // ties on start[0] (two accepted MUMs starting at the same reference position: a trimmed remnant of length 2 next to a
    // later MUM) make the order implementation-defined: the reference calls std::sort (libstdc++ introsort, unstable) with
    // operator< on start[0] (src/TMum.cpp:151) at every one of these calls.  The same algorithm on (start0, id) records in the
    // same initial order makes the same comparisons and moves, hence the same permutation - whatever the element type.
    bool kv_initial = true;                             // kv = the records in the initial order (as gathered above)
    std::vector<int64_t> tkeys;
    static const bool prof_sort = getenv("PB200_PROFILE_HOST") != nullptr;
    auto literal_sort = [&]() {
        const double tl0 = now_s();
        if (!kv_initial)                                // (final_mums_ still holds the initial order)
            parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
                for (size_t i = (size_t)c * per; i < std::min(M, (size_t)(c + 1) * per); ++i) kv[i] = std::make_pair(mum_start_[mums_[final_mums_[i]].off], final_mums_[i]);
            });
        literal_std_sort_by_first(kv.data(), M, threads_, tkeys.data(), tkeys.size());     // = std::sort(kv.begin(), kv.end(), by .first) (parallel.h)
        const double tl1 = now_s();
        parallel_chunks(M > 32768 ? threads_ : 1, nch, [&](long c) {
            for (size_t i = (size_t)c * per; i < std::min(M, (size_t)(c + 1) * per); ++i) final_mums_[i] = kv[i].second;
        });
        final_sorted_ = false;                         // the next call sorts again, like the reference
        if (prof_sort) fprintf(stderr, "[pb200 literal sort ms] %.2f (+ %.2f)\n", (tl1 - tl0) * 1e3, (now_s() - tl1) * 1e3);
    };
    auto has_ties = [&]() {                             // (kv ascending here) + the tied keys for the literal call
        tkeys.clear();
        for (size_t i = 1; i < M; ++i) if (kv[i].first == kv[i - 1].first && (tkeys.empty() || tkeys.back() != kv[i].first)) tkeys.push_back(kv[i].first);
        return !tkeys.empty();
    };
    final_sorted_ = true;
    if (descents == 0) {
        if (has_ties()) literal_sort();
        return;
    }
    std::vector<std::pair<int64_t, int>> tmp(M);
    kv_initial = false;
    if (descents == 1) {
        // the usual shape: the anchors (ascending) followed by the recursion's MUMs (ascending): one merge
        std::merge(kv.begin(), kv.begin() + (long)split, kv.begin() + (long)split, kv.end(), tmp.begin(),
                   [](const std::pair<int64_t, int>& x, const std::pair<int64_t, int>& y) { return x.first < y.first; });
        kv.swap(tmp);
    } else {
        // LSD byte radix sort of (start0, id) pairs: with distinct keys the result is the unique ascending order
        for (int shift = 0; shift < 64 && (maxkey >> shift) != 0; shift += 8) {
            size_t cnt[257] = {0};
            for (size_t i = 0; i < M; ++i) cnt[((uint64_t)kv[i].first >> shift & 0xff) + 1]++;
            for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
            for (size_t i = 0; i < M; ++i) tmp[cnt[(uint64_t)kv[i].first >> shift & 0xff]++] = kv[i];
            kv.swap(tmp);
        }
    }
    if (has_ties()) { literal_sort(); return; }
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
    if (final_min_length_ > rvalue) return;        // no MUM is short enough to be examined by the loop below
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
            stats_.mums_filtered++;
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
    const float dd = prm_.diag_diff;
    if (!(dd > 1.0f)) {
        // Default diagdiff (<= 1): every MUM either joins the open cluster or closes it and opens the next one, so the open cluster's
        // last MUM is always nt-1 and all its members share one orientation vector (joining requires f == F(back) == F(front)).
        // The join decision of the loop below is then a pure function of the pair (nt-1, nt): evaluate the pairs in parallel and
        // cut the sorted MUM list into runs.  Any MUM shorter than `random` (skipped by the loop) sends us to the literal loop.
        std::vector<uint8_t> joins((size_t)M, 0);
        std::vector<int64_t> sl((size_t)M);
        std::atomic<int> short_mum(0);
        const long per = 2048;
        parallel_chunks(M > 16384 ? threads_ : 1, ((long)M + per - 1) / per, [&](long c) {
            const long i0 = c * per, i1 = std::min<long>((long)M, (c + 1) * per);
            for (long i = i0; i < i1; ++i) {
                const MumRec& m = mums_[final_mums_[i]];
                sl[i] = m.length;
                if (m.length < prm_.random) short_mum.store(1, std::memory_order_relaxed);
                if (i == 0) continue;
                const MumRec& b = mums_[final_mums_[i - 1]];
                const int64_t* sn = &mum_start_[m.off];
                const int64_t* sb = &mum_start_[b.off];
                const uint8_t* fn = &mum_fwd_[m.off];
                const uint8_t* fb = &mum_fwd_[b.off];
                float max_length_region = 0, min_length_region = (float)(prm_.d + 10);
                bool add = true;
                for (int k = 0; k < n_; ++k) {
                    const int f = fn[k];
                    const int64_t gap = sn[k] - (sb[k] + b.length);
                    const int64_t rgap = sb[k] - (sn[k] + m.length);
                    if (f && (float)gap > max_length_region) max_length_region = (float)gap;
                    else if (!f && (float)rgap > max_length_region) max_length_region = (float)gap;   // sic (src/parsnp.cpp:2608-2611)
                    if (f && (float)gap < min_length_region) min_length_region = (float)gap;
                    else if (!f && (float)rgap < min_length_region) min_length_region = (float)rgap;
                    if (f != (int)fb[k]) add = false;
                    else if (f && gap < 0) add = false;
                    else if (!f && gap >= 0) add = false;
                    else if (f && gap > prm_.d) add = false;
                    else if (!f && rgap > prm_.d) add = false;
                    if (!add) break;
                }
                if (add) {
                    if (min_length_region == 0) min_length_region = 1;
                    if (max_length_region == 0) max_length_region = 1;
                    add = min_length_region / max_length_region >= 1.0 - dd;
                }
                joins[i] = add ? 1 : 0;
            }
        });
        if (!short_mum.load()) {
            for (int64_t a = 0; a < M;) {
                int64_t b = a + 1;
                while (b < M && joins[b]) ++b;
                out.emplace_back();
                ClusterRec& c = out.back();
                c.type = 1;
                c.length = 0;
                c.mums.resize((size_t)(b - a));
                for (int64_t i = a; i < b; ++i) { c.length += sl[i]; c.mums[(size_t)(i - a)] = (int)i; }
                const MumRec& first = mums_[final_mums_[a]];
                const MumRec& last = mums_[final_mums_[b - 1]];
                c.start.assign(&mum_start_[first.off], &mum_start_[first.off] + n_);
                c.end.resize(n_);
                for (int k = 0; k < n_; ++k) c.end[k] = mum_start_[last.off + k] + last.length;
                a = b;
            }
            return;
        }
    }
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
            stats_.clusters_filtered++;
            stats_.mums_filtered += (int64_t)cl[cs].mums.size();
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

// ------------------------------------------------------------------ setUnalignableRegions (src/parsnp.cpp:2310-2382)
// The reference walks mumlayout bit by bit, round-robin over the genomes: in every round each genome contributes its next run
// of clear bits [startpos, endpos] (setting them on the way); a record is written when startpos != endpos (a single clear
// bit is skipped), and the walk stops the first time the LAST genome has no run left - whatever the others still hold.
// Restated with word-level scans on the final layout (runs are found, not set: every round resumes behind its last run).
void Aligner::unaligned_regions(std::vector<int32_t>& genome, std::vector<int64_t>& start, std::vector<int64_t>& end) const {
    genome.clear(); start.clear(); end.clear();
    std::vector<int64_t> lastpos((size_t)n_, 0);
    std::vector<uint8_t> exhausted((size_t)n_, 0);           // a trailing run without a closing set bit is re-scanned as all ones
    for (bool stop = false; !stop;) {
        for (int k = 0; k < n_; ++k) {
            const BitRow& row = truth_.layout[k];
            const int64_t size = len_[k] + 1;                // mumlayout[k].size(), bit len_[k] is the sentinel
            int64_t startpos = -1, endpos = -1;
            if (!exhausted[k]) {
                // first clear bit at or after lastpos
                int64_t m = lastpos[k];
                while (m < size && row.get(m)) {
                    const int64_t r = row.run_up(m, size);
                    m += r ? r : 1;
                }
                if (m < size) {
                    startpos = m;
                    const int64_t nx = row.next_set(m, size);    // first set bit behind the run (size: none)
                    endpos = nx - 1;
                    if (nx < size) lastpos[k] = endpos + 1;
                    else exhausted[k] = 1;                        // (cannot happen: the sentinel closes every run)
                } else {
                    exhausted[k] = 1;
                }
            }
            if (startpos != endpos) { genome.push_back(k); start.push_back(startpos); end.push_back(endpos); }
            else if (startpos == -1 && k == n_ - 1) stop = true;
        }
    }
}

// ------------------------------------------------------------------ recursion discovered by the engine (cuda/recursion.cuh)
namespace {
// minsize(slength) of an expression for slength < N, computed once per process and expression (the calculator is a string
// interpreter: ~0.2 us per value)
std::shared_ptr<const std::vector<int32_t>> minsize_table(const std::string& expr, int N, int threads) {
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<const std::vector<int32_t>>> cache;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(expr);
        if (it != cache.end()) return it->second;
    }
    std::shared_ptr<std::vector<int32_t>> tab(new std::vector<int32_t>((size_t)N, 0));
    const MinSizeExpr ex(expr);
    std::atomic<int> bad(0);
    const long per = 256;
    parallel_chunks(threads, (N + per - 1) / per, [&](long c) {
        for (long sl = c * per; sl < std::min<long>(N, (c + 1) * per); ++sl) {
            try { (*tab)[(size_t)sl] = ex(sl); } catch (...) { bad.store(1); }
        }
    });
    if (bad.load()) return nullptr;              // (an expression that fails for some length: left to the path that evaluates on demand)
    std::lock_guard<std::mutex> lk(mu);
    cache[expr] = tab;
    return tab;
}
}  // namespace

// The initial regions are cut into a few slices in reference order; the first one is followed here, the others on their own
// thread while the replay already consumes what is published (every task waits for the slice of its regions only).
bool Aligner::discover_on_device() {
    const size_t R = initial_regions_.size();
    const int TABN = 4160;                       // covers every window the shared-memory search kernel takes (<= 4096 bases)
    disc_tab_ = minsize_table(prm_.mums, TABN, threads_);
    if (!disc_tab_) return false;
    const long per = 4096;
    if (!anchors_on_device_) {                   // (else the engine already holds them: it made them)
        disc_coords_.resize(R * 2 * (size_t)n_);
        parallel_chunks(R > 16384 ? threads_ : 1, ((long)R + per - 1) / per, [&](long c) {
            for (size_t i = (size_t)c * per; i < std::min(R, (size_t)(c + 1) * per); ++i)
                std::memcpy(&disc_coords_[i * 2 * (size_t)n_], rstart(initial_regions_[i]), sizeof(int64_t) * 2 * (size_t)n_);
        });
    }
    // slices: a small first one (the replay starts as soon as it is there), then growing
    const char* es = getenv("PB200_DISCOVERY_SLICES");
    size_t K = es ? (size_t)std::max(1, atoi(es)) : 1;
    if (anchors_on_device_) K = 1;               // (the engine is already following the whole recursion)       // (measured on configs[1]: slices cost more in launches and syncs than the overlap returns)
    K = std::min(K, R);
    disc_begin_.assign(K + 1, 0);
    {
        double tot = 0, run = 0;
        for (size_t k = 0; k < K; ++k) tot += 1.0 + (double)k;
        for (size_t k = 0; k < K; ++k) { run += 1.0 + (double)k; disc_begin_[k + 1] = (size_t)((double)R * run / tot); }
        disc_begin_[K] = R;
    }
    slice_cache_.clear();
    for (size_t k = 0; k < K; ++k) slice_cache_.emplace_back(new CandCache);
    slice_of_initial_.assign(R, 0);
    for (size_t k = 0; k < K; ++k)
        for (size_t i = disc_begin_[k]; i < disc_begin_[k + 1]; ++i) slice_of_initial_[i] = (int)k;
    slices_ready_ = 0;
    stats_.spec_slices = (int64_t)K;
    if (!discover_slice(0)) { slice_cache_.clear(); slice_of_initial_.clear(); stats_.spec_slices = 0; return false; }
    {
        std::lock_guard<std::mutex> lk(slice_mu_);
        slices_ready_ = 1;
    }
    if (K > 1)
        spec_thread_ = std::thread([this, K] {
            try {
                for (size_t k = 1; k < K; ++k) {
                    discover_slice((int)k);               // (false = left empty: the replay searches those regions on demand)
                    {
                        std::lock_guard<std::mutex> lk(slice_mu_);
                        slices_ready_ = (int)k + 1;
                    }
                    slice_cv_.notify_all();
                }
            } catch (...) {
                std::lock_guard<std::mutex> lk(slice_mu_);
                spec_error_ = std::current_exception();
                slices_ready_ = (int)K;                    // nobody waits for ever; run() rethrows
                slice_cv_.notify_all();
            }
        });
    return true;
}

bool Aligner::discover_slice(int k) {
    const double t0 = now_s();
    const size_t i0 = disc_begin_[(size_t)k], R = disc_begin_[(size_t)k + 1] - i0;
    const long per = 4096;
    std::vector<const uint64_t*> rows((size_t)n_);
    std::vector<int64_t> nwords((size_t)n_);
    for (int g = 0; g < n_; ++g) { rows[(size_t)g] = truth_.layout[(size_t)g].words(); nwords[(size_t)g] = truth_.layout[(size_t)g].nwords(); }
    RecursionRequest rq;
    rq.n = n_; rq.coords = disc_coords_.data() + i0 * 2 * (size_t)n_; rq.nregions = (int)R; rq.layout = rows.data(); rq.layout_words = nwords.data();
    rq.upload_layout = k == 0;
    rq.resume = anchors_on_device_;                   // (later slices run beside the replay, which writes the layout: the engine keeps its scratch copy)
    rq.q = prm_.q; rq.p = prm_.p; rq.minsize_tab = disc_tab_->data(); rq.minsize_n = (int)disc_tab_->size();
    RecursionResult res;
    if (R == 0) return true;
    {
        std::lock_guard<std::mutex> lk(backend_mu_);
        if (!be_->discover_recursion(rq, res)) return false;
    }
    const double t1 = now_s();
    // ---- the result as a candidate cache: one window per region; coordinates, window records, hashes and candidates come from the
    // engine in their final form (views into its pinned staging memory; with several slices they are copied, the next slice
    // overwrites them)
    CandCache& C = *slice_cache_[(size_t)k];
    const size_t NR = res.nregions;
    const bool copy = slice_cache_.size() > 1;
    C.rp.n = n_;
    C.entries.resize(NR);
    C.wins.resize(NR);
    C.chunks.emplace_back();
    CandBatch& cb = C.chunks.back();
    cb.nq = n_ - 1;
    cb.off.assign(1, 0);
    const size_t NC = res.ncands;
    if (copy) {
        C.rp.coord.assign(res.coords, res.coords + NR * 2 * (size_t)n_);
        cb.k.assign(res.k, res.k + NC); cb.lon.assign(res.lon, res.lon + NC);
        cb.sp.assign(res.sp, res.sp + NC * (size_t)(n_ - 1)); cb.fwd.assign(res.fwd, res.fwd + NC * (size_t)(n_ - 1));
    } else {
        C.rp.ext_coord = res.coords;
        cb.vk = res.k; cb.vlon = res.lon; cb.vsp = res.sp; cb.vfwd = res.fwd; cb.vcount = NC;
    }
    dev_ = DeviceDecisions();
    if (!copy && res.flags && res.parent && res.acc_shift && res.acc_len && res.dropped == 0 && res.nfw <= res.fw_cap && !getenv("PB200_NO_DEVICE_FINAL")) {
        dev_.valid = true;
        dev_.nregions = NR; dev_.ncands = NC; dev_.nfw = res.nfw;
        dev_.coords = res.coords; dev_.slen = res.slen; dev_.wins = res.wins; dev_.k = res.k; dev_.lon = res.lon; dev_.sp = res.sp; dev_.fwd = res.fwd;
        dev_.flags = res.flags; dev_.parent = res.parent; dev_.acc_shift = res.acc_shift; dev_.acc_len = res.acc_len; dev_.fw = res.fw;
    }
    // (the index over the regions' coordinates: built right away, or - when the replay may take most gaps from the engine as final -
    //  after the gaps have been classified, over the regions it will really look up: build_region_index)
    std::vector<uint8_t> valid_local;
    std::vector<uint64_t> hashes_local;
    const bool defer_index = dev_.valid && k == 0;
    std::vector<uint8_t>& valid = defer_index ? disc_valid_ : valid_local;
    std::vector<uint64_t>& hashes = defer_index ? disc_hashes_ : hashes_local;
    valid.assign(NR, 0);
    hashes.assign(res.hashes, res.hashes + NR);
    const long nblk = ((long)NR + per - 1) / per;
    std::vector<int64_t> blk_searched((size_t)nblk + 1, 0), blk_cands((size_t)nblk + 1, 0);
    parallel_chunks(NR > 16384 ? threads_ : 1, nblk, [&](long c) {
        int64_t ns = 0, nc = 0;
        for (size_t r = (size_t)c * per; r < std::min(NR, (size_t)(c + 1) * per); ++r) {
            CacheEntry e; e.region = (int)r; e.first_win = (int64_t)r; e.nwin = 1;
            C.entries[r] = e;
            WinRec w = res.wins[r];
            if (w.ncand < 0 || (uint64_t)w.cand_off + (uint64_t)w.ncand > NC) w.ncand = -1;       // not searched (or its candidates did not fit)
            valid[r] = w.ncand >= 0;
            if (!valid[r]) w.ncand = 0;                               // stays out of the index: the replay searches it on demand
            else { ++ns; nc += w.ncand; }
            C.wins[r] = w;
        }
        blk_searched[(size_t)c] = ns; blk_cands[(size_t)c] = nc;
    });
    int64_t searched = 0, cands = 0;
    for (long c = 0; c < nblk; ++c) { searched += blk_searched[(size_t)c]; cands += blk_cands[(size_t)c]; }
    disc_index_deferred_ = defer_index;
    C.sorted_valid = defer_index ? &disc_valid_ : nullptr;  // (entry r = region r of the engine's list, ascending start[0])
    if (!defer_index) C.map.build_parallel(hashes, valid, threads_);
    std::lock_guard<std::mutex> lk(backend_mu_);           // (the statistics are shared with the replay's on-demand searches)
    stats_.spec_regions += searched;
    stats_.regions_searched += searched;
    stats_.windows_searched += searched;
    stats_.candidates += cands;
    stats_.spec_levels += res.levels;
    stats_.spec_deferred += res.deferred + res.dropped;
    stats_.t_spec_search += t1 - t0;
    stats_.t_spec_host += now_s() - t1;
    return true;
}

// ------------------------------------------------------------------ main sequence (src/parsnp.cpp:3187-3273)
bool Aligner::run() {
    double t0 = now_s();
    set_initial_clusters();
    if (!prm_.anchors_only) {
        stats_.host_threads = threads_;
        if (speculate_ && !initial_regions_.empty() && pipeline_ && !getenv("PB200_NO_DEVICE_RECURSION")) replay_prepare_async();
        if (speculate_ && !initial_regions_.empty() && pipeline_ && discover_on_device()) {
            // the engine followed the recursion itself (cuda/recursion.cuh): one published "slice" holds every predicted region
        } else if (speculate_ && !initial_regions_.empty()) {
            // slices of the initial regions (reference order): the speculation thread discovers and searches slice k+1 while the
            // replay below consumes slice k
            const char* es = getenv("PB200_SPEC_SLICES");
            size_t K = es ? (size_t)std::max(1, atoi(es)) : std::min<size_t>(16, initial_regions_.size() / 512 + 1);
            K = std::min(K, initial_regions_.size());
            stats_.spec_slices = (int64_t)K;
            // private copies for the speculation thread (the replay appends to rp_ and sets bits in truth_): parallel memcpy
            frozen_rp_.n = n_;
            frozen_rp_.coord.resize(rp_.coord.size());
            frozen_rp_.slen.resize(rp_.slen.size());
            spec_world_.layout.resize(truth_.layout.size());
            {
                const size_t CH = (size_t)1 << 20, cb = rp_.coord.size() * sizeof(int64_t), sb = rp_.slen.size() * sizeof(int64_t);
                const long nc = (long)((cb + CH - 1) / CH), ns = (long)((sb + CH - 1) / CH), nl = (long)truth_.layout.size();
                parallel_chunks(threads_, nc + ns + nl, [&](long c) {
                    if (c < nc) std::memcpy((char*)frozen_rp_.coord.data() + (size_t)c * CH, (const char*)rp_.coord.data() + (size_t)c * CH, std::min(CH, cb - (size_t)c * CH));
                    else if (c < nc + ns) std::memcpy((char*)frozen_rp_.slen.data() + (size_t)(c - nc) * CH, (const char*)rp_.slen.data() + (size_t)(c - nc) * CH,
                                                       std::min(CH, sb - (size_t)(c - nc) * CH));
                    else spec_world_.layout[(size_t)(c - nc - ns)] = truth_.layout[(size_t)(c - nc - ns)];
                });
            }
            slice_regions_.assign(K, std::vector<int>());
            slice_of_initial_.resize(initial_regions_.size());
            for (size_t i = 0; i < initial_regions_.size(); ++i) {
                const size_t k = i * K / initial_regions_.size();
                slice_regions_[k].push_back(initial_regions_[i]);
                slice_of_initial_[i] = (int)k;
            }
            slice_cache_.clear();
            for (size_t k = 0; k < K; ++k) slice_cache_.emplace_back(new CandCache);
            slices_ready_ = 0;
            if (pipeline_ && !getenv("PB200_SPEC_SYNC")) spec_thread_ = std::thread([this] { speculation_thread_main(); });   // (PB200_SPEC_SYNC: tools/host_bench.py times the replay alone)
            else {
                // lock step (collectives inside the search): an error must surface HERE, on every rank at the same call - a rank
                // that went on with a half-filled cache would issue searches, i.e. collectives, that the others never join
                speculation_thread_main();
                if (spec_error_) std::rethrow_exception(spec_error_);
            }
        }
        try {
            do_work_exact();
        } catch (...) {
            if (spec_thread_.joinable()) spec_thread_.join();
            throw;
        }
        if (spec_thread_.joinable()) spec_thread_.join();
        if (spec_error_) std::rethrow_exception(spec_error_);
    }
    if (all_mums_.empty()) { stats_.t_total = now_s() - t0; return false; }
    double t1 = now_s();
    final_mums_ = all_mums_;
    final_sorted_ = false;
    if (sorted_hint_.size() != all_mums_.size()) sorted_hint_.clear();
    const bool prof = getenv("PB200_PROFILE_HOST") != nullptr;
    double tl[6] = {now_s(), 0, 0, 0, 0, 0};
    if (prm_.random) filter_random1();
    tl[1] = now_s();
    set_final_clusters(clusters_);
    tl[2] = now_s();
    filter_clusters_simple(clusters_);
    tl[3] = now_s();
    set_final_clusters(clusters_);
    tl[4] = now_s();
    set_inter_cluster_regions(clusters_);
    tl[5] = now_s();
    if (prof) fprintf(stderr, "[pb200 lcb ms] copy %.2f filter_random1 %.2f set_final %.2f filter_simple %.2f set_final %.2f inter %.2f\n", (tl[0] - t1) * 1e3,
                      (tl[1] - tl[0]) * 1e3, (tl[2] - tl[1]) * 1e3, (tl[3] - tl[2]) * 1e3, (tl[4] - tl[3]) * 1e3, (tl[5] - tl[4]) * 1e3);
    stats_.t_lcb = now_s() - t1;
    stats_.t_total = now_s() - t0;
    return true;
}

}  // namespace pb200
