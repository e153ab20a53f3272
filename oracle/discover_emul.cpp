// oracle/discover_emul.cpp - TEST INFRASTRUCTURE ONLY: a CPU restatement of the engine's device-resident discovery of the
// recursion (parsnp_b200/csrc/cuda/recursion.cuh: recursion_level_kernel + accept_region + seed_lists_kernel and the sorted
// delivery of CudaEngine::rec_finish) behind SearchBackend::discover_recursion, over the csgmum search of oracle/ref_backend.cpp.
//
// The product's parallel replay (parsnp_b200/csrc/host/replay.cpp) takes the engine's accept decisions as FINAL for the gaps
// where they cannot depend on the order.  That logic lives on the host but needs a backend that delivers such decisions; with
// this one the CPU fuzzers (tools/fuzz_replay.py, tools/fuzz_host.py with PB200_TEST_BACKEND=3) exercise it without a GPU:
// level by level, regions of a level in RANDOM order (the device's CTAs take them in no particular order), the two regions
// of a gap pair in the reference's order, decisions on a scratch copy of mumlayout, flags / parents / accepted shifts and
// lengths / writes outside the region reported exactly as the kernel does.  Never linked into libparsnp_b200.so.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <vector>
#include "../parsnp_b200/csrc/common.h"

namespace pb200_oracle {
pb200::SearchBackend* make_ref_backend();

namespace {
struct Bits {
    std::vector<uint64_t> w;
    bool get(int64_t i) const { return (w[(size_t)(i >> 6)] >> (i & 63)) & 1ull; }
    void set(int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) w[(size_t)(i >> 6)] |= 1ull << (i & 63); }
    int64_t run_up(int64_t a, int64_t b) const { int64_t i = a; while (i < b && get(i)) ++i; return i - a; }
    int64_t run_down(int64_t a, int64_t b) const { int64_t i = b; while (i > a && get(i - 1)) --i; return b - i; }
    int64_t prev_set(int64_t i) const { while (i >= 0 && !get(i)) --i; return i; }
    int64_t next_set(int64_t i, int64_t limit) const { while (i < limit && !get(i)) ++i; return i; }
};
struct Reg {
    std::vector<int64_t> S, E;
    int64_t slen = 0;
    int minsize = 0, parent = -1, ncand = -1;
    uint32_t flags = 0;
    std::vector<int32_t> k, lon, sp, shift, alen;
    std::vector<uint8_t> fwd;
};
struct Entry { int id; bool pair, second_first; };
}  // namespace

class DiscoverEmulBackend : public pb200::SearchBackend {
public:
    DiscoverEmulBackend() : inner_(make_ref_backend()) {}
    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override {
        n_ = n; seq_.assign(seq, seq + n); len_.assign(len, len + n);
        inner_->set_genomes(n, seq, len);
    }
    void search(const pb200::WindowTask* tasks, int ntasks, const int64_t* coords, pb200::CandBatch& out) override { inner_->search(tasks, ntasks, coords, out); }

    bool discover_recursion(const pb200::RecursionRequest& rq, pb200::RecursionResult& out) override {
        if (rq.resume || rq.n != n_ || rq.nregions <= 0 || !rq.upload_layout) return false;
        const int n = n_, nq = n - 1;
        const char* ec = getenv("PB200_EMUL_MAXLEN");            // the device's largest window class (longer regions are left to the host)
        const int64_t maxlen = ec ? atoll(ec) : 4096;
        const char* es = getenv("PB200_EMUL_SEED");
        std::mt19937 rng(es ? (unsigned)atoi(es) : 12345u);
        std::vector<Bits> bits((size_t)n);
        for (int g = 0; g < n; ++g) bits[(size_t)g].w.assign(rq.layout[g], rq.layout[g] + rq.layout_words[g]);
        regs_.clear();
        fw_.clear();
        auto classify = [&](Reg& r) {            // rec::class_of: true = the device takes it
            int64_t sl = 500000000;
            for (int g = 0; g < n; ++g) sl = std::min(sl, r.E[(size_t)g] - r.S[(size_t)g]);
            r.slen = sl;
            r.minsize = sl >= 0 && sl < rq.minsize_n ? rq.minsize_tab[sl] : 0;
            const int64_t L0 = r.E[0] - r.S[0];
            return !(r.minsize < 4 || L0 > rq.p || L0 <= 0 || L0 > maxlen);
        };
        std::vector<Entry> cur, next;
        const size_t R = (size_t)rq.nregions;
        std::vector<uint8_t> ok(R);
        for (size_t r = 0; r < R; ++r) {
            Reg x;
            const int64_t* s = rq.coords + r * 2 * (size_t)n;
            x.S.assign(s, s + n); x.E.assign(s + n, s + 2 * n);
            ok[r] = classify(x);
            regs_.push_back(std::move(x));
        }
        for (size_t r = 0; r < R; ++r) {         // seed_lists_kernel: (right side of anchor i, left side of anchor i+1) = one entry, the second first
            const int64_t* s = rq.coords + r * 2 * (size_t)n;
            const bool head = r + 1 < R && s[2 * n] == s[0] - 1 && s[2 * n + n] == s[n];
            if (head && ok[r] && ok[r + 1]) { cur.push_back(Entry{(int)r, true, true}); ++r; continue; }
            if (ok[r]) cur.push_back(Entry{(int)r, false, false});
        }
        int levels = 0;
        while (!cur.empty()) {
            std::shuffle(cur.begin(), cur.end(), rng);
            next.clear();
            for (const Entry& e : cur) {
                const int first = e.id + (e.pair && e.second_first ? 1 : 0);
                const int second = e.pair ? e.id + (e.second_first ? 0 : 1) : -1;
                process(rq, bits, first, false, next);
                if (second >= 0) { regs_[(size_t)second].flags |= pb200::REC_SECOND; process(rq, bits, second, true, next); }
            }
            cur.swap(next);
            ++levels;
        }
        // ---- delivery: ascending start[0] (stable), candidates regrouped, parents as sorted positions
        const size_t NR = regs_.size();
        std::vector<int> perm(NR), inv(NR);
        for (size_t i = 0; i < NR; ++i) perm[i] = (int)i;
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return regs_[(size_t)a].S[0] < regs_[(size_t)b].S[0]; });
        for (size_t i = 0; i < NR; ++i) inv[(size_t)perm[i]] = (int)i;
        o_coords_.assign(NR * 2 * (size_t)n, 0); o_slen_.assign(NR, 0); o_hash_.assign(NR, 0); o_wins_.assign(NR, pb200::WindowRec());
        o_flags_.assign(NR, 0); o_parent_.assign(NR, -1);
        o_k_.clear(); o_lon_.clear(); o_sp_.clear(); o_fwd_.clear(); o_shift_.clear(); o_alen_.clear();
        int64_t searched = 0, deferred = 0;
        for (size_t i = 0; i < NR; ++i) {
            const Reg& r = regs_[(size_t)perm[i]];
            int64_t* c = &o_coords_[i * 2 * (size_t)n];
            for (int g = 0; g < n; ++g) { c[g] = r.S[(size_t)g]; c[n + g] = r.E[(size_t)g]; }
            o_slen_[i] = r.slen;
            o_hash_[i] = pb200::region_coords_hash(c, 2 * n);
            pb200::WindowRec w;
            w.ref_start = r.S[0]; w.ref_len = r.E[0] - r.S[0]; w.cand_off = (int64_t)o_k_.size(); w.ncand = r.ncand; w.chunk = 0;
            o_wins_[i] = w;
            o_flags_[i] = r.flags;
            o_parent_[i] = r.parent < 0 ? -1 : inv[(size_t)r.parent];
            if (r.ncand < 0) { ++deferred; continue; }
            ++searched;
            o_k_.insert(o_k_.end(), r.k.begin(), r.k.end()); o_lon_.insert(o_lon_.end(), r.lon.begin(), r.lon.end());
            o_sp_.insert(o_sp_.end(), r.sp.begin(), r.sp.end()); o_fwd_.insert(o_fwd_.end(), r.fwd.begin(), r.fwd.end());
            o_shift_.insert(o_shift_.end(), r.shift.begin(), r.shift.end()); o_alen_.insert(o_alen_.end(), r.alen.begin(), r.alen.end());
        }
        (void)nq;
        out = pb200::RecursionResult();
        out.nregions = NR; out.ncands = o_k_.size();
        out.coords = o_coords_.data(); out.slen = o_slen_.data(); out.hashes = o_hash_.data(); out.wins = o_wins_.data();
        out.k = o_k_.data(); out.lon = o_lon_.data(); out.sp = o_sp_.data(); out.fwd = o_fwd_.data();
        out.flags = o_flags_.data(); out.parent = o_parent_.data(); out.acc_shift = o_shift_.data(); out.acc_len = o_alen_.data();
        out.fw = fw_.data(); out.nfw = fw_.size() / 3; out.fw_cap = 65536;
        out.levels = levels; out.deferred = deferred; out.dropped = 0; out.searched = searched;
        return true;
    }

private:
    // one region: search (small_window), then rec::accept_region on the scratch layout
    void process(const pb200::RecursionRequest& rq, std::vector<Bits>& bits, int id, bool second, std::vector<Entry>& next) {
        const int n = n_, nq = n - 1;
        {
            Reg& r = regs_[(size_t)id];
            pb200::WindowTask t;
            t.ref_start = r.S[0]; t.ref_len = r.E[0] - r.S[0]; t.coord_off = 0; t.minsize = r.minsize; t.pad = 0;
            std::vector<int64_t> coords;
            for (int g = 1; g < n; ++g) coords.push_back(r.S[(size_t)g]);
            for (int g = 1; g < n; ++g) coords.push_back(r.E[(size_t)g] - r.S[(size_t)g]);
            pb200::CandBatch cb;
            inner_->search(&t, 1, coords.data(), cb);
            cb.compact(1);
            r.k.assign(cb.k.begin(), cb.k.end()); r.lon.assign(cb.lon.begin(), cb.lon.end());
            r.sp.assign(cb.sp.begin(), cb.sp.end()); r.fwd.assign(cb.fwd.begin(), cb.fwd.end());
            r.ncand = (int)r.k.size();
            r.shift.assign(r.k.size(), -1); r.alen.assign(r.k.size(), 0);
        }
        const int nc = regs_[(size_t)id].ncand;
        std::vector<int> acc;
        std::vector<std::vector<int64_t>> acc_st;
        uint32_t rflags = 0;
        std::vector<int64_t> st((size_t)n), prev_end;
        std::vector<uint8_t> fw((size_t)n);
        for (int c = 0; c < nc; ++c) {
            const Reg& r = regs_[(size_t)id];
            const int64_t LON = r.lon[(size_t)c];
            bool fail = false, rev = false;
            for (int g = 0; g < n; ++g) {
                const int64_t off = g == 0 ? r.k[(size_t)c] : r.sp[(size_t)c * nq + (g - 1)];
                fw[(size_t)g] = g == 0 ? 1 : r.fwd[(size_t)c * nq + (g - 1)];
                const int64_t rl = r.E[(size_t)g] - r.S[(size_t)g];
                if ((uint64_t)(off + 1) > (uint64_t)(uint32_t)rl) fail = true;
                int64_t s = r.S[(size_t)g] + off;
                if (!fw[(size_t)g]) s = len_[(size_t)g] - (s + LON);
                if (s + LON > len_[(size_t)g] || s < 0) fail = true;
                st[(size_t)g] = s;
                rev |= !fw[(size_t)g];
            }
            if (fail || LON < 5) continue;
            if (rev) rflags |= 1;
            int64_t length = LON, shift = 0;
            for (int g = 0; g < n && length > 0; ++g) {
                const int64_t t1 = bits[(size_t)g].run_up(st[(size_t)g] + shift, st[(size_t)g] + shift + length);
                shift += t1; length -= t1;
                const int64_t t2 = bits[(size_t)g].run_down(st[(size_t)g] + shift, st[(size_t)g] + shift + length);
                length -= t2;
            }
            if (length < 2) continue;
            bool badmum = false;
            for (int g = 1; g < n && !badmum; ++g) {
                if (fw[(size_t)g]) continue;
                const uint8_t* g0 = seq_[0] + st[0] + shift;
                const uint8_t* gk = seq_[(size_t)g] + st[(size_t)g] + shift;
                for (int64_t x = 0; x < length && !badmum; ++x) {
                    const uint8_t a = gk[length - 1 - x];
                    const uint8_t cm = a == 'A' ? 'T' : a == 'T' ? 'A' : a == 'C' ? 'G' : a == 'G' ? 'C' : 'N';
                    badmum = cm != g0[x];
                }
            }
            if (badmum) continue;
            for (int g = 0; g < n; ++g) {
                bits[(size_t)g].set(st[(size_t)g] + shift, st[(size_t)g] + shift + length);
                if (!fw[(size_t)g]) { fw_.push_back(g); fw_.push_back((int32_t)(st[(size_t)g] + shift)); fw_.push_back((int32_t)length); }
            }
            if (!acc.empty()) {
                bool bad = false;
                for (int g = 0; g < n; ++g) {
                    // (the kernel computes the previous end from the FORWARD formula whatever the strand: such regions are flagged anyway)
                    const int pc = acc.back();
                    const int64_t poff = g == 0 ? r.k[(size_t)pc] : r.sp[(size_t)pc * nq + (g - 1)];
                    const int64_t pend = r.S[(size_t)g] + poff + r.shift[(size_t)pc] + r.alen[(size_t)pc];
                    if (st[(size_t)g] + shift < pend) bad = true;
                }
                if (bad) rflags |= 4;
            }
            Reg& rw = regs_[(size_t)id];
            rw.shift[(size_t)c] = (int32_t)shift; rw.alen[(size_t)c] = (int32_t)length;
            acc.push_back(c);
            std::vector<int64_t> fin((size_t)n);
            for (int g = 0; g < n; ++g) fin[(size_t)g] = st[(size_t)g] + shift;
            acc_st.push_back(std::move(fin));
        }
        if (second && !acc.empty()) rflags |= 2;
        regs_[(size_t)id].flags |= rflags;
        if (acc.empty()) return;
        // determineRegion around every accepted MUM after ALL accepts of the region; (left side of MUM a, right side of MUM a-1) = a pair
        bool have_p = false;
        Reg pend;
        auto push_one = [&](Reg& x) {
            x.parent = id;
            const bool take = classify_child(rq, x);
            const int nid = (int)regs_.size();
            regs_.push_back(x);
            if (take) next.push_back(Entry{nid, false, false});
        };
        for (size_t a = 0; a < acc.size(); ++a) {
            const int64_t length = regs_[(size_t)id].alen[(size_t)acc[a]];
            Reg L, Rr;
            L.S.resize((size_t)n); L.E.resize((size_t)n); Rr.S.resize((size_t)n); Rr.E.resize((size_t)n);
            int64_t lsl = 500000000, rsl = 500000000;
            for (int g = 0; g < n; ++g) {
                const int64_t s = acc_st[a][(size_t)g];
                int64_t cp = bits[(size_t)g].prev_set(s - 1);
                if (cp < 0) cp = 0;
                L.S[(size_t)g] = cp + 1; L.E[(size_t)g] = s - 1;
                const int64_t en = s + length;
                int64_t cq = en + 1;
                if (cq < len_[(size_t)g]) cq = bits[(size_t)g].next_set(cq, len_[(size_t)g]);
                Rr.S[(size_t)g] = en + 1; Rr.E[(size_t)g] = cq - 1;
                lsl = std::min(lsl, L.E[(size_t)g] - L.S[(size_t)g]);
                rsl = std::min(rsl, Rr.E[(size_t)g] - Rr.S[(size_t)g]);
            }
            const bool pl = lsl > rq.q, pp = have_p;
            if (pl && pp) {
                L.parent = id; pend.parent = id;
                const bool ta = classify_child(rq, L), tb = classify_child(rq, pend);
                const int nid = (int)regs_.size();
                regs_.push_back(L);
                regs_.push_back(pend);
                if (ta && tb) next.push_back(Entry{nid, true, false});
                else { if (ta) next.push_back(Entry{nid, false, false}); if (tb) next.push_back(Entry{nid + 1, false, false}); }
            } else if (pl) push_one(L);
            else if (pp) push_one(pend);
            have_p = rsl > rq.q;
            pend = Rr;
        }
        if (have_p) push_one(pend);
    }
    bool classify_child(const pb200::RecursionRequest& rq, Reg& r) const {
        const char* ec = getenv("PB200_EMUL_MAXLEN");
        const int64_t maxlen = ec ? atoll(ec) : 4096;
        int64_t sl = 500000000;
        for (int g = 0; g < n_; ++g) sl = std::min(sl, r.E[(size_t)g] - r.S[(size_t)g]);
        r.slen = sl;
        r.minsize = sl >= 0 && sl < rq.minsize_n ? rq.minsize_tab[sl] : 0;
        const int64_t L0 = r.E[0] - r.S[0];
        return !(r.minsize < 4 || L0 > rq.p || L0 <= 0 || L0 > maxlen);
    }

    std::unique_ptr<pb200::SearchBackend> inner_;
    int n_ = 0;
    std::vector<const uint8_t*> seq_;
    std::vector<int64_t> len_;
    std::vector<Reg> regs_;
    std::vector<int32_t> fw_;
    std::vector<int64_t> o_coords_, o_slen_;
    std::vector<uint64_t> o_hash_;
    std::vector<pb200::WindowRec> o_wins_;
    std::vector<uint32_t> o_flags_;
    std::vector<int32_t> o_parent_, o_k_, o_lon_, o_sp_, o_shift_, o_alen_;
    std::vector<uint8_t> o_fwd_;
};

pb200::SearchBackend* make_discover_backend() { return new DiscoverEmulBackend(); }

}  // namespace pb200_oracle
