// setMums1 loop D (validation, TMum constructor, trim, reverse-strand check, mumlayout update: src/parsnp.cpp:1713-1842,
// 1399-1477, src/TMum.cpp:13-72) and determineRegion (src/parsnp.cpp:1199-1290), written once over a layout ACCESS POLICY:
//   DirectAccess   one bitmap (anchors, the sequential replay, the speculative passes)
//   TaskAccess     a replay task's view of the shared bitmaps (replay.cpp): its own span directly, everything else through
//                  the ownership rules that keep the parallel replay identical to the reference's sequential order
// A policy provides  get / run_up / run_down / prev_set / next_set  per genome and  commit(st, length)  = "set the accepted
// MUM's bits in every genome" (fw: its strand per genome).
#pragma once
#include "aligner.h"

namespace pb200 {

inline uint8_t comp_base(uint8_t c) {          // Aligner::reversec (src/parsnp.cpp:1294-1393) on the ingest alphabet
    switch (c) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'C': return 'G';
        case 'G': return 'C';
        default: return 'N';
    }
}

struct DirectAccess {
    std::vector<BitRow>& L;
    bool atomic;                                // concurrent writers of neighbouring bits (speculative passes)
    inline bool get(int g, int64_t i) { return L[g].get(i); }
    inline int64_t run_up(int g, int64_t a, int64_t b) { return L[g].run_up(a, b); }
    inline int64_t run_down(int g, int64_t a, int64_t b) { return L[g].run_down(a, b); }
    inline int64_t prev_set(int g, int64_t i) { return L[g].prev_set(i); }
    inline int64_t next_set(int g, int64_t i, int64_t limit) { return L[g].next_set(i, limit); }
    inline void commit(const int64_t* st, int64_t length, int n, const uint8_t*) {
        for (int k = 0; k < n; ++k) {
            if (atomic) L[k].set_range_atomic(st[k], st[k] + length);
            else L[k].set_range(st[k], st[k] + length);
        }
    }
};

// determineRegion into tmp coordinate buffers; returns slength (TRegion ctor, src/LCR.cpp:16-37)
template <class Acc>
inline int64_t det_region_impl(Acc& acc, const std::vector<int64_t>& len, int n, const int64_t* mstart, int64_t mlen, bool left,
                               int64_t* S, int64_t* E) {
    int64_t sl = 500000000;
    for (int i = 0; i < n; ++i) {
        if (left) {
            int64_t cp = acc.prev_set(i, mstart[i] - 1);
            if (cp < 0) cp = 0;
            S[i] = cp + 1;
            E[i] = mstart[i] - 1;
        } else {
            int64_t en = mstart[i] + mlen;
            int64_t cp = en + 1;
            if (cp < len[i]) cp = acc.next_set(i, cp, len[i]);
            S[i] = en + 1;
            E[i] = cp - 1;
        }
        sl = std::min(sl, E[i] - S[i]);
    }
    return sl;
}

template <class Acc>
int64_t Aligner::det_region_t(Acc& acc, const int64_t* mstart, int64_t mlen, bool left, int64_t* S, int64_t* E) const {
    return det_region_impl(acc, len_, n_, mstart, mlen, left, S, E);
}

template <class Acc>
void Aligner::accept_candidates_t(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& C, int cache_idx, Acc& acc,
                                  MumPool& mp, std::vector<int>& found, bool trace) {
    const CacheEntry& ce = C.entries[cache_idx];
    const int nq = n_ - 1;
    int64_t st_buf[64];
    uint8_t fw_buf[64];
    std::vector<int64_t> st_vec;
    std::vector<uint8_t> fw_vec;
    int64_t* st = st_buf;
    uint8_t* fw = fw_buf;
    if (n_ > 64) { st_vec.resize(n_); fw_vec.resize(n_); st = st_vec.data(); fw = fw_vec.data(); }
    for (int wi = 0; wi < ce.nwin; ++wi) {
        const WinRec& win = C.wins[ce.first_win + wi];
        const CandBatch& cb = C.chunks[win.chunk];
        if (trace) trace_.emplace_back(win.ref_start, win.ref_len);
        for (int32_t c = 0; c < win.ncand; ++c) {
            const int64_t ci = win.cand_off + c;
            const int64_t LON = cb.LON()[ci];
            bool bad = false;
            // Mum.DSP is 1-based (src/parsnp.cpp:1671,1681); range pre-check in unsigned arithmetic (1723)
            uint64_t dsp0 = (uint64_t)((int64_t)cb.K()[ci] + 1 + win.ref_start);
            if ((uint64_t)(dsp0 - (uint64_t)rs[0]) > (uint64_t)(uint32_t)(re[0] - rs[0])) bad = true;
            st[0] = (int64_t)dsp0 - 1;
            fw[0] = 1;
            // shortcut: a candidate whose reference interval is already covered trims to nothing in the first pass of the trim loop
            // below whatever the other genomes hold, and nothing before that point has a side effect
            if (!bad && st[0] >= 0 && st[0] + LON <= len_[0] && acc.get(0, st[0]) && acc.run_up(0, st[0], st[0] + LON) == LON) continue;
            // TMum ctor (src/TMum.cpp:13-72): a reverse-strand start is mirrored on the WHOLE genome length; the ctor's `ok` ends up
            // false as soon as one genome's interval leaves its sequence (a middle-genome failure makes the reference throw;
            // unreachable).  Range pre-check and ctor are fused into one pass; both only ever skip the candidate.
            bool any_fail = st[0] + LON > len_[0] || st[0] < 0;
            const int32_t* spj = cb.SP() + ci * nq;
            const uint8_t* fwj = cb.FWD() + ci * nq;
            for (int j = 1; j < n_; ++j) {
                const uint64_t dsp = (uint64_t)((int64_t)spj[j - 1] + 1 + rs[j]);
                bad |= (uint64_t)(dsp - (uint64_t)rs[j]) > (uint64_t)(uint32_t)(re[j] - rs[j]);
                int64_t s = (int64_t)dsp - 1;
                const uint8_t f = fwj[j - 1];
                if (!f) s = len_[j] - (s + LON);
                any_fail |= (s + LON > len_[j]) | (s < 0);
                st[j] = s;
                fw[j] = f;
            }
            if (bad || any_fail || LON < 5) continue;
            // trim (src/parsnp.cpp:1399-1477): every trim shifts ALL genomes, strand ignored
            int64_t length = LON;
            for (int j = 0; j < n_; ++j) {
                int64_t t1 = acc.run_up(j, st[j], st[j] + length);
                if (t1) { for (int i = 0; i < n_; ++i) st[i] += t1; length -= t1; }
                int64_t t2 = acc.run_down(j, st[j], st[j] + length);
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
            acc.commit(st, length, n_, fw);
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

}  // namespace pb200
