// Bridge to the reference's vendored libMUSCLE 3.7 (third party, linked unchanged: third_party/build_muscle.py).
// Same settings as the reference's wrapper (src/MuscleInterface.cpp:38-49).
#include "xmfa.h"
#ifdef PB200_HAVE_MUSCLE
#include "libMUSCLE/muscle.h"
#include "libMUSCLE/params.h"
#include "libMUSCLE/seq.h"
#include "libMUSCLE/seqvect.h"
#include "libMUSCLE/msa.h"
#include "libMUSCLE/threadstorage.h"
namespace muscle { extern void MUSCLE(SeqVect& v, MSA& msaOut); }
#endif

namespace pb200 {

bool muscle_available() {
#ifdef PB200_HAVE_MUSCLE
    return true;
#else
    return false;
#endif
}

bool muscle_align(const std::vector<std::string>& seqs, std::vector<std::string>& out) {
#ifdef PB200_HAVE_MUSCLE
    using namespace muscle;
    g_SeqType.get() = SEQTYPE_DNA;
    g_uMaxIters.get() = 1;
    g_bStable.get() = true;
    g_bVerbose.get() = false;
    g_bQuiet.get() = true;
    g_SeqWeight1.get() = SEQWEIGHT_ClustalW;
    SetMaxIters(g_uMaxIters.get());
    SetSeqWeightMethod(g_SeqWeight1.get());
    g_ulMaxSecs.get() = 0;
    SeqVect sv;
    for (size_t i = 0; i < seqs.size(); i++) {
        Seq s;
        s.SetId((unsigned)i);
        s.SetName("seq00000");
        s.resize(seqs[i].size());
        std::copy(seqs[i].begin(), seqs[i].end(), s.begin());
        sv.AppendSeq(s);
    }
    MSA msa;
    MUSCLE(sv, msa);
    out.clear();
    out.resize(msa.GetSeqCount());
    for (size_t i = 0; i < msa.GetSeqCount(); i++) {
        unsigned idx = msa.GetSeqIndex((unsigned)i);
        out[i] = std::string(msa.GetSeqBuffer(idx), msa.GetColCount());
    }
    return true;
#else
    (void)seqs; (void)out;
    return false;
#endif
}

}  // namespace pb200
