/* parsnp_b200.h - C ABI of the B200-native MUM + LCB engine (drop-in for the hot path of marbl/parsnp).
 *
 * Plain C: pointers, sizes, integer error codes; no C++/torch types.  All functions return 0 on success and a
 * negative pb200_status on error (pb200_last_error() gives the message).  The library is CUDA-only: every entry
 * point that computes fails with PB200_ERR_NO_CUDA when no sm_100-class device is usable - there is no CPU path.
 *
 * What each entry point replaces in the reference (file:line relative to the marbl/parsnp tree):
 *
 *   pb200_genomes_create / pb200_genomes_free
 *       the in-memory `vector<string> genomes` handed to the Aligner (src/parsnp.cpp:3141, 3180) - ASCII
 *       A,C,G,T,N only, i.e. *after* the ingest rules of src/parsnp.cpp:2999-3133.  Copies the texts to HBM.
 *
 *   pb200_search_windows
 *       boundary B2 = the csgmum C interface as used by Aligner::setMums1 for ONE reference window:
 *       new_CSG/build_CSG/find_leaves/free_CSG (src/csgmum/csg.h:70-74), Find_UM/Intersect_UM/Merge_Master
 *       (src/csgmum/mum.h:49-54) and the candidate emission loop (src/parsnp.cpp:1570-1695), batched over many
 *       windows per call.  Output = the `Mum{DSP,LON,forward}` candidates (src/csgmum/mum.h:27-31), 0-based.
 *
 *   pb200_align / pb200_align_resident
 *       boundary B1 minus file I/O = the sequence main() runs between "genomes in memory" and "final LCB list":
 *       Aligner::setInitialClusters, doWork, filterRandom1, setFinalClusters, filterRandomClustersSimple1,
 *       setFinalClusters, setInterClusterRegions (src/parsnp.cpp:3187-3273).  Output = this->mums and
 *       this->clusters as they stand when Aligner::writeOutput starts (src/parsnp.cpp:505).
 *
 *   pb200_minsize
 *       Converter()+Calculator() as used at src/parsnp.cpp:1502-1514.
 *
 *   pb200_mumi
 *       Aligner::setMumi (src/parsnp.cpp:1869-2115), the calcmumi=1 mode the Python driver runs first (parsnp:1365-1373).
 *
 *   pb200_comm_set / pb200_comm_clear  (multi-GPU; one process per GPU)
 *       no reference counterpart (the reference has no distributed backend): query genomes are sharded over
 *       ranks for the anchor scan, windows are sharded for the recursion; the caller's collectives carry the exchange.
 */
#ifndef PARSNP_B200_H
#define PARSNP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PB200_OK = 0,
    PB200_ERR_NO_CUDA = -1,      /* no usable CUDA device / extension built without kernels */
    PB200_ERR_ARG = -2,
    PB200_ERR_CUDA = -3,         /* a CUDA call failed; see pb200_last_error() */
    PB200_ERR_INTERNAL = -4,
    PB200_ERR_NO_MUMS = -5       /* reference: "NO MUMS FOUND" (src/parsnp.cpp:3223-3229) */
} pb200_status;

typedef struct pb200_genomes pb200_genomes;   /* genome texts resident in HBM on one device */
typedef struct pb200_result pb200_result;     /* MUM list + LCB list of one alignment */

/* ini keys of the reference (src/parsnp.cpp:2866-2901, template.ini) that the MUM+LCB path reads */
typedef struct {
    int32_t c;             /* [LCB] c          min LCB length                      (template default 21) */
    int32_t d;             /* [LCB] d          max gap between MUMs of an LCB      (300) */
    int32_t q;             /* [LCB] q          min region length for recursion     (30) */
    int64_t p;             /* [LCB] p          reference window length             (15000000) */
    float diagdiff;        /* [LCB] diagdiff                                       (0.12) */
    int32_t filter;        /* [MUM] filter                                         (1) */
    int32_t anchors_only;  /* [MUM] anchorsonly                                    (0) */
    const char* anchors;   /* [MUM] anchors    min anchor length expression        ("1.1*(Log(S))") */
    const char* mums;      /* [MUM] mums       min MUM length expression           ("1.1*(Log(S))") */
    int32_t flags;         /* PB200_FLAG_* */
    int32_t reserved;
} pb200_params;

#define PB200_FLAG_TRACE_WINDOWS 1   /* record the sequence of searched windows (tests) */
#define PB200_FLAG_NO_SPECULATION 2  /* skip the speculative batching pass: every region is searched on demand */
#define PB200_FLAG_UNALIGNED 4       /* also list the unaligned regions (ini [LCB] unaligned=1) */

void pb200_params_default(pb200_params* p);
const char* pb200_last_error(void);
const char* pb200_version(void);

/* 1 when a CUDA device of compute capability >= 10.0 is present and the kernels are loadable, else 0 */
int pb200_cuda_available(void);
/* number of CUDA devices this process can see (CUDA_VISIBLE_DEVICES applies); 0 without a driver.  parsnp_b200_core uses it to
 * spread the concurrent copies the Python driver launches (parsnp:1572-1597) over the GPUs of a box. */
int pb200_device_count(void);

/* ---- genomes ---- */
/* seqs[i] = ASCII text of genome i (A,C,G,T,N), lens[i] its length; genome 0 is the reference.
 * Host buffers are copied to device `device` (pageable or pinned); they must stay valid until
 * pb200_genomes_free (the host orchestrator reads them for the reverse-strand verification). */
int pb200_genomes_create(int device, int n, const uint8_t* const* seqs, const int64_t* lens, pb200_genomes** out);
void pb200_genomes_free(pb200_genomes* g);

/* ---- B2: batched window search ---- */
typedef struct {
    int64_t ref_start;     /* window start in genome 0 */
    int64_t ref_len;       /* window length */
    int64_t coord_off;     /* offset into `coords`: q_start[n-1] then q_len[n-1] (genomes 1..n-1) */
    int32_t minsize;
    int32_t pad;
} pb200_window;

/* Runs index build + both-strand scan of every query region + fold + emission for each window.
 * Outputs are malloc'ed by the library (free with pb200_free_buffer):
 *   cand_off[ntasks+1]; k[ncand], lon[ncand]; sp[ncand*(n-1)], fwd[ncand*(n-1)] */
int pb200_search_windows(pb200_genomes* g, int ntasks, const pb200_window* tasks, const int64_t* coords, int64_t ncoords,
                         int64_t** cand_off, int32_t** k, int32_t** lon, int32_t** sp, uint8_t** fwd);
void pb200_free_buffer(void* p);

/* ---- B1: the whole MUM + LCB path ----
 * Return value PB200_OK, or PB200_ERR_NO_MUMS when the anchor search found nothing (the reference then writes
 * "NO MUMS FOUND" and exits 0, src/parsnp.cpp:3223-3229): in BOTH cases *out is a valid result (empty lists for NO_MUMS,
 * its stats filled) that the caller must release with pb200_result_free.  On every other negative status *out is untouched.
 * An invalid minimum-length expression (e.g. "S/0") is reported as PB200_ERR_ARG (the reference's calculator calls exit(1),
 * src/Converter.cpp:252; a library must not end its host process). */
int pb200_align_resident(pb200_genomes* g, const pb200_params* prm, pb200_result** out);
/* convenience: upload + align + free the device copy (the end-to-end call a user makes with host buffers) */
int pb200_align(int device, int n, const uint8_t* const* seqs, const int64_t* lens, const pb200_params* prm,
                pb200_result** out);

int pb200_result_n(const pb200_result* r);                 /* number of genomes */
int64_t pb200_result_num_mums(const pb200_result* r);
/* length[m], slength[m], start[m*n], end[m*n], fwd[m*n] - any pointer may be NULL */
int pb200_result_mums(const pb200_result* r, int64_t* length, int64_t* slength, int64_t* start, int64_t* end, uint8_t* fwd);
/* the same arrays without a copy: pointers into the result, valid until pb200_result_free (any pointer argument may be NULL).
 * end[] = start[] + length (TMum::end) is not stored: it is materialised by the first call that asks for it. */
int pb200_result_mums_view(const pb200_result* r, const int64_t** length, const int64_t** slength, const int64_t** start,
                           const int64_t** end, const uint8_t** fwd);
int64_t pb200_result_num_clusters(const pb200_result* r);
/* type[c] (1 = LCB, 0 = inter-cluster record), nmums[c], length[c], start[c*n], end[c*n]; order = this->clusters */
int pb200_result_clusters(const pb200_result* r, int32_t* type, int64_t* nmums, int64_t* length, int64_t* start, int64_t* end);
/* MUMs of every cluster (Cluster::mums): off[c]..off[c+1] index into idx[], idx = positions in the MUM list; returns the
 * total number of indices (call with NULLs first). Inter-cluster records (type 0) have none. */
int pb200_result_cluster_mums(const pb200_result* r, int64_t* off, int64_t* idx);
/* unaligned regions (PB200_FLAG_UNALIGNED): the records Aligner::setUnalignableRegions (src/parsnp.cpp:2310-2382) writes to
 * parsnp.unalign, in its order: genome[r] (0-based), start[r], end[r] (`>genome+1:start-end`; the sequence printed is
 * genomes[genome].substr(start, end-start)).  Returns the number of records; any pointer may be NULL. */
int64_t pb200_result_unaligned(const pb200_result* r, int32_t* genome, int64_t* start, int64_t* end);
/* searched windows in order (PB200_FLAG_TRACE_WINDOWS): pairs (ref_start, ref_len) */
int64_t pb200_result_num_trace(const pb200_result* r);
int pb200_result_trace(const pb200_result* r, int64_t* pairs);
/* named counters/timers; returns number of values written */
int pb200_result_stats(const pb200_result* r, double* values, int cap);
const char* pb200_stats_names(void);                       /* comma-separated names matching pb200_result_stats */
void pb200_result_free(pb200_result* r);

/* ---- MUMi mode (ini calcmumi=1): Aligner::setMumi (src/parsnp.cpp:1869-2115) ----
 * out[j-1] = MUMi distance of query j to the reference = 1 - (reference positions of the first window covered by
 * MUMs >= 15 bp) / window length, with the reference's length-ratio rule; the value parsnp_core prints as
 * "<j>:<%f>" into all.mumi.  out must hold n-1 doubles. */
int pb200_mumi(pb200_genomes* g, const pb200_params* prm, double* out);

/* ---- minsize ---- */
int pb200_minsize(const char* expr, int64_t slength);

/* ---- engine profiling hooks (bench.py roofline) ---- */
/* last per-kernel-group device times of the engine in ms (CUDA events on the engine stream) */
int pb200_engine_timers(pb200_genomes* g, double* values, int cap);
const char* pb200_engine_timer_names(void);
void pb200_engine_reset_timers(pb200_genomes* g);

/* ---- multi-GPU (one process per GPU; every rank calls pb200_align_resident with the same genomes and parameters) ----
 * The library shards the search (queries for large windows, windows for the recursion batches)
 * and calls back for the collectives, which the host side implements with torch.distributed (NCCL over NVLink on the
 * GPUs, gloo in the CPU tests).  Callbacks return 0 on success.  `on_device` = the pointers are device memory of this
 * rank's GPU (the library has synchronised its stream before the call; the callback must complete the collective
 * before returning). */
typedef int (*pb200_allgather_cb)(void* user, const void* send, void* recv, int64_t bytes_per_rank, int on_device);
typedef int (*pb200_allreduce_cb)(void* user, void* buf_i32, int64_t count, int is_max, int on_device);   /* int32 min / max */
typedef int (*pb200_bcast_cb)(void* user, void* buf, int64_t bytes, int root, int on_device);
/* Every rank must then call pb200_align* with the same genomes and parameters: the ranks run in lock step and every search
 * starts with a 16-byte all-gather that compares their window lists - ranks that are out of step get PB200_ERR_INTERNAL
 * ("not searching the same windows"), never each other's candidates.
 * bcast_index != 0: the window index is built on rank 0 and broadcast; 0: every rank rebuilds it */
int pb200_comm_set(pb200_genomes* g, int rank, int world, pb200_allgather_cb ag, pb200_allreduce_cb ar, pb200_bcast_cb bc,
                   void* user, int bcast_index);
void pb200_comm_clear(pb200_genomes* g);

#ifdef __cplusplus
}
#endif
#endif
