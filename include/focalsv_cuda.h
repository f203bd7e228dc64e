/*
 * focalsv_cuda.h — C ABI of libfocalsv_cuda.so (B200 / sm_100a).
 *
 * Drop-in boundary for FocalSV's alignment-DP hot path.  Every entry point
 * uses plain pointers and sizes only (no C++ / torch types) so it can be bound
 * with ctypes from the reference's Python call sites:
 *   focalsv/4_sv_calling/Dippav/DipPAV_variant_call.py:103-112   (asm5)
 *   focalsv/TRA_INV_DUP_call/Target/call_DUP_from_contigs.py:114-126 (asm10)
 *   focalsv/TRA_INV_DUP_call/Target/align_ins2ref.py:64-71      (map-hifi/pb/ont)
 * and from C at the one in-tree native call site
 *   software/hifiasm-0.16.1/Correct.cpp:7658-7705 (afine_gap_alignment).
 *
 * The arithmetic contract is the reference's ksw2:
 *   software/hifiasm-0.16.1/ksw2.h:23-32   ksw_extz_t           -> fsv_result
 *   software/hifiasm-0.16.1/ksw2.h:8-17    KSW_EZ_* flags       -> FSV_EZ_*
 *   software/hifiasm-0.16.1/ksw2.h:54-55   ksw_extz2_sse(...)   -> fsv_ksw_extz2 / fsv_align_batch (q2<0)
 *   software/hifiasm-0.16.1/ksw2.h:60-61   ksw_extd2_sse(...)   -> fsv_ksw_extd2 / fsv_align_batch
 * Results (every fsv_result field and every CIGAR word) are bit-identical to
 * ksw_extz2_sse, including its band-rounding and tie-break behaviour: PINNED against the reference's own file compiled
 * in place (oracle/_ref, tests/test_oracle_vs_ref.py, tests/golden/kat_extz2.npz).
 * ksw_extd2_sse: the reference tree holds only its PROTOTYPE (ksw2.h:60-61); the body is minimap2 2.24's
 * ksw2_extd2_sse.c, a dependency that is absent from /root/reference and from this image.  The dual-affine results are
 * bit-identical to this repository's restatement of that published algorithm (oracle/ksw2_oracle.c, fsvo_extd2), which is
 * anchored (extd2(q,e,q,e) == extz2(q,e) on every field; global score == an independent two-piece Gotoh DP; every
 * CIGAR re-scores to ez.score) but PARITY-UNPINNED at the source level.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device fsv_init
 * fails with FSV_ERR_NO_DEVICE.
 */
#ifndef FOCALSV_CUDA_H_
#define FOCALSV_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSV_ABI_VERSION 4

/* ksw2.h:6 */
#define FSV_NEG_INF (-0x40000000)

/* ksw2.h:8-14 — same numeric values so reference flag words pass through. */
#define FSV_EZ_SCORE_ONLY  0x01
#define FSV_EZ_RIGHT       0x02
#define FSV_EZ_GENERIC_SC  0x04
#define FSV_EZ_APPROX_MAX  0x08
#define FSV_EZ_APPROX_DROP 0x10
#define FSV_EZ_EXTZ_ONLY   0x40
#define FSV_EZ_REV_CIGAR   0x80

/* error codes (0 = ok, negative = error; nothing ever aborts or throws) */
enum {
    FSV_OK = 0,
    FSV_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: there is no CPU path */
    FSV_ERR_CUDA = -2,        /* a CUDA runtime call failed; see fsv_last_error */
    FSV_ERR_INVALID = -3,     /* bad argument (null pointer, negative size, ...) */
    FSV_ERR_NOMEM = -4,       /* device or host allocation failed */
    FSV_ERR_CIGAR_CAP = -5,   /* cigar arena too small; *cigar_used = words needed */
    FSV_ERR_SCORING = -6,     /* scoring outside the range ksw2's int8 lanes can hold */
    FSV_ERR_STATE = -7        /* batch object used out of order */
};

/* Scoring of one batch.  Mirrors the (m, mat, q, e, q2, e2) arguments of
 * ksw2.h:54-61.  mat is the m*m matrix in row-major order; only m = 5 is
 * used by the reference (Correct.cpp:7670-7672).  q2 < 0 selects the
 * single-affine recurrence (ksw_extz2_sse); q2 >= 0 the dual-affine one
 * (ksw_extd2_sse). */
typedef struct fsv_scoring {
    int8_t m;
    int8_t q, e;
    int8_t q2, e2;
    int8_t reserved[3];
    int8_t mat[32];           /* first m*m entries used */
} fsv_scoring;

/* One (query segment, target segment) task.  Sequences are uint8 codes
 * 0..m-1 (Correct.cpp:7676-7685 encoding: A,C,G,T -> 0..3, other -> 4)
 * living in caller-owned arenas; offsets are in bytes. */
typedef struct fsv_task {
    int64_t q_off;            /* offset of query[0] in the query arena */
    int64_t t_off;            /* offset of target[0] in the target arena */
    int32_t qlen, tlen;
    int32_t w;                /* band width, <0 = max(qlen,tlen) (ksw2_extz2_sse.c:72) */
    int32_t zdrop;            /* <0 disables (ksw2.h:170) */
    int32_t end_bonus;
    int32_t flag;             /* FSV_EZ_* */
} fsv_task;

/* Field-for-field ksw_extz_t (ksw2.h:23-32) minus the owned pointer. */
typedef struct fsv_result {
    int32_t max;              /* ksw_extz_t.max (31-bit unsigned field) */
    int32_t zdropped;
    int32_t max_q, max_t;
    int32_t mqe, mqe_t;
    int32_t mte, mte_q;
    int32_t score;
    int32_t reach_end;
    int32_t n_cigar;
    int32_t status;           /* per-task: 0 ok, FSV_ERR_SCORING if it was reset like ksw2_extz2_sse.c:82 */
    int64_t cigar_off;        /* index (uint32 words) of this task's CIGAR in the arena */
    int64_t cells;            /* in-band DP cells actually processed: sum_r (en0-st0+1) */
} fsv_result;

typedef struct fsv_stats {
    int64_t tasks;            /* tasks processed since init */
    int64_t cells;            /* in-band cells since init */
    int64_t fill_launches;    /* DP fill kernel launches */
    int64_t backtrack_launches;
    int64_t other_launches;   /* pack / scan / gather kernels */
    int64_t exact_path_tasks; /* tasks re-run on the int8-exact kernel */
    int64_t h2d_bytes, d2h_bytes;
    double  fill_ms;          /* CUDA-event time of fill kernels (last run) */
    double  backtrack_ms;     /* CUDA-event time of backtrack kernels (last run) */
    double  total_ms;         /* CUDA-event time of the last fsv_batch_run */
    int64_t traceback_bytes;  /* traceback bytes written by the last run */
    int64_t segmented_tasks;  /* last run: long tasks cut into segments that ran on separate SMs */
    int64_t segment_fallbacks;/* last run: of those, tasks whose boundary check failed and that were run again whole */
} fsv_stats;

typedef struct fsv_ctx fsv_ctx;
typedef struct fsv_batch fsv_batch;

/* ---- context --------------------------------------------------------- */
/* One context per device per host thread/process (thread-compatible, not
 * thread-safe).  device < 0 selects the current CUDA device. */
int fsv_init(int device, fsv_ctx** ctx);
void fsv_destroy(fsv_ctx* ctx);
const char* fsv_strerror(int code);
const char* fsv_last_error(const fsv_ctx* ctx);
int fsv_abi_version(void);
int fsv_device_count(void);
int fsv_get_stats(const fsv_ctx* ctx, fsv_stats* out);
/* tunables (FSV_ERR_INVALID for unknown keys or bad values):
 *   "traceback_budget_bytes"  size of the traceback page pool (0 = auto)
 *   "traceback_page_bytes"    page size of the pool (default 32 MiB)
 *   "lazy_min_pages"          tasks with at least this many traceback pages take them as they advance (default 16; 0 = off)
 *   "lazy_fill_pct"           such a task starts only while the projected peak of those running stays below this share of the pool (default 65)
 *   "pool_stall_ms"           watchdog of the lazy pool (default 60000)
 *   "segment_min_diags"       tasks with at least this many antidiagonals are cut into segments that run on separate
 *                             CTAs (0 = off, -1 = auto, the default: tasks whose chain would outlast 60 % of the batch)
 *   "segment_rows"            antidiagonals per segment (0 = auto: 4 x the cold-start lead, whole traceback pages)
 *   "segment_warm_pct"        cold-start lead of a segment in percent of the band width (default 300: 2w to forget the start, w more until most lanes of the band entered after that; a boundary that does not verify is repaired: one segment runs again, so a shorter lead trades cells for repairs)
 *   "segment_pool_pct"        share of the traceback pool the segmented tasks may hold (default 45)
 *   "segment_slots"           1 = long tasks beyond the pool share re-use the static pages of earlier ones in turn (default), 0 = they stay whole
 *   "segment_align_pages"     1 = segments are whole traceback pages (default), 0 = any multiple of 1024 antidiagonals (experiment)
 *   "segment_pool_pct_bound"  the pool share when the batch's traceback does not fit the pool and it is throughput-bound (default 25)
 *   "segment_extz"            auto mode: 1 = extension (EXTZ_ONLY) tasks are segmented too (default), 0 = global tasks only
 *   "ew_kernel"               which tasks run on the edge-warp fill kernel (left-aligned traceback, no wildcard bases, band >= 32): 1 = those whose band needs
 *                             up to 4 main warps (default), 2 = all of them, 0 = none (everything on the one-vector-per-thread kernel)
 *   "force_exact"             1 = int8-exact general kernel only
 *   "exact_smem_lanes", "force_excl"   kernel experiments */
int fsv_set_option(fsv_ctx* ctx, const char* key, int64_t value);

/* ---- one-call batch API (host buffers in, host buffers out) ----------
 * Replaces a loop of ksw_extz2_sse / ksw_extd2_sse calls (ksw2.h:54-61).
 * All buffers are caller-owned.  On FSV_ERR_CIGAR_CAP the fsv_result array is
 * fully valid (n_cigar set) and *cigar_used holds the words required. */
int fsv_align_batch(fsv_ctx* ctx, const fsv_scoring* sc,
                    const uint8_t* query_arena, size_t query_bytes,
                    const uint8_t* target_arena, size_t target_bytes,
                    const fsv_task* tasks, size_t n_tasks,
                    fsv_result* out,
                    uint32_t* cigar_arena, size_t cigar_cap, size_t* cigar_used);

/* ---- staged batch API (device-resident timing, double buffering) ------
 * create = H2D of arenas + task table; run = kernels only, inputs resident
 * in HBM; fetch = D2H of results + compact CIGAR arena. */
int fsv_batch_create(fsv_ctx* ctx, const fsv_scoring* sc,
                     const uint8_t* query_arena, size_t query_bytes,
                     const uint8_t* target_arena, size_t target_bytes,
                     const fsv_task* tasks, size_t n_tasks, fsv_batch** batch);
int fsv_batch_run(fsv_batch* batch);
int fsv_batch_fetch(fsv_batch* batch, fsv_result* out,
                    uint32_t* cigar_arena, size_t cigar_cap, size_t* cigar_used);
void fsv_batch_destroy(fsv_batch* batch);
/* Per-task device timeline of the last run (GPU globaltimer, ns): start_end_ns[2*i] = when task i got its
 * traceback pages and started, [2*i+1] = when its CIGAR was written.  For schedule analysis / tracing. */
int fsv_batch_timeline(fsv_batch* batch, int64_t* start_end_ns);
/* How the batch was planned (host only, valid after fsv_batch_create): plan[i] = FSV_PLAN_* bits of task i,
 * warps per task in bits 8..11, number of segments in bits 16..  For parity sampling and schedule analysis. */
#define FSV_PLAN_DPX       0x1   /* runs on the DPX fill kernel */
#define FSV_PLAN_GENERAL   0x2   /* runs on the general int8-exact kernel */
#define FSV_PLAN_SEGMENTED 0x4   /* cut into segments that run on separate CTAs */
#define FSV_PLAN_EXCLUSIVE 0x8   /* runs on the exclusive (one CTA per SM) launch */
#define FSV_PLAN_EDGE_WARP 0x10  /* (with FSV_PLAN_DPX) runs on the edge-warp variant of the DPX fill kernel */
int fsv_batch_plan(const fsv_batch* batch, int32_t* plan);

/* ---- Level 1: the pipeline hook -------------------------------
 * What FocalSV does around its `minimap2 -a -x <preset> --cs -r2k ref.fa contigs.fa` call
 * (focalsv/4_sv_calling/Dippav/DipPAV_variant_call.py:97-137, call_DUP_from_contigs.py:114-126,
 * align_ins2ref.py:64-71): every haplotype contig of a region against that region's reference window, one record
 * per (contig, window) pair carrying what the consumers read from pysam today (reference_start, reference_end,
 * is_reverse, mapping_quality, cigartuples; extract_contig_signature_CCS.py:343-356).  The scoring is the preset's
 * (minimap2 2.24 values), the band is minimap2's bw*1.5+1 for `-r<bw>`.
 * The reference is passed once as codes (0..3, other 4); windows are [region_start[i], region_end[i]) into it and are
 * NOT copied on the host.  contig i = contig_codes[contig_off[i] .. +contig_len[i]).  Each pair is one global
 * dual-affine task (ksw_extd2_sse with the preset's z-drop); pos = region_start, ref_end = pos + reference bases
 * the CIGAR consumes.  Errors as fsv_align_batch (FSV_ERR_CIGAR_CAP: records valid, *cigar_used = words needed). */
typedef struct fsv_preset {
    char    name[16];         /* "asm5", "asm10", "map-hifi", "map-pb", "map-ont", "hifiasm" (Correct.h:1194-1199) */
    int32_t a, b, q, e, q2, e2;   /* q2 < 0: single-affine */
    int32_t zdrop, zdrop_inv;
    int32_t bw, bw_long;
    int32_t sc_ambi, end_bonus;
} fsv_preset;
/* Host only.  FSV_ERR_INVALID for an unknown name; `sc` (may be NULL) receives the 5x5 matrix minimap2's
 * ksw_gen_simple_mat builds for the preset. */
int fsv_preset_lookup(const char* name, fsv_preset* out, fsv_scoring* sc);

typedef struct fsv_record {
    int64_t pos;              /* reference_start (0-based) */
    int64_t ref_end;          /* reference_end */
    int64_t cigar_off;        /* index (uint32 words) into the CIGAR arena */
    int32_t n_cigar;
    int32_t query_length;
    int32_t score;
    int32_t zdropped;
    int32_t is_reverse;       /* always 0: strand selection belongs to seeding (row f2) */
    int32_t mapq;             /* placeholder: 60, as a unique full-length contig alignment gets; 0 when zdropped (the CIGAR
                               * then ends at the maximum cell and consumes only a prefix of the contig) */
} fsv_record;
int fsv_realign_regions(fsv_ctx* ctx, const uint8_t* ref_codes, size_t ref_len,
                        const int64_t* region_start, const int64_t* region_end,
                        const uint8_t* contig_codes, size_t contig_bytes,
                        const int64_t* contig_off, const int32_t* contig_len, size_t n,
                        const char* preset, int bw, int flag,
                        fsv_record* out, uint32_t* cigar_arena, size_t cigar_cap, size_t* cigar_used);

/* ---- next row: CIGAR -> DEL / INS signatures on the device -------------
 * What the reference does with every aligned contig right after the alignment
 * (focalsv/4_sv_calling/Dippav/extract_contig_signature_CCS.py:14-127, extract_sig_from_cigar): keep D / I
 * operations >= min_svlen and fold neighbouring ones of the same contig (:49-127).  Runs on the CIGARs of a batch
 * that has been run (fsv_batch_run) while they are still in HBM; only the signatures are copied back.
 * ref_start[i] = reference coordinate of target[0] of task i (NULL = 0).  Per task: its DELs, then its INSs, in
 * CIGAR order.  On a too-small `out` returns FSV_ERR_CIGAR_CAP with *n_out = records needed. */
typedef struct fsv_signature {
    int32_t task;             /* index in the caller's task array */
    int32_t svtype;           /* 0 = DEL, 1 = INS */
    int64_t pos;              /* reference offset (sig[2]) */
    int32_t svlen;            /* sig[3] */
    int32_t read_start;       /* contig offsets (sig[5], sig[6]) */
    int32_t read_end;
    int32_t pad_;
} fsv_signature;
int fsv_batch_signatures(fsv_batch* batch, const int64_t* ref_start, int min_svlen,
                         fsv_signature* out, size_t cap, size_t* n_out);

/* ---- next row: batched global edit distance -----------------------------
 * What the reference computes with edlib.align(seq1, seq2)["editDistance"] (default mode NW, k = -1: the plain
 * unit-cost Levenshtein distance) when it de-duplicates INS alleles
 * (focalsv/4_sv_calling/Dippav/remove_redundancy.py:57-63, remove_redundancy_region_based.py:123-128).
 * Sequences are raw bytes compared for equality (ASCII or codes); at most 8 distinct byte values may occur in one
 * call (DNA + N), otherwise FSV_ERR_INVALID.  dist[i] = distance of pair i; an empty sequence gives the other's length. */
typedef struct fsv_pair {
    int64_t a_off, b_off;     /* byte offsets into the two arenas */
    int32_t a_len, b_len;
} fsv_pair;
int fsv_edit_distance_batch(fsv_ctx* ctx, const uint8_t* a_arena, size_t a_bytes, const uint8_t* b_arena, size_t b_bytes,
                            const fsv_pair* pairs, size_t n_pairs, int32_t* dist);

/* ---- single-task convenience, argument-for-argument ksw2.h:54-61 ------
 * (km dropped; ez -> fsv_result + caller-owned cigar buffer). */
int fsv_ksw_extz2(fsv_ctx* ctx, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                  int8_t m, const int8_t* mat, int8_t q, int8_t e, int w, int zdrop,
                  int end_bonus, int flag, fsv_result* ez, uint32_t* cigar, int cigar_cap);
int fsv_ksw_extd2(fsv_ctx* ctx, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                  int8_t m, const int8_t* mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w,
                  int zdrop, int end_bonus, int flag, fsv_result* ez, uint32_t* cigar, int cigar_cap);

/* ---- next row f2 (host only): seeding, chaining, task decomposition ------
 * The caller of ksw2 inside `minimap2 -a -x asm5 --cs -r2k` (DipPAV_variant_call.py:103): minimap2 2.24's sketch.c /
 * chain.c / align.c (mm_align1), a dependency that is not in the reference tree, restated from the published algorithm
 * (PARITY UNPINNED; same strand, one chain, global ends): (w,k)-minimizers of both sequences, seeds occurring more than
 * max_occ times dropped, colinear chaining with minimap2's gap cost, then the pair is cut at chain anchors at least
 * min_fill bases apart.  The pieces tile [0,qlen) x [0,tlen) in order; each is one global DP task (a piece with an empty
 * side is a pure gap), so a 200 kb x 200 kb pair becomes a few hundred small fills plus one rectangle per structural
 * variant.  FSV_ERR_CIGAR_CAP with *n_pieces = entries needed when `cap` is too small. */
typedef struct fsv_piece {
    int32_t q_beg, q_end;     /* query  [q_beg, q_end) */
    int32_t t_beg, t_end;     /* target [t_beg, t_end) */
} fsv_piece;
int fsv_chain_pieces(const uint8_t* query, int32_t qlen, const uint8_t* target, int32_t tlen,
                     int k, int w, int max_occ, int max_gap, int min_fill,
                     fsv_piece* pieces, size_t cap, size_t* n_pieces, int32_t* chain_score, int32_t* n_anchors);

/* Second version (host only, PARITY UNPINNED like fsv_chain_pieces): both strands, several chains per pair, extension ends.
 * chains[0] is the primary alignment; the others are supplementary (their query interval overlaps the accepted ones by less
 * than half).  Coordinates of a chain with strand = 1 refer to the REVERSE COMPLEMENT of the query.  pieces[piece_off ..
 * piece_off + n_pieces) tile the core [q_beg, q_end) x [t_beg, t_end) in order (global fills, as fsv_chain_pieces);
 * lq / lt (rq / rt) are the query / target bases before (after) the core to offer to an extension task
 * (FSV_EZ_EXTZ_ONLY; the left one on reversed sequences with FSV_EZ_RIGHT | FSV_EZ_REV_CIGAR), as mm_align1 does.
 * sub_score = best score of a chain that was dropped as a secondary of this one (mapq).
 * FSV_ERR_CIGAR_CAP with *n_chains / *n_pieces = entries needed when a buffer is too small. */
typedef struct fsv_chain_opts {
    int32_t k, w;             /* minimizer k-mer and window (asm5: 19, 19) */
    int32_t max_occ;          /* seeds occurring more often in the target are dropped */
    int32_t max_gap;          /* chaining / extension reach (the preset's bw_long for assemblies) */
    int32_t min_fill;         /* min_ksw_len: anchors at least this far apart bound a fill (200) */
    int32_t max_chains;       /* primary + supplementary alignments kept per pair */
    int32_t min_chain_score, min_anchors;      /* minimap2 -m 40 -n 3 */
    int32_t a, q, e, end_bonus;                /* match score, gap open / extend of the preset: size of the extension offers */
} fsv_chain_opts;
typedef struct fsv_chain {
    int32_t strand, score, n_anchors, sub_score;
    int32_t q_beg, q_end, t_beg, t_end;
    int32_t piece_off, n_pieces;
    int32_t lq, lt, rq, rt;
} fsv_chain;
int fsv_chain_pair(const uint8_t* query, int32_t qlen, const uint8_t* target, int32_t tlen, const fsv_chain_opts* opts,
                   fsv_chain* chains, size_t chain_cap, size_t* n_chains,
                   fsv_piece* pieces, size_t piece_cap, size_t* n_pieces);

/* Stitch the CIGARs of one pair's pieces, in order, into one (host only): task_of[i] = index of piece i's task in `res`
 * (its CIGAR is copied) or -1 (a piece with an empty side: a pure I / D of the other side's length); neighbouring
 * operations of the same kind are merged.  FSV_ERR_CIGAR_CAP with *n_out = words needed when `cap` is too small. */
int fsv_stitch_cigars(const fsv_piece* pieces, const int32_t* task_of, size_t n_pieces,
                      const fsv_result* res, const uint32_t* cigar_arena,
                      uint32_t* out, size_t cap, size_t* n_out);

/* ---- roofline denominator ---------------------------------------------
 * Measured issue rate (32-bit lane-ops per second, whole device) of a dependency-free
 * stream of the instruction class the fill kernels are built from:
 * kind 0 VIADD.16x2, 1 VIMNMX3.S16x2, 2 VIADDMNMX.S16x2, 3 LOP3, 4 IMAD, 5 the fill
 * kernel's mix, 6 PRMT, 7 VIMNMX3 + IMAD interleaved (both integer pipes). */
int fsv_measure_int_peak(fsv_ctx* ctx, int kind, double* lane_ops_per_s);

/* ---- host-side helpers (no device work) ------------------------------ */
/* In-band cell count of one task if it runs to completion:
 * sum over r of (en0 - st0 + 1) with the bounds of ksw2_extz2_sse.c:102-110. */
int64_t fsv_task_cells(int32_t qlen, int32_t tlen, int32_t w);
/* Longest-processing-time-first binning of tasks over n_bins:
 * bin_of[i] receives the bin of task i.  Deterministic. */
int fsv_lpt_bins(const fsv_task* tasks, size_t n_tasks, int n_bins, int32_t* bin_of);

#ifdef __cplusplus
}
#endif
#endif /* FOCALSV_CUDA_H_ */
