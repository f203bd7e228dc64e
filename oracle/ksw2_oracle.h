/* ksw2_oracle.h — CPU oracle (TEST INFRASTRUCTURE ONLY; see ksw2_oracle.c). */
#ifndef FSV_KSW2_ORACLE_H_
#define FSV_KSW2_ORACLE_H_
#include <stdint.h>
#include "../include/focalsv_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
/* counters used to size the GPU design (how often the int8 lanes clamp or wrap) */
typedef struct fsvo_diag {
    int64_t clamp_inband;  /* lanes inside [st0,en0] where z was cut by the max-score clamp */
    int64_t clamp_oob;     /* same, in the rounded lanes outside [st0,en0] with t <= r */
    int64_t clamp_top;     /* same, in lanes above the first row (t > r): never feed a band cell */
    int64_t wraps;         /* int8 additions whose exact result left [-128,127] */
} fsvo_diag;

/* Both return n_cigar (>= 0) or -1 on allocation failure.  At most cigar_cap
 * words are copied to `cigar`; ez->n_cigar always holds the full length. */
int fsvo_extz2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m, const int8_t* mat,
               int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag,
               fsv_result* ez, uint32_t* cigar, int cigar_cap, fsvo_diag* dg);
int fsvo_extd2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m, const int8_t* mat,
               int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag,
               fsv_result* ez, uint32_t* cigar, int cigar_cap, fsvo_diag* dg);
int64_t fsvo_task_cells(int qlen, int tlen, int w);
extern int fsvo_force_scalar;   /* 1 = use the scalar lane loop even without diagnostics */
int32_t fsvo_gotoh2_global(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m,
                           const int8_t* mat, int q, int e, int q2, int e2);
int32_t fsvo_score_cigar(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m,
                         const int8_t* mat, int q, int e, int q2, int e2,
                         const uint32_t* cigar, int n_cigar, int* q_used, int* t_used);
/* batch driver with a thread pool (bench.py cpu_baseline leg): runs tasks
 * [0,n) with `threads` workers, one task per thread at a time. */
int fsvo_run_batch(const fsv_scoring* sc, const uint8_t* qarena, const uint8_t* tarena,
                   const fsv_task* tasks, int64_t n, int threads, fsv_result* out,
                   uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used);
/* global unit-cost edit distance (what edlib.align(a, b)["editDistance"] returns, remove_redundancy.py:57-63) */
int32_t fsvo_edit_distance(int alen, const uint8_t* a, int blen, const uint8_t* b);
#ifdef __cplusplus
}
#endif
#endif
