/*
 * ref_shim.c — TEST INFRASTRUCTURE ONLY.
 * Thin adapter compiled TOGETHER WITH the reference's own, unmodified
 * software/hifiasm-0.16.1/ksw2_extz2_sse.c (in place, from /root/reference;
 * no reference source is copied into this repository).  Output goes to
 * oracle/_ref/libksw2_ref.so (git-ignored).  It converts ksw_extz_t
 * (ksw2.h:23-32) into the flat fsv_result used by the tests.
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "ksw2.h"                      /* -I/root/reference/software/hifiasm-0.16.1 */
#include "../include/focalsv_cuda.h"

static int64_t cells_done(int qlen, int tlen, int w, const ksw_extz_t* ez, int flag)
{   /* no instrumentation inside the reference: count the antidiagonals a full run covers */
    int64_t n = 0; int r;
    (void)ez; (void)flag;
    if (qlen <= 0 || tlen <= 0) return 0;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    for (r = 0; r < qlen + tlen - 1; ++r) {
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
        if (en > ((r + w) >> 1)) en = (r + w) >> 1;
        if (st > en) break;
        n += en - st + 1;
    }
    return n;
}

int fsvref_extz2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m, const int8_t* mat,
                 int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag,
                 fsv_result* out, uint32_t* cigar, int cigar_cap)
{
    ksw_extz_t ez;
    memset(&ez, 0, sizeof(ez));
    ksw_extz2_sse(0, qlen, query, tlen, target, m, mat, q, e, w, zdrop, end_bonus, flag, &ez);
    out->max = (int32_t)ez.max; out->zdropped = ez.zdropped;
    out->max_q = ez.max_q; out->max_t = ez.max_t;
    out->mqe = ez.mqe; out->mqe_t = ez.mqe_t; out->mte = ez.mte; out->mte_q = ez.mte_q;
    out->score = ez.score; out->reach_end = ez.reach_end; out->n_cigar = ez.n_cigar;
    out->status = 0; out->cigar_off = 0;
    out->cells = cells_done(qlen, tlen, w, &ez, flag);
    if (ez.n_cigar > 0 && cigar)
        memcpy(cigar, ez.cigar, (size_t)(ez.n_cigar < cigar_cap ? ez.n_cigar : cigar_cap) * 4);
    free(ez.cigar);
    return out->n_cigar;
}

/* thread-pool batch driver for the CPU baseline ("kind": "reference"): batch_pool.h (largest first, memory gate) */
#include "batch_pool.h"
static void ref_run_one(const fsv_scoring* sc, const uint8_t* qa, const uint8_t* ta, const fsv_task* t,
                        fsv_result* out, uint32_t* cig, int cap)
{
    fsvref_extz2(t->qlen, qa + t->q_off, t->tlen, ta + t->t_off, sc->m, sc->mat, sc->q, sc->e,
                 t->w, t->zdrop, t->end_bonus, t->flag, out, cig, cap);
}

int fsvref_run_batch(const fsv_scoring* sc, const uint8_t* qarena, const uint8_t* tarena,
                     const fsv_task* tasks, int64_t n, int threads, fsv_result* out,
                     uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used)
{
    return bp_run_batch(ref_run_one, sc, qarena, tarena, tasks, n, threads, out, cigar_arena, cigar_cap, cigar_used);
}
