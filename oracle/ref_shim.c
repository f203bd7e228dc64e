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

/* thread-pool batch driver for the CPU baseline ("kind": "reference") */
typedef struct {
    const fsv_scoring* sc; const uint8_t *qa, *ta; const fsv_task* tasks; int64_t n;
    fsv_result* out; uint32_t** cig; volatile int64_t next; pthread_mutex_t mu;
} pool_t;

static void* worker(void* arg)
{
    pool_t* p = (pool_t*)arg;
    for (;;) {
        int64_t i; const fsv_task* t; int cap;
        pthread_mutex_lock(&p->mu); i = p->next++; pthread_mutex_unlock(&p->mu);
        if (i >= p->n) break;
        t = &p->tasks[i];
        cap = t->qlen + t->tlen + 2;
        p->cig[i] = (t->flag & FSV_EZ_SCORE_ONLY) ? 0 : (uint32_t*)malloc((size_t)cap * 4);
        fsvref_extz2(t->qlen, p->qa + t->q_off, t->tlen, p->ta + t->t_off, p->sc->m, p->sc->mat, p->sc->q, p->sc->e,
                     t->w, t->zdrop, t->end_bonus, t->flag, &p->out[i], p->cig[i], cap);
    }
    return 0;
}

int fsvref_run_batch(const fsv_scoring* sc, const uint8_t* qarena, const uint8_t* tarena,
                     const fsv_task* tasks, int64_t n, int threads, fsv_result* out,
                     uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used)
{
    pool_t p; pthread_t* th; int i; int64_t k, used = 0;
    if (threads < 1) threads = 1;
    p.sc = sc; p.qa = qarena; p.ta = tarena; p.tasks = tasks; p.n = n; p.out = out; p.next = 0;
    p.cig = (uint32_t**)calloc((size_t)n + 1, sizeof(uint32_t*));
    pthread_mutex_init(&p.mu, 0);
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (i = 0; i < threads; ++i) pthread_create(&th[i], 0, worker, &p);
    for (i = 0; i < threads; ++i) pthread_join(th[i], 0);
    for (k = 0; k < n; ++k) {
        out[k].cigar_off = used;
        if (p.cig[k]) {
            if (cigar_arena && used + out[k].n_cigar <= cigar_cap)
                memcpy(cigar_arena + used, p.cig[k], (size_t)out[k].n_cigar * 4);
            free(p.cig[k]);
        }
        used += out[k].n_cigar;
    }
    if (cigar_used) *cigar_used = used;
    free(th); free(p.cig); pthread_mutex_destroy(&p.mu);
    return 0;
}
