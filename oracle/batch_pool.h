/*
 * batch_pool.h — TEST INFRASTRUCTURE ONLY (shared by ksw2_oracle.c and ref_shim.c).
 *
 * Thread-pool driver of the CPU arms: one task per thread at a time (SURVEY 8d "CPU baseline timing"),
 * tasks handed out LARGEST FIRST (longest-processing-time order, so the tail of a batch does not leave
 * threads idle behind one long task drawn last), and a memory gate: the traceback matrix of a task is
 * (qlen+tlen-1) * n_col*16 bytes (ksw2_extz2_sse.c:75-76,94), 7 GB for a 1.1 Mb x 1.1 Mb band-3001 pair,
 * so a task starts only while the tracebacks in flight stay below a budget (FSVO_MEM_GB, default half of
 * physical memory); a task larger than the budget runs alone.
 */
#ifndef FSV_BATCH_POOL_H
#define FSV_BATCH_POOL_H
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../include/focalsv_cuda.h"

typedef void (*bp_run_fn)(const fsv_scoring* sc, const uint8_t* qa, const uint8_t* ta, const fsv_task* t,
                          fsv_result* out, uint32_t* cig, int cap);

typedef struct {
    const fsv_scoring* sc; const uint8_t *qa, *ta; const fsv_task* tasks; int64_t n;
    fsv_result* out; uint32_t** cig; int64_t* order; int64_t* need; uint8_t* taken;
    int64_t next; int64_t inflight, budget; int running;
    pthread_mutex_t mu; pthread_cond_t cv; bp_run_fn run;
} bp_pool_t;

static int64_t bp_cells_est(const fsv_task* t)
{
    int64_t mn = t->qlen < t->tlen ? t->qlen : t->tlen, w = t->w < 0 ? (t->qlen > t->tlen ? t->qlen : t->tlen) : t->w;
    if (t->qlen <= 0 || t->tlen <= 0) return 0;
    return ((int64_t)t->qlen + t->tlen - 1) * (mn < w + 1 ? mn : w + 1);
}
static int64_t bp_tb_bytes(const fsv_task* t)
{
    int64_t mn = t->qlen < t->tlen ? t->qlen : t->tlen, w = t->w < 0 ? (t->qlen > t->tlen ? t->qlen : t->tlen) : t->w;
    int64_t n_col;
    if (t->qlen <= 0 || t->tlen <= 0 || (t->flag & FSV_EZ_SCORE_ONLY)) return 0;
    n_col = ((mn < w + 1 ? mn : w + 1) + 15) / 16 + 1;
    return ((int64_t)t->qlen + t->tlen - 1) * n_col * 16;
}

static const int64_t* bp_sort_key;
static int bp_cmp_desc(const void* a, const void* b)
{
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    if (bp_sort_key[x] != bp_sort_key[y]) return bp_sort_key[x] > bp_sort_key[y] ? -1 : 1;
    return x < y ? -1 : x > y;
}

static void* bp_worker(void* arg)
{
    bp_pool_t* p = (bp_pool_t*)arg;
    for (;;) {
        int64_t k, i = -1; const fsv_task* t; int cap;
        pthread_mutex_lock(&p->mu);
        for (;;) {
            while (p->next < p->n && p->taken[p->next]) ++p->next;          /* head = first task not handed out */
            if (p->next >= p->n) break;
            /* memory gate: the largest waiting task whose traceback fits beside those in flight (anything, if nothing runs) */
            for (k = p->next; k < p->n && k < p->next + 8192; ++k)
                if (!p->taken[k] && (p->running == 0 || p->inflight + p->need[p->order[k]] <= p->budget)) { i = p->order[k]; p->taken[k] = 1; break; }
            if (i >= 0) break;
            pthread_cond_wait(&p->cv, &p->mu);
        }
        if (i < 0) { pthread_mutex_unlock(&p->mu); break; }
        p->inflight += p->need[i]; ++p->running;
        pthread_mutex_unlock(&p->mu);
        t = &p->tasks[i];
        cap = t->qlen + t->tlen + 2;
        p->cig[i] = (t->flag & FSV_EZ_SCORE_ONLY) ? 0 : (uint32_t*)malloc((size_t)(cap > 0 ? cap : 1) * 4);
        p->run(p->sc, p->qa, p->ta, t, &p->out[i], p->cig[i], cap);
        pthread_mutex_lock(&p->mu);
        p->inflight -= p->need[i]; --p->running;
        pthread_cond_broadcast(&p->cv);
        pthread_mutex_unlock(&p->mu);
    }
    return 0;
}

/* Runs the n tasks on `threads` threads; results in caller order, CIGARs packed into cigar_arena in caller order. */
static int bp_run_batch(bp_run_fn run, const fsv_scoring* sc, const uint8_t* qarena, const uint8_t* tarena,
                        const fsv_task* tasks, int64_t n, int threads, fsv_result* out,
                        uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used)
{
    bp_pool_t p; pthread_t* th; int i; int64_t k, used = 0; int64_t* est;
    const char* env = getenv("FSVO_MEM_GB");
    if (threads < 1) threads = 1;
    memset(&p, 0, sizeof p);
    p.sc = sc; p.qa = qarena; p.ta = tarena; p.tasks = tasks; p.n = n; p.out = out; p.run = run;
    p.cig = (uint32_t**)calloc((size_t)n + 1, sizeof(uint32_t*));
    p.order = (int64_t*)malloc(((size_t)n + 1) * 8); p.need = (int64_t*)malloc(((size_t)n + 1) * 8);
    est = (int64_t*)malloc(((size_t)n + 1) * 8); p.taken = (uint8_t*)calloc((size_t)n + 1, 1);
    for (k = 0; k < n; ++k) { p.order[k] = k; est[k] = bp_cells_est(&tasks[k]); p.need[k] = bp_tb_bytes(&tasks[k]); }
    bp_sort_key = est;                       /* (the drivers are not re-entrant; they are called from one Python thread) */
    qsort(p.order, (size_t)n, 8, bp_cmp_desc);
    p.budget = env ? (int64_t)(atof(env) * 1e9) : (int64_t)sysconf(_SC_PHYS_PAGES) * (int64_t)sysconf(_SC_PAGE_SIZE) / 2;
    pthread_mutex_init(&p.mu, 0); pthread_cond_init(&p.cv, 0);
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (i = 0; i < threads; ++i) pthread_create(&th[i], 0, bp_worker, &p);
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
    free(th); free(p.cig); free(p.order); free(p.need); free(est); free(p.taken);
    pthread_mutex_destroy(&p.mu); pthread_cond_destroy(&p.cv);
    return 0;
}
#endif
