/*
 * afine_gap_probe.c — C caller of libfocalsv_cuda.so shaped like the reference's one in-tree native call site,
 * afine_gap_alignment (software/hifiasm-0.16.1/Correct.cpp:7658-7705): ASCII reads in, the caller's c2n table encodes
 * them (forward, or reversed for the backward strand), a 5x5 matrix with a zero wildcard row/column is built from
 * (sc_mch, sc_mis), ONE alignment call is made with (gapo, gape, bandLen, zdrop, end_bonus, mode), and nine ksw_extz_t
 * fields are read back.  The only difference to the reference's body is the call itself: fsv_ksw_extz2 instead of
 * ksw_extz2_sse (ksw2.h:54-55), with a caller-owned CIGAR buffer instead of the krealloc'ed ez.cigar.
 *
 * Built and run by tests/test_boundary.py (gcc, links the shared object directly; no Python in between):
 *   afine_gap_probe <cases.txt>      one case per line: qseq tseq strand mode end_bonus
 * prints per case: global_score extension_score q_boundary_score q_boundary_t t_boundary_score t_boundary_q max_t max_q dropped n_cigar cigar
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "focalsv_cuda.h"

/* hifiasm's constants (Correct.h:1194-1199) */
#define P_FORWARD 0
#define P_BACKWARD 1
#define P_MATCH 2
#define P_MISMATCH 4
#define P_GAPO 4
#define P_GAPE 2
#define P_ZDROP 400
#define P_BAND 500

typedef struct {
    long long max_q_pos, max_t_pos, global_score, extention_score, q_boundary_score, q_boundary_t_coordinate,
              t_boundary_score, t_boundary_q_coordinate, droped;
    int n_cigar;
} probe_out;

static int probe_afine_gap_alignment(fsv_ctx* ctx, const char* qseq, uint8_t* qnum, int ql, const char* tseq, uint8_t* tnum, int tl,
                                     const uint8_t* c2n, int strand, int sc_mch, int sc_mis, int gapo, int gape, int bandLen,
                                     int zdrop, int end_bonus, int mode, probe_out* o, uint32_t* cigar, int cigar_cap)
{
    int8_t mat[25];
    int i, j, rc, a = sc_mch, b = sc_mis < 0 ? sc_mis : -sc_mis;
    fsv_result ez;
    for (i = 0; i < 5; ++i)
        for (j = 0; j < 5; ++j) mat[i * 5 + j] = (int8_t)((i == 4 || j == 4) ? 0 : i == j ? a : b);
    memset(&ez, 0, sizeof ez);
    for (i = 0; i < tl; ++i) tnum[i] = c2n[(uint8_t)tseq[strand == P_FORWARD ? i : tl - i - 1]];
    for (i = 0; i < ql; ++i) qnum[i] = c2n[(uint8_t)qseq[strand == P_FORWARD ? i : ql - i - 1]];
    rc = fsv_ksw_extz2(ctx, ql, qnum, tl, tnum, 5, mat, (int8_t)gapo, (int8_t)gape, bandLen, zdrop, end_bonus, mode, &ez, cigar, cigar_cap);
    if (rc != FSV_OK) return rc;
    o->global_score = ez.score; o->extention_score = ez.max;
    o->q_boundary_score = ez.mqe; o->q_boundary_t_coordinate = ez.mqe_t;
    o->t_boundary_score = ez.mte; o->t_boundary_q_coordinate = ez.mte_q;
    o->max_t_pos = ez.max_t; o->max_q_pos = ez.max_q; o->droped = ez.zdropped; o->n_cigar = ez.n_cigar;
    return FSV_OK;
}

int main(int argc, char** argv)
{
    static char q[1 << 16], t[1 << 16];
    static uint8_t qn[1 << 16], tn[1 << 16];
    static uint32_t cig[1 << 17];
    uint8_t c2n[256];
    fsv_ctx* ctx = 0;
    FILE* fp;
    int strand, mode, eb, rc, k;
    if (argc < 2) { fprintf(stderr, "usage: %s cases.txt\n", argv[0]); return 2; }
    memset(c2n, 4, sizeof c2n);
    c2n['A'] = c2n['a'] = 0; c2n['C'] = c2n['c'] = 1; c2n['G'] = c2n['g'] = 2; c2n['T'] = c2n['t'] = 3;
    rc = fsv_init(-1, &ctx);
    if (rc != FSV_OK) { fprintf(stderr, "fsv_init: %s\n", fsv_strerror(rc)); return 3; }
    fp = fopen(argv[1], "r");
    if (!fp) { perror(argv[1]); return 2; }
    while (fscanf(fp, "%65535s %65535s %d %d %d", q, t, &strand, &mode, &eb) == 5) {
        probe_out o;
        memset(&o, 0, sizeof o);
        rc = probe_afine_gap_alignment(ctx, q, qn, (int)strlen(q), t, tn, (int)strlen(t), c2n, strand, P_MATCH, P_MISMATCH, P_GAPO, P_GAPE,
                                       P_BAND, P_ZDROP, eb, mode, &o, cig, (int)(sizeof cig / 4));
        if (rc != FSV_OK) { fprintf(stderr, "fsv_ksw_extz2: %s (%s)\n", fsv_strerror(rc), fsv_last_error(ctx)); return 4; }
        printf("%lld %lld %lld %lld %lld %lld %lld %lld %lld %d ", o.global_score, o.extention_score, o.q_boundary_score,
               o.q_boundary_t_coordinate, o.t_boundary_score, o.t_boundary_q_coordinate, o.max_t_pos, o.max_q_pos, o.droped, o.n_cigar);
        for (k = 0; k < o.n_cigar; ++k) printf("%u%c", cig[k] >> 4, "MIDNSHP=XB"[cig[k] & 0xf]);
        printf("\n");
    }
    fclose(fp);
    fsv_destroy(ctx);
    return 0;
}
