/* oracle/port/lr2rmats_port.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * A plain-C, single-threaded restatement of the lr2rmats per-alignment hot path,
 * written from the behaviour of the reference (Xinglab/lr2rmats @ 0072f71), with
 * its sequential cursors and back-scans kept sequential on purpose: the CUDA
 * product uses closed forms / parallel folds, and this file is what they are
 * checked against.  Every function cites the reference lines it follows.
 *
 * Parity pin: this port is itself checked byte-for-byte against the compiled
 * reference binary (oracle/_ref/lr2rmats) -- tests/test_oracle_pin.py, golden
 * outputs under tests/golden/.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lr2rmats_port.h"

#define OP(c)    ((c) & 0xfu)
#define OLEN(c)  ((int)((c) >> 4))
enum { C_M = 0, C_I, C_D, C_N, C_S, C_H, C_P, C_EQ, C_X, C_B };

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s ? s : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static int iabs(int x) { return x < 0 ? -x : x; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ======================================================================== filter */

/* bam_cigar2rlen (htslib/sam.c:334): reference-consuming ops are M D N = X */
static int ref_len(const uint32_t *c, int n)
{
    int i, l = 0;
    for (i = 0; i < n; ++i) {
        unsigned op = OP(c[i]);
        if (op == C_M || op == C_D || op == C_N || op == C_EQ || op == C_X) l += OLEN(c[i]);
    }
    return l;
}

/* remove_overlap, bam_filter.c:49-59 (0-based pos against 1-based GTF coords, kept as is) */
static int rm_overlap(int tid, int pos, int rlen, const lrb_anno *rm)
{
    int i;
    if (!rm) return 0;
    for (i = 0; i < rm->n_trans; ++i) {
        if (tid == rm->tid[i] && !(pos > rm->end[i] || rm->start[i] > pos + rlen - 1)) return 1;
        if (tid < rm->tid[i]) return 0;
    }
    return 0;
}

/* gtf_filter, bam_filter.c:61-86.  returns 1 = drop */
static int filter_one(const lrb_batch *b, int64_t r, const lrb_anno *rm, const lrb_filter_params *p, int *score, int *intron_n)
{
    const uint32_t *c = b->cigar + b->cigar_off[r];
    int n_c = (int)(b->cigar_off[r + 1] - b->cigar_off[r]);
    int i, del_len = 0;
    if (b->flag[r] & 4) return 1;                                   /* :63 */
    *intron_n = 0;
    for (i = 0; i < n_c; ++i) {                                     /* :68-71 */
        if (OP(c[i]) == C_N) (*intron_n)++;
        else if (OP(c[i]) == C_D) del_len += OLEN(c[i]);
    }
    int cigar_qlen = b->l_qseq[r];                                  /* :73-76 */
    if (n_c > 0) {
        unsigned op0 = OP(c[0]), op1 = OP(c[n_c - 1]);
        if (op0 == C_S || op0 == C_H) cigar_qlen -= OLEN(c[0]);
        if (n_c > 1 && (op1 == C_S || op1 == C_H)) cigar_qlen -= OLEN(c[n_c - 1]);
    }
    if ((cigar_qlen + 0.0) / b->l_qseq[r] < p->cov_rate) return 1;  /* :77  double quotient vs promoted float */
    int ed = b->nm[r];                                              /* :79-81 */
    if ((cigar_qlen - ed + del_len) < p->map_qual * cigar_qlen) return 1;   /* :82  float product */
    if (rm_overlap(b->tid[r], b->pos[r], ref_len(c, n_c), rm)) return 1;    /* :83 */
    *score = cigar_qlen - ed + del_len;                             /* :84 */
    return 0;
}

/* the qname-run state machine of bam_filter(), bam_filter.c:129-159 */
int orc_filter(const lrb_batch *b, const lrb_anno *rm, const lrb_filter_params *p, lrb_filter_result *out)
{
    int64_t n = b->n, r, nk = 0;
    uint8_t *pass = xcalloc(n, 1);
    int32_t *score = xcalloc(n, 4), *intron = xcalloc(n, 4);
    uint32_t *keep = xmalloc(n * 4);
    int have = 0; uint64_t lq = 0; int64_t best = -1;
    int b_score = 0, s_score = 0, b_intron = 0;
    for (r = 0; r < n; ++r) {
        int sc = 0, in = 0;
        if (filter_one(b, r, rm, p, &sc, &in)) continue;            /* :131 */
        pass[r] = 1; score[r] = sc; intron[r] = in;
        if (have && b->qname_hash[r] == lq) {                       /* :133 */
            if (sc > b_score) { best = r; s_score = b_score; b_score = sc; b_intron = in; }
            else if (sc > s_score) s_score = sc;
        } else {
            if (have && s_score < p->sec_rat * b_score && b_intron >= p->min_intron_n) keep[nk++] = (uint32_t)best;   /* :144-148 */
            best = r; b_score = sc; s_score = 0; b_intron = in; lq = b->qname_hash[r]; have = 1;
        }
    }
    if (have && s_score < p->sec_rat * b_score && b_intron >= p->min_intron_n) keep[nk++] = (uint32_t)best;           /* :155-159 */
    out->n = n; out->pass = pass; out->score = score; out->intron_n = intron; out->n_keep = nk; out->keep_idx = keep;
    return 0;
}
void orc_free_filter(lrb_filter_result *r)
{
    free((void *)r->pass); free((void *)r->score); free((void *)r->intron_n); free((void *)r->keep_idx);
    memset(r, 0, sizeof *r);
}

/* ===================================================================== CIGAR walk */

/* gen_exon, bam2gtf.c:31-78; returns exon count written to es/ee (capacity n_c+1) */
static int walk_cigar(const uint32_t *c, int n_c, int pos, const lrb_exon_params *p, int *es, int *ee)
{
    int n = 0, i, start = pos + 1, end = start - 1;
    for (i = 0; i < n_c; ++i) {
        int l = OLEN(c[i]);
        switch (OP(c[i])) {
        case C_N:
            if (l >= p->min_intron) {
                if (n == 0 || (end - start + 1) >= p->min_exon) { es[n] = start; ee[n] = end; n++; }
                start = end + l + 1;
            }
            end += l; break;
        case C_D:
            if (l > p->max_delet) {
                if (n == 0 || (end - start + 1) >= p->min_exon) { es[n] = start; ee[n] = end; n++; }
                start = end + l + 1;
            }
            end += l; break;
        case C_M: case C_EQ: case C_X: end += l; break;
        default: break;                                             /* I S H P B and unknown codes: ignored */
        }
    }
    es[n] = start; ee[n] = end; n++;                                /* :76 last exon always */
    return n;
}

/* set_trans_name's sort_exon (gtf.c:37-51,94-100): ascending (start,end).  CIGAR chains are
 * already ascending; insertion sort keeps this a no-op for them and exact for GTF chains. */
static void sort_chain(int *es, int *ee, int n)
{
    int i, j;
    for (i = 1; i < n; ++i) {
        int s = es[i], e = ee[i];
        for (j = i - 1; j >= 0 && (es[j] > s || (es[j] == s && ee[j] > e)); --j) { es[j + 1] = es[j]; ee[j + 1] = ee[j]; }
        es[j + 1] = s; ee[j + 1] = e;
    }
}

/* read_bam_trans / the bam2gtf loop, bam2gtf.c:89-110,150-156 */
int orc_bam2gtf(const lrb_batch *b, const uint32_t *sel, int64_t n_sel, const lrb_exon_params *p, lrb_exon_result *out)
{
    int64_t n = sel ? n_sel : b->n, r, cap = 0, tot = 0;
    for (r = 0; r < n; ++r) { int64_t k = sel ? sel[r] : r; cap += (b->cigar_off[k + 1] - b->cigar_off[k]) + 1; }
    uint32_t *off = xmalloc((n + 1) * 4), *ridx = sel ? xmalloc(n * 4) : NULL;
    int32_t *es = xmalloc(cap * 4), *ee = xmalloc(cap * 4), *tid = xmalloc(n * 4);
    uint8_t *rev = xmalloc(n);
    for (r = 0; r < n; ++r) {
        int64_t k = sel ? sel[r] : r;
        off[r] = (uint32_t)tot;
        if (ridx) ridx[r] = (uint32_t)k;
        tid[r] = b->tid[k];
        rev[r] = b->xs[k] == 0 ? ((b->flag[k] & 16) != 0) : (b->xs[k] == '+' ? 0 : 1);   /* bam2gtf.c:35-37 */
        if (b->flag[k] & 4) continue;                               /* gen_trans :82 */
        int ne = walk_cigar(b->cigar + b->cigar_off[k], (int)(b->cigar_off[k + 1] - b->cigar_off[k]), b->pos[k], p, es + tot, ee + tot);
        sort_chain(es + tot, ee + tot, ne);
        tot += ne;
    }
    off[n] = (uint32_t)tot;
    out->n_reads = n; out->read_idx = ridx; out->tid = tid; out->is_rev = rev; out->exon_off = off; out->exon_start = es; out->exon_end = ee;
    return 0;
}
void orc_free_exon(lrb_exon_result *r)
{
    free((void *)r->read_idx); free((void *)r->tid); free((void *)r->is_rev); free((void *)r->exon_off);
    free((void *)r->exon_start); free((void *)r->exon_end);
    memset(r, 0, sizeof *r);
}

/* ============================================================ transcripts & merge */

typedef struct {
    int tid, is_rev, start, end;      /* trans_t.tid/is_rev/start/end */
    int n; const int *es, *ee;        /* exon chain (read-only view) */
    int fs, le;                       /* exon[0].start, exon[n-1].end (mutable copies) */
    int cov; uint32_t src;            /* src: row of the list that was folded */
} mt_t;
#define XS_(t, i) ((i) == 0 ? (t)->fs : (t)->es[i])
#define XE_(t, i) ((i) == (t)->n - 1 ? (t)->le : (t)->ee[i])

typedef struct { mt_t *t; int64_t n, m; } mlist_t;
static void ml_push(mlist_t *L, const mt_t *t)
{
    if (L->n == L->m) { L->m = L->m ? L->m * 2 : 1024; L->t = xrealloc(L->t, L->m * sizeof(mt_t)); }
    L->t[L->n++] = *t;
}

/* check_iden, gtf.c:54-92.  Returns 0 identical, 2 "partial", -1 different (1 is never produced) */
static int chain_iden(const mt_t *t1, const mt_t *t2, int ss_dis, int end_dis)
{
    const mt_t *l, *s; int full;
    if (t1->n > t2->n) { l = t1; s = t2; full = 0; }
    else if (t1->n < t2->n) { l = t2; s = t1; full = 0; }
    else { l = t1; s = t2; full = 1; }
    int i, j;
    if (full) {
        if (iabs(XS_(l, 0) - XS_(s, 0)) > end_dis) return -1;
        for (i = 0; i < l->n - 1; ++i) {
            if (iabs(XE_(l, i) - XE_(s, i)) > ss_dis) return -1;
            if (iabs(XS_(l, i + 1) - XS_(s, i + 1)) > ss_dis) return -1;
        }
        if (iabs(XE_(l, l->n - 1) - XE_(s, s->n - 1)) > end_dis) return -1;
        return 0;
    } else {
        int pm = -1;
        if (iabs(XS_(l, 0) - XS_(s, 0)) > end_dis) return -1;
        for (i = 0; i < l->n - 1; ++i) {
            if (iabs(XE_(l, i) - XE_(s, 0)) <= ss_dis && iabs(XS_(l, i + 1) - XS_(s, 1)) <= ss_dis) {
                pm = 2;
                for (i = i + 1, j = 1; i < l->n - 1 && j < s->n - 1; ++i, ++j) {
                    if (iabs(XE_(l, i) - XE_(s, j)) > ss_dis) return -1;
                    if (iabs(XS_(l, i + 1) - XS_(s, j + 1)) > ss_dis) return -1;
                }
                break;
            }
        }
        if (iabs(XE_(l, l->n - 1) - XE_(s, s->n - 1)) > end_dis) return -1;
        return pm;
    }
}

/* exon_overlap_frac, update_gtf.c:80-89 (double quotient returned as float) */
static float ovlp_frac(int s1, int e1, int s2, int e2)
{
    if (s1 > e2 || s2 > e1) return 0.0f;
    int ov = imin(e1, e2) - imax(s1, s2) + 1;
    int ml = imin(e1 - s1 + 1, e2 - s2 + 1);
    return (float)(ov / (ml + 0.0));
}

/* merge_trans + merge_trans1/2, update_gtf.c:98-163.  returns 1 if t was absorbed/dropped */
static int fold_one(const mt_t *t, mlist_t *T, const lrb_update_params *p)
{
    int64_t i;
    for (i = T->n - 1; i >= 0; --i) {
        mt_t *E = T->t + i;
        if (t->tid > E->tid || t->start > E->end) return 0;                          /* :148 */
        if (p->force_strand && t->is_rev != E->is_rev) continue;                     /* :149 */
        if (t->n == 1 && E->n == 1) {                                                /* merge_trans2 :122-140 */
            if (iabs(t->fs - E->fs) > p->end_dis) continue;
            if (iabs(t->le - E->le) > p->end_dis) continue;
            if (ovlp_frac(t->fs, t->le, E->fs, E->le) >= p->single_exon_ovlp_frac) {
                E->cov++;
                if (t->fs < E->fs) { E->fs = t->fs; E->start = t->fs; }
                if (t->le > E->le) { E->le = t->le; E->end = t->le; }
                return 1;
            }
        } else if (t->n > 1 && E->n > 1) {                                           /* merge_trans1 :98-119 */
            int ret = chain_iden(t, E, p->ss_dis, p->end_dis);
            if (ret == 0) {
                E->cov++;
                if (t->fs < E->fs) { E->fs = t->fs; E->start = t->fs; }
                if (t->le > E->le) { E->le = t->le; E->end = t->le; }
                return 1;
            } else if (ret == 2) return 1;
        }
    }
    return 0;
}

static void merged_out(const mlist_t *T, lrb_merged_list *o)
{
    int64_t n = T->n, i;
    uint32_t *cand = xmalloc(n * 4);
    int32_t *cov = xmalloc(n * 4), *tt = xmalloc(n * 4), *ts = xmalloc(n * 4), *te = xmalloc(n * 4), *fs = xmalloc(n * 4), *le = xmalloc(n * 4);
    for (i = 0; i < n; ++i) {
        const mt_t *e = T->t + i;
        cand[i] = e->src; cov[i] = e->cov; tt[i] = e->tid; ts[i] = e->start; te[i] = e->end; fs[i] = e->fs; le[i] = e->le;
    }
    o->n = n; o->cand = cand; o->cov = cov; o->t_tid = tt; o->t_start = ts; o->t_end = te; o->first_start = fs; o->last_end = le;
}
static void merged_free(lrb_merged_list *o)
{
    free((void *)o->cand); free((void *)o->cov); free((void *)o->t_tid); free((void *)o->t_start); free((void *)o->t_end);
    free((void *)o->first_start); free((void *)o->last_end);
}

/* ================================================================== classification */

typedef struct {
    int tid, is_rev, start, end, n;
    const int *es, *ee;
    uint8_t *fl;                       /* slot flags LRB_F_* */
    int lfull, rfull, lnoth, rnoth, full, known, known_site, unreliable, sj_checked;
    int ref;
} rd_t;

static int ex_ovlp(int s1, int e1, int s2, int e2) { return !(s1 > e2 || s2 > e1); }   /* update_gtf.c:91-95 */

/* check_full, update_gtf.c:629-681 */
static void full_check(rd_t *t, const lrb_anno *A, int a, int level)
{
    const int *as = A->exon_start + A->exon_off[a], *ae = A->exon_end + A->exon_off[a];
    int an = (int)(A->exon_off[a + 1] - A->exon_off[a]);
    if (t->lfull && t->rfull) return;
    int i = t->n - 1, j = an - 1, ii;
    if (level == 1) {
        if (!t->lfull && t->ee[0] == ae[0]) t->lfull = 1;
        if (!t->rfull && t->es[i] == as[j]) t->rfull = 1;
    } else if (level == 2) {
        if (!t->lfull && ex_ovlp(t->es[0], t->ee[0], as[0], ae[0])) t->lfull = 1;
        if (!t->rfull && ex_ovlp(t->es[i], t->ee[i], as[j], ae[j])) t->rfull = 1;
    } else if (level == 3 || level == 4) {
        if (!t->lfull) {
            if (ex_ovlp(t->es[0], t->ee[0], as[0], ae[0])) t->lfull = 1;
            else for (ii = 0; ii < an; ++ii) if (ex_ovlp(t->es[0], t->ee[0], as[ii], ae[ii])) { t->lnoth = 0; break; }
        }
        if (level == 3 && !t->rfull) {
            if (ex_ovlp(t->es[i], t->ee[i], as[j], ae[j])) t->rfull = 1;
            else for (ii = 0; ii < an; ++ii) if (ex_ovlp(t->es[i], t->ee[i], as[ii], ae[ii])) { t->rnoth = 0; break; }
        }
    }
}

/* check_splice_site, update_gtf.c:717-779 (acceptor compare against exon[j].start kept) */
static int site_check(rd_t *t, const lrb_anno *A, int a, int dis)
{
    const int *as = A->exon_start + A->exon_off[a], *ae = A->exon_end + A->exon_off[a];
    int an = (int)(A->exon_off[a + 1] - A->exon_off[a]);
    int bam_ovlp = 0, iden = 0, all = (t->n - 1) * 2, i, j;
    int os = imax(t->start, A->start[a]), oe = imin(t->end, A->end[a]);
    for (i = 0; i < t->n - 1; ++i) {
        if (t->ee[i] >= os && t->ee[i] <= oe) bam_ovlp++;
        if (t->es[i + 1] >= os && t->es[i + 1] <= oe) bam_ovlp++;
    }
    for (i = 0; i < an - 1; ++i) {
        if (ae[i] >= os && ae[i] <= oe)
            for (j = 0; j < t->n - 1; ++j)
                if (iabs(ae[i] - t->ee[j]) <= dis) { iden++; t->fl[j] &= ~LRB_F_NOVEL_DON; }
        if (as[i + 1] >= os && as[i + 1] <= oe)
            for (j = 0; j < t->n - 1; ++j)
                if (iabs(as[i + 1] - t->es[j]) <= dis) { iden++; t->fl[j] &= ~LRB_F_NOVEL_ACC; }
    }
    for (i = 0; i < an; ++i)
        for (j = 0; j < t->n; ++j)
            if (iabs(as[i] - t->es[j]) <= dis && iabs(ae[i] - t->ee[j]) <= dis) t->fl[j] &= ~LRB_F_NOVEL_EXON;
    for (i = 0; i < an - 1; ++i)
        for (j = 0; j < t->n - 1; ++j)
            if (iabs(ae[i] - t->ee[j]) <= dis && iabs(as[i + 1] - t->es[j + 1]) <= dis) t->fl[j] &= ~LRB_F_NOVEL_JUNC;
    if (all == bam_ovlp && bam_ovlp == iden) { t->known = 1; return 1; }
    if (iden > 0) { t->known_site = 1; return 2; }
    return 0;
}

/* check_with_anno_trans, update_gtf.c:786-835 (sequential cursor kept) */
static void anno_check(rd_t *t, const lrb_anno *A, int *cursor, const lrb_update_params *p)
{
    int i, ref = -1;
    for (i = *cursor; i < A->n_trans; ++i) {
        int cmp;
        if (t->tid < A->tid[i] || (t->tid == A->tid[i] && t->end <= A->start[i])) cmp = -1;          /* comp_trans :786-790 */
        else if (A->tid[i] < t->tid || (A->tid[i] == t->tid && A->end[i] <= t->start)) cmp = 1;
        else cmp = 0;
        if (cmp < 0) break;
        if (cmp > 0) { if (*cursor == i) ++(*cursor); continue; }
        int an = (int)(A->exon_off[i + 1] - A->exon_off[i]);
        full_check(t, A, i, p->full_level);
        if (t->n == 1 && an == 1) {
            if (ovlp_frac(t->es[0], t->ee[0], A->exon_start[A->exon_off[i]], A->exon_end[A->exon_off[i]]) >= p->single_exon_ovlp_frac) {
                ref = i; t->known = 1; break;
            }
        } else if (t->n > 1 && an > 1) {
            int ret = site_check(t, A, i, p->ss_dis);
            if (ret == 1) { ref = i; break; }
            else if (ret == 2) ref = i;
        }
    }
    t->ref = ref;
    if (ref != -1) t->is_rev = A->is_rev[ref];                      /* :823-833 strand flip */
    /* set_full, :683-696 */
    int l = p->full_level;
    if (l == 5) t->full = 1;
    else if (l == 4) t->full = (t->lfull || t->lnoth);
    else if (l == 3) t->full = ((t->lfull || t->lnoth) && (t->rfull || t->rnoth));
    else t->full = (t->lfull && t->rfull);
}

/* check_short_sj1, update_gtf.c:589-603 */
static int sj_one(int tid, int start, int end, const lrb_sj *S, int64_t i, const lrb_update_params *p)
{
    while (i < S->n) {
        if (S->tid[i] > tid || (S->tid[i] == tid && S->don[i] >= end)) return 0;
        if (iabs(S->don[i] - start) <= p->ss_dis && iabs(S->acc[i] - end) <= p->ss_dis) {
            int cnt = p->use_multi ? S->uniq_c[i] + S->multi_c[i] : S->uniq_c[i];
            if (cnt >= p->min_sj_cnt) return 1;
        }
        i++;
    }
    return 0;
}

/* check_with_short_sj + check_short_sj, update_gtf.c:609-627,698-709 */
static int sj_check(rd_t *t, const lrb_sj *S, int64_t *cursor, const lrb_update_params *p)
{
    int64_t i = *cursor; int j, ret = 1, done = 0, res = 0;
    while (i < S->n) {
        if (S->tid[i] < t->tid || (S->tid[i] == t->tid && S->acc[i] <= t->start)) { i++; *cursor = i; }
        else if (S->tid[i] > t->tid || (S->tid[i] == t->tid && S->don[i] >= t->end)) { res = 0; done = 1; break; }
        else {
            for (j = 0; j < t->n - 1; ++j)
                if ((t->fl[j] & LRB_F_NOVEL_JUNC) && sj_one(t->tid, t->ee[j] + 1, t->es[j + 1] - 1, S, i, p) == 0) {
                    t->fl[j] |= LRB_F_UNRELIABLE; ret = 0;
                }
            res = ret; done = 1; break;
        }
    }
    if (!done) res = 0;
    t->unreliable = 1 - res; t->sj_checked = 1;
    return res;
}

/* ---- growable u32/i32 vectors ---- */
typedef struct { uint32_t *v; int64_t n, m; } uvec;
static void uv_push(uvec *a, uint32_t x) { if (a->n == a->m) { a->m = a->m ? a->m * 2 : 1024; a->v = xrealloc(a->v, a->m * 4); } a->v[a->n++] = x; }

typedef struct { int tid, key1, key2, score, type, is_rev; } set_ent;
typedef struct { set_ent *v; int64_t n, m; } setvec;
static void sv_push(setvec *a, set_ent e) { if (a->n == a->m) { a->m = a->m ? a->m * 2 : 1024; a->v = xrealloc(a->v, a->m * sizeof(set_ent)); } a->v[a->n++] = e; }

/* the four back-scan "sets" of print_trans_summary: add_simp_gene / _exon / _site / _sj,
 * update_gtf.c:181-295.  match test precedes the stop test; gene equality ignores tid. */
static void set_add(setvec *S, set_ent e, int by_gene, int add)
{
    int64_t i;
    for (i = S->n - 1; i >= 0; --i) {
        set_ent *x = S->v + i;
        int eq = by_gene ? (x->key1 == e.key1) : (x->tid == e.tid && x->key1 == e.key1 && x->key2 == e.key2);
        if (eq) { x->score += add; return; }
        if (e.tid > x->tid) break;
    }
    e.score = add;
    sv_push(S, e);
}

int orc_update(const lrb_exon_result *C, const lrb_anno *A, const lrb_sj *S, const lrb_update_params *p, lrb_update_result *out)
{
    int64_t n = C->n_reads, r, ne = C->exon_off[n];
    memset(out, 0, sizeof *out);
    /* bam_T copy (read_bam_trans initial state, bam2gtf.c:97-103) */
    uint32_t *off = xmalloc((n + 1) * 4); memcpy(off, C->exon_off, (n + 1) * 4);
    int32_t *es = xmalloc(ne * 4), *ee = xmalloc(ne * 4), *tid = xmalloc(n * 4), *ref = xmalloc(n * 4);
    memcpy(es, C->exon_start, ne * 4); memcpy(ee, C->exon_end, ne * 4); memcpy(tid, C->tid, n * 4);
    uint8_t *rev = xmalloc(n), *fl = xmalloc(ne);
    uint32_t *cls = xcalloc(n, 4), *ridx = NULL;
    if (C->read_idx) { ridx = xmalloc(n * 4); memcpy(ridx, C->read_idx, n * 4); }
    uvec known = {0}, unrec = {0}, nv_read = {0}, nv_lo = {0}, nv_n = {0}, nv_piece = {0};
    mlist_t upd = {0};
    int cursor = 0; int64_t sj_cursor = 0; int64_t sj_n = S ? S->n : 0;

    for (r = 0; r < n; ++r) {
        int en = (int)(off[r + 1] - off[r]), j;
        if (en == 0) { free(off); free(es); free(ee); free(tid); free(ref); free(rev); free(fl); free(cls); free(ridx); return LRB_E_UNMAPPED; }
        for (j = 0; j < en; ++j)
            fl[off[r] + j] = (j < en - 1) ? (LRB_F_NOVEL_EXON | LRB_F_NOVEL_DON | LRB_F_NOVEL_ACC | LRB_F_NOVEL_JUNC) : LRB_F_NOVEL_EXON;
        rd_t t; memset(&t, 0, sizeof t);
        t.tid = C->tid[r]; t.is_rev = C->is_rev[r]; t.n = en; t.es = es + off[r]; t.ee = ee + off[r]; t.fl = fl + off[r];
        t.start = t.es[0]; t.end = t.ee[en - 1]; t.lnoth = 1; t.rnoth = 1;
        anno_check(&t, A, &cursor, p);                                               /* check_trans :942 */
        int sj_ok = 1;
        if (t.full && !t.known && t.known_site && sj_n > 0) sj_ok = sj_check(&t, S, &sj_cursor, p);
        rev[r] = (uint8_t)t.is_rev; ref[r] = t.ref;
        cls[r] = (t.known ? LRB_C_KNOWN : 0) | (t.known_site ? LRB_C_KNOWN_SITE : 0) | (t.unreliable ? LRB_C_UNRELIABLE : 0) |
                 (t.full ? LRB_C_FULL : 0) | (t.lfull ? LRB_C_LFULL : 0) | (t.rfull ? LRB_C_RFULL : 0) |
                 (t.lnoth ? LRB_C_LNOTH : 0) | (t.rnoth ? LRB_C_RNOTH : 0) | (t.sj_checked ? LRB_C_SJ_CHECKED : 0);
        if (!t.full) continue;                                                       /* :943 */
        if (t.known) { uv_push(&known, (uint32_t)r); continue; }
        if (!t.known_site) { uv_push(&unrec, (uint32_t)r); continue; }
        if (sj_ok) {                                                                 /* :947-950 */
            mt_t m = { t.tid, t.is_rev, t.start, t.end, en, t.es, t.ee, t.es[0], t.ee[en - 1], 1, (uint32_t)nv_read.n };
            uv_push(&nv_read, (uint32_t)r); uv_push(&nv_lo, 0); uv_push(&nv_n, (uint32_t)en); uv_push(&nv_piece, (uint32_t)-1);
            if (!fold_one(&m, &upd, p)) ml_push(&upd, &m);
        } else if (p->split_trans) {                                                 /* split_trans :837-913 */
            int last = 0, k = 0, has_novel = 0, has_known = 0, i;
            for (i = 0; i <= en - 1; ++i) {
                int at_end = (i == en - 1);
                if (!at_end) { if (t.fl[i] & LRB_F_NOVEL_JUNC) has_novel = 1; else has_known = 1; }
                if (at_end || (t.fl[i] & LRB_F_UNRELIABLE)) {
                    if (has_novel && has_known && i - last >= 1) {
                        int pn = i - last + 1;
                        /* pieces keep tid=start=end=is_rev=0 (calloc'd, never set_trans_name'd) -- SURVEY Q14 */
                        mt_t m = { 0, 0, 0, 0, pn, t.es + last, t.ee + last, t.es[last], t.ee[i], 1, (uint32_t)nv_read.n };
                        uv_push(&nv_read, (uint32_t)r); uv_push(&nv_lo, (uint32_t)last); uv_push(&nv_n, (uint32_t)pn); uv_push(&nv_piece, (uint32_t)k);
                        k++;
                        if (!fold_one(&m, &upd, p)) ml_push(&upd, &m);
                    }
                    last = i + 1; has_novel = 0; has_known = 0;
                }
            }
        }
    }

    out->ex.n_reads = n; out->ex.read_idx = ridx; out->ex.tid = tid; out->ex.is_rev = rev; out->ex.exon_off = off;
    out->ex.exon_start = es; out->ex.exon_end = ee;
    out->cls = cls; out->ref_anno = ref; out->exon_flag = fl;
    out->n_known = known.n; out->known_idx = known.v ? known.v : xmalloc(4);
    out->n_unrecog = unrec.n; out->unrecog_idx = unrec.v ? unrec.v : xmalloc(4);
    out->novel.n = nv_read.n;
    out->novel.read = nv_read.v ? nv_read.v : xmalloc(4); out->novel.exon_lo = nv_lo.v ? nv_lo.v : xmalloc(4);
    out->novel.exon_n = nv_n.v ? nv_n.v : xmalloc(4); out->novel.piece = (int32_t *)(nv_piece.v ? nv_piece.v : xmalloc(4));
    merged_out(&upd, &out->updated);

    if (p->want_summary) {                                                           /* print_trans_summary :421-587 */
        setvec G = {0}, E = {0}, D = {0}, Ac = {0}, J = {0}, KG = {0};
        int64_t i; int partial = 0;
        for (i = 0; i < upd.n; ++i) {
            const mt_t *u = upd.t + i;
            uint32_t c = u->src, rr = out->novel.read[c], lo = out->novel.exon_lo[c];
            int pn = (int)out->novel.exon_n[c], j;
            int is_piece = out->novel.piece[c] >= 0;
            const uint8_t *f = fl + off[rr] + lo;
            int etid = tid[rr], erev = rev[rr];
            int ttid = u->tid, trev = u->is_rev;
            set_ent g = { ttid, A->gene[ref[rr]], 0, 0, 0, 0 };
            set_add(&G, g, 1, 0);
            partial += is_piece;
            for (j = 0; j < pn; ++j)
                if (f[j] & LRB_F_NOVEL_EXON) {
                    set_ent e = { etid, XS_(u, j), XE_(u, j), 0, pn > 1 ? ((j == 0 || j == pn - 1) ? 0 : 1) : 2, erev };
                    set_add(&E, e, 0, u->cov);
                }
            for (j = 0; j < pn - 1; ++j) if (f[j] & LRB_F_NOVEL_DON) { set_ent e = { ttid, XE_(u, j), 0, 0, 0, trev }; set_add(&D, e, 0, 0); }
            for (j = 0; j < pn - 1; ++j) if (f[j] & LRB_F_NOVEL_ACC) { set_ent e = { ttid, XS_(u, j + 1), 0, 0, 0, trev }; set_add(&Ac, e, 0, 0); }
            for (j = 0; j < pn - 1; ++j) if (f[j] & LRB_F_NOVEL_JUNC) { set_ent e = { ttid, XE_(u, j), XS_(u, j + 1), 0, 0, trev }; set_add(&J, e, 0, 1); }
        }
        int32_t *s = out->summary;
        s[LRB_S_UPD_GENES] = (int32_t)G.n; s[LRB_S_NOVEL_TRANS] = (int32_t)upd.n; s[LRB_S_NOVEL_PARTIAL] = partial;
        s[LRB_S_NOVEL_FULL] = (int32_t)upd.n - partial;
        s[LRB_S_NOVEL_EXONS] = (int32_t)E.n; s[LRB_S_NOVEL_SITES] = (int32_t)(D.n + Ac.n); s[LRB_S_NOVEL_JUNC] = (int32_t)J.n;
        mlist_t uk = {0}, ur = {0}, uu = {0}, un = {0};
        for (r = 0; r < n; ++r) {                                                    /* :501-526 */
            int en = (int)(off[r + 1] - off[r]);
            mt_t m = { tid[r], rev[r], es[off[r]], ee[off[r] + en - 1], en, es + off[r], ee + off[r], es[off[r]], ee[off[r] + en - 1], 1, (uint32_t)r };
            if (cls[r] & LRB_C_KNOWN) {
                s[LRB_S_KNOWN_TRANS]++;
                set_ent g = { tid[r], A->gene[ref[r]], 0, 0, 0, 0 }; set_add(&KG, g, 1, 0);
                if (!fold_one(&m, &uk, p)) ml_push(&uk, &m);
            } else if (cls[r] & LRB_C_KNOWN_SITE) {
                if (cls[r] & LRB_C_UNRELIABLE) { s[LRB_S_NOVEL_UNRELIABLE]++; if (!fold_one(&m, &uu, p)) ml_push(&uu, &m); }
                else { s[LRB_S_NOVEL_RELIABLE]++; if (!fold_one(&m, &ur, p)) ml_push(&ur, &m); }
            } else { s[LRB_S_UNRECOG]++; if (!fold_one(&m, &un, p)) ml_push(&un, &m); }
        }
        s[LRB_S_KNOWN_GENES] = (int32_t)KG.n; s[LRB_S_UNIQ_KNOWN] = (int32_t)uk.n;
        s[LRB_S_NOVEL_BAM] = s[LRB_S_NOVEL_RELIABLE] + s[LRB_S_NOVEL_UNRELIABLE];
        s[LRB_S_UNIQ_RELIABLE] = (int32_t)ur.n; s[LRB_S_UNIQ_UNRELIABLE] = (int32_t)uu.n; s[LRB_S_UNIQ_UNRECOG] = (int32_t)un.n;
        int64_t nb = E.n;
        int32_t *bt = xmalloc(nb * 4), *bs = xmalloc(nb * 4), *be = xmalloc(nb * 4), *bsc = xmalloc(nb * 4);
        uint8_t *bty = xmalloc(nb), *brv = xmalloc(nb);
        for (i = 0; i < nb; ++i) { bt[i] = E.v[i].tid; bs[i] = E.v[i].key1; be[i] = E.v[i].key2; bsc[i] = E.v[i].score; bty[i] = (uint8_t)E.v[i].type; brv[i] = (uint8_t)E.v[i].is_rev; }
        out->bed.n = nb; out->bed.tid = bt; out->bed.start = bs; out->bed.end = be; out->bed.score = bsc; out->bed.type = bty; out->bed.is_rev = brv;
        free(G.v); free(E.v); free(D.v); free(Ac.v); free(J.v); free(KG.v);
        free(uk.t); free(ur.t); free(uu.t); free(un.t);
    }
    free(upd.t);
    return 0;
}

void orc_free_update(lrb_update_result *r)
{
    orc_free_exon(&r->ex);
    free((void *)r->cls); free((void *)r->ref_anno); free((void *)r->exon_flag);
    free((void *)r->known_idx); free((void *)r->unrecog_idx);
    free((void *)r->novel.read); free((void *)r->novel.exon_lo); free((void *)r->novel.exon_n); free((void *)r->novel.piece);
    merged_free(&r->updated);
    free((void *)r->bed.tid); free((void *)r->bed.start); free((void *)r->bed.end); free((void *)r->bed.score);
    free((void *)r->bed.type); free((void *)r->bed.is_rev);
    memset(r, 0, sizeof *r);
}

/* uniq_trans, unique_gtf.c:73-84 */
int orc_unique(const lrb_exon_result *C, const lrb_update_params *p, lrb_unique_result *out)
{
    int64_t n = C->n_reads, r, ne = C->exon_off[n];
    memset(out, 0, sizeof *out);
    uint32_t *off = xmalloc((n + 1) * 4); memcpy(off, C->exon_off, (n + 1) * 4);
    int32_t *es = xmalloc(ne * 4), *ee = xmalloc(ne * 4), *tid = xmalloc(n * 4);
    memcpy(es, C->exon_start, ne * 4); memcpy(ee, C->exon_end, ne * 4); memcpy(tid, C->tid, n * 4);
    uint8_t *rev = xmalloc(n); memcpy(rev, C->is_rev, n);
    uint32_t *ridx = NULL;
    if (C->read_idx) { ridx = xmalloc(n * 4); memcpy(ridx, C->read_idx, n * 4); }
    mlist_t U = {0}; uvec sh = {0};
    for (r = 0; r < n; ++r) {
        int en = (int)(off[r + 1] - off[r]);
        if (en == 0) { free(off); free(es); free(ee); free(tid); free(rev); free(ridx); free(U.t); free(sh.v); return LRB_E_UNMAPPED; }
        mt_t m = { tid[r], rev[r], es[off[r]], ee[off[r] + en - 1], en, es + off[r], ee + off[r], es[off[r]], ee[off[r] + en - 1], 1, (uint32_t)r };
        if (!fold_one(&m, &U, p)) ml_push(&U, &m); else uv_push(&sh, (uint32_t)r);
    }
    out->ex.n_reads = n; out->ex.read_idx = ridx; out->ex.tid = tid; out->ex.is_rev = rev; out->ex.exon_off = off;
    out->ex.exon_start = es; out->ex.exon_end = ee;
    merged_out(&U, &out->uniq);
    out->n_shared = sh.n; out->shared_idx = sh.v ? sh.v : xmalloc(4);
    free(U.t);
    return 0;
}
void orc_free_unique(lrb_unique_result *r)
{
    orc_free_exon(&r->ex); merged_free(&r->uniq); free((void *)r->shared_idx);
    memset(r, 0, sizeof *r);
}


/* ------------------------------------------------------------------------------------------------ bam2sj
 * Restates bam2sj_core (parse_bam.c:896-924), gen_sj (:402-442), sj_sch_group (:339-351) and sj_update_group (:353-380):
 * the junction array is kept by linear search from its end + memmove, with the reference's own (lopsided) comparison. */
typedef struct { int tid, don, acc, uniq_c, multi_c; } psj_t;
static int psj_search(const psj_t *S, int n, const psj_t *x, int *hit)
{
    *hit = 0;
    for (int i = n - 1; i >= 0; --i) {
        if (S[i].tid == x->tid && S[i].don == x->don && S[i].acc == x->acc) { *hit = 1; return i; }
        else if (S[i].tid < x->tid || S[i].don < x->don || (S[i].don == x->don && S[i].acc < x->acc)) return i + 1;   /* parse_bam.c:348 */
    }
    return 0;
}
int orc_bam2sj(const lrb_batch *b, const uint8_t *is_uniq, const lrb_sj_params *p, lrb_sj *out)
{
    psj_t *S = NULL; int n = 0, m = 0;
    for (int64_t r = 0; r < b->n; ++r) {
        if (b->flag[r] & 4) continue;                                   /* bam_unmap, :909 */
        const int uq = is_uniq ? (is_uniq[r] != 0) : 0;
        if (p->pair_only && !(b->flag[r] & 2)) continue;                /* bam_is_prop / PAIR_T, :914 */
        int end = b->pos[r];                                            /* start - 1, start = pos + 1 */
        for (uint64_t k = b->cigar_off[r]; k < b->cigar_off[r + 1]; ++k) {
            const uint32_t w = b->cigar[k], op = w & 15u; const int l = (int)(w >> 4);
            if (op == 3) {
                if (l >= p->min_intron) {
                    psj_t x = {b->tid[r], end + 1, end + l, uq, 1 - uq};
                    int hit, i = psj_search(S, n, &x, &hit);
                    if (hit) { S[i].uniq_c += x.uniq_c; S[i].multi_c += x.multi_c; }
                    else {
                        if (n == m) { m = m ? 2 * m : 1024; S = (psj_t *)realloc(S, (size_t)m * sizeof *S); }
                        memmove(S + i + 1, S + i, (size_t)(n - i) * sizeof *S);
                        S[i] = x; ++n;
                    }
                }
                end += l;
            } else if (op == 0 || op == 7 || op == 8 || op == 2) end += l;
        }
    }
    int32_t *t = (int32_t *)malloc((size_t)(n ? n : 1) * 5 * sizeof(int32_t));
    for (int i = 0; i < n; ++i) { t[i] = S[i].tid; t[n + i] = S[i].don; t[2 * n + i] = S[i].acc; t[3 * n + i] = S[i].uniq_c; t[4 * n + i] = S[i].multi_c; }
    free(S);
    out->n = n; out->tid = t; out->don = t + n; out->acc = t + 2 * n; out->uniq_c = t + 3 * n; out->multi_c = t + 4 * n;
    return 0;
}
void orc_free_sj(lrb_sj *r) { free((void *)r->tid); memset(r, 0, sizeof *r); }
