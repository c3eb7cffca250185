/*
 * nb_oracle.c -- CPU ORACLE for the numbskull Gibbs / weight-learning hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product path (numbskull_b200/) never links,
 * imports or calls anything in this directory.
 *
 * It is a plain-C restatement of the reference's numba algorithm, operating on
 * the reference's own packed record arrays (numbskull/numbskulltypes.py:11-39)
 * so the same numpy buffers can be handed to the reference, to this oracle and
 * to the CUDA library.  Each function cites the reference lines it follows
 * (paths relative to the reference checkout).
 *
 * Parity pin: the random stream is MT19937 seeded by Knuth's LCG and turned
 * into 53-bit doubles exactly the way numba's runtime does it, with separate
 * "np" and "py" generators (np.random.rand in draw_sample, random.random in
 * sample_and_sgd).  With nthreads == 1 and the same seed this oracle therefore
 * reproduces the reference's count / var_value / weight_value arrays
 * bit-for-bit; tests/golden/ holds arrays produced by the real numba code
 * (tests/golden/make_golden.py) and tests/test_oracle_golden.py checks them.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#pragma pack(push, 1)
typedef struct { uint8_t isFixed; double initialValue; } nbo_weight;            /*  9 B */
typedef struct { int8_t isEvidence; int64_t initialValue; int16_t dataType;
                 int64_t cardinality; int64_t vtf_offset; } nbo_variable;        /* 27 B */
typedef struct { int16_t factorFunction; int64_t weightId; double featureValue;
                 int64_t arity; int64_t ftv_offset; } nbo_factor;                /* 34 B */
typedef struct { int64_t vid; int64_t dense_equal_to; } nbo_ftv;                 /* 16 B */
typedef struct { int64_t value; int64_t factor_index_offset;
                 int64_t factor_index_length; } nbo_vtf;                         /* 24 B */
#pragma pack(pop)

/* ------------------------------------------------------------------ */
/* MT19937 (Matsumoto & Nishimura 1998), seeded/consumed like numba.   */
/* ------------------------------------------------------------------ */
#define MT_N 624
#define MT_M 397
typedef struct { uint32_t mt[MT_N]; int idx; } nbo_mt;

void nbo_mt_seed(nbo_mt *s, uint32_t seed)
{
    for (int i = 0; i < MT_N; i++) {
        s->mt[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    s->idx = MT_N;
}

static void mt_twist(nbo_mt *s)
{
    uint32_t *mt = s->mt;
    for (int i = 0; i < MT_N; i++) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % MT_N] & 0x7fffffffu);
        uint32_t x = mt[(i + MT_M) % MT_N] ^ (y >> 1);
        if (y & 1u) x ^= 0x9908b0dfu;
        mt[i] = x;
    }
    s->idx = 0;
}

static uint32_t mt_u32(nbo_mt *s)
{
    if (s->idx >= MT_N) mt_twist(s);
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

double nbo_mt_double(nbo_mt *s)
{
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* Per-thread generator pair: np.random.* and random.* are distinct in numba. */
typedef struct { nbo_mt np; nbo_mt py; } nbo_rng;

void nbo_rng_seed(nbo_rng *r, uint32_t seed)
{
    nbo_mt_seed(&r->np, seed);
    nbo_mt_seed(&r->py, seed);
}

/* ------------------------------------------------------------------ */
/* Graph view: the arrays FactorGraph owns (factorgraph.py:30-73).     */
/* ------------------------------------------------------------------ */
typedef struct {
    const nbo_weight *weight;     int64_t n_weight;
    const nbo_variable *variable; int64_t n_var;
    const nbo_factor *factor;     int64_t n_factor;
    const nbo_ftv *fmap;          int64_t n_fmap;
    const nbo_vtf *vmap;          int64_t n_vmap;
    const int64_t *factor_index;  int64_t n_findex;
    const int64_t *cstart;        /* n_var + 1 */
    int64_t *count;
    int64_t *var_value;           /* copy 0 */
    int64_t *var_value_evid;      /* copy 0 */
    double *weight_value;         /* copy 0 */
} nbo_graph;

enum {
    F_NOOP = -1, F_IMPLY_NATURAL = 0, F_OR = 1, F_AND = 2, F_EQUAL = 3, F_ISTRUE = 4,
    F_LINEAR = 7, F_RATIO = 8, F_LOGICAL = 9, F_AND_CAT = 12, F_IMPLY_MLN = 13,
    F_OR_CAT = 14, F_EQUAL_CAT_CONST = 15, F_IMPLY_NATURAL_CAT = 16, F_IMPLY_MLN_CAT = 17,
    F_DP_CLASS_PRIOR = 18, F_DP_LF_PRIOR = 19, F_DP_LF_PROPENSITY = 20, F_DP_LF_ACCURACY = 21,
    F_DP_LF_CLASS_PROPENSITY = 22, F_DP_DEP_FIXING = 23, F_DP_DEP_REINFORCING = 24,
    F_DP_DEP_EXCLUSIVE = 25, F_DP_DEP_SIMILAR = 26, F_UFO = 30
};

/* Value of fmap slot l as eval_factor sees it: the sampled variable is forced
 * to `value`, everything else comes from the chain (inference.py:164-165 and
 * every other branch). */
static inline int64_t member_value(const nbo_ftv *fmap, int64_t l, int64_t var_samp,
                                   int64_t value, const int64_t *vals)
{
    int64_t vid = fmap[l].vid;
    return vid == var_samp ? value : vals[vid];
}

/* The three *_MLN/_CAT implication branches read the head through
 * var_value[var_copy][l] with l an fmap SLOT, not a variable id
 * (inference.py:243,277,292).  Restated as coded; out-of-range slots (numba
 * would read out of bounds) yield 0 and set *err = 2. */
static inline int64_t head_value_as_coded(const nbo_ftv *fmap, int64_t l, int64_t var_samp,
                                          int64_t value, const int64_t *vals,
                                          int64_t n_var, int *err)
{
    if (fmap[l].vid == var_samp) return value;
    if (l < 0 || l >= n_var) { if (err) *err = 2; return 0; }
    return vals[l];
}

/* inference.py:149-413 eval_factor.  Returns the factor function value as a
 * double (numba unifies the int/float returns to float64).  *err: 1 = factor
 * function not implemented (inference.py:410-413), 2 = see above. */
double nbo_eval_factor(const nbo_graph *g, int64_t factor_id, int64_t var_samp,
                       int64_t value, const int64_t *vals, int *err)
{
    const nbo_factor *fac = &g->factor[factor_id];
    const nbo_ftv *fmap = g->fmap;
    const int64_t s = fac->ftv_offset, e = s + fac->arity;
    int64_t l, v, head;

    switch (fac->factorFunction) {
    case F_NOOP:                                   /* :160-161 */
        return 0;
    case F_IMPLY_NATURAL:                          /* :162-176 */
        for (l = s; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) == 0) return 0;
        head = member_value(fmap, e - 1, var_samp, value, vals);
        return head ? 1 : -1;
    case F_OR:                                     /* :177-183 */
        for (l = s; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) == 1) return 1;
        return -1;
    case F_EQUAL:                                  /* :184-192 */
        v = member_value(fmap, s, var_samp, value, vals);
        for (l = s + 1; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) != v) return -1;
        return 1;
    case F_AND:
    case F_ISTRUE:                                 /* :193-200 */
        for (l = s; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) == 0) return -1;
        return 1;
    case F_LINEAR: {                               /* :201-211 */
        int64_t res = 0;
        head = member_value(fmap, e - 1, var_samp, value, vals);
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) == head) res++;
        return (double)res;
    }
    case F_RATIO: {                                /* :212-222 */
        int64_t res = 1;
        head = member_value(fmap, e - 1, var_samp, value, vals);
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) == head) res++;
        return log((double)res);
    }
    case F_LOGICAL:                                /* :223-231 */
        head = member_value(fmap, e - 1, var_samp, value, vals);
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) == head) return 1;
        return 0;
    case F_IMPLY_MLN:                              /* :232-246 */
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) == 0) return 1;
        head = head_value_as_coded(fmap, e - 1, var_samp, value, vals, g->n_var, err);
        return head ? 1 : 0;
    case F_AND_CAT:
    case F_EQUAL_CAT_CONST:                        /* :251-258 */
        for (l = s; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) != fmap[l].dense_equal_to) return 0;
        return 1;
    case F_OR_CAT:                                 /* :259-265 */
        for (l = s; l < e; l++)
            if (member_value(fmap, l, var_samp, value, vals) == fmap[l].dense_equal_to) return 1;
        return -1;
    case F_IMPLY_NATURAL_CAT:                      /* :266-280 */
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) != fmap[l].dense_equal_to) return 0;
        head = head_value_as_coded(fmap, e - 1, var_samp, value, vals, g->n_var, err);
        return head == fmap[e - 1].dense_equal_to ? 1 : -1;
    case F_IMPLY_MLN_CAT:                          /* :281-295 */
        for (l = s; l < e - 1; l++)
            if (member_value(fmap, l, var_samp, value, vals) != fmap[l].dense_equal_to) return 1;
        head = head_value_as_coded(fmap, e - 1, var_samp, value, vals, g->n_var, err);
        return head == fmap[e - 1].dense_equal_to ? 1 : 0;
    case F_DP_CLASS_PRIOR:                         /* :301-305 */
        return member_value(fmap, s, var_samp, value, vals) == 1 ? 1 : -1;
    case F_DP_LF_PRIOR:                            /* :306-315 */
        v = member_value(fmap, s, var_samp, value, vals);
        return v == 2 ? -1 : (v == 0 ? 0 : 1);
    case F_DP_LF_PROPENSITY: {                     /* :316-320 */
        int64_t abstain = g->variable[fmap[s].vid].cardinality - 1;
        return member_value(fmap, s, var_samp, value, vals) == abstain ? 0 : 1;
    }
    case F_DP_LF_ACCURACY: {                       /* :321-332 */
        int64_t y = member_value(fmap, s, var_samp, value, vals);
        int64_t lf = member_value(fmap, s + 1, var_samp, value, vals);
        int64_t abstain = g->variable[fmap[s + 1].vid].cardinality - 1;
        if (lf == abstain) return 0;
        return y == lf ? 1 : -1;
    }
    case F_DP_LF_CLASS_PROPENSITY: {               /* :333-347 */
        int64_t y = member_value(fmap, s, var_samp, value, vals);
        int64_t lf = member_value(fmap, s + 1, var_samp, value, vals);
        int64_t abstain = g->variable[fmap[s + 1].vid].cardinality - 1;
        if (lf == abstain) return 0;
        return y == 1 ? 1 : -1;
    }
    case F_DP_DEP_FIXING:                          /* :348-364 */
    case F_DP_DEP_REINFORCING: {                   /* :365-381 */
        int64_t y = member_value(fmap, s, var_samp, value, vals);
        int64_t l1 = member_value(fmap, s + 1, var_samp, value, vals);
        int64_t l2 = member_value(fmap, s + 2, var_samp, value, vals);
        int64_t abstain = g->variable[fmap[s + 1].vid].cardinality - 1;
        if (l1 == abstain) return l2 != 1 ? -1 : 0;
        if (fac->factorFunction == F_DP_DEP_FIXING) {
            if (l1 == 0 && l2 == 1 && y == 1) return 1;
            if (l1 == 1 && l2 == 0 && y == 0) return 1;
        } else {
            if (l1 == 0 && l2 == 0 && y == 0) return 1;
            if (l1 == 1 && l2 == 1 && y == 1) return 1;
        }
        return 0;
    }
    case F_DP_DEP_EXCLUSIVE: {                     /* :382-388 */
        int64_t l1 = member_value(fmap, s, var_samp, value, vals);
        int64_t l2 = member_value(fmap, s + 1, var_samp, value, vals);
        int64_t abstain = g->variable[fmap[s].vid].cardinality - 1;
        return (l1 == abstain || l2 == abstain) ? 0 : -1;
    }
    case F_DP_DEP_SIMILAR:                         /* :389-394 */
        return member_value(fmap, s, var_samp, value, vals) ==
               member_value(fmap, s + 1, var_samp, value, vals) ? 1 : 0;
    case F_UFO:                                    /* :399-405 */
        v = member_value(fmap, s, var_samp, value, vals);
        if (v == 0) return 0;
        return (double)member_value(fmap, s + v - 1, var_samp, value, vals);
    default:                                       /* :410-413 */
        if (err) *err = 1;
        return 0;
    }
}

/* inference.py:55-71 potential.  Note featureValue is NOT used here. */
double nbo_potential(const nbo_graph *g, int64_t var_samp, int64_t value,
                     const int64_t *vals, int *err)
{
    const nbo_variable *var = &g->variable[var_samp];
    int64_t off = var->dataType == 0 ? 0 : value;
    const nbo_vtf *vtf = &g->vmap[var->vtf_offset + off];
    double p = 0.0;
    for (int64_t k = vtf->factor_index_offset;
         k < vtf->factor_index_offset + vtf->factor_index_length; k++) {
        int64_t fid = g->factor_index[k];
        p += g->weight_value[g->factor[fid].weightId] *
             nbo_eval_factor(g, fid, var_samp, value, vals, err);
    }
    return p;
}

/* inference.py:36-52 draw_sample: exp, inclusive scan, one uniform, first
 * k with Z[k] >= z (np.argmax of an all-False mask is 0). */
static int64_t draw_sample(const nbo_graph *g, int64_t var_samp, const int64_t *vals,
                           double *Z, nbo_rng *rng, int *err)
{
    int64_t card = g->variable[var_samp].cardinality;
    for (int64_t k = 0; k < card; k++) Z[k] = exp(nbo_potential(g, var_samp, k, vals, err));
    for (int64_t k = 1; k < card; k++) Z[k] += Z[k - 1];
    double z = nbo_mt_double(&rng->np) * Z[card - 1];
    for (int64_t k = 0; k < card; k++)
        if (Z[k] >= z) return k;
    return 0;
}

/* inference.py:10-33 gibbsthread. */
int nbo_gibbsthread(const nbo_graph *g, int64_t shard, int64_t nshards, double *Z,
                    int sample_evidence, int burnin, nbo_rng *rng)
{
    int err = 0;
    int64_t nvar = g->n_var;
    int64_t start = (shard * nvar) / nshards, end = ((shard + 1) * nvar) / nshards;
    for (int64_t v = start; v < end; v++) {
        const nbo_variable *var = &g->variable[v];
        if (var->isEvidence == 4) continue;
        if (var->isEvidence == 0 || sample_evidence) {
            int64_t k = draw_sample(g, v, g->var_value, Z, rng, &err);
            g->var_value[v] = k;
            if (!burnin) {
                if (var->cardinality == 2) g->count[g->cstart[v]] += k;
                else g->count[g->cstart[v] + k] += 1;
            }
        }
    }
    return err;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* learning.py:34-43 get_factor_id_range. */
static void bucket_range(const nbo_graph *g, int64_t v, int64_t val, int64_t *s, int64_t *e)
{
    const nbo_variable *var = &g->variable[v];
    const nbo_vtf *vtf = &g->vmap[var->vtf_offset + (var->dataType == 0 ? 0 : val)];
    *s = vtf->factor_index_offset;
    *e = *s + vtf->factor_index_length;
}

/* learning.py:46-125 sample_and_sgd. */
static void sample_and_sgd(const nbo_graph *g, int64_t v, double step, int regularization,
                           double reg_param, double truncation, double *Z, int64_t *fids,
                           int learn_non_evidence, nbo_rng *rng, int *err)
{
    const nbo_variable *var = &g->variable[v];
    int64_t evidence, proposal;
    if (var->isEvidence != 1) evidence = draw_sample(g, v, g->var_value_evid, Z, rng, err);
    else evidence = var->initialValue;
    g->var_value_evid[v] = evidence;
    proposal = draw_sample(g, v, g->var_value, Z, rng, err);
    g->var_value[v] = proposal;
    if (!learn_non_evidence && var->isEvidence != 1) return;

    int64_t s0, e0, s;
    bucket_range(g, v, evidence, &s0, &e0);
    if (evidence != proposal) {
        int64_t s1, e1;
        bucket_range(g, v, proposal, &s1, &e1);
        s = (e0 - s0) + (e1 - s1);
        memcpy(fids, g->factor_index + s0, (size_t)(e0 - s0) * sizeof(int64_t));
        memcpy(fids + (e0 - s0), g->factor_index + s1, (size_t)(e1 - s1) * sizeof(int64_t));
        qsort(fids, (size_t)s, sizeof(int64_t), cmp_i64);
    } else {
        s = e0 - s0;
        memcpy(fids, g->factor_index + s0, (size_t)s * sizeof(int64_t));
    }

    int truncate = 0;
    if (regularization == 1) truncate = nbo_mt_double(&rng->py) < 1.0 / truncation;

    int64_t last = -1;
    for (int64_t i = 0; i < s; i++) {
        int64_t fid = fids[i];
        if (fid == last) continue;
        last = fid;
        int64_t wid = g->factor[fid].weightId;
        if (g->weight[wid].isFixed) continue;
        double p0 = nbo_eval_factor(g, fid, v, evidence, g->var_value_evid, err);
        double p1 = nbo_eval_factor(g, fid, v, proposal, g->var_value, err);
        double gradient = (p1 - p0) * g->factor[fid].featureValue;
        double w = g->weight_value[wid];
        if (regularization == 2) {
            w *= 1.0 / (1.0 + reg_param * step);
            w -= step * gradient;
        } else if (regularization == 1) {
            w -= step * gradient;
            if (truncate) {
                double l1delta = reg_param * step * truncation;
                if (w > 0) { w = w - l1delta; if (w < 0) w = 0; }
                else       { w = w + l1delta; if (w > 0) w = 0; }
            }
        } else {
            w -= step * gradient;
        }
        g->weight_value[wid] = w;
    }
}

/* learning.py:12-31 learnthread. */
int nbo_learnthread(const nbo_graph *g, int64_t shard, int64_t nshards, double step,
                    int regularization, double reg_param, double truncation, double *Z,
                    int64_t *fids, int learn_non_evidence, nbo_rng *rng)
{
    int err = 0;
    int64_t nvar = g->n_var;
    int64_t start = (shard * nvar) / nshards, end = ((shard + 1) * nvar) / nshards;
    for (int64_t v = start; v < end; v++) {
        if (g->variable[v].isEvidence == 4) continue;
        sample_and_sgd(g, v, step, regularization, reg_param, truncation, Z, fids,
                       learn_non_evidence, rng, &err);
    }
    return err;
}

/* ------------------------------------------------------------------ */
/* Epoch drivers (factorgraph.py:13-24 run_pool, :129-208).            */
/* Threads own contiguous variable ranges and race on shared arrays    */
/* exactly like the reference's Hogwild pool.                          */
/* ------------------------------------------------------------------ */
typedef struct {
    const nbo_graph *g; int64_t shard, nshards; int learn;
    int sample_evidence, burnin;
    double step; int regularization; double reg_param, truncation; int learn_non_evidence;
    double *Z; int64_t *fids; nbo_rng *rng; int err;
} shard_job;

static void *shard_main(void *p)
{
    shard_job *j = (shard_job *)p;
    if (j->learn)
        j->err = nbo_learnthread(j->g, j->shard, j->nshards, j->step, j->regularization,
                                 j->reg_param, j->truncation, j->Z, j->fids,
                                 j->learn_non_evidence, j->rng);
    else
        j->err = nbo_gibbsthread(j->g, j->shard, j->nshards, j->Z, j->sample_evidence,
                                 j->burnin, j->rng);
    return NULL;
}

static int run_pool(shard_job *jobs, int nthreads)
{
    int err = 0;
    if (nthreads == 1) {
        shard_main(&jobs[0]);
    } else {
        pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; i++) pthread_create(&t[i], NULL, shard_main, &jobs[i]);
        for (int i = 0; i < nthreads; i++) pthread_join(t[i], NULL);
        free(t);
    }
    for (int i = 0; i < nthreads; i++) err |= jobs[i].err;
    return err;
}

typedef struct { int nthreads; int64_t maxcard, maxbucket; double *Z; int64_t *fids;
                 nbo_rng *rng; shard_job *jobs; } nbo_pool;

nbo_pool *nbo_pool_create(const nbo_graph *g, int nthreads, uint32_t seed)
{
    nbo_pool *p = (nbo_pool *)calloc(1, sizeof(nbo_pool));
    p->nthreads = nthreads;
    p->maxcard = 1; p->maxbucket = 1;
    for (int64_t i = 0; i < g->n_var; i++)
        if (g->variable[i].cardinality > p->maxcard) p->maxcard = g->variable[i].cardinality;
    for (int64_t i = 0; i < g->n_vmap; i++)
        if (g->vmap[i].factor_index_length > p->maxbucket) p->maxbucket = g->vmap[i].factor_index_length;
    p->Z = (double *)calloc((size_t)(nthreads * p->maxcard), sizeof(double));
    p->fids = (int64_t *)calloc((size_t)(nthreads * 2 * p->maxbucket), sizeof(int64_t));
    p->rng = (nbo_rng *)calloc((size_t)nthreads, sizeof(nbo_rng));
    p->jobs = (shard_job *)calloc((size_t)nthreads, sizeof(shard_job));
    /* thread 0 is the caller's thread in the reference (threads == 1 runs
     * inline); worker threads get their own streams. */
    for (int i = 0; i < nthreads; i++) nbo_rng_seed(&p->rng[i], seed + (uint32_t)i);
    return p;
}

void nbo_pool_destroy(nbo_pool *p)
{
    if (!p) return;
    free(p->Z); free(p->fids); free(p->rng); free(p->jobs); free(p);
}

/* factorgraph.py:129-175: `epochs` sweeps of gibbsthread over the pool. */
int nbo_gibbs_epochs(const nbo_graph *g, nbo_pool *p, int64_t epochs, int sample_evidence,
                     int burnin)
{
    int err = 0;
    for (int64_t ep = 0; ep < epochs; ep++) {
        for (int i = 0; i < p->nthreads; i++) {
            shard_job *j = &p->jobs[i];
            memset(j, 0, sizeof(*j));
            j->g = g; j->shard = i; j->nshards = p->nthreads; j->learn = 0;
            j->sample_evidence = sample_evidence; j->burnin = burnin;
            j->Z = p->Z + (size_t)i * (size_t)p->maxcard; j->rng = &p->rng[i];
        }
        err |= run_pool(p->jobs, p->nthreads);
    }
    return err;
}

/* factorgraph.py:188-206: learning epochs with stepsize *= decay; returns the
 * final stepsize through *stepsize. */
int nbo_learn_epochs(const nbo_graph *g, nbo_pool *p, int64_t epochs, double *stepsize,
                     double decay, int regularization, double reg_param, double truncation,
                     int learn_non_evidence)
{
    int err = 0;
    for (int64_t ep = 0; ep < epochs; ep++) {
        for (int i = 0; i < p->nthreads; i++) {
            shard_job *j = &p->jobs[i];
            memset(j, 0, sizeof(*j));
            j->g = g; j->shard = i; j->nshards = p->nthreads; j->learn = 1;
            j->step = *stepsize; j->regularization = regularization;
            j->reg_param = reg_param; j->truncation = truncation;
            j->learn_non_evidence = learn_non_evidence;
            j->Z = p->Z + (size_t)i * (size_t)p->maxcard;
            j->fids = p->fids + (size_t)i * 2 * (size_t)p->maxbucket;
            j->rng = &p->rng[i];
        }
        err |= run_pool(p->jobs, p->nthreads);
        *stepsize *= decay;
    }
    return err;
}

/* ------------------------------------------------------------------ */
/* dataloading.py:16-81 compute_var_map (integer work, bit-exact).     */
/* ------------------------------------------------------------------ */
void nbo_compute_var_map(nbo_variable *variables, int64_t n_var, const nbo_factor *factors,
                         int64_t n_factor, const nbo_ftv *fmap, int64_t n_fmap, nbo_vtf *vmap,
                         int64_t n_vmap, int64_t *factor_index, const uint8_t *domain_mask,
                         const int64_t *factors_to_skip, int64_t n_skip)
{
    /* :20-30 implicit domains */
    for (int64_t i = 0; i < n_var; i++) {
        if (variables[i].dataType == 0 || domain_mask[i]) continue;
        for (int64_t k = 0; k < variables[i].cardinality; k++)
            vmap[variables[i].vtf_offset + k].value = k;
    }
    /* :33-38 bucket lengths from every fmap entry */
    for (int64_t j = 0; j < n_fmap; j++) {
        int64_t vid = fmap[j].vid;
        int64_t val = variables[vid].dataType == 1 ? fmap[j].dense_equal_to : 0;
        vmap[variables[vid].vtf_offset + val].factor_index_length += 1;
    }
    /* :40-46 exclusive scan */
    int64_t last_len = 0, last_off = 0;
    for (int64_t i = 0; i < n_vmap; i++) {
        vmap[i].factor_index_offset = last_off + last_len;
        last_len = vmap[i].factor_index_length;
        last_off = vmap[i].factor_index_offset;
    }
    /* :48-65 scatter factor ids in factor order */
    int64_t *offsets = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_vmap > 0 ? n_vmap : 1));
    for (int64_t i = 0; i < n_vmap; i++) offsets[i] = vmap[i].factor_index_offset;
    int64_t fts = 0;
    for (int64_t i = 0; i < n_factor; i++) {
        if (fts < n_skip && factors_to_skip[fts] == i) { fts++; continue; }
        for (int64_t j = factors[i].ftv_offset; j < factors[i].ftv_offset + factors[i].arity; j++) {
            int64_t vid = fmap[j].vid;
            int64_t val = variables[vid].dataType == 1 ? fmap[j].dense_equal_to : 0;
            int64_t b = variables[vid].vtf_offset + val;
            factor_index[offsets[b]++] = i;
        }
    }
    free(offsets);
    /* :67-81 sort + unique each bucket; offsets are not compacted */
    for (int64_t i = 0; i < n_vmap; i++) {
        int64_t off = vmap[i].factor_index_offset, len = vmap[i].factor_index_length;
        qsort(factor_index + off, (size_t)len, sizeof(int64_t), cmp_i64);
        int64_t n = 0, last = -1;
        for (int64_t k = 0; k < len; k++) {
            int64_t fid = factor_index[off + k];
            if (fid == last) continue;
            last = fid;
            factor_index[off + n++] = fid;
        }
        vmap[i].factor_index_length = n;
    }
}

/* ------------------------------------------------------------------ */
/* Exact marginals by enumeration (tiny graphs only).  Uses            */
/* eval_factor(var_samp = -1), i.e. every member read from the state,  */
/* the way SURVEY.md section 4 pins the statistical tests.             */
/* out has cstart[n_var] entries laid out like `count` / `marginals`.  */
/* Evidence variables (isEvidence == 1) are clamped iff clamp_evidence.*/
/* ------------------------------------------------------------------ */
int nbo_exact_marginals(const nbo_graph *g, int clamp_evidence, double *out, double *logZ)
{
    int64_t n = g->n_var, total = g->cstart[n];
    int64_t *state = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    double *acc = (double *)calloc((size_t)(total > 0 ? total : 1), sizeof(double));
    double Zsum = 0.0;
    int err = 0;
    for (int64_t i = 0; i < n; i++)
        if (clamp_evidence && g->variable[i].isEvidence == 1) state[i] = g->variable[i].initialValue;
    for (;;) {
        double en = 0.0;
        for (int64_t f = 0; f < g->n_factor; f++)
            en += g->weight_value[g->factor[f].weightId] *
                  nbo_eval_factor(g, f, -1, 0, state, &err);
        double w = exp(en);
        Zsum += w;
        for (int64_t i = 0; i < n; i++) {
            if (g->variable[i].cardinality == 2) acc[g->cstart[i]] += w * (double)state[i];
            else acc[g->cstart[i] + state[i]] += w;
        }
        /* odometer over the free variables */
        int64_t i = 0;
        for (; i < n; i++) {
            if (clamp_evidence && g->variable[i].isEvidence == 1) continue;
            if (++state[i] < g->variable[i].cardinality) break;
            state[i] = 0;
        }
        if (i == n) break;
    }
    for (int64_t k = 0; k < total; k++) out[k] = acc[k] / Zsum;
    if (logZ) *logZ = log(Zsum);
    free(state); free(acc);
    return err;
}

int nbo_sizeof_rng(void) { return (int)sizeof(nbo_rng); }
int nbo_sizeof_graph(void) { return (int)sizeof(nbo_graph); }
