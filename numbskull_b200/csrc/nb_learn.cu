// Weight learning sweep (learning.py:12-125 learnthread / sample_and_sgd).
//
// Per colour (split into mini-batches, see DESIGN.md "learning"): every owned
// variable samples the evidence chain and the free chain in the same pass over
// its row, then walks the row again for the per-factor gradient
// (f(proposal | free) - f(evidence | evid)) * featureValue.  Gradients and visit
// counts are REDUCED BY WEIGHT ID -- block-private shared-memory tables, flushed
// to per-block partials that the last block sums in block order -- and the
// SGD / L2-shrink / L1-truncated-gradient update is applied once per mini-batch.
// Graphs whose weight table does not fit in shared memory accumulate into a
// global table instead and apply with a separate kernel.
#include <algorithm>
#include <cmath>

#include "nb_eval.cuh"

#define NB_LEARN_SMEM_W 2048
#define NB_LEARN_MAX_BLOCKS (148 * 2)
#define NB_LEARN_THREADS 256

struct LearnArgs {
    const uint32_t *vmeta;
    const int64_t *slice_ptr;
    const uint32_t *twords;
    const int64_t *wrow_ptr;
    const uint32_t *wwords;
    const int64_t *inc_ptr;
    const uint2 *inc;
    const uint32_t *rng_id;
    const nb_val_t *vinit;
    nb_val_t *val_free;
    nb_val_t *val_evid;
    double *weight;
    const uint8_t *wfixed;
    int64_t n_trows;
    int W;
    uint64_t seed, epoch;
    double step, reg_param, truncation;
    int regularization, learn_non_evidence;
    // reduction targets
    long long *g_grad;    // [W] global fixed-point table (large-W path)
    uint32_t *g_cnt;      // [W]
    long long *p_grad;    // [blocks][W] per-block partials (shared-memory path; read as int32 by the truth-table kernels)
    uint32_t *p_cnt;
    uint32_t *done;       // completion counter
    // truth-table rows
    const int64_t *tt_ptr;
    const uint4 *tt;
    const uint32_t *tt_base;
    const uint32_t *tt_wid;   // weight id per quad (the quads themselves inline the weight VALUE for the Gibbs sweep)
    int32_t *gi_grad;     // [W] global integer table (large-W path)
};

// ---------------------------------------------------------------------------
// closed-form application of n per-visit updates (learning.py:110-125):
//   L2: each visit does w = w * s - step * g_i, s = 1 / (1 + reg_param * step),
//       so n visits give w * s^n - step * sum_i g_i s^(n-i); the g_i are spread
//       evenly over the batch, i.e. sum_i g_i s^(n-i) ~= G * (1 - s^n) / (n (1 - s)).
//   L1: w -= step * G, then the m truncating visits' soft-threshold, merged.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double nb_apply_update(double w, double G, uint32_t cnt, int regularization, double step,
                                                  double reg_param, double truncation)
{
    if (regularization == 2) {
        if (cnt == 0) return w - step * G;
        double s = 1.0 / (1.0 + reg_param * step);
        double sn = pow(s, (double)cnt);
        double geo = (s < 1.0) ? (1.0 - sn) / ((double)cnt * (1.0 - s)) : 1.0;
        return w * sn - step * G * geo;
    }
    w -= step * G;
    if (regularization == 1 && cnt > 0) {
        double l1 = reg_param * step * truncation * (double)cnt;
        w = w > 0.0 ? fmax(0.0, w - l1) : fmin(0.0, w + l1);
    }
    return w;
}

// Generic rows accumulate gradients in 64-bit fixed point (2^-20 units): integer addition is
// associative, so per-weight sums do not depend on the order in which threads, blocks or atomics
// land -- the reduction by weight id is deterministic without sorting the incidences.
#define NB_GRAD_UNIT 1048576.0
typedef long long nb_fix_t;
template <bool SMEM>
struct GradSink {
    nb_fix_t *grad;
    uint32_t *cnt;
    __device__ __forceinline__ void add(uint32_t wid, double g, uint32_t c)
    {
        if (g != 0.0) atomicAdd((unsigned long long *)(grad + wid), (unsigned long long)__double2ll_rn(g * NB_GRAD_UNIT));
        if (c) atomicAdd(cnt + wid, c);
    }
};

// second pass over a row: gradient of every visited incidence with a learnable weight
template <bool WIDE, bool SMEM>
__device__ inline void nb_row_gradient(const NbRow &r, int len, uint32_t self, uint32_t meta, int ev, int prop,
                                       const LearnArgs &a, uint32_t cnt_inc, GradSink<SMEM> &sink,
                                       int first_inc, int inc_stride, const uint2 *inc_list, int n_inc)
{
    const bool cat = NB_META_DTYPE(meta) == 1;
    if (inc_list == nullptr) {
        int pos = 0, cur = -1;
        while (pos < len) {
            NbHdr h = nb_read_hdr<WIDE>(r, pos);
            if (h.code == C_MARK) cur = (int)h.wid;
            else if (!h.fixed && (!cat || cur == ev || cur == prop)) {
                int mpos = nb_member_pos<WIDE>(h, pos);
                double f1 = nb_eval_incidence(r, h, mpos, self, prop, a.val_free);
                double f0 = nb_eval_incidence(r, h, mpos, self, ev, a.val_evid);
                double feat = h.feat ? nb_read_feature<WIDE>(r, h, pos) : 1.0;
                sink.add(h.wid, (f1 - f0) * feat, cnt_inc);
            }
            pos += nb_inc_words<WIDE>(h);
        }
    } else {
        for (int i = first_inc; i < n_inc; i += inc_stride) {
            uint2 ent = inc_list[i];
            int pos = (int)ent.x, cur = (int)ent.y;
            NbHdr h = nb_read_hdr<WIDE>(r, pos);
            if (h.fixed || (cat && cur != ev && cur != prop)) continue;
            int mpos = nb_member_pos<WIDE>(h, pos);
            double f1 = nb_eval_incidence(r, h, mpos, self, prop, a.val_free);
            double f0 = nb_eval_incidence(r, h, mpos, self, ev, a.val_evid);
            double feat = h.feat ? nb_read_feature<WIDE>(r, h, pos) : 1.0;
            sink.add(h.wid, (f1 - f0) * feat, cnt_inc);
        }
    }
}

// ---------------------------------------------------------------------------
// block epilogue of the shared-memory path: flush the block's table, and let
// the last block to finish sum the partials in block order and apply the update
// ---------------------------------------------------------------------------
template <class T>
__device__ inline void nb_flush_and_apply(const LearnArgs &a, T *s_grad, uint32_t *s_cnt, double unit)
{
    __shared__ bool s_last;
    __syncthreads();
    T *part = reinterpret_cast<T *>(a.p_grad);
    T *pg = part + (size_t)blockIdx.x * a.W;
    uint32_t *pc = a.p_cnt + (size_t)blockIdx.x * a.W;
    for (int w = threadIdx.x; w < a.W; w += blockDim.x) { pg[w] = s_grad[w]; pc[w] = s_cnt[w]; }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int w = threadIdx.x; w < a.W; w += blockDim.x) {
        if (a.wfixed[w]) continue;
        double G = 0.0;
        uint32_t n = 0;
        T Gi = 0;
        for (unsigned b = 0; b < gridDim.x; b++) {      // integer sums: exact, order-independent
            Gi += __ldcg(part + (size_t)b * a.W + w);
            n += __ldcg(a.p_cnt + (size_t)b * a.W + w);
        }
        G = (double)Gi * unit;
        if (Gi != 0 || n != 0)
            a.weight[w] = nb_apply_update(a.weight[w], G, n, a.regularization, a.step, a.reg_param, a.truncation);
    }
    if (threadIdx.x == 0) *a.done = 0;
}

// ---------------------------------------------------------------------------
// thread path
// ---------------------------------------------------------------------------
template <bool WIDE, bool SMEM>
__global__ void __launch_bounds__(NB_LEARN_THREADS) k_learn_thread(LearnArgs a, int beg0, int end0, int beg1, int end1)
{
    extern __shared__ unsigned char s_raw[];
    nb_fix_t *s_grad = (nb_fix_t *)s_raw;
    uint32_t *s_cnt = (uint32_t *)(s_raw + sizeof(nb_fix_t) * (size_t)(SMEM ? a.W : 0));
    if (SMEM) {
        for (int w = threadIdx.x; w < a.W; w += blockDim.x) { s_grad[w] = 0; s_cnt[w] = 0u; }
        __syncthreads();
    }
    GradSink<SMEM> sink{SMEM ? s_grad : a.g_grad, SMEM ? s_cnt : a.g_cnt};

    // two id ranges per launch: the colour's FAST-class rows and its GEN-class rows
    const int64_t n0 = end0 - beg0, ntot = n0 + (end1 - beg1);
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < ntot; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nid = it < n0 ? beg0 + it : beg1 + (it - n0);
        const uint32_t meta = a.vmeta[nid];
        const int evid = NB_META_EVID(meta);
        if (!NB_META_VALID(meta) || evid == 4) continue;                   // learning.py:24-26
        NbRow r = nb_thread_row(a.twords, a.slice_ptr, nid);
        const int len = NB_META_ROWLEN(meta);
        const uint32_t self = (uint32_t)nid, id = a.rng_id[nid];
        int ev;
        if (evid != 1) {                                                    // :53-57
            NbUniforms rng(id, a.epoch, NB_TAG_EVID, a.seed);
            ev = nb_sample_row<WIDE>(r, len, self, meta, a.val_evid, a.weight, rng);
        } else {
            ev = (int)a.vinit[nid];                                         // :60-61
        }
        a.val_evid[nid] = (nb_val_t)ev;                                     // :63
        NbUniforms rng(id, a.epoch, NB_TAG_FREE, a.seed);
        int prop = nb_sample_row<WIDE>(r, len, self, meta, a.val_free, a.weight, rng);   // :65-67
        a.val_free[nid] = (nb_val_t)prop;                                   // :69
        if (!a.learn_non_evidence && evid != 1) continue;                   // :70-71
        uint32_t cnt_inc = 1;
        if (a.regularization == 1) {                                        // :90
            NbUniforms tr(id, a.epoch, NB_TAG_TRUNC, a.seed);
            cnt_inc = tr.next() < 1.0 / a.truncation ? 1u : 0u;
        } else if (a.regularization != 2) cnt_inc = 0;
        nb_row_gradient<WIDE, SMEM>(r, len, self, meta, ev, prop, a, cnt_inc, sink, 0, 1, nullptr, 0);
    }
    if (SMEM) nb_flush_and_apply<nb_fix_t>(a, s_grad, s_cnt, 1.0 / NB_GRAD_UNIT);
}

// ---------------------------------------------------------------------------
// warp path: one long row per warp
// ---------------------------------------------------------------------------
// per-value energies of a warp row into e[4] (dataType 0, card <= 4) or the shared array se
template <bool WIDE>
__device__ inline void nb_warp_energies_l(const NbRow &r, const uint2 *inc, int n_inc, uint32_t self, uint32_t meta,
                                          const nb_val_t *__restrict__ vals, const double *__restrict__ weight,
                                          double e[4], double *se)
{
    const int lane = threadIdx.x & 31, card = NB_META_CARD(meta);
    const bool small = NB_META_DTYPE(meta) == 0 && card <= 4;
    if (!small) {
        for (int k = lane; k < card; k += 32) se[k] = 0.0;
        __syncwarp();
    }
    for (int i = lane; i < n_inc; i += 32) {
        uint2 ent = inc[i];
        int pos = (int)ent.x;
        NbHdr h = nb_read_hdr<WIDE>(r, pos);
        int mpos = nb_member_pos<WIDE>(h, pos);
        double w = weight[h.wid];
        if (small) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < card) e[k] += w * nb_eval_incidence(r, h, mpos, self, k, vals);
        } else if (NB_META_DTYPE(meta) == 0) {
            for (int k = 0; k < card; k++) atomicAdd(&se[k], w * nb_eval_incidence(r, h, mpos, self, k, vals));
        } else {
            atomicAdd(&se[(int)ent.y], w * nb_eval_incidence(r, h, mpos, self, (int)ent.y, vals));
        }
    }
    if (small) {
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = nb_warp_sum(e[k]);
    } else {
        __syncwarp();
    }
}

template <bool WIDE>
__device__ inline int nb_warp_sample_l(const NbRow &r, const uint2 *inc, int n_inc, uint32_t self, uint32_t meta,
                                       const nb_val_t *__restrict__ vals, const double *__restrict__ weight, double *se,
                                       NbUniforms &rng)
{
    const int card = NB_META_CARD(meta);
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    nb_warp_energies_l<WIDE>(r, inc, n_inc, self, meta, vals, weight, e, se);
    int k;
    if (NB_META_DTYPE(meta) == 0 && card <= 4) k = nb_draw_small(e, card, rng.next());
    else {
        NbReservoir res;
        for (int j = 0; j < card; j++)
            if (res.add(se[j], 1.0, rng.next32())) res.pick = j;
        k = res.pick;
    }
    __syncwarp();
    return k;
}

#define NB_LWARPS (NB_LEARN_THREADS / 32)

template <bool WIDE, bool SMEM>
__global__ void __launch_bounds__(NB_LEARN_THREADS) k_learn_warp(LearnArgs a, int wbeg, int wend)
{
    extern __shared__ unsigned char s_raw[];
    __shared__ double s_e[NB_LWARPS][NB_MAX_CARD + 1];
    nb_fix_t *s_grad = (nb_fix_t *)s_raw;
    uint32_t *s_cnt = (uint32_t *)(s_raw + sizeof(nb_fix_t) * (size_t)(SMEM ? a.W : 0));
    if (SMEM) {
        for (int w = threadIdx.x; w < a.W; w += blockDim.x) { s_grad[w] = 0; s_cnt[w] = 0u; }
        __syncthreads();
    }
    GradSink<SMEM> sink{SMEM ? s_grad : a.g_grad, SMEM ? s_cnt : a.g_cnt};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int64_t wr = (int64_t)wbeg + blockIdx.x * (int64_t)NB_LWARPS + warp; wr < wend;
         wr += (int64_t)gridDim.x * NB_LWARPS) {
        const int64_t nid = a.n_trows + wr;
        const uint32_t meta = a.vmeta[nid];
        const int evid = NB_META_EVID(meta);
        if (!NB_META_VALID(meta) || evid == 4) continue;
        NbRow r = nb_warp_row(a.wwords, a.wrow_ptr, wr);
        const uint2 *inc = a.inc + a.inc_ptr[wr];
        const int n_inc = (int)(a.inc_ptr[wr + 1] - a.inc_ptr[wr]);
        const uint32_t self = (uint32_t)nid, id = a.rng_id[nid];
        int ev;
        if (evid != 1) {
            NbUniforms rng(id, a.epoch, NB_TAG_EVID, a.seed);
            ev = nb_warp_sample_l<WIDE>(r, inc, n_inc, self, meta, a.val_evid, a.weight, s_e[warp], rng);
        } else {
            ev = (int)a.vinit[nid];
        }
        NbUniforms rng(id, a.epoch, NB_TAG_FREE, a.seed);
        int prop = nb_warp_sample_l<WIDE>(r, inc, n_inc, self, meta, a.val_free, a.weight, s_e[warp], rng);
        if (lane == 0) { a.val_evid[nid] = (nb_val_t)ev; a.val_free[nid] = (nb_val_t)prop; }
        if (!a.learn_non_evidence && evid != 1) continue;
        uint32_t cnt_inc = 1;
        if (a.regularization == 1) {
            NbUniforms tr(id, a.epoch, NB_TAG_TRUNC, a.seed);
            cnt_inc = tr.next() < 1.0 / a.truncation ? 1u : 0u;
        } else if (a.regularization != 2) cnt_inc = 0;
        nb_row_gradient<WIDE, SMEM>(r, 0, self, meta, ev, prop, a, cnt_inc, sink, lane, 32, inc, n_inc);
    }
    if (SMEM) nb_flush_and_apply<nb_fix_t>(a, s_grad, s_cnt, 1.0 / NB_GRAD_UNIT);
}

// ---------------------------------------------------------------------------
// truth-table rows (Boolean variable, arity <= 3, unit featureValue): one warp per
// SELL slice with a uniform trip count.  f(k) = f(0) + k (f(1) - f(0)) comes from
// the two tables, so the gradient is an INTEGER; the per-weight sums are exact and
// independent of the order of accumulation (warp REDUX -> shared int table ->
// per-block partials summed in block order, or an integer global table).
// ---------------------------------------------------------------------------
struct GradSinkI {
    int32_t *grad;
    uint32_t *cnt;
    __device__ __forceinline__ void add(uint32_t wid, int g, uint32_t c)
    {
        if (g) atomicAdd(grad + wid, g);
        if (c) atomicAdd(cnt + wid, c);
    }
};


template <bool SMEM>
__global__ void __launch_bounds__(NB_LEARN_THREADS) k_learn_tt(LearnArgs a, int beg, int end, uint32_t kfree,
                                                               uint32_t kevid, uint32_t ktrunc)
{
    extern __shared__ unsigned char s_raw[];
    int32_t *s_grad = (int32_t *)s_raw;
    uint32_t *s_cnt = (uint32_t *)(s_raw + sizeof(int32_t) * (size_t)(SMEM ? a.W : 0));
    if (SMEM) {
        for (int w = threadIdx.x; w < a.W; w += blockDim.x) { s_grad[w] = 0; s_cnt[w] = 0u; }
        __syncthreads();
    }
    GradSinkI sink{SMEM ? s_grad : a.gi_grad, SMEM ? s_cnt : a.g_cnt};
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = blockIdx.x * (int64_t)NB_LWARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * NB_LWARPS;
    const nb_val_t *__restrict__ vF = a.val_free;
    const nb_val_t *__restrict__ vE = a.val_evid;
    const double *__restrict__ weight = a.weight;

    for (int64_t s = (beg >> 5) + warp_global; s < (((int64_t)end + 31) >> 5); s += n_warps) {
        const int64_t nid = (s << 5) + lane;
        const uint32_t meta = a.vmeta[nid];
        const uint32_t rid = a.rng_id[nid];
        const int evid = NB_META_EVID(meta);
        const bool valid = nid >= beg && nid < end && NB_META_VALID(meta) && evid != 4;   // learning.py:24-26
        const int64_t q0 = a.tt_ptr[s];
        const int n = (int)((a.tt_ptr[s + 1] - q0) >> 5);
        const uint4 *qp = a.tt + q0 + lane;
        const uint32_t *bp = a.tt_base + q0 + lane;
        const uint32_t *wp = a.tt_wid + q0 + lane;

        // ---- pass 1: e1 - e0 under both chains (current weights: gathered by id, the inlined
        //      values are only refreshed for the Gibbs sweep) ----
        double dF = 0.0, dE = 0.0;
        for (int j = 0; j < n; j += 2) {
            uint4 q[2];
            uint32_t wid[2];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                q[t] = (j + t < n) ? __ldg(qp + (size_t)(j + t) * 32)
                                   : make_uint4((uint32_t)nid, (uint32_t)nid, NB_TT_NEUTRAL | NB_TT_FIXED_BIT, 0u);
                wid[t] = (j + t < n) ? __ldg(wp + (size_t)(j + t) * 32) : 0u;
            }
            int xf[2][2], xe[2][2];
            double w[2];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                xf[t][0] = vF[q[t].x]; xf[t][1] = vF[q[t].y];
                xe[t][0] = vE[q[t].x]; xe[t][1] = vE[q[t].y];
                w[t] = weight[wid[t]];
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                dF = fma(w[t], (double)((int)((q[t].z >> (3 * nb_tt_index(xf[t][0], xf[t][1]))) & 7u) - 2), dF);
                dE = fma(w[t], (double)((int)((q[t].z >> (3 * nb_tt_index(xe[t][0], xe[t][1]))) & 7u) - 2), dE);
            }
        }
        // ---- both samples (learning.py:53-69) ----
        int ev;
        if (evid != 1) {
            const double u = nb_philox2x32_u53(rid, (uint32_t)a.epoch, kevid);
            ev = u <= (double)(1.0f / (1.0f + __expf((float)dE))) ? 0 : 1;
        } else {
            ev = (int)a.vinit[nid];
        }
        const double uf = nb_philox2x32_u53(rid, (uint32_t)a.epoch, kfree);
        const int prop = uf <= (double)(1.0f / (1.0f + __expf((float)dF))) ? 0 : 1;
        if (valid) { a.val_evid[nid] = (nb_val_t)ev; a.val_free[nid] = (nb_val_t)prop; }
        const bool active = valid && (a.learn_non_evidence || evid == 1);        // :70-71
        uint32_t cinc = 1;
        if (a.regularization == 1)                                                // :90
            cinc = nb_philox2x32_u53(rid, (uint32_t)a.epoch, ktrunc) < 1.0 / a.truncation ? 1u : 0u;
        else if (a.regularization != 2) cinc = 0;

        // ---- pass 2: integer gradient of every learnable incidence (:97-125) ----
        if (__ballot_sync(FULL, active) == 0u) continue;
        for (int j = 0; j < n; j++) {
            const uint4 q = __ldg(qp + (size_t)j * 32);
            const uint32_t b = __ldg(bp + (size_t)j * 32);
            const uint32_t wid = __ldg(wp + (size_t)j * 32);
            const int iF = nb_tt_index(vF[q.x], vF[q.y]), iE = nb_tt_index(vE[q.x], vE[q.y]);
            // this variable's own slots read its just-written values; the tables ignore them
            const int fF = ((int)((b >> (2 * iF)) & 3u) - 1) + prop * ((int)((q.z >> (3 * iF)) & 7u) - 2);
            const int fE = ((int)((b >> (2 * iE)) & 3u) - 1) + ev * ((int)((q.z >> (3 * iE)) & 7u) - 2);
            const bool contrib = active && !(q.z & NB_TT_FIXED_BIT);
            const unsigned who = __ballot_sync(FULL, contrib);
            if (who == 0u) continue;
            const int leader = __ffs(who) - 1;
            const uint32_t w0 = __shfl_sync(FULL, wid, leader);
            if (__all_sync(FULL, !contrib || wid == w0)) {
                const int G = __reduce_add_sync(FULL, contrib ? fF - fE : 0);
                const unsigned Cn = __reduce_add_sync(FULL, contrib ? cinc : 0u);
                if (lane == leader) sink.add(w0, G, Cn);
            } else if (contrib) {
                sink.add(wid, fF - fE, cinc);
            }
        }
    }
    if (SMEM) nb_flush_and_apply<int32_t>(a, s_grad, s_cnt, 1.0);
}

// Same algorithm, one WARP per row with the lanes striding over the row's quads: used when the
// rows are long (data-programming models: a label variable with ~100 labelling functions), where
// the mini-batches are too small to fill the GPU with one thread per row.
template <bool SMEM>
__global__ void __launch_bounds__(NB_LEARN_THREADS) k_learn_tt_row(LearnArgs a, int beg, int end, uint32_t kfree,
                                                                   uint32_t kevid, uint32_t ktrunc)
{
    extern __shared__ unsigned char s_raw[];
    int32_t *s_grad = (int32_t *)s_raw;
    uint32_t *s_cnt = (uint32_t *)(s_raw + sizeof(int32_t) * (size_t)(SMEM ? a.W : 0));
    if (SMEM) {
        for (int w = threadIdx.x; w < a.W; w += blockDim.x) { s_grad[w] = 0; s_cnt[w] = 0u; }
        __syncthreads();
    }
    GradSinkI sink{SMEM ? s_grad : a.gi_grad, SMEM ? s_cnt : a.g_cnt};
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = blockIdx.x * (int64_t)NB_LWARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * NB_LWARPS;
    const nb_val_t *__restrict__ vF = a.val_free;
    const nb_val_t *__restrict__ vE = a.val_evid;
    const double *__restrict__ weight = a.weight;

    for (int64_t nid = (int64_t)beg + warp_global; nid < end; nid += n_warps) {
        const uint32_t meta = a.vmeta[nid];
        const int evid = NB_META_EVID(meta);
        if (!NB_META_VALID(meta) || evid == 4) continue;                         // learning.py:24-26
        const uint32_t rid = a.rng_id[nid];
        const int64_t s = nid >> 5;
        const int64_t q0 = a.tt_ptr[s];
        const int n = (int)((a.tt_ptr[s + 1] - q0) >> 5);
        const uint4 *qp = a.tt + q0 + (nid & 31);
        const uint32_t *bp = a.tt_base + q0 + (nid & 31);
        const uint32_t *wp = a.tt_wid + q0 + (nid & 31);
        double dF = 0.0, dE = 0.0;
        for (int j = lane; j < n; j += 32) {
            const uint4 q = __ldg(qp + (size_t)j * 32);
            const double w = weight[__ldg(wp + (size_t)j * 32)];
            dF = fma(w, (double)((int)((q.z >> (3 * nb_tt_index(vF[q.x], vF[q.y]))) & 7u) - 2), dF);
            dE = fma(w, (double)((int)((q.z >> (3 * nb_tt_index(vE[q.x], vE[q.y]))) & 7u) - 2), dE);
        }
        dF = nb_warp_sum(dF);
        dE = nb_warp_sum(dE);
        int ev;
        if (evid != 1) {
            const double u = nb_philox2x32_u53(rid, (uint32_t)a.epoch, kevid);
            ev = u <= (double)(1.0f / (1.0f + __expf((float)dE))) ? 0 : 1;
        } else {
            ev = (int)a.vinit[nid];
        }
        const double uf = nb_philox2x32_u53(rid, (uint32_t)a.epoch, kfree);
        const int prop = uf <= (double)(1.0f / (1.0f + __expf((float)dF))) ? 0 : 1;
        __syncwarp();
        if (lane == 0) { a.val_evid[nid] = (nb_val_t)ev; a.val_free[nid] = (nb_val_t)prop; }
        if (!a.learn_non_evidence && evid != 1) continue;                        // :70-71
        uint32_t cinc = 1;
        if (a.regularization == 1)
            cinc = nb_philox2x32_u53(rid, (uint32_t)a.epoch, ktrunc) < 1.0 / a.truncation ? 1u : 0u;
        else if (a.regularization != 2) cinc = 0;
        for (int j = lane; j < n; j += 32) {
            const uint4 q = __ldg(qp + (size_t)j * 32);
            if (q.z & NB_TT_FIXED_BIT) continue;
            const uint32_t b = __ldg(bp + (size_t)j * 32);
            // slots that point at the variable itself are ignored by the tables
            const int iF = nb_tt_index(vF[q.x], vF[q.y]), iE = nb_tt_index(vE[q.x], vE[q.y]);
            const int fF = ((int)((b >> (2 * iF)) & 3u) - 1) + prop * ((int)((q.z >> (3 * iF)) & 7u) - 2);
            const int fE = ((int)((b >> (2 * iE)) & 3u) - 1) + ev * ((int)((q.z >> (3 * iE)) & 7u) - 2);
            sink.add(__ldg(wp + (size_t)j * 32), fF - fE, cinc);
        }
    }
    if (SMEM) nb_flush_and_apply<int32_t>(a, s_grad, s_cnt, 1.0);
}

__global__ void k_apply_global_int(LearnArgs a)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.W) return;
    int G = a.gi_grad[w];
    uint32_t n = a.g_cnt[w];
    if (G == 0 && n == 0u) return;
    a.gi_grad[w] = 0;
    a.g_cnt[w] = 0u;
    if (a.wfixed[w]) return;
    a.weight[w] = nb_apply_update(a.weight[w], (double)G, n, a.regularization, a.step, a.reg_param, a.truncation);
}

// large-W path: apply the global table and clear it
__global__ void k_apply_global(LearnArgs a)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.W) return;
    long long G = a.g_grad[w];
    uint32_t n = a.g_cnt[w];
    if (G == 0 && n == 0u) return;
    a.g_grad[w] = 0;
    a.g_cnt[w] = 0u;
    if (a.wfixed[w]) return;
    a.weight[w] = nb_apply_update(a.weight[w], (double)G * (1.0 / NB_GRAD_UNIT), n, a.regularization, a.step, a.reg_param,
                                  a.truncation);
}

// visits per weight of one colour (upper bound: every incidence of a learnable row)
struct RowRanges {
    int beg[4], end[4];    // thread-row id ranges (PAIR, FAST, CAT, GEN)
    int wbeg, wend;        // warp rows
};

template <bool WIDE>
__global__ void k_visit_histogram(LearnArgs a, RowRanges rr, uint32_t *hist)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t nid = -1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t n = rr.end[k] - rr.beg[k];
        if (nid < 0 && i >= 0 && i < n) nid = rr.beg[k] + i;
        if (nid < 0) i -= n;
    }
    if (nid < 0) {
        if (i >= rr.wend - rr.wbeg) return;
        nid = a.n_trows + rr.wbeg + i;
    }
    const uint32_t meta = a.vmeta[nid];
    const int evid = NB_META_EVID(meta);
    if (!NB_META_VALID(meta) || evid == 4) return;
    if (!a.learn_non_evidence && evid != 1) return;
    NbRow r = nid < a.n_trows ? nb_thread_row(a.twords, a.slice_ptr, nid) : nb_warp_row(a.wwords, a.wrow_ptr, nid - a.n_trows);
    int len = nid < a.n_trows ? NB_META_ROWLEN(meta)
                              : (int)(a.wrow_ptr[nid - a.n_trows + 1] - a.wrow_ptr[nid - a.n_trows]);
    int pos = 0;
    while (pos < len) {
        NbHdr h = nb_read_hdr<WIDE>(r, pos);
        if (h.code != C_MARK && !h.fixed) atomicAdd(hist + h.wid, 1u);
        pos += nb_inc_words<WIDE>(h);
    }
}

__global__ void k_max_u32(const uint32_t *x, int n, uint32_t *out)
{
    uint32_t m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, x[i]);
    atomicMax(out, m);
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
static LearnArgs learn_args(nb_graph *g)
{
    LearnArgs a;
    memset(&a, 0, sizeof(a));
    a.vmeta = g->d_vmeta; a.slice_ptr = g->d_slice_ptr; a.twords = g->d_twords;
    a.wrow_ptr = g->d_wrow_ptr; a.wwords = g->d_wwords; a.inc_ptr = g->d_inc_ptr; a.inc = g->d_inc;
    a.rng_id = g->d_rng_id; a.vinit = g->d_vinit; a.val_free = g->d_val[0]; a.val_evid = g->d_val[1];
    a.weight = g->d_weight; a.wfixed = g->d_wfixed; a.n_trows = g->n_trows; a.W = (int)g->W;
    a.g_grad = g->d_grad; a.g_cnt = g->d_nvis; a.p_grad = g->d_gpart; a.p_cnt = g->d_npart; a.done = g->d_done;
    a.tt_ptr = g->d_tt_ptr; a.tt = g->d_tt; a.tt_base = g->d_tt_base; a.tt_wid = g->d_tt_wid; a.gi_grad = g->d_gradi;
    return a;
}

static int ensure_learn_buffers(nb_graph *g, bool smem)
{
    if (!g->d_done) NB_TRY(nb_alloc(g, &g->d_done, 4));
    if (!g->d_grad) {
        NB_TRY(nb_alloc(g, &g->d_grad, (size_t)g->W));
        NB_TRY(nb_alloc(g, &g->d_gradi, (size_t)g->W));
        NB_TRY(nb_alloc(g, &g->d_nvis, (size_t)g->W));
    }
    if (smem && !g->d_gpart) {
        NB_TRY(nb_alloc(g, &g->d_gpart, (size_t)NB_LEARN_MAX_BLOCKS * (size_t)g->W));
        NB_TRY(nb_alloc(g, &g->d_npart, (size_t)NB_LEARN_MAX_BLOCKS * (size_t)g->W));
    }
    return NB_OK;
}

// max over weights of the gradient visits one sweep of colour c can make
static int color_visit_bounds(nb_graph *g, LearnArgs a, std::vector<int64_t> &out)
{
    out.assign((size_t)g->n_colors, 0);
    uint32_t *d_max;
    NB_TRY(nb_alloc(g, &d_max, 1));
    for (int c = 0; c < g->n_colors; c++) {
        const NbColorRange &cr = g->colors[(size_t)c];
        RowRanges rr;
        rr.beg[0] = cr.p_beg; rr.end[0] = cr.p_end; rr.beg[1] = cr.f_beg; rr.end[1] = cr.f_end;
        rr.beg[2] = cr.c_beg; rr.end[2] = cr.c_end; rr.beg[3] = cr.t_beg; rr.end[3] = cr.t_end;
        rr.wbeg = cr.w_beg; rr.wend = cr.w_end;
        int64_t n = (cr.p_end - cr.p_beg) + (cr.f_end - cr.f_beg) + (cr.c_end - cr.c_beg) + (cr.t_end - cr.t_beg) +
                    (cr.w_end - cr.w_beg);
        if (n == 0) continue;
        NB_CUDA(cudaMemsetAsync(g->d_nvis, 0, (size_t)g->W * 4, g->stream));
        NB_CUDA(cudaMemsetAsync(d_max, 0, 4, g->stream));
        unsigned grid = (unsigned)((n + 255) / 256);
        if (g->wide) k_visit_histogram<true><<<grid, 256, 0, g->stream>>>(a, rr, g->d_nvis);
        else k_visit_histogram<false><<<grid, 256, 0, g->stream>>>(a, rr, g->d_nvis);
        k_max_u32<<<64, 256, 0, g->stream>>>(g->d_nvis, (int)g->W, d_max);
        uint32_t m = 0;
        NB_CUDA(cudaMemcpyAsync(&m, d_max, 4, cudaMemcpyDeviceToHost, g->stream));
        NB_CUDA(cudaStreamSynchronize(g->stream));
        out[(size_t)c] = m;
    }
    NB_CUDA(cudaMemsetAsync(g->d_nvis, 0, (size_t)g->W * 4, g->stream));
    return NB_OK;
}

template <bool WIDE, bool SMEM>
static int launch_learn_range(nb_graph *g, const LearnArgs &a, int pb, int pe, int fb, int fe, int cb, int ce, int tb, int te,
                              int wb, int we, bool long_rows)
{
    const size_t smem_tt = SMEM ? (size_t)g->W * 8 : 0;     // int32 sums + counts
    const size_t smem = SMEM ? (size_t)g->W * 12 : 0;       // 64-bit fixed-point sums + counts
    const unsigned apply_grid = (unsigned)((g->W + 255) / 256);
    const int tt_range[2][2] = {{pb, pe}, {fb, fe}};                // PAIR rows, then FAST rows: both have TT quads
    for (int pass = 0; pass < 2; pass++) {
        const int rb = tt_range[pass][0], re = tt_range[pass][1];
        if (re <= rb) continue;
        const uint32_t kf = nb_fold_key(a.seed, a.epoch, NB_TAG_FREE), ke = nb_fold_key(a.seed, a.epoch, NB_TAG_EVID),
                       kt = nb_fold_key(a.seed, a.epoch, NB_TAG_TRUNC);
        if (long_rows) {     // one warp per row
            int64_t need = ((int64_t)(re - rb) + NB_LWARPS - 1) / NB_LWARPS;
            unsigned grid = (unsigned)std::min<int64_t>(need, SMEM ? NB_LEARN_MAX_BLOCKS : 148 * 16);
            k_learn_tt_row<SMEM><<<grid, NB_LEARN_THREADS, smem_tt, g->stream>>>(a, rb, re, kf, ke, kt);
        } else {             // one thread per row, one warp per SELL slice
            int64_t need = ((((int64_t)re + 31) >> 5) - (rb >> 5) + NB_LWARPS - 1) / NB_LWARPS;
            unsigned grid = (unsigned)std::min<int64_t>(need, SMEM ? NB_LEARN_MAX_BLOCKS : 148 * 16);
            k_learn_tt<SMEM><<<grid, NB_LEARN_THREADS, smem_tt, g->stream>>>(a, rb, re, kf, ke, kt);
        }
        g->launches++;
        if (!SMEM) { k_apply_global_int<<<apply_grid, 256, 0, g->stream>>>(a); g->launches++; }
    }
    if (te > tb || ce > cb) {
        int64_t need = ((int64_t)(te - tb) + (ce - cb) + NB_LEARN_THREADS - 1) / NB_LEARN_THREADS;
        unsigned grid = (unsigned)std::min<int64_t>(need, SMEM ? NB_LEARN_MAX_BLOCKS : (1 << 30));
        k_learn_thread<WIDE, SMEM><<<grid, NB_LEARN_THREADS, smem, g->stream>>>(a, cb, ce, tb, te);
        g->launches++;
        if (!SMEM) { k_apply_global<<<apply_grid, 256, 0, g->stream>>>(a); g->launches++; }
    }
    if (we > wb) {
        int64_t need = ((int64_t)(we - wb) + NB_LWARPS - 1) / NB_LWARPS;
        unsigned grid = (unsigned)std::min<int64_t>(need, SMEM ? NB_LEARN_MAX_BLOCKS : (1 << 30));
        k_learn_warp<WIDE, SMEM><<<grid, NB_LEARN_THREADS, smem, g->stream>>>(a, wb, we);
        g->launches++;
        if (!SMEM) { k_apply_global<<<apply_grid, 256, 0, g->stream>>>(a); g->launches++; }
    }
    return NB_OK;
}

// per-graph cache of the visit bounds (depends on learn_non_evidence)
static int learn_prepare(nb_graph *g, LearnArgs &a, std::vector<int64_t> &vmax, int learn_non_evidence)
{
    const bool smem = g->W <= NB_LEARN_SMEM_W;
    NB_TRY(ensure_learn_buffers(g, smem));
    a = learn_args(g);
    a.learn_non_evidence = learn_non_evidence;
    if (g->learn_vmax_flag != learn_non_evidence || (int)g->learn_vmax.size() != g->n_colors) {
        NB_TRY(color_visit_bounds(g, a, g->learn_vmax));
        g->learn_vmax_flag = learn_non_evidence;
    }
    vmax = g->learn_vmax;
    return NB_OK;
}

// Rows of colour c whose original id falls into block `b` of `nb` (blocks = runs of id windows).
// The learning epoch walks the blocks in increasing id order and, inside a block, the colours:
// weights and chains advance together through the graph like in the reference's ascending-id
// scan (learning.py:20-31), instead of one whole colour (all of a variable type) at a time.
static int learn_block_of_color(nb_graph *g, const LearnArgs &a, int c, int b, int nb)
{
    const bool smem = g->W <= NB_LEARN_SMEM_W;
    const NbColorRange &cr = g->colors[(size_t)c];
    const int64_t w_lo = g->n_win * (int64_t)b / nb, w_hi = g->n_win * (int64_t)(b + 1) / nb;
    const size_t row = (size_t)g->n_win + 1;
    auto range = [&](int cls, int &beg, int &end) {
        const int32_t *ws = g->win_start.data() + (size_t)(cls * (g->n_colors + 1) + c) * row;
        beg = ws[w_lo];
        end = ws[w_hi];
    };
    int pb, pe, fb, fe, cb, ce, tb, te, wb, we;
    range(NB_CLASS_PAIR, pb, pe);
    range(NB_CLASS_FAST, fb, fe);
    range(NB_CLASS_CAT, cb, ce);      // categorical rows: learned by the generic thread kernel
    range(NB_CLASS_GEN, tb, te);
    range(NB_CLASS_WARP, wb, we);
    wb -= (int)g->n_trows;          // warp rows are addressed by their index
    we -= (int)g->n_trows;
    if (pe <= pb && fe <= fb && ce <= cb && te <= tb && we <= wb) return NB_OK;
    const int64_t rows = (cr.p_end - cr.p_beg) + (cr.f_end - cr.f_beg) + (cr.c_end - cr.c_beg) + (cr.t_end - cr.t_beg) +
                         (cr.w_end - cr.w_beg);
    // truth-table rows of this colour average >= 16 incidences: spread each row over a warp
    const bool long_rows = ((cr.p_end - cr.p_beg) + (cr.f_end - cr.f_beg)) > 0 && cr.edges >= 16 * rows;
    if (g->wide) {
        if (smem) return launch_learn_range<true, true>(g, a, pb, pe, fb, fe, cb, ce, tb, te, wb, we, long_rows);
        return launch_learn_range<true, false>(g, a, pb, pe, fb, fe, cb, ce, tb, te, wb, we, long_rows);
    }
    if (smem) return launch_learn_range<false, true>(g, a, pb, pe, fb, fe, cb, ce, tb, te, wb, we, long_rows);
    return launch_learn_range<false, false>(g, a, pb, pe, fb, fe, cb, ce, tb, te, wb, we, long_rows);
}

// Mini-batch size: at most this many visits of any one weight between two applications.  The
// reference applies every visit immediately (learning.py:110-125); a batch of n visits with a
// stale weight moves it by about step * n * E[g], which must stay small against the scale on
// which the gradient itself changes or strongly driven transients overshoot (and L1 then traps
// the weights at 0).  step * n <= 0.25 keeps the batched trajectory on the per-visit one
// (measured against the CPU reference port on the labelling-function model, tools/lf_check.py).
static int64_t default_batch_visits(double step, int64_t batch_visits)
{
    return batch_visits > 0 ? batch_visits : (int64_t)std::max(1.0, std::floor(0.25 / std::max(std::fabs(step), 1e-12)));
}

// number of id blocks per epoch: enough that no weight collects more than the batch bound in a block
static int block_count(const nb_graph *g, const std::vector<int64_t> &vmax, int64_t bv)
{
    int64_t total = 0;
    for (int64_t v : vmax) total += v;       // a weight can be visited from every colour inside a block
    int64_t nb = std::max<int64_t>(1, (total + bv - 1) / bv);
    return (int)std::min<int64_t>(nb, std::max<int64_t>(1, g->n_win));
}

int nb_learn_block_count(nb_graph *g, double step, int learn_non_evidence, int64_t batch_visits, int *n_blocks)
{
    LearnArgs a;
    std::vector<int64_t> vmax;
    NB_TRY(learn_prepare(g, a, vmax, learn_non_evidence));
    *n_blocks = block_count(g, vmax, default_batch_visits(step, batch_visits));
    return NB_OK;
}

int nb_learn_color(nb_graph *g, int color, int block, int n_blocks, double step, int regularization, double reg_param,
                   double truncation, int learn_non_evidence, uint64_t seed, uint64_t epoch)
{
    if (color < 0 || color >= g->n_colors) return NB_OK;   // a colour this rank does not own
    if (n_blocks < 1 || block < 0 || block >= n_blocks) NB_FAIL(NB_ERR_INVALID, "block %d of %d", block, n_blocks);
    LearnArgs a;
    std::vector<int64_t> vmax;
    NB_TRY(learn_prepare(g, a, vmax, learn_non_evidence));
    a.seed = seed; a.reg_param = reg_param; a.truncation = truncation; a.regularization = regularization;
    a.step = step; a.epoch = epoch;
    NB_TRY(learn_block_of_color(g, a, color, block, n_blocks));
    g->weights_version++;
    NB_CUDA(cudaGetLastError());
    return NB_OK;
}

int nb_run_learn(nb_graph *g, int64_t n_epochs, double *stepsize, double decay, int regularization, double reg_param,
                 double truncation, int learn_non_evidence, uint64_t seed, int64_t batch_visits)
{
    if (n_epochs <= 0) return NB_OK;
    LearnArgs a;
    std::vector<int64_t> vmax;
    NB_TRY(learn_prepare(g, a, vmax, learn_non_evidence));
    a.seed = seed; a.reg_param = reg_param; a.truncation = truncation; a.regularization = regularization;
    double step = *stepsize;
    for (int64_t ep = 0; ep < n_epochs; ep++) {
        a.step = step;
        a.epoch = g->epoch_counter++;
        const int nb = block_count(g, vmax, default_batch_visits(step, batch_visits));
        for (int b = 0; b < nb; b++)
            for (int c = 0; c < g->n_colors; c++) NB_TRY(learn_block_of_color(g, a, c, b, nb));
        g->weights_version++;
        NB_CUDA(cudaGetLastError());
        step *= decay;   // factorgraph.py:206
    }
    *stepsize = step;
    return NB_OK;
}
