/*
 * pa_model_mega.cu -- the whole decode step of a handful of sequences (<= 8, one new token each) as
 * ONE persistent cooperative kernel (the batch-1 case of BASELINE configs[0]).
 *
 * With one to eight rows every op of gpt2_forward (paged_infer.c:646-728) is a read of its weights:
 * 0.5 GB per step at GPT-2 124M = 77 us at HBM speed, while a chain of 88 dependent launches costs
 * ~9 us each whatever they do.  So the step is one grid of one 512-thread CTA per SM that walks the
 * ops in order and meets at a grid barrier wherever an op needs what other CTAs produced:
 *
 *   embedding | per layer { [ln1 -> smem] QKV rows + KV append | attention |
 *               [atty -> smem] attproj + residual | [ln2 -> smem] fc + GELU | [fch -> smem] fcproj + residual } |
 *   [lnf -> smem] LM head | sampler
 *
 * Projections: weight-streaming GEMV -- a warp owns 1/2/4 output features (or, for rows of 4C
 * floats, a pair of warps one feature), its lanes stream those weight rows with 16-byte loads that
 * bypass L1, the M input rows sit in shared memory (every CTA normalises / stages them for itself,
 * so layernorm costs no barrier), fp32 FMA, warp-shuffle reduction, bias / GELU / residual /
 * page-slot scatter in the epilogue -- the arithmetic of pa_gemv_kernel (pa_qkv.cu).
 * Attention: up to 512 tokens the warps of ONE CTA share a (sequence, head) and merge their chunks
 * through shared memory; beyond that the chunks spread over the grid, partial (o, m, l) go to a
 * workspace and are merged after one more barrier.  hs/4 lanes per token, 16-byte loads through the
 * block table, online softmax from the reference's -10000 start (paged_infer.c:187).
 * Barrier: one release-add + relaxed polls + one acquire fence on a counter in L2 that only ever
 * grows (also across launches); between ARRIVING and WAITING a CTA requests whatever the next phase
 * needs that no other CTA produces (weights, bias, layernorm parameters, old K/V).  Activations
 * written by one CTA and read by another after a barrier are read through L2 (ld.global.cg).
 * Roofline: HBM (weights read once); measured numbers and the history of the kernel in DESIGN 4.6 and
 * profiles/r01_model_step.md.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "pa_internal.h"
#include "pa_model_dev.cuh"
#include "pa_ptx.cuh"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {

constexpr int kWarps = PA_MEGA_WARPS_PER_SM;       // 16: at 8 (two per scheduler) every dependent instruction was paid at full latency
constexpr int kThreads = kWarps * 32;
constexpr int kMaxM = PA_MEGA_MAX_SEQS;
constexpr int kLoads = 12;                 // 16-byte weight loads in flight per lane (registers: 512 threads leave 128 each)
constexpr float kMaxInit = -10000.0f;      // paged_infer.c:187

// All CTAs of the (cooperative, hence co-resident) grid meet: the barrier orders the CTA's stores
// before thread 0's release-add; the acquire-poll plus the second barrier orders everybody's later
// loads after the other CTAs' stores.  The counter only grows (over launches too, so nothing resets
// it): barrier i of this launch completes at base + (i+1)*gridDim, compared modulo 2^32.
// The barrier is split: between arriving and waiting a CTA requests whatever the next phase needs
// that does not depend on other CTAs (weights, biases, layernorm parameters, old K/V), so those
// loads are issued and travel while the grid meets.
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned* bar, unsigned base, unsigned& passed) {
    if (threadIdx.x == 0) {
        const unsigned target = base + (passed + 1) * gridDim.x;
        // relaxed polls and ONE acquire fence at the end: an acquire load invalidates the SM's whole L1
        // every time it executes, under the feet of the loads the CTA has in flight
        unsigned seen;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while ((int)(seen - target) < 0);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    ++passed;
}

// The kernel body is deliberately COMPACT: every phase runs its code once per layer on a few warps
// per scheduler, so instruction count is time -- a first version with one inlined GEMV per call site
// and feature count (47 k instructions, 755 KB) spent ~10 us per phase.  Hence ONE loop over the
// projection phases with one copy of each routine, and the batch bound as a template parameter.

// ---- projections: out(m, n) for the M rows in shared memory and every feature n of w (N, K) -------
// A warp pass covers `feat` (1, 2 or 4: the fewest that cover N in one pass of the grid's warps)
// output features; its lanes interleave the 16-byte chunks of those weight rows and keep kLoads
// loads in flight.  Load u * feat + f of a batch is chunk u of feature f; a feature's chunks
// accumulate into 4 / feat slots (independent FMA chains) that are folded at the end.
__device__ __forceinline__ int pick_feat_shift(int N, int K) {
    const int nw = gridDim.x * kWarps;
    int sh = N <= nw ? 0 : (N <= 2 * nw ? 1 : 2);
    // ... but a warp's first batch should cover its whole rows when it can (the rest is not prefetched)
    while (sh > 0 && (kLoads >> sh) * 128 < K) --sh;
    return sh;
}
// K split over a pair of neighbouring warps: when a single feature per warp would need more than one batch
// of loads for its row (fcproj: K = 4C) and half the warps would idle anyway, warps 2i and 2i+1 of a CTA
// take the two halves of feature i's row and add their sums through shared memory.
__device__ __forceinline__ int pick_ksplit(int N, int K, int sh) {
    return sh == 0 && (K >> 2) > kLoads * 32 && 2 * N <= (int)gridDim.x * kWarps && (K & 7) == 0;
}
// The weight loads of a warp's FIRST batch are issued between arriving at the grid barrier in front of
// the phase and waiting on it: weights never depend on another CTA, so they stream in from HBM while
// the grid meets.  Read-only, read-once: straight from L2, no L1 line to allocate.
__device__ __forceinline__ float4 ld_weight(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// One specialisation per feature count (SH = log2): with everything about the batch's shape known at
// compile time a load is an add and an LDG, and an FMA group an LDS and four FFMAs -- with run-time
// shapes the same loop was ~900 instructions of address arithmetic and branches per pass, and at
// two warps per scheduler those are paid at full latency.  Load u*FEAT + f = chunk u of feature f.
template <int SH>
__device__ __forceinline__ void gemv_issue_t(float4 (&wv)[kLoads], const float* __restrict__ w, int N, int K, int k4len, int n0, int cb) {
    constexpr int FEAT = 1 << SH, UN = kLoads >> SH;
#pragma unroll
    for (int f = 0; f < FEAT; ++f) {
        const float4* row = reinterpret_cast<const float4*>(w + (size_t)min(n0 + f, N - 1) * K) + cb;
#pragma unroll
        for (int u = 0; u < UN; ++u)
            wv[u * FEAT + f] = (cb + 32 * u < k4len && n0 < N) ? ld_weight(row + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// w: first float of the warp's K range in row 0; K: floats between rows; k4len: 16-byte chunks of the range
__device__ __forceinline__ void gemv_issue(float4 (&wv)[kLoads], const float* __restrict__ w, int N, int K, int k4len, int sh, int n0, int cb) {
    if (sh == 0) gemv_issue_t<0>(wv, w, N, K, k4len, n0, cb);
    else if (sh == 1) gemv_issue_t<1>(wv, w, N, K, k4len, n0, cb);
    else gemv_issue_t<2>(wv, w, N, K, k4len, n0, cb);
}
// acc[m][slot] += w . x for one batch; a feature's chunks spread over 4 / FEAT slots (independent FMA
// chains), folded by the caller.  Chunks past the row carry zero weights: their x index is clamped.
template <int SH, int MAXM>
__device__ __forceinline__ void gemv_fma_t(const float4 (&wv)[kLoads], const float4* xs4, int cb, int K4, int k4len, int M, float (&acc)[MAXM][4]) {
    constexpr int FEAT = 1 << SH, UN = kLoads >> SH;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const int c = min(cb + 32 * u, k4len - 1);
#pragma unroll
        for (int mm = 0; mm < MAXM; ++mm) {
            if (mm < M) {
                const float4 xv = xs4[mm * K4 + c];
#pragma unroll
                for (int f = 0; f < FEAT; ++f) {
                    constexpr int kSpread = 4 / FEAT;
                    const int slot = f + FEAT * (u % kSpread);
                    const float4 wq = wv[u * FEAT + f];
                    acc[mm][slot] = fmaf(wq.w, xv.w, fmaf(wq.z, xv.z, fmaf(wq.y, xv.y, fmaf(wq.x, xv.x, acc[mm][slot]))));
                }
            }
        }
    }
}

struct StepSmem { int kv_start[kMaxM], kv_end[kMaxM], slot[kMaxM]; };

// ---- attention, phase 1: a warp per (sequence, head, chunk); LPT = hs/4 lanes per token ---------
template <int LPT>
struct KvBatch {                           // one batch of a warp's unit: UN iterations x TPI tokens, this lane's 16 bytes of each
    static constexpr int TPI = 32 / LPT, UN = 4;
    float4 k[UN], v[UN];
};
// Requests K/V of the tokens [tb, tb + TPI*UN) of the unit.  only_new = false: every token except the
// step's new one (position last-1, still being written by the QKV phase) -- issued BEFORE the barrier;
// only_new = true: just that one, after the barrier.
template <int LPT>
__device__ __forceinline__ void attn_issue(KvBatch<LPT>& kb, const pa_mega_args& a, const float* pool_k, const float* pool_v,
                                           const int* tbl, int h, int tb, int t1, int last, bool only_new) {
    constexpr int TPI = KvBatch<LPT>::TPI, UN = KvBatch<LPT>::UN;
    const int lane = threadIdx.x & 31, sub = lane / LPT, li = lane % LPT;
#pragma unroll
    for (int i = 0; i < UN; ++i) {
        const int t = tb + i * TPI + sub;
        const bool is_new = t == last - 1;
        if (t < t1 && is_new == only_new) {
            const size_t off = ((size_t)((tbl[t >> a.bs_shift] << a.bs_shift) + (t & (a.bs - 1)))) * a.C + h * a.hs;
            kb.k[i] = __ldcg(reinterpret_cast<const float4*>(pool_k + off) + li);
            kb.v[i] = __ldcg(reinterpret_cast<const float4*>(pool_v + off) + li);
        } else if (!only_new) {
            kb.k[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            kb.v[i] = kb.k[i];
        }
    }
}
// unit u -> (sequence*NH + head, chunk, token range); false when the chunk lies beyond the sequence
// Two shapes of the split: GLOBAL -- chunks of a.chunk_tokens spread over the whole grid, partials to the
// workspace, merged after a grid barrier (long contexts); LOCAL (a.local_attn) -- the warps of ONE CTA
// share a (sequence, head), kWarps chunks of ceil(len / kWarps) tokens, merged through shared memory
// right away: no partial traffic, no merge phase, one grid barrier less per layer.
__device__ __forceinline__ bool attn_unit(const pa_mega_args& a, const StepSmem& st, int u, int& sh, int& c, int& t0, int& t1, int& last) {
    c = u % a.max_chunks; sh = u / a.max_chunks;
    if (sh >= a.M * a.NH) { t0 = t1 = last = 0; return false; }
    const int s = sh / a.NH;
    last = st.kv_end[s];
    const int chunk = a.local_attn ? (last - st.kv_start[s] + kWarps - 1) / kWarps : a.chunk_tokens;
    t0 = st.kv_start[s] + c * chunk;
    t1 = min(last, t0 + chunk);
    return t0 < last;
}
template <int LPT>
__device__ __forceinline__ void attn_partials(KvBatch<LPT>& kb, const pa_mega_args& a, const StepSmem& st, const float* pool_k, const float* pool_v,
                                              float (*ps)[132]) {
    constexpr int TPI = KvBatch<LPT>::TPI, UN = KvBatch<LPT>::UN;
    const int lane = threadIdx.x & 31, sub = lane / LPT, li = lane % LPT;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    const int hs = a.hs, C = a.C;
    const int n_units = a.M * a.NH * a.max_chunks;
    bool have = true;                                               // kb holds the first batch of unit gw (all but the new token)
    for (int u = gw; u < n_units; u += nw) {              // (LOCAL: the CTA's warps run this loop in lockstep, one (sequence, head) at a time)
        int sh, c, t0, t1, last;
        const bool valid = attn_unit(a, st, u, sh, c, t0, t1, last);                  // warp-uniform
        if (!valid) have = false;
        if (!valid && !a.local_attn) continue;
        const int h = sh % a.NH, s = sh / a.NH;
        const int* tbl = a.table + (size_t)s * a.tstride;
        const float4 q4 = valid ? __ldcg(reinterpret_cast<const float4*>(a.q + (size_t)s * C + h * hs) + li) : make_float4(0.f, 0.f, 0.f, 0.f);
        float m_run = kMaxInit, l_run = 0.0f;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tb = t0; tb < t1; tb += TPI * UN) {              // (no iteration for an empty chunk)
            if (!have) attn_issue<LPT>(kb, a, pool_k, pool_v, tbl, h, tb, t1, last, false);
            have = false;
            attn_issue<LPT>(kb, a, pool_k, pool_v, tbl, h, tb, t1, last, true);
            // the batch's scores, then ONE rescale of the running state for all of them
            float sc[UN];
            float m_new = m_run;
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                float dot = fmaf(q4.w, kb.k[i].w, fmaf(q4.z, kb.k[i].z, fmaf(q4.y, kb.k[i].y, q4.x * kb.k[i].x)));
#pragma unroll
                for (int d = LPT / 2; d >= 1; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
                sc[i] = (tb + i * TPI + sub < t1) ? dot * a.scale : -INFINITY;      // uniform over the token's LPT lanes
                m_new = fmaxf(m_new, sc[i]);
            }
            const float alpha = expf(m_run - m_new);
            l_run *= alpha; o4.x *= alpha; o4.y *= alpha; o4.z *= alpha; o4.w *= alpha;
            m_run = m_new;
#pragma unroll 1
            for (int i = 0; i < UN; ++i) {                          // (not unrolled: one expf body; sc/kb indexed through a switch-free select)
                float s_i = sc[0];
                float4 v_i = kb.v[0];
#pragma unroll
                for (int jx = 1; jx < UN; ++jx) if (jx == i) { s_i = sc[jx]; v_i = kb.v[jx]; }
                const float e = expf(s_i - m_new);                  // exp(-inf) = 0 for the padding
                l_run += e;
                o4.x = fmaf(e, v_i.x, o4.x); o4.y = fmaf(e, v_i.y, o4.y); o4.z = fmaf(e, v_i.z, o4.z); o4.w = fmaf(e, v_i.w, o4.w);
            }
        }
        if (TPI == 2) {                                             // the two token sub-groups of the warp -> one partial
            const float m_o = __shfl_xor_sync(0xffffffffu, m_run, LPT), l_o = __shfl_xor_sync(0xffffffffu, l_run, LPT);
            float4 o_o;
            o_o.x = __shfl_xor_sync(0xffffffffu, o4.x, LPT); o_o.y = __shfl_xor_sync(0xffffffffu, o4.y, LPT);
            o_o.z = __shfl_xor_sync(0xffffffffu, o4.z, LPT); o_o.w = __shfl_xor_sync(0xffffffffu, o4.w, LPT);
            const float m_t = fmaxf(m_run, m_o);
            const float wa = expf(m_run - m_t), wb = expf(m_o - m_t);
            l_run = fmaf(l_run, wa, l_o * wb);
            o4.x = fmaf(o4.x, wa, o_o.x * wb); o4.y = fmaf(o4.y, wa, o_o.y * wb);
            o4.z = fmaf(o4.z, wa, o_o.z * wb); o4.w = fmaf(o4.w, wa, o_o.w * wb);
            m_run = m_t;
        }
        if (!a.local_attn) {
            if (sub == 0) {
                float* pr = a.part + (size_t)(sh * a.max_chunks + c) * (hs + 4);
                reinterpret_cast<float4*>(pr)[li] = o4;
                if (li == 0) { pr[hs] = m_run; pr[hs + 1] = l_run; }
            }
        } else {
            // the CTA's kWarps partials (an empty chunk leaves the neutral state) -> shared memory, merged in chunk order
            if (sub == 0) {
                reinterpret_cast<float4*>(ps[c])[li] = o4;
                if (li == 0) { ps[c][hs] = m_run; ps[c][hs + 1] = l_run; }
            }
            __syncthreads();
            if ((int)threadIdx.x < hs && sh < a.M * a.NH) {
                const int dd = threadIdx.x;
                float m_tot = kMaxInit;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) m_tot = fmaxf(m_tot, ps[w][hs]);
                float l_tot = 0.0f, o = 0.0f;
#pragma unroll 4
                for (int w = 0; w < kWarps; ++w) {
                    const float wgt = expf(ps[w][hs] - m_tot);
                    l_tot = fmaf(ps[w][hs + 1], wgt, l_tot);
                    o = fmaf(ps[w][dd], wgt, o);
                }
                a.atty[(size_t)s * C + h * hs + dd] = o * ((l_tot == 0.0f) ? 0.0f : 1.0f / l_tot);      // :213
            }
            __syncthreads();
        }
    }
}

// ---- attention, phase 2: a warp per (sequence, head) merges its chunks' partials in chunk order ----
// lane i weighs partial i (one expf per lane), the weights travel by shuffle; every lane owns the
// head dims lane, lane + 32, ...
template <int LPT>
__device__ __forceinline__ void attn_merge(const pa_mega_args& a, const StepSmem& st) {
    constexpr int UN = 8, NJ = LPT / 8;            // partials per batch; head dims per lane (head_dim / 32)
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    const int hs = a.hs;
    for (int sh = gw; sh < a.M * a.NH; sh += nw) {
        const int s = sh / a.NH, h = sh % a.NH;
        const int len = st.kv_end[s] - st.kv_start[s];
        const int n_part = (len + a.chunk_tokens - 1) / a.chunk_tokens;
        const float* pr = a.part + (size_t)sh * a.max_chunks * (hs + 4);
        float m_tot = kMaxInit;
        if (n_part > UN) {                                          // several batches: the maxima first
            for (int i = lane; i < n_part; i += 32) m_tot = fmaxf(m_tot, __ldcg(pr + (size_t)i * (hs + 4) + hs));
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) m_tot = fmaxf(m_tot, __shfl_xor_sync(0xffffffffu, m_tot, d));
        }
        float o[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) o[j] = 0.0f;
        float l_tot = 0.0f;
        for (int ib = 0; ib < n_part; ib += UN) {
            float po[UN][NJ];
#pragma unroll
            for (int i = 0; i < UN; ++i) {                          // the batch's o values in flight at once
                const float* pp = pr + (size_t)min(ib + i, n_part - 1) * (hs + 4);
#pragma unroll
                for (int j = 0; j < NJ; ++j) po[i][j] = __ldcg(pp + lane + 32 * j);
            }
            const int mine = ib + (lane & (UN - 1));                // lanes 0..15 (and their mirrors) weigh partial ib + lane
            const float* pp = pr + (size_t)min(mine, n_part - 1) * (hs + 4);
            const float pm = mine < n_part ? __ldcg(pp + hs) : kMaxInit;
            const float pl = mine < n_part ? __ldcg(pp + hs + 1) : 0.0f;
            if (n_part <= UN) {                                     // one batch: its maxima arrive with everything else
                m_tot = pm;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) m_tot = fmaxf(m_tot, __shfl_xor_sync(0xffffffffu, m_tot, d));
                m_tot = fmaxf(m_tot, kMaxInit);
            }
            const float wgt = mine < n_part ? expf(pm - m_tot) : 0.0f;
            const float lw = pl * wgt;
#pragma unroll
            for (int i = 0; i < UN; ++i) {                          // chunk order
                const float w_i = __shfl_sync(0xffffffffu, wgt, i);
                l_tot += __shfl_sync(0xffffffffu, lw, i);
#pragma unroll
                for (int j = 0; j < NJ; ++j) o[j] = fmaf(po[i][j], w_i, o[j]);
            }
        }
        const float inv = (l_tot == 0.0f) ? 0.0f : 1.0f / l_tot;        // :213
        float* out = a.atty + (size_t)s * a.C + h * hs;
#pragma unroll
        for (int j = 0; j < NJ; ++j) out[lane + 32 * j] = o[j] * inv;
    }
}

enum { PH_QKV = 0, PH_ATTPROJ = 1, PH_FC = 2, PH_FCPROJ = 3, PH_LM = 4 };
struct PhaseDesc { const float *w, *bias; int N, K; };
__device__ __forceinline__ PhaseDesc phase_desc(const pa_mega_args& a, int ph) {
    const int C = a.C, l = ph >> 2;
    PhaseDesc d;
    if (ph >= 4 * a.L) { d.w = a.wte; d.bias = nullptr; d.N = a.V; d.K = C; return d; }
    switch (ph & 3) {
        case PH_QKV: d.w = a.qkvw + (size_t)l * 3 * C * C; d.bias = a.qkvb + (size_t)l * 3 * C; d.N = 3 * C; d.K = C; break;
        case PH_ATTPROJ: d.w = a.attprojw + (size_t)l * C * C; d.bias = a.attprojb + (size_t)l * C; d.N = C; d.K = C; break;
        case PH_FC: d.w = a.fcw + (size_t)l * 4 * C * C; d.bias = a.fcb + (size_t)l * 4 * C; d.N = 4 * C; d.K = C; break;
        default: d.w = a.fcprojw + (size_t)l * 4 * C * C; d.bias = a.fcprojb + (size_t)l * C; d.N = C; d.K = 4 * C; break;
    }
    return d;
}

// LPT: lanes per token (head_dim / 4); MAXM: bound of the batch (unrolling of the row loops);
// LNREGS: registers per lane of a layernorm row (32 * LNREGS >= C)
template <int LPT, int MAXM, int LNREGS>
__global__ void __launch_bounds__(kThreads, 1)
pa_decode_step_mega_kernel(const pa_mega_args a) {
    extern __shared__ __align__(16) float xs[];          // [M][4C] staged / normalised input rows | layernorm weight [C] | bias [C]
    __shared__ PaSampleSmem<kThreads> samp;
    __shared__ StepSmem st;
    __shared__ float ln_red[2 * kWarps];
    __shared__ float ks_buf[kWarps / 2][kMaxM];              // K split: a warp pair's second partial sums
    __shared__ __align__(16) float attn_ps[kWarps][132];     // LOCAL attention: the CTA's partials (o[hs], m, l)
    const int M = a.M, C = a.C;
    float* ln_ws = xs + (size_t)M * 4 * C;
    float* ln_bs = ln_ws + C;
    // the next layernorm's weight and bias -> shared memory with fire-and-forget 16-byte copies, issued before a
    // barrier and waited for after it (they are parameters: nothing to wait for but HBM)
    auto ln_params_issue = [&](const float* lw, const float* lb) {
        for (int i = threadIdx.x; i < (C >> 2); i += kThreads) {
            cp_async16(smem_u32(ln_ws + 4 * i), lw + 4 * i, 16);
            cp_async16(smem_u32(ln_bs + 4 * i), lb + 4 * i, 16);
        }
        cp_async_commit();
    };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
    unsigned passed = 0;
    int n_stamp = 0;
    auto stamp = [&]() {            // PA_MEGA_DEBUG: CTA 0 records when it reaches each barrier and when it leaves it
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.dbg[n_stamp++] = t;
        }
    };
#define GRID_ARRIVE() do { stamp(); grid_arrive(a.bar); } while (0)
#define GRID_WAIT() do { grid_wait(a.bar, a.bar_base, passed); stamp(); } while (0)
    float4 wv[kLoads];               // a warp's first batch of weight loads for the NEXT projection, in flight across barriers
    // ... and the bias of the output this lane will finish in that first pass (lane (f, m) -> feature n0 + f)
    auto bias_issue = [&](const PhaseDesc& pd, int shift, int ksp) {
        const int n = (ksp ? gw >> 1 : gw << shift) + (lane >> 3);
        return (pd.bias && (lane >> 3) < (1 << shift) && n < pd.N && !(ksp && (gw & 1))) ? __ldg(pd.bias + n) : 0.0f;
    };
    // the warp's first batch of the phase: feature(s) and K range as the projection loop below will take them
    auto first_issue = [&](const PhaseDesc& pd, int shift, int ksp) {
        const int k4 = pd.K >> 2, k4len = ksp ? k4 >> 1 : k4;
        gemv_issue(wv, pd.w + (ksp ? (size_t)(gw & 1) * (k4len << 2) : 0), pd.N, pd.K, k4len, shift, ksp ? gw >> 1 : gw << shift, lane);
    };
    KvBatch<LPT> kb;                 // likewise the first batch of K/V for the attention phase

    if (threadIdx.x < M) {
        st.kv_start[threadIdx.x] = a.kv_start[threadIdx.x];
        st.kv_end[threadIdx.x] = a.kv_end[threadIdx.x];
        st.slot[threadIdx.x] = a.slots[threadIdx.x];
    }
    // encoder_forward (:24-46): CTA m writes row m of the residual stream
    if ((int)blockIdx.x < M) {
        const float* e = a.wte + (size_t)a.tokens[blockIdx.x] * C;
        const float* ps = a.wpe + (size_t)a.positions[blockIdx.x] * C;
        for (int i = threadIdx.x; i < C; i += kThreads) a.x[(size_t)blockIdx.x * C + i] = e[i] + ps[i];
    }
    GRID_ARRIVE();
    PhaseDesc d = phase_desc(a, 0);
    int sh = pick_feat_shift(d.N, d.K), ks = pick_ksplit(d.N, d.K, sh);
    first_issue(d, sh, ks);
    float bias_first = bias_issue(d, sh, ks);
    ln_params_issue(a.ln1w, a.ln1b);
    GRID_WAIT();

    const int n_phases = 4 * a.L + 1;
    for (int ph = 0; ph < n_phases; ++ph) {
        const int l = ph >> 2, kind = ph == n_phases - 1 ? PH_LM : (ph & 3);
        float* pool_k = a.pool_k + (size_t)l * a.layer_stride;
        float* pool_v = a.pool_v + (size_t)l * a.layer_stride;
        const int N = d.N, K = d.K, K4 = K >> 2, feat = 1 << sh;

        // ---- the phase's input rows -> shared memory: layernorm of the residual stream (ln1 :703, ln2 :718,
        // lnf :724), or the previous op's output as it is (atty, fch)
        if (kind == PH_ATTPROJ || kind == PH_FCPROJ) {
            const float4* src = reinterpret_cast<const float4*>(kind == PH_ATTPROJ ? a.atty : a.fch);
            for (int i = threadIdx.x; i < M * K4; i += kThreads) reinterpret_cast<float4*>(xs)[i] = __ldcg(src + i);
        } else {
            // layernorm_forward (:49-89), 8 / MAXM warps per row (a single warp would leave every latency exposed)
            constexpr int W = kWarps / MAXM, PER = (LNREGS + W - 1) / W;
            const int row = warp / W, wr = warp % W;
            const bool act = row < M;
            const float* xr = a.x + (size_t)row * C;
            float v[PER];
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int c = wr * 32 + lane + 32 * W * i;
                v[i] = (act && c < C) ? __ldcg(xr + c) : 0.0f;
                sum += v[i];
            }
#pragma unroll
            for (int dd = 16; dd >= 1; dd >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dd);
            if (lane == 0) ln_red[warp] = sum;
            cp_async_wait<0>();
            __syncthreads();                                  // the partial sums, and everybody's copies of the layernorm parameters
            float tot = 0.0f;
#pragma unroll
            for (int k = 0; k < W; ++k) tot += ln_red[row * W + k];
            const float mean = tot / C;
            float var = 0.0f;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int c = wr * 32 + lane + 32 * W * i;
                const float dlt = v[i] - mean;
                if (c < C) var += dlt * dlt;
            }
#pragma unroll
            for (int dd = 16; dd >= 1; dd >>= 1) var += __shfl_xor_sync(0xffffffffu, var, dd);
            if (lane == 0) ln_red[kWarps + warp] = var;
            __syncthreads();
            float totv = 0.0f;
#pragma unroll
            for (int k = 0; k < W; ++k) totv += ln_red[kWarps + row * W + k];
            const float rstd = 1.0f / sqrtf(totv / C + 1e-5f);           // eps, :56
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int c = wr * 32 + lane + 32 * W * i;
                if (act && c < C) xs[(size_t)row * C + c] = (rstd * (v[i] - mean)) * ln_ws[c] + ln_bs[c];
            }
        }
        __syncthreads();

        // ---- the projection (matmul_forward :92-114 on M rows): weight-streaming GEMV ------------------
        {
            const float* res = (kind == PH_ATTPROJ || kind == PH_FCPROJ) ? a.x : nullptr;     // residual_forward :253-257
            // the warp's share: feature(s) n0.. and a K range (all of K, or one half of it with a K split)
            const int half = ks ? (gw & 1) : 0, k4len = ks ? K4 >> 1 : K4;
            const float* wk = d.w + (size_t)half * (k4len << 2);
            const float4* xs4 = reinterpret_cast<const float4*>(xs) + half * k4len;
            const int n_stride = ks ? nw >> 1 : nw << sh;
            bool have = true;                                // wv holds the first batch of the first pass
            for (int n0 = ks ? gw >> 1 : gw << sh; n0 < N; n0 += n_stride) {
                // lane (f, m) finishes output (m, n0 + f): its bias and residual are requested now, used after the reduction
                const int f = lane >> 3, m = lane & 7;
                const bool mine = f < feat && m < M && n0 + f < N && half == 0;
                const float bv = have ? bias_first : ((mine && d.bias) ? __ldg(d.bias + n0 + f) : 0.0f);
                const float rv = (mine && res) ? __ldcg(res + (size_t)m * C + n0 + f) : 0.0f;
                // the rows of the warp's passes after this one -> L2 while this pass computes (two passes ahead;
                // the first pass also requests the one right after it)
                for (int ahead = have ? 1 : 2; ahead <= 2; ++ahead) {
                    const int np = n0 + ahead * n_stride;
                    if (np < N) {
                        const int lines = k4len >> 3;        // 128-byte lines of the warp's K range per row
                        for (int i = lane; i < feat * lines; i += 32) {
                            const int pf = i / lines, pl = i - pf * lines;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(wk + (size_t)min(np + pf, N - 1) * K + pl * 32));
                        }
                    }
                }
                float acc[MAXM][4];
#pragma unroll
                for (int mm = 0; mm < MAXM; ++mm)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mm][q] = 0.0f;
                for (int cb = lane; cb < k4len; cb += 32 * (kLoads >> sh)) {
                    if (!have) gemv_issue(wv, wk, N, K, k4len, sh, n0, cb);
                    have = false;
                    if (sh == 0) gemv_fma_t<0, MAXM>(wv, xs4, cb, K4, k4len, M, acc);
                    else if (sh == 1) gemv_fma_t<1, MAXM>(wv, xs4, cb, K4, k4len, M, acc);
                    else gemv_fma_t<2, MAXM>(wv, xs4, cb, K4, k4len, M, acc);
                }
                float v = 0.0f;
#pragma unroll
                for (int mm = 0; mm < MAXM; ++mm) {
                    if (mm < M) {
                        // fold the slots of a feature, then the lanes
                        if (feat <= 2) { acc[mm][0] += acc[mm][2]; acc[mm][1] += acc[mm][3]; }
                        if (feat == 1) acc[mm][0] += acc[mm][1];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q < feat) {
#pragma unroll
                                for (int dd = 16; dd >= 1; dd >>= 1) acc[mm][q] += __shfl_xor_sync(0xffffffffu, acc[mm][q], dd);
                                if (q == f && mm == m) v = acc[mm][q];
                            }
                        }
                    }
                }
                if (ks) {                                    // the pair's second half -> first half (lane m holds row m's sum after the reduction)
                    if (half && lane < M) {
#pragma unroll
                        for (int mm = 0; mm < MAXM; ++mm) if (mm == lane) ks_buf[warp >> 1][lane] = acc[mm][0];
                    }
                    named_bar_sync(1 + (warp >> 1), 64);
                    if (!half && mine) v += ks_buf[warp >> 1][m];
                    named_bar_sync(1 + (warp >> 1), 64);     // (the buffer is free again for the pair's next pass)
                }
                if (mine) {
                    const int n = n0 + f;
                    v += bv;
                    if (kind == PH_QKV) {                    // Q to the dense buffer, K and V straight to the token's page slot (:706-710)
                        if (n < C) a.q[(size_t)m * C + n] = v;
                        else {
                            const size_t slot_off = (size_t)st.slot[m] * C;
                            if (n < 2 * C) pool_k[slot_off + (n - C)] = v;
                            else pool_v[slot_off + (n - 2 * C)] = v;
                        }
                    } else if (kind == PH_FC) a.fch[(size_t)m * 4 * C + n] = pa_gelu(v);      // :719-720
                    else if (kind == PH_LM) a.logits[(size_t)m * a.Vp + n] = v;                // :726
                    else a.x[(size_t)m * C + n] = v + rv;                                      // :716-717, :721-722
                }
            }
        }

        // ---- paged attention between the QKV projection and attproj (:713-715) -------------------------
        if (kind == PH_QKV) {
            GRID_ARRIVE();
            int ush, uc, t0, t1, last;
            if (attn_unit(a, st, gw, ush, uc, t0, t1, last))
                attn_issue<LPT>(kb, a, pool_k, pool_v, a.table + (size_t)(ush / a.NH) * a.tstride, ush % a.NH, t0, t1, last, false);
            GRID_WAIT();
            attn_partials<LPT>(kb, a, st, pool_k, pool_v, attn_ps);
            if (!a.local_attn) {
                GRID_ARRIVE();
                GRID_WAIT();
                attn_merge<LPT>(a, st);
            }
        }
        GRID_ARRIVE();
        if (ph + 1 < n_phases) {
            d = phase_desc(a, ph + 1);
            sh = pick_feat_shift(d.N, d.K);
            ks = pick_ksplit(d.N, d.K, sh);
            first_issue(d, sh, ks);
            bias_first = bias_issue(d, sh, ks);
            // (this phase read ln_ws/ln_bs before its __syncthreads at the latest; nobody reads them again before the barrier)
            if (kind == PH_ATTPROJ) ln_params_issue(a.ln2w + (size_t)l * C, a.ln2b + (size_t)l * C);
            else if (kind == PH_FCPROJ) {
                if (ph + 2 < n_phases) ln_params_issue(a.ln1w + (size_t)(l + 1) * C, a.ln1b + (size_t)(l + 1) * C);
                else ln_params_issue(a.lnfw, a.lnfb);
            }
        }
        GRID_WAIT();
    }
    if (a.dbg) {                     // PA_MEGA_DEBUG: 16 empty barriers back to back = the bare cost of one
        for (int i = 0; i < 16; ++i) { GRID_ARRIVE(); GRID_WAIT(); }
    }
    // softmax_forward + sample_mult on each row of logits
    if ((int)blockIdx.x < M)
        pa_sample_row<kThreads>(a.logits + (size_t)blockIdx.x * a.Vp, a.V, a.use_coins ? a.coins[blockIdx.x] : -1.0f, a.next + blockIdx.x, samp);
}

typedef void (*MegaKernel)(const pa_mega_args);
template <int LPT, int LNREGS>
MegaKernel pick_kernel_m(int M) {
    if (M <= 1) return pa_decode_step_mega_kernel<LPT, 1, LNREGS>;
    if (M <= 2) return pa_decode_step_mega_kernel<LPT, 2, LNREGS>;
    if (M <= 4) return pa_decode_step_mega_kernel<LPT, 4, LNREGS>;
    return pa_decode_step_mega_kernel<LPT, 8, LNREGS>;
}
MegaKernel pick_kernel(int M, int C, int hs) {
    if (hs == 64) return C <= 1024 ? pick_kernel_m<16, 32>(M) : pick_kernel_m<16, kLnMaxPerLane>(M);
    return C <= 1024 ? pick_kernel_m<32, 32>(M) : pick_kernel_m<32, kLnMaxPerLane>(M);
}

}  // namespace

extern "C" size_t pa_cu_model_mega_smem(int M, int C, int hs, int block_size) {
    if (M < 1 || M > kMaxM || (C & 3) || C > 32 * kLnMaxPerLane || (hs != 64 && hs != 128)) return 0;
    if (block_size < 1 || (block_size & (block_size - 1))) return 0;          // pages are addressed with shifts
    const size_t bytes = ((size_t)M * 4 * C + 2 * (size_t)C) * sizeof(float);        // input rows + layernorm weight and bias
    return bytes <= 200 * 1024 ? bytes : 0;
}

extern "C" int pa_cu_model_mega_step(const pa_mega_args* a, void* stream) {
    const size_t smem = pa_cu_model_mega_smem(a->M, a->C, a->hs, a->bs);
    if (!smem) return PA_ERR_UNSUPPORTED;
    MegaKernel fn = pick_kernel(a->M, a->C, a->hs);
    // (each instantiation needs its own opt-in for more than 48 KB of dynamic shared memory)
    // (and the opt-in is per device: one process may drive several GPUs)
    int cur_dev = 0;
    CU_CHECK(cudaGetDevice(&cur_dev));
    static MegaKernel attr_fn[64];
    static size_t attr_smem[64];
    static int attr_dev[64];
    static int attr_n = 0;
    int ai = 0;
    while (ai < attr_n && !(attr_fn[ai] == fn && attr_dev[ai] == cur_dev)) ++ai;
    if (ai == attr_n || attr_smem[ai] < smem) {
        CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (ai == attr_n && attr_n < 64) ++attr_n;
        if (ai < 64) { attr_fn[ai] = fn; attr_smem[ai] = smem; attr_dev[ai] = cur_dev; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    pa_mega_args args = *a;
    void* kargs[] = {&args};
    // cooperative: the launch fails instead of deadlocking the grid barrier if the grid could not be co-resident
    static unsigned long long* d_dbg = nullptr;
    const bool dbg = getenv("PA_MEGA_DEBUG") != nullptr;
    if (dbg && !d_dbg) CU_CHECK(cudaMalloc((void**)&d_dbg, 4096 * sizeof(unsigned long long)));
    args.dbg = dbg ? d_dbg : nullptr;
    CU_CHECK(cudaLaunchCooperativeKernel((const void*)fn, dim3(a->sm_count), dim3(kThreads), kargs, smem, s));
    if (dbg) {        // per phase (averaged over the layers): ns of work before the barrier, ns inside the barrier
        static unsigned long long hst[4096];
        CU_CHECK(cudaStreamSynchronize(s));
        CU_CHECK(cudaMemcpy(hst, d_dbg, sizeof(hst), cudaMemcpyDeviceToHost));
        const int per_layer = a->local_attn ? 5 : 6, n = 1 + per_layer * a->L + 1;       // barriers
        const char* names6[6] = {"qkv", "attn", "merge", "attproj", "fc", "fcproj"};
        const char* names5[5] = {"qkv", "attn+merge", "attproj", "fc", "fcproj"};
        const char* const* names = a->local_attn ? names5 : names6;
        double work[6] = {0}, wait[6] = {0};
        for (int l = 0; l < a->L; ++l)
            for (int p = 0; p < per_layer; ++p) {
                const int b = 1 + l * per_layer + p;                   // barrier index: stamps 2b (arrive), 2b+1 (leave)
                work[p] += (double)(hst[2 * b] - hst[2 * b - 1]);
                wait[p] += (double)(hst[2 * b + 1] - hst[2 * b]);
            }
        fprintf(stderr, "mega dbg: embed barrier %lld ns;", (long long)(hst[1] - hst[0]));
        for (int p = 0; p < per_layer; ++p) fprintf(stderr, " %s %.0f+%.0f", names[p], work[p] / a->L, wait[p] / a->L);
        fprintf(stderr, "; bare barrier %.0f ns", (double)(hst[2 * (n + 16) - 1] - hst[2 * n - 1]) / 16.0);
        fprintf(stderr, "; lm head %lld+%lld; total %lld ns\n", (long long)(hst[2 * (n - 1)] - hst[2 * (n - 1) - 1]),
                (long long)(hst[2 * (n - 1) + 1] - hst[2 * (n - 1)]), (long long)(hst[2 * (n - 1) + 1] - hst[0]));
    }
    return PA_OK;
}
