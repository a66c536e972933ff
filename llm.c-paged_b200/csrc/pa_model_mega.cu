/*
 * pa_model_mega.cu -- the whole decode step of a handful of sequences (<= 8, one new token each) as
 * ONE persistent cooperative kernel (the batch-1 case of BASELINE configs[0]).
 *
 * With one to eight rows every op of gpt2_forward (paged_infer.c:646-728) is a read of its weights:
 * 0.5 GB per step at GPT-2 124M = 77 us at HBM speed, while a chain of 88 dependent launches costs
 * ~9 us each whatever they do.  So the step is one grid of one CTA per SM that walks the ops in
 * order and meets at a grid barrier (one release-add + acquire-poll on a counter in L2) wherever an
 * op needs what other CTAs produced:
 *
 *   embedding | per layer { [ln1 -> smem] QKV rows + KV append | attention partials | merge |
 *               [atty -> smem] attproj + residual | [ln2 -> smem] fc + GELU | [fch -> smem] fcproj + residual } |
 *   [lnf -> smem] LM head | sampler
 *
 * Projections: weight-streaming GEMV -- a warp owns 1/2/4 output features, its lanes stream those
 * weight rows with 16-byte loads (all loads of a row in flight before the first FMA), the M input
 * rows sit in shared memory (every CTA normalises / stages them for itself, so layernorm costs no
 * barrier), fp32 FMA, warp-shuffle reduction, bias / GELU / residual / page-slot scatter in the
 * epilogue -- the arithmetic of pa_gemv_kernel (pa_qkv.cu).  Attention: a warp per (sequence, head,
 * chunk of tokens) walks the block table with 16-byte loads, hs/4 lanes per token, online softmax
 * from the reference's -10000 start (paged_infer.c:187), partial (o, m, l) to a workspace; a warp
 * per (sequence, head) merges the chunks in order.  Activations written by one CTA and read by
 * another after a barrier are read through L2 (ld.global.cg).  Roofline: HBM (weights read once).
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "pa_internal.h"
#include "pa_model_dev.cuh"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxM = PA_MEGA_MAX_SEQS;
constexpr int kLoads = 24;                 // 16-byte weight loads in flight per lane
constexpr float kMaxInit = -10000.0f;      // paged_infer.c:187

// All CTAs of the (cooperative, hence co-resident) grid meet: the barrier orders the CTA's stores
// before thread 0's release-add; the acquire-poll plus the second barrier orders everybody's later
// loads after the other CTAs' stores.  The counter only grows: barrier i completes at (i+1)*gridDim.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& passed) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = (passed + 1) * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
    }
    __syncthreads();
    ++passed;
}

// M rows of K floats, global (written by other CTAs before the last barrier) -> shared memory
__device__ __forceinline__ void stage_rows(float* xs, const float* src, int M, int K) {
    const int K4 = K >> 2;
    for (int i = threadIdx.x; i < M * K4; i += kThreads)
        reinterpret_cast<float4*>(xs)[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
    __syncthreads();
}
// layernorm of the M rows of x into shared memory: warp m takes row m (M <= 8 warps)
__device__ __forceinline__ void ln_rows(float* xs, const float* x, const float* w, const float* b, int M, int C) {
    const int warp = threadIdx.x >> 5;
    if (warp < M) pa_layernorm_row<true>(xs + (size_t)warp * C, x + (size_t)warp * C, w, b, C, threadIdx.x & 31);
    __syncthreads();
}

// out(m, n) for the M rows in shared memory and every feature n of w (N, K): FEAT features per warp
// pass, lanes interleave the 16-byte chunks of a row.  epi(m, n, dot) finishes one output.
template <int FEAT, typename Epi>
__device__ __forceinline__ void gemv_rows(const float* __restrict__ w, int N, int K, int M, const float* xs, Epi epi) {
    constexpr int UN = kLoads / FEAT;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    const int K4 = K >> 2;
    const float4* xs4 = reinterpret_cast<const float4*>(xs);
    for (int n0 = gw * FEAT; n0 < N; n0 += nw * FEAT) {
        float acc[FEAT][kMaxM];
#pragma unroll
        for (int f = 0; f < FEAT; ++f)
#pragma unroll
            for (int m = 0; m < kMaxM; ++m) acc[f][m] = 0.0f;
        const float4* wr[FEAT];
#pragma unroll
        for (int f = 0; f < FEAT; ++f) wr[f] = reinterpret_cast<const float4*>(w + (size_t)min(n0 + f, N - 1) * K);
        for (int cb = lane; cb < K4; cb += 32 * UN) {
            float4 wv[UN][FEAT];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int c = cb + 32 * u;
#pragma unroll
                for (int f = 0; f < FEAT; ++f) wv[u][f] = c < K4 ? __ldg(wr[f] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int c = cb + 32 * u;
                if (c < K4) {
#pragma unroll
                    for (int m = 0; m < kMaxM; ++m) {
                        if (m < M) {
                            const float4 xv = xs4[m * K4 + c];
#pragma unroll
                            for (int f = 0; f < FEAT; ++f)
                                acc[f][m] = fmaf(wv[u][f].w, xv.w, fmaf(wv[u][f].z, xv.z, fmaf(wv[u][f].y, xv.y, fmaf(wv[u][f].x, xv.x, acc[f][m]))));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < kMaxM; ++m) {
            if (m < M) {
#pragma unroll
                for (int f = 0; f < FEAT; ++f)
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) acc[f][m] += __shfl_xor_sync(0xffffffffu, acc[f][m], d);
            }
        }
        // lane (f, m) finishes output (m, n0 + f)
        const int f = lane >> 3, m = lane & 7;
        if (f < FEAT && m < M && n0 + f < N) {
            float v = 0.0f;
#pragma unroll
            for (int ff = 0; ff < FEAT; ++ff)
#pragma unroll
                for (int mm = 0; mm < kMaxM; ++mm)
                    if (ff == f && mm == m) v = acc[ff][mm];
            epi(m, n0 + f, v);
        }
    }
}
// features per warp pass: the fewest that still cover N in one pass of the grid's warps
template <typename Epi>
__device__ __forceinline__ void gemv_auto(const float* __restrict__ w, int N, int K, int M, const float* xs, Epi epi) {
    const int nw = gridDim.x * kWarps;
    if (N <= nw) gemv_rows<1>(w, N, K, M, xs, epi);
    else if (N <= 2 * nw) gemv_rows<2>(w, N, K, M, xs, epi);
    else gemv_rows<4>(w, N, K, M, xs, epi);
}

// ---- attention, phase 1: a warp per (sequence, head, chunk); LPT = hs/4 lanes per token ---------
template <int LPT>
__device__ __forceinline__ void attn_partials(const pa_mega_args& a, const float* pool_k, const float* pool_v) {
    constexpr int TPI = 32 / LPT;          // tokens per warp iteration (sub-groups of the warp)
    constexpr int UN = 8;                  // iterations in flight
    const int lane = threadIdx.x & 31, sub = lane / LPT, li = lane % LPT;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    const int hs = a.hs, C = a.C;
    const int n_units = a.M * a.NH * a.max_chunks;
    for (int u = gw; u < n_units; u += nw) {
        const int c = u % a.max_chunks, sh = u / a.max_chunks, h = sh % a.NH, s = sh / a.NH;
        const int first = a.kv_start[s], last = a.kv_end[s];
        const int t0 = first + c * a.chunk_tokens;
        if (t0 >= last) continue;                                   // warp-uniform
        const int t1 = min(last, t0 + a.chunk_tokens);
        const int* tbl = a.table + (size_t)s * a.tstride;
        const float4 q4 = __ldcg(reinterpret_cast<const float4*>(a.q + (size_t)s * C + h * hs) + li);
        float m_run = kMaxInit, l_run = 0.0f;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tb = t0; tb < t1; tb += TPI * UN) {
            float4 k4[UN], v4[UN];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int t = tb + i * TPI + sub;
                if (t < t1) {
                    const size_t off = ((size_t)tbl[t / a.bs] * a.bs + (t % a.bs)) * C + h * hs;
                    k4[i] = __ldcg(reinterpret_cast<const float4*>(pool_k + off) + li);
                    v4[i] = __ldcg(reinterpret_cast<const float4*>(pool_v + off) + li);
                } else {
                    k4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    v4[i] = k4[i];
                }
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int t = tb + i * TPI + sub;
                float dot = fmaf(q4.w, k4[i].w, fmaf(q4.z, k4[i].z, fmaf(q4.y, k4[i].y, q4.x * k4[i].x)));
#pragma unroll
                for (int d = LPT / 2; d >= 1; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
                if (t < t1) {                                       // uniform over the token's LPT lanes
                    const float sc = dot * a.scale;
                    const float m_new = fmaxf(m_run, sc);
                    const float alpha = expf(m_run - m_new), e = expf(sc - m_new);
                    l_run = l_run * alpha + e;
                    o4.x = fmaf(e, v4[i].x, o4.x * alpha); o4.y = fmaf(e, v4[i].y, o4.y * alpha);
                    o4.z = fmaf(e, v4[i].z, o4.z * alpha); o4.w = fmaf(e, v4[i].w, o4.w * alpha);
                    m_run = m_new;
                }
            }
        }
        float* pr = a.part + ((size_t)(sh * a.max_chunks + c) * TPI + sub) * (hs + 4);
        reinterpret_cast<float4*>(pr)[li] = o4;
        if (li == 0) { pr[hs] = m_run; pr[hs + 1] = l_run; }
    }
}

// ---- attention, phase 2: a warp per (sequence, head) merges its chunks' partials in order --------
template <int LPT>
__device__ __forceinline__ void attn_merge(const pa_mega_args& a) {
    constexpr int TPI = 32 / LPT;
    constexpr int UN = 8;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    const int hs = a.hs;
    for (int sh = gw; sh < a.M * a.NH; sh += nw) {
        const int s = sh / a.NH, h = sh % a.NH;
        const int len = a.kv_end[s] - a.kv_start[s];
        const int n_part = ((len + a.chunk_tokens - 1) / a.chunk_tokens) * TPI;
        const float* pr = a.part + (size_t)sh * a.max_chunks * TPI * (hs + 4);
        float m_tot = kMaxInit;
        for (int i = lane; i < n_part; i += 32) m_tot = fmaxf(m_tot, __ldcg(pr + (size_t)i * (hs + 4) + hs));
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m_tot = fmaxf(m_tot, __shfl_xor_sync(0xffffffffu, m_tot, d));
        // lane owns dims lane, lane + 32, ... of the head
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        float l_tot = 0.0f;
        for (int ib = 0; ib < n_part; ib += UN) {
            float pm[UN], pl[UN], po[UN][4];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const float* pp = pr + (size_t)min(ib + i, n_part - 1) * (hs + 4);
                pm[i] = __ldcg(pp + hs); pl[i] = __ldcg(pp + hs + 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) po[i][j] = (lane + 32 * j < hs) ? __ldcg(pp + lane + 32 * j) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                if (ib + i < n_part) {
                    const float wgt = expf(pm[i] - m_tot);
                    l_tot = fmaf(pl[i], wgt, l_tot);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = fmaf(po[i][j], wgt, o[j]);
                }
            }
        }
        const float inv = (l_tot == 0.0f) ? 0.0f : 1.0f / l_tot;        // :213
        float* out = a.atty + (size_t)s * a.C + h * hs;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (lane + 32 * j < hs) out[lane + 32 * j] = o[j] * inv;
    }
}

template <int LPT>
__global__ void __launch_bounds__(kThreads, 1)
pa_decode_step_mega_kernel(const pa_mega_args a) {
    extern __shared__ __align__(16) float xs[];          // [M][4C] staged / normalised input rows
    __shared__ PaSampleSmem<kThreads> samp;
    const int M = a.M, C = a.C;
    unsigned passed = 0;
    int n_stamp = 0;
    auto stamp = [&]() {            // PA_MEGA_DEBUG: CTA 0 records when it reaches each barrier and when it leaves it
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.dbg[n_stamp++] = t;
        }
    };
#define GRID_SYNC() do { stamp(); grid_sync(a.bar, passed); stamp(); } while (0)

    // encoder_forward (:24-46): CTA m writes row m of the residual stream
    if ((int)blockIdx.x < M) {
        const float* e = a.wte + (size_t)a.tokens[blockIdx.x] * C;
        const float* ps = a.wpe + (size_t)a.positions[blockIdx.x] * C;
        for (int i = threadIdx.x; i < C; i += kThreads) a.x[(size_t)blockIdx.x * C + i] = e[i] + ps[i];
    }
    GRID_SYNC();

    for (int l = 0; l < a.L; ++l) {
        float* pool_k = a.pool_k + (size_t)l * a.layer_stride;
        float* pool_v = a.pool_v + (size_t)l * a.layer_stride;
        // ln1 -> QKV projection; Q to the dense buffer, K and V straight to the token's page slot (:703-710)
        ln_rows(xs, a.x, a.ln1w + (size_t)l * C, a.ln1b + (size_t)l * C, M, C);
        {
            const float* bias = a.qkvb + (size_t)l * 3 * C;
            gemv_auto(a.qkvw + (size_t)l * 3 * C * C, 3 * C, C, M, xs, [&](int m, int n, float v) {
                v += bias[n];
                if (n < C) a.q[(size_t)m * C + n] = v;
                else {
                    const size_t slot_off = (size_t)a.slots[m] * C;
                    if (n < 2 * C) pool_k[slot_off + (n - C)] = v;
                    else pool_v[slot_off + (n - 2 * C)] = v;
                }
            });
        }
        GRID_SYNC();
        attn_partials<LPT>(a, pool_k, pool_v);
        GRID_SYNC();
        attn_merge<LPT>(a);
        GRID_SYNC();
        // x += atty . attprojw^T + attprojb (:716-717)
        stage_rows(xs, a.atty, M, C);
        {
            const float* bias = a.attprojb + (size_t)l * C;
            gemv_auto(a.attprojw + (size_t)l * C * C, C, C, M, xs, [&](int m, int n, float v) {
                float* xp = a.x + (size_t)m * C + n;
                *xp = v + bias[n] + __ldcg(xp);
            });
        }
        GRID_SYNC();
        // fch = gelu(ln2(x) . fcw^T + fcb) (:718-720)
        ln_rows(xs, a.x, a.ln2w + (size_t)l * C, a.ln2b + (size_t)l * C, M, C);
        {
            const float* bias = a.fcb + (size_t)l * 4 * C;
            gemv_auto(a.fcw + (size_t)l * 4 * C * C, 4 * C, C, M, xs, [&](int m, int n, float v) {
                a.fch[(size_t)m * 4 * C + n] = pa_gelu(v + bias[n]);
            });
        }
        GRID_SYNC();
        // x += fch . fcprojw^T + fcprojb (:721-722)
        stage_rows(xs, a.fch, M, 4 * C);
        {
            const float* bias = a.fcprojb + (size_t)l * C;
            gemv_auto(a.fcprojw + (size_t)l * 4 * C * C, C, 4 * C, M, xs, [&](int m, int n, float v) {
                float* xp = a.x + (size_t)m * C + n;
                *xp = v + bias[n] + __ldcg(xp);
            });
        }
        GRID_SYNC();
    }
    // final layernorm, logits = lnf . wte^T (:724-726), then softmax + sample_mult per row
    ln_rows(xs, a.x, a.lnfw, a.lnfb, M, C);
    gemv_rows<4>(a.wte, a.V, C, M, xs, [&](int m, int n, float v) { a.logits[(size_t)m * a.Vp + n] = v; });
    GRID_SYNC();
    if ((int)blockIdx.x < M)
        pa_sample_row<kThreads>(a.logits + (size_t)blockIdx.x * a.Vp, a.V, a.coins ? a.coins[blockIdx.x] : -1.0f, a.next + blockIdx.x, samp);
}

}  // namespace

extern "C" size_t pa_cu_model_mega_smem(int M, int C, int hs) {
    if (M < 1 || M > kMaxM || (C & 3) || C > 32 * kLnMaxPerLane || (hs != 64 && hs != 128)) return 0;
    const size_t bytes = (size_t)M * 4 * C * sizeof(float);
    return bytes <= 200 * 1024 ? bytes : 0;
}

extern "C" int pa_cu_model_mega_step(const pa_mega_args* a, void* stream) {
    const size_t smem = pa_cu_model_mega_smem(a->M, a->C, a->hs);
    if (!smem) return PA_ERR_UNSUPPORTED;
    auto fn = a->hs == 64 ? pa_decode_step_mega_kernel<16> : pa_decode_step_mega_kernel<32>;
    static size_t attr_smem[2] = {0, 0};
    size_t& cur = attr_smem[a->hs == 64 ? 0 : 1];
    if (smem > cur) {
        CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaMemsetAsync(a->bar, 0, sizeof(unsigned), s));
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
        const int per_layer = 6, n = 1 + per_layer * a->L + 1;       // barriers
        const char* names[per_layer] = {"qkv", "attn", "merge", "attproj", "fc", "fcproj"};
        double work[per_layer] = {0}, wait[per_layer] = {0};
        for (int l = 0; l < a->L; ++l)
            for (int p = 0; p < per_layer; ++p) {
                const int b = 1 + l * per_layer + p;                   // barrier index: stamps 2b (arrive), 2b+1 (leave)
                work[p] += (double)(hst[2 * b] - hst[2 * b - 1]);
                wait[p] += (double)(hst[2 * b + 1] - hst[2 * b]);
            }
        fprintf(stderr, "mega dbg: embed barrier %lld ns;", (long long)(hst[1] - hst[0]));
        for (int p = 0; p < per_layer; ++p) fprintf(stderr, " %s %.0f+%.0f", names[p], work[p] / a->L, wait[p] / a->L);
        fprintf(stderr, "; lm head %lld+%lld; total %lld ns\n", (long long)(hst[2 * (n - 1)] - hst[2 * (n - 1) - 1]),
                (long long)(hst[2 * (n - 1) + 1] - hst[2 * (n - 1)]), (long long)(hst[2 * (n - 1) + 1] - hst[0]));
    }
    return PA_OK;
}
