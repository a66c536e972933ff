/*
 * pa_model_dev.cuh -- device code shared by the per-op kernels of pa_model.cu and the persistent
 * small-batch step kernel of pa_model_mega.cu: the GELU of gelu_forward and softmax_forward +
 * sample_mult for one row of logits on a CTA.  Same statements in both users, so the two routes
 * through a decode step agree to the last bit in these ops.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

constexpr int kLnMaxPerLane = 64;          // the layernorm kernels keep a row in registers: C <= 32 * 64 = 2048

// gelu_forward, paged_infer.c:243-251 (tanh form)
static __device__ __forceinline__ float pa_gelu(float v) {
    const float cube = 0.044715f * v * v * v;
    return 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + cube)));
}

// ---- softmax_forward + sample_mult (paged_infer.c:259-286, :838-848) fused, one CTA of NT threads per row.
// maxval starts at -10000 as the reference's does; the probabilities are exp(l - max) / sum;
// sample_mult returns the first index whose running sum exceeds the coin.  Here the comparison is
// made against coin * sum (no division per element): a warp owns a contiguous range of the
// vocabulary (its lanes interleave, so reads coalesce), the range sums are scanned in index order
// and the warp whose range holds the crossing finds it with 32-wide inclusive scans -- an index can
// differ from the sequential reference only when the coin lies within rounding distance of a
// boundary of the distribution.  The probabilities are never written.  coin < 0 selects argmax
// (the FIRST maximum, as the reference's strict > keeps it).
template <int NT>
struct PaSampleSmem {
    float red[NT / 32];
    int redi[NT / 32];
    float wsum[NT / 32];
    int pick;
};
template <int NT>
static __device__ __forceinline__ void pa_sample_row(const float* __restrict__ l, int V, float coin, int* __restrict__ next_out,
                                                     PaSampleSmem<NT>& sm) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = ((V + NW - 1) / NW + 31) & ~31;
    const int c0 = min(V, warp * chunk), c1 = min(V, c0 + chunk);
    constexpr int kNone = 0x7fffffff;
    float mx = -10000.0f;                                           // :270
    int arg = kNone;
    for (int i = c0 + lane; i < c1; i += 32) { const float v = l[i]; if (v > mx) { mx = v; arg = i; } }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, d);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    if (lane == 0) { sm.red[warp] = mx; sm.redi[warp] = arg; }
    if (tid == 0) sm.pick = -1;
    __syncthreads();
    float maxval = sm.red[0];
    int argmax = sm.redi[0];
#pragma unroll
    for (int w = 1; w < NW; ++w)
        if (sm.red[w] > maxval || (sm.red[w] == maxval && sm.redi[w] < argmax)) { maxval = sm.red[w]; argmax = sm.redi[w]; }
    if (argmax == kNone) argmax = 0;
    if (coin < 0.0f) {                                              // uniform over the CTA
        if (tid == 0) *next_out = argmax;
        return;
    }
    float part = 0.0f;
    for (int i = c0 + lane; i < c1; i += 32) part += expf(l[i] - maxval);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) sm.wsum[warp] = part;
    __syncthreads();
    // running sums over the warp ranges, formed identically by every thread
    float before = 0.0f, upto = 0.0f, total = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w == warp) before = total;
        total += sm.wsum[w];
        if (w == warp) upto = total;
    }
    const float target = coin * total;
    if (target >= before && target < upto && c0 < c1) {             // the crossing lies in this warp's range (exactly one warp)
        float cdf = before;
        int pick = -1;
        for (int b = c0; b < c1 && pick < 0; b += 32) {
            const int i = b + lane;
            float sc = i < c1 ? expf(l[i] - maxval) : 0.0f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {                      // inclusive scan over the 32 entries
                const float t = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += t;
            }
            const float incl = cdf + sc;
            const unsigned hit = __ballot_sync(0xffffffffu, i < c1 && target < incl);
            if (hit) pick = b + __ffs(hit) - 1;
            cdf = __shfl_sync(0xffffffffu, incl, 31);
        }
        if (pick < 0) pick = c1 - 1;                                // the range sum and the scan round differently
        if (lane == 0) sm.pick = pick;
    }
    __syncthreads();
    if (tid == 0) *next_out = sm.pick >= 0 ? sm.pick : V - 1;       // "in case of rounding errors", :847
}
