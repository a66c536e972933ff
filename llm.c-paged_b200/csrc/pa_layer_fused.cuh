/* pa_layer_fused.cuh -- parameters of the resident per-layer grid (pa_layer_fused.cu), shared with its host in pa_model.cu */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

// one phase of the resident grid
struct pa_fused_phase {
    int kind;                     // 0: projection, 1: layernorm rows
    // projection: out = act(x . w^T + bias) + residual; columns >= n_dense go to the page slots (fused KV append)
    const CUtensorMap* tm_x;      // activation (rows, K), box 32 x 128, in device memory
    const CUtensorMap* tm_w;      // weight (N, K), box 32 x 64
    const float* bias;
    float* out;
    const float* residual;
    int out_stride, res_stride, N, K, n_dense, act;
    int tiles_n, n_split;         // tiles_n * n_split <= CTAs
    int slot;                     // arrival-counter set of this phase kind (monotonic across launches)
    unsigned gen;                 // completed uses of that set before this launch
    // layernorm: ln_out[r] = layernorm(ln_in[r]) for the M rows
    const float* ln_in;
    float* ln_out;
    const float* ln_w;
    const float* ln_b;
};
struct pa_fused_params {
    pa_fused_phase ph[6];
    int n_phases, M, C;
    float* pool_k;                // this layer's K / V pools and the step's slot mapping (QKV phase)
    float* pool_v;
    const int* slots;
    float4* ws;                   // split-K partial tiles [tile][split][float4 column][row]
    unsigned* ws_cnt;             // [6][kMaxTiles] arrival counters, never reset
    unsigned* bar;                // grid barrier counter, never reset
    unsigned bar_gen;             // barrier generations completed before this launch
    unsigned long long* dbg;      // optional timeline of CTA 0 (PA_FUSED_DEBUG=1): globaltimer at entry, then per phase: work done, barrier passed
};

