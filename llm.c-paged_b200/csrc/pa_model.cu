/*
 * pa_model.cu -- the rest of the decode step around the paged-attention path:
 * token + position embedding, layernorm, the dense projections (through pa_cu_linear: tensor-core
 * 3xTF32 GEMM with bias / GELU / residual epilogues), final layernorm, LM head and the sampler,
 * driven over ALL layers (the reference fork stops at `l < 1`, paged_infer.c:659) with one KV pool
 * per layer behind one block table.
 *
 * Reference functions restated on the device (all paged_infer.c): encoder_forward :24-46,
 * layernorm_forward :49-89, gelu_forward :243-251 (GEMM epilogue), residual_forward :253-257 (GEMM
 * epilogue), softmax_forward :259-286 + sample_mult :838-848 (one fused kernel: the probabilities
 * are never materialised), and the order of gpt2_forward :696-728.
 *
 * Parameters live in ONE device buffer in the checkpoint's tensor order (paged_infer.c:441-488):
 * wte, wpe, ln1w, ln1b, qkvw, qkvb, attprojw, attprojb, ln2w, ln2b, fcw, fcb, fcprojw, fcprojb,
 * lnfw, lnfb.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pa_internal.h"
#include "pa_pdl.cuh"
#include "pa_model_dev.cuh"
#include "pa_layer_fused.cuh"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

struct pa_model {
    pa_handle* h;
    pa_model_config cfg;
    int C, V, L, maxT, Vp;                 // Vp: logits row stride (V rounded up to a multiple of 4)
    float* params;                         // device, checkpoint order
    size_t n_params;
    const float *wte, *wpe, *ln1w, *ln1b, *qkvw, *qkvb, *attprojw, *attprojb, *ln2w, *ln2b, *fcw, *fcb, *fcprojw,
        *fcprojb, *lnfw, *lnfb;
    int max_batch;                         // most tokens (and sequences) in one step
    float *x, *ln, *q, *atty, *fch, *logits;     // device activations for one step
    int* d_io;                             // device: tokens[ntok] | positions[ntok] | last_row[nseq] | next[nseq]
    float* d_coins;
    int* h_io;                             // pinned mirror
    float* h_coins;
    float* mega_part;                      // attention split workspace of the persistent small-batch kernel
    size_t mega_part_floats;
    unsigned* mega_bar;                    // its grid-barrier counter (monotonic over launches) ...
    unsigned mega_bar_base;                // ... and its value when the next launch starts
    int mega_refused;                      // a cooperative launch failed once: stay on the chain (automatic choice only)
    int next_on_device;                    // the step's sampled tokens are wanted in device memory too (pa_group gathers them with NCCL)
    int* d_next_cur;                       // device: sampled tokens of the step in flight ...
    int* h_next_cur;                       // ... their pinned host copy ...
    int nseq_cur;                          // ... and how many (0: no step in flight)
    // resident per-layer grid (pa_layer_fused.cu): tensor maps in device memory, split-K workspace, never-reset counters
    CUtensorMap* fz_maps;                  // [3 + 4 L]: atty, ln, fch | per layer attprojw, fcw, fcprojw, qkvw
    float4* fz_ws;
    unsigned* fz_cnt;                      // [4 slots][max tiles] arrival counters | [1] grid barrier
    unsigned fz_gen[4];                    // completed uses of each counter slot
    unsigned fz_bar_gen;                   // grid-barrier generations completed
    int fz_state;                          // 0 not set up, 1 ready, -1 unavailable (shape outside its domain / set-up failed)
};

namespace {

// ---- encoder_forward (paged_infer.c:24-46) for one new token per sequence ------------------------
__global__ void pa_embed_kernel(float* __restrict__ x, const int* __restrict__ tokens, const int* __restrict__ positions,
                                const float* __restrict__ wte, const float* __restrict__ wpe, int C) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.x;
    const float* e = wte + (size_t)tokens[s] * C;
    const float* ps = wpe + (size_t)positions[s] * C;
    for (int i = threadIdx.x; i < C; i += blockDim.x) x[(size_t)s * C + i] = e[i] + ps[i];
}

// ---- layernorm_forward (paged_infer.c:49-89): one 128-thread CTA per row, the row held in registers ----
// (a single warp per row leaves every latency of the load -> sum -> variance -> scale chain exposed: 5.6 us
// for a 768-float row; four warps share it and meet twice through shared memory)
constexpr int kLnThreads = 128, kLnPerThread = 32 * kLnMaxPerLane / kLnThreads;       // C <= 2048
__device__ __forceinline__ void layernorm_row_block(float* __restrict__ o, const float* __restrict__ x, const float* __restrict__ weight,
                                                    const float* __restrict__ bias, int C) {
    __shared__ float red[2][kLnThreads / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float v[kLnPerThread];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnPerThread; ++i) {
        const int c = tid + kLnThreads * i;
        v[i] = c < C ? x[c] : 0.0f;
        sum += v[i];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) red[0][warp] = sum;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) tot += red[0][w];
    const float m = tot / C;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnPerThread; ++i) {
        const int c = tid + kLnThreads * i;
        const float dlt = v[i] - m;
        if (c < C) var += dlt * dlt;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    if (lane == 0) red[1][warp] = var;
    __syncthreads();
    float totv = 0.0f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) totv += red[1][w];
    const float s = 1.0f / sqrtf(totv / C + 1e-5f);              // eps, :56
#pragma unroll
    for (int i = 0; i < kLnPerThread; ++i) {
        const int c = tid + kLnThreads * i;
        if (c < C) o[c] = (s * (v[i] - m)) * weight[c] + bias[c];
    }
}
__global__ void __launch_bounds__(kLnThreads)
pa_layernorm_kernel(float* __restrict__ out, const float* __restrict__ inp, const float* __restrict__ weight,
                    const float* __restrict__ bias, int rows, int C) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    layernorm_row_block(out + (size_t)row * C, inp + (size_t)row * C, weight, bias, C);
}

// the same over gathered input rows (final layernorm of each sequence's last position), compact output
__global__ void __launch_bounds__(kLnThreads)
pa_layernorm_rows_kernel(float* __restrict__ out, const float* __restrict__ inp, const int* __restrict__ rows_in,
                         const float* __restrict__ weight, const float* __restrict__ bias, int rows, int C) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    layernorm_row_block(out + (size_t)row * C, inp + (size_t)rows_in[row] * C, weight, bias, C);
}

// ---- softmax_forward + sample_mult (paged_infer.c:259-286, :838-848) fused: one CTA per row of logits
// (pa_sample_row in pa_model_dev.cuh)
constexpr int kSampleThreads = 512;
__global__ void __launch_bounds__(kSampleThreads)
pa_sample_kernel(const float* __restrict__ logits, int stride, int V, const float* __restrict__ coins, int* __restrict__ next) {
    __shared__ PaSampleSmem<kSampleThreads> sm;
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    pa_sample_row<kSampleThreads>(logits + (size_t)row * stride, V, coins ? coins[row] : -1.0f, next + row, sm);
}

size_t param_count(int V, int maxT, int L, int C) {
    return (size_t)V * C + (size_t)maxT * C + (size_t)L * (2 * C + 3 * (size_t)C * C + 3 * C + (size_t)C * C + C + 2 * C +
                                                            4 * (size_t)C * C + 4 * C + 4 * (size_t)C * C + C) + 2 * C;
}

// synthetic random-init weights on the device (no checkpoint is available offline): N(0, std) from a
// counter-based hash; layernorm weights 1, biases 0 -- GPT-2's initialisation scheme
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}
__global__ void pa_init_normal_kernel(float* p, size_t n, float stdv, float mean, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float u1 = (hash32(seed + 2 * i) >> 8) * (1.0f / 16777216.0f) + 1e-7f;
        const float u2 = (hash32(seed + 2 * i + 1) >> 8) * (1.0f / 16777216.0f);
        p[i] = mean + stdv * sqrtf(-2.0f * logf(u1)) * cosf(6.28318530718f * u2);
    }
}

// the decode step of <= 8 sequences through the persistent kernel; the step tables are already on the device
int mega_step(pa_model* m, int nseq, const int* tok, const int* pos, const float* coins, int* next_mapped, cudaStream_t s) {
    pa_handle* h = m->h;
    const pa_step_layout& L = h->step;
    pa_mega_args a;
    a.wte = m->wte; a.wpe = m->wpe; a.ln1w = m->ln1w; a.ln1b = m->ln1b; a.qkvw = m->qkvw; a.qkvb = m->qkvb;
    a.attprojw = m->attprojw; a.attprojb = m->attprojb; a.ln2w = m->ln2w; a.ln2b = m->ln2b; a.fcw = m->fcw; a.fcb = m->fcb;
    a.fcprojw = m->fcprojw; a.fcprojb = m->fcprojb; a.lnfw = m->lnfw; a.lnfb = m->lnfb;
    a.C = m->C; a.NH = h->cfg.n_heads; a.hs = h->cfg.head_dim; a.L = m->L; a.V = m->V; a.Vp = m->Vp;
    a.M = nseq; a.use_coins = coins != nullptr; a.next = next_mapped;
    for (int i = 0; i < nseq; ++i) { a.tokens[i] = tok[i]; a.positions[i] = pos[i]; a.coins[i] = coins ? coins[i] : -1.0f; }
    a.x = m->x; a.q = m->q; a.atty = m->atty; a.fch = m->fch; a.logits = m->logits;
    a.pool_k = h->pool_k; a.pool_v = h->pool_v; a.layer_stride = h->layer_stride;
    a.kv_end = h->d_step + L.off_kv_end; a.kv_start = h->d_step + L.off_kv_start;
    a.slots = h->d_step + L.off_slot; a.table = h->d_step + L.off_table;
    a.tstride = L.tstride; a.bs = h->cfg.block_size;
    a.bs_shift = 0;
    while ((1 << a.bs_shift) < a.bs) ++a.bs_shift;
    a.scale = (float)(1.0 / sqrtf((float)h->cfg.head_dim));            // :174
    a.sm_count = h->sm_count;
    // attention split: about one (sequence, head, chunk) unit per warp of the grid, chunks of >= 16 tokens
    long long tokens = 0;
    int max_len = 1;
    for (int i = 0; i < nseq; ++i) {
        const int len = h->h_step[L.off_kv_end + i] - h->h_step[L.off_kv_start + i];
        tokens += len;
        if (len > max_len) max_len = len;
    }
    const long long warps = (long long)h->sm_count * PA_MEGA_WARPS_PER_SM;
    long long chunk = (tokens * a.NH + warps - 1) / warps;
    chunk = ((chunk + 15) / 16) * 16;
    if (chunk < 16) chunk = 16;
    a.chunk_tokens = (int)chunk;
    a.max_chunks = (max_len + a.chunk_tokens - 1) / a.chunk_tokens;
    // contexts up to 32 tokens per warp of a CTA (measured: 2636 vs 2441 tokens/s at 256, 2026 vs 2122 at 1024): one CTA per (sequence, head), merged in shared memory
    static const int local_max = getenv("PA_MEGA_LOCAL_ATTN_MAX") ? atoi(getenv("PA_MEGA_LOCAL_ATTN_MAX")) : 32 * PA_MEGA_WARPS_PER_SM;
    a.local_attn = max_len <= local_max;
    if (a.local_attn) a.max_chunks = PA_MEGA_WARPS_PER_SM;
    const int tpi = 32 / (a.hs / 4);
    const size_t need = (size_t)nseq * a.NH * a.max_chunks * tpi * (a.hs + 4);
    if (need > m->mega_part_floats) {
        CU_CHECK(cudaStreamSynchronize(s));
        cudaFree(m->mega_part);
        m->mega_part = nullptr; m->mega_part_floats = 0;
        CU_CHECK(cudaMalloc((void**)&m->mega_part, need * 2 * sizeof(float)));
        m->mega_part_floats = need * 2;
    }
    if (!m->mega_bar) {
        CU_CHECK(cudaMalloc((void**)&m->mega_bar, sizeof(unsigned)));
        CU_CHECK(cudaMemsetAsync(m->mega_bar, 0, sizeof(unsigned), s));
        m->mega_bar_base = 0;
    }
    // the barrier counter only ever grows: this launch counts from where the last one stopped
    a.part = m->mega_part; a.bar = m->mega_bar; a.bar_base = m->mega_bar_base;
    const int rc = pa_cu_model_mega_step(&a, s);
    if (rc == PA_OK) m->mega_bar_base += (unsigned)(2 + (a.local_attn ? 5 : 6) * m->L + (getenv("PA_MEGA_DEBUG") ? 16 : 0)) * (unsigned)h->sm_count;      // (only a launch that ran counts)
    return rc;
}

// ---- the resident per-layer grid: set-up (once per model) and one launch -------------------------------------------
enum { FZ_ATTY = 0, FZ_LN = 1, FZ_FCH = 2, FZ_W0 = 3 };      // map indices; per layer: +0 attprojw, +1 fcw, +2 fcprojw, +3 qkvw
enum { FZ_SLOT_ATTPROJ = 0, FZ_SLOT_FC = 1, FZ_SLOT_FCPROJ = 2, FZ_SLOT_QKV = 3 };

bool fused_domain(const pa_model* m, int ntok) {
    const int C = m->C, sms = m->h->sm_count;
    return ntok >= 1 && ntok <= 128 && (C % 64) == 0 && C <= 2048 && (4 * C) / 64 <= sms && sms <= pa_cu_layer_fused_max_tiles() &&
           (size_t)m->h->smem_optin >= pa_cu_layer_fused_smem();
}

int fused_setup(pa_model* m) {
    if (m->fz_state != 0) return m->fz_state > 0 ? PA_OK : PA_ERR_UNSUPPORTED;
    m->fz_state = -1;
    const int C = m->C, L = m->L, sms = m->h->sm_count;
    const int rows = m->max_batch < 128 ? 128 : m->max_batch;      // (the activation buffers are at least this tall: see pa_model_create)
    const int n_maps = FZ_W0 + 4 * L;
    std::vector<CUtensorMap> maps(n_maps);
    int rc = pa_cu_make_map_2d(&maps[FZ_ATTY], m->atty, m->max_batch, C, C, 128);
    if (rc == PA_OK) rc = pa_cu_make_map_2d(&maps[FZ_LN], m->ln, m->max_batch, C, C, 128);
    if (rc == PA_OK) rc = pa_cu_make_map_2d(&maps[FZ_FCH], m->fch, m->max_batch, 4 * C, 4 * C, 128);
    for (int l = 0; l < L && rc == PA_OK; ++l) {
        rc = pa_cu_make_map_2d(&maps[FZ_W0 + 4 * l + 0], m->attprojw + (size_t)l * C * C, C, C, C, 64);
        if (rc == PA_OK) rc = pa_cu_make_map_2d(&maps[FZ_W0 + 4 * l + 1], m->fcw + (size_t)l * 4 * C * C, 4 * C, C, C, 64);
        if (rc == PA_OK) rc = pa_cu_make_map_2d(&maps[FZ_W0 + 4 * l + 2], m->fcprojw + (size_t)l * 4 * C * C, C, 4 * C, 4 * C, 64);
        if (rc == PA_OK) rc = pa_cu_make_map_2d(&maps[FZ_W0 + 4 * l + 3], m->qkvw + (size_t)l * 3 * C * C, 3 * C, C, C, 64);
    }
    (void)rows;
    if (rc != PA_OK) return rc;
    const size_t cnt_words = (size_t)4 * pa_cu_layer_fused_max_tiles() + 32;
    cudaError_t e = cudaMalloc((void**)&m->fz_maps, (size_t)n_maps * sizeof(CUtensorMap));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->fz_ws, (size_t)sms * 16 * 128 * sizeof(float4));      // one 64-column partial tile per CTA
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->fz_cnt, cnt_words * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemcpy(m->fz_maps, maps.data(), (size_t)n_maps * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(m->fz_cnt, 0, cnt_words * sizeof(unsigned));
    if (e != cudaSuccess) {
        pa_set_error("fused layer grid: set-up failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PA_ERR_NOMEM;
    }
    memset(m->fz_gen, 0, sizeof(m->fz_gen));
    m->fz_bar_gen = 0;
    m->fz_state = 1;
    return PA_OK;
}

// one launch: [attproj, ln2, fc, fcproj] of layer l_post (when >= 0) followed by [ln1, QKV + KV append] of layer l_pre (when < L)
int fused_launch(pa_model* m, int l_post, int l_pre, int ntok, cudaStream_t s) {
    pa_handle* h = m->h;
    const int C = m->C, sms = h->sm_count, mt = pa_cu_layer_fused_max_tiles();
    pa_fused_params P;
    memset(&P, 0, sizeof(P));
    int n = 0;
    auto proj = [&](int slot, int map_x, int map_w, const float* bias, float* out, int out_stride, int N, int K, int n_dense,
                    const float* residual, int act) {
        pa_fused_phase& ph = P.ph[n++];
        ph.kind = 0;
        ph.tm_x = m->fz_maps + map_x; ph.tm_w = m->fz_maps + map_w;
        ph.bias = bias; ph.out = out; ph.out_stride = out_stride; ph.N = N; ph.K = K; ph.n_dense = n_dense;
        ph.residual = residual; ph.res_stride = out_stride; ph.act = act;
        ph.n_split = pa_cu_layer_fused_split(N, K, sms, &ph.tiles_n);
        ph.slot = slot; ph.gen = m->fz_gen[slot]++;
    };
    auto lnorm = [&](const float* w, const float* b) {
        pa_fused_phase& ph = P.ph[n++];
        ph.kind = 1;
        ph.ln_in = m->x; ph.ln_out = m->ln; ph.ln_w = w; ph.ln_b = b;
    };
    if (l_post >= 0) {
        const int l = l_post;
        proj(FZ_SLOT_ATTPROJ, FZ_ATTY, FZ_W0 + 4 * l + 0, m->attprojb + (size_t)l * C, m->x, C, C, C, C, m->x, 0);              // :716-717
        lnorm(m->ln2w + (size_t)l * C, m->ln2b + (size_t)l * C);                                                               // :718
        proj(FZ_SLOT_FC, FZ_LN, FZ_W0 + 4 * l + 1, m->fcb + (size_t)l * 4 * C, m->fch, 4 * C, 4 * C, C, 4 * C, nullptr, 1);     // :719-720
        proj(FZ_SLOT_FCPROJ, FZ_FCH, FZ_W0 + 4 * l + 2, m->fcprojb + (size_t)l * C, m->x, C, C, 4 * C, C, m->x, 0);            // :721-722
    }
    if (l_pre < m->L) {
        const int l = l_pre;
        lnorm(m->ln1w + (size_t)l * C, m->ln1b + (size_t)l * C);                                                               // :703
        proj(FZ_SLOT_QKV, FZ_LN, FZ_W0 + 4 * l + 3, m->qkvb + (size_t)l * 3 * C, m->q, C, 3 * C, C, C, nullptr, 0);             // :706 + add_to_cache :710
        P.pool_k = h->pool_k + (size_t)l * h->layer_stride;
        P.pool_v = h->pool_v + (size_t)l * h->layer_stride;
        P.slots = h->d_step + h->step.off_slot;
    }
    P.n_phases = n; P.M = ntok; P.C = C;
    P.ws = m->fz_ws; P.ws_cnt = m->fz_cnt; P.bar = m->fz_cnt + (size_t)4 * mt;
    P.bar_gen = m->fz_bar_gen;
    m->fz_bar_gen += (unsigned)(n - 1);
    // the CTAs wait for each other: co-resident by construction (grid = SMs, one CTA per SM); cooperative (the driver checks
    // it) whenever the chain's launch overlap is switched off anyway
    const int rc = pa_cu_layer_fused_launch(&P, sms, !(pa_pdl_enabled && pa_pdl_gate), s);
    if (rc == PA_OK) h->launches++;
    return rc;
}

}  // namespace

extern "C" {

size_t pa_model_param_count(const pa_model_config* c) {
    return c ? param_count(c->vocab_size, c->max_seq_len, c->n_layers, c->channels) : 0;
}

int pa_model_create(pa_handle* h, const pa_model_config* cfg, const float* params_host, unsigned long long seed,
                    int max_batch, pa_model** out) {
    if (!h || !cfg || !out) { pa_set_error("pa_model_create: NULL argument"); return PA_ERR_INVALID; }
    *out = nullptr;
    if (h->host_only || !h->pool_k) { pa_set_error("pa_model_create: handle has no device; there is no CPU fallback"); return PA_ERR_NO_DEVICE; }
    if (cfg->channels != h->C || cfg->n_heads != h->cfg.n_heads || cfg->n_layers != h->cfg.n_layers) {
        pa_set_error("pa_model_create: model (C=%d NH=%d L=%d) does not match the KV handle (C=%d NH=%d L=%d)", cfg->channels,
                     cfg->n_heads, cfg->n_layers, h->C, h->cfg.n_heads, h->cfg.n_layers);
        return PA_ERR_INVALID;
    }
    if (cfg->channels > 32 * kLnMaxPerLane || cfg->vocab_size < 1 || cfg->max_seq_len < 1 || max_batch < 1) {
        pa_set_error("pa_model_create: unsupported geometry");
        return PA_ERR_INVALID;
    }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    pa_model* m = (pa_model*)calloc(1, sizeof(pa_model));
    if (!m) { pa_set_error("pa_model_create: out of host memory"); return PA_ERR_NOMEM; }
    m->h = h; m->cfg = *cfg;
    const int C = m->C = cfg->channels, V = m->V = cfg->vocab_size, L = m->L = cfg->n_layers, maxT = m->maxT = cfg->max_seq_len;
    m->Vp = (V + 3) & ~3;
    m->max_batch = max_batch;
    m->n_params = param_count(V, maxT, L, C);
    cudaError_t e = cudaMalloc((void**)&m->params, m->n_params * sizeof(float));
    const size_t B = (size_t)max_batch;
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->x, B * C * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->ln, B * C * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->q, B * C * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->atty, B * C * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->fch, B * 4 * C * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->logits, B * m->Vp * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_io, B * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_coins, B * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&m->h_io, B * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&m->h_coins, B * sizeof(float));
    if (e != cudaSuccess) {
        pa_set_error("pa_model_create: allocation failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        pa_model_destroy(m);
        return PA_ERR_NOMEM;
    }
    float* p = m->params;
    m->wte = p; p += (size_t)V * C;
    m->wpe = p; p += (size_t)maxT * C;
    m->ln1w = p; p += (size_t)L * C;
    m->ln1b = p; p += (size_t)L * C;
    m->qkvw = p; p += (size_t)L * 3 * C * C;
    m->qkvb = p; p += (size_t)L * 3 * C;
    m->attprojw = p; p += (size_t)L * C * C;
    m->attprojb = p; p += (size_t)L * C;
    m->ln2w = p; p += (size_t)L * C;
    m->ln2b = p; p += (size_t)L * C;
    m->fcw = p; p += (size_t)L * 4 * C * C;
    m->fcb = p; p += (size_t)L * 4 * C;
    m->fcprojw = p; p += (size_t)L * C * 4 * C;
    m->fcprojb = p; p += (size_t)L * C;
    m->lnfw = p; p += C;
    m->lnfb = p; p += C;
    cudaStream_t s = (cudaStream_t)h->stream;
    if (params_host) {
        CU_CHECK(cudaMemcpyAsync(m->params, params_host, m->n_params * sizeof(float), cudaMemcpyHostToDevice, s));
    } else {
        // GPT-2 style random init: weights N(0, 0.02), layernorm weights 1, every bias 0
        struct { const float* ptr; size_t n; float stdv, mean; } parts[] = {
            {m->wte, (size_t)V * C, 0.02f, 0.f}, {m->wpe, (size_t)maxT * C, 0.02f, 0.f},
            {m->ln1w, (size_t)L * C, 0.f, 1.f}, {m->ln1b, (size_t)L * C, 0.f, 0.f},
            {m->qkvw, (size_t)L * 3 * C * C, 0.02f, 0.f}, {m->qkvb, (size_t)L * 3 * C, 0.f, 0.f},
            {m->attprojw, (size_t)L * C * C, 0.02f, 0.f}, {m->attprojb, (size_t)L * C, 0.f, 0.f},
            {m->ln2w, (size_t)L * C, 0.f, 1.f}, {m->ln2b, (size_t)L * C, 0.f, 0.f},
            {m->fcw, (size_t)L * 4 * C * C, 0.02f, 0.f}, {m->fcb, (size_t)L * 4 * C, 0.f, 0.f},
            {m->fcprojw, (size_t)L * 4 * C * C, 0.02f, 0.f}, {m->fcprojb, (size_t)L * C, 0.f, 0.f},
            {m->lnfw, (size_t)C, 0.f, 1.f}, {m->lnfb, (size_t)C, 0.f, 0.f}};
        uint64_t sd = seed * 0x9e3779b97f4a7c15ull + 1;
        for (auto& pt : parts) {
            pa_init_normal_kernel<<<592, 256, 0, s>>>(const_cast<float*>(pt.ptr), pt.n, pt.stdv, pt.mean, sd);
            sd += 2 * pt.n + 17;
        }
        CU_CHECK(cudaGetLastError());
    }
    CU_CHECK(cudaStreamSynchronize(s));
    *out = m;
    return PA_OK;
}

/* N(mean, stdv) from the counter-based hash above, written by the device: synthetic pools and activations of
 * any size without a host copy (a 160 GB pool is filled at HBM speed instead of over PCIe); the same
 * (seed, index) always gives the same value, so a checker can regenerate any element on the host side */
int pa_fill_normal(float* dev, size_t n, float stdv, float mean, unsigned long long seed, void* stream) {
    if (!dev) { pa_set_error("pa_fill_normal: NULL pointer"); return PA_ERR_INVALID; }
    if (n == 0) return PA_OK;
    pa_init_normal_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(dev, n, stdv, mean, (uint64_t)seed);
    CU_CHECK(cudaGetLastError());
    return PA_OK;
}

void pa_model_destroy(pa_model* m) {
    if (!m) return;
    cudaFree(m->params); cudaFree(m->x); cudaFree(m->ln); cudaFree(m->q); cudaFree(m->atty); cudaFree(m->fch);
    cudaFree(m->logits); cudaFree(m->d_io); cudaFree(m->d_coins); cudaFree(m->mega_part); cudaFree(m->mega_bar);
    cudaFree(m->fz_maps); cudaFree(m->fz_ws); cudaFree(m->fz_cnt);
    if (m->h_io) cudaFreeHost(m->h_io);
    if (m->h_coins) cudaFreeHost(m->h_coins);
    free(m);
}

float* pa_model_params(pa_model* m) { return m ? m->params : nullptr; }
float* pa_model_logits(pa_model* m, int* stride) {
    if (!m) return nullptr;
    if (stride) *stride = m->Vp;
    return m->logits;
}

/* The forward pass for a step of the batch: sequence seq_ids[i] receives n_new[i] >= 1 tokens
 * (its prompt, a chunk of it, or one decode token; packed in step order in `tokens`) at its next
 * positions; next_tokens[i] is sampled from the logits of its LAST new position with coins[i] in
 * [0,1) (NULL: argmax).  Runs: page choice + table mirror, embedding, then per layer ln1 -> QKV
 * projection with fused KV append -> paged attention (decode kernel when every sequence has one
 * new token, else the causal prefill kernel) -> attproj (+residual) -> ln2 -> fc (+GELU) -> fcproj
 * (+residual); final layernorm, LM head and sampler on the last rows only. */
static int model_forward_impl(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins, int nseq);

/* Enqueue the whole step on the handle's stream and return: no host synchronisation.  pa_model_wait
 * finishes it.  A host driving several GPUs queues every GPU's step (and the NCCL gather behind it,
 * pa_group_model_step) before it waits for any of them. */
int pa_model_forward_async(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins, int nseq) {
    if (m && m->nseq_cur) { pa_set_error("pa_model_forward_async: the previous step was not finished with pa_model_wait"); return PA_ERR_INVALID; }
    const int rc = model_forward_impl(m, seq_ids, n_new, tokens, coins, nseq);
    pa_pdl_gate = 1;          // whatever path the step left by, the gate does not outlive it
    return rc;
}
int pa_model_wait(pa_model* m, int* next_tokens) {
    if (!m || !m->nseq_cur) { pa_set_error("pa_model_wait: no step in flight"); return PA_ERR_INVALID; }
    const int n = m->nseq_cur;
    m->nseq_cur = 0;
    CU_CHECK(cudaSetDevice(m->h->cfg.device));
    CU_CHECK(cudaStreamSynchronize((cudaStream_t)m->h->stream));
    if (next_tokens) memcpy(next_tokens, m->h_next_cur, (size_t)n * sizeof(int));
    return PA_OK;
}
int* pa_model_next_tokens_dev(pa_model* m) { return m ? m->d_next_cur : nullptr; }
void pa_model_want_device_tokens(pa_model* m, int on) { if (m) m->next_on_device = on ? 1 : 0; }
pa_handle* pa_model_handle(pa_model* m) { return m ? m->h : nullptr; }

int pa_model_forward(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins, int nseq,
                     int* next_tokens) {
    if (!next_tokens) { pa_set_error("pa_model_forward: bad arguments"); return PA_ERR_INVALID; }
    const int rc = pa_model_forward_async(m, seq_ids, n_new, tokens, coins, nseq);
    if (rc != PA_OK) return rc;
    return pa_model_wait(m, next_tokens);
}

static int model_forward_impl(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins, int nseq) {
    if (!m || !seq_ids || !n_new || !tokens || nseq < 1) {
        pa_set_error("pa_model_forward: bad arguments");
        return PA_ERR_INVALID;
    }
    pa_handle* h = m->h;
    const int C = m->C, L = m->L, V = m->V;
    long long ntok_ll = 0;
    int max_q = 0;
    for (int i = 0; i < nseq; ++i) {
        if (n_new[i] < 1) { pa_set_error("pa_model_forward: n_new[%d] = %d (every sequence of the step needs a token)", i, n_new[i]); return PA_ERR_INVALID; }
        ntok_ll += n_new[i];
        if (n_new[i] > max_q) max_q = n_new[i];
    }
    if (ntok_ll > m->max_batch || nseq > m->max_batch) {
        pa_set_error("pa_model_forward: %lld tokens in the step, the model was created for %d", ntok_ll, m->max_batch);
        return PA_ERR_INVALID;
    }
    const int ntok = (int)ntok_ll;
    for (int row = 0; row < ntok; ++row)
        if (tokens[row] < 0 || tokens[row] >= V) { pa_set_error("pa_model_forward: token %d out of range", tokens[row]); return PA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)h->stream;
    pa_pdl_enabled = h->tune[PA_TUNE_NO_PDL] ? 0 : 1;
    // The page choice comes FIRST: pa_step_begin may bring a swapped-out sequence back (its cached length is 0
    // until then), so the positions of the new tokens are only known from the step tables it builds:
    // position of a sequence's first new token = kv_end - n_new.
    int rc = pa_step_begin(h, seq_ids, n_new, nseq);
    if (rc != PA_OK) return rc;
    // host io layout: tokens[ntok] | positions[ntok] | last_row[nseq] | next[nseq]
    int* h_tok = m->h_io, *h_pos = m->h_io + ntok, *h_last = m->h_io + 2 * ntok, *h_next = m->h_io + 2 * ntok + nseq;
    for (int i = 0, row = 0; i < nseq; ++i) {
        const int end = h->h_step[h->step.off_kv_end + i];
        const int pos0 = end - n_new[i];
        if (pos0 < 0 || pos0 != h->step_pre_len[i] || end > m->maxT) {
            if (end > m->maxT) pa_set_error("pa_model_forward: sequence %d would reach position %d of %d", seq_ids[i], end, m->maxT);
            else pa_set_error("pa_model_forward: sequence %d changed length while the step was built (%d cached, %d new, %d before)", seq_ids[i], end, n_new[i], h->step_pre_len[i]);
            pa_step_rollback(h);
            return PA_ERR_INVALID;
        }
        for (int j = 0; j < n_new[i]; ++j, ++row) {
            h_tok[row] = tokens[row];
            h_pos[row] = pos0 + j;
        }
        h_last[i] = row - 1;
        if (coins) m->h_coins[i] = coins[i];
    }
    // A handful of sequences with one new token each: the whole step is ONE persistent kernel (pa_model_mega.cu)
    const int model_path = h->tune[PA_TUNE_MODEL_PATH];
    bool use_mega = false;
    if (model_path != 1 && model_path != 3) {
        const size_t mega_smem = (max_q == 1 && nseq <= (model_path == 2 ? PA_MEGA_MAX_SEQS : PA_MEGA_AUTO_SEQS)) ? pa_cu_model_mega_smem(nseq, C, h->cfg.head_dim, h->cfg.block_size) : 0;
        use_mega = mega_smem && (size_t)h->smem_optin >= mega_smem + 1024 && !(m->mega_refused && model_path != 2);
        if (!use_mega && model_path == 2) {
            pa_set_error("pa_model_forward: the persistent step kernel takes at most %d sequences of one new token each (head_dim 64 or 128)", PA_MEGA_MAX_SEQS);
            pa_step_rollback(h);
            return PA_ERR_UNSUPPORTED;
        }
    }
    // Overlapping launches (programmatic dependent launch) along the step's chain of kernels.  While the split-K
    // projections were cluster launches the overlap cost time beyond 128 tokens (early-resident successors got in
    // their way); with the workspace split it pays at every size measured (2.84 -> 2.74 ms at 256 tokens).
    static const int pdl_max_tokens = getenv("PA_PDL_MAX_TOKENS") ? atoi(getenv("PA_PDL_MAX_TOKENS")) : (1 << 30);
    pa_pdl_gate = ntok <= pdl_max_tokens;
    rc = pa_step_upload(h, s);
    if (rc != PA_OK) return rc;
    if (use_mega) {
        // token ids, positions and coins travel as kernel arguments, the sampled tokens come back through
        // mapped pinned memory: the step is ONE table copy, ONE launch and ONE synchronisation
        // (with next_on_device they go to device memory first -- NCCL gathers from there -- and are copied back)
        int* d_next_mega = m->d_io + 2 * ntok + nseq;
        rc = mega_step(m, nseq, h_tok, h_pos, coins ? m->h_coins : nullptr, m->next_on_device ? d_next_mega : h_next, s);
        if (rc == PA_OK) {
            h->launches += 1;
            if (m->next_on_device) CU_CHECK(cudaMemcpyAsync(h_next, d_next_mega, (size_t)nseq * sizeof(int), cudaMemcpyDeviceToHost, s));
            m->d_next_cur = d_next_mega; m->h_next_cur = h_next; m->nseq_cur = nseq;
            return PA_OK;
        }
        if (model_path == 2) return rc;
        // chosen automatically and the cooperative launch was refused (no co-residency to be had here): the chain
        // of per-op kernels takes this step and the following ones
        cudaGetLastError();
        m->mega_refused = 1;
    }
    int* d_tok = m->d_io, *d_pos = m->d_io + ntok, *d_last = m->d_io + 2 * ntok, *d_next = m->d_io + 2 * ntok + nseq;
    CU_CHECK(cudaMemcpyAsync(m->d_io, m->h_io, (size_t)(2 * ntok + nseq) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (coins) CU_CHECK(cudaMemcpyAsync(m->d_coins, m->h_coins, (size_t)nseq * sizeof(float), cudaMemcpyHostToDevice, s));
    CU_CHECK(pa_launch_pdl(pa_embed_kernel, dim3(ntok), dim3(256), 0, s, 1, m->x, (const int*)d_tok, (const int*)d_pos, m->wte, m->wpe, C));
    const int path = h->tune[PA_TUNE_GEMM_PATH];
    const int ln_grid = ntok;                      // one CTA per row
    long launches = 1;
    // model_path 3 (opt-in): everything between two attention launches as ONE resident grid (pa_layer_fused.cu), 2
    // launches per layer instead of 7.  Measured SLOWER than the chain (cfg2: 1.675 vs 1.528 ms per step): with
    // programmatic dependent launch the chain's launch overheads were already hidden, and a projection's ~8-10 us are
    // its own latency chain (first TMA round trip, k-slabs at 0.4 us, split-K publish / poll / reduce), which a phase
    // of the resident grid pays just the same, plus a grid barrier (profiles/r02_model_step.md).
    bool use_fused = model_path == 3 && path == 0 && fused_domain(m, ntok);
    if (use_fused && fused_setup(m) != PA_OK) use_fused = false;
    if (model_path == 3 && !use_fused) {
        pa_set_error("pa_model_forward: the resident per-layer grid takes steps of at most 128 tokens, C a multiple of 64 up to 2048, the automatic projection path");
        pa_step_rollback(h);
        return PA_ERR_UNSUPPORTED;
    }
    if (use_fused) {
        rc = fused_launch(m, -1, 0, ntok, s);                                  // ln1 + QKV of layer 0
        for (int l = 0; l < L && rc == PA_OK; ++l) {
            rc = max_q == 1 ? pa_decode(h, l, m->q, C, m->atty, C, s) : pa_prefill(h, l, m->q, C, m->atty, C, s);
            if (rc == PA_OK) rc = fused_launch(m, l, l + 1, ntok, s);         // the rest of layer l, ln1 + QKV of layer l + 1
        }
        if (rc != PA_OK) return rc;
    }
    for (int l = 0; l < L && !use_fused; ++l) {
        CU_CHECK(pa_launch_pdl(pa_layernorm_kernel, dim3(ln_grid), dim3(kLnThreads), 0, s, 1, m->ln, (const float*)m->x, m->ln1w + (size_t)l * C, m->ln1b + (size_t)l * C, ntok, C));
        rc = pa_qkv_append(h, l, m->ln, C, m->qkvw + (size_t)l * 3 * C * C, m->qkvb + (size_t)l * 3 * C, m->q, C, s);
        if (rc != PA_OK) return rc;
        rc = max_q == 1 ? pa_decode(h, l, m->q, C, m->atty, C, s) : pa_prefill(h, l, m->q, C, m->atty, C, s);
        if (rc != PA_OK) return rc;
        // x += atty . attprojw^T + attprojb      (matmul_forward + residual_forward, :716-717)
        rc = pa_cu_linear(m->atty, C, m->attprojw + (size_t)l * C * C, m->attprojb + (size_t)l * C, m->x, C, ntok, C, C, m->x, C, 0, path, s);
        if (rc != PA_OK) return rc;
        CU_CHECK(pa_launch_pdl(pa_layernorm_kernel, dim3(ln_grid), dim3(kLnThreads), 0, s, 1, m->ln, (const float*)m->x, m->ln2w + (size_t)l * C, m->ln2b + (size_t)l * C, ntok, C));
        // fch = gelu(ln . fcw^T + fcb)           (:719-720)
        rc = pa_cu_linear(m->ln, C, m->fcw + (size_t)l * 4 * C * C, m->fcb + (size_t)l * 4 * C, m->fch, 4 * C, ntok, 4 * C, C, nullptr, 0, 1, path, s);
        if (rc != PA_OK) return rc;
        // x += fch . fcprojw^T + fcprojb         (:721-722)
        rc = pa_cu_linear(m->fch, 4 * C, m->fcprojw + (size_t)l * 4 * C * C, m->fcprojb + (size_t)l * C, m->x, C, ntok, C, 4 * C, m->x, C, 0, path, s);
        if (rc != PA_OK) return rc;
        launches += 5;
    }
    // only each sequence's last new position feeds the LM head: final layernorm over the gathered rows
    CU_CHECK(pa_launch_pdl(pa_layernorm_rows_kernel, dim3(nseq), dim3(kLnThreads), 0, s, 1, m->ln, (const float*)m->x, (const int*)d_last, m->lnfw, m->lnfb, nseq, C));
    rc = pa_cu_linear(m->ln, C, m->wte, nullptr, m->logits, m->Vp, nseq, V, C, nullptr, 0, 0, path, s);       // logits = lnf . wte^T (:726)
    if (rc != PA_OK) return rc;
    CU_CHECK(pa_launch_pdl(pa_sample_kernel, dim3(nseq), dim3(kSampleThreads), 0, s, 1, (const float*)m->logits, m->Vp, V, (const float*)(coins ? m->d_coins : nullptr), d_next));
    h->launches += launches + 3;        // (pa_qkv_append / pa_decode / pa_prefill count themselves)
    CU_CHECK(cudaMemcpyAsync(h_next, d_next, (size_t)nseq * sizeof(int), cudaMemcpyDeviceToHost, s));
    m->d_next_cur = d_next; m->h_next_cur = h_next; m->nseq_cur = nseq;
    return PA_OK;
}

/* One decode step: one new token per sequence (pa_model_forward with n_new = 1). */
int pa_model_decode_step(pa_model* m, const int* seq_ids, const int* tokens, const float* coins, int nseq, int* next_tokens) {
    if (!m || nseq < 1 || nseq > m->max_batch) { pa_set_error("pa_model_decode_step: bad arguments (nseq=%d, max_batch=%d)", nseq, m ? m->max_batch : 0); return PA_ERR_INVALID; }
    std::vector<int> ones(nseq, 1);
    return pa_model_forward(m, seq_ids, ones.data(), tokens, coins, nseq, next_tokens);
}

}  // extern "C"
