/*
 * pa_layer_fused.cu -- everything of a decode step BETWEEN two paged-attention launches as ONE resident grid
 * (gpt2_forward paged_infer.c:696-728 for a step of at most 128 new tokens):
 *
 *      attproj + residual | layernorm 2 | fc + GELU | fcproj + residual | layernorm 1 of the next layer | QKV + KV append
 *
 * instead of six kernels (four split-K projections of ~10 us each whose tensor pipe is 14 % busy, two layernorms):
 * the projections are latency chains -- launch, first TMA round trip, a few k-slabs of MMAs, split-K publish /
 * arrive / reduce, store -- and most of that is per LAUNCH, not per flop.  Here one grid of one CTA per SM walks the
 * phases; a phase is the tile loop of pa_gemm_tc.cu (tcgen05 kind::tf32 with the 3xTF32 split: A = the activation
 * slab from TMEM as raw and lo columns, B = the weight slab and its lo copy from shared memory, leading term chunked
 * over 256 floats of K, partial tiles of the K splits reduced in split order through an L2 workspace) on the CTA's
 * (tile, split) of that projection, or a layernorm row; the mbarrier rings, the TMEM allocation and the tensor maps
 * live across phases, and the phases meet at a grid barrier (release-add + relaxed polls + acquire fence, ~1.3 us)
 * instead of a kernel boundary.  The attention stays its own launch: it is the HBM-bound half of the step and
 * wants every SM's shared memory for its page ring.
 *
 * MEASURED (B200, GPT-2 124M, 64 sequences x 1024 ctx): correct (logits 2e-6 of the oracle, 29 launches per step instead
 * of 88) but SLOWER than the chain it was meant to replace -- 1.675 vs 1.528 ms per step.  Timeline of CTA 0
 * (PA_FUSED_DEBUG=1, ns: work + barrier): attproj 7000 + 2200, layernorm 3500 + 1600, fc 9700 + 2000, fcproj 9900 + 2100,
 * layernorm 3400 + 1600, QKV 8800.  A projection phase costs what the stand-alone kernel costs: its time is its own
 * latency chain (first TMA round trip ~2 us, k-slabs at ~0.4 us each -- twelve N=64 MMA instructions at ~63 cycles --,
 * split-K publish / poll / reduce ~2.5 us, stores), not the launch, which programmatic dependent launch had already
 * hidden in the chain; the grid barriers and the layernorm phases come on top.  So it is OPT-IN (PA_TUNE_MODEL_PATH=3)
 * and the chain stays the default; what it would take to win is in profiles/r02_model_step.md.
 *
 * Arithmetic = pa_gemm3x_kernel's and pa_layernorm_kernel's, so the logits keep the chain's tolerance (L x 1e-5).
 * All CTAs are co-resident by construction (grid = SMs, one CTA per SM); the launch is cooperative whenever the
 * chain's launch overlap is off, like the workspace split of pa_gemm_tc.cu.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pa_internal.h"
#include "pa_pdl.cuh"
#include "pa_ptx.cuh"
#include "pa_layer_fused.cuh"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {

// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t addr) { return smem_desc(addr, 16, 1024, 2); }

constexpr int kBM = 128;          // rows per CTA = TMEM lanes (the whole step: at most 128 tokens)
constexpr int kBN = 64;           // columns per tile
constexpr int kBK = 32;           // floats per k-slab = one 128-byte swizzle row
constexpr int kChunk = 8;         // k-slabs per main-accumulator chunk (256 floats of K)
constexpr int kStages = 6;        // shared-memory ring: x 16 KB | w 8 KB | w_lo 8 KB per stage
constexpr int kAStages = 4;       // TMEM ring of A slabs
constexpr int kXBytes = kBM * 128, kWBytes = kBN * 128, kStageBytes = kXBytes + 2 * kWBytes;
constexpr int kACols = 2 * kBK;
constexpr int kMain = 0, kSmall = 2 * kBN, kA = 3 * kBN;      // TMEM columns: main x2 | small | A ring
static_assert(3 * kBN + kAStages * kACols <= 512, "TMEM columns");
constexpr size_t kFusedSmem = 1024 + (size_t)kStages * kStageBytes + 256 + kBN * 4 + 64;
constexpr int kMaxTiles = 160;    // arrival counters per phase slot (a projection has at most one tile per SM)

__device__ __forceinline__ float gelu_tanh(float x) {           // gelu_forward, paged_infer.c:243-251
    const float k = 0.7978845608028654f;                        // sqrtf(2/pi)
    const float cube = 0.044715f * x * x * x;
    return 0.5f * x * (1.0f + tanhf(k * (x + cube)));
}
__device__ __forceinline__ float tf32_lo(float a) {
    return a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);
}

// layernorm_forward (paged_infer.c:49-89) of one row by 128 threads, the row held in registers: the arithmetic of
// pa_layernorm_kernel (pa_model.cu)
constexpr int kLnPer = 16;        // C <= 2048
__device__ __forceinline__ void layernorm_row_128(float* __restrict__ o, const float* __restrict__ x, const float* __restrict__ weight,
                                                  const float* __restrict__ bias, int C, float* red /* [8] shared */) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float v[kLnPer];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnPer; ++i) {
        const int c = tid + 128 * i;
        v[i] = c < C ? __ldcg(x + c) : 0.0f;
        sum += v[i];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) red[warp] = sum;
    named_bar_sync(1, 128);
    float tot = 0.0f;
#pragma unroll
    for (int w = 0; w < 4; ++w) tot += red[w];
    const float m = tot / C;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnPer; ++i) {
        const int c = tid + 128 * i;
        const float dlt = v[i] - m;
        if (c < C) var += dlt * dlt;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    if (lane == 0) red[4 + warp] = var;
    named_bar_sync(1, 128);
    float totv = 0.0f;
#pragma unroll
    for (int w = 0; w < 4; ++w) totv += red[4 + w];
    const float s = 1.0f / sqrtf(totv / C + 1e-5f);              // eps, :56
#pragma unroll
    for (int i = 0; i < kLnPer; ++i) {
        const int c = tid + 128 * i;
        if (c < C) o[c] = (s * (v[i] - m)) * weight[c] + bias[c];
    }
    named_bar_sync(1, 128);       // `red` is free for the next row
}

}  // namespace

namespace {

__global__ void __launch_bounds__(192, 1)
pa_layer_fused_kernel(const pa_fused_params P) {
    constexpr uint32_t kIdesc = instr_desc(kBM, kBN, 0, 0);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * kStageBytes);
    uint64_t* full = bars;                       // TMA bytes of a stage landed
    uint64_t* split = bars + kStages;            // A in TMEM and w_lo in shared memory are ready
    uint64_t* empty = bars + 2 * kStages;        // the MMAs that read the stage have completed
    uint64_t* done = bars + 3 * kStages;         // every MMA of the phase has completed
    uint64_t* chunk_done = done + 1;             // [2] the chunk in main accumulator b is complete
    uint64_t* chunk_free = done + 3;             // [2] the splitter threads have taken it into registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 5);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);      // [kBN] the tile's bias
    float* ln_red = bias_s + kBN;                                  // [8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (P.dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); P.dbg[n_stamp++] = t; }
    };
    stamp();
    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&split[s]), 128);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(done), 1);
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&chunk_done[b]), 1); mbar_init(smem_u32(&chunk_free[b]), 128); }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the activations (atty from the attention kernel, x from the previous resident grid) belong to the previous
    // kernel of the stream until it has completed
    pdl_wait();
    stamp();

    // cumulative counters: the rings and the chunk accumulators simply keep going across phases, so every role
    // derives ring slot and mbarrier parity from the same running totals
    int gs = 0;          // k-slabs this CTA has put through its stage ring so far
    int gch = 0;         // main-accumulator chunks so far
    int gdone = 0;       // projection phases this CTA took part in
    unsigned passed = 0; // grid barriers of this launch

    for (int pi = 0; pi < P.n_phases; ++pi) {
        const pa_fused_phase& ph = P.ph[pi];
        if (ph.kind == 1) {
            // ================================ layernorm rows ==================================
            if (warp < 4) {
                for (int row = blockIdx.x; row < P.M; row += gridDim.x)
                    layernorm_row_128(ph.ln_out + (size_t)row * P.C, ph.ln_in + (size_t)row * P.C, ph.ln_w, ph.ln_b, P.C, ln_red);
            }
        } else if ((int)blockIdx.x < ph.tiles_n * ph.n_split) {
            // ================================ projection tile ==================================
            const int tile = blockIdx.x % ph.tiles_n, krank = blockIdx.x / ph.tiles_n, n_split = ph.n_split;
            const int n0 = tile * kBN;
            const int total_slabs = (ph.K + kBK - 1) / kBK;
            const int slab0 = (int)((long long)krank * total_slabs / n_split);
            const int n_slabs = (int)((long long)(krank + 1) * total_slabs / n_split) - slab0;     // >= 1
            const int n_chunks = (n_slabs + kChunk - 1) / kChunk;
            if (tid < kBN) bias_s[tid] = (ph.bias && n0 + tid < ph.N) ? __ldg(ph.bias + n0 + tid) : 0.0f;
            const bool kv_phase = ph.n_dense < ph.N;
            const int my_slot = (kv_phase && warp < 4 && tid < P.M) ? __ldg(P.slots + tid) : 0;

            if (warp == 4) {
                // ---------------------------- TMA producer ----------------------------
                const bool leader = elect_one();
                for (int s = 0; s < n_slabs; ++s) {
                    const int g = gs + s, st = g % kStages;
                    if (g >= kStages) mbar_wait(smem_u32(&empty[st]), ((g / kStages) - 1) & 1);
                    unsigned char* stage = base + st * kStageBytes;
                    const uint32_t bar = smem_u32(&full[st]);
                    if (leader) {
                        mbar_arrive_expect_tx(bar, kXBytes + kWBytes);
                        tma_box_2d(smem_u32(stage + kXBytes), ph.tm_w, (slab0 + s) * kBK, n0, bar);
                        tma_box_2d(smem_u32(stage), ph.tm_x, (slab0 + s) * kBK, 0, bar);          // rows past the matrix read as zero
                    }
                    __syncwarp();
                }
            } else if (warp == 5) {
                // ----------------------------- MMA issuer -----------------------------
                const bool leader = elect_one();
                for (int s = 0; s < n_slabs; ++s) {
                    const int g = gs + s, st = g % kStages;
                    const int c = gch + s / kChunk, cb = c & 1;
                    const bool chunk_start = (s % kChunk) == 0;
                    // main accumulator cb is reused every second chunk: the splitter threads must have read it
                    if (chunk_start && c >= 2) mbar_wait(smem_u32(&chunk_free[cb]), ((c >> 1) - 1) & 1);
                    mbar_wait(smem_u32(&split[st]), (g / kStages) & 1);
                    tc_fence_after();
                    const uint32_t w_lo32 = smem_desc_lo(smem_u32(base + st * kStageBytes + kXBytes), 16);      // + 2 per k-step (pa_ptx.cuh)
                    const uint32_t wlo_lo32 = w_lo32 + (kWBytes >> 4);
                    constexpr uint32_t w_hi32 = smem_desc_hi(1024, 2);
                    const uint32_t a_raw = tmem_base + kA + (g % kAStages) * kACols;
                    const uint32_t a_lo = a_raw + kBK;
                    const uint32_t d_main = tmem_base + kMain + cb * kBN;
                    const uint32_t d_small = tmem_base + kSmall;
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < kBK / 8; ++ks) {
                            mma_tf32_ts_lohi(d_main, a_raw + ks * 8, w_lo32 + ks * 2, w_hi32, kIdesc, (chunk_start && ks == 0) ? 0u : 1u);
                            mma_tf32_ts_lohi(d_small, a_lo + ks * 8, w_lo32 + ks * 2, w_hi32, kIdesc, (s > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts_lohi(d_small, a_raw + ks * 8, wlo_lo32 + ks * 2, w_hi32, kIdesc, 1u);
                        }
                        tc_commit(smem_u32(&empty[st]));
                        if ((s % kChunk) == kChunk - 1 || s == n_slabs - 1) tc_commit(smem_u32(&chunk_done[cb]));
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(smem_u32(done));
                __syncwarp();
            } else {
                // ----------------------- splitter, then epilogue ----------------------
                float acc[kBN];
                const int r = tid;                                   // row of the step = TMEM lane
                const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll
                for (int i = 0; i < kBN; ++i) acc[i] = 0.0f;
                int next_chunk = 0;
                auto take_chunk = [&](int ch) {                  // main accumulator of chunk ch -> registers (round-to-nearest adds)
                    const int c = gch + ch, cb = c & 1;
                    mbar_wait(smem_u32(&chunk_done[cb]), (c >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < kBN; cc += 32) {
                        float v[32];
                        tmem_ld32(tmem_base + lane_off + kMain + cb * kBN + cc, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[cc + i] += v[i];
                    }
                    tc_fence_before();
                    mbar_arrive(smem_u32(&chunk_free[cb]));
                };
                for (int s = 0; s < n_slabs; ++s) {
                    const int g = gs + s, st = g % kStages;
                    if ((s % kChunk) == kChunk / 2 && s >= kChunk) take_chunk(next_chunk++);
                    // the A ring in TMEM is shorter than the stage ring: slab g reuses the columns of slab g - kAStages,
                    // whose MMAs signal the `empty` barrier of the stage that slab used
                    if (g >= kAStages) {
                        const int gp = g - kAStages;
                        mbar_wait(smem_u32(&empty[gp % kStages]), (gp / kStages) & 1);
                    }
                    mbar_wait(smem_u32(&full[st]), (g / kStages) & 1);
                    tc_fence_after();
                    const unsigned char* xs = base + st * kStageBytes;
                    float a[kBK];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {                // row r of the swizzled slab: chunk c sits at c ^ (r % 8)
                        const float4 v = *reinterpret_cast<const float4*>(xs + r * 128 + ((c ^ (r & 7)) << 4));
                        a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
                    }
                    const uint32_t a_tmem = tmem_base + lane_off + kA + (g % kAStages) * kACols;
                    tmem_st32(a_tmem, a);
#pragma unroll
                    for (int i = 0; i < kBK; ++i) a[i] = tf32_lo(a[i]);
                    tmem_st32(a_tmem + kBK, a);
                    // w_lo slab: element-wise over the flat (swizzled) buffer
                    const float4* wsrc = reinterpret_cast<const float4*>(xs + kXBytes);
                    float4* wl = reinterpret_cast<float4*>(const_cast<unsigned char*>(xs) + kXBytes + kWBytes);
#pragma unroll
                    for (int i = 0; i < kWBytes / 16 / 128; ++i) {
                        const float4 v = wsrc[tid + i * 128];
                        wl[tid + i * 128] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                    }
                    fence_proxy_async_smem();      // generic-proxy stores -> visible to the MMA
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(smem_u32(&split[st]));
                }
                while (next_chunk < n_chunks) take_chunk(next_chunk++);      // the last one or two chunks
                mbar_wait(smem_u32(done), gdone & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < kBN; cc += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_off + kSmall + cc, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[cc + i] += v[i];
                }
                tc_fence_before();

                // ---- epilogue: bias, activation, residual, then the dense row or the token's page slot -------
                const int m = r;
                const size_t slot_off = (size_t)my_slot * P.C;
                auto emit4 = [&](int c4, float4 av, float4 rv) {
                    const int n = n0 + 4 * c4;
                    if (n >= ph.N) return;
                    float o[4] = {av.x, av.y, av.z, av.w};
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * c4);
                    o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
                    if (ph.act == 1) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = gelu_tanh(o[e]);
                    }
                    float* dst;
                    if (n < ph.n_dense) dst = ph.out + (size_t)m * ph.out_stride + n;
                    else if (n - ph.n_dense < P.C) dst = P.pool_k + slot_off + (n - ph.n_dense);
                    else dst = P.pool_v + slot_off + (n - ph.n_dense - P.C);
                    if (ph.residual) { o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w; }
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                };
                const int rows = P.M < kBM ? P.M : kBM;
                // split-K through the L2 workspace: publish the partial rows, wait until all n_split partials of the
                // tile are there, then reduce and store THIS CTA's share of the tile's float4 columns in split order
                float4* part = P.ws + ((size_t)tile * n_split + krank) * (kBN / 4) * kBM;
                if (n_split > 1 && r < rows) {
#pragma unroll
                    for (int c4 = 0; c4 < kBN / 4; ++c4)
                        __stcg(part + c4 * kBM + r, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
                }
                const int c4b = krank * (kBN / 4) / n_split, nsh = (krank + 1) * (kBN / 4) / n_split - c4b;      // this CTA's share
                float4 resv[kBN / 8];
                if (ph.residual && r < rows) {
#pragma unroll
                    for (int i = 0; i < kBN / 8; ++i) {
                        const int n = n0 + 4 * (c4b + i);
                        if (i < nsh && n + 3 < ph.N) resv[i] = __ldcg(reinterpret_cast<const float4*>(ph.residual + (size_t)m * ph.res_stride + n));
                    }
                }
                named_bar_sync(1, 128);          // the warpgroup's stores (and bias_s) before thread 0's release
                if (n_split > 1) {
                    unsigned* cnt = P.ws_cnt + ph.slot * kMaxTiles + tile;
                    if (tid == 0) {
                        const unsigned target = (ph.gen + 1) * (unsigned)n_split;      // the counter only grows: n_split per use
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
                        unsigned seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                        } while ((int)(seen - target) < 0);
                    }
                    named_bar_sync(1, 128);
                    float4* stage4 = reinterpret_cast<float4*>(base);       // the stage ring is idle now
                    const int per_split = nsh * kBM;
                    if (r < rows) {
                        for (int sp = 0; sp < n_split; ++sp) {
                            const float4* src = P.ws + (((size_t)tile * n_split + sp) * (kBN / 4) + c4b) * kBM + r;
                            for (int i = 0; i < nsh; ++i) cp_async16(smem_u32(stage4 + sp * per_split + i * kBM + r), src + i * kBM, 16);
                        }
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                    named_bar_sync(1, 128);
                    if (r < rows) {
                        for (int i = 0; i < nsh; ++i) {
                            float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            for (int sp = 0; sp < n_split; ++sp) {                        // split order: deterministic
                                const float4 v = stage4[sp * per_split + i * kBM + r];
                                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                            }
                            float4 rv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                            for (int u = 0; u < kBN / 8; ++u) if (u == i) rv = resv[u];
                            emit4(c4b + i, sum, rv);
                        }
                    }
                } else if (r < rows) {
#pragma unroll
                    for (int c4 = 0; c4 < kBN / 4; ++c4) {
                        const int n = n0 + 4 * c4;
                        const float4 rv = (ph.residual && n + 3 < ph.N) ? __ldcg(reinterpret_cast<const float4*>(ph.residual + (size_t)m * ph.res_stride + n))
                                                                         : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        emit4(c4, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]), rv);
                    }
                }
            }
            gs += n_slabs;
            gch += n_chunks;
            gdone += 1;
        }
        // ================================== the phases meet ===================================
        __syncthreads();
        stamp();
        if (pi + 1 < P.n_phases) {
            // this phase's global stores (generic proxy) are read by the next phase's TMA loads (async proxy) in
            // OTHER CTAs: proxy fence, then the release / acquire pair of the barrier
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.bar) : "memory");
                const unsigned target = (P.bar_gen + passed + 1) * gridDim.x;
                unsigned seen;
                do {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.bar) : "memory");
                } while ((int)(seen - target) < 0);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncthreads();
            ++passed;
            stamp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

extern "C" {

size_t pa_cu_layer_fused_smem(void) { return kFusedSmem; }
int pa_cu_layer_fused_max_tiles(void) { return kMaxTiles; }

/* K splits of a projection inside the resident grid: the same rule as pa_cu_gemm_tc (cover the machine, at least two
 * k-slabs per CTA, at most 16 = float4 columns of a 64-wide tile), but from (N, K, SMs) alone -- never from M -- so
 * that the never-reset arrival counters advance by the same amount every launch. */
int pa_cu_layer_fused_split(int N, int K, int sms, int* tiles_n) {
    const int tiles = (N + kBN - 1) / kBN;
    const int total_slabs = (K + kBK - 1) / kBK;
    int n_split = tiles <= sms ? sms / tiles : 1;
    if (n_split > 16) n_split = 16;
    if (n_split > total_slabs / 2) n_split = total_slabs / 2;
    if (n_split < 1) n_split = 1;
    if (tiles_n) *tiles_n = tiles;
    return n_split;
}

int pa_cu_layer_fused_launch(const pa_fused_params* p, int sms, int cooperative, void* stream) {
    static std::atomic<unsigned long long> attr_done{0};
    CU_CHECK(pa_optin_smem(attr_done, pa_layer_fused_kernel, (int)kFusedSmem));
    static const bool dbg = getenv("PA_FUSED_DEBUG") != nullptr;
    static unsigned long long* d_dbg = nullptr;
    pa_fused_params q = *p;
    if (dbg) {
        if (!d_dbg) CU_CHECK(cudaMalloc((void**)&d_dbg, 64 * sizeof(unsigned long long)));
        q.dbg = d_dbg;
    }
    pa_launch_cooperative = cooperative ? 1 : 0;
    const cudaError_t e = pa_launch_pdl(pa_layer_fused_kernel, dim3(sms), dim3(192), kFusedSmem, (cudaStream_t)stream, 1, q);
    pa_launch_cooperative = 0;
    CU_CHECK(e);
    if (dbg) {      // ns from kernel entry: dependency resolved, then per phase (work done, barrier passed)
        unsigned long long hst[64];
        CU_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        CU_CHECK(cudaMemcpy(hst, d_dbg, sizeof(hst), cudaMemcpyDeviceToHost));
        fprintf(stderr, "fused dbg (%d phases): wait %lld;", p->n_phases, (long long)(hst[1] - hst[0]));
        for (int i = 0; i < p->n_phases; ++i) {
            const long long w = (long long)(hst[2 + 2 * i] - hst[1 + 2 * i]);
            const long long b = i + 1 < p->n_phases ? (long long)(hst[3 + 2 * i] - hst[2 + 2 * i]) : 0;
            fprintf(stderr, " %s[%d x %d] %lld+%lld", p->ph[i].kind ? "ln" : "proj", p->ph[i].tiles_n, p->ph[i].n_split, w, b);
        }
        fprintf(stderr, "; total %lld ns\n", (long long)(hst[2 * p->n_phases] - hst[0]));
    }
    return PA_OK;
}

}  // extern "C"
