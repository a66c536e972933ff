/*
 * pa_gemm_tc.cu -- fp32-accurate tensor-core GEMM for the projections next to the attention path
 * (SURVEY 8f.1/8f.2):   out[m][n] = bias[n] + sum_k x[m][k] * w[n][k]      (paged_infer.c:92-114)
 *
 * tcgen05.mma kind::tf32 with the 3xTF32 split: every fp32 operand a is used as
 * a_hi (what the tensor core sees: the top 19 bits) and a_lo = a - a_hi, and
 *      x.w  ~=  x_hi.w_hi + x_hi.w_lo + x_lo.w_hi          (the dropped x_lo.w_lo term is 2^-22 relative)
 * That keeps the result inside the path's 1e-5 tolerance at a third of the TF32 tensor rate --
 * still an order of magnitude above fp32 FFMA.  One more thing is needed for that: the tensor
 * core's fp32 accumulate TRUNCATES (measured: results biased towards zero by ~3e-8 of the
 * accumulator per MMA, coherently, i.e. 1.3e-5 at K=1600 if everything is summed in one TMEM
 * accumulator).  So the leading term x_hi.w_hi is accumulated in TMEM only over chunks of 256
 * floats of K, and the chunks are added in registers with round-to-nearest fp32 (the partial
 * sums of different chunks have unrelated signs, so the bias no longer adds up); the two small
 * terms go to a second TMEM accumulator whose truncation is 2^-11 further down.
 *
 * One CTA per (128 rows of x, BN columns of out):
 *   warp 4       TMA producer: per 32-float k-slab one box of x (128 rows) and one of w (BN rows),
 *                128-byte swizzle, 4-stage mbarrier ring
 *   warps 0-3    splitter, then epilogue: thread r copies row r of the x slab into TMEM as the A
 *                operand (raw = hi, and lo) and the warpgroup writes the w_lo slab next to the raw w
 *                slab in shared memory (element-wise, so the swizzled layout is irrelevant)
 *   warp 5       MMA issuer: 3 MMAs per 8-deep k-step, A from TMEM, B from shared memory
 * Epilogue: thread r reads row r of the accumulator from TMEM, adds the bias and stores BN
 * contiguous floats -- to the dense output row, or (fused KV append) to the token's page slot.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "pa_internal.h"
#include "pa_pdl.cuh"
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

// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t addr) { return smem_desc(addr, 16, 1024, 2); }

constexpr int kBM = 128;          // rows per CTA = TMEM lanes
constexpr int kBK = 32;           // floats per k-slab = one 128-byte swizzle row
constexpr int kMaxSplit = 4;      // CTAs of a cluster splitting K
constexpr int kChunk = 8;          // k-slabs per main-accumulator chunk (256 floats of K)

struct GemmTcParams {
    const float* bias;       // (N) or NULL
    float* out;              // dense destination, column n < n_dense of row m at out + m*out_stride + n
    float* pool_k;           // destinations of columns >= n_dense (fused KV append), or NULL
    float* pool_v;
    const int* slots;        // [M] slot_mapping
    int M, N, K;
    int out_stride;
    int n_dense;
    int C;
    int terms;               // 3: 3xTF32 (fp32-accurate), 1: plain TF32 (reduced precision)
    const float* residual;   // optional: out = act(acc + bias) + residual[m][n] (may alias out)
    int res_stride;
    int act;                 // 0 none, 1 GELU (tanh form, paged_infer.c:243-251)
    unsigned long long* dbg; // optional timeline of CTA (0,0,0) (PA_GEMM_DEBUG=1), NULL normally
    // split-K through an L2-resident workspace (gridDim.z CTAs per tile, no cluster): partial tiles
    // [tile][split][float4 column][row], and one arrival counter per tile.  Two counter sets alternate
    // between launches: a launch counts in ws_cnt and zeroes ws_cnt_next for its successor (which
    // cannot touch it before this grid has completed), so nobody has to wait for a reset.
    float4* ws;
    unsigned* ws_cnt;
    unsigned* ws_cnt_next;
};

__device__ __forceinline__ float gelu_tanh(float x) {           // gelu_forward, paged_infer.c:243-251
    const float k = 0.7978845608028654f;                        // sqrtf(2/pi)
    const float cube = 0.044715f * x * x * x;
    return 0.5f * x * (1.0f + tanhf(k * (x + cube)));
}
// the part of an fp32 value the TF32 datapath drops
__device__ __forceinline__ float tf32_lo(float a) {
    return a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);
}

template <int BN>
struct GemmCfg {
    static constexpr int kStages = BN == 64 ? 6 : 4;     // shared-memory ring: 32 KB (BN 64) / 48 KB (BN 128) per stage, 192 KB in all
    static constexpr int kAStages = BN == 64 ? 4 : 2;    // TMEM ring of A slabs (what is left of the 512 columns)
    static constexpr int kXBytes = kBM * 128;            // one k-slab of x: 128 rows x 128 B
    static constexpr int kWBytes = BN * 128;
    static constexpr int kStageBytes = kXBytes + 2 * kWBytes;     // x | w | w_lo
    static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 256 + BN * 4;
    static constexpr int kACols = 2 * kBK;               // A operand per slab: 32 raw + 32 lo columns
    // TMEM columns: main accumulator (hi.hi) x2 (chunks alternate) | small-term accumulator | A ring
    static constexpr int kMain = 0, kSmall = 2 * BN, kA = 3 * BN;
    static constexpr int kCols = 3 * BN + kAStages * kACols;
    static constexpr int kTmemCols = kCols <= 256 ? 256 : 512;
    static_assert(kCols <= 512, "TMEM columns");
    // split-K reduction scratch in the leader's (idle) stage buffers: [peer][float4 column][row]
    // (cluster splits are only launched with BN = 64: three peers' tiles fit the leader's ring)
};


// grid (N tiles, M tiles, n_split); the n_split CTAs of a cluster share an output tile and split K.
template <int BN>
__global__ void __launch_bounds__(192, 1)
pa_gemm3x_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const GemmTcParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages, kAStages = Cfg::kAStages;
    constexpr uint32_t kIdesc = instr_desc(kBM, BN, 0, 0);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * Cfg::kStageBytes);
    uint64_t* full = bars;                       // TMA bytes of a stage landed
    uint64_t* split = bars + kStages;            // A in TMEM and w_lo in shared memory are ready
    uint64_t* empty = bars + 2 * kStages;        // the MMAs that read the stage have completed
    uint64_t* done = bars + 3 * kStages;         // every MMA has completed
    uint64_t* chunk_done = done + 1;             // [2] the chunk in main accumulator b is complete
    uint64_t* chunk_free = done + 3;             // [2] the splitter threads have taken it into registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 5);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);      // [BN] this tile's bias, fetched while the pipeline fills

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    auto stamp = [&](int i) { if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[i] = t; } };
    stamp(0);
    pdl_launch_dependents();      // the next kernel of the step may start its own prologue
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
    const int n_split = gridDim.z;
    const int krank = blockIdx.z;                 // = the CTA's rank in its cluster when K is split over a cluster
    const int total_slabs = (p.K + kBK - 1) / kBK;
    const int slab0 = (int)((long long)krank * total_slabs / n_split);
    const int n_slabs = (int)((long long)(krank + 1) * total_slabs / n_split) - slab0;     // >= 1: the host keeps n_split <= total_slabs

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
    if (tid < BN) bias_s[tid] = (p.bias && n0 + tid < p.N) ? __ldg(p.bias + n0 + tid) : 0.0f;
    if (warp == 4) {
        tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    stamp(1);
    // page slot of this thread's row (fused KV append), fetched long before the epilogue needs it
    const int my_slot = (p.slots && warp < 4 && m0 + tid < p.M) ? __ldg(p.slots + m0 + tid) : 0;
    // The weights are never written inside a chain of dependent launches, so the producer requests the
    // w boxes of the first ring pass BEFORE waiting for the previous kernel: they stream in from HBM
    // while that kernel drains.  (`full` expects x + w bytes; the x box follows after the wait.)
    const int n_pre = n_slabs < kStages ? n_slabs : kStages;
    if (warp == 4) {
        if (elect_one()) {
            for (int s = 0; s < n_pre; ++s) {
                const uint32_t bar = smem_u32(&full[s]);
                mbar_arrive_expect_tx(bar, Cfg::kXBytes + Cfg::kWBytes);
                tma_box_2d(smem_u32(base + s * Cfg::kStageBytes + Cfg::kXBytes), &tm_w, (slab0 + s) * kBK, n0, bar);
            }
        }
        __syncwarp();
    }
    // everything above touched only weights, the step tables (mirrored before the chain started) and
    // on-chip state; x (and the buffers written below) belong to the previous kernel until it has completed
    pdl_wait();
    if (p.ws && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid < 128) {
        p.ws_cnt_next[tid] = 0;
        p.ws_cnt_next[tid + 128] = 0;
    }

    if (warp == 4) {
        // ================================ TMA producer ========================================
        const bool leader = elect_one();
        for (int s = 0; s < n_slabs; ++s) {
            const int st = s % kStages, j = s / kStages;
            if (j > 0) mbar_wait(smem_u32(&empty[st]), (j - 1) & 1);
            unsigned char* stage = base + st * Cfg::kStageBytes;
            const uint32_t bar = smem_u32(&full[st]);
            if (leader) {
                if (s >= n_pre) {
                    mbar_arrive_expect_tx(bar, Cfg::kXBytes + Cfg::kWBytes);
                    tma_box_2d(smem_u32(stage + Cfg::kXBytes), &tm_w, (slab0 + s) * kBK, n0, bar);
                }
                tma_box_2d(smem_u32(stage), &tm_x, (slab0 + s) * kBK, m0, bar);          // rows/columns past the matrix read as zero
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ================================= MMA issuer =========================================
        const bool leader = elect_one();
        for (int s = 0; s < n_slabs; ++s) {
            const int st = s % kStages, j = s / kStages;
            const int ch = s / kChunk, cb = ch & 1;
            const bool chunk_start = (s % kChunk) == 0;
            // main accumulator cb is reused every second chunk: the splitter threads must have read it
            if (chunk_start && ch >= 2) { mbar_wait(smem_u32(&chunk_free[cb]), ((ch >> 1) - 1) & 1); }
            mbar_wait(smem_u32(&split[st]), j & 1);
            tc_fence_after();
            // descriptor low words (address field + LBO) of the w and w_lo slabs, advanced by 2 (32 bytes) per k-step;
            // the high word is constant: one uniform add per instruction instead of smem_desc()'s shift / mask / or chain
            const uint32_t w_lo32 = smem_desc_lo(smem_u32(base + st * Cfg::kStageBytes + Cfg::kXBytes), 16);
            const uint32_t wlo_lo32 = w_lo32 + (Cfg::kWBytes >> 4);
            constexpr uint32_t w_hi32 = smem_desc_hi(1024, 2);
            const uint32_t a_raw = tmem_base + Cfg::kA + (s % kAStages) * Cfg::kACols;
            const uint32_t a_lo = a_raw + kBK;
            const uint32_t d_main = tmem_base + Cfg::kMain + cb * BN;
            const uint32_t d_small = tmem_base + Cfg::kSmall;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < kBK / 8; ++ks) {
                    mma_tf32_ts_lohi(d_main, a_raw + ks * 8, w_lo32 + ks * 2, w_hi32, kIdesc, (chunk_start && ks == 0) ? 0u : 1u);
                    if (p.terms == 3) {
                        mma_tf32_ts_lohi(d_small, a_lo + ks * 8, w_lo32 + ks * 2, w_hi32, kIdesc, (s > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ts_lohi(d_small, a_raw + ks * 8, wlo_lo32 + ks * 2, w_hi32, kIdesc, 1u);
                    }
                }
                tc_commit(smem_u32(&empty[st]));
                if ((s % kChunk) == kChunk - 1 || s == n_slabs - 1) tc_commit(smem_u32(&chunk_done[cb]));
            }
            __syncwarp();
        }
        if (leader) tc_commit(smem_u32(done));
        __syncwarp();
    }

    float acc[BN];                                       // splitter threads: this CTA's share of the output row
    const int r = tid;                                   // row of the tile = TMEM lane (threads 0..127)
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    if (warp < 4) {
        // ============================ splitter, then epilogue ==================================
#pragma unroll
        for (int i = 0; i < BN; ++i) acc[i] = 0.0f;
        const int n_chunks = (n_slabs + kChunk - 1) / kChunk;
        int next_chunk = 0;                              // chunks are taken in order
        auto take_chunk = [&](int ch) {                  // main accumulator of chunk ch -> registers (round-to-nearest adds)
            const int cb = ch & 1;
            mbar_wait(smem_u32(&chunk_done[cb]), (ch >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN; c += 32) {
                float v[32];
                tmem_ld32(tmem_base + lane_off + Cfg::kMain + cb * BN + c, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c + i] += v[i];
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&chunk_free[cb]));
        };
        for (int s = 0; s < n_slabs; ++s) {
            const int st = s % kStages, j = s / kStages;
            // half a chunk later the MMAs of the previous chunk have long completed: no stall
            if ((s % kChunk) == kChunk / 2 && s >= kChunk) take_chunk(next_chunk++);
            // the A ring in TMEM is shorter than the stage ring: slab s reuses the columns of slab
            // s - kAStages, whose MMAs signal the `empty` barrier of the stage that slab used
            if (s >= kAStages) {
                const int sp = s - kAStages;
                mbar_wait(smem_u32(&empty[sp % kStages]), (sp / kStages) & 1);
            }
            mbar_wait(smem_u32(&full[st]), j & 1);
            if (s == 0) stamp(2);
            tc_fence_after();
            const unsigned char* xs = base + st * Cfg::kStageBytes;
            float a[kBK];
#pragma unroll
            for (int c = 0; c < 8; ++c) {                // row r of the swizzled slab: chunk c sits at c ^ (r % 8)
                const float4 v = *reinterpret_cast<const float4*>(xs + r * 128 + ((c ^ (r & 7)) << 4));
                a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
            }
            const uint32_t a_tmem = tmem_base + lane_off + Cfg::kA + (s % kAStages) * Cfg::kACols;
            tmem_st32(a_tmem, a);
            if (p.terms == 3) {
#pragma unroll
                for (int i = 0; i < kBK; ++i) a[i] = tf32_lo(a[i]);
                tmem_st32(a_tmem + kBK, a);
                // w_lo slab: element-wise over the flat (swizzled) buffer
                const float4* ws = reinterpret_cast<const float4*>(xs + Cfg::kXBytes);
                float4* wl = reinterpret_cast<float4*>(const_cast<unsigned char*>(xs) + Cfg::kXBytes + Cfg::kWBytes);
#pragma unroll
                for (int i = 0; i < Cfg::kWBytes / 16 / 128; ++i) {
                    const float4 v = ws[tid + i * 128];
                    wl[tid + i * 128] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                }
                fence_proxy_async_smem();      // generic-proxy stores -> visible to the MMA
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&split[st]));
        }
        stamp(3);
        while (next_chunk < n_chunks) take_chunk(next_chunk++);      // the last one or two chunks
        mbar_wait(smem_u32(done), 0);
        stamp(4);
        tc_fence_after();
        if (p.terms == 3) {
#pragma unroll
            for (int c = 0; c < BN; c += 32) {
                float v[32];
                tmem_ld32(tmem_base + lane_off + Cfg::kSmall + c, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c + i] += v[i];
            }
        }
        tc_fence_before();
    }

    // ---- epilogue for 4 columns of row r: bias, activation, residual, then the dense row or the
    // token's page slot (a float4 lies entirely in one destination: the Q | K | V boundaries are
    // multiples of 32 whenever this kernel is chosen) ----------------------------------------------
    const int m = m0 + r;
    const size_t slot_off = (size_t)my_slot * p.C;
    // rv: the residual values of a full group, fetched by the caller (ahead of time where it can)
    auto emit4 = [&](int c4, float4 a, float4 rv) {
        const int n = n0 + 4 * c4;
        if (n >= p.N) return;
        float o[4] = {a.x, a.y, a.z, a.w};
        const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * c4);
        o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
        if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = gelu_tanh(o[e]);
        }
        float* dst;
        if (n < p.n_dense) dst = p.out + (size_t)m * p.out_stride + n;
        else if (n - p.n_dense < p.C) dst = p.pool_k + slot_off + (n - p.n_dense);
        else dst = p.pool_v + slot_off + (n - p.n_dense - p.C);
        if (n + 3 < p.N) {
            if (p.residual) { o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w; }
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {                                         // last, partial group of a row whose length is not a multiple of 4
            const float* res = p.residual ? p.residual + (size_t)m * p.res_stride + n : nullptr;
            for (int e = 0; e < 4 && n + e < p.N; ++e) dst[e] = o[e] + (res ? res[e] : 0.0f);
        }
    };

    __syncwarp();
    if (n_split > 1 && p.ws) {
        // ---- split-K through the L2 workspace: every CTA of the tile publishes its partial rows,
        // waits until all n_split partials are there (the grid is co-resident: the host keeps
        // tiles * n_split <= SMs), then reduces and stores ITS share of the tile's float4 columns --
        // the sum runs in split order, so the result does not depend on arrival order -------------
        if (warp < 4) {
            const int tile = blockIdx.y * gridDim.x + blockIdx.x;
            const int rows = (p.M - m0) < kBM ? (p.M - m0) : kBM;             // valid rows of this tile
            float4* part = p.ws + ((size_t)tile * n_split + krank) * (BN / 4) * kBM;
            if (r < rows) {
#pragma unroll
                for (int c4 = 0; c4 < BN / 4; ++c4)
                    __stcg(part + c4 * kBM + r, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
            }
            // this CTA's share of the tile's float4 columns, and the residual values it will need
            const int c4b = krank * (BN / 4) / n_split, nsh = (krank + 1) * (BN / 4) / n_split - c4b;      // 1..8 columns
            float4 resv[BN / 8];
            if (p.residual && r < rows) {
#pragma unroll
                for (int i = 0; i < BN / 8; ++i) {
                    const int n = n0 + 4 * (c4b + i);
                    if (i < nsh && n + 3 < p.N) resv[i] = __ldcg(reinterpret_cast<const float4*>(p.residual + (size_t)m * p.res_stride + n));
                }
            }
            // publish: the barrier orders the warpgroup's stores before thread 0's release
            named_bar_sync(1, 128);
            unsigned* cnt = p.ws_cnt + tile;
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                } while (seen < (unsigned)n_split);
            }
            named_bar_sync(1, 128);
            stamp(5);
            // all n_split partials of the share -> shared memory (the stage ring is idle now) in ONE
            // round trip: 16-byte async copies (L2 only), [split][column][row]
            float4* stage4 = reinterpret_cast<float4*>(base);
            const int per_split = nsh * kBM;
            if (r < rows) {                                                   // thread r moves row r of every (split, column)
                for (int sp = 0; sp < n_split; ++sp) {
                    const float4* src = p.ws + (((size_t)tile * n_split + sp) * (BN / 4) + c4b) * kBM + r;
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
                    for (int u = 0; u < BN / 8; ++u) if (u == i) rv = resv[u];
                    emit4(c4b + i, sum, rv);
                }
            }
        }
    } else {
        // ---- split-K over a cluster: the peers hand their partial rows to the leader through
        // distributed shared memory; the leader adds them in rank order (deterministic) -----------
        if (n_split > 1) {
            cluster_sync_all();                          // the leader's stage buffers are idle from here on
            if (warp < 4 && krank > 0) {
                const uint32_t dst = mapa(smem_u32(base + (size_t)(krank - 1) * kBM * BN * 4), 0);
#pragma unroll
                for (int c4 = 0; c4 < BN / 4; ++c4)
                    st_cluster_v4(dst + (uint32_t)(c4 * kBM + r) * 16u, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
            }
            cluster_sync_all();
            if (warp < 4 && krank == 0) {
                for (int pr = 0; pr < n_split - 1; ++pr) {
                    const float4* src = reinterpret_cast<const float4*>(base + (size_t)pr * kBM * BN * 4);
#pragma unroll
                    for (int c4 = 0; c4 < BN / 4; ++c4) {
                        const float4 v = src[c4 * kBM + r];
                        acc[4 * c4] += v.x; acc[4 * c4 + 1] += v.y; acc[4 * c4 + 2] += v.z; acc[4 * c4 + 3] += v.w;
                    }
                }
            }
        }
        stamp(5);
        if (warp < 4 && krank == 0 && m < p.M) {
#pragma unroll
            for (int g = 0; g < BN / 4; g += 8) {        // 8 residual loads in flight, then 8 stores
                float4 rv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n = n0 + 4 * (g + i);
                    rv[i] = (p.residual && n + 3 < p.N) ? __ldcg(reinterpret_cast<const float4*>(p.residual + (size_t)m * p.res_stride + n))
                                                        : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    emit4(g + i, make_float4(acc[4 * (g + i)], acc[4 * (g + i) + 1], acc[4 * (g + i) + 2], acc[4 * (g + i) + 3]), rv[i]);
            }
        }
    }
    stamp(6);
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---- host side ---------------------------------------------------------------------------------
// (rows, K) fp32 matrix with a row stride; box = 32 columns x box_rows rows, 128-byte swizzle, zero fill outside
int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int row_stride, int box_rows) {
    pa_encode_tiled_fn enc = pa_get_encode_tiled();
    if (!enc) { pa_set_error("cuTensorMapEncodeTiled not available from the driver"); return PA_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pa_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return PA_ERR_CUDA; }
    return PA_OK;
}

// small cache of encoded maps: the same weight matrices and activation buffers come back every step
struct MapKey { const void* ptr; int rows, K, stride, box_rows; };
struct MapCache {
    static constexpr int N = 256;      // a 12-layer model alone has 49 weight matrices x up to two tile widths
    MapKey key[N];
    CUtensorMap map[N];
    int used = 0, next = 0;
    std::mutex mu;
};
MapCache g_maps;
int get_map(CUtensorMap* out, const float* ptr, int rows, int K, int row_stride, int box_rows) {
    std::lock_guard<std::mutex> lock(g_maps.mu);
    for (int i = 0; i < g_maps.used; ++i) {
        const MapKey& k = g_maps.key[i];
        if (k.ptr == ptr && k.rows == rows && k.K == K && k.stride == row_stride && k.box_rows == box_rows) {
            *out = g_maps.map[i];
            return PA_OK;
        }
    }
    int rc = make_map(out, ptr, rows, K, row_stride, box_rows);
    if (rc != PA_OK) return rc;
    const int i = g_maps.used < MapCache::N ? g_maps.used++ : (g_maps.next++ % MapCache::N);
    g_maps.key[i] = MapKey{ptr, rows, K, row_stride, box_rows};
    g_maps.map[i] = *out;
    return PA_OK;
}

// ---- split-K workspace: one per (device, stream) -- launches on a stream are serialised (also under
// programmatic dependent launch: the reduction runs after griddepcontrol.wait), so they can share it.
constexpr int kWsMaxSplit = 16;       // <= BN / 4 float4 columns of a tile: every CTA of a tile reduces at least one
constexpr int kWsMaxCtas = 256;       // partial tiles in the workspace (>= SMs of the device)
struct SplitWs { int dev; cudaStream_t stream; float4* ws; unsigned* cnt; int parity; };
static_assert(kWsMaxCtas == 256, "the kernel zeroes the next counter set with 128 threads x 2");
struct SplitWsCache {
    static constexpr int N = 16;
    SplitWs e[N];
    int used = 0;
    std::mutex mu;
};
SplitWsCache g_ws;
// false when the cache is full or the allocation fails (the caller then splits over a cluster instead);
// hands out the counter set of this launch and flips to the other one for the next
bool get_split_ws(int dev, cudaStream_t s, GemmTcParams* p, bool* cooperative) {
    constexpr int BN = 128;        // sized for the widest tile
    std::lock_guard<std::mutex> lock(g_ws.mu);
    SplitWs* w = nullptr;
    for (int i = 0; i < g_ws.used && !w; ++i)
        if (g_ws.e[i].dev == dev && g_ws.e[i].stream == s) w = &g_ws.e[i];
    if (!w) {
        if (g_ws.used == SplitWsCache::N) return false;
        SplitWs n{dev, s, nullptr, nullptr, 0};
        const size_t bytes = (size_t)kWsMaxCtas * kBM * BN * sizeof(float);
        if (cudaMalloc((void**)&n.ws, bytes) != cudaSuccess || cudaMalloc((void**)&n.cnt, 2 * kWsMaxCtas * sizeof(unsigned)) != cudaSuccess ||
            cudaMemsetAsync(n.cnt, 0, 2 * kWsMaxCtas * sizeof(unsigned), s) != cudaSuccess) {
            cudaGetLastError();
            if (n.ws) cudaFree(n.ws);
            if (n.cnt) cudaFree(n.cnt);
            return false;
        }
        g_ws.e[g_ws.used] = n;
        w = &g_ws.e[g_ws.used++];
    }
    p->ws = w->ws;
    p->ws_cnt = w->cnt + w->parity * kWsMaxCtas;
    p->ws_cnt_next = w->cnt + (1 - w->parity) * kWsMaxCtas;      // (the sets swap in split_ws_launched, once the launch went through)
    // The CTAs of a workspace split wait for each other, which is safe while the grid has the device to itself
    // (one stream of projections per device -- the handle's).  Once a second stream of the same device has run
    // such launches, two grids could starve each other of SMs: from then on the launches are cooperative (the
    // driver makes each grid resident as a whole), which costs the launch overlap (measured +8 % per step).
    int others = 0;
    for (int i = 0; i < g_ws.used; ++i) others += g_ws.e[i].dev == dev && g_ws.e[i].stream != s;
    static const char* coop_env = getenv("PA_GEMM_COOP");
    *cooperative = coop_env ? atoi(coop_env) != 0 : others > 0;
    return true;
}

}  // namespace
/* the handle that owned `stream` is going away: its split-K workspace with it */
extern "C" void pa_cu_gemm_stream_released(int dev, void* stream) {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    for (int i = 0; i < g_ws.used; ++i) {
        if (g_ws.e[i].dev == dev && g_ws.e[i].stream == (cudaStream_t)stream) {
            cudaFree(g_ws.e[i].ws);
            cudaFree(g_ws.e[i].cnt);
            g_ws.e[i] = g_ws.e[--g_ws.used];
            return;
        }
    }
}
namespace {

// the launch that was handed the current counter set is in the stream: the next one takes the other set
void split_ws_launched(int dev, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    for (int i = 0; i < g_ws.used; ++i)
        if (g_ws.e[i].dev == dev && g_ws.e[i].stream == s) { g_ws.e[i].parity ^= 1; return; }
}

template <int BN>
int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tw, const GemmTcParams& p, int n_split, bool cluster, bool cooperative, cudaStream_t s) {
    using Cfg = GemmCfg<BN>;
    auto fn = pa_gemm3x_kernel<BN>;
    static std::atomic<unsigned long long> attr_done{0};       // per instantiation; one bit per device
    CU_CHECK(pa_optin_smem(attr_done, fn, (int)Cfg::kSmem));
    pa_launch_cooperative = cooperative ? 1 : 0;
    const cudaError_t e = pa_launch_pdl(fn, dim3((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, n_split), dim3(192), Cfg::kSmem, s, cluster ? n_split : 1, tx, tw, p);
    pa_launch_cooperative = 0;
    CU_CHECK(e);
    return PA_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

/* (rows, K) fp32 matrix -> tensor map with a 32-column x box_rows box, 128-byte swizzle (for pa_layer_fused.cu) */
extern "C" int pa_cu_make_map_2d(void* map_out, const float* ptr, int rows, int K, int row_stride, int box_rows) {
    return make_map(reinterpret_cast<CUtensorMap*>(map_out), ptr, rows, K, row_stride, box_rows);
}

/* out (M,N) = x (M,K) . w (N,K)^T + bias on the tensor cores; columns >= n_dense go to the page
 * slots (fused KV append) when pool_k is given.  terms = 3 (3xTF32, fp32-accurate) or 1 (TF32).
 * PA_ERR_UNSUPPORTED when the shape is outside the kernel's domain (caller falls back to SIMT). */
extern "C" int pa_cu_gemm_tc(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                             int M, int N, int K, int n_dense, float* pool_k, float* pool_v, const int* slots, int C,
                             int terms, int n_split_override, const float* residual, int res_stride, int act, void* stream) {
    if (M <= 0 || N <= 0) return PA_OK;
    if ((K & 3) || (x_stride & 3) || (out_stride & 3) || !aligned16(x) || !aligned16(w) || !aligned16(out) ||
        (residual && ((res_stride & 3) || !aligned16(residual))))
        return PA_ERR_UNSUPPORTED;
    if (pool_k && (N & 3)) return PA_ERR_UNSUPPORTED;
    // a 32-column group of the epilogue must not straddle the Q | K | V boundaries
    if (pool_k && ((n_dense & 31) || (C & 31))) return PA_ERR_UNSUPPORTED;
    GemmTcParams p;
    p.bias = bias; p.out = out; p.pool_k = pool_k; p.pool_v = pool_v; p.slots = slots;
    p.M = M; p.N = N; p.K = K; p.out_stride = out_stride; p.n_dense = pool_k ? n_dense : N; p.C = C;
    p.terms = terms == 1 ? 1 : 3;
    p.residual = residual; p.res_stride = res_stride; p.act = act;
    static unsigned long long* d_dbg = nullptr;
    p.dbg = nullptr;
    if (getenv("PA_GEMM_DEBUG")) { if (!d_dbg) cudaMalloc((void**)&d_dbg, 64); p.dbg = d_dbg; }
    // Small M (decode): a handful of CTAs each walking all of K is latency-bound (a CTA turns a
    // k-slab around in ~0.4 us at BN = 64 / ~0.5 us at BN = 128, bounded by its shared-memory
    // traffic), so K is split until the grid covers the machine: n_split CTAs per tile, each with at
    // least two k-slabs, all co-resident (tiles * n_split <= SMs: they wait for each other), partials
    // reduced through the L2 workspace.  Tiles are 64 columns wide while that leaves room to split;
    // when 64-wide tiles alone (nearly) fill the machine, 128-wide tiles halve their number -- less
    // shared-memory traffic per flop, and K can be split again.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > kWsMaxCtas) sms = kWsMaxCtas;
    const int total_slabs = (K + kBK - 1) / kBK;
    const long long m_tiles = (M + kBM - 1) / kBM;
    static const int bn_env = getenv("PA_GEMM_BN") ? atoi(getenv("PA_GEMM_BN")) : 0;      // experiments: force 64 or 128
    int BN = 64;
    if (n_split_override >= 0 && N >= 128) {
        const long long tiles64 = (long long)((N + 63) / 64) * m_tiles;
        if (tiles64 * 2 > sms) BN = 128;
        if (bn_env == 64 || bn_env == 128) BN = bn_env;
    }
    const long long tiles = (long long)((N + BN - 1) / BN) * m_tiles;
    const int ws_cap = tiles <= sms ? (int)(sms / tiles) : 1;          // co-residency bound
    int n_split = 1;
    bool cluster = false, cooperative = false;
    if (n_split_override == 0) {
        n_split = ws_cap < kWsMaxSplit ? ws_cap : kWsMaxSplit;
        if (n_split > total_slabs / 2) n_split = total_slabs / 2;
        if (n_split < 1) n_split = 1;
    } else if (n_split_override > 0) {                                  // forced split count (clamped to what can run)
        n_split = n_split_override < kWsMaxSplit ? n_split_override : kWsMaxSplit;
        if (n_split > ws_cap) n_split = ws_cap;
        if (n_split > total_slabs) n_split = total_slabs;
    } else {                                                            // negative: K split over a cluster of 2 or 4 CTAs (BN = 64)
        cluster = true;
        n_split = -n_split_override >= kMaxSplit ? kMaxSplit : 2;
        while (n_split > 1 && total_slabs < n_split) n_split /= 2;
    }
    p.ws = nullptr; p.ws_cnt = nullptr; p.ws_cnt_next = nullptr;
    if (n_split > 1 && !cluster) {
        if (!get_split_ws(dev, (cudaStream_t)stream, &p, &cooperative)) {             // no workspace: clusters of 2 or 4
            if (BN == 128) n_split = 1;
            else {
                cluster = true;
                n_split = (n_split >= 4 && tiles * 4 <= (sms / 37) * 32) ? 4 : 2;
                while (n_split > 1 && (total_slabs < n_split || tiles * n_split > sms)) n_split /= 2;
            }
        }
    }
    if (n_split == 1) cluster = false;
    CUtensorMap tx, tw;
    int rc = get_map(&tx, x, M, K, x_stride, kBM);
    if (rc == PA_OK) rc = get_map(&tw, w, N, K, K, BN);
    if (rc != PA_OK) return rc;
    cooperative = cooperative && n_split > 1 && !cluster;
    rc = BN == 128 ? launch_gemm<128>(tx, tw, p, n_split, false, cooperative, (cudaStream_t)stream)
                   : launch_gemm<64>(tx, tw, p, n_split, cluster, cooperative, (cudaStream_t)stream);
    if (rc == PA_OK && p.ws) split_ws_launched(dev, (cudaStream_t)stream);
    if (p.dbg) {      // ns since kernel entry: TMEM ready, first slab landed, splitter done, MMAs done, reduction done, stores issued
        unsigned long long hst[8];
        cudaDeviceSynchronize();
        cudaMemcpy(hst, d_dbg, 56, cudaMemcpyDeviceToHost);
        fprintf(stderr, "gemm dbg split=%d:", n_split);
        for (int i = 1; i < 7; ++i) fprintf(stderr, " t%d=%lld", i, (long long)(hst[i] - hst[0]));
        fprintf(stderr, "\n");
    }
    return rc;
}
