/*
 * pa_prefill_tc.cu -- tensor-core variant of the causal multi-row paged attention (prefill):
 * tcgen05.mma kind::tf32 with accumulators in TMEM, K/V pages gathered by TMA tensor-map box
 * copies (one box per page and 32-column block, 128-byte hardware swizzle) completing on
 * mbarriers.  Used only where the contraction really is dense: many query rows per sequence.
 *
 * REDUCED PRECISION: Q, K, V and the probabilities enter the tensor cores as TF32 (10-bit
 * mantissa, fp32 accumulate) and exp is ex2.approx.  This path is opt-in
 * (PA_TUNE_PREFILL_PATH = 3) and has its own tolerance, stated in tests/gpu_common.py
 * (TC_REL_TOL); the default prefill path is the fp32 SIMT kernel in pa_prefill.cu (1e-5).
 *
 * Semantics: attention_paged rows (paged_infer.c:163-240), as pa_prefill.cu.
 *
 * One CTA per (head, sequence, tile of 128 query rows) = one TMEM lane per query row:
 *   warps 0..4*NWG-1  softmax warpgroups; warpgroup g owns key tiles g, g+NWG, ... with its own
 *                     online-softmax state (m, l, o) -- the states are merged at the end, so the
 *                     warpgroups never wait for each other.  A thread reads its row of S from
 *                     TMEM (tcgen05.ld), writes P back over it (tcgen05.st) and accumulates the
 *                     P.V tile from TMEM into registers.
 *   warp 4*NWG        TMA producer (+ TMEM allocation)
 *   warp 4*NWG+1      MMA issuer: S = Q.K^T (A = Q from TMEM, B = K from shared memory, K-major)
 *                     and O += P.V (A = P from TMEM, B = V from shared memory, MN-major)
 * Shared memory tiles are [32-column block][row][32 floats] with a 128-byte swizzle: for K that
 * is the K-major SW128 layout (rows = keys = N, 16-byte chunks XOR row%8), for V the MN-major
 * layout of a 32-bit operand (rows = keys = K, 32-byte chunks XOR row%4,
 * CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the same box shapes serve both, the tensor maps and
 * descriptors differ.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

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

constexpr float kMaxInit = -10000.0f;   // paged_infer.c:187
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kBM = 128;                // query rows per CTA = TMEM lanes

struct TcParams {
    const float* q;
    float* out;
    const int* kv_end;
    const int* kv_start;
    const int* q_row0;
    const int* table;
    int B, C, NH, bs, tstride, q_stride, out_stride;
    int n_tiles, layer;
    float sl2;              // scale * log2(e): scores are handled in the exp2 domain
};



// tile_lin -> (sequence, q tile): warp-parallel scan over ceil(nq/128)
__device__ __forceinline__ void find_tile(const TcParams& p, int tile_lin, int& seq, int& qt, int& n_qt) {
    const int lane = threadIdx.x & 31;
    int run = 0;
    seq = -1; qt = 0; n_qt = 0;
    for (int c = 0; c < p.B; c += 32) {
        const int i = c + lane;
        int n = 0;
        if (i < p.B) n = (p.q_row0[i + 1] - p.q_row0[i] + kBM - 1) / kBM;
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, run + incl > tile_lin);
        if (hit) {
            const int l = __ffs(hit) - 1;
            const int excl = __shfl_sync(0xffffffffu, incl - n, l);
            seq = c + l;
            qt = tile_lin - run - excl;
            n_qt = __shfl_sync(0xffffffffu, n, l);
            return;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

template <int HS, int BN, int NWG, int NST, int SBUF>
struct TcCfg {
    static constexpr int kThreads = NWG * 128 + 64;
    // two CTAs per SM when their tiles fit (one-warpgroup CTAs overlap through co-residency)
    static constexpr int kMinBlocks = (NWG == 1 && 2 * (1024 + 2 * NST * BN * HS * 4 + 512) <= 227 * 1024 && 2 * (HS + SBUF * BN + HS) <= 512) ? 2 : 1;
    static constexpr int kQBytes = kBM * HS * 4;
    static constexpr int kKVBytes = BN * HS * 4;
    static constexpr int kTileBytes = 2 * NST * kKVBytes;      // NST-deep rings of K and V tiles
    static constexpr int kBarBytes = ((2 * SBUF + 1) * NWG + 4 * NST) * 8 + 32;
    static constexpr size_t kSmem = 1024 + kTileBytes + kBarBytes;     // 1024: manual alignment slack
    static constexpr int kCols = HS + NWG * (SBUF * BN + HS);      // Q | S[NWG][SBUF] | O[NWG]
    static constexpr int kLag = NWG * SBUF - 1;              // Q.K^T runs this many key tiles ahead of P.V
    static_assert(NST > kLag, "the K ring must hold the tiles whose Q.K^T has been issued ahead");
    static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
    static_assert(HS + NWG * SBUF * BN + NWG * HS <= 512, "TMEM columns");
    static_assert((NWG - 1) * kBM * (HS + 2) * 4 <= 2 * NST * kKVBytes, "merge scratch must fit the K/V buffers");
};

template <int HS, int BN, int NWG, int NST, int SBUF>
__global__ void __launch_bounds__(NWG * 128 + 64, (TcCfg<HS, BN, NWG, NST, SBUF>::kMinBlocks))
pa_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const TcParams p) {
    using Cfg = TcCfg<HS, BN, NWG, NST, SBUF>;
    constexpr int DB = HS / 32;                         // 32-column blocks per row
    constexpr uint32_t kIdescQK = instr_desc(kBM, BN, 0, 0);
    constexpr uint32_t kIdescPV = instr_desc(kBM, HS, 0, 1);

    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* Ks = base;                                           // [NST][DB][BN][128 B]
    unsigned char* Vs = Ks + NST * Cfg::kKVBytes;                       // [NST][DB][BN][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + NST * Cfg::kKVBytes);
    uint64_t* k_full = bars;                 // [NST] TMA bytes of a K tile landed
    uint64_t* v_full = bars + NST;           // [NST]
    uint64_t* k_empty = bars + 2 * NST;      // [NST] the Q.K^T that read the K tile has completed
    uint64_t* v_empty = bars + 3 * NST;      // [NST] the P.V that read the V tile has completed
    uint64_t* s_full = bars + 4 * NST;       // [NWG][SBUF] Q.K^T committed: S readable
    uint64_t* p_ready = s_full + NWG * SBUF; // [NWG][SBUF] the warpgroup wrote P over S
    uint64_t* o_full = p_ready + NWG * SBUF; // [NWG] P.V committed: O holds this tile too
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NWG);
    int* s_unit = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kProducerWarp = NWG * 4, kMmaWarp = NWG * 4 + 1;

    const int h = blockIdx.x / p.n_tiles;
    const int tile_lin = blockIdx.x - h * p.n_tiles;
    if (warp == 0) {
        int seq, qt, n_qt;
        find_tile(p, tile_lin, seq, qt, n_qt);
        if (lane == 0) { s_unit[0] = seq; s_unit[1] = n_qt - 1 - qt; }     // heaviest q tile first
    }
    if (tid == 0) {
        for (int b = 0; b < NST; ++b) {
            mbar_init(smem_u32(&k_full[b]), 1);
            mbar_init(smem_u32(&v_full[b]), 1);
            mbar_init(smem_u32(&k_empty[b]), 1);
            mbar_init(smem_u32(&v_empty[b]), 1);
        }
        for (int b = 0; b < NWG * SBUF; ++b) {
            mbar_init(smem_u32(&s_full[b]), 1);
            mbar_init(smem_u32(&p_ready[b]), 128);
        }
        for (int b = 0; b < NWG; ++b) mbar_init(smem_u32(&o_full[b]), 1);
        mbar_fence_init();
    }
    if (warp == kProducerWarp) {
        tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int seq = s_unit[0];
    const int qt = s_unit[1];

    int n_kt = 0, rows = 0, row0 = 0, nq = 0, kv_start = 0, kv_end = 0, j0 = 0, k_begin = 0;
    if (seq >= 0) {
        row0 = p.q_row0[seq];
        nq = p.q_row0[seq + 1] - row0;
        kv_start = p.kv_start[seq];
        kv_end = p.kv_end[seq];
        j0 = qt * kBM;
        rows = min(kBM, nq - j0);
        const int lim_last = kv_end - (nq - 1 - (j0 + rows - 1));
        k_begin = (kv_start / BN) * BN;
        n_kt = lim_last > k_begin ? (lim_last - k_begin + BN - 1) / BN : 0;
    }

    // ---- Q tile -> TMEM columns [0, HS): thread = query row = TMEM lane.  Q is the A operand of
    // every S = Q.K^T instruction; read from TMEM it costs no shared-memory bandwidth (from
    // shared memory the 4 KB of A per 8-deep k-step would exceed what the SM can feed the MMA).
    if (warp < NWG * 4 && n_kt > 0) {
        const int g = warp >> 2, wq = warp & 3;
        const int r = wq * 32 + lane;
        constexpr int kColsPerWg = HS / NWG;             // each warpgroup stores its share of the columns
        const bool ok = r < rows;
        const float* src = p.q + (size_t)(row0 + j0 + (ok ? r : 0)) * p.q_stride + h * HS + g * kColsPerWg;
#pragma unroll
        for (int c = 0; c < kColsPerWg; c += 32) {
            float qv[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(src + c + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                qv[i] = t.x; qv[i + 1] = t.y; qv[i + 2] = t.z; qv[i + 3] = t.w;
            }
            tmem_st32(tmem_base + ((uint32_t)(wq * 32) << 16) + g * kColsPerWg + c, qv);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == kProducerWarp) {
        // ================================ TMA producer ========================================
        if (n_kt > 0) {
            const int* tbl = p.table + (size_t)seq * p.tstride;
            const int n_pages = (kv_end + p.bs - 1) / p.bs;
            const int ppt = BN / p.bs;                                   // pages per key tile
            // page ids of a tile: one lane each, fetched one tile ahead of their use
            auto fetch_pages = [&](int it) {
                const int pg = (k_begin + it * BN) / p.bs + lane;
                return (lane < ppt && it < n_kt) ? __ldg(tbl + min(pg, n_pages - 1)) : 0;   // pages past the last one repeat it (their keys are masked)
            };
            int page_next = fetch_pages(0);
            const uint32_t page_bytes = (uint32_t)p.bs * 128u;
            const bool leader = elect_one();
            for (int it = 0; it < n_kt; ++it) {
                const int st = it % NST, j = it / NST;
                const int page_cur = page_next;
                page_next = fetch_pages(it + 1);
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    // the ring slot is free once the MMA that read its previous content has completed
                    if (j > 0) mbar_wait(smem_u32(kv == 0 ? &k_empty[st] : &v_empty[st]), (j - 1) & 1);
                    const uint32_t bar = smem_u32(kv == 0 ? &k_full[st] : &v_full[st]);
                    const uint32_t dst0 = smem_u32((kv == 0 ? Ks : Vs) + st * Cfg::kKVBytes);
                    const CUtensorMap* map = kv == 0 ? &tm_k : &tm_v;
                    if (leader) mbar_arrive_expect_tx(bar, Cfg::kKVBytes);
                    for (int pi = 0; pi < ppt; ++pi) {
                        const int row = __shfl_sync(0xffffffffu, page_cur, pi) * p.bs;
                        if (leader) {
#pragma unroll
                            for (int db = 0; db < DB; ++db)
                                tma_box_3d(dst0 + db * (BN * 128) + pi * page_bytes, map, h * HS + db * 32, row, p.layer, bar);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================= MMA issuer =========================================
        if (n_kt > 0) {
            const bool leader = elect_one();
            auto issue_pv = [&](int it) {
                const int b = it % NWG, j = it / NWG, st = it % NST;
                const int sb = b * SBUF + (j % SBUF);              // S/P buffer of this tile
                mbar_wait(smem_u32(&v_full[st]), (it / NST) & 1);
                mbar_wait(smem_u32(&p_ready[sb]), (j / SBUF) & 1);
                tc_fence_after();
                const uint32_t v_lo32 = smem_desc_lo(smem_u32(Vs + st * Cfg::kKVBytes), BN * 128);      // + 64 (1024 bytes) per k-step (pa_ptx.cuh)
                constexpr uint32_t v_hi32 = smem_desc_hi(512, 1);
                const uint32_t p_tmem = tmem_base + HS + sb * BN;
                const uint32_t o_tmem = tmem_base + HS + NWG * SBUF * BN + b * HS;
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < BN / 8; ++ks)          // 8 keys per instruction = two 4-row swizzle groups
                        mma_tf32_ts_lohi(o_tmem, p_tmem + ks * 8, v_lo32 + ks * 64, v_hi32, kIdescPV, (j > 0 || ks > 0) ? 1u : 0u);
                    tc_commit(smem_u32(&o_full[b]));
                    tc_commit(smem_u32(&v_empty[st]));
                }
                __syncwarp();
            };
            for (int it = 0; it < n_kt; ++it) {
                const int b = it % NWG, st = it % NST;
                const int sb = b * SBUF + ((it / NWG) % SBUF);
                mbar_wait(smem_u32(&k_full[st]), (it / NST) & 1);
                tc_fence_after();
                const uint32_t k_lo32 = smem_desc_lo(smem_u32(Ks + st * Cfg::kKVBytes), 16);
                constexpr uint32_t k_hi32 = smem_desc_hi(1024, 2);
                // the P.V that read this S/P buffer last was issued SBUF tiles of this warpgroup ago, before
                // this instruction in program order: the tensor pipe executes them in order
                const uint32_t s_tmem = tmem_base + HS + sb * BN;
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HS / 8; ++ks) {        // 8 floats (32 B) of the head dimension per instruction
                        const uint32_t koff = (ks >> 2) * (BN * 8) + (ks & 3) * 2;      // descriptor units (16 bytes)
                        mma_tf32_ts_lohi(s_tmem, tmem_base + ks * 8, k_lo32 + koff, k_hi32, kIdescQK, ks > 0);
                    }
                    tc_commit(smem_u32(&s_full[sb]));
                    tc_commit(smem_u32(&k_empty[st]));
                }
                __syncwarp();
                if (it >= Cfg::kLag) issue_pv(it - Cfg::kLag);
            }
            for (int it = max(0, n_kt - Cfg::kLag); it < n_kt; ++it) issue_pv(it);
        }
    } else {
        // ============================== softmax warpgroups ====================================
        const int g = warp >> 2;                         // warpgroup
        const int wq = warp & 3;                         // TMEM lane quarter of this warp
        const int r = wq * 32 + lane;                    // query row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const int lim = kv_end - (nq - 1 - (j0 + r));    // this row sees keys [kv_start, lim)
        const int lim_first = kv_end - (nq - 1 - j0);
        const uint32_t s_tmem0 = tmem_base + lane_off + HS + g * SBUF * BN;
        const uint32_t o_tmem = tmem_base + lane_off + HS + NWG * SBUF * BN + g * HS;
        // exp2-domain online softmax.  The O accumulator stays in TMEM across this warpgroup's key
        // tiles; it is rescaled only when the row maximum has grown by more than 2^8 since the
        // maximum in use (probabilities stay <= 256, far inside fp32/tf32 range), so the common
        // tile costs one TMEM read of S and one write of P per thread.
        float m_run = kMaxInit * kLog2e, l_run = 0.0f;
        int n_mine = 0;
        // One key tile of this warpgroup.  MASK is a compile-time flag and the two instances are
        // reached through a real branch: only tiles that touch the window start or a row's causal
        // limit pay for the per-element compares (as predicated code they would cost issue slots
        // on every tile).
        auto tile = [&](auto mask_tag, int it) {
            constexpr bool MASK = decltype(mask_tag)::value;
            const int j = it / NWG;
            const int g0 = k_begin + it * BN;
            const int sb = g * SBUF + (j % SBUF);
            const uint32_t s_tmem = s_tmem0 + (j % SBUF) * BN;
            mbar_wait(smem_u32(&s_full[sb]), (j / SBUF) & 1);
            tc_fence_after();
            float sv[BN];
#pragma unroll
            for (int c = 0; c < BN; c += 32) tmem_ld32(s_tmem + c, sv + c);
            tmem_wait_ld();
            if (MASK) {
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    const int key = g0 + i;
                    if (key < kv_start || key >= lim) sv[i] = -INFINITY;
                }
            }
            // the scale is positive, so the maximum can be taken on the raw scores
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < BN; ++i) mx = fmaxf(mx, sv[i]);
            const float m_new = fmaxf(m_run, mx * p.sl2);
            const bool grow = (m_new - m_run) > 8.0f;
            const float m_use = grow ? m_new : m_run;
            const float neg_m = -m_use;
            float psum = 0.0f;
#pragma unroll
            for (int i = 0; i < BN; ++i) {
                const float e = ex2_approx(fmaf(sv[i], p.sl2, neg_m));
                sv[i] = e;
                psum += e;
            }
#pragma unroll
            for (int c = 0; c < BN; c += 32) tmem_st32(s_tmem + c, sv + c);
            const float alpha = grow ? ex2_approx(m_run - m_new) : 1.0f;
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                // rescale this warpgroup's O: its previous P.V must have completed, and the next one
                // cannot start before p_ready below
                mbar_wait(smem_u32(&o_full[g]), (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < HS; c += 32) {
                    float ov[32];
                    tmem_ld32(o_tmem + c, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ov[i] *= alpha;
                    tmem_st32(o_tmem + c, ov);
                }
            }
            l_run = l_run * alpha + psum;
            m_run = m_use;
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&p_ready[sb]));
        };
        for (int it = g; it < n_kt; it += NWG, ++n_mine) {
            const int g0 = k_begin + it * BN;
            if ((g0 < kv_start) || (g0 + BN > lim_first)) tile(std::true_type{}, it);
            else tile(std::false_type{}, it);
        }
        // ---- this warpgroup's O out of TMEM -----------------------------------------------------
        float o[HS];
        if (n_mine > 0) {
            mbar_wait(smem_u32(&o_full[g]), (n_mine - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < HS; c += 32) tmem_ld32(o_tmem + c, o + c);
            tmem_wait_ld();
        } else {
#pragma unroll
            for (int i = 0; i < HS; ++i) o[i] = 0.0f;
        }
        tc_fence_before();

        // ---- merge the warpgroups' states, normalise, store ----------------------------------
        if (NWG > 1) {
            // a warpgroup leaves its loop when ITS last P.V has completed; the other one may still
            // be using the K/V buffers that double as merge scratch, so meet first
            asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
            float* scratch = reinterpret_cast<float*>(Ks);               // [NWG-1][128][HS+2]
            if (g > 0) {
                float* dst = scratch + ((g - 1) * kBM + r) * (HS + 2);
#pragma unroll
                for (int i = 0; i < HS; ++i) dst[i] = o[i];
                dst[HS] = m_run;
                dst[HS + 1] = l_run;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
            if (g == 0) {
#pragma unroll 1
                for (int og = 1; og < NWG; ++og) {
                    const float* src = scratch + ((og - 1) * kBM + r) * (HS + 2);
                    const float m1 = src[HS], l1 = src[HS + 1];
                    const float m = fmaxf(m_run, m1);
                    const float w0 = ex2_approx(m_run - m), w1 = ex2_approx(m1 - m);
                    l_run = l_run * w0 + l1 * w1;
#pragma unroll
                    for (int i = 0; i < HS; ++i) o[i] = o[i] * w0 + src[i] * w1;
                    m_run = m;
                }
            }
        }
        if (g == 0 && r < rows && seq >= 0) {
            const float inv = (l_run == 0.0f) ? 0.0f : 1.0f / l_run;
            float* dst = p.out + (size_t)(row0 + j0 + r) * p.out_stride + h * HS;
#pragma unroll
            for (int i = 0; i < HS; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarp) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct TcState {
    CUtensorMap tm_k, tm_v;
    bool ready;
};

// pool viewed as (layer, row = page*bs + slot, column) fp32; box = one page x 32 columns, 128-byte swizzle
int make_pool_map(CUtensorMap* map, float* pool, const pa_handle* h, CUtensorMapSwizzle swizzle) {
    pa_encode_tiled_fn enc = pa_get_encode_tiled();
    if (!enc) { pa_set_error("cuTensorMapEncodeTiled not available from the driver"); return PA_ERR_CUDA; }
    const cuuint64_t rows = (cuuint64_t)h->cfg.max_blocks * h->cfg.block_size;
    cuuint64_t dims[3] = {(cuuint64_t)h->C, rows, (cuuint64_t)h->cfg.n_layers};
    cuuint64_t strides[2] = {(cuuint64_t)h->C * 4, (cuuint64_t)h->layer_stride * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)h->cfg.block_size, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pool, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pa_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return PA_ERR_CUDA; }
    return PA_OK;
}

template <int HS, int BN, int NWG, int NST, int SBUF>
int launch_tc(const TcState* st, const TcParams& p, cudaStream_t s) {
    using Cfg = TcCfg<HS, BN, NWG, NST, SBUF>;
    auto fn = pa_prefill_tc_kernel<HS, BN, NWG, NST, SBUF>;
    static std::atomic<unsigned long long> attr_done{0};       // per instantiation; one bit per device
    CU_CHECK(pa_optin_smem(attr_done, fn, (int)Cfg::kSmem));
    fn<<<(unsigned)((long long)p.n_tiles * p.NH), Cfg::kThreads, Cfg::kSmem, s>>>(st->tm_k, st->tm_v, p);
    CU_CHECK(cudaGetLastError());
    return PA_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" void pa_cu_prefill_tc_release(pa_handle* h) {
    free(h->tc_state);
    h->tc_state = nullptr;
}

// PA_OK = launched; PA_ERR_UNSUPPORTED = outside the kernel's domain
extern "C" int pa_cu_prefill_tc(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride,
                                void* stream) {
    const pa_step_layout& L = h->step;
    const int hs = h->cfg.head_dim, bs = h->cfg.block_size;
    if (!(hs == 64 || hs == 128)) return PA_ERR_UNSUPPORTED;
    // keys per tile: head_dim 64 runs best with 64-key tiles, two 96 KB CTAs per SM (measured,
    // profiles/r01_prefill.md); 128-key tiles on request
    const int BN = (hs == 64 && h->tune[PA_TUNE_TC_KEY_TILE] == 128) ? 128 : 64;
    // a page must be whole 8-row swizzle groups and divide the key tile
    if (bs < 8 || (bs & (bs - 1)) || bs > BN) return PA_ERR_UNSUPPORTED;
    if ((h->C % 4) || (q_stride % 4) || (out_stride % 4) || !aligned16(q) || !aligned16(out)) return PA_ERR_UNSUPPORTED;
    TcState* st = (TcState*)h->tc_state;
    if (!st) {
        void* mem = nullptr;
        if (posix_memalign(&mem, 64, sizeof(TcState)) != 0) { pa_set_error("out of host memory"); return PA_ERR_NOMEM; }
        st = (TcState*)mem;
        memset(st, 0, sizeof(*st));
        h->tc_state = st;
    }
    if (!st->ready) {
        // K is a K-major operand (16-byte swizzle chunks), V an MN-major one (32-byte chunks)
        int rc = make_pool_map(&st->tm_k, h->pool_k, h, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == PA_OK) rc = make_pool_map(&st->tm_v, h->pool_v, h, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc != PA_OK) return rc;
        st->ready = true;
    }
    TcParams p;
    p.q = q; p.out = out;
    p.kv_end = h->d_step + L.off_kv_end;
    p.kv_start = h->d_step + L.off_kv_start;
    p.q_row0 = h->d_step + L.off_q_row0;
    p.table = h->d_step + L.off_table;
    p.B = L.nseq; p.C = h->C; p.NH = h->cfg.n_heads; p.bs = bs;
    p.tstride = L.tstride; p.q_stride = q_stride; p.out_stride = out_stride;
    p.layer = layer;
    p.sl2 = (float)(1.0 / sqrtf((float)hs)) * kLog2e;
    long long n_tiles = 0;
    const int* qr = h->h_step + L.off_q_row0;
    for (int i = 0; i < L.nseq; ++i) n_tiles += (qr[i + 1] - qr[i] + kBM - 1) / kBM;
    if (n_tiles == 0) return PA_OK;
    if (n_tiles * p.NH > 0x7fffffffLL) return PA_ERR_UNSUPPORTED;
    p.n_tiles = (int)n_tiles;
    cudaStream_t s = (cudaStream_t)stream;
    // measured (profiles/r01_prefill.md): one softmax warpgroup with two S/P buffers beats two
    // warpgroups with one buffer each for both head sizes (head_dim 64: two such CTAs per SM)
    const int want_wg = h->tune[PA_TUNE_TC_WARPGROUPS];
    const int nwg = want_wg == 2 ? 2 : 1;
    int rc;
    // K/V rings are 3 tiles deep (the load of tile i+3 starts when the MMAs of tile i have completed:
    // with 2 the tensor pipe waited a full L2 round trip every other tile)
    // One-warpgroup CTAs keep two S/P buffers, so Q.K^T of the next tile runs while the softmax of the
    // current one does; two-warpgroup CTAs alternate warpgroups instead (TMEM has no room for both).
    if (hs == 64 && BN == 128) rc = nwg == 1 ? launch_tc<64, 128, 1, 3, 2>(st, p, s) : launch_tc<64, 128, 2, 3, 1>(st, p, s);
    else if (hs == 64) rc = nwg == 1 ? launch_tc<64, 64, 1, 3, 2>(st, p, s) : launch_tc<64, 64, 2, 3, 1>(st, p, s);
    else rc = nwg == 1 ? launch_tc<128, 64, 1, 3, 2>(st, p, s) : launch_tc<128, 64, 2, 3, 1>(st, p, s);
    if (rc == PA_OK) h->launches++;
    return rc;
}
