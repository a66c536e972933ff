/*
 * pa_prefill_tc3.cu -- fp32-ACCURATE tensor-core prefill: the causal multi-row paged attention of
 * attention_paged (paged_infer.c:163-240) on tcgen05 with the 3xTF32 split, inside the path's 1e-5 tolerance.
 *
 * The tensor core reads the top 19 bits of an fp32 operand (a_hi); with a_lo = a - a_hi (exact in fp32)
 *      a.b  ~=  a_lo.b_hi + a_hi.b_lo + a_hi.b_hi          (the dropped a_lo.b_lo term is 2^-22 relative)
 * for both contractions of the kernel, S = Q.K^T and O_tile = P.V.  What keeps it at fp32 accuracy:
 *   - the two small products are issued FIRST into a fresh TMEM accumulator and the leading one last: the
 *     tensor core's fp32 accumulate truncates, so only the hs/8 (resp. BN/8) leading k-steps of a key tile
 *     round at full magnitude;
 *   - the running output is NOT accumulated in TMEM across key tiles (thousands of truncating steps at 32k
 *     context): every key tile's P.V lands in a fresh accumulator and is added to the row's running output in
 *     registers with round-to-nearest fp32 (o = o * alpha + o_tile, the online-softmax rescale);
 *   - softmax in fp32, running max from the reference's -10000, `sum == 0 -> 0`; exp as one FMA + ex2.approx in the
 *     exp2 domain (2^-22 relative; PA_PREFILL_TC3_EXPF=1 switches to expf as the reference writes it).
 * Measured against the oracle: see tests/test_gpu_parity.py (prefill cases, path 4) and profiles/r02_prefill.md.
 *
 * PERSISTENT: one CTA per SM walks its column of a host-built schedule of units (sequence, tile of 128 query rows =
 * TMEM lanes, head) -- see build_schedule() -- and nothing drains between units (see the kernel's comment).
 * 12 warps (three warpgroups):
 *   warps 0-3   softmax warpgroup: thread = query row.  Reads its row of S (tcgen05.ld), online softmax, writes
 *               P and P_lo back into TMEM as the A operands of P.V (tcgen05.st), folds the PREVIOUS tile's
 *               P.V result into its register accumulator while the tensor core works on the current one, and
 *               stores a finished unit's rows (after the next unit's first key tile)
 *   warps 4-7   splitter warpgroup: K_lo / V_lo tiles next to the raw ones the TMA delivered (element-wise over
 *               the flat swizzled buffers: same layout, other base address); the next unit's Q tile: prefetched
 *               into the staging tile (cp.async), split and moved to TMEM at the unit boundary
 *   warp 8      TMA producer (+ TMEM allocation): one tensor-map box per page and 32-column block
 *   warp 9      MMA issuer (one elected lane): S = Q.K^T (A = Q, Q_lo from TMEM; B = K, K_lo from shared memory,
 *               K-major SW128) and O_tile = P.V (A = P, P_lo from TMEM; B = V, V_lo MN-major SW128/32B atoms)
 * Shared memory: 3-deep rings of K, K_lo, V, V_lo (192 KB) + the 32 KB Q / O staging tile.
 * TMEM columns: Q hs | Q_lo hs | SBUF x (S/P BN | P_lo BN) | OBUF x O_tile hs  <= 512.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

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
constexpr int kBM = 128;                // query rows per CTA = TMEM lanes
constexpr float kLog2e = 1.4426950408889634f;

struct Tc3Params {
    const float* q;
    float* out;
    const int* kv_end;
    const int* kv_start;
    const int* q_row0;
    const int* table;
    const int2* units;      // the schedule: [n_rows][gridDim.x], CTA c's j-th unit at [j][c]; x = sequence (-1: none), y = q tile | head << 16
    unsigned long long* dbg;   // optional timeline (PA_PREFILL_TC3_TIMELINE=1)
    int bs, tstride, q_stride, out_stride;
    int n_rows, layer;
    float scale;
    float sl2;              // scale * log2(e)
};

// The split a = hi + lo with hi = a ROUNDED to tf32 (cvt.rna: nearest, low 13 bits zero) rather than the truncation
// the tensor core would apply to a raw fp32 operand: |lo| <= 2^-11 |a| instead of 2^-10, so what the tensor core
// drops of lo (it keeps lo's top 10 mantissa bits) is 2^-21 |a| instead of 2^-20.  Both halves are therefore
// materialised (hi written over the raw operand), and a - hi is exact in fp32.
// (integer form of cvt.rna.tf32.f32 -- nearest, ties away from zero, on the sign-magnitude bit pattern: an add and
// a mask on the integer pipe; the cvt instruction runs on the quarter-rate conversion unit the exponentials need)
__device__ __forceinline__ float tf32_hi(float a) {
    return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}

template <int HS, int BN, int NST, int SBUF, int OBUF>
struct Tc3Cfg {
    // whole warpgroups (setmaxnreg works on warpgroups): softmax | splitter | producer, MMA issuer, two idle warps
    static constexpr int kThreads = 3 * 128;
    // registers per thread after setmaxnreg: the launch gives 65536 / 384 = 168
    static constexpr int kRegSoftmax = 232, kRegOther = 96;
    static_assert((kRegSoftmax + 2 * kRegOther) * 128 <= 65536, "register file");
    static constexpr int kKVBytes = BN * HS * 4;
    static constexpr int kRingBytes = 4 * NST * kKVBytes;      // NST-deep rings of K, K_lo, V, V_lo tiles
    static constexpr int kStageCols = 64;                      // the Q / O staging tile: 128 rows x 64 columns fp32
    static constexpr int kStageBytes = kBM * kStageCols * 4;
    static constexpr int kNumBars = 6 * NST + 2 * SBUF + 2 * OBUF + 2;
    static constexpr int kBarBytes = kNumBars * 8 + 16;
    static constexpr size_t kSmem = 1024 + kRingBytes + kStageBytes + kBarBytes;     // 1024: manual alignment slack
    static constexpr int kQ = 0, kQlo = HS, kSP = 2 * HS, kO = 2 * HS + SBUF * 2 * BN;
    static constexpr int kCols = kO + OBUF * HS;
    static constexpr int kLag = SBUF - 1;                    // Q.K^T runs this many key tiles ahead of P.V
    static_assert(HS % kStageCols == 0, "head_dim in staging tiles");
    static_assert(BN == 32 || BN == 64, "the key mask of a tile is one or two 32-bit words");
    static_assert(NST > kLag, "the K ring must hold the tiles whose Q.K^T has been issued ahead");
    static_assert(kCols <= 512, "TMEM columns");
    static_assert(kSmem <= 227 * 1024, "shared memory");
};

// byte offset of 16-byte chunk `c` (0..15) of row `r` in the staging tile: rows of 256 bytes, chunks XOR-swizzled by
// the row so that both access patterns are bank-conflict free -- a thread walking its own row (8 consecutive rows per
// quarter-warp: 8 different chunk positions) and a warp reading two whole rows (16 lanes x 16 bytes each)
__device__ __forceinline__ uint32_t stage_off(int r, int c) { return (uint32_t)(r * 256 + ((c ^ (r & 15)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    // (volatile, no memory clobber: ordered against the other asm statements, but several may be in flight)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// bits [lo, hi) of a 32-bit word, for any integers lo, hi
__device__ __forceinline__ uint32_t bit_range32(int lo, int hi) {
    lo = max(lo, 0); hi = min(hi, 32);
    if (hi <= lo) return 0u;
    return (0xffffffffu >> (32 - hi)) & (0xffffffffu << lo);
}

// A PERSISTENT kernel: one CTA per SM walks its column of the schedule (units = (sequence, tile of 128 query rows,
// head), dealt to the CTAs by the host, longest first) and NOTHING drains between units: all rings and mbarrier phases
// keep counting (base_it), the producer and the splitter run ahead into the next unit's keys, the next Q tile waits
// in the staging tile and replaces the current one in TMEM as soon as the current unit's last Q.K^T has completed, the
// issuer goes on alternating Q.K^T(i) and P.V(i - 1) across the boundary, and the softmax warpgroup folds in a unit's
// last P.V and stores its rows after the softmax of the next unit's first key tile.
template <int HS, int BN, int NST, int SBUF, int OBUF, bool EXPF>
__global__ void __launch_bounds__(384, 1)
pa_prefill_tc3_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const Tc3Params p) {
    using Cfg = Tc3Cfg<HS, BN, NST, SBUF, OBUF>;
    constexpr int DB = HS / 32;                         // 32-column blocks per row
    constexpr int NCH = HS / Cfg::kStageCols;           // staging tiles per row
    constexpr uint32_t kIdescQK = instr_desc(kBM, BN, 0, 0);
    constexpr uint32_t kIdescPV = instr_desc(kBM, HS, 0, 1);

    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* Ks = base;                                           // [NST][DB][BN][128 B]
    unsigned char* Kl = Ks + NST * Cfg::kKVBytes;                       // K_lo, same layout
    unsigned char* Vs = Kl + NST * Cfg::kKVBytes;
    unsigned char* Vl = Vs + NST * Cfg::kKVBytes;
    unsigned char* Stg = Vl + NST * Cfg::kKVBytes;                      // [128][64] fp32, swizzled (stage_off)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Stg + Cfg::kStageBytes);
    uint64_t* k_full = bars;                 // [NST] TMA bytes of a K tile landed
    uint64_t* v_full = bars + NST;           // [NST]
    uint64_t* k_split = bars + 2 * NST;      // [NST] K_lo written (128 splitter threads)
    uint64_t* v_split = bars + 3 * NST;      // [NST]
    uint64_t* k_empty = bars + 4 * NST;      // [NST] the Q.K^T that read K / K_lo has completed
    uint64_t* v_empty = bars + 5 * NST;      // [NST] the P.V that read V / V_lo has completed
    uint64_t* s_full = bars + 6 * NST;       // [SBUF] Q.K^T committed: S readable
    uint64_t* p_ready = s_full + SBUF;       // [SBUF] the softmax threads wrote P, P_lo (128 arrivals)
    uint64_t* o_full = p_ready + SBUF;       // [OBUF] P.V committed: the tile's O readable
    uint64_t* o_free = o_full + OBUF;        // [OBUF] the softmax threads have taken it into registers (128 arrivals)
    uint64_t* q_ready = o_free + OBUF;       // [1] a unit's Q tile is in TMEM and has left the staging tile (128 arrivals); one phase per unit
    uint64_t* o_stored = q_ready + 1;        // [1] a unit's output has left the staging tile (128 softmax threads); one phase per unit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_stored + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kSplitWarp0 = 4, kProducerWarp = 8, kMmaWarp = 9;

    // developer timeline: 8 stamps per CTA, then (CTA 0 only) 3 per key tile
    auto stamp = [&](int slot) {
        if (p.dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[(size_t)blockIdx.x * 8 + slot] = t; }
    };
    auto stamp_tile = [&](int gi, int k) {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && gi < 160) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[(size_t)gridDim.x * 8 + gi * 3 + k] = t;
        }
    };
    if (tid == 0) stamp(0);
    if (tid == 0) {
        for (int b = 0; b < NST; ++b) {
            mbar_init(smem_u32(&k_full[b]), 1);
            mbar_init(smem_u32(&v_full[b]), 1);
            mbar_init(smem_u32(&k_split[b]), 128);
            mbar_init(smem_u32(&v_split[b]), 128);
            mbar_init(smem_u32(&k_empty[b]), 1);
            mbar_init(smem_u32(&v_empty[b]), 1);
        }
        for (int b = 0; b < SBUF; ++b) {
            mbar_init(smem_u32(&s_full[b]), 1);
            mbar_init(smem_u32(&p_ready[b]), 128);
        }
        for (int b = 0; b < OBUF; ++b) {
            mbar_init(smem_u32(&o_full[b]), 1);
            mbar_init(smem_u32(&o_free[b]), 128);
        }
        mbar_init(smem_u32(q_ready), 128);
        mbar_init(smem_u32(o_stored), 128);
        mbar_fence_init();
    }
    if (warp == kProducerWarp) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) stamp(1);

    // the CTA's j-th unit; n_kt == 0 (no key visible to any of its rows) is legal: its rows are zero, no role touches
    // a ring or a barrier for it
    struct Geo { int seq, h, n_kt, rows, row_g, kv_start, kv_end, k_begin, lim_first; };
    auto geo = [&](int j) {
        Geo g;
        g.seq = -1;
        if (j >= p.n_rows) return g;
        const int2 u = __ldg(p.units + (size_t)j * gridDim.x + blockIdx.x);
        if (u.x < 0) return g;
        g.seq = u.x;
        g.h = (int)((unsigned)u.y >> 16);
        const int j0 = (u.y & 0xffff) * kBM;
        const int row0 = __ldg(p.q_row0 + g.seq), nq = __ldg(p.q_row0 + g.seq + 1) - row0;
        g.kv_start = __ldg(p.kv_start + g.seq);
        g.kv_end = __ldg(p.kv_end + g.seq);
        g.rows = min(kBM, nq - j0);
        g.row_g = row0 + j0;
        g.lim_first = g.kv_end - (nq - 1 - j0);                   // row r of the tile sees keys [kv_start, min(kv_end, lim_first + r))
        const int lim_last = g.lim_first + g.rows - 1;
        g.k_begin = (g.kv_start / BN) * BN;
        g.n_kt = lim_last > g.k_begin ? (lim_last - g.k_begin + BN - 1) / BN : 0;
        return g;
    };

    // Q tile -> TMEM through the staging tile; executed by one warpgroup (warp quarter wq = the 32 query rows = TMEM
    // lanes the warp owns, so only the warp itself has to agree on its part of the staging tile).
    // global -> staging tile, 64 columns of the warp's 32 rows: 16-byte async copies, two whole rows per warp
    // instruction (a thread reading its own row straight from global memory touches 32 lines per instruction:
    // measured 3 us per Q tile).  Rows past the tile's last are zero-filled.
    const uint32_t stg = smem_u32(Stg);
    auto fetch_q = [&](const Geo& g, int ch) {
        const int wq = warp & 3;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const int r = wq * 32 + 2 * k + (lane >> 4), c = lane & 15;
            const bool ok = r < g.rows;
            const float* src = p.q + (size_t)(g.row_g + (ok ? r : 0)) * p.q_stride + g.h * HS + ch * Cfg::kStageCols + c * 4;
            cp_async16(stg + stage_off(r, c), src, ok ? 16 : 0);
        }
        cp_async_commit();
    };
    // staging tile -> TMEM: thread = query row = TMEM lane; rounded-to-tf32 columns to [kQ, kQ + HS), the remainders
    // to [kQlo, kQlo + HS).  `prev_qk` >= 0: the key tile whose Q.K^T (the last one issued with the previous Q) must
    // have completed before Q is replaced -- its k_empty commit; the next completion on that ring slot needs the new
    // Q, so the phase cannot run away.
    auto q_to_tmem = [&](const Geo& g, int prev_qk) {
        const int wq = warp & 3;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (ch > 0) fetch_q(g, ch);
            cp_async_wait<0>();
            __syncwarp();
            if (ch == 0 && prev_qk >= 0) mbar_wait(smem_u32(&k_empty[prev_qk % NST]), (prev_qk / NST) & 1);
            tc_fence_after();
            const int r = wq * 32 + lane;
#pragma unroll
            for (int c = 0; c < Cfg::kStageCols; c += 16) {
                float qv[16], qh[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = lds128(stg + stage_off(r, c / 4 + i));
                    qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) { qh[i] = tf32_hi(qv[i]); qv[i] -= qh[i]; }
                tmem_st16(tmem_base + ((uint32_t)(wq * 32) << 16) + Cfg::kQ + ch * Cfg::kStageCols + c, qh);
                tmem_st16(tmem_base + ((uint32_t)(wq * 32) << 16) + Cfg::kQlo + ch * Cfg::kStageCols + c, qv);
            }
            tmem_wait_st();
            __syncwarp();                      // every lane has read its row: the warp's part of the tile may be refilled
        }
        tc_fence_before();
        mbar_arrive(smem_u32(q_ready));
    };

    // Registers follow the roles: the kernel is compiled for 65536 / threads per thread; the splitter and the producer /
    // issuer warpgroups hand most of theirs back and the softmax warpgroup -- a query row's running output (hs floats)
    // plus a key tile of scores per thread -- takes them (setmaxnreg).
    // Every role walks the same list of units; `live` counts those with key tiles (the phases of q_ready / o_stored),
    // base_it the key tiles so far (the ring slots and the phases of everything else).
    if (warp >= kProducerWarp) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegOther));
      if (warp == kProducerWarp) {
        // ================================ TMA producer ========================================
        int base_it = 0;
        const bool leader = elect_one();
        for (int j = 0;; ++j) {
            const Geo g = geo(j);
            if (g.seq < 0) break;
            if (g.n_kt == 0) continue;
            const int* tbl = p.table + (size_t)g.seq * p.tstride;
            const int n_pages = (g.kv_end + p.bs - 1) / p.bs;
            const int ppt = BN / p.bs;                                   // pages per key tile
            auto fetch_pages = [&](int it) {
                const int pg = (g.k_begin + it * BN) / p.bs + lane;
                return (lane < ppt && it < g.n_kt) ? __ldg(tbl + min(pg, n_pages - 1)) : 0;   // pages past the last one repeat it (their keys are masked)
            };
            int page_next = fetch_pages(0);
            const uint32_t page_bytes = (uint32_t)p.bs * 128u;
            for (int it = 0; it < g.n_kt; ++it) {
                const int gi = base_it + it, st = gi % NST, use = gi / NST;
                const int page_cur = page_next;
                page_next = fetch_pages(it + 1);
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    // the ring slot (raw and lo) is free once the MMA that read its previous content has completed
                    if (use > 0) mbar_wait(smem_u32(kv == 0 ? &k_empty[st] : &v_empty[st]), (use - 1) & 1);
                    const uint32_t bar = smem_u32(kv == 0 ? &k_full[st] : &v_full[st]);
                    const uint32_t dst0 = smem_u32((kv == 0 ? Ks : Vs) + st * Cfg::kKVBytes);
                    const CUtensorMap* map = kv == 0 ? &tm_k : &tm_v;
                    if (leader) mbar_arrive_expect_tx(bar, Cfg::kKVBytes);
                    for (int pi = 0; pi < ppt; ++pi) {
                        const int row = __shfl_sync(0xffffffffu, page_cur, pi) * p.bs;
                        if (leader) {
#pragma unroll
                            for (int db = 0; db < DB; ++db)
                                tma_box_3d(dst0 + db * (BN * 128) + pi * page_bytes, map, g.h * HS + db * 32, row, p.layer, bar);
                        }
                    }
                    __syncwarp();
                }
            }
            base_it += g.n_kt;
        }
      } else if (warp == kMmaWarp) {
        // ================================= MMA issuer =========================================
        int base_it = 0, live = 0;
        const bool leader = elect_one();
        auto issue_pv = [&](int gi) {                 // gi: key tile counted over the CTA's units
            const int st = gi % NST, sb = gi % SBUF, ob = gi % OBUF;
            mbar_wait(smem_u32(&v_split[st]), (gi / NST) & 1);           // V landed and V_lo written
            mbar_wait(smem_u32(&p_ready[sb]), (gi / SBUF) & 1);
            if (gi >= OBUF) mbar_wait(smem_u32(&o_free[ob]), (gi / OBUF - 1) & 1);      // the tile that used this O buffer is in registers
            tc_fence_after();
            // descriptors: low word = address field (+ LBO), advanced by (byte offset >> 4) per instruction; high word constant
            const uint32_t v_lo = smem_desc_lo(smem_u32(Vs + st * Cfg::kKVBytes), BN * 128), vl_lo = smem_desc_lo(smem_u32(Vl + st * Cfg::kKVBytes), BN * 128);
            constexpr uint32_t v_hi = smem_desc_hi(512, 1);
            const uint32_t p_tmem = tmem_base + Cfg::kSP + sb * 2 * BN, pl_tmem = p_tmem + BN;
            const uint32_t o_tmem = tmem_base + Cfg::kO + ob * HS;
            if (leader) {
                // small products first (fresh accumulator), the leading one last
#pragma unroll
                for (int ks = 0; ks < BN / 8; ++ks) {          // 8 keys per instruction = two 4-row swizzle groups (1024 bytes)
                    mma_tf32_ts_lohi(o_tmem, pl_tmem + ks * 8, v_lo + ks * 64, v_hi, kIdescPV, ks > 0 ? 1u : 0u);
                    mma_tf32_ts_lohi(o_tmem, p_tmem + ks * 8, vl_lo + ks * 64, v_hi, kIdescPV, 1u);
                }
#pragma unroll
                for (int ks = 0; ks < BN / 8; ++ks)
                    mma_tf32_ts_lohi(o_tmem, p_tmem + ks * 8, v_lo + ks * 64, v_hi, kIdescPV, 1u);
                tc_commit(smem_u32(&o_full[ob]));
                tc_commit(smem_u32(&v_empty[st]));
            }
            __syncwarp();
        };
        for (int j = 0;; ++j) {
            const Geo g = geo(j);
            if (g.seq < 0) break;
            if (g.n_kt == 0) continue;
            mbar_wait(smem_u32(q_ready), live & 1);                          // this unit's Q, Q_lo are in TMEM
            tc_fence_after();
            for (int it = 0; it < g.n_kt; ++it) {
                const int gi = base_it + it, st = gi % NST, sb = gi % SBUF;
                mbar_wait(smem_u32(&k_split[st]), (gi / NST) & 1);           // K landed and K_lo written
                tc_fence_after();
                const uint32_t k_lo = smem_desc_lo(smem_u32(Ks + st * Cfg::kKVBytes), 16), kl_lo = smem_desc_lo(smem_u32(Kl + st * Cfg::kKVBytes), 16);
                constexpr uint32_t k_hi = smem_desc_hi(1024, 2);
                // the P.V that read this S/P buffer last was issued SBUF tiles ago, before this instruction in
                // program order: the tensor pipe executes them in order
                const uint32_t s_tmem = tmem_base + Cfg::kSP + sb * 2 * BN;
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HS / 8; ++ks) {        // 8 floats (32 B) of the head dimension per instruction
                        constexpr int kBlk = BN * 128 / 16;      // a 32-column block, in descriptor units
                        const uint32_t koff = (ks >> 2) * kBlk + (ks & 3) * 2;
                        mma_tf32_ts_lohi(s_tmem, tmem_base + Cfg::kQlo + ks * 8, k_lo + koff, k_hi, kIdescQK, ks > 0 ? 1u : 0u);
                        mma_tf32_ts_lohi(s_tmem, tmem_base + Cfg::kQ + ks * 8, kl_lo + koff, k_hi, kIdescQK, 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < HS / 8; ++ks) {
                        constexpr int kBlk = BN * 128 / 16;
                        const uint32_t koff = (ks >> 2) * kBlk + (ks & 3) * 2;
                        mma_tf32_ts_lohi(s_tmem, tmem_base + Cfg::kQ + ks * 8, k_lo + koff, k_hi, kIdescQK, 1u);
                    }
                    tc_commit(smem_u32(&s_full[sb]));
                    tc_commit(smem_u32(&k_empty[st]));
                }
                __syncwarp();
                if (gi >= Cfg::kLag) issue_pv(gi - Cfg::kLag);               // (may be the previous unit's last tile)
            }
            base_it += g.n_kt;
            ++live;
        }
        for (int gi = max(0, base_it - Cfg::kLag); gi < base_it; ++gi) issue_pv(gi);
      }       // (the last two warps only complete the warpgroup)
    } else if (warp >= kSplitWarp0) {
        // ============================== splitter warpgroup ====================================
        // per key tile K -> (K_hi in place, K_lo beside it), V likewise: element-wise over the flat swizzled tile; and
        // every Q tile but the CTA's first (that one is the softmax warpgroup's, which has nothing else to do then):
        // prefetched into the staging tile as soon as that is free, moved to TMEM at the boundary
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegOther));
        const int t = tid - kSplitWarp0 * 32;
        int base_it = 0, live = 0;
        Geo g = geo(0);
        int j = 0;
        while (g.seq >= 0 && g.n_kt == 0) g = geo(++j);                      // the first unit with key tiles
        while (g.seq >= 0) {
            if (live > 0) q_to_tmem(g, base_it - 1);                         // (prefetched during the previous unit)
            Geo gn = geo(++j);                                                // the next one
            while (gn.seq >= 0 && gn.n_kt == 0) gn = geo(++j);
            bool prefetch_due = gn.seq >= 0;
            for (int it = 0; it < g.n_kt; ++it) {
                const int gi = base_it + it, st = gi % NST;
                const uint32_t par = (gi / NST) & 1;
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    mbar_wait(smem_u32(kv == 0 ? &k_full[st] : &v_full[st]), par);
                    float4* src = reinterpret_cast<float4*>((kv == 0 ? Ks : Vs) + st * Cfg::kKVBytes);
                    float4* dst = reinterpret_cast<float4*>((kv == 0 ? Kl : Vl) + st * Cfg::kKVBytes);
#pragma unroll 4
                    for (int i = 0; i < Cfg::kKVBytes / 16 / 128; ++i) {
                        const float4 v = src[t + i * 128];
                        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                        src[t + i * 128] = hi;                                         // hi over the raw tile
                        dst[t + i * 128] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    }
                    fence_proxy_async_smem();      // generic-proxy stores -> visible to the MMA
                    mbar_arrive(smem_u32(kv == 0 ? &k_split[st] : &v_split[st]));
                }
                if (prefetch_due) {
                    // The staging tile is free for the next unit's Q once this unit's Q has left it (live == 0: the
                    // softmax warpgroup's q_ready; otherwise this warp's own copy, above) and the previous unit's
                    // output has gone through it (the softmax warpgroup stores it after its first key tile of this
                    // unit).  Looked at after every key tile, waited for only after the unit's last: the key tiles
                    // must not queue up behind the store.
                    const uint32_t bar = smem_u32(live == 0 ? q_ready : o_stored);
                    const uint32_t par = live == 0 ? 0u : (uint32_t)((live - 1) & 1);
                    bool free_now = mbar_test(bar, par);
                    free_now = __all_sync(0xffffffffu, free_now);
                    if (!free_now && it == g.n_kt - 1) { mbar_wait(bar, par); free_now = true; }
                    if (free_now) { fetch_q(gn, 0); prefetch_due = false; }
                }
            }
            base_it += g.n_kt;
            ++live;
            g = gn;
        }
    } else {
        // ================================ softmax warpgroup ====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kRegSoftmax));
        const int wq = warp & 3;                         // TMEM lane quarter of this warp
        const int r = wq * 32 + lane;                    // query row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        float o[HS];                                      // the row's running output; rescaled and extended one key tile behind the softmax
#pragma unroll
        for (int i = 0; i < HS; ++i) o[i] = 0.0f;
        float alpha_pend = 1.0f;                          // rescale that belongs to the tile whose P.V is still in flight
        float m_run = 0.0f, l_run = 0.0f;
        // the unit whose key tiles are being processed
        int j = -1;                                       // its index
        int unit_end = 0;                                 // the key tile (counted over the CTA) at which the next unit starts
        int g0 = 0, lim = 0, lim_min = 0, kv_start = 0, h_cur = 0, rows_cur = 0, row_cur = 0;
        // the finished unit whose last P.V is still to be folded in and whose rows are still to be stored
        bool fin_due = false;
        int fin_h = 0, fin_rows = 0, fin_row = 0;
        float fin_l = 0.0f;

        // One loop over the CTA's key tiles, ONE copy of each stage in the instruction stream (the kernel is large;
        // code that runs once per unit would be fetched cold every time): [softmax of tile gi] then [fold in tile
        // gi - 1, and store its unit if that was the unit's last tile].
        for (int gi = 0;; ++gi) {
            bool have_tile = true;
            if (gi == unit_end) {
                // the unit is finished (its last P.V still in flight); step to the next one with key tiles
                if (gi > 0) { fin_due = true; fin_h = h_cur; fin_rows = rows_cur; fin_row = row_cur; fin_l = l_run; }
                Geo g;
                for (;;) {
                    g = geo(++j);
                    if (g.seq < 0 || g.n_kt > 0) break;
                    // no key visible to the tile: zeros (`sum == 0 -> 0`), straight to global memory
                    if (r < g.rows) {
                        float* dst = p.out + (size_t)(g.row_g + r) * p.out_stride + g.h * HS;
#pragma unroll
                        for (int i = 0; i < HS; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
                }
                have_tile = g.seq >= 0;
                if (have_tile) {
                    if (gi == 0) {                         // the CTA's first Q tile (the others are the splitter warpgroup's)
                        fetch_q(g, 0);
                        q_to_tmem(g, -1);
                    }
                    unit_end = gi + g.n_kt;
                    g0 = g.k_begin;
                    kv_start = g.kv_start;
                    lim = min(g.kv_end, g.lim_first + r);  // this row sees keys [kv_start, lim) (rows past the tile's last: nothing beyond the cache)
                    lim_min = min(g.kv_end, g.lim_first);  // the unit's first row: the smallest limit
                    h_cur = g.h; rows_cur = g.rows; row_cur = g.row_g;
                    m_run = EXPF ? kMaxInit : kMaxInit * kLog2e;
                    l_run = 0.0f;
                }
            }
            float alpha = 1.0f;
            if (have_tile) {
                const int sb = gi % SBUF;
                const uint32_t s_tmem = tmem_base + lane_off + Cfg::kSP + sb * 2 * BN;
                mbar_wait(smem_u32(&s_full[sb]), (gi / SBUF) & 1);
                tc_fence_after();
                if (gi == 0 && tid == 0) stamp(2);
                stamp_tile(gi, 0);
                float sv[BN];
#pragma unroll
                for (int c = 0; c < BN; c += 32) tmem_ld32(s_tmem + c, sv + c);
                tmem_wait_ld();
                // Only tiles that touch the window start or the causal limit of the unit's first row (the smallest) pay
                // for masking -- a pass of its own over the scores (a bit mask of the keys the row sees, one test + select
                // per key) rather than a second copy of the whole tile body, which would be fetched cold on the diagonal.
                if (g0 < kv_start || g0 + BN > lim_min) {
#pragma unroll
                    for (int w = 0; w < BN / 32; ++w) {
                        const uint32_t mask = bit_range32(kv_start - g0 - 32 * w, lim - g0 - 32 * w);
#pragma unroll
                        for (int i = 0; i < 32; ++i) sv[32 * w + i] = (mask & (1u << i)) ? sv[32 * w + i] : -INFINITY;
                    }
                }
                // Two softmax flavours (EXPF, chosen on the host):
                //  true : s = (q.k) * scale, expf(s - max) as the reference writes it (paged_infer.c:197-208);
                //  false: the same in the exp2 domain, one FMA + ex2.approx per key (relative error 2^-22 plus the rounding
                //         of an argument of magnitude < 30: <= 2e-6 on keys whose weight is 2^-30, <= 4e-7 on keys that
                //         matter) -- the default.
                // Either way the running max starts at the reference's -10000 (m_run is kept in the domain in use).
                const float kscale = EXPF ? p.scale : p.sl2;
                // (four independent chains for the row maximum and the row sum: with one warp per scheduler a 64-deep
                // dependent chain of 4-cycle operations is 256 exposed cycles per tile)
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    float s = sv[i];
                    if (EXPF) s *= kscale;
                    sv[i] = s;
                    mx4[i & 3] = fmaxf(mx4[i & 3], s);
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                const float m_new = fmaxf(m_run, EXPF ? mx : mx * kscale);     // (the scale is positive: max of the raw scores)
                alpha = EXPF ? expf(m_run - m_new) : ex2_approx(m_run - m_new);      // 1 when the maximum did not move
                const float neg_m = -m_new;
                float ps4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < BN; c += 16) {               // 16 columns at a time: half the live registers of a 32-wide store
                    float ph[16], pl[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e = EXPF ? expf(sv[c + i] - m_new) : ex2_approx(fmaf(sv[c + i], kscale, neg_m));   // masked keys: exp(-inf) = 0
                        ps4[i & 3] += e;
                        ph[i] = tf32_hi(e);
                        pl[i] = e - ph[i];
                    }
                    tmem_st16(s_tmem + c, ph);                      // P_hi
                    tmem_st16(s_tmem + BN + c, pl);                 // P_lo
                }
                const float psum = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
                l_run = l_run * alpha + psum;
                m_run = m_new;
                g0 += BN;
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&p_ready[sb]));
                if (gi == 0 && tid == 0) stamp(3);
                stamp_tile(gi, 1);
            }
            if (gi > 0) {
                // fold key tile gi - 1's P.V result into the register accumulator: o = o * alpha(gi - 1) + O_tile(gi - 1);
                // it has had this tile's softmax to complete
                const int ob = (gi - 1) % OBUF;
                mbar_wait(smem_u32(&o_full[ob]), ((gi - 1) / OBUF) & 1);
                tc_fence_after();
                const uint32_t o_tmem = tmem_base + lane_off + Cfg::kO + ob * HS;
                // 64 columns per round trip (two loads in flight, ONE wait): four serialised TMEM round trips per tile
                // were a tenth of the tile's time
                constexpr int FC = HS == 64 ? 64 : 32;          // (head_dim 128: o[] alone is 128 registers)
#pragma unroll
                for (int c = 0; c < HS; c += FC) {
                    float ov[FC];
#pragma unroll
                    for (int cc = 0; cc < FC; cc += 32) tmem_ld32(o_tmem + c + cc, ov + cc);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < FC; ++i) o[c + i] = fmaf(o[c + i], alpha_pend, ov[i]);
                }
                tc_fence_before();
                mbar_arrive(smem_u32(&o_free[ob]));
                if (fin_due) {
                    // Normalise and store the finished unit's rows, and clear the accumulator.  Through the staging tile,
                    // 64 columns of the warp's own 32 rows at a time: each thread writes its row, the warp reads two whole
                    // rows per instruction and stores them as four full lines (a thread storing its own row straight to
                    // global memory touches 32 lines per instruction: measured 2 us per Q tile).  The staging tile is
                    // free: the Q of the unit in progress has left it (its first S has been seen), and the splitter
                    // fetches the Q after that only once o_stored says this output is out.
                    fin_due = false;
                    const float inv = (fin_l == 0.0f) ? 0.0f : 1.0f / fin_l;      // :213
                    float* dst = p.out + (size_t)(fin_row + wq * 32 + (lane >> 4)) * p.out_stride + fin_h * HS + (lane & 15) * 4;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            sts128(stg + stage_off(r, c), make_float4(o[ch * 64 + 4 * c] * inv, o[ch * 64 + 4 * c + 1] * inv,
                                                                      o[ch * 64 + 4 * c + 2] * inv, o[ch * 64 + 4 * c + 3] * inv));
                        __syncwarp();
                        float4 v[16];                       // (all loads in flight, then all stores: o[] is dead here)
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = lds128(stg + stage_off(wq * 32 + 2 * k + (lane >> 4), lane & 15));
                        float* d = dst + ch * 64;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (wq * 32 + 2 * k + (lane >> 4) < fin_rows) *reinterpret_cast<float4*>(d) = v[k];
                            d += 2 * (size_t)p.out_stride;
                        }
                        __syncwarp();
                    }
                    mbar_arrive(smem_u32(o_stored));
#pragma unroll
                    for (int i = 0; i < HS; ++i) o[i] = 0.0f;
                }
                stamp_tile(gi - 1, 2);
            }
            alpha_pend = alpha;
            if (!have_tile) break;
        }
        if (tid == 0) { stamp(4); if (p.dbg) p.dbg[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)unit_end; }
    }

    tc_fence_before();
    __syncthreads();
    if (tid == 0) stamp(5);
    if (warp == kProducerWarp) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct Tc3State {
    CUtensorMap tm_k, tm_v;
    bool ready;
    // the schedule of the current step (built once per uploaded step: every layer's call reuses it)
    int2* d_units;
    int2* h_units;               // pinned
    size_t units_cap;            // in int2
    cudaEvent_t copied;          // the list has left h_units
    unsigned long sched_uploads; // h->step_uploads it was built from
    int sched_bn, sched_grid, sched_rows;
};

// pool viewed as (layer, row = page*bs + slot, column) fp32; box = one page x 32 columns, 128-byte swizzle
int make_pool_map3(CUtensorMap* map, float* pool, const pa_handle* h, CUtensorMapSwizzle swizzle) {
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

// The schedule.  Two things pull against each other:
//  - balance wants the units sorted longest first and dealt to the CTAs boustrophedon (row 0 left to right, row 1 right
//    to left, ...: what a CTA gets more than its neighbour in one row it gets less in the next) -- within 1% of the
//    greedy longest-first assignment on every shape tried;
//  - HBM traffic wants the q tiles of one (sequence, head) to run AT THE SAME TIME on neighbouring SMs: they walk the
//    same keys in step, so each K/V tile comes from HBM once and from L2 for the others.  Sorted by length, the CTAs
//    of a row hold the same q tile of different (sequence, head)s instead, share nothing, and every unit streams its
//    keys from HBM: measured 8-12% slower (the kernel runs at the power cap, and that traffic is power).
// So: the q tiles go into kBuckets classes by length, longest class first; inside a class the order is (sequence,
// head, q tile) -- heads OUTSIDE the q tiles; and the list is dealt boustrophedon.  Balance stays >= 0.97 of ideal on
// the shapes of profiles/r02_prefill.md, a ragged batch included.  Laid out [row][cta]: a CTA reads its column.
constexpr int kBuckets = 8;
// (pure host code: no CUDA call, so that the CPU test suite can check it; `flat` NULL = sizes only)
int make_schedule(const pa_handle* h, int BN, int max_ctas, int2* flat, size_t cap, int* grid_out, int* rows_out) {
    const pa_step_layout& L = h->step;
    const int NH = h->cfg.n_heads;
    *grid_out = 0; *rows_out = 0;
    if (!h->h_step || L.nseq < 1) return PA_OK;
    const int* qr = h->h_step + L.off_q_row0;
    const int* ke = h->h_step + L.off_kv_end;
    const int* ks = h->h_step + L.off_kv_start;
    struct Tile { int seq, qt, n_kt; };
    std::vector<Tile> tiles;
    for (int i = 0; i < L.nseq; ++i) {
        const int nq = qr[i + 1] - qr[i], n_qt = (nq + kBM - 1) / kBM;
        if (n_qt > 0xffff) { pa_set_error("pa_prefill: more than 65535 query tiles in one sequence"); return PA_ERR_UNSUPPORTED; }
        const int k_begin = (ks[i] / BN) * BN;
        for (int t = 0; t < n_qt; ++t) {
            const int last_row = std::min(nq, (t + 1) * kBM) - 1;
            const int lim_last = ke[i] - (nq - 1 - last_row);
            tiles.push_back({i, t, lim_last > k_begin ? (lim_last - k_begin + BN - 1) / BN : 0});
        }
    }
    const long long n_units = (long long)tiles.size() * NH;
    if (n_units == 0) return PA_OK;
    int max_kt = 1;
    for (const Tile& t : tiles) max_kt = std::max(max_kt, t.n_kt);
    auto bucket = [&](const Tile& t) { return (int)((long long)t.n_kt * kBuckets / (max_kt + 1)); };
    std::stable_sort(tiles.begin(), tiles.end(), [&](const Tile& a, const Tile& b) {
        const int ba = bucket(a), bb = bucket(b);
        if (ba != bb) return ba > bb;
        if (a.seq != b.seq) return a.seq < b.seq;
        return a.qt > b.qt; });
    const int grid = (int)std::min<long long>(n_units, std::max(1, max_ctas));
    const size_t rows = (size_t)((n_units + grid - 1) / grid), n = rows * grid;
    *grid_out = grid; *rows_out = (int)rows;
    if (!flat) return PA_OK;
    if (cap < n) { pa_set_error("pa_prefill_schedule: %zu entries needed, room for %zu", n, cap); return PA_ERR_INVALID; }
    size_t k = 0;
    for (size_t i0 = 0; i0 < tiles.size();) {
        size_t i1 = i0 + 1;                            // a run of one sequence's q tiles of one class
        while (i1 < tiles.size() && tiles[i1].seq == tiles[i0].seq && bucket(tiles[i1]) == bucket(tiles[i0])) ++i1;
        for (int hd = 0; hd < NH; ++hd)
            for (size_t i = i0; i < i1; ++i, ++k) {
                const size_t row = k / grid, pos = k % grid;
                flat[row * grid + ((row & 1) ? grid - 1 - pos : pos)] = make_int2(tiles[i].seq, tiles[i].qt | (hd << 16));
            }
        i0 = i1;
    }
    for (; k < n; ++k) {
        const size_t row = k / grid, pos = k % grid;
        flat[row * grid + ((row & 1) ? grid - 1 - pos : pos)] = make_int2(-1, 0);
    }
    return PA_OK;
}

// build the step's schedule and send it to the device (once per uploaded step)
int build_schedule(pa_handle* h, Tc3State* st, int BN, cudaStream_t s) {
    st->sched_grid = 0; st->sched_rows = 0;
    st->sched_uploads = h->step_uploads; st->sched_bn = BN;
    int grid = 0, rows = 0;
    int rc = make_schedule(h, BN, h->sm_count, nullptr, 0, &grid, &rows);
    if (rc != PA_OK || grid == 0) return rc;
    const size_t n = (size_t)rows * grid;
    if (st->units_cap < n) {
        CU_CHECK(cudaStreamSynchronize(s));            // (rare: nothing may still read the old list)
        if (st->d_units) cudaFree(st->d_units);
        if (st->h_units) cudaFreeHost(st->h_units);
        st->d_units = nullptr; st->h_units = nullptr; st->units_cap = 0;
        const size_t cap = n + n / 2 + 1024;
        CU_CHECK(cudaMalloc((void**)&st->d_units, cap * sizeof(int2)));
        CU_CHECK(cudaMallocHost((void**)&st->h_units, cap * sizeof(int2)));
        st->units_cap = cap;
    }
    if (!st->copied) CU_CHECK(cudaEventCreateWithFlags(&st->copied, cudaEventDisableTiming));
    else CU_CHECK(cudaEventSynchronize(st->copied));          // the previous step's copy has left the pinned buffer (long ago)
    rc = make_schedule(h, BN, h->sm_count, st->h_units, st->units_cap, &grid, &rows);
    if (rc != PA_OK) return rc;
    CU_CHECK(cudaMemcpyAsync(st->d_units, st->h_units, n * sizeof(int2), cudaMemcpyHostToDevice, s));
    CU_CHECK(cudaEventRecord(st->copied, s));
    st->sched_grid = grid; st->sched_rows = rows;
    return PA_OK;
}

template <int HS, int BN, int NST, int SBUF, int OBUF, bool EXPF>
int launch_tc3(const Tc3State* st, const Tc3Params& p, cudaStream_t s) {
    using Cfg = Tc3Cfg<HS, BN, NST, SBUF, OBUF>;
    auto fn = pa_prefill_tc3_kernel<HS, BN, NST, SBUF, OBUF, EXPF>;
    static std::atomic<unsigned long long> attr_done{0};       // per instantiation; one bit per device
    CU_CHECK(pa_optin_smem(attr_done, fn, (int)Cfg::kSmem));
    const unsigned n_ctas = (unsigned)st->sched_grid;
    static const bool timeline = getenv("PA_PREFILL_TC3_TIMELINE") != nullptr;
    if (timeline) {
        // developer aid: per-CTA stamps -> how full the SMs were and how long a CTA takes to get going; per key tile
        // stamps of CTA 0's softmax thread 0 (stderr; synchronises)
        Tc3Params q = p;
        const size_t n = (size_t)n_ctas * 8 + 3 * 160;
        CU_CHECK(cudaMalloc((void**)&q.dbg, n * sizeof(unsigned long long)));
        CU_CHECK(cudaMemsetAsync(q.dbg, 0, n * sizeof(unsigned long long), s));
        fn<<<n_ctas, Cfg::kThreads, Cfg::kSmem, s>>>(st->tm_k, st->tm_v, q);
        CU_CHECK(cudaGetLastError());
        CU_CHECK(cudaStreamSynchronize(s));
        std::vector<unsigned long long> hst(n);
        CU_CHECK(cudaMemcpy(hst.data(), q.dbg, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CU_CHECK(cudaFree(q.dbg));
        unsigned long long t0 = ~0ull, t1 = 0, first_end = ~0ull;
        double busy = 0, tiles = 0, first_s = 0, first_p = 0;
        for (unsigned c = 0; c < n_ctas; ++c) {
            const unsigned long long* e = &hst[(size_t)c * 8];
            t0 = std::min(t0, e[0]); t1 = std::max(t1, e[5]); first_end = std::min(first_end, e[5]);
            busy += (double)(e[5] - e[0]);
            tiles += (double)e[6];
            first_s += (double)(e[2] - e[0]); first_p += (double)(e[3] - e[0]);
        }
        fprintf(stderr, "tc3 timeline: %u CTAs x %d units, span %.1f us (first CTA done after %.1f), SM occupancy %.3f, key tiles/CTA %.1f = %.0f ns each; first S after %.0f ns, first P after %.0f\n",
                n_ctas, p.n_rows, (t1 - t0) * 1e-3, (first_end - t0) * 1e-3, busy / ((double)(t1 - t0) * n_ctas), tiles / n_ctas,
                busy / tiles, first_s / n_ctas, first_p / n_ctas);
        fprintf(stderr, "  CTA 0, softmax thread 0, ns from entry per key tile: S seen / P written / O folded\n ");
        const unsigned long long* e2 = &hst[(size_t)n_ctas * 8];
        for (int i = 0; i < 160 && e2[3 * i]; ++i)
            fprintf(stderr, " [%d] %lld %lld %lld", i, (long long)(e2[3 * i] - hst[0]), (long long)(e2[3 * i + 1] - hst[0]), (long long)(e2[3 * i + 2] - hst[0]));
        fprintf(stderr, "\n");
        return PA_OK;
    }
    fn<<<n_ctas, Cfg::kThreads, Cfg::kSmem, s>>>(st->tm_k, st->tm_v, p);
    CU_CHECK(cudaGetLastError());
    return PA_OK;
}

bool aligned16_3(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" PA_API int pa_prefill_schedule(pa_handle* h, int key_tile, int max_ctas, int* units, size_t cap, int* n_ctas, int* rows) {
    if (!h || !n_ctas || !rows || !(key_tile == 32 || key_tile == 64)) { pa_set_error("pa_prefill_schedule: bad arguments"); return PA_ERR_INVALID; }
    if (h->step.nseq < 1) { pa_set_error("pa_prefill_schedule: no step (pa_step_begin)"); return PA_ERR_INVALID; }
    if (max_ctas <= 0) max_ctas = h->sm_count > 0 ? h->sm_count : 148;
    return make_schedule(h, key_tile, max_ctas, reinterpret_cast<int2*>(units), cap, n_ctas, rows);
}

extern "C" void pa_cu_prefill_tc3_release(pa_handle* h) {
    Tc3State* st = (Tc3State*)h->tc3_state;
    if (st && st->d_units) cudaFree(st->d_units);
    if (st && st->h_units) cudaFreeHost(st->h_units);
    if (st && st->copied) cudaEventDestroy(st->copied);
    free(h->tc3_state);
    h->tc3_state = nullptr;
}

// PA_OK = launched; PA_ERR_UNSUPPORTED = outside the kernel's domain
extern "C" int pa_cu_prefill_tc3(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride,
                                 void* stream) {
    const pa_step_layout& L = h->step;
    const int hs = h->cfg.head_dim, bs = h->cfg.block_size;
    if (!(hs == 64 || hs == 128)) return PA_ERR_UNSUPPORTED;
    if (h->cfg.n_heads > 0xffff) return PA_ERR_UNSUPPORTED;
    const int BN = hs == 64 ? 64 : 32;         // four tiles (K, K_lo, V, V_lo) x 3 stages must fit 227 KB
    // a page must be whole 8-row swizzle groups and divide the key tile
    if (bs < 8 || (bs & (bs - 1)) || bs > BN) return PA_ERR_UNSUPPORTED;
    if ((h->C % 4) || (q_stride % 4) || (out_stride % 4) || !aligned16_3(q) || !aligned16_3(out)) return PA_ERR_UNSUPPORTED;
    Tc3State* st = (Tc3State*)h->tc3_state;
    if (!st) {
        void* mem = nullptr;
        if (posix_memalign(&mem, 64, sizeof(Tc3State)) != 0) { pa_set_error("out of host memory"); return PA_ERR_NOMEM; }
        st = (Tc3State*)mem;
        memset(st, 0, sizeof(*st));
        h->tc3_state = st;
    }
    if (!st->ready) {
        // K is a K-major operand (16-byte swizzle chunks), V an MN-major one (32-byte chunks)
        int rc = make_pool_map3(&st->tm_k, h->pool_k, h, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == PA_OK) rc = make_pool_map3(&st->tm_v, h->pool_v, h, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc != PA_OK) return rc;
        st->ready = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (st->sched_uploads != h->step_uploads || st->sched_bn != BN) {
        rc = build_schedule(h, st, BN, s);
        if (rc != PA_OK) { st->sched_uploads = 0; return rc; }
    }
    if (st->sched_grid == 0) return PA_OK;
    Tc3Params p;
    p.q = q; p.out = out;
    p.kv_end = h->d_step + L.off_kv_end;
    p.kv_start = h->d_step + L.off_kv_start;
    p.q_row0 = h->d_step + L.off_q_row0;
    p.table = h->d_step + L.off_table;
    p.units = st->d_units;
    p.dbg = nullptr;
    p.bs = bs;
    p.tstride = L.tstride; p.q_stride = q_stride; p.out_stride = out_stride;
    p.n_rows = st->sched_rows;
    p.layer = layer;
    p.scale = (float)(1.0 / sqrtf((float)hs));          // paged_infer.c:174
    p.sl2 = p.scale * kLog2e;
    static const bool expf_exact = getenv("PA_PREFILL_TC3_EXPF") && atoi(getenv("PA_PREFILL_TC3_EXPF")) != 0;
    // (A second softmax warpgroup alternating key tiles was built, tested and measured slower -- 125 vs 133 TFLOP/s at
    // 16 x 2048: the softmax is not what bounds the kernel, the 48 MMA instructions per key tile are -- and removed
    // when the staging tile took the shared memory its merge needed; profiles/r02_prefill.md.)
    if (hs == 64) rc = expf_exact ? launch_tc3<64, 64, 3, 2, 2, true>(st, p, s) : launch_tc3<64, 64, 3, 2, 2, false>(st, p, s);      // TMEM 128 + 256 + 128 = 512 columns
    else rc = expf_exact ? launch_tc3<128, 32, 3, 2, 1, true>(st, p, s) : launch_tc3<128, 32, 3, 2, 1, false>(st, p, s);           // TMEM 256 + 128 + 128 = 512 columns
    if (rc == PA_OK) h->launches++;
    return rc;
}
