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
 * One CTA per (head, sequence, tile of 128 query rows) = one TMEM lane per query row; 12 warps (three warpgroups):
 *   warps 0-3   softmax warpgroup: thread = query row.  Reads its row of S (tcgen05.ld), online softmax, writes
 *               P and P_lo back into TMEM as the A operands of P.V (tcgen05.st), and folds the PREVIOUS tile's
 *               P.V result into its register accumulator while the tensor core works on the current one
 *   warps 4-7   splitter warpgroup: K_lo / V_lo tiles next to the raw ones the TMA delivered (element-wise over
 *               the flat swizzled buffers: same layout, other base address)
 *   warp 8      TMA producer (+ TMEM allocation): one tensor-map box per page and 32-column block
 *   warp 9      MMA issuer (one elected lane): S = Q.K^T (A = Q, Q_lo from TMEM; B = K, K_lo from shared memory,
 *               K-major SW128) and O_tile = P.V (A = P, P_lo from TMEM; B = V, V_lo MN-major SW128/32B atoms)
 * TMEM columns: Q hs | Q_lo hs | SBUF x (S/P BN | P_lo BN) | OBUF x O_tile hs  <= 512.
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
constexpr int kBM = 128;                // query rows per CTA = TMEM lanes
constexpr float kLog2e = 1.4426950408889634f;

struct Tc3Params {
    const float* q;
    float* out;
    const int* kv_end;
    const int* kv_start;
    const int* q_row0;
    const int* table;
    int B, C, NH, bs, tstride, q_stride, out_stride;
    int n_tiles, layer;
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

// tile_lin -> (sequence, q tile): warp-parallel scan over ceil(nq/128)
__device__ __forceinline__ void find_tile3(const Tc3Params& p, int tile_lin, int& seq, int& qt, int& n_qt) {
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

template <int HS, int BN, int NST, int SBUF, int OBUF, int NWG>
struct Tc3Cfg {
    // whole warpgroups (setmaxnreg works on warpgroups): NWG x softmax | splitter | producer, MMA issuer, two idle warps
    static constexpr int kThreads = (NWG + 2) * 128;
    // registers per thread after setmaxnreg: the launch gives 65536 / kThreads (168 at 384 threads, 128 at 512)
    static constexpr int kRegSoftmax = NWG == 1 ? 232 : 184, kRegOther = NWG == 1 ? 96 : 72;
    static_assert(NWG == 1 || NWG == 2, "one or two softmax warpgroups");
    static_assert((NWG * kRegSoftmax + 2 * kRegOther) * 128 <= 65536, "register file");
    static_assert(NWG == 1 || (SBUF == NWG), "with two softmax warpgroups each owns one S/P buffer");
    static_assert((NWG - 1) * 128 * (HS + 2) * 4 <= NST * BN * HS * 4, "merge scratch must fit the K ring");
    static constexpr int kKVBytes = BN * HS * 4;
    static constexpr int kTileBytes = 4 * NST * kKVBytes;      // NST-deep rings of K, K_lo, V, V_lo tiles
    static constexpr int kNumBars = 6 * NST + 2 * SBUF + 2 * OBUF;
    static constexpr int kBarBytes = kNumBars * 8 + 32;
    static constexpr size_t kSmem = 1024 + kTileBytes + kBarBytes;     // 1024: manual alignment slack
    static constexpr int kQ = 0, kQlo = HS, kSP = 2 * HS, kO = 2 * HS + SBUF * 2 * BN;
    static constexpr int kCols = kO + OBUF * HS;
    static constexpr int kLag = SBUF - 1;                    // Q.K^T runs this many key tiles ahead of P.V
    static_assert(NST > kLag, "the K ring must hold the tiles whose Q.K^T has been issued ahead");
    static_assert(kCols <= 512, "TMEM columns");
    static_assert(kSmem <= 227 * 1024, "shared memory");
};

template <int HS, int BN, int NST, int SBUF, int OBUF, int NWG, bool EXPF>
__global__ void __launch_bounds__((NWG + 2) * 128, 1)
pa_prefill_tc3_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const Tc3Params p) {
    using Cfg = Tc3Cfg<HS, BN, NST, SBUF, OBUF, NWG>;
    constexpr int DB = HS / 32;                         // 32-column blocks per row
    constexpr uint32_t kIdescQK = instr_desc(kBM, BN, 0, 0);
    constexpr uint32_t kIdescPV = instr_desc(kBM, HS, 0, 1);

    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* Ks = base;                                           // [NST][DB][BN][128 B]
    unsigned char* Kl = Ks + NST * Cfg::kKVBytes;                       // K_lo, same layout
    unsigned char* Vs = Kl + NST * Cfg::kKVBytes;
    unsigned char* Vl = Vs + NST * Cfg::kKVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vl + NST * Cfg::kKVBytes);
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + OBUF);
    int* s_unit = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kSplitWarp0 = 4 * NWG, kProducerWarp = 4 * NWG + 4, kMmaWarp = 4 * NWG + 5;

    const int h = blockIdx.x / p.n_tiles;
    const int tile_lin = blockIdx.x - h * p.n_tiles;
    if (warp == 0) {
        int seq, qt, n_qt;
        find_tile3(p, tile_lin, seq, qt, n_qt);
        if (lane == 0) { s_unit[0] = seq; s_unit[1] = n_qt - 1 - qt; }     // heaviest q tile first
    }
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
        mbar_fence_init();
    }
    if (warp == kProducerWarp) tmem_alloc<512>(tmem_slot);
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

    // ---- Q tile -> TMEM: raw columns [0, HS) and lo columns [HS, 2 HS); thread = query row = TMEM lane.
    // The softmax and splitter warpgroups share the columns (warpgroup g stores the 32-column blocks g, g+2, ...).
    if (warp < 8 && n_kt > 0) {
        const int g = warp >> 2, wq = warp & 3;
        const int r = wq * 32 + lane;
        const bool ok = r < rows;
        const float* src = p.q + (size_t)(row0 + j0 + (ok ? r : 0)) * p.q_stride + h * HS;
#pragma unroll
        for (int c = g * 32; c < HS; c += 64) {
            float qv[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(src + c + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                qv[i] = t.x; qv[i + 1] = t.y; qv[i + 2] = t.z; qv[i + 3] = t.w;
            }
            float qh[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { qh[i] = tf32_hi(qv[i]); qv[i] -= qh[i]; }
            tmem_st32(tmem_base + ((uint32_t)(wq * 32) << 16) + Cfg::kQ + c, qh);
            tmem_st32(tmem_base + ((uint32_t)(wq * 32) << 16) + Cfg::kQlo + c, qv);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Registers follow the roles: the kernel is compiled for 168 per thread (384 threads); the splitter and the
    // producer / issuer warpgroups hand most of theirs back and the softmax warpgroup -- a query row's running output
    // (hs floats) plus a key tile of scores per thread -- takes them (setmaxnreg).
    if (warp >= kProducerWarp) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegOther));
      if (warp == kProducerWarp) {
        // ================================ TMA producer ========================================
        if (n_kt > 0) {
            const int* tbl = p.table + (size_t)seq * p.tstride;
            const int n_pages = (kv_end + p.bs - 1) / p.bs;
            const int ppt = BN / p.bs;                                   // pages per key tile
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
                    // the ring slot (raw and lo) is free once the MMA that read its previous content has completed
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
                const int st = it % NST, sb = it % SBUF, ob = it % OBUF;
                mbar_wait(smem_u32(&v_split[st]), (it / NST) & 1);           // V landed and V_lo written
                mbar_wait(smem_u32(&p_ready[sb]), (it / SBUF) & 1);
                if (it >= OBUF) mbar_wait(smem_u32(&o_free[ob]), (it / OBUF - 1) & 1);      // the tile that used this O buffer is in registers
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
            for (int it = 0; it < n_kt; ++it) {
                const int st = it % NST, sb = it % SBUF;
                mbar_wait(smem_u32(&k_split[st]), (it / NST) & 1);           // K landed and K_lo written
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
                if (it >= Cfg::kLag) issue_pv(it - Cfg::kLag);
            }
            for (int it = max(0, n_kt - Cfg::kLag); it < n_kt; ++it) issue_pv(it);
        }
      }       // (warps 10 and 11 only complete the third warpgroup)
    } else if (warp >= kSplitWarp0) {
        // ============================== splitter warpgroup ====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegOther));
        // K -> (K_hi in place, K_lo beside it), V likewise: element-wise over the flat swizzled tile
        const int t = tid - kSplitWarp0 * 32;
        for (int it = 0; it < n_kt; ++it) {
            const int st = it % NST;
            const uint32_t par = (it / NST) & 1;
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
        }
    } else {
        // ============================== softmax warpgroup(s) ===================================
        // With two warpgroups, warpgroup g owns the key tiles g, g+2, ... with its own online-softmax state (m, l, o);
        // the states are merged at the end, so the warpgroups never wait for each other.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kRegSoftmax));
        const int g = warp >> 2;                         // softmax warpgroup
        const int wq = warp & 3;                         // TMEM lane quarter of this warp
        const int r = wq * 32 + lane;                    // query row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const int lim = min(kv_end, kv_end - (nq - 1 - (j0 + r)));    // this row sees keys [kv_start, lim) (rows past the tile's last: nothing beyond the cache)
        const int lim_first = kv_end - (nq - 1 - j0);
        float m_run = EXPF ? kMaxInit : kMaxInit * kLog2e, l_run = 0.0f;
        float o[HS];
#pragma unroll
        for (int i = 0; i < HS; ++i) o[i] = 0.0f;
        float alpha_pend = 1.0f;                          // rescale that belongs to the tile whose P.V is still in flight

        // fold tile `it`'s P.V result into the register accumulator: o = o * alpha(it) + O_tile(it)
        auto take_o = [&](int it, float alpha) {
            const int ob = it % OBUF;
            mbar_wait(smem_u32(&o_full[ob]), (it / OBUF) & 1);
            tc_fence_after();
            const uint32_t o_tmem = tmem_base + lane_off + Cfg::kO + ob * HS;
            if (NWG == 1) {
                // 64 columns per round trip (two loads in flight, ONE wait): the register budget of the single softmax
                // warpgroup allows it, and four serialised TMEM round trips per tile were a tenth of the tile's time
#pragma unroll
                for (int c = 0; c < HS; c += 64) {
                    float ov[64];
                    tmem_ld32(o_tmem + c, ov);
                    tmem_ld32(o_tmem + c + 32, ov + 32);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 64; ++i) o[c + i] = fmaf(o[c + i], alpha, ov[i]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < HS; c += 16) {
                    float ov[16];
                    tmem_ld16(o_tmem + c, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[c + i] = fmaf(o[c + i], alpha, ov[i]);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&o_free[ob]));
        };

        // One key tile.  MASK is a compile-time flag and the two instances are reached through a real branch:
        // only tiles that touch the window start or a row's causal limit pay for the per-element compares.
        auto tile = [&](auto mask_tag, int it) {
            constexpr bool MASK = decltype(mask_tag)::value;
            const int g0 = k_begin + it * BN;
            const int sb = it % SBUF;
            const uint32_t s_tmem = tmem_base + lane_off + Cfg::kSP + sb * 2 * BN;
            mbar_wait(smem_u32(&s_full[sb]), (it / SBUF) & 1);
            tc_fence_after();
            float sv[BN];
#pragma unroll
            for (int c = 0; c < BN; c += 32) tmem_ld32(s_tmem + c, sv + c);
            tmem_wait_ld();
            // Two softmax flavours (EXPF, chosen on the host):
            //  true : s = (q.k) * scale, expf(s - max) as the reference writes it (paged_infer.c:197-208);
            //  false: the same in the exp2 domain, one FMA + ex2.approx per key (relative error 2^-22 plus the rounding
            //         of an argument of magnitude < 30: <= 2e-6 on keys whose weight is 2^-30, <= 4e-7 on keys that
            //         matter) -- the default: the softmax warpgroup, not the tensor core, bounds this kernel at head_dim 64.
            // Either way the running max starts at the reference's -10000 (m_run is kept in the domain in use).
            const float kscale = EXPF ? p.scale : p.sl2;
            // (four independent chains for the row maximum and the row sum: with one warp per scheduler a 64-deep
            // dependent chain of 4-cycle operations is 256 exposed cycles per tile)
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < BN; ++i) {
                float s = sv[i];
                if (MASK) {
                    const int key = g0 + i;
                    if (key < kv_start || key >= lim) s = -INFINITY;
                }
                if (EXPF) s *= kscale;
                sv[i] = s;
                mx4[i & 3] = fmaxf(mx4[i & 3], s);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const float m_new = fmaxf(m_run, EXPF ? mx : mx * kscale);     // (the scale is positive: max of the raw scores)
            const float alpha = EXPF ? expf(m_run - m_new) : ex2_approx(m_run - m_new);      // 1 when the maximum did not move
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
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&p_ready[sb]));
            // this warpgroup's previous tile's P.V has had this tile's softmax to complete
            if (it >= NWG) take_o(it - NWG, alpha_pend);
            alpha_pend = alpha;
        };
        int last_mine = -1;
        for (int it = g; it < n_kt; it += NWG) {
            const int g0 = k_begin + it * BN;
            if ((g0 < kv_start) || (g0 + BN > lim_first)) tile(std::true_type{}, it);
            else tile(std::false_type{}, it);
            last_mine = it;
        }
        if (last_mine >= 0) take_o(last_mine, alpha_pend);
        if (NWG > 1) {
            // merge the warpgroups' states: warpgroup 1 hands (o, m, l) of its rows to warpgroup 0 through the K ring
            // (every MMA has completed: each warpgroup waited for its last P.V, and the tensor pipe runs in order)
            asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
            float* scratch = reinterpret_cast<float*>(Ks);               // [128][HS + 2]
            if (g == 1) {
                float* dst = scratch + r * (HS + 2);
#pragma unroll
                for (int i = 0; i < HS; ++i) dst[i] = o[i];
                dst[HS] = m_run;
                dst[HS + 1] = l_run;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NWG * 128) : "memory");
            if (g == 0) {
                const float* src = scratch + r * (HS + 2);
                const float m1 = src[HS], l1 = src[HS + 1];
                const float m = fmaxf(m_run, m1);
                const float w0 = EXPF ? expf(m_run - m) : ex2_approx(m_run - m), w1 = EXPF ? expf(m1 - m) : ex2_approx(m1 - m);
                l_run = l_run * w0 + l1 * w1;
#pragma unroll
                for (int i = 0; i < HS; ++i) o[i] = o[i] * w0 + src[i] * w1;
                m_run = m;
            }
        }
        if (g == 0 && r < rows && seq >= 0) {
            const float inv = (l_run == 0.0f) ? 0.0f : 1.0f / l_run;      // :213
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
        tmem_dealloc<512>(tmem_base);
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct Tc3State {
    CUtensorMap tm_k, tm_v;
    bool ready;
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

template <int HS, int BN, int NST, int SBUF, int OBUF, int NWG, bool EXPF>
int launch_tc3(const Tc3State* st, const Tc3Params& p, cudaStream_t s) {
    using Cfg = Tc3Cfg<HS, BN, NST, SBUF, OBUF, NWG>;
    auto fn = pa_prefill_tc3_kernel<HS, BN, NST, SBUF, OBUF, NWG, EXPF>;
    static std::atomic<unsigned long long> attr_done{0};       // per instantiation; one bit per device
    CU_CHECK(pa_optin_smem(attr_done, fn, (int)Cfg::kSmem));
    fn<<<(unsigned)((long long)p.n_tiles * p.NH), Cfg::kThreads, Cfg::kSmem, s>>>(st->tm_k, st->tm_v, p);
    CU_CHECK(cudaGetLastError());
    return PA_OK;
}

bool aligned16_3(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" void pa_cu_prefill_tc3_release(pa_handle* h) {
    free(h->tc3_state);
    h->tc3_state = nullptr;
}

// PA_OK = launched; PA_ERR_UNSUPPORTED = outside the kernel's domain
extern "C" int pa_cu_prefill_tc3(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride,
                                 void* stream) {
    const pa_step_layout& L = h->step;
    const int hs = h->cfg.head_dim, bs = h->cfg.block_size;
    if (!(hs == 64 || hs == 128)) return PA_ERR_UNSUPPORTED;
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
    Tc3Params p;
    p.q = q; p.out = out;
    p.kv_end = h->d_step + L.off_kv_end;
    p.kv_start = h->d_step + L.off_kv_start;
    p.q_row0 = h->d_step + L.off_q_row0;
    p.table = h->d_step + L.off_table;
    p.B = L.nseq; p.C = h->C; p.NH = h->cfg.n_heads; p.bs = bs;
    p.tstride = L.tstride; p.q_stride = q_stride; p.out_stride = out_stride;
    p.layer = layer;
    p.scale = (float)(1.0 / sqrtf((float)hs));          // paged_infer.c:174
    p.sl2 = p.scale * kLog2e;
    long long n_tiles = 0;
    const int* qr = h->h_step + L.off_q_row0;
    for (int i = 0; i < L.nseq; ++i) n_tiles += (qr[i + 1] - qr[i] + kBM - 1) / kBM;
    if (n_tiles == 0) return PA_OK;
    if (n_tiles * p.NH > 0x7fffffffLL) return PA_ERR_UNSUPPORTED;
    p.n_tiles = (int)n_tiles;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    static const bool expf_exact = getenv("PA_PREFILL_TC3_EXPF") && atoi(getenv("PA_PREFILL_TC3_EXPF")) != 0;
    // One softmax warpgroup.  A second one alternating key tiles (each with its own S/P and O buffers; head_dim 64 only:
    // head_dim 128 has ONE O buffer in TMEM, which would serialise them) is built and tested
    // (PA_PREFILL_TC3_WARPGROUPS=2) but measured SLOWER: 125 vs 133 TFLOP/s at 16 x 2048, 159 vs 166 at 16 x 4096 -- the
    // softmax is not what bounds the kernel; the 48 N=64 MMA instructions per key tile are (a 128x64x8 tf32 instruction
    // takes ~63 cycles, twice its ideal: measured in pa_gemm_tc.cu), and the extra warpgroup costs registers.
    static const int nwg64 = getenv("PA_PREFILL_TC3_WARPGROUPS") ? atoi(getenv("PA_PREFILL_TC3_WARPGROUPS")) : 1;
    if (hs == 64 && nwg64 == 2) rc = expf_exact ? launch_tc3<64, 64, 3, 2, 2, 2, true>(st, p, s) : launch_tc3<64, 64, 3, 2, 2, 2, false>(st, p, s);
    else if (hs == 64) rc = expf_exact ? launch_tc3<64, 64, 3, 2, 2, 1, true>(st, p, s) : launch_tc3<64, 64, 3, 2, 2, 1, false>(st, p, s);      // TMEM 128 + 256 + 128 = 512 columns
    else rc = expf_exact ? launch_tc3<128, 32, 3, 2, 1, 1, true>(st, p, s) : launch_tc3<128, 32, 3, 2, 1, 1, false>(st, p, s);                  // TMEM 256 + 128 + 128 = 512 columns
    if (rc == PA_OK) h->launches++;
    return rc;
}
