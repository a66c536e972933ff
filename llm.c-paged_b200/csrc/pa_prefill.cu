/*
 * pa_prefill.cu -- causal multi-row paged attention (prompt prefill, chunked prefill and the
 * T-row window of the reference's attention_paged), fp32 SIMT, tiled ("flash" style).
 *
 *   pa_prefill_tiled_kernel<HS, BM, BN>
 *       one CTA per (head, sequence, tile of BM query rows); K/V tiles of BN keys are gathered
 *       through the block table into shared memory with 16-byte cp.async (double buffered),
 *       S = Q K^T and O += P V are register-tiled fp32 FMA contractions, the softmax is the
 *       online form of paged_infer.c:187-236 (running max from -10000, sum==0 -> 0, expf).
 *
 * Semantics: rows t of attention_paged (paged_infer.c:163-240); query row j of a sequence with
 * nq new rows sees cached tokens [kv_start, kv_end - (nq-1-j)).  Any block size (the gather is
 * per key row), head_dim 64 or 128; everything else stays on pa_attn_rows_kernel.
 *
 * This is the default (fp32, tolerance 1e-5) prefill path.  The tensor-core variant lives in
 * pa_prefill_tc.cu and has its own stated tolerance.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

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

struct PrefillParams {
    const float* pool_k;    // layer base, rows of C floats, row = page*bs + slot
    const float* pool_v;
    const float* q;         // packed query rows (step order), head h at +h*HS
    float* out;
    const int* kv_end;      // [B] keys visible to the LAST query row of the sequence
    const int* kv_start;    // [B]
    const int* q_row0;      // [B+1] first packed query row of each sequence (NULL: one row each)
    const int* table;       // [B][tstride]
    int B, C, NH, bs, tstride, q_stride, out_stride;
    int n_tiles;            // sum over sequences of ceil(nq / BM)
    float scale;
};

// tile_lin -> (sequence, q tile inside it): warp-parallel scan over ceil(nq/BM)
template <int BM>
__device__ __forceinline__ void find_tile(const PrefillParams& p, int tile_lin, int& seq, int& qt, int& n_qt) {
    const int lane = threadIdx.x & 31;
    int run = 0;
    seq = -1; qt = 0; n_qt = 0;
    for (int c = 0; c < p.B; c += 32) {
        const int i = c + lane;
        int n = 0;
        if (i < p.B) {
            const int nq = p.q_row0 ? p.q_row0[i + 1] - p.q_row0[i] : 1;
            n = (nq + BM - 1) / BM;
        }
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

template <int HS, int BM, int BN>
struct PrefillCfg {
    static constexpr int kThreads = 256;
    static constexpr int TX = BN / 4;                 // threads across the keys of a tile (4 keys each)
    static constexpr int TY = kThreads / TX;          // thread rows
    static constexpr int RM = BM / TY;                // query rows per thread
    static constexpr int DC = HS / (4 * TX);          // float4 output column groups per thread
    static constexpr int LDQ = HS + 4;                // padded row strides (floats): conflict-free LDS.128
    static constexpr int LDP = BN + 4;
    static constexpr int kQFloats = BM * LDQ;
    static constexpr int kKVFloats = BN * LDQ;
    static constexpr int kPFloats = BM * LDP;
    static constexpr size_t kSmem = (size_t)(kQFloats + 4 * kKVFloats + kPFloats) * sizeof(float);
    static_assert(TX == 16, "a query row's threads must be one half-warp");
    static_assert(BM % TY == 0 && HS % (4 * TX) == 0, "tile shape");
};

template <int HS, int BM, int BN>
__global__ void __launch_bounds__(256, 1)
pa_prefill_tiled_kernel(const PrefillParams p) {
    using Cfg = PrefillCfg<HS, BM, BN>;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, RM = Cfg::RM, DC = Cfg::DC, LDQ = Cfg::LDQ, LDP = Cfg::LDP;
    constexpr int HS4 = HS / 4;

    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                               // [BM][LDQ]
    float* Ks = Qs + Cfg::kQFloats;                 // [2][BN][LDQ]
    float* Vs = Ks + 2 * Cfg::kKVFloats;            // [2][BN][LDQ]
    float* Ps = Vs + 2 * Cfg::kKVFloats;            // [BM][LDP]
    __shared__ int s_unit[3];

    // heads are the slowest index: CTAs resident together work on the same head's K/V columns,
    // and the heaviest q tile of a sequence (the last one) is scheduled first
    const int h = blockIdx.x / p.n_tiles;
    const int tile_lin = blockIdx.x - h * p.n_tiles;
    if (threadIdx.x < 32) {
        int seq, qt, n_qt;
        find_tile<BM>(p, tile_lin, seq, qt, n_qt);
        if (threadIdx.x == 0) { s_unit[0] = seq; s_unit[1] = n_qt - 1 - qt; s_unit[2] = n_qt; }
    }
    __syncthreads();
    const int seq = s_unit[0];
    if (seq < 0) return;
    const int qt = s_unit[1];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int row0 = p.q_row0 ? p.q_row0[seq] : seq;
    const int nq = p.q_row0 ? p.q_row0[seq + 1] - row0 : 1;
    const int kv_start = p.kv_start[seq];
    const int kv_end = p.kv_end[seq];
    const int* tbl = p.table + (size_t)seq * p.tstride;
    const int j0 = qt * BM;                                   // first query row of the tile
    const int rows = min(BM, nq - j0);
    // row j sees keys [kv_start, kv_end - (nq-1-j)); the tile's last row sees the most
    const int lim_last = kv_end - (nq - 1 - (j0 + rows - 1));
    const int k_begin = (kv_start / BN) * BN;
    const int n_kt = lim_last > k_begin ? (lim_last - k_begin + BN - 1) / BN : 0;

    // ---- async loads -----------------------------------------------------------------------
    auto load_q = [&]() {
        for (int i = tid; i < BM * HS4; i += Cfg::kThreads) {
            const int r = i / HS4, c4 = i - r * HS4;
            const bool ok = r < rows;
            const float* src = p.q + (size_t)(row0 + j0 + (ok ? r : 0)) * p.q_stride + h * HS + c4 * 4;
            cp_async16(smem_u32(Qs + r * LDQ + c4 * 4), src, ok ? 16 : 0);
        }
    };
    auto load_kv = [&](int kt, int buf) {
        const int g0 = k_begin + kt * BN;
        float* kd = Ks + buf * Cfg::kKVFloats;
        float* vd = Vs + buf * Cfg::kKVFloats;
        for (int i = tid; i < BN * HS4; i += Cfg::kThreads) {
            const int r = i / HS4, c4 = i - r * HS4;
            const int g = g0 + r;
            const bool ok = g < kv_end;                       // rows past the sequence are zero-filled
            const int gg = ok ? g : 0;
            const int pg = gg / p.bs;
            const size_t off = ((size_t)__ldg(tbl + pg) * p.bs + (gg - pg * p.bs)) * p.C + h * HS + c4 * 4;
            cp_async16(smem_u32(kd + r * LDQ + c4 * 4), p.pool_k + off, ok ? 16 : 0);
            cp_async16(smem_u32(vd + r * LDQ + c4 * 4), p.pool_v + off, ok ? 16 : 0);
        }
    };

    float o[RM][DC][4];
    float m_run[RM], l_run[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        m_run[i] = kMaxInit;
        l_run[i] = 0.0f;
#pragma unroll
        for (int c = 0; c < DC; ++c) o[i][c][0] = o[i][c][1] = o[i][c][2] = o[i][c][3] = 0.0f;
    }

    load_q();
    if (n_kt > 0) load_kv(0, 0);
    cp_async_commit();

    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) load_kv(kt + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const float* Kb = Ks + buf * Cfg::kKVFloats;
        const float* Vb = Vs + buf * Cfg::kKVFloats;
        // ---- S = Q K^T : rows ty + TY*i, keys tx + TX*j ------------------------------------
        float s[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
#pragma unroll 4
        for (int d = 0; d < HS; d += 4) {
            float4 kf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) kf[j] = *reinterpret_cast<const float4*>(Kb + (tx + TX * j) * LDQ + d);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                const float4 qf = *reinterpret_cast<const float4*>(Qs + (ty + TY * i) * LDQ + d);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qf.x, kf[j].x, s[i][j]);
                    s[i][j] = fmaf(qf.y, kf[j].y, s[i][j]);
                    s[i][j] = fmaf(qf.z, kf[j].z, s[i][j]);
                    s[i][j] = fmaf(qf.w, kf[j].w, s[i][j]);
                }
            }
        }
        // ---- online softmax per row --------------------------------------------------------
        const int g0 = k_begin + kt * BN;
        // masking is needed only on tiles that touch the window start or some row's causal limit
        const int lim_first = kv_end - (nq - 1 - j0);
        const bool need_mask = (g0 < kv_start) || (g0 + BN > lim_first);
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int r = ty + TY * i;
            const int lim = kv_end - (nq - 1 - (j0 + r));    // rows >= `rows` get lim > lim_last: harmless, never stored
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = s[i][j] * p.scale;
                if (need_mask) {
                    const int g = g0 + tx + TX * j;
                    if (g < kv_start || g >= lim) v = -INFINITY;
                }
                s[i][j] = v;
                mx = fmaxf(mx, v);
            }
#pragma unroll
            for (int d = TX / 2; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            const float m_new = fmaxf(m_run[i], mx);
            const float alpha = expf(m_run[i] - m_new);
            float psum = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = expf(s[i][j] - m_new);
                Ps[r * LDP + tx + TX * j] = e;
                psum += e;
            }
            l_run[i] = l_run[i] * alpha + psum;              // per-thread partial of the row sum
            m_run[i] = m_new;
#pragma unroll
            for (int c = 0; c < DC; ++c) {
                o[i][c][0] *= alpha; o[i][c][1] *= alpha; o[i][c][2] *= alpha; o[i][c][3] *= alpha;
            }
        }
        __syncwarp();        // a row's probabilities are written and read by the same half-warp
        // ---- O += P V : rows ty + TY*i, columns tx*4 + 64*c ---------------------------------
#pragma unroll 2
        for (int n = 0; n < BN; n += 4) {
            float4 pf[RM];
#pragma unroll
            for (int i = 0; i < RM; ++i) pf[i] = *reinterpret_cast<const float4*>(Ps + (ty + TY * i) * LDP + n);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int c = 0; c < DC; ++c) {
                    const float4 vf = *reinterpret_cast<const float4*>(Vb + (n + jj) * LDQ + tx * 4 + 4 * TX * c);
#pragma unroll
                    for (int i = 0; i < RM; ++i) {
                        const float pw = jj == 0 ? pf[i].x : (jj == 1 ? pf[i].y : (jj == 2 ? pf[i].z : pf[i].w));
                        o[i][c][0] = fmaf(pw, vf.x, o[i][c][0]);
                        o[i][c][1] = fmaf(pw, vf.y, o[i][c][1]);
                        o[i][c][2] = fmaf(pw, vf.z, o[i][c][2]);
                        o[i][c][3] = fmaf(pw, vf.w, o[i][c][3]);
                    }
                }
            }
        }
        __syncthreads();     // buffer `buf` is refilled by the loads issued next iteration
    }
    cp_async_wait<0>();

    // ---- epilogue: finish the row sums, normalise, store ------------------------------------
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int r = ty + TY * i;
        float l = l_run[i];
#pragma unroll
        for (int d = TX / 2; d >= 1; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
        const float inv = (l == 0.0f) ? 0.0f : 1.0f / l;
        if (r < rows) {
            float* dst = p.out + (size_t)(row0 + j0 + r) * p.out_stride + h * HS + tx * 4;
#pragma unroll
            for (int c = 0; c < DC; ++c)
                *reinterpret_cast<float4*>(dst + 4 * TX * c) =
                    make_float4(o[i][c][0] * inv, o[i][c][1] * inv, o[i][c][2] * inv, o[i][c][3] * inv);
        }
    }
}

template <int HS, int BM, int BN>
int launch_tiled(const PrefillParams& pp, int n_tiles, cudaStream_t s) {
    using Cfg = PrefillCfg<HS, BM, BN>;
    auto fn = pa_prefill_tiled_kernel<HS, BM, BN>;
    static std::atomic<unsigned long long> attr_done{0};       // per instantiation; one bit per device
    CU_CHECK(pa_optin_smem(attr_done, fn, (int)Cfg::kSmem));
    PrefillParams p = pp;
    p.n_tiles = n_tiles;
    fn<<<(unsigned)((long long)n_tiles * p.NH), Cfg::kThreads, Cfg::kSmem, s>>>(p);
    CU_CHECK(cudaGetLastError());
    return PA_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Returns PA_OK when the tiled kernel was launched, PA_ERR_UNSUPPORTED when the shape is outside
// its domain (the caller then uses the generic rows kernel), another error otherwise.
extern "C" int pa_cu_prefill_tiled(pa_handle* h, int layer, const float* q, int q_stride, float* out,
                                   int out_stride, int all_new_rows, void* stream) {
    const pa_step_layout& L = h->step;
    const int hs = h->cfg.head_dim;
    if (!(hs == 64 || hs == 128)) return PA_ERR_UNSUPPORTED;
    if ((h->C % 4) || (q_stride % 4) || (out_stride % 4) || !aligned16(q) || !aligned16(out)) return PA_ERR_UNSUPPORTED;
    PrefillParams pp;
    pp.pool_k = h->pool_k + (size_t)layer * h->layer_stride;
    pp.pool_v = h->pool_v + (size_t)layer * h->layer_stride;
    pp.q = q; pp.out = out;
    pp.kv_end = h->d_step + L.off_kv_end;
    pp.kv_start = h->d_step + L.off_kv_start;
    pp.q_row0 = all_new_rows ? h->d_step + L.off_q_row0 : nullptr;
    pp.table = h->d_step + L.off_table;
    pp.B = L.nseq; pp.C = h->C; pp.NH = h->cfg.n_heads; pp.bs = h->cfg.block_size;
    pp.tstride = L.tstride; pp.q_stride = q_stride; pp.out_stride = out_stride;
    pp.scale = (float)(1.0 / sqrtf((float)hs));          // paged_infer.c:174
    pp.n_tiles = 0;
    const int BM = hs == 64 ? 128 : 64;
    // q tiles from the host copy of the step tables (still the current step)
    long long n_tiles = 0;
    if (all_new_rows) {
        const int* qr = h->h_step + L.off_q_row0;
        for (int i = 0; i < L.nseq; ++i) n_tiles += (qr[i + 1] - qr[i] + BM - 1) / BM;
    } else {
        n_tiles = L.nseq;
    }
    if (n_tiles == 0) return PA_OK;
    if (n_tiles * pp.NH > 0x7fffffffLL) return PA_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = hs == 64 ? launch_tiled<64, 128, 64>(pp, (int)n_tiles, s) : launch_tiled<128, 64, 64>(pp, (int)n_tiles, s);
    if (rc == PA_OK) h->launches++;
    return rc;
}
