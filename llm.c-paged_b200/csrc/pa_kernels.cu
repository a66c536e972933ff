/*
 * pa_kernels.cu -- hand-written sm_100a kernels of the paged KV-cache path and their launchers.
 *
 *   pa_append_kernel          KV append: scatter new K/V rows into their page slots
 *                             (semantics: add_to_cache copy loop, paged_infer.c:548-566)
 *   pa_decode_stream_kernel   paged decode attention, HBM-bound streaming design:
 *                             flat page stream split evenly over a persistent grid ("stream-K over
 *                             pages"), K/V page tiles staged in shared memory by the TMA bulk-copy
 *                             engine (cp.async.bulk + mbarrier ring, one producer warp), fp32 SIMT
 *                             QK^T / online softmax / PV with warp-shuffle reductions, split
 *                             partials merged in-kernel by the last-arriving CTA
 *                             (semantics: row t=T-1 of attention_paged, paged_infer.c:182-236)
 *   pa_attn_rows_kernel       generic fp32 SIMT causal attention through the block table: any
 *                             head size / block size / alignment, any number of query rows per
 *                             sequence (prefill; fallback decode)
 *                             (semantics: attention_paged, paged_infer.c:163-240)
 *
 * Arithmetic notes for parity: running max starts at -10000.0f
 * (paged_infer.c:187); 1/sum with sum==0 -> 0 (:213); scale = (float)(1.0/sqrtf(hs)) is computed
 * on the host (:174); expf, never __expf; no fast-math.
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

// =============================================================================================
// KV append
// =============================================================================================
// One CTA row per new token; threads copy 16-byte chunks of the K row and the V row to
// pool[slot][0..C).  slot = page*block_size + row_in_page comes from the mirrored slot mapping.
template <bool kVec4>
__global__ void __launch_bounds__(256)
pa_append_kernel(const float* __restrict__ k_src, const float* __restrict__ v_src, int src_stride,
                 float* __restrict__ pool_k, float* __restrict__ pool_v,
                 const int* __restrict__ slot_mapping, int n_tokens, int C) {
    for (int tok = blockIdx.y; tok < n_tokens; tok += gridDim.y) {
        const size_t dst = (size_t)slot_mapping[tok] * C;
        const size_t src = (size_t)tok * src_stride;
        if (kVec4) {
            const int n4 = C >> 2;
            const float4* ks = reinterpret_cast<const float4*>(k_src + src);
            const float4* vs = reinterpret_cast<const float4*>(v_src + src);
            float4* kd = reinterpret_cast<float4*>(pool_k + dst);
            float4* vd = reinterpret_cast<float4*>(pool_v + dst);
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
                kd[i] = __ldg(ks + i);
                vd[i] = __ldg(vs + i);
            }
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) {
                pool_k[dst + i] = k_src[src + i];
                pool_v[dst + i] = v_src[src + i];
            }
        }
    }
}


// =============================================================================================
// decode: flat page-stream kernel
// =============================================================================================
// Work decomposition ("stream-K over pages"): the pages of the whole batch, times the head
// groups, form one flat list of U units.  The list is cut into RANGES, each owned by one virtual
// CTA: the first Us units into Gs equal static ranges (range c is the first job of physical CTA
// c), the tail U-Us into D small ranges that the physical CTAs claim with an atomic counter as
// they run dry (tapered self-scheduling: SMs do not all stream at the same speed, and a static
// split makes everyone wait for the slowest).  A range may start and end anywhere, also in the
// middle of a sequence; every (sequence, head group) piece inside a range is a SEGMENT that
// yields a partial (m, l, o).  Because both the range index and the row index grow
// monotonically along the list, `range + hg*B + row` is a unique partial slot.
struct DecodeParams {
    float* pool_k;            // layer base, [max_blocks][BS][C]
    float* pool_v;
    const float* q;           // (B, q_stride)
    float* out;               // (B, out_stride)
    const float* k_new;       // fused append: this step's K row of sequence i at k_new + i*new_stride
    const float* v_new;       //               (NULL: the rows are already in the pool)
    const int* kv_end;        // [B]
    const int* kv_start;      // [B]
    const int* cum_pages;     // [B+1]
    const int* table;         // [B][tstride]
    float* ws;                // partial slots
    int* counters;            // [n_hg*B] arrival counters (self-resetting)
    int* sched;               // [0] next dynamic range, [1] finished CTAs (self-resetting)
    unsigned long long* dbg;  // optional per-CTA timeline (8 words per CTA), NULL normally
    long long U;              // work units = n_hg * P
    long long Us;             // units in the static part
    int Gs;                   // static ranges (= physical CTAs that get one)
    int D;                    // dynamic ranges
    int B, C, hpg, n_hg, W;   // W = hpg*HS floats per tile row
    int tstride, q_stride, out_stride, new_stride;
    int P;                    // total pages of the batch
    int n_stages;
    int n_cons;               // consumer threads (multiple of 32)
    int slot_floats;          // floats per partial slot
    int n_cum_smem;           // B+1 when the prefix sums are staged in shared memory, else 0
    float scale;
};

constexpr int kFlagFirst = 1;   // first page of a (sequence, head-group) segment in this range
constexpr int kFlagLast = 2;    // last page of the segment in this range
constexpr int kFlagNew = 4;     // last page of the sequence and its last row is this step's token
constexpr int kFlagEnd = 8;     // no more work for this CTA
constexpr int kSegQ = 4;        // depth of the consumer -> merger queue of finished segments

// Transposed butterfly: N per-token partial sums per lane, reduced over the 2*D lanes that share
// a head.  Each step the lanes trade half of their values, so the whole reduction costs ~N
// shuffles instead of N*log2(lanes).  Afterwards the lane holds max(1, N0/lanes) finished sums,
// the first of them for token `tok`.
template <int N, int D>
struct XReduce {
    template <int BS>
    static __device__ __forceinline__ void run(float (&v)[BS], int gl, int& tok) {
        if constexpr (D >= 1) {
            if constexpr (N > 1) {
                const bool upper = (gl & D) != 0;
#pragma unroll
                for (int i = 0; i < N / 2; ++i) {
                    const float send = upper ? v[i] : v[i + N / 2];
                    const float keep = upper ? v[i + N / 2] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, D);
                }
                tok += upper ? N / 2 : 0;
                XReduce<N / 2, D / 2>::run(v, gl, tok);
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], D);
                XReduce<1, D / 2>::run(v, gl, tok);
            }
        }
    }
};

template <int LPH>
__device__ __forceinline__ float group_max(float x) {
#pragma unroll
    for (int d = LPH / 2; d >= 1; d >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, d));
    return x;
}
template <int LPH>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
    for (int d = LPH / 2; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// n units dealt to g ranges as [c*n/g, (c+1)*n/g): the range that owns unit x
__device__ __forceinline__ int even_owner(long long x, long long n, int g) {
    return (int)(((x + 1) * (long long)g - 1) / n);
}
__device__ __forceinline__ int range_of_unit(const DecodeParams& p, long long x) {
    if (x < p.Us) return even_owner(x, p.Us, p.Gs);
    return p.Gs + even_owner(x - p.Us, p.U - p.Us, p.D);
}
__device__ __forceinline__ void range_bounds(const DecodeParams& p, int r, long long& lo, long long& hi) {
    if (r < p.Gs) {
        lo = (long long)r * p.Us / p.Gs;
        hi = (long long)(r + 1) * p.Us / p.Gs;
    } else {
        const long long n = p.U - p.Us;
        lo = p.Us + (long long)(r - p.Gs) * n / p.D;
        hi = p.Us + (long long)(r - p.Gs + 1) * n / p.D;
    }
}

template <int HS, int BS>
__global__ void __launch_bounds__(320, 1)
pa_decode_stream_kernel(const DecodeParams p) {
    constexpr int LPH = HS / 4;                       // lanes per head (one float4 column each)
    constexpr int NV = (BS >= LPH) ? BS / LPH : 1;    // finished scores per lane after the reduce
    constexpr int REP = (BS >= LPH) ? 1 : LPH / BS;   // lanes holding the same token's score
    static_assert(LPH == 16 || LPH == 32, "head size 64 or 128");
    static_assert(BS >= 4 && BS <= 32 && (BS & (BS - 1)) == 0, "block size 4..32, power of two");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = p.W;
    const int W4 = W >> 2;
    const int tile_floats = BS * W;
    float* tiles = reinterpret_cast<float*>(smem_raw);                          // [n_stages][BS*W]
    float* qbuf = tiles + (size_t)p.n_stages * tile_floats;                      // [n_stages][W]
    float* psm = qbuf + (size_t)p.n_stages * W;                                  // [n_cons/LPH][BS]
    int4* meta = reinterpret_cast<int4*>(psm + (p.n_cons / LPH) * BS);           // [n_stages][2]
    int4* segdesc = meta + 2 * p.n_stages;                                       // [kSegQ][2] finished-segment queue
    uint64_t* bars = reinterpret_cast<uint64_t*>(segdesc + 2 * kSegQ);           // full[], empty[], segq_full[], segq_empty[]
    uint64_t* segq = bars + 2 * p.n_stages;
    int* cum_sm = reinterpret_cast<int*>(segq + 2 * kSegQ);                      // [n_cum_smem] (+ kv_start, kv_end)

    const int tid = threadIdx.x;
    const int n_cons_warps = p.n_cons >> 5;
    unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 8 : nullptr;

    if (tid == 0) {
        if (dbg) dbg[0] = global_ns();
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                          // full: producer's expect_tx arrive
            mbar_init(smem_u32(&bars[p.n_stages + s]), n_cons_warps);  // empty: one arrive per consumer warp
        }
        for (int s = 0; s < kSegQ; ++s) {
            mbar_init(smem_u32(&segq[s]), n_cons_warps);               // segment posted by every consumer warp
            mbar_init(smem_u32(&segq[kSegQ + s]), 1);                  // descriptor taken by the merger
        }
        mbar_fence_init();
    }
    // Let the next kernel in the stream start its own prologue as SMs drain (programmatic
    // dependent launch); it blocks in griddepcontrol.wait until this grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // The step tables were mirrored by a copy that precedes the previous kernel in stream order,
    // so they may be read before the dependency wait: stage the per-sequence prefix sums and
    // bounds in shared memory (one round trip instead of a dependent chain per lookup).
    for (int i = tid; i < p.n_cum_smem; i += blockDim.x) cum_sm[i] = __ldg(p.cum_pages + i);
    if (p.n_cum_smem) {
        for (int i = tid; i < p.B; i += blockDim.x) {
            cum_sm[p.n_cum_smem + i] = __ldg(p.kv_start + i);
            cum_sm[p.n_cum_smem + p.B + i] = __ldg(p.kv_end + i);
        }
    }
    __syncthreads();
    const int* cum = p.n_cum_smem ? cum_sm : p.cum_pages;
    const int* kvs = p.n_cum_smem ? cum_sm + p.n_cum_smem : p.kv_start;
    const int* kve = p.n_cum_smem ? cum_sm + p.n_cum_smem + p.B : p.kv_end;

    if (tid >= p.n_cons && tid < p.n_cons + 32) {
        // =================================== producer warp ===================================
        // Walks this CTA's ranges batch by batch (32 pages per batch, one page per lane).  The
        // lookups of the NEXT batch (block ids) and the claim of the NEXT dynamic range are
        // issued before the current batch's copies, so their latency hides behind the ring.
        const int lane = tid & 31;
        const bool leader = elect_one();                // issues every copy and barrier operation
        int stage = 0;
        uint32_t phase = 0;
        long long t_empty = 0;
        bool first_issue = true;
        const int n_ranges = p.Gs + p.D;

        struct Batch { int range, base, cnt, row, hg, blk, lo, hi, flags, r_first, nsegs; };
        auto resolve = [&](int range, int base) {
            Batch bt;
            bt.range = range; bt.base = base; bt.cnt = 0;
            bt.row = bt.hg = bt.blk = bt.lo = bt.hi = bt.flags = bt.r_first = 0; bt.nsegs = 1;
            if (range >= n_ranges) return bt;
            long long u_begin, u_end;
            range_bounds(p, range, u_begin, u_end);
            const int n_units = (int)(u_end - u_begin);
            bt.cnt = max(0, min(32, n_units - base));
            if (base + lane < n_units) {
                const long long u = u_begin + base + lane;
                bt.hg = (int)(u / p.P);
                const int f = (int)(u - (long long)bt.hg * p.P);
                int a = 0, b = p.B;                   // largest row with cum[row] <= f
                while (b - a > 1) {
                    const int mid = (a + b) >> 1;
                    if (cum[mid] <= f) a = mid; else b = mid;
                }
                bt.row = a;
                const int cum0 = cum[a], cum1 = cum[a + 1];
                const int start = kvs[a], end = kve[a];
                const int pg = start / BS + (f - cum0);
                bt.blk = __ldg(p.table + (size_t)a * p.tstride + pg);
                bt.lo = (f == cum0) ? start % BS : 0;
                bt.hi = min(BS, end - pg * BS);
                if (f == cum0 || base + lane == 0) bt.flags |= kFlagFirst;
                if (f == cum1 - 1 || base + lane == n_units - 1) bt.flags |= kFlagLast;
                if (f == cum1 - 1 && p.k_new != nullptr) bt.flags |= kFlagNew;
                const long long seg0 = (long long)bt.hg * p.P + cum0;
                const long long seg1 = (long long)bt.hg * p.P + cum1 - 1;
                bt.r_first = range_of_unit(p, seg0);
                bt.nsegs = range_of_unit(p, seg1) - bt.r_first + 1;
            }
            return bt;
        };
        // where the batch after (range, base) starts; consumes the reserve claim at a range end
        int claim = 0;                                  // lane 0: one dynamic range held in reserve
        auto advance = [&](int range, int base, int& n_range, int& n_base) {
            long long u_begin, u_end;
            range_bounds(p, range, u_begin, u_end);
            if (base + 32 < (int)(u_end - u_begin)) { n_range = range; n_base = base + 32; return; }
            n_range = p.Gs + __shfl_sync(0xffffffffu, claim, 0);
            n_base = 0;
            if (n_range < n_ranges && lane == 0) claim = atomicAdd(p.sched, 1);   // refill the reserve
        };

        Batch cur = resolve(blockIdx.x, 0);             // first job: this CTA's static range
        // q / k_new / v_new / out belong to the previous kernel until it has completed
        // (and so do the scheduler words, the arrival counters and the partial slots)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (lane == 0) claim = atomicAdd(p.sched, 1);
        while (cur.range < n_ranges) {
            int n_range, n_base;
            advance(cur.range, cur.base, n_range, n_base);
            const Batch nxt = resolve(n_range, n_base);     // loads in flight during the copies below
            for (int j = 0; j < cur.cnt; ++j) {
                const int j_row = __shfl_sync(0xffffffffu, cur.row, j);
                const int j_hg = __shfl_sync(0xffffffffu, cur.hg, j);
                const int j_blk = __shfl_sync(0xffffffffu, cur.blk, j);
                const int j_lo = __shfl_sync(0xffffffffu, cur.lo, j);
                const int j_hi = __shfl_sync(0xffffffffu, cur.hi, j);
                const int j_flags = __shfl_sync(0xffffffffu, cur.flags, j);
                const int j_rfirst = __shfl_sync(0xffffffffu, cur.r_first, j);
                const int j_nsegs = __shfl_sync(0xffffffffu, cur.nsegs, j);
                const size_t page_off = (size_t)j_blk * BS * p.C + (size_t)j_hg * W;
                const bool first = (j_flags & kFlagFirst) != 0;
                // rows that come from the pool; with a fused append the sequence's newest row
                // comes straight from this step's k/v rows instead
                const int pool_rows = (j_flags & kFlagNew) ? j_hi - 1 : j_hi;
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    const uint32_t full = smem_u32(&bars[stage]);
                    const uint32_t dst = smem_u32(tiles + (size_t)stage * tile_floats);
                    const float* src = (kv == 0 ? p.pool_k : p.pool_v) + page_off;
                    if (leader) {
                        if (dbg) {
                            if (first_issue) { dbg[1] = global_ns(); first_issue = false; }
                            const long long c0 = clock64();
                            mbar_wait(smem_u32(&bars[p.n_stages + stage]), phase ^ 1);
                            t_empty += clock64() - c0;
                        } else {
                            mbar_wait(smem_u32(&bars[p.n_stages + stage]), phase ^ 1);   // slot free
                        }
                        if (kv == 0) {
                            meta[2 * stage] = make_int4(j_row, j_hg, j_lo | (j_hi << 8), j_flags);
                            meta[2 * stage + 1] = make_int4(j_rfirst, j_nsegs, cur.range, j_blk);
                        }
                        uint32_t bytes = (uint32_t)j_hi * W * 4u;
                        if (kv == 0 && first) bytes += W * 4u;
                        mbar_arrive_expect_tx(full, bytes);
                    }
                    __syncwarp();
                    if (W == p.C) {            // whole rows: the valid part of the page is contiguous
                        if (leader && pool_rows > 0) tma_bulk_g2s(dst, src, (uint32_t)pool_rows * W * 4u, full);
                    } else {                   // a column slice: one bulk copy per row
                        for (int r = 0; r < pool_rows; ++r)
                            if (leader) tma_bulk_g2s(dst + r * W * 4u, src + (size_t)r * p.C, W * 4u, full);
                    }
                    if ((j_flags & kFlagNew) && leader)
                        tma_bulk_g2s(dst + (uint32_t)(j_hi - 1) * W * 4u,
                                     (kv == 0 ? p.k_new : p.v_new) + (size_t)j_row * p.new_stride + (size_t)j_hg * W,
                                     W * 4u, full);
                    if (kv == 0 && first && leader)
                        tma_bulk_g2s(smem_u32(qbuf + (size_t)stage * W),
                                     p.q + (size_t)j_row * p.q_stride + (size_t)j_hg * W, W * 4u, full);
                    __syncwarp();
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
            }
            cur = nxt;
        }
        if (leader) {
            // tell the consumers there is nothing more, then take part in resetting the counters
            mbar_wait(smem_u32(&bars[p.n_stages + stage]), phase ^ 1);
            meta[2 * stage] = make_int4(0, 0, 0, kFlagEnd);
            mbar_arrive(smem_u32(&bars[stage]));
            if (dbg) dbg[4] = (unsigned long long)t_empty;
            const int done = atomicAdd(p.sched + 1, 1);
            if (done == (int)gridDim.x - 1) {      // every CTA has made its last (failed) claim
                p.sched[0] = 0;
                p.sched[1] = 0;
            }
        }
        return;
    }

    if (tid >= p.n_cons + 32) {
        // ==================================== merger warp ====================================
        // Takes finished segments off the queue so the consumers never wait for a round trip:
        // publishes the partial (fence + arrival count) and, when it is the last of its
        // (sequence, head group), merges all partials into the output row.
        const int lane = tid & 31;
        int qs = 0;
        uint32_t qph = 0;
        for (;;) {
            mbar_wait(smem_u32(&segq[qs]), qph);
            const int4 d0 = segdesc[2 * qs];
            const int4 d1 = segdesc[2 * qs + 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&segq[kSegQ + qs]));
            if (++qs == kSegQ) { qs = 0; qph ^= 1; }
            if (d1.y) break;                                   // end of this CTA's work
            const int row = d0.x, hg = d0.y, r_first = d0.z, nsegs = d0.w;
            int last = 0;
            if (lane == 0) {
                __threadfence();                               // partial stores before the arrival count
                last = atomicAdd(p.counters + hg * p.B + row, 1) == nsegs - 1;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (!last) continue;
            __threadfence();
            const float* base = p.ws + (size_t)(r_first + hg * p.B + row) * p.slot_floats;
            // Online merge, one pass: every lane owns float4 columns lane, lane+32, ...; they are
            // handled four at a time, and the (m, l, o) of four partials are loaded together, so
            // a typical 3-4 way split costs two round trips to L2.
            constexpr int kGrp = 4;
            for (int c0 = 0; c0 < W4; c0 += 32 * kGrp) {
                float4 o[kGrp];
                float Ls[kGrp], M[kGrp];
#pragma unroll
                for (int k = 0; k < kGrp; ++k) { o[k] = make_float4(0.f, 0.f, 0.f, 0.f); Ls[k] = 0.0f; M[k] = kMaxInit; }
                for (int i0 = 0; i0 < nsegs; i0 += 4) {
                    float2 ml[4][kGrp];
                    float4 oi[4][kGrp];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float* sl = base + (size_t)min(i0 + u, nsegs - 1) * p.slot_floats;
#pragma unroll
                        for (int k = 0; k < kGrp; ++k) {
                            const int c = min(c0 + lane + 32 * k, W4 - 1);
                            ml[u][k] = __ldcg(reinterpret_cast<const float2*>(sl + W) + c / LPH);
                            oi[u][k] = __ldcg(reinterpret_cast<const float4*>(sl) + c);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (i0 + u < nsegs) {
#pragma unroll
                            for (int k = 0; k < kGrp; ++k) {
                                const float m_new = fmaxf(M[k], ml[u][k].x);
                                const float so = expf(M[k] - m_new), w = expf(ml[u][k].x - m_new);
                                Ls[k] = fmaf(ml[u][k].y, w, Ls[k] * so);
                                o[k].x = fmaf(oi[u][k].x, w, o[k].x * so); o[k].y = fmaf(oi[u][k].y, w, o[k].y * so);
                                o[k].z = fmaf(oi[u][k].z, w, o[k].z * so); o[k].w = fmaf(oi[u][k].w, w, o[k].w * so);
                                M[k] = m_new;
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < kGrp; ++k) {
                    const int c = c0 + lane + 32 * k;
                    if (c < W4) {
                        const float inv = (Ls[k] == 0.0f) ? 0.0f : 1.0f / Ls[k];
                        *reinterpret_cast<float4*>(p.out + (size_t)row * p.out_stride + (size_t)hg * W + c * 4) =
                            make_float4(o[k].x * inv, o[k].y * inv, o[k].z * inv, o[k].w * inv);
                    }
                }
            }
            if (lane == 0) p.counters[hg * p.B + row] = 0;     // ready for the next launch
        }
        return;
    }

    // ======================================= consumers =======================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // rows of the batch that own no page at all produce zeros (paged_infer.c never hits this:
    // T >= 1 always sees at least its own key)
    for (int r = blockIdx.x; r < p.B; r += gridDim.x) {
        if (cum[r + 1] == cum[r])
            for (int c = tid; c < p.C; c += p.n_cons) p.out[(size_t)r * p.out_stride + c] = 0.0f;
    }

    const int lane = tid & 31;
    const bool col_valid = tid < W4;
    const int c4 = col_valid ? tid : W4 - 1;     // padded lanes shadow the last column
    const int hl = tid / LPH;                    // head within the tile
        const int gl = tid % LPH;                    // lane within the head
    float* my_p = psm + hl * BS;

    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float m_run = kMaxInit, l_run = 0.0f;
    int stage = 0;
    uint32_t phase = 0;
    long long t_full = 0, t_seg = 0;
    unsigned long long n_tiles = 0;
    int qs = 0;                                  // finished-segment queue position
    uint32_t qph = 0;

    for (;;) {
        // ------------------------------- K tile: scores -------------------------------------
        if (dbg && tid == 0) {
            const long long c0 = clock64();
            mbar_wait(smem_u32(&bars[stage]), phase);
            t_full += clock64() - c0;
            if (n_tiles == 0) dbg[2] = global_ns();
            n_tiles += 2;
        } else {
            mbar_wait(smem_u32(&bars[stage]), phase);
        }
        const int4 mt0 = meta[2 * stage];
        const int4 mt1 = meta[2 * stage + 1];
        const int lo = mt0.z & 0xff, hi = (mt0.z >> 8) & 0xff, flags = mt0.w;
        if (flags & kFlagEnd) break;
        if (flags & kFlagFirst) {
            qv = reinterpret_cast<const float4*>(qbuf + (size_t)stage * W)[c4];
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            m_run = kMaxInit;
            l_run = 0.0f;
        }
        // fused append: the newest row sits in the tile (copied from this step's k row); store it
        // to its page slot -- nobody reads it from the pool during this launch
        const size_t new_off = ((size_t)mt1.w * BS + (hi - 1)) * p.C + (size_t)mt0.y * W + c4 * 4;
        float part[BS];
        {
            const float4* kt = reinterpret_cast<const float4*>(tiles + (size_t)stage * tile_floats) + c4;
#pragma unroll
            for (int t = 0; t < BS; ++t) {
                const float4 k4 = kt[t * W4];
                part[t] = fmaf(qv.w, k4.w, fmaf(qv.z, k4.z, fmaf(qv.y, k4.y, qv.x * k4.x)));
            }
            if ((flags & kFlagNew) && col_valid)
                *reinterpret_cast<float4*>(p.pool_k + new_off) = kt[(hi - 1) * W4];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[p.n_stages + stage]));     // K slot free again
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }

        int tok = 0;
        XReduce<BS, LPH / 2>::run(part, gl, tok);
        float s[NV];
        float mx = -INFINITY;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int t = tok + v;
            s[v] = (t >= lo && t < hi) ? part[v] * p.scale : -INFINITY;
            mx = fmaxf(mx, s[v]);
        }
        mx = group_max<LPH>(mx);
        const float m_new = fmaxf(m_run, mx);
        const float alpha = expf(m_run - m_new);
        float psum = 0.0f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float e = expf(s[v] - m_new);
            my_p[tok + v] = e;
            psum += e;
        }
        if (REP > 1 && (gl % REP) != 0) psum = 0.0f;     // replicated lanes count once
        psum = group_sum<LPH>(psum);
        l_run = l_run * alpha + psum;
        m_run = m_new;
        acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
        __syncwarp();                                    // my_p visible to the head's lanes

        // ------------------------------- V tile: weighted sum -------------------------------
        if (dbg && tid == 0) {
            const long long c0 = clock64();
            mbar_wait(smem_u32(&bars[stage]), phase);
            t_full += clock64() - c0;
        } else {
            mbar_wait(smem_u32(&bars[stage]), phase);
        }
        {
            const float4* vt = reinterpret_cast<const float4*>(tiles + (size_t)stage * tile_floats) + c4;
            const float4* p4 = reinterpret_cast<const float4*>(my_p);
#pragma unroll
            for (int t4 = 0; t4 < BS / 4; ++t4) {
                const float4 pw = p4[t4];
                const float pj[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = t4 * 4 + j;
                    if (t < hi) {            // rows >= hi were never written by the copy
                        const float4 v4 = vt[t * W4];
                        acc.x = fmaf(pj[j], v4.x, acc.x);
                        acc.y = fmaf(pj[j], v4.y, acc.y);
                        acc.z = fmaf(pj[j], v4.z, acc.z);
                        acc.w = fmaf(pj[j], v4.w, acc.w);
                    }
                }
            }
            if ((flags & kFlagNew) && col_valid)
                *reinterpret_cast<float4*>(p.pool_v + new_off) = vt[(hi - 1) * W4];
        }
        __syncwarp();                                    // everyone done with my_p and the V tile
        if (lane == 0) mbar_arrive(smem_u32(&bars[p.n_stages + stage]));
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }

        // ------------------------------- end of a segment -----------------------------------
        if (flags & kFlagLast) {
            const long long c_seg = dbg ? clock64() : 0;
            const int row = mt0.x, hg = mt0.y, r_first = mt1.x, nsegs = mt1.y, range = mt1.z;
            float* out_ptr = p.out + (size_t)row * p.out_stride + (size_t)hg * W + c4 * 4;
            if (nsegs == 1) {
                const float inv = (l_run == 0.0f) ? 0.0f : 1.0f / l_run;
                if (col_valid)
                    *reinterpret_cast<float4*>(out_ptr) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
            } else {
                // partial (m, l, unnormalised o) to this range's slot, then hand the segment to the
                // merger warp and carry on with the next tile
                mbar_wait(smem_u32(&segq[kSegQ + qs]), qph ^ 1);          // queue slot free
                float* slot = p.ws + (size_t)(range + hg * p.B + row) * p.slot_floats;
                if (col_valid) {
                    __stcg(reinterpret_cast<float4*>(slot) + c4, acc);
                    if (gl == 0) __stcg(reinterpret_cast<float2*>(slot + W) + hl, make_float2(m_run, l_run));
                }
                if (tid == 0) {
                    segdesc[2 * qs] = make_int4(row, hg, r_first, nsegs);
                    segdesc[2 * qs + 1] = make_int4(range, 0, 0, 0);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&segq[qs]));
                if (++qs == kSegQ) { qs = 0; qph ^= 1; }
            }
            if (dbg) t_seg += clock64() - c_seg;
        }
    }
    // tell the merger warp to finish
    mbar_wait(smem_u32(&segq[kSegQ + qs]), qph ^ 1);
    if (tid == 0) segdesc[2 * qs + 1] = make_int4(0, 1, 0, 0);
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&segq[qs]));
    if (dbg && tid == 0) {
        dbg[3] = global_ns();
        dbg[5] = (unsigned long long)t_full;
        dbg[6] = (unsigned long long)t_seg;
        dbg[7] = n_tiles;
    }
}

// =============================================================================================
// generic causal rows (prefill; fallback decode): one warp per (query row, head)
// =============================================================================================
struct RowsParams {
    const float* pool_k;
    const float* pool_v;
    const float* q;
    float* out;
    const int* kv_end;      // [B] keys visible to the LAST query row of the sequence
    const int* kv_start;    // [B]
    const int* q_row0;      // [B+1] first packed query row of each sequence (NULL: one row per sequence)
    const int* table;
    int B, C, NH, hs, bs, tstride, q_stride, out_stride;
    int n_rows;             // total query rows
    float scale;
};

constexpr int kMaxHsPerLane = 8;   // hs <= 256

__global__ void __launch_bounds__(128)
pa_attn_rows_kernel(const RowsParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.n_rows * p.NH) return;
    const int row = warp / p.NH;
    const int h = warp - row * p.NH;
    // sequence of this row
    int seq = row, j = 0, nq = 1;
    if (p.q_row0) {
        int a = 0, b = p.B;
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (p.q_row0[mid] <= row) a = mid; else b = mid;
        }
        seq = a;
        j = row - p.q_row0[seq];
        nq = p.q_row0[seq + 1] - p.q_row0[seq];
    }
    const int first = p.kv_start[seq];
    const int last = p.kv_end[seq] - (nq - 1 - j);     // row j sees keys [first, last)
    const int* tbl = p.table + (size_t)seq * p.tstride;
    const float* qh = p.q + (size_t)row * p.q_stride + h * p.hs;
    float* oh = p.out + (size_t)row * p.out_stride + h * p.hs;

    float o[kMaxHsPerLane];
#pragma unroll
    for (int i = 0; i < kMaxHsPerLane; ++i) o[i] = 0.0f;
    float m_run = kMaxInit, l_run = 0.0f;

    for (int g0 = first; g0 < last; g0 += 32) {
        const int g = g0 + lane;
        float s = -INFINITY;
        const float* vrow = nullptr;
        if (g < last) {
            const size_t off = ((size_t)tbl[g / p.bs] * p.bs + (g % p.bs)) * p.C + h * p.hs;
            const float* krow = p.pool_k + off;
            vrow = p.pool_v + off;
            // the reference's evaluation order, unfused (paged_infer.c:193-197): with large
            // logits one ulp of the score already moves the softmax weights visibly
            float dot = 0.0f;
            for (int i = 0; i < p.hs; ++i) dot = __fadd_rn(dot, __fmul_rn(qh[i], krow[i]));
            s = __fmul_rn(dot, p.scale);
        }
        float mx = s;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        const float m_new = fmaxf(m_run, mx);
        const float alpha = expf(m_run - m_new);
        const float e = expf(s - m_new);
        float esum = e;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, d);
        l_run = l_run * alpha + esum;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < kMaxHsPerLane; ++i) o[i] *= alpha;
        const int cnt = min(32, last - g0);
        for (int t = 0; t < cnt; ++t) {
            const float et = __shfl_sync(0xffffffffu, e, t);
            const float* vt = reinterpret_cast<const float*>(
                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(vrow), t));
#pragma unroll
            for (int i = 0; i < kMaxHsPerLane; ++i) {
                const int d = lane + 32 * i;
                if (d < p.hs) o[i] = fmaf(et, vt[d], o[i]);
            }
        }
    }
    const float inv = (l_run == 0.0f) ? 0.0f : 1.0f / l_run;
#pragma unroll
    for (int i = 0; i < kMaxHsPerLane; ++i) {
        const int d = lane + 32 * i;
        if (d < p.hs) oh[d] = o[i] * inv;
    }
}

// =============================================================================================
// launch helpers
// =============================================================================================
typedef void (*decode_fn_t)(const DecodeParams);

decode_fn_t pick_decode(int hs, int bs) {
#define PA_CASE(H, Bk) if (hs == H && bs == Bk) return pa_decode_stream_kernel<H, Bk>;
    PA_CASE(64, 4) PA_CASE(64, 8) PA_CASE(64, 16) PA_CASE(64, 32)
    PA_CASE(128, 4) PA_CASE(128, 8) PA_CASE(128, 16) PA_CASE(128, 32)
#undef PA_CASE
    return nullptr;
}

struct DecodePlan {
    int hpg, n_hg, W, n_stages, n_cons, grid, slot_floats;
    int Gs, D, n_cum_smem;
    long long U, Us;
    size_t smem;
};

constexpr int kMaxCumSmem = 1024;   // prefix sums + bounds staged in shared memory up to this many rows

size_t decode_smem_bytes(int bs, int W, int n_stages, int n_cons, int lph, int n_cum) {
    size_t b = (size_t)n_stages * bs * W * 4;          // tiles
    b += (size_t)n_stages * W * 4;                     // q slots
    b += (size_t)(n_cons / lph) * bs * 4;              // probabilities
    b += (size_t)(n_stages + kSegQ) * 2 * sizeof(int4);      // tile meta + finished-segment queue
    b += (size_t)(n_stages + kSegQ) * 2 * sizeof(uint64_t);  // barriers
    b += (size_t)n_cum * 3 * 4;                        // prefix sums of pages, kv_start, kv_end
    return b;
}

// Choose heads per tile, ring depth and CTAs per SM for this step.  Measured on B200 (round 1,
// profiles/r01_tile_sweep.txt):
//  * a CTA turns over one (K,V) tile pair per ~0.9 us whatever the tile width (the consumer warps
//    work on a tile in lock-step), so narrow tiles need several CTAs per SM to keep up with HBM:
//    pairs >= 80 KB -> 1 CTA/SM, >= 36 KB -> 2, >= 18 KB -> 3, else 4;
//  * 3 stages for tiles >= 40 KB, else 4: more bulk copies in flight cost DRAM locality
//    (3 x 48 KB beat 4 x 48 KB; 2 CTAs x 4 x 20 KB beat 1 CTA x 8 x 20 KB by 1.5x);
//  * the widest tile that still gives every CTA slot a unit of work; when the batch is too
//    small for that, the narrowest tile (most parallelism).
bool plan_decode(const pa_handle* h, int total_pages, int B, DecodePlan* plan) {
    const int hs = h->cfg.head_dim, bs = h->cfg.block_size, NH = h->cfg.n_heads;
    const int lph = hs / 4;
    const int n_cum = (B + 1 <= kMaxCumSmem) ? B + 1 : 0;
    const int want_hpg = h->tune[PA_TUNE_HEADS_PER_TILE];
    const int want_stages = h->tune[PA_TUNE_STAGES];
    const int smem_sm = h->smem_per_sm > 0 ? h->smem_per_sm : h->smem_optin + 1024;
    int best = 0, best_stages = 0, best_per_sm = 1;
    for (int hpg = NH; hpg >= 1; --hpg) {
        if (NH % hpg) continue;
        if (want_hpg > 0 && hpg != want_hpg) continue;
        const int W = hpg * hs;
        const int n_cons = ((W / 4) + 31) & ~31;
        if (n_cons > 256) continue;
        const size_t tile = (size_t)bs * W * 4;
        const size_t per_stage = tile + (size_t)W * 4 + 2 * sizeof(int4) + 16;
        const size_t fixed = (size_t)(n_cons / lph) * bs * 4 + 64 + kSegQ * 48 + (size_t)n_cum * 12;
        int per_sm = 2 * tile >= 80 * 1024 ? 1 : (2 * tile >= 36 * 1024 ? 2 : (2 * tile >= 18 * 1024 ? 3 : 4));
        const int regs_limit = 65536 / (168 * (n_cons + 64));            // register file
        if (per_sm > regs_limit) per_sm = regs_limit > 0 ? regs_limit : 1;
        int stages = 0;
        for (; per_sm >= 1; --per_sm) {
            size_t budget = (size_t)smem_sm / per_sm - 1024;
            if (budget > (size_t)h->smem_optin) budget = h->smem_optin;
            if (budget <= fixed) continue;
            stages = (int)((budget - fixed) / per_stage);
            const int cap = want_stages > 0 ? want_stages : (tile >= 40 * 1024 ? 3 : 4);
            if (stages > cap) stages = cap;
            if (stages >= 2) break;
        }
        if (stages < 2) continue;
        best = hpg;
        best_stages = stages;
        best_per_sm = per_sm;
        if ((long long)(NH / hpg) * total_pages >= (long long)h->sm_count * per_sm) break;   // enough units: keep the wide tile
    }
    if (best == 0) return false;
    plan->hpg = best;
    plan->n_hg = NH / best;
    plan->W = best * hs;
    plan->n_cons = ((plan->W / 4) + 31) & ~31;
    plan->n_stages = best_stages;
    plan->n_cum_smem = n_cum;
    plan->smem = decode_smem_bytes(bs, plan->W, best_stages, plan->n_cons, lph, n_cum);
    const long long units = (long long)plan->n_hg * total_pages;
    long long grid = (long long)h->sm_count * best_per_sm;
    if (h->tune[PA_TUNE_GRID] > 0) grid = h->tune[PA_TUNE_GRID];
    if (grid > (long long)h->sm_count * 8) grid = (long long)h->sm_count * 8;
    if (grid > units) grid = units;
    if (grid < 1) grid = 1;
    plan->grid = (int)grid;
    plan->slot_floats = (plan->W + 2 * plan->hpg + 3) & ~3;
    // static part first, a tail of small ranges claimed dynamically
    plan->U = units;
    plan->Gs = (int)grid;
    plan->Us = units;
    plan->D = 0;
    // default: all static.  The dynamic tail balances the SMs' streaming time, but every range is
    // one more partial to merge, and the merges of the last sequences sit on the critical path
    // (measured: slower than the static split for every setting tried; kept as a knob).
    int pct = h->tune[PA_TUNE_STATIC_PCT] > 0 ? h->tune[PA_TUNE_STATIC_PCT] : 100;
    int dyn_units = h->tune[PA_TUNE_DYN_UNITS] > 0 ? h->tune[PA_TUNE_DYN_UNITS] : 2;
    if (pct < 100 && units >= 4 * grid) {
        long long us = units * pct / 100;
        long long d = (units - us + dyn_units - 1) / dyn_units;
        const long long d_cap = (long long)h->sm_count * 15;
        if (d > d_cap) d = d_cap;
        if (d >= 1) { plan->Us = us; plan->D = (int)d; }
    }
    return true;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_rows(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride,
                bool all_new_rows, cudaStream_t s) {
    const pa_step_layout& L = h->step;
    if (h->cfg.head_dim > 32 * kMaxHsPerLane) {
        pa_set_error("head_dim %d > %d unsupported", h->cfg.head_dim, 32 * kMaxHsPerLane);
        return PA_ERR_UNSUPPORTED;
    }
    RowsParams rp;
    rp.pool_k = h->pool_k + (size_t)layer * h->layer_stride;
    rp.pool_v = h->pool_v + (size_t)layer * h->layer_stride;
    rp.q = q; rp.out = out;
    rp.kv_end = h->d_step + L.off_kv_end;
    rp.kv_start = h->d_step + L.off_kv_start;
    rp.q_row0 = all_new_rows ? h->d_step + L.off_q_row0 : nullptr;
    rp.table = h->d_step + L.off_table;
    rp.B = L.nseq; rp.C = h->C; rp.NH = h->cfg.n_heads; rp.hs = h->cfg.head_dim; rp.bs = h->cfg.block_size;
    rp.tstride = L.tstride; rp.q_stride = q_stride; rp.out_stride = out_stride;
    rp.n_rows = all_new_rows ? L.ntok : L.nseq;
    rp.scale = (float)(1.0 / sqrtf((float)h->cfg.head_dim));
    if (rp.n_rows == 0) return PA_OK;
    const long long warps = (long long)rp.n_rows * rp.NH;
    const int threads = 128;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    pa_attn_rows_kernel<<<(unsigned)blocks, threads, 0, s>>>(rp);
    CU_CHECK(cudaGetLastError());
    h->launches++;
    return PA_OK;
}

// =============================================================================================
// decode, small work: one CTA per (sequence, head)
// =============================================================================================
// A handful of sequences (the reference's own decode loop is batch 1) is a latency problem, not a
// bandwidth one: the stream kernel's persistent grid, page scheduler and cross-CTA merge cost ~10 us
// before the first byte matters.  Here the 16 warps of ONE CTA share a (sequence, head): chunks of
// ceil(len / 16) tokens, hs/4 lanes per token (16-byte loads through the block table), online
// softmax from the reference's -10000 start, the 16 partial (o, m, l) merged through shared memory
// in chunk order.  With k_new/v_new the step's new token is read from there and its head slice is
// stored to the page slot by this CTA (fused append).  Same arithmetic as the persistent step
// kernel's attention (pa_model_mega.cu).
struct SmallParams {
    float* pool_k;
    float* pool_v;
    const float* q;
    const float* k_new;      // fused append (or NULL)
    const float* v_new;
    float* out;
    const int* kv_end;
    const int* kv_start;
    const int* slots;        // [B] slot of the new token (fused append)
    const int* table;
    int C, NH, hs, bs, tstride, q_stride, out_stride, new_stride;
    float scale;
};
constexpr int kSmallWarps = 16;
template <int LPT>
__global__ void __launch_bounds__(kSmallWarps * 32)
pa_decode_small_kernel(const SmallParams p) {
    constexpr int TPI = 32 / LPT, UN = 4;
    __shared__ __align__(16) float ps[kSmallWarps][LPT * 4 + 4];
    pdl_launch_dependents();
    pdl_wait();               // q, the step tables and the pool may belong to the previous kernel of the stream
    const int h = blockIdx.x, s = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LPT, li = lane % LPT;
    const int hs = p.hs, C = p.C;
    const int first = p.kv_start[s], last = p.kv_end[s];
    const int chunk = (last - first + kSmallWarps - 1) / kSmallWarps;
    const int t0 = first + warp * chunk, t1 = min(last, t0 + chunk);
    const int* tbl = p.table + (size_t)s * p.tstride;
    const float4 q4 = *(reinterpret_cast<const float4*>(p.q + (size_t)s * p.q_stride + h * hs) + li);
    // the new token's K/V head slice: from the step's rows, and to its page slot
    const float4* kn = p.k_new ? reinterpret_cast<const float4*>(p.k_new + (size_t)s * p.new_stride + h * hs) : nullptr;
    const float4* vn = p.k_new ? reinterpret_cast<const float4*>(p.v_new + (size_t)s * p.new_stride + h * hs) : nullptr;
    if (kn && threadIdx.x < LPT) {
        const size_t off = (size_t)p.slots[s] * C + h * hs;
        reinterpret_cast<float4*>(p.pool_k + off)[threadIdx.x] = kn[threadIdx.x];
        reinterpret_cast<float4*>(p.pool_v + off)[threadIdx.x] = vn[threadIdx.x];
    }
    float m_run = kMaxInit, l_run = 0.0f;
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tb = t0; tb < t1; tb += TPI * UN) {
        float4 k4[UN], v4[UN];
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int t = tb + i * TPI + sub;
            if (t < t1) {
                if (kn && t == last - 1) { k4[i] = kn[li]; v4[i] = vn[li]; }
                else {
                    const size_t off = ((size_t)tbl[t / p.bs] * p.bs + (t % p.bs)) * C + h * hs;
                    k4[i] = __ldg(reinterpret_cast<const float4*>(p.pool_k + off) + li);
                    v4[i] = __ldg(reinterpret_cast<const float4*>(p.pool_v + off) + li);
                }
            } else {
                k4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                v4[i] = k4[i];
            }
        }
        float sc[UN];
        float m_new = m_run;
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            float dot = fmaf(q4.w, k4[i].w, fmaf(q4.z, k4[i].z, fmaf(q4.y, k4[i].y, q4.x * k4[i].x)));
#pragma unroll
            for (int d = LPT / 2; d >= 1; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
            sc[i] = (tb + i * TPI + sub < t1) ? dot * p.scale : -INFINITY;
            m_new = fmaxf(m_new, sc[i]);
        }
        const float alpha = expf(m_run - m_new);
        l_run *= alpha; o4.x *= alpha; o4.y *= alpha; o4.z *= alpha; o4.w *= alpha;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const float e = expf(sc[i] - m_new);                    // exp(-inf) = 0 for the padding
            l_run += e;
            o4.x = fmaf(e, v4[i].x, o4.x); o4.y = fmaf(e, v4[i].y, o4.y); o4.z = fmaf(e, v4[i].z, o4.z); o4.w = fmaf(e, v4[i].w, o4.w);
        }
    }
    if (TPI > 1) {                                                  // the token sub-groups of the warp -> one partial
#pragma unroll
        for (int d = LPT; d < 32; d <<= 1) {
            const float m_o = __shfl_xor_sync(0xffffffffu, m_run, d), l_o = __shfl_xor_sync(0xffffffffu, l_run, d);
            float4 o_o;
            o_o.x = __shfl_xor_sync(0xffffffffu, o4.x, d); o_o.y = __shfl_xor_sync(0xffffffffu, o4.y, d);
            o_o.z = __shfl_xor_sync(0xffffffffu, o4.z, d); o_o.w = __shfl_xor_sync(0xffffffffu, o4.w, d);
            const float m_t = fmaxf(m_run, m_o);
            const float wa = expf(m_run - m_t), wb = expf(m_o - m_t);
            // (lower sub-group first in both partners, so they agree bit for bit)
            const bool lo = (lane & d) == 0;
            const float la = lo ? l_run : l_o, lb = lo ? l_o : l_run, xa = lo ? wa : wb, xb = lo ? wb : wa;
            l_run = fmaf(la, xa, lb * xb);
            o4.x = fmaf(lo ? o4.x : o_o.x, xa, (lo ? o_o.x : o4.x) * xb); o4.y = fmaf(lo ? o4.y : o_o.y, xa, (lo ? o_o.y : o4.y) * xb);
            o4.z = fmaf(lo ? o4.z : o_o.z, xa, (lo ? o_o.z : o4.z) * xb); o4.w = fmaf(lo ? o4.w : o_o.w, xa, (lo ? o_o.w : o4.w) * xb);
            m_run = m_t;
        }
    }
    if (sub == 0) {
        reinterpret_cast<float4*>(ps[warp])[li] = o4;
        if (li == 0) { ps[warp][hs] = m_run; ps[warp][hs + 1] = l_run; }
    }
    __syncthreads();
    if ((int)threadIdx.x < hs) {
        const int d = threadIdx.x;
        float m_tot = kMaxInit;
#pragma unroll
        for (int w = 0; w < kSmallWarps; ++w) m_tot = fmaxf(m_tot, ps[w][hs]);
        float l_tot = 0.0f, o = 0.0f;
#pragma unroll 4
        for (int w = 0; w < kSmallWarps; ++w) {                     // chunk order
            const float wgt = expf(ps[w][hs] - m_tot);
            l_tot = fmaf(ps[w][hs + 1], wgt, l_tot);
            o = fmaf(ps[w][d], wgt, o);
        }
        p.out[(size_t)s * p.out_stride + h * hs + d] = o * ((l_tot == 0.0f) ? 0.0f : 1.0f / l_tot);     // :213
    }
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

static int check_compute(pa_handle* h, int layer, const char* who) {
    if (!h) { pa_set_error("%s: NULL handle", who); return PA_ERR_INVALID; }
    if (h->host_only || !h->pool_k) {
        pa_set_error("%s: handle has no device (host-only); there is no CPU fallback", who);
        return PA_ERR_NO_DEVICE;
    }
    if (layer < 0 || layer >= h->cfg.n_layers) { pa_set_error("%s: layer %d out of range", who, layer); return PA_ERR_INVALID; }
    if (h->step.nseq < 1) { pa_set_error("%s: no step (call pa_step_begin)", who); return PA_ERR_INVALID; }
    if (!h->step.uploaded) { pa_set_error("%s: step tables not uploaded (call pa_step_upload)", who); return PA_ERR_INVALID; }
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) { pa_set_error("%s: cudaSetDevice failed", who); return PA_ERR_CUDA; }
    pa_pdl_enabled = h->tune[PA_TUNE_NO_PDL] ? 0 : 1;      // this thread's launches follow THIS handle's switch
    return PA_OK;
}

int pa_append(pa_handle* h, int layer, const float* k, const float* v, int row_stride, void* stream) {
    int rc = check_compute(h, layer, "pa_append");
    if (rc != PA_OK) return rc;
    const pa_step_layout& L = h->step;
    if (L.ntok == 0) return PA_OK;
    if (!k || !v || row_stride < h->C) { pa_set_error("pa_append: bad source"); return PA_ERR_INVALID; }
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)h->stream;
    float* pk = h->pool_k + (size_t)layer * h->layer_stride;
    float* pv = h->pool_v + (size_t)layer * h->layer_stride;
    const int* slots = h->d_step + L.off_slot;
    const bool vec = (h->C % 4 == 0) && (row_stride % 4 == 0) && aligned16(k) && aligned16(v);
    const int per_row = vec ? h->C / 4 : h->C;
    dim3 grid((per_row + 255) / 256, L.ntok > 65535 ? 65535 : L.ntok);
    if (vec) pa_append_kernel<true><<<grid, 256, 0, s>>>(k, v, row_stride, pk, pv, slots, L.ntok, h->C);
    else pa_append_kernel<false><<<grid, 256, 0, s>>>(k, v, row_stride, pk, pv, slots, L.ntok, h->C);
    CU_CHECK(cudaGetLastError());
    h->launches++;
    return PA_OK;
}

static int decode_impl(pa_handle* h, int layer, const float* q, int q_stride, const float* k_new,
                       const float* v_new, int new_stride, float* out, int out_stride, void* stream,
                       const char* who) {
    int rc = check_compute(h, layer, who);
    if (rc != PA_OK) return rc;
    if (!q || !out || q_stride < h->C || out_stride < h->C) { pa_set_error("%s: bad q/out", who); return PA_ERR_INVALID; }
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)h->stream;
    const pa_step_layout& L = h->step;
    const bool fused = k_new != nullptr;
    if (fused && (!v_new || new_stride < h->C || L.ntok != L.nseq || L.max_q != 1)) {
        pa_set_error("%s: fused append needs exactly one new token per sequence in the step", who);
        return PA_ERR_INVALID;
    }
    const int hs = h->cfg.head_dim, bs = h->cfg.block_size;
    const int path = h->tune[PA_TUNE_DECODE_PATH];
    decode_fn_t fn = pick_decode(hs, bs);
    bool stream_ok = fn != nullptr && (q_stride % 4 == 0) && (out_stride % 4 == 0) && aligned16(q) && aligned16(out);
    if (fused) stream_ok = stream_ok && (new_stride % 4 == 0) && aligned16(k_new) && aligned16(v_new);
    if (path == 1 && !stream_ok) {
        pa_set_error("%s: stream kernel needs head_dim 64/128, block_size 4/8/16/32 and 16-byte aligned rows", who);
        return PA_ERR_UNSUPPORTED;
    }
    // small work: one CTA per (sequence, head) (path 3; chosen by itself below a measured amount of KV)
    const bool small_ok = (hs == 32 || hs == 64 || hs == 128) && (q_stride % 4 == 0) && (h->C % 4 == 0) && aligned16(q) &&
                          (!fused || ((new_stride % 4 == 0) && aligned16(k_new) && aligned16(v_new))) && L.nseq <= 65535;
    if (path == 3 && !small_ok) {
        pa_set_error("%s: the small-batch kernel needs head_dim 32/64/128 and 16-byte aligned rows", who);
        return PA_ERR_UNSUPPORTED;
    }
    static const long long small_max = getenv("PA_DECODE_SMALL_MAX") ? atoll(getenv("PA_DECODE_SMALL_MAX")) : 8192;       // token-heads / 12: measured crossover ~98 k (tools/decode_small_sweep.py)
    if (small_ok && (path == 3 || (path == 0 && (long long)L.total_pages * bs * h->cfg.n_heads <= small_max * 12))) {
        SmallParams sp;
        sp.pool_k = h->pool_k + (size_t)layer * h->layer_stride;
        sp.pool_v = h->pool_v + (size_t)layer * h->layer_stride;
        sp.q = q; sp.k_new = k_new; sp.v_new = v_new; sp.out = out;
        sp.kv_end = h->d_step + L.off_kv_end; sp.kv_start = h->d_step + L.off_kv_start;
        sp.slots = h->d_step + L.off_slot; sp.table = h->d_step + L.off_table;
        sp.C = h->C; sp.NH = h->cfg.n_heads; sp.hs = hs; sp.bs = bs; sp.tstride = L.tstride;
        sp.q_stride = q_stride; sp.out_stride = out_stride; sp.new_stride = new_stride;
        sp.scale = (float)(1.0 / sqrtf((float)hs));          // paged_infer.c:174
        const dim3 grid(h->cfg.n_heads, L.nseq), block(kSmallWarps * 32);
        if (hs == 64) CU_CHECK(pa_launch_pdl(pa_decode_small_kernel<16>, grid, block, 0, s, 1, sp));
        else if (hs == 128) CU_CHECK(pa_launch_pdl(pa_decode_small_kernel<32>, grid, block, 0, s, 1, sp));
        else CU_CHECK(pa_launch_pdl(pa_decode_small_kernel<8>, grid, block, 0, s, 1, sp));
        h->launches++;
        h->tune[PA_TUNE_LAST_HPG] = h->tune[PA_TUNE_LAST_STAGES] = h->tune[PA_TUNE_LAST_GRID] = 0;
        return PA_OK;
    }
    DecodePlan plan;
    if (path != 2 && path != 3 && stream_ok && plan_decode(h, L.total_pages, L.nseq, &plan)) {
        DecodeParams dp;
        dp.pool_k = h->pool_k + (size_t)layer * h->layer_stride;
        dp.pool_v = h->pool_v + (size_t)layer * h->layer_stride;
        dp.q = q; dp.out = out;
        dp.k_new = k_new; dp.v_new = v_new; dp.new_stride = new_stride;
        dp.kv_end = h->d_step + L.off_kv_end;
        dp.kv_start = h->d_step + L.off_kv_start;
        dp.cum_pages = h->d_step + L.off_cum_pages;
        dp.table = h->d_step + L.off_table;
        dp.ws = h->d_ws; dp.counters = h->d_counters; dp.sched = h->d_counters + h->n_counters;
        dp.dbg = nullptr;
        if (h->tune[PA_TUNE_DEBUG_TIMELINE]) {
            if (!h->d_dbg) {
                CU_CHECK(cudaMalloc((void**)&h->d_dbg, (size_t)h->sm_count * 8 * 8 * sizeof(unsigned long long)));
            }
            CU_CHECK(cudaMemsetAsync(h->d_dbg, 0, (size_t)h->sm_count * 8 * 8 * sizeof(unsigned long long), s));
            dp.dbg = (unsigned long long*)h->d_dbg;
            h->dbg_ctas = plan.grid;
        }
        dp.P = L.total_pages > 0 ? L.total_pages : 1;
        dp.U = plan.U; dp.Us = plan.Us; dp.Gs = plan.Gs; dp.D = plan.D;
        dp.B = L.nseq; dp.C = h->C; dp.hpg = plan.hpg; dp.n_hg = plan.n_hg; dp.W = plan.W;
        dp.tstride = L.tstride; dp.q_stride = q_stride; dp.out_stride = out_stride;
        dp.n_stages = plan.n_stages; dp.n_cons = plan.n_cons; dp.slot_floats = plan.slot_floats;
        dp.n_cum_smem = plan.n_cum_smem;
        dp.scale = (float)(1.0 / sqrtf((float)hs));          // paged_infer.c:174
        if ((size_t)(plan.Gs + plan.D + plan.n_hg * L.nseq) * plan.slot_floats > h->ws_floats) {
            pa_set_error("%s: split workspace too small", who);
            return PA_ERR_INVALID;
        }
        if (h->decode_attr_fn != (void*)fn) {      // once per handle (one geometry per handle)
            CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
            h->decode_attr_fn = (void*)fn;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(plan.grid);
        cfg.blockDim = dim3(plan.n_cons + 64);     // consumers + producer warp + merger warp
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = h->tune[PA_TUNE_NO_PDL] ? 0 : 1;
        CU_CHECK(cudaLaunchKernelEx(&cfg, fn, dp));
        h->launches++;
        h->tune[PA_TUNE_LAST_HPG] = plan.hpg;
        h->tune[PA_TUNE_LAST_STAGES] = plan.n_stages;
        h->tune[PA_TUNE_LAST_GRID] = plan.grid;
        return PA_OK;
    }
    h->tune[PA_TUNE_LAST_HPG] = h->tune[PA_TUNE_LAST_STAGES] = h->tune[PA_TUNE_LAST_GRID] = 0;
    if (fused) {
        rc = pa_append(h, layer, k_new, v_new, new_stride, s);
        if (rc != PA_OK) return rc;
    }
    return launch_rows(h, layer, q, q_stride, out, out_stride, false, s);
}

int pa_debug_timeline(pa_handle* h, unsigned long long* out, int max_ctas) {
    if (!h || !h->d_dbg || !out) { pa_set_error("pa_debug_timeline: enable PA_TUNE_DEBUG_TIMELINE and run a decode first"); return PA_ERR_INVALID; }
    int n = h->dbg_ctas < max_ctas ? h->dbg_ctas : max_ctas;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    CU_CHECK(cudaDeviceSynchronize());
    CU_CHECK(cudaMemcpy(out, h->d_dbg, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return n;
}

int pa_decode(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride, void* stream) {
    return decode_impl(h, layer, q, q_stride, nullptr, nullptr, 0, out, out_stride, stream, "pa_decode");
}

int pa_decode_append(pa_handle* h, int layer, const float* q, const float* k, const float* v, int row_stride,
                     float* out, int out_stride, void* stream) {
    if (!k || !v) { pa_set_error("pa_decode_append: NULL k/v"); return PA_ERR_INVALID; }
    return decode_impl(h, layer, q, row_stride, k, v, row_stride, out, out_stride, stream, "pa_decode_append");
}

int pa_prefill(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride, void* stream) {
    int rc = check_compute(h, layer, "pa_prefill");
    if (rc != PA_OK) return rc;
    if (!q || !out || q_stride < h->C || out_stride < h->C) { pa_set_error("pa_prefill: bad q/out"); return PA_ERR_INVALID; }
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)h->stream;
    const int path = h->tune[PA_TUNE_PREFILL_PATH];
    if (path == 3) {      // tensor-core TF32 variant: opt-in, own tolerance
        rc = pa_cu_prefill_tc(h, layer, q, q_stride, out, out_stride, (void*)s);
        if (rc == PA_ERR_UNSUPPORTED)
            pa_set_error("pa_prefill: tcgen05 kernel needs head_dim 64/128, block_size 8..128 (power of two), 16-byte aligned rows");
        return rc;
    }
    // fp32-accurate tensor-core kernel (tcgen05 3xTF32, tolerance 1e-5): forced with 4, and the automatic choice (0)
    // wherever its domain allows.  Since it became persistent there is no crossover left against the tiled SIMT kernel:
    // it is as fast at 1 x 16 rows and 1.6-2.7 x faster from 16 x 128 on, and 4-6 x on few rows over a long cache
    // (64 x 4 rows on 1024 cached tokens: 0.165 vs 0.891 ms; profiles/r02_prefill.md).  PA_PREFILL_TC3_MIN_ROWS
    // restores a threshold.
    static const int tc3_min_rows = getenv("PA_PREFILL_TC3_MIN_ROWS") ? atoi(getenv("PA_PREFILL_TC3_MIN_ROWS")) : 1;
    if (path == 4 || (path == 0 && h->step.ntok >= tc3_min_rows)) {
        rc = pa_cu_prefill_tc3(h, layer, q, q_stride, out, out_stride, (void*)s);
        if (rc != PA_ERR_UNSUPPORTED) return rc;
        if (path == 4) {
            pa_set_error("pa_prefill: the 3xTF32 tcgen05 kernel needs head_dim 64/128, block_size 8..64 (head_dim 128: 8..32; a power of two), 16-byte aligned rows");
            return rc;
        }
    }
    if (path != 2) {
        rc = pa_cu_prefill_tiled(h, layer, q, q_stride, out, out_stride, 1, (void*)s);
        if (rc != PA_ERR_UNSUPPORTED) return rc;
        if (path == 1) { pa_set_error("pa_prefill: tiled kernel needs head_dim 64/128 and 16-byte aligned rows"); return rc; }
    }
    return launch_rows(h, layer, q, q_stride, out, out_stride, true, s);
}

struct HostPipe {            // streams and events of the staged host-buffer pipeline (pa_decode_step_host_async)
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> ev;      // per layer: input landed, kernel done, output copied
    std::vector<bool> used;
    int n_layers = 0;
    // completion tickets (pa_decode_step_host_mark / _wait): a ring of events, two per ticket (kernel stream, D2H stream)
    static constexpr int kTickets = 8;
    cudaEvent_t tick[kTickets][2] = {};
    long long next_ticket = 1;
    // pinned host pointer -> device alias, remembered (cudaPointerGetAttributes costs ~1 us per call)
    struct Alias { const void* host = nullptr; void* dev = nullptr; bool pinned = false; unsigned gen = 0; };
    Alias alias[4];
    int alias_next = 0;
    // whole-step staging (pa_decode_step_host_layers_async): two sets of device buffers so that the input copy of
    // step n+1 and the output copy of step n-1 run beside the kernels of step n
    float* st_in[2] = {nullptr, nullptr};
    float* st_out[2] = {nullptr, nullptr};
    size_t st_in_floats = 0, st_out_floats = 0;
    cudaEvent_t st_ev[2][3] = {};      // per set: inputs landed, kernels done, outputs copied
    bool st_used[2] = {false, false};
    int st_next = 0;
};
static HostPipe* host_pipe_get(pa_handle* h) {
    if (!h->host_pipe) h->host_pipe = new HostPipe();
    return (HostPipe*)h->host_pipe;
}
// pinned (page-locked) host memory and its address in the device's address space, or {false} for pageable memory
static HostPipe::Alias host_alias(pa_handle* h, const void* p) {
    HostPipe* hp = host_pipe_get(h);
    for (auto& a : hp->alias) if (a.host == p && a.gen == pa_host_free_generation) return a;      // (a pa_host_free since then: ask again)
    HostPipe::Alias a;
    a.host = p;
    a.gen = pa_host_free_generation;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost) { a.pinned = true; a.dev = attr.devicePointer; }
    cudaGetLastError();
    if (a.pinned) { hp->alias[hp->alias_next] = a; hp->alias_next = (hp->alias_next + 1) % 4; }    // (pageable memory may be freed and pinned later: not remembered)
    return a;
}
void pa_cu_host_pipe_release(pa_handle* h) {
    HostPipe* hp = (HostPipe*)h->host_pipe;
    if (!hp) return;
    for (auto& t : hp->tick) for (auto e : t) if (e) cudaEventDestroy(e);
    for (auto& t : hp->st_ev) for (auto e : t) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { cudaFree(hp->st_in[i]); cudaFree(hp->st_out[i]); }
    for (auto e : hp->ev) if (e) cudaEventDestroy(e);
    if (hp->h2d) cudaStreamDestroy(hp->h2d);
    if (hp->d2h) cudaStreamDestroy(hp->d2h);
    delete hp;
    h->host_pipe = nullptr;
}
/* everything queued by pa_decode_step_host_async (kernels and both copy directions) has completed */
int pa_decode_step_host_sync(pa_handle* h) {
    if (!h || h->host_only) { pa_set_error("pa_decode_step_host_sync: no device"); return PA_ERR_NO_DEVICE; }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    CU_CHECK(cudaStreamSynchronize((cudaStream_t)h->stream));
    HostPipe* hp = (HostPipe*)h->host_pipe;
    if (hp && hp->h2d) { CU_CHECK(cudaStreamSynchronize(hp->h2d)); CU_CHECK(cudaStreamSynchronize(hp->d2h)); }
    return PA_OK;
}
/* A completion ticket for everything pa_decode_step_host_async has queued so far: the host can queue the NEXT
 * step (its tables, its layers) right away and wait for THIS step's outputs afterwards -- the device never sits
 * idle while the host turns a step around.  Tickets are positive and valid until 8 newer ones were made. */
int pa_decode_step_host_mark(pa_handle* h) {
    if (!h || h->host_only) { pa_set_error("pa_decode_step_host_mark: no device"); return PA_ERR_NO_DEVICE; }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    HostPipe* hp = host_pipe_get(h);
    const long long t = hp->next_ticket;
    cudaEvent_t* ev = hp->tick[t % HostPipe::kTickets];
    for (int i = 0; i < 2; ++i)
        if (!ev[i]) CU_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    CU_CHECK(cudaEventRecord(ev[0], (cudaStream_t)h->stream));
    CU_CHECK(cudaEventRecord(ev[1], hp->d2h ? hp->d2h : (cudaStream_t)h->stream));
    hp->next_ticket = t + 1;
    return (int)(t & 0x3fffffff);
}
int pa_decode_step_host_wait(pa_handle* h, int ticket) {
    if (!h || h->host_only) { pa_set_error("pa_decode_step_host_wait: no device"); return PA_ERR_NO_DEVICE; }
    HostPipe* hp = (HostPipe*)h->host_pipe;
    const long long newest = hp ? ((hp->next_ticket - 1) & 0x3fffffff) : 0;
    const long long age = (newest - ticket) & 0x3fffffff;
    if (!hp || ticket < 1 || age >= HostPipe::kTickets || !hp->tick[ticket % HostPipe::kTickets][0]) {
        pa_set_error("pa_decode_step_host_wait: ticket %d is not one of the last %d made", ticket, HostPipe::kTickets);
        return PA_ERR_INVALID;
    }
    CU_CHECK(cudaEventSynchronize(hp->tick[ticket % HostPipe::kTickets][0]));
    CU_CHECK(cudaEventSynchronize(hp->tick[ticket % HostPipe::kTickets][1]));
    return PA_OK;
}
static int decode_step_host_impl(pa_handle* h, int layer, const float* qkv_host, float* out_host, bool sync);
int pa_decode_step_host(pa_handle* h, int layer, const float* qkv_host, float* out_host) {
    return decode_step_host_impl(h, layer, qkv_host, out_host, true);
}
/* Same, stream-ordered on the handle's stream without the final synchronisation: PINNED buffers only
 * (pa_host_alloc); the caller calls pa_decode_step_host_sync(h) before reading out_host or reusing
 * qkv_host.  Lets a host queue several layers / steps behind each other. */
int pa_decode_step_host_async(pa_handle* h, int layer, const float* qkv_host, float* out_host) {
    return decode_step_host_impl(h, layer, qkv_host, out_host, false);
}
/* Every layer of the step in ONE call (one trip through the binding instead of n_layers): layer l reads its
 * q|k|v rows at qkv_host + l * qkv_layer_stride floats (0: the same rows for every layer) and writes
 * out_host + l * out_layer_stride. */
static int host_pipe_streams(pa_handle* h, HostPipe* hp) {
    if (hp->h2d) return PA_OK;
    CU_CHECK(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
    CU_CHECK(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
    return PA_OK;
}
int pa_decode_step_host_layers_async(pa_handle* h, const float* qkv_host, size_t qkv_layer_stride, float* out_host,
                                     size_t out_layer_stride) {
    if (!h || h->host_only) { pa_set_error("pa_decode_step_host_layers_async: no device; there is no CPU fallback"); return PA_ERR_NO_DEVICE; }
    if (!qkv_host || !out_host) { pa_set_error("pa_decode_step_host_layers_async: NULL buffer"); return PA_ERR_INVALID; }
    const int mode = h->tune[PA_TUNE_NO_ZEROCOPY];
    const int L = h->cfg.n_layers;
    const size_t C = h->C, n = h->step.nseq;
    const HostPipe::Alias a_in = host_alias(h, qkv_host), a_out = host_alias(h, out_host);
    // Whole-step staging (mode 0 auto, 3 forced): ONE input copy per step on a copy stream, the layers' kernels on
    // device-resident rows with their launch overlap intact (no event between them), ONE output copy per step on a
    // second copy stream.  Measured against the kernels reading host memory themselves (mode 2, zero-copy): PCIe
    // latency inside every layer's kernel costs ~7 us per layer; staged, a step costs what it costs device-resident.
    // Which of the two wins depends on the step: zero-copy pays PCIe latency once per layer (~7 us), staging pays for
    // its copies' writes into an HBM the kernels already saturate (~4 us per MB of inputs; measured on B200: 64
    // sequences x 12 layers 0.820 ms staged vs 0.894 zero-copy, 256 x 12 layers 1.835 vs 1.770).  Auto takes the cheaper.
    const double in_mb = (double)(qkv_layer_stride ? (size_t)L : 1) * n * 3 * C * sizeof(float) / 1e6;
    const bool staged_pays = mode == 3 || in_mb * 4.0 < (double)L * 7.0;
    const bool staged = (mode == 0 || mode == 3) && staged_pays && a_in.pinned && a_out.pinned && h->step.nseq >= 1 &&
                        h->step.ntok == h->step.nseq && (qkv_layer_stride == 0 || qkv_layer_stride >= n * 3 * C) && out_layer_stride >= n * C;
    if (!staged) {
        if (mode == 3) { pa_set_error("pa_decode_step_host_layers_async: whole-step staging needs pinned buffers and per-layer output rows"); return PA_ERR_INVALID; }
        for (int l = 0; l < L; ++l) {
            const int rc = decode_step_host_impl(h, l, qkv_host + (size_t)l * qkv_layer_stride, out_host + (size_t)l * out_layer_stride, false);
            if (rc != PA_OK) return rc;
        }
        return PA_OK;
    }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    HostPipe* hp = host_pipe_get(h);
    int rc = host_pipe_streams(h, hp);
    if (rc != PA_OK) return rc;
    cudaStream_t s = (cudaStream_t)h->stream;
    const size_t in_layers = qkv_layer_stride ? (size_t)L : 1;
    const size_t in_floats = in_layers * n * 3 * C, out_floats = (size_t)L * n * C;
    if (in_floats > hp->st_in_floats || out_floats > hp->st_out_floats) {
        CU_CHECK(cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            cudaFree(hp->st_in[i]); cudaFree(hp->st_out[i]);
            hp->st_in[i] = hp->st_out[i] = nullptr;
            CU_CHECK(cudaMalloc((void**)&hp->st_in[i], in_floats * sizeof(float)));
            CU_CHECK(cudaMalloc((void**)&hp->st_out[i], out_floats * sizeof(float)));
            hp->st_used[i] = false;
            for (auto& e : hp->st_ev[i]) if (!e) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        hp->st_in_floats = in_floats; hp->st_out_floats = out_floats;
    }
    const int b = hp->st_next;
    hp->st_next ^= 1;
    cudaEvent_t ev_in = hp->st_ev[b][0], ev_k = hp->st_ev[b][1], ev_out = hp->st_ev[b][2];
    if (hp->st_used[b]) {
        CU_CHECK(cudaStreamWaitEvent(hp->h2d, ev_k, 0));      // the kernels that last read this input set
        CU_CHECK(cudaStreamWaitEvent(s, ev_out, 0));          // the copy that last read this output set
    }
    if (qkv_layer_stride == 0 || qkv_layer_stride == n * 3 * C) {
        CU_CHECK(cudaMemcpyAsync(hp->st_in[b], qkv_host, in_floats * sizeof(float), cudaMemcpyHostToDevice, hp->h2d));
    } else {
        CU_CHECK(cudaMemcpy2DAsync(hp->st_in[b], n * 3 * C * sizeof(float), qkv_host, qkv_layer_stride * sizeof(float),
                                   n * 3 * C * sizeof(float), (size_t)L, cudaMemcpyHostToDevice, hp->h2d));
    }
    CU_CHECK(cudaEventRecord(ev_in, hp->h2d));
    if (!h->step.uploaded) { rc = pa_step_upload(h, s); if (rc != PA_OK) return rc; }
    CU_CHECK(cudaStreamWaitEvent(s, ev_in, 0));
    for (int l = 0; l < L; ++l) {
        const float* in_l = hp->st_in[b] + (qkv_layer_stride ? (size_t)l * n * 3 * C : 0);
        rc = pa_decode_append(h, l, in_l, in_l + C, in_l + 2 * C, (int)(3 * C), hp->st_out[b] + (size_t)l * n * C, (int)C, s);
        if (rc != PA_OK) return rc;
    }
    CU_CHECK(cudaEventRecord(ev_k, s));
    CU_CHECK(cudaStreamWaitEvent(hp->d2h, ev_k, 0));
    if (out_layer_stride == n * C) {
        CU_CHECK(cudaMemcpyAsync(out_host, hp->st_out[b], out_floats * sizeof(float), cudaMemcpyDeviceToHost, hp->d2h));
    } else {
        CU_CHECK(cudaMemcpy2DAsync(out_host, out_layer_stride * sizeof(float), hp->st_out[b], n * C * sizeof(float), n * C * sizeof(float),
                                   (size_t)L, cudaMemcpyDeviceToHost, hp->d2h));
    }
    CU_CHECK(cudaEventRecord(ev_out, hp->d2h));
    hp->st_used[b] = true;
    return PA_OK;
}
static int decode_step_host_impl(pa_handle* h, int layer, const float* qkv_host, float* out_host, bool sync) {
    if (!h || h->host_only) { pa_set_error("pa_decode_step_host: no device; there is no CPU fallback"); return PA_ERR_NO_DEVICE; }
    if (!qkv_host || !out_host) { pa_set_error("pa_decode_step_host: NULL buffer"); return PA_ERR_INVALID; }
    const pa_step_layout& L = h->step;
    if (L.nseq < 1 || L.ntok != L.nseq) { pa_set_error("pa_decode_step_host: needs a step with one new token per sequence"); return PA_ERR_INVALID; }
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return PA_ERR_CUDA;
    const size_t C = h->C, n = L.nseq;
    cudaStream_t s = (cudaStream_t)h->stream;
    int rc;
    if (!h->step.uploaded) { rc = pa_step_upload(h, s); if (rc != PA_OK) return rc; }
    const HostPipe::Alias a_in = host_alias(h, qkv_host), a_out = host_alias(h, out_host);
    const bool in_pinned = a_in.pinned, out_pinned = a_out.pinned;
    const float* in_alias = (const float*)a_in.dev;
    float* out_alias = (float*)a_out.dev;
    if (in_alias && out_alias && h->tune[PA_TUNE_NO_ZEROCOPY] != 1) {
        /* Pinned host buffers are mapped into the device address space: the kernel's bulk copies
         * pull the q / k / v rows over PCIe themselves (each row is read exactly once) and the
         * output rows are stored straight to host memory -- no staging copies, one launch. */
        rc = pa_decode_append(h, layer, in_alias, in_alias + C, in_alias + 2 * C, (int)(3 * C), out_alias, (int)C, s);
        if (rc != PA_OK) return rc;
        if (sync) CU_CHECK(cudaStreamSynchronize(s));
        return PA_OK;
    }
    if (!sync && (!in_pinned || !out_pinned)) {
        pa_set_error("pa_decode_step_host_async: pageable host buffers need the synchronous entry (they are staged through one pinned buffer)");
        return PA_ERR_INVALID;
    }
    if (!sync) {
        /* pinned, zero-copy switched off: a three-stream pipeline.  The H2D copy of layer l+1 runs on its
         * own stream while the kernel of layer l computes, and the D2H copy of layer l's output overlaps the
         * kernel of layer l+1; per-layer staging regions and events keep the queued layers apart. */
        rc = pa_cu_ensure_stage(h, (size_t)h->cfg.n_layers * n * 4 * C);
        if (rc != PA_OK) return rc;
        HostPipe* hp = host_pipe_get(h);
        rc = host_pipe_streams(h, hp);
        if (rc != PA_OK) return rc;
        if (hp->ev.empty()) {
            hp->n_layers = h->cfg.n_layers;
            hp->ev.resize((size_t)3 * hp->n_layers);
            for (auto& e : hp->ev) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            hp->used.assign(hp->n_layers, false);
        }
        cudaEvent_t ev_in = hp->ev[3 * layer], ev_k = hp->ev[3 * layer + 1], ev_out = hp->ev[3 * layer + 2];
        float* d_qkv_l = h->d_stage + (size_t)layer * n * 4 * C;
        float* d_out_l = d_qkv_l + n * 3 * C;
        if (hp->used[layer]) {                       // the previous step's kernel / D2H of this layer's regions
            CU_CHECK(cudaStreamWaitEvent(hp->h2d, ev_k, 0));
            CU_CHECK(cudaStreamWaitEvent(s, ev_out, 0));
        }
        CU_CHECK(cudaMemcpyAsync(d_qkv_l, qkv_host, n * 3 * C * sizeof(float), cudaMemcpyHostToDevice, hp->h2d));
        CU_CHECK(cudaEventRecord(ev_in, hp->h2d));
        CU_CHECK(cudaStreamWaitEvent(s, ev_in, 0));
        rc = pa_decode_append(h, layer, d_qkv_l, d_qkv_l + C, d_qkv_l + 2 * C, (int)(3 * C), d_out_l, (int)C, s);
        if (rc != PA_OK) return rc;
        CU_CHECK(cudaEventRecord(ev_k, s));
        CU_CHECK(cudaStreamWaitEvent(hp->d2h, ev_k, 0));
        CU_CHECK(cudaMemcpyAsync(out_host, d_out_l, n * C * sizeof(float), cudaMemcpyDeviceToHost, hp->d2h));
        CU_CHECK(cudaEventRecord(ev_out, hp->d2h));
        hp->used[layer] = true;
        return PA_OK;
    }
    rc = pa_cu_ensure_stage(h, n * 4 * C);
    if (rc != PA_OK) return rc;
    float* d_qkv = h->d_stage;
    float* d_out = h->d_stage + n * 3 * C;
    /* pageable host memory is staged through the pinned buffer; pinned memory goes straight */
    const float* src = qkv_host;
    if (!in_pinned) { memcpy(h->h_stage, qkv_host, n * 3 * C * sizeof(float)); src = h->h_stage; }
    CU_CHECK(cudaMemcpyAsync(d_qkv, src, n * 3 * C * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = pa_decode_append(h, layer, d_qkv, d_qkv + C, d_qkv + 2 * C, (int)(3 * C), d_out, (int)C, s);
    if (rc != PA_OK) return rc;
    float* dst = out_pinned ? out_host : h->h_stage + n * 3 * C;
    CU_CHECK(cudaMemcpyAsync(dst, d_out, n * C * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU_CHECK(cudaStreamSynchronize(s));
    if (!out_pinned) memcpy(out_host, dst, n * C * sizeof(float));
    return PA_OK;
}

}  // extern "C"
