/*
 * pa_qkv.cu -- the step immediately before the paged-attention path: the QKV
 * projection of the step's new tokens, with the KV append fused into its epilogue.
 *
 *   out[m][n] = bias[n] + sum_i x[m][i] * w[n][i]        (matmul_forward, paged_infer.c:92-114)
 *
 * For columns n < C (Q) the result goes to a dense (rows, C) buffer; for C <= n < 3C (K, V) the
 * epilogue writes straight to the token's page slot pool[slot_mapping[m]][n - C | n - 2C], so the
 * new K/V rows never make a round trip through a (B,T,3C) activation buffer and no separate
 * append kernel runs (the reference does matmul_cached -> add_to_cache, paged_infer.c:706-710).
 *
 * Default: the 3xTF32 tensor-core GEMM in pa_gemm_tc.cu (fp32-accurate).  This file holds the
 * fp32 SIMT kernel used for shapes outside that kernel's domain (unaligned rows, gathered input
 * rows) and as the cross-check: register-tiled, both operands K-contiguous.  The reference
 * accumulates sequentially from the bias (:105-110); both kernels accumulate in tiles and add
 * the bias last -- within the 1e-5 tolerance of the path (tests/test_gpu_qkv.py).
 *
 * Also exported with the reference's own names for the call at paged_infer.c:703-706:
 * matmul_forward / matmul_cached (host or device pointers, (B,T,OC) output layout).
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "pa_internal.h"
#include "pa_pdl.cuh"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {

struct QkvParams {
    const float* x;          // input rows, row m at x + in_row(m)*x_stride
    const int* in_rows;      // optional gather of input rows (NULL: identity)
    const float* w;          // (N, K) row-major
    const float* bias;       // (N) or NULL
    float* out;              // dense destination: column n < n_dense of row m at out + out_row(m)*out_stride + n
    const int* out_rows;     // optional scatter of dense output rows (NULL: identity)
    float* pool_k;           // K/V destination for columns n >= n_dense (NULL: everything is dense)
    float* pool_v;
    const int* slots;        // [M] slot_mapping
    int M, N, K;
    int x_stride, out_stride;
    int n_dense;             // columns [0, n_dense) are dense; then C columns of K, then C columns of V
    int C;
    const float* residual;   // optional: out = act(acc + bias) + residual[m][n] (may alias out)
    int res_stride;
    int act;                 // 0 none, 1 GELU (tanh form, paged_infer.c:243-251)
};

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int kThreads = 256;            // 16 x 16 threads, 4 x 4 outputs each

__global__ void __launch_bounds__(kThreads)
pa_qkv_kernel(const QkvParams p) {
    __shared__ __align__(16) float As[2][BK][BM + 4];     // x tile, transposed: [k][m]
    __shared__ __align__(16) float Bs[2][BK][BN + 4];     // w tile, transposed: [k][n]
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // each thread loads one float4 of the x tile and one of the w tile per k-slab
    const int lr = tid >> 2;                 // tile row 0..63
    const int lk = (tid & 3) * 4;            // k offset inside the slab
    const int xm = m0 + lr, wn = n0 + lr;
    const float* xrow = nullptr;
    if (xm < p.M) xrow = p.x + (size_t)(p.in_rows ? p.in_rows[xm] : xm) * p.x_stride;
    const float* wrow = wn < p.N ? p.w + (size_t)wn * p.K : nullptr;
    const bool vec = (p.K & 3) == 0 && (p.x_stride & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15) == 0;

    auto load = [&](const float* row, int k) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row) {
            if (vec && k + 3 < p.K) v = __ldg(reinterpret_cast<const float4*>(row + k));
            else {
                if (k < p.K) v.x = row[k];
                if (k + 1 < p.K) v.y = row[k + 1];
                if (k + 2 < p.K) v.z = row[k + 2];
                if (k + 3 < p.K) v.w = row[k + 3];
            }
        }
        return v;
    };
    auto stash = [&](int buf, const float4& a, const float4& b) {
        As[buf][lk][lr] = a.x; As[buf][lk + 1][lr] = a.y; As[buf][lk + 2][lr] = a.z; As[buf][lk + 3][lr] = a.w;
        Bs[buf][lk][lr] = b.x; Bs[buf][lk + 1][lr] = b.y; Bs[buf][lk + 2][lr] = b.z; Bs[buf][lk + 3][lr] = b.w;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    float4 ra = load(xrow, lk), rb = load(wrow, lk);
    stash(0, ra, rb);
    __syncthreads();
    const int n_slabs = (p.K + BK - 1) / BK;
    for (int s = 0; s < n_slabs; ++s) {
        const int buf = s & 1;
        if (s + 1 < n_slabs) {
            ra = load(xrow, (s + 1) * BK + lk);
            rb = load(wrow, (s + 1) * BK + lk);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (s + 1 < n_slabs) {
            stash(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }

    // ---- epilogue: bias, then dense store or scatter to the token's page slot -----------------
    const int n = n0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        float* dense = p.out ? p.out + (size_t)(p.out_rows ? p.out_rows[m] : m) * p.out_stride : nullptr;
        const size_t slot_off = p.slots ? (size_t)p.slots[m] * p.C : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n + j;
            if (nn >= p.N) continue;
            float v = acc[i][j] + (p.bias ? p.bias[nn] : 0.0f);
            if (p.act == 1) {
                const float cube = 0.044715f * v * v * v;
                v = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + cube)));
            }
            if (p.residual) v += p.residual[(size_t)m * p.res_stride + nn];
            if (nn < p.n_dense) {
                dense[nn] = v;
            } else {
                const int c = nn - p.n_dense;
                if (c < p.C) p.pool_k[slot_off + c] = v;
                else p.pool_v[slot_off + (c - p.C)] = v;
            }
        }
    }
}

// ---- small-M path (decode with a handful of sequences): a weight-streaming GEMV --------------------
// With M <= 4 rows the projection is a pure read of the weight matrix (3C*C*4 bytes for QKV): one
// warp per group of 4 output features streams their weight rows with 16-byte loads (24+ loads in
// flight per lane), the M input rows sit in shared memory, fp32 FMA, warp-shuffle reduction, same
// epilogue as the tiled kernels.  Roofline: HBM (weights are read exactly once).
constexpr int kGemvMaxM = 4;
constexpr int kGemvFeat = 4;         // output features per warp pass
__global__ void __launch_bounds__(256)
pa_gemv_kernel(const QkvParams p) {
    extern __shared__ __align__(16) float xs[];          // [M][K]
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31;
    const int K4 = p.K >> 2;
    for (int i = tid; i < p.M * K4; i += blockDim.x) {
        const int m = i / K4, c = i - m * K4;
        const float* row = p.x + (size_t)(p.in_rows ? p.in_rows[m] : m) * p.x_stride;
        reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(row) + c);
    }
    __syncthreads();
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int n0 = warp_global * kGemvFeat; n0 < p.N; n0 += n_warps * kGemvFeat) {
        float acc[kGemvFeat][kGemvMaxM];
#pragma unroll
        for (int f = 0; f < kGemvFeat; ++f)
#pragma unroll
            for (int m = 0; m < kGemvMaxM; ++m) acc[f][m] = 0.0f;
        for (int c = lane; c < K4; c += 32) {
            float4 wv[kGemvFeat];
#pragma unroll
            for (int f = 0; f < kGemvFeat; ++f) {
                const int n = min(n0 + f, p.N - 1);
                wv[f] = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)n * p.K) + c);
            }
#pragma unroll
            for (int m = 0; m < kGemvMaxM; ++m) {
                if (m < p.M) {
                    const float4 xv = reinterpret_cast<const float4*>(xs)[m * K4 + c];
#pragma unroll
                    for (int f = 0; f < kGemvFeat; ++f)
                        acc[f][m] = fmaf(wv[f].w, xv.w, fmaf(wv[f].z, xv.z, fmaf(wv[f].y, xv.y, fmaf(wv[f].x, xv.x, acc[f][m]))));
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kGemvFeat; ++f)
#pragma unroll
            for (int m = 0; m < kGemvMaxM; ++m)
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) acc[f][m] += __shfl_xor_sync(0xffffffffu, acc[f][m], d);
        // lane (f, m) finishes output (m, n0 + f)
        const int f = lane >> 2, m = lane & 3;
        if (f < kGemvFeat && m < p.M && n0 + f < p.N) {
            float v = 0.0f;
#pragma unroll
            for (int ff = 0; ff < kGemvFeat; ++ff)
#pragma unroll
                for (int mm = 0; mm < kGemvMaxM; ++mm)
                    if (ff == f && mm == m) v = acc[ff][mm];
            const int nn = n0 + f;
            v += p.bias ? p.bias[nn] : 0.0f;
            if (p.act == 1) {
                const float cube = 0.044715f * v * v * v;
                v = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + cube)));
            }
            if (p.residual) v += p.residual[(size_t)m * p.res_stride + nn];
            if (nn < p.n_dense) {
                p.out[(size_t)(p.out_rows ? p.out_rows[m] : m) * p.out_stride + nn] = v;
            } else {
                const size_t slot_off = (size_t)p.slots[m] * p.C;
                const int c = nn - p.n_dense;
                if (c < p.C) p.pool_k[slot_off + c] = v;
                else p.pool_v[slot_off + (c - p.C)] = v;
            }
        }
    }
}

bool gemv_ok(const QkvParams& p) {
    return p.M >= 1 && p.M <= kGemvMaxM && (p.K & 3) == 0 && (p.x_stride & 3) == 0 &&
           ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15) == 0 &&
           (size_t)p.M * p.K * sizeof(float) <= 96 * 1024;
}
int launch_gemv(const QkvParams& p, cudaStream_t s) {
    static std::atomic<unsigned long long> attr_done{0};       // one bit per device
    CU_CHECK(pa_optin_smem(attr_done, pa_gemv_kernel, 96 * 1024));
    const int groups = (p.N + kGemvFeat - 1) / kGemvFeat;            // warp passes needed
    int blocks = (groups + 7) / 8;
    if (blocks > 592) blocks = 592;                                   // 4 CTAs per SM: the rest loops
    CU_CHECK(pa_launch_pdl(pa_gemv_kernel, dim3(blocks), dim3(256), (size_t)p.M * p.K * sizeof(float), s, 1, p));
    return PA_OK;
}

int launch(const QkvParams& p, cudaStream_t s) {
    if (p.M <= 0 || p.N <= 0) return PA_OK;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM);
    CU_CHECK(pa_launch_pdl(pa_qkv_kernel, grid, dim3(kThreads), 0, s, 1, p));
    return PA_OK;
}

}  // namespace

extern "C" {

/* QKV projection of the step's new tokens with the KV append fused into the epilogue.
 * x: (ntok, C) rows in step order; w: (3C, C); bias: (3C) or NULL; q_out: (ntok, C). */
int pa_qkv_append(pa_handle* h, int layer, const float* x, int x_stride, const float* w, const float* bias,
                  float* q_out, int q_stride, void* stream) {
    if (!h) { pa_set_error("pa_qkv_append: NULL handle"); return PA_ERR_INVALID; }
    if (h->host_only || !h->pool_k) { pa_set_error("pa_qkv_append: handle has no device; there is no CPU fallback"); return PA_ERR_NO_DEVICE; }
    if (layer < 0 || layer >= h->cfg.n_layers) { pa_set_error("pa_qkv_append: layer %d out of range", layer); return PA_ERR_INVALID; }
    const pa_step_layout& L = h->step;
    if (L.nseq < 1 || !L.uploaded) { pa_set_error("pa_qkv_append: no uploaded step (pa_step_begin + pa_step_upload)"); return PA_ERR_INVALID; }
    if (!x || !w || !q_out || x_stride < h->C || q_stride < h->C) { pa_set_error("pa_qkv_append: bad arguments"); return PA_ERR_INVALID; }
    if (L.ntok == 0) return PA_OK;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    pa_pdl_enabled = h->tune[PA_TUNE_NO_PDL] ? 0 : 1;
    QkvParams p;
    p.x = x; p.in_rows = nullptr; p.w = w; p.bias = bias;
    p.out = q_out; p.out_rows = nullptr;
    p.pool_k = h->pool_k + (size_t)layer * h->layer_stride;
    p.pool_v = h->pool_v + (size_t)layer * h->layer_stride;
    p.slots = h->d_step + L.off_slot;
    p.M = L.ntok; p.N = 3 * h->C; p.K = h->C;
    p.x_stride = x_stride; p.out_stride = q_stride;
    p.n_dense = h->C; p.C = h->C;
    p.residual = nullptr; p.res_stride = 0; p.act = 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)h->stream;
    const int path = h->tune[PA_TUNE_GEMM_PATH];
    int rc = PA_ERR_UNSUPPORTED;
    if ((path == 0 || path == 4) && gemv_ok(p)) {       // a handful of rows: stream the weights once
        rc = launch_gemv(p, s);
        if (rc == PA_OK) h->launches++;
        return rc;
    }
    if (path == 4) { pa_set_error("pa_qkv_append: the GEMV path takes at most %d rows", kGemvMaxM); return PA_ERR_UNSUPPORTED; }
    if (path != 1) {       // tensor cores: 3xTF32 keeps fp32 accuracy; plain TF32 only on request
        rc = pa_cu_gemm_tc(x, x_stride, w, bias, q_out, q_stride, p.M, p.N, p.K, p.n_dense, p.pool_k, p.pool_v, p.slots,
                           p.C, path == 3 ? 1 : 3, h->tune[PA_TUNE_GEMM_SPLIT_K], nullptr, 0, 0, (void*)s);
        if (rc == PA_ERR_UNSUPPORTED && path >= 2) {
            pa_set_error("pa_qkv_append: tcgen05 GEMM needs C %% 32 == 0 and 16-byte aligned rows");
            return rc;
        }
    }
    if (rc == PA_ERR_UNSUPPORTED) rc = launch(p, s);
    if (rc == PA_OK) h->launches++;
    return rc;
}

/* out = act(x.w^T + bias) + residual on device pointers; path as PA_TUNE_GEMM_PATH */
int pa_cu_linear(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                 int M, int N, int K, const float* residual, int res_stride, int act, int path, void* stream) {
    if (!x || !w || !out || M < 0 || N < 0 || K < 1) { pa_set_error("pa_cu_linear: bad arguments"); return PA_ERR_INVALID; }
    int rc = PA_ERR_UNSUPPORTED;
    if (path == 0 || path == 4) {
        QkvParams g;
        g.x = x; g.in_rows = nullptr; g.w = w; g.bias = bias; g.out = out; g.out_rows = nullptr;
        g.pool_k = g.pool_v = nullptr; g.slots = nullptr;
        g.M = M; g.N = N; g.K = K; g.x_stride = x_stride; g.out_stride = out_stride; g.n_dense = N; g.C = 0;
        g.residual = residual; g.res_stride = res_stride; g.act = act;
        if (gemv_ok(g)) return launch_gemv(g, (cudaStream_t)stream);
        if (path == 4) { pa_set_error("pa_cu_linear: the GEMV path takes at most %d rows", kGemvMaxM); return PA_ERR_UNSUPPORTED; }
    }
    if (path != 1) {
        rc = pa_cu_gemm_tc(x, x_stride, w, bias, out, out_stride, M, N, K, N, nullptr, nullptr, nullptr, 0,
                           path == 3 ? 1 : 3, 0, residual, res_stride, act, stream);
        if (rc != PA_ERR_UNSUPPORTED) return rc;
    }
    QkvParams p;
    p.x = x; p.in_rows = nullptr; p.w = w; p.bias = bias;
    p.out = out; p.out_rows = nullptr;
    p.pool_k = p.pool_v = nullptr; p.slots = nullptr;
    p.M = M; p.N = N; p.K = K;
    p.x_stride = x_stride; p.out_stride = out_stride;
    p.n_dense = N; p.C = 0;
    p.residual = residual; p.res_stride = res_stride; p.act = act;
    return launch(p, (cudaStream_t)stream);
}

/* plain fp32 GEMM with bias on device pointers: out (M, N) = x (M, K) . w (N, K)^T + bias */
int pa_matmul_bias(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                   int M, int N, int K, void* stream) {
    return pa_cu_linear(x, x_stride, w, bias, out, out_stride, M, N, K, nullptr, 0, 0, 0, stream);
}

/* ---- reference names (paged_infer.c:92-160), host or device pointers ------------------------ */
namespace {
struct DevView {              // a device view of a caller buffer: the buffer itself, or a staged copy
    float* d = nullptr;
    float* host = nullptr;
    size_t bytes = 0;
    bool owned = false;
    bool open(const float* p, size_t n_floats, bool copy_in) {
        bytes = n_floats * sizeof(float);
        if (!p) return true;
        if (pa_cu_is_device_ptr(p)) { d = const_cast<float*>(p); return true; }
        host = const_cast<float*>(p);
        owned = true;
        if (cudaMalloc((void**)&d, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); d = nullptr; return false; }
        if (copy_in && cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return false; }
        return true;
    }
    bool copy_back() { return !owned || cudaMemcpy(host, d, bytes, cudaMemcpyDeviceToHost) == cudaSuccess; }
    ~DevView() { if (owned && d) cudaFree(d); }
};
}  // namespace

/* paged_infer.c:92-114: out (B,T,OC) = inp (B,T,C) . weight (OC,C)^T + bias */
void matmul_forward(float* out, float* inp, float* weight, float* bias, int B, int T, int C, int OC) {
    if (!out || !inp || !weight || B < 1 || T < 1 || C < 1 || OC < 1) { fprintf(stderr, "matmul_forward: invalid arguments\n"); return; }
    DevView x, w, bv, o;
    const size_t rows = (size_t)B * T;
    if (!x.open(inp, rows * C, true) || !w.open(weight, (size_t)OC * C, true) || !bv.open(bias, OC, true) ||
        !o.open(out, rows * OC, false)) { fprintf(stderr, "matmul_forward: device staging failed\n"); return; }
    if (pa_matmul_bias(x.d, C, w.d, bv.d, o.d, OC, (int)rows, OC, C, nullptr) != PA_OK ||
        cudaStreamSynchronize(0) != cudaSuccess || !o.copy_back())
        fprintf(stderr, "matmul_forward: %s\n", pa_last_error());
}

/* paged_infer.c:117-160: Q (columns [0,C)) for every row of the window, K and V (columns [C,3C))
 * for the last row of each batch entry only; everything else in `out` is left as it was. */
void matmul_cached(float* out, float* inp, float* weight, float* bias, int B, int T, int C, int OC) {
    if (!out || !inp || !weight || B < 1 || T < 1 || C < 1 || OC < 3 * C) { fprintf(stderr, "matmul_cached: invalid arguments\n"); return; }
    DevView x, w, bv, o;
    const size_t rows = (size_t)B * T;
    if (!x.open(inp, rows * C, true) || !w.open(weight, (size_t)3 * C * C, true) || !bv.open(bias, (size_t)3 * C, true) ||
        !o.open(out, rows * OC, true)) { fprintf(stderr, "matmul_cached: device staging failed\n"); return; }
    std::vector<int> last(B);
    for (int b = 0; b < B; b++) last[b] = b * T + T - 1;
    int* d_last = nullptr;
    if (cudaMalloc((void**)&d_last, (size_t)B * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(d_last, last.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d_last);
        fprintf(stderr, "matmul_cached: device staging failed\n");
        return;
    }
    QkvParams p;
    p.x = x.d; p.in_rows = nullptr; p.w = w.d; p.bias = bv.d;
    p.out = o.d; p.out_rows = nullptr;
    p.pool_k = p.pool_v = nullptr; p.slots = nullptr;
    p.M = (int)rows; p.N = C; p.K = C; p.x_stride = C; p.out_stride = OC; p.n_dense = C; p.C = 0;
    p.residual = nullptr; p.res_stride = 0; p.act = 0;
    int rc = launch(p, 0);                                   // Q for all rows
    if (rc == PA_OK) {
        p.in_rows = p.out_rows = d_last;
        p.w = w.d + (size_t)C * C; p.bias = bv.d ? bv.d + C : nullptr;
        p.out = o.d + C;
        p.M = B; p.N = 2 * C; p.n_dense = 2 * C;
        rc = launch(p, 0);                                   // K, V for the last rows
    }
    if (rc != PA_OK || cudaStreamSynchronize(0) != cudaSuccess || !o.copy_back())
        fprintf(stderr, "matmul_cached: %s\n", pa_last_error());
    cudaFree(d_last);
}

}  // extern "C"
