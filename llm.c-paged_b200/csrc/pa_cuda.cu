/*
 * pa_cuda.cu -- the thin C-ABI layer over the CUDA runtime: device selection, the KV page pool,
 * pinned step-table ring + its single-copy mirror, staging for the host-buffer entry points and
 * the small plumbing calls a plain-C host needs (no cuda_runtime.h on the host side).
 *
 * No PyTorch, no CPU fallback: every compute entry fails with PA_ERR_NO_DEVICE / PA_ERR_CUDA.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pa_internal.h"

#define CU_CHECK(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pa_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PA_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

namespace {
struct StepRing {
    static const int N = 4;
    int* buf[N];
    size_t cap[N];
    cudaEvent_t ev[N];
    bool pending[N];
    int next;
    int cur;
    bool pinned;
};
}  // namespace

extern "C" {

int pa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

/* the device the plumbing calls below (pa_dev_alloc, pa_memcpy_*, pa_stream_create, ...) act on: a host that drives
 * several GPUs from one thread switches before it allocates on another one (handles switch by themselves) */
int pa_set_device(int device) {
    CU_CHECK(cudaSetDevice(device));
    return PA_OK;
}

int pa_cu_init(pa_handle* h) {
    int n = pa_device_count();
    if (n <= 0) {
        pa_set_error("no CUDA device visible: libpaged_attn has no CPU fallback");
        return PA_ERR_NO_DEVICE;
    }
    if (h->cfg.device < 0 || h->cfg.device >= n) {
        pa_set_error("device %d out of range (%d visible)", h->cfg.device, n);
        return PA_ERR_INVALID;
    }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, h->cfg.device));
    if (prop.major < 10) {
        pa_set_error("device %d is sm_%d%d; this library is built for sm_100a only", h->cfg.device, prop.major, prop.minor);
        return PA_ERR_NO_DEVICE;
    }
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    h->max_pitch = prop.memPitch;
    cudaStream_t s;
    CU_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    h->stream = (void*)s;
    return PA_OK;
}

int pa_cu_alloc_pool(pa_handle* h) {
    CU_CHECK(cudaSetDevice(h->cfg.device));
    size_t bytes = (size_t)h->cfg.n_layers * h->layer_stride * sizeof(float);
    /* Reference host code writes pages through KVBlock.keys / .values directly (block_manager_test.c:9-29).  For such
     * callers a manager made by create_block_manager() can keep its pool in MANAGED memory (PA_COMPAT_HOST_PAGES=1): the
     * same pointers are valid on the host and on the device (pages migrate on touch).  Everything else uses plain device
     * memory -- the TMA / bulk-copy kernels want their pages resident in HBM. */
    const char* hostp = getenv("PA_COMPAT_HOST_PAGES");
    const bool managed = h->compat && hostp && atoi(hostp) != 0;
    cudaError_t e = managed ? cudaMallocManaged((void**)&h->pool_k, bytes) : cudaMalloc((void**)&h->pool_k, bytes);
    if (e == cudaSuccess) e = managed ? cudaMallocManaged((void**)&h->pool_v, bytes) : cudaMalloc((void**)&h->pool_v, bytes);
    if (e != cudaSuccess) {
        pa_set_error("KV pool: cudaMalloc of 2 x %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return PA_ERR_NOMEM;
    }
    CU_CHECK(cudaMemsetAsync(h->pool_k, 0, bytes, (cudaStream_t)h->stream));
    CU_CHECK(cudaMemsetAsync(h->pool_v, 0, bytes, (cudaStream_t)h->stream));
    /* split-decode workspace: (max CTAs + max_seqs) partial slots of (C + 2*NH) floats, and one
     * arrival counter per (sequence, head) -- see pa_decode_stream_kernel */
    size_t max_ctas = (size_t)h->sm_count * 24;   /* static + dynamic ranges */
    if (h->max_heads < h->cfg.n_heads) h->max_heads = h->cfg.n_heads;
    h->ws_floats = (max_ctas + (size_t)h->cfg.max_seqs) * ((size_t)h->C + 2 * (size_t)h->max_heads + 8);
    h->n_counters = (size_t)h->cfg.max_seqs * h->max_heads;
    CU_CHECK(cudaMalloc((void**)&h->d_ws, h->ws_floats * sizeof(float)));
    /* + 2 scheduler words (next dynamic range, finished CTAs) */
    CU_CHECK(cudaMalloc((void**)&h->d_counters, (h->n_counters + 2) * sizeof(int)));
    CU_CHECK(cudaMemsetAsync(h->d_counters, 0, (h->n_counters + 2) * sizeof(int), (cudaStream_t)h->stream));
    CU_CHECK(cudaStreamSynchronize((cudaStream_t)h->stream));
    return PA_OK;
}

void pa_cu_release(pa_handle* h) {
    StepRing* r = (StepRing*)h->step_ring;
    if (!h->host_only && h->stream) {
        cudaSetDevice(h->cfg.device);
        cudaStreamSynchronize((cudaStream_t)h->stream);
    }
    if (r) {
        for (int i = 0; i < StepRing::N; i++) {
            if (r->buf[i]) { if (r->pinned) cudaFreeHost(r->buf[i]); else free(r->buf[i]); }
            if (r->pinned && r->ev[i]) cudaEventDestroy(r->ev[i]);
        }
        delete r;
        h->step_ring = NULL;
    }
    if (h->host_only) return;
    pa_cu_prefill_tc_release(h);
    pa_cu_prefill_tc3_release(h);
    pa_cu_host_pipe_release(h);
    cudaFree(h->pool_k); cudaFree(h->pool_v);
    cudaFree(h->d_step); cudaFree(h->d_ws); cudaFree(h->d_counters);
    cudaFree(h->d_stage);
    cudaFree(h->d_dbg);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->stream) {
        cudaStreamSynchronize((cudaStream_t)h->stream);
        pa_cu_gemm_stream_released(h->cfg.device, h->stream);
        cudaStreamDestroy((cudaStream_t)h->stream);
    }
    h->pool_k = h->pool_v = NULL;
}

int* pa_cu_step_host_buffer(pa_handle* h, size_t ints) {
    StepRing* r = (StepRing*)h->step_ring;
    if (!r) {
        r = new StepRing();
        memset(r, 0, sizeof(*r));
        r->pinned = !h->host_only;
        h->step_ring = r;
    }
    if (r->pinned) cudaSetDevice(h->cfg.device);
    int i = r->next;
    r->next = (r->next + 1) % StepRing::N;
    if (r->pinned && r->pending[i]) {     /* the upload that last used this buffer must be done */
        cudaEventSynchronize(r->ev[i]);
        r->pending[i] = false;
    }
    if (r->cap[i] < ints) {
        size_t cap = ints + ints / 2 + 1024;
        if (r->buf[i]) { if (r->pinned) cudaFreeHost(r->buf[i]); else free(r->buf[i]); r->buf[i] = NULL; }
        if (r->pinned) {
            if (cudaMallocHost((void**)&r->buf[i], cap * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                pa_set_error("pinned step buffer: cudaMallocHost(%zu) failed", cap * sizeof(int));
                r->cap[i] = 0;
                return NULL;
            }
            if (!r->ev[i]) cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming);
        } else {
            r->buf[i] = (int*)malloc(cap * sizeof(int));
            if (!r->buf[i]) { pa_set_error("step buffer: out of host memory"); r->cap[i] = 0; return NULL; }
        }
        r->cap[i] = cap;
    }
    r->cur = i;
    h->step.uploaded = 0;
    return r->buf[i];
}

int pa_cu_step_upload(pa_handle* h, void* stream) {
    StepRing* r = (StepRing*)h->step_ring;
    if (!r || !h->h_step) { pa_set_error("pa_step_upload: no step tables"); return PA_ERR_INVALID; }
    CU_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)h->stream;
    size_t ints = (size_t)h->step.total_ints;
    if (h->d_step_cap_ints < ints) {
        /* growing the mirror is rare; make sure nothing still reads the old one */
        CU_CHECK(cudaDeviceSynchronize());
        cudaFree(h->d_step);
        h->d_step = NULL;
        size_t cap = ints + ints / 2 + 1024;
        CU_CHECK(cudaMalloc((void**)&h->d_step, cap * sizeof(int)));
        h->d_step_cap_ints = cap;
    }
    CU_CHECK(cudaMemcpyAsync(h->d_step, h->h_step, ints * sizeof(int), cudaMemcpyHostToDevice, s));
    CU_CHECK(cudaEventRecord(r->ev[r->cur], s));
    r->pending[r->cur] = true;
    h->step.uploaded = 1;
    h->step_uploads++;
    return PA_OK;
}

int pa_cu_ensure_stage(pa_handle* h, size_t floats) {
    if (h->stage_floats >= floats) return PA_OK;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    CU_CHECK(cudaStreamSynchronize((cudaStream_t)h->stream));
    if (h->h_stage) cudaFreeHost(h->h_stage);
    cudaFree(h->d_stage);
    h->h_stage = NULL; h->d_stage = NULL; h->stage_floats = 0;
    size_t cap = floats + floats / 4;
    CU_CHECK(cudaMallocHost((void**)&h->h_stage, cap * sizeof(float)));
    CU_CHECK(cudaMalloc((void**)&h->d_stage, cap * sizeof(float)));
    h->stage_floats = cap;
    return PA_OK;
}

int pa_cu_copy_page_rows(pa_handle* h, int src_page, int dst_page, int rows) {
    if (h->host_only || !h->pool_k || rows <= 0) return PA_OK;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)h->stream;
    const size_t page = (size_t)h->cfg.block_size * h->C, bytes = (size_t)rows * h->C * sizeof(float);
    for (int l = 0; l < h->cfg.n_layers; l++) {
        const size_t base = (size_t)l * h->layer_stride;
        CU_CHECK(cudaMemcpyAsync(h->pool_k + base + dst_page * page, h->pool_k + base + src_page * page, bytes, cudaMemcpyDeviceToDevice, s));
        CU_CHECK(cudaMemcpyAsync(h->pool_v + base + dst_page * page, h->pool_v + base + src_page * page, bytes, cudaMemcpyDeviceToDevice, s));
    }
    CU_CHECK(cudaStreamSynchronize(s));
    return PA_OK;
}

int pa_cu_swap_page(pa_handle* h, int page, float* host_k, float* host_v, int to_host) {
    if (h->host_only || !h->pool_k) return PA_OK;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)h->stream;
    const size_t page_bytes = (size_t)h->cfg.block_size * h->C * sizeof(float);
    const size_t lpitch = h->layer_stride * sizeof(float);
    float* dk = h->pool_k + (size_t)page * h->cfg.block_size * h->C;
    float* dv = h->pool_v + (size_t)page * h->cfg.block_size * h->C;
    /* one 2-D copy moves a page of every layer (pitch = layer stride) while the pitch is one the runtime
     * accepts (cudaDeviceProp::memPitch, ~2 GiB: pools beyond ~700 k tokens at C = 768 exceed it); beyond
     * that, one plain copy per layer */
    if (lpitch <= h->max_pitch) {
        if (to_host) {
            CU_CHECK(cudaMemcpy2DAsync(host_k, page_bytes, dk, lpitch, page_bytes, h->cfg.n_layers, cudaMemcpyDeviceToHost, s));
            CU_CHECK(cudaMemcpy2DAsync(host_v, page_bytes, dv, lpitch, page_bytes, h->cfg.n_layers, cudaMemcpyDeviceToHost, s));
        } else {
            CU_CHECK(cudaMemcpy2DAsync(dk, lpitch, host_k, page_bytes, page_bytes, h->cfg.n_layers, cudaMemcpyHostToDevice, s));
            CU_CHECK(cudaMemcpy2DAsync(dv, lpitch, host_v, page_bytes, page_bytes, h->cfg.n_layers, cudaMemcpyHostToDevice, s));
        }
    } else {
        const size_t pf = page_bytes / sizeof(float);
        for (int l = 0; l < h->cfg.n_layers; l++) {
            float* lk = dk + (size_t)l * h->layer_stride;
            float* lv = dv + (size_t)l * h->layer_stride;
            if (to_host) {
                CU_CHECK(cudaMemcpyAsync(host_k + l * pf, lk, page_bytes, cudaMemcpyDeviceToHost, s));
                CU_CHECK(cudaMemcpyAsync(host_v + l * pf, lv, page_bytes, cudaMemcpyDeviceToHost, s));
            } else {
                CU_CHECK(cudaMemcpyAsync(lk, host_k + l * pf, page_bytes, cudaMemcpyHostToDevice, s));
                CU_CHECK(cudaMemcpyAsync(lv, host_v + l * pf, page_bytes, cudaMemcpyHostToDevice, s));
            }
        }
    }
    return PA_OK;          /* enqueued on the handle's stream; pa_cu_swap_sync waits */
}
int pa_cu_swap_sync(pa_handle* h) {
    if (h->host_only || !h->pool_k) return PA_OK;
    CU_CHECK(cudaSetDevice(h->cfg.device));
    CU_CHECK(cudaStreamSynchronize((cudaStream_t)h->stream));
    return PA_OK;
}

int pa_cu_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

/* ---- plumbing for plain-C hosts ------------------------------------------------------------ */
void* pa_dev_alloc(size_t bytes) {
    void* p = NULL;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        pa_set_error("pa_dev_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    return p;
}
void pa_dev_free(void* p) { if (p) cudaFree(p); }
void* pa_host_alloc(size_t bytes) {
    void* p = NULL;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        pa_set_error("pa_host_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    return p;
}
/* bumped whenever pinned memory goes away: remembered (host pointer -> device alias) pairs are only trusted within a generation */
unsigned pa_host_free_generation = 0;
void pa_host_free(void* p) { if (p) { ++pa_host_free_generation; cudaFreeHost(p); } }
int pa_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    CU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (!stream) CU_CHECK(cudaStreamSynchronize(0));
    return PA_OK;
}
int pa_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    CU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (!stream) CU_CHECK(cudaStreamSynchronize(0));
    return PA_OK;
}
int pa_memset(void* dst, int value, size_t bytes, void* stream) {
    CU_CHECK(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return PA_OK;
}
void* pa_stream_create(void) {
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return (void*)s;
}
void pa_stream_destroy(void* stream) {
    if (!stream) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaStreamSynchronize((cudaStream_t)stream);
    pa_cu_gemm_stream_released(dev, stream);
    cudaStreamDestroy((cudaStream_t)stream);
}
int pa_stream_sync(void* stream) { CU_CHECK(cudaStreamSynchronize((cudaStream_t)stream)); return PA_OK; }
int pa_device_sync(void) { CU_CHECK(cudaDeviceSynchronize()); return PA_OK; }
void* pa_event_create(void) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return (void*)e;
}
void pa_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int pa_event_record(void* ev, void* stream) { CU_CHECK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream)); return PA_OK; }
float pa_event_elapsed_ms(void* start, void* stop) {
    float ms = -1.0f;
    if (cudaEventSynchronize((cudaEvent_t)stop) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    return ms;
}
int pa_flush_l2(void* scratch, size_t bytes, void* stream) {
    CU_CHECK(cudaMemsetAsync(scratch, 0, bytes, (cudaStream_t)stream));
    return PA_OK;
}

}  // extern "C"
