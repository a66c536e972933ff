// tools/membench.cu -- "speed of light" probe for the paged-read access pattern on this box.
// Streams n_pages pages of page_bytes each (shuffled order, 12 disjoint regions rotated to defeat
// L2) with (a) the TMA bulk-copy + mbarrier ring used by pa_decode_stream_kernel but NO compute,
// (b) plain 16-byte LDGs.  Prints GB/s so the decode kernel's achieved bandwidth can be read
// against what the memory system gives this pattern.   Build: see tools/Makefile
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// chunk_bytes: size of each bulk copy (page split into page_bytes/chunk_bytes copies on one barrier)
__global__ void tma_stream(const char* base, const int* order, int n_pages, int page_bytes, int chunk_bytes,
                           int stages, float* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = (uint64_t*)(sm + (size_t)stages * page_bytes);
    const int per = (n_pages + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n_pages, p0 + per);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mb_init(s32(&bars[s]), 1); mb_init(s32(&bars[stages + s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {            // producer
        int st = 0; uint32_t ph = 0;
        for (int p = p0; p < p1; p++) {
            mb_wait(s32(&bars[stages + st]), ph ^ 1);
            mb_expect(s32(&bars[st]), page_bytes);
            const char* src = base + (size_t)order[p] * page_bytes;
            for (int c = 0; c < page_bytes; c += chunk_bytes) bulk(s32(sm + (size_t)st * page_bytes + c), src + c, chunk_bytes, s32(&bars[st]));
            if (++st == stages) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 0) {      // consumer: touch one word, release
        int st = 0; uint32_t ph = 0; float acc = 0;
        for (int p = p0; p < p1; p++) {
            mb_wait(s32(&bars[st]), ph);
            acc += *(float*)(sm + (size_t)st * page_bytes);
            mb_arrive(s32(&bars[stages + st]));
            if (++st == stages) { st = 0; ph ^= 1; }
        }
        if (acc == 123.456f) *sink = acc;
    }
}

__global__ void ldg_stream(const char* base, const int* order, int n_pages, int page_bytes, float* sink) {
    const int per = (n_pages + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n_pages, p0 + per);
    float acc = 0;
    for (int p = p0; p < p1; p++) {
        const float4* src = (const float4*)(base + (size_t)order[p] * page_bytes);
        const int n4 = page_bytes / 16;
        for (int i = threadIdx.x; i < n4; i += blockDim.x * 4) {
            float4 a = __ldg(src + i), b = make_float4(0, 0, 0, 0), c = b, d = b;
            if (i + blockDim.x < n4) b = __ldg(src + i + blockDim.x);
            if (i + 2 * blockDim.x < n4) c = __ldg(src + i + 2 * blockDim.x);
            if (i + 3 * blockDim.x < n4) d = __ldg(src + i + 3 * blockDim.x);
            acc += a.x + b.y + c.z + d.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main(int argc, char** argv) {
    const int page_bytes = argc > 1 ? atoi(argv[1]) : 49152;
    const int n_pages = argc > 2 ? atoi(argv[2]) : 8192;      // 403 MB like cfg2 (K+V pages)
    const int regions = 12;
    const size_t region_bytes = (size_t)n_pages * page_bytes;
    char* buf; CK(cudaMalloc(&buf, region_bytes * regions)); CK(cudaMemset(buf, 1, region_bytes * regions));
    float* sink; CK(cudaMalloc(&sink, 4));
    std::vector<int> seq(n_pages), shuf(n_pages);
    for (int i = 0; i < n_pages; i++) seq[i] = shuf[i] = i;
    std::mt19937 g(1234); std::shuffle(shuf.begin(), shuf.end(), g);
    int *d_seq, *d_shuf; CK(cudaMalloc(&d_seq, n_pages * 4)); CK(cudaMalloc(&d_shuf, n_pages * 4));
    CK(cudaMemcpy(d_seq, seq.data(), n_pages * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_shuf, shuf.data(), n_pages * 4, cudaMemcpyHostToDevice));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    CK(cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
        for (int r = 0; r < regions; r++) launch(r);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 4 * regions;
        for (int r = 0; r < reps; r++) launch(r % regions);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
        printf("%-60s %8.1f GB/s  %7.2f us/launch\n", name, region_bytes * (double)reps / (ms * 1e-3) / 1e9, ms * 1e3 / reps);
    };
    char name[256];
    for (int shuffled = 0; shuffled < 2; shuffled++) {
        const int* order = shuffled ? d_shuf : d_seq;
        for (int stages : {2, 3, 4}) for (int chunk : {page_bytes, page_bytes / 4, page_bytes / 16}) for (int grid : {sms, 2 * sms}) {
            size_t smem = (size_t)stages * page_bytes + 2 * stages * 8 + 16;
            if (grid == 2 * sms && smem * 2 > prop.sharedMemPerMultiprocessor) continue;
            if (smem > prop.sharedMemPerBlockOptin) continue;
            snprintf(name, sizeof name, "tma %s page=%d stages=%d chunk=%d grid=%d", shuffled ? "shuffled" : "sequential", page_bytes, stages, chunk, grid);
            run(name, [&](int r) { tma_stream<<<grid, 64, smem>>>(buf + (size_t)r * region_bytes, order, n_pages, page_bytes, chunk, stages, sink); });
        }
        for (int tpb : {256, 512, 1024}) for (int mult : {1, 2, 4}) {
            snprintf(name, sizeof name, "ldg %s page=%d threads=%d grid=%d", shuffled ? "shuffled" : "sequential", page_bytes, tpb, sms * mult);
            run(name, [&](int r) { ldg_stream<<<sms * mult, tpb>>>(buf + (size_t)r * region_bytes, order, n_pages, page_bytes, sink); });
        }
    }
    return 0;
}
