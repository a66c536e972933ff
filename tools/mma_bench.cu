// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128, K=8) as a function of N, operand
// source (A from TMEM or shared memory) and accumulator dependence.  One CTA per SM, one issuing
// thread.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = TS one accumulator, 1 = TS two accumulators alternating, 2 = SS one accumulator, 3 = bf16 SS (K=16)
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float*)base)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (mode >= 4) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t b_smem = smem_u32(base + 32 * 1024);
        const int warp = threadIdx.x >> 5;
        const int n_issuers = mode == 5 ? 2 : 1;
        long long t0 = 0, t1 = 0, t2 = 0;
        if (warp < n_issuers) {
            const uint32_t d0 = tm + warp * 128;
            const uint64_t b0 = desc_k128(b_smem), b1 = desc_k128(b_smem + 32), b2 = desc_k128(b_smem + 64), b3 = desc_k128(b_smem + 96);
            uint32_t pred = 0;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
            t0 = clock64();
            for (int i = 0; i < reps; i += 4) {
                if (pred) {
                    const uint32_t dd = mode == 6 ? tm + ((i >> 2) & 3) * 64 : d0;
                    mma_ts(dd, tm + 448, b0, idesc, i > 12);
                    mma_ts(dd, tm + 456, b1, idesc, 1);
                    mma_ts(dd, tm + 464, b2, idesc, 1);
                    mma_ts(dd, tm + 472, b3, idesc, 1);
                }
                __syncwarp();
            }
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
            if (warp == 0) {
                asm volatile("{\n.reg .pred P1;\nW4:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D4;\nbra W4;\nD4:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
                t2 = clock64();
                if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = (t1 - t0) / n_issuers; out[1] = (t2 - t0) / n_issuers; }
            }
        }
    } else if (threadIdx.x == 0) {
        const uint32_t fmt = mode == 3 ? 1u : 2u;      // bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a_smem = smem_u32(base), b_smem = smem_u32(base + 32 * 1024);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            const uint32_t koff = (i & 3) * 32;
            if (mode == 0) mma_ts(tm, tm + 448 + (i & 3) * 8, desc_k128(b_smem + koff), idesc, i > 0);
            else if (mode == 1) mma_ts(tm + (i & 1) * 256 * 0 + (i & 1) * (N <= 128 ? 128 : 0), tm + 448 + (i & 3) * 8, desc_k128(b_smem + koff), idesc, i > 1);
            else if (mode == 2) mma_ss(tm, desc_k128(a_smem + koff), desc_k128(b_smem + koff), idesc, i > 0);
            else mma_f16_ss(tm, desc_k128(a_smem + koff), desc_k128(b_smem + koff), idesc, i > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[] = {"tf32 A=TMEM 1 acc", "tf32 A=TMEM 2 acc", "tf32 A=smem 1 acc", "bf16 A=smem 1 acc (K=16)", "tf32 TS elect unrolled", "tf32 TS 2 issuing warps", "tf32 TS 4 acc unrolled"};
    for (int mode = 4; mode >= 4 && mode < 5; ++mode)
        for (int N : {16, 32, 64, 128, 256}) {
            if ((mode == 1 || mode == 5) && N > 128) continue;
            if (mode == 6 && N > 64) continue;
            if (mode < 4 && N != 64 && N != 256) continue;
            printf("mode %d N %d ...\n", mode, N); fflush(stdout);
            const int reps = 512;
            long long h[2];
            for (int it = 0; it < 2; ++it) {
                bench<<<148, 128, 100 * 1024>>>(N, mode, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-26s N=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (floor %5.1f)\n", names[mode], N, (double)h[0] / reps,
                   (double)h[1] / reps, 128.0 * N / 256.0);
            fflush(stdout);
        }
    return 0;
}
