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
    if (mode >= 7) {
        // fresh operands: every instruction reads another 2 KB B slice (a 64 KB region, K-major SW128 tiles of 64 rows) and
        // another 8-column A slice (256 TMEM columns); mode 7 one issuing warp, mode 8 two (own accumulators, own barrier)
        __shared__ uint64_t bar2[2];
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t b_smem = smem_u32(base);
        const int warp = threadIdx.x >> 5;
        const int n_issuers = mode == 8 ? 2 : 1;
        __shared__ volatile int stop_flag;
        if (threadIdx.x == 0) stop_flag = 0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        long long t0 = 0, t1 = 0, t2 = 0;
        if (warp < n_issuers) {
            const uint32_t d0 = tm + warp * 64;                    // accumulators at columns 0 / 64
            const uint32_t a0 = tm + 128 + warp * 128;            // A slices at columns 128.. / 256..
            const uint32_t blo = ((b_smem + warp * 32768) & 0x3ffff) >> 4 | (1u << 16);
            const uint32_t bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
            uint32_t pred = 0;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
            t0 = clock64();
            for (int i = 0; i < reps; i += 16) {
                if (pred) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // slice j of a 32 KB region: tile (j >> 2) of 8 KB, 32-byte column block (j & 3)
                        const uint32_t lo = blo + ((((i >> 4) & 1) * 16384 + (j >> 2) * 8192 / 2 + (j & 3) * 32) >> 4);
                        asm volatile("{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}\n"
                                     ::"r"(d0), "r"(a0 + ((j * 8) & 127)), "r"(lo), "r"(bhi), "r"(idesc), "r"((uint32_t)(i + j > 0)) : "memory");
                    }
                }
                __syncwarp();
            }
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[warp])) : "memory");
            t1 = clock64();
            asm volatile("{\n.reg .pred P1;\nW7:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D7;\nbra W7;\nD7:\n}\n" ::"r"(smem_u32(&bar2[warp])) : "memory");
            t2 = clock64();
            if (blockIdx.x == 0 && warp == 0 && (threadIdx.x & 31) == 0) { out[0] = (t1 - t0) / n_issuers; out[1] = (t2 - t0) / n_issuers; }
            if (warp == 0) stop_flag = 1;
        } else if (mode == 9) {
            // interference: the other three warps copy 16-byte vectors around a 32 KB region of shared memory
            float4* reg = reinterpret_cast<float4*>(base + 64 * 1024);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int k = threadIdx.x;
            while (!stop_flag) {
#pragma unroll
                for (int u = 0; u < 8; ++u) { const float4 v = reg[(k + u * 96) & 2047]; acc.x += v.x; reg[(k + u * 96 + 1024) & 2047] = v; }
                k += 7;
            }
            if (acc.x == 123.456f) out[1] = 0;
        } else if (mode == 10) {
            // interference: the other three warps read and write 32-column groups of their TMEM lanes (columns 384..511)
            const uint32_t lane_base = tm + (((uint32_t)(warp & 3) * 32) << 16) + 384;
            uint32_t r[32];
            while (!stop_flag) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                 : "r"(lane_base + u * 32) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                                 ::"r"(lane_base + u * 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                                   "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
    } else if (mode >= 4) {
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
    const char* names[] = {"tf32 A=TMEM 1 acc", "tf32 A=TMEM 2 acc", "tf32 A=smem 1 acc", "bf16 A=smem 1 acc (K=16)", "tf32 TS elect unrolled", "tf32 TS 2 issuing warps", "tf32 TS 4 acc unrolled",
                           "tf32 TS fresh operands", "tf32 TS fresh, 2 issuers", "fresh + smem traffic", "fresh + TMEM ld/st traffic"};
    for (int mode : {7, 9, 10})
        for (int N : {16, 32, 64, 128, 256}) {
            if ((mode == 1 || mode == 5) && N > 128) continue;
            if (mode >= 7 && N != 64 && N != 128) continue;
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
