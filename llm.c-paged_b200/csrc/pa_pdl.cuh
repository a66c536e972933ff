/*
 * pa_pdl.cuh -- programmatic dependent launch for the chain of small kernels of a decode step.
 * A kernel launched through pa_launch_pdl may start (be scheduled, run its prologue) while its
 * predecessor in the stream is still running; it MUST execute pdl_wait() before it reads anything
 * an earlier kernel wrote or writes anything an earlier kernel may read -- the wait returns once
 * the predecessor grid has completed and its memory is visible (transitively the whole chain).
 * pdl_launch_dependents() at the top lets the successor's launch overlap this kernel.
 */
#pragma once
#include <cuda_runtime.h>

#include <atomic>

// More than 48 KB of dynamic shared memory is a per-DEVICE opt-in of a kernel: `done` is the call site's (one
// per kernel instantiation) bit mask of the devices that already have it.  One process may drive several GPUs
// (pa_group_create), so a process-wide "done once" flag is not enough.
template <typename F>
static inline cudaError_t pa_optin_smem(std::atomic<unsigned long long>& done, F fn, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_relaxed) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_relaxed);
    return e;
}

extern "C" __thread int pa_pdl_enabled;     /* per host thread; set from the handle's PA_TUNE_NO_PDL at each launching entry */
extern "C" __thread int pa_pdl_gate;        /* per-step gate set by the caller of the chain (pa_model_forward) */

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern "C" __thread int pa_launch_cooperative;      /* set around a launch whose CTAs wait for each other (split-K through the L2 workspace) */

template <typename... KArgs, typename... Args>
static inline cudaError_t pa_launch_pdl(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        int cluster_z, Args... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    int n = 0;
    if (pa_pdl_enabled && pa_pdl_gate) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = 1;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = (unsigned)cluster_z;
        ++n;
    }
    if (pa_launch_cooperative) {        // co-residency of the whole grid is checked at launch instead of assumed
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, fn, KArgs(args)...);
}
