/*
 * pa_group.cu -- the one exchange step of the path (north_star: "NCCL over NVLink is used only to gather
 * sampled tokens or logits"): sequences are sharded over the GPUs of one box, every GPU owns its
 * block manager, page pool and tables, and after the sampler the ranks all-gather their int32 next tokens
 * (or their last-position logits when sampling is centralised).
 *
 * Two ways to form a group, same calls afterwards:
 *   (a) pa_group_create  -- ONE process drives n GPUs (plain-C hosts: examples/generate_multi.c): one handle
 *       and one stream per GPU, ncclCommInitAll;
 *   (b) pa_group_join    -- one process per GPU (torchrun, MPI): rank 0 makes a 128-byte id
 *       (pa_comm_unique_id), the launcher hands it to every rank, each joins with its own handle
 *       (ncclCommInitRank).
 * The gather is a plain ncclAllGather ENQUEUED ON THE HANDLE'S STREAM behind the sampler kernel: no host
 * synchronisation between the model step and the collective; the host waits once per step.
 *
 * NCCL is not a link-time dependency: libnccl.so.2 is opened on first use (a process that already has one
 * loaded -- torch's -- shares it), so single-GPU hosts without NCCL can still load libpaged_attn.so.
 * Reference precedent: none in C (the only collective use is Python DDP, train_gpt2.py:400-412).
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_open() {
    const char* names[] = {getenv("PA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
#define PA_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name))
    PA_SYM(GetUniqueId, "ncclGetUniqueId");
    PA_SYM(CommInitRank, "ncclCommInitRank");
    PA_SYM(CommInitAll, "ncclCommInitAll");
    PA_SYM(CommDestroy, "ncclCommDestroy");
    PA_SYM(AllGather, "ncclAllGather");
    PA_SYM(GroupStart, "ncclGroupStart");
    PA_SYM(GroupEnd, "ncclGroupEnd");
    PA_SYM(GetErrorString, "ncclGetErrorString");
    PA_SYM(GetVersion, "ncclGetVersion");
#undef PA_SYM
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllGather ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.GetErrorString) {
        dlclose(g_nccl.lib);
        g_nccl.lib = nullptr;
    }
}
int nccl_ready(const char* who) {
    std::call_once(g_nccl_once, nccl_open);
    if (!g_nccl.lib) {
        pa_set_error("%s: libnccl.so.2 could not be opened (%s); multi-GPU groups need NCCL (PA_NCCL_LIB names another path)", who,
                     dlerror() ? dlerror() : "symbols missing");
        return PA_ERR_UNSUPPORTED;
    }
    return PA_OK;
}
#define NCCL_CHECK(call)                                                                             \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) {                                                                     \
            pa_set_error("%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__);     \
            return PA_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

}  // namespace

struct pa_group {
    int world = 1, rank0 = 0;             // world size; rank of local member 0 (local members hold consecutive ranks)
    bool owns_handles = false;            // (a): the group created the handles
    std::vector<pa_handle*> members;
    std::vector<ncclComm_t> comms;
    // per local member: device + pinned buffers of world * cap ints for the gathered tokens
    std::vector<int*> d_all, h_all;
    int cap = 0;
    // overlapped gather (pa_group_model_step_overlapped): per member a side stream, and per step parity a staging
    // copy of the member's own tokens, a gathered buffer (device + pinned) and two events
    struct Side {
        cudaStream_t stream = nullptr;
        int* d_local[2] = {nullptr, nullptr};
        int* d_all[2] = {nullptr, nullptr};
        int* h_all[2] = {nullptr, nullptr};
        cudaEvent_t sampled[2] = {nullptr, nullptr}, gathered[2] = {nullptr, nullptr};
    };
    std::vector<Side> side;
    int side_cap = 0, parity = 0, pending = -1, pending_nseq = 0;      // pending: parity of a gather in flight, -1 none
};

extern "C" {

int pa_nccl_version(void) {
    if (nccl_ready("pa_nccl_version") != PA_OK || !g_nccl.GetVersion) return 0;
    int v = 0;
    return g_nccl.GetVersion(&v) == ncclSuccess ? v : 0;
}

int pa_comm_unique_id(void* id128) {
    if (!id128) { pa_set_error("pa_comm_unique_id: NULL buffer"); return PA_ERR_INVALID; }
    int rc = nccl_ready("pa_comm_unique_id");
    if (rc != PA_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == PA_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return PA_OK;
}

static int group_buffers(pa_group* g, int cap) {
    if (cap <= g->cap) return PA_OK;
    for (size_t i = 0; i < g->members.size(); ++i) {
        CU_CHECK(cudaSetDevice(g->members[i]->cfg.device));
        CU_CHECK(cudaStreamSynchronize((cudaStream_t)g->members[i]->stream));
        if (g->d_all[i]) cudaFree(g->d_all[i]);
        if (g->h_all[i]) cudaFreeHost(g->h_all[i]);
        g->d_all[i] = nullptr; g->h_all[i] = nullptr;
        CU_CHECK(cudaMalloc((void**)&g->d_all[i], (size_t)g->world * cap * sizeof(int)));
        CU_CHECK(cudaMallocHost((void**)&g->h_all[i], (size_t)g->world * cap * sizeof(int)));
    }
    g->cap = cap;
    return PA_OK;
}

int pa_group_create(const pa_config* cfg, int n_gpus, const int* devices, pa_group** out) {
    if (!cfg || !out || n_gpus < 1) { pa_set_error("pa_group_create: bad arguments"); return PA_ERR_INVALID; }
    *out = nullptr;
    if (n_gpus > pa_device_count()) { pa_set_error("pa_group_create: %d GPUs asked for, %d visible", n_gpus, pa_device_count()); return PA_ERR_NO_DEVICE; }
    int rc = n_gpus > 1 ? nccl_ready("pa_group_create") : PA_OK;
    if (rc != PA_OK) return rc;
    pa_group* g = new pa_group();
    g->world = n_gpus; g->rank0 = 0; g->owns_handles = true;
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) {
        devs[i] = devices ? devices[i] : i;
        pa_config c = *cfg;
        c.device = devs[i];
        pa_handle* h = nullptr;
        rc = pa_create(&c, &h);
        if (rc != PA_OK) { pa_group_destroy(g); return rc; }
        g->members.push_back(h);
    }
    g->comms.assign(n_gpus, nullptr);
    g->d_all.assign(n_gpus, nullptr);
    g->h_all.assign(n_gpus, nullptr);
    if (n_gpus > 1) {
        ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), n_gpus, devs.data());
        if (r != ncclSuccess) {
            pa_set_error("pa_group_create: ncclCommInitAll: %s", g_nccl.GetErrorString(r));
            pa_group_destroy(g);
            return PA_ERR_CUDA;
        }
    }
    *out = g;
    return PA_OK;
}

int pa_group_join(pa_handle* h, const void* id128, int rank, int world, pa_group** out) {
    if (!h || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) { pa_set_error("pa_group_join: bad arguments"); return PA_ERR_INVALID; }
    *out = nullptr;
    if (h->host_only) { pa_set_error("pa_group_join: host-only handle"); return PA_ERR_NO_DEVICE; }
    int rc = world > 1 ? nccl_ready("pa_group_join") : PA_OK;
    if (rc != PA_OK) return rc;
    pa_group* g = new pa_group();
    g->world = world; g->rank0 = rank; g->owns_handles = false;
    g->members.push_back(h);
    g->comms.assign(1, nullptr);
    g->d_all.assign(1, nullptr);
    g->h_all.assign(1, nullptr);
    if (world > 1) {
        CU_CHECK(cudaSetDevice(h->cfg.device));
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&g->comms[0], world, id, rank);
        if (r != ncclSuccess) {
            pa_set_error("pa_group_join: ncclCommInitRank: %s", g_nccl.GetErrorString(r));
            delete g;
            return PA_ERR_CUDA;
        }
    }
    *out = g;
    return PA_OK;
}

void pa_group_destroy(pa_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->members.size(); ++i) {
        pa_handle* h = g->members[i];
        if (h && !h->host_only) {
            cudaSetDevice(h->cfg.device);
            cudaStreamSynchronize((cudaStream_t)h->stream);
        }
        if (i < g->comms.size() && g->comms[i]) g_nccl.CommDestroy(g->comms[i]);
        if (i < g->d_all.size() && g->d_all[i]) cudaFree(g->d_all[i]);
        if (i < g->h_all.size() && g->h_all[i]) cudaFreeHost(g->h_all[i]);
        if (i < g->side.size()) {
            pa_group::Side& sd = g->side[i];
            if (sd.stream) cudaStreamSynchronize(sd.stream);
            for (int b = 0; b < 2; ++b) {
                if (sd.d_local[b]) cudaFree(sd.d_local[b]);
                if (sd.d_all[b]) cudaFree(sd.d_all[b]);
                if (sd.h_all[b]) cudaFreeHost(sd.h_all[b]);
                if (sd.sampled[b]) cudaEventDestroy(sd.sampled[b]);
                if (sd.gathered[b]) cudaEventDestroy(sd.gathered[b]);
            }
            if (sd.stream) cudaStreamDestroy(sd.stream);
        }
        if (g->owns_handles) pa_destroy(h);
    }
    delete g;
}

int pa_group_size(pa_group* g) { return g ? g->world : 0; }
int pa_group_local_count(pa_group* g) { return g ? (int)g->members.size() : 0; }
int pa_group_rank(pa_group* g, int i) { return (g && i >= 0 && i < (int)g->members.size()) ? g->rank0 + i : -1; }
pa_handle* pa_group_handle(pa_group* g, int i) { return (g && i >= 0 && i < (int)g->members.size()) ? g->members[i] : nullptr; }

static int gather_impl(pa_group* g, const void* const* send, void* const* recv, size_t count, ncclDataType_t type, size_t elt, const char* who) {
    if (!g || !send || !recv) { pa_set_error("%s: bad arguments", who); return PA_ERR_INVALID; }
    const int n = (int)g->members.size();
    if (g->world == 1) {        // a group of one: the "gather" is a copy on the member's stream
        CU_CHECK(cudaSetDevice(g->members[0]->cfg.device));
        if (send[0] != recv[0]) CU_CHECK(cudaMemcpyAsync(recv[0], send[0], count * elt, cudaMemcpyDeviceToDevice, (cudaStream_t)g->members[0]->stream));
        return PA_OK;
    }
    if (n > 1) NCCL_CHECK(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        // (with one communicator per device in ONE process the calls of a collective must sit inside a group)
        ncclResult_t r = g_nccl.AllGather(send[i], recv[i], count, type, g->comms[i], (cudaStream_t)g->members[i]->stream);
        if (r != ncclSuccess) {
            if (n > 1) g_nccl.GroupEnd();
            pa_set_error("%s: ncclAllGather: %s", who, g_nccl.GetErrorString(r));
            return PA_ERR_CUDA;
        }
        g->members[i]->launches++;
    }
    if (n > 1) NCCL_CHECK(g_nccl.GroupEnd());
    return PA_OK;
}

/* send[i]: n_per_rank int32 in device memory of local member i; recv[i]: world * n_per_rank int32 there,
 * rank-major.  Stream-ordered on each member's stream; returns without synchronising. */
int pa_group_gather_tokens(pa_group* g, const int* const* send, int* const* recv, int n_per_rank) {
    if (n_per_rank < 1) { pa_set_error("pa_group_gather_tokens: n_per_rank < 1"); return PA_ERR_INVALID; }
    return gather_impl(g, (const void* const*)send, (void* const*)recv, (size_t)n_per_rank, ncclInt32, sizeof(int), "pa_group_gather_tokens");
}
/* the same for fp32 rows (each rank's last-position logits: n_floats_per_rank = nseq * row stride) */
int pa_group_gather_logits(pa_group* g, const float* const* send, float* const* recv, size_t n_floats_per_rank) {
    if (n_floats_per_rank < 1) { pa_set_error("pa_group_gather_logits: nothing to gather"); return PA_ERR_INVALID; }
    return gather_impl(g, (const void* const*)send, (void* const*)recv, n_floats_per_rank, ncclFloat32, sizeof(float), "pa_group_gather_logits");
}

/* One decode step of the whole group: every local member's model takes one new token per sequence
 * (nseq sequences each: seq_ids[i], tokens[i], coins[i] or NULL for argmax), the sampled tokens of ALL
 * ranks are gathered behind the sampler on the members' streams, and the host waits once per member.
 * all_next: world * nseq ints, rank-major (what every rank needs to feed the next step / detokenise). */
int pa_group_model_step(pa_group* g, pa_model* const* models, const int* const* seq_ids, const int* const* tokens,
                        const float* const* coins, int nseq, int* all_next) {
    if (!g || !models || !seq_ids || !tokens || nseq < 1 || !all_next) { pa_set_error("pa_group_model_step: bad arguments"); return PA_ERR_INVALID; }
    const int n = (int)g->members.size();
    int rc = group_buffers(g, nseq);
    if (rc != PA_OK) return rc;
    std::vector<int> ones(nseq, 1);
    std::vector<const int*> send(n);
    for (int i = 0; i < n; ++i) {
        if (!models[i] || pa_model_handle(models[i]) != g->members[i]) { pa_set_error("pa_group_model_step: model %d does not belong to member %d", i, i); return PA_ERR_INVALID; }
        pa_model_want_device_tokens(models[i], 1);
        rc = pa_model_forward_async(models[i], seq_ids[i], ones.data(), tokens[i], coins ? coins[i] : nullptr, nseq);
        if (rc != PA_OK) {
            for (int j = 0; j < i; ++j) pa_model_wait(models[j], nullptr);
            return rc;
        }
        send[i] = pa_model_next_tokens_dev(models[i]);
    }
    rc = pa_group_gather_tokens(g, send.data(), g->d_all.data(), nseq);
    for (int i = 0; i < n && rc == PA_OK; ++i) {
        if (cudaSetDevice(g->members[i]->cfg.device) != cudaSuccess ||
            cudaMemcpyAsync(g->h_all[i], g->d_all[i], (size_t)g->world * nseq * sizeof(int), cudaMemcpyDeviceToHost,
                            (cudaStream_t)g->members[i]->stream) != cudaSuccess) {
            pa_set_error("pa_group_model_step: copy of the gathered tokens failed");
            rc = PA_ERR_CUDA;
        }
    }
    for (int i = 0; i < n; ++i) {
        const int rw = pa_model_wait(models[i], nullptr);
        if (rc == PA_OK) rc = rw;
    }
    if (rc == PA_OK) memcpy(all_next, g->h_all[0], (size_t)g->world * nseq * sizeof(int));
    return rc;
}


static int side_buffers(pa_group* g, int cap) {
    const int n = (int)g->members.size();
    if ((int)g->side.size() != n) g->side.assign(n, pa_group::Side());
    if (cap <= g->side_cap) return PA_OK;
    for (int i = 0; i < n; ++i) {
        pa_group::Side& sd = g->side[i];
        CU_CHECK(cudaSetDevice(g->members[i]->cfg.device));
        if (!sd.stream) CU_CHECK(cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking));
        CU_CHECK(cudaStreamSynchronize(sd.stream));
        CU_CHECK(cudaStreamSynchronize((cudaStream_t)g->members[i]->stream));
        for (int b = 0; b < 2; ++b) {
            if (sd.d_local[b]) cudaFree(sd.d_local[b]);
            if (sd.d_all[b]) cudaFree(sd.d_all[b]);
            if (sd.h_all[b]) cudaFreeHost(sd.h_all[b]);
            sd.d_local[b] = sd.d_all[b] = sd.h_all[b] = nullptr;
            CU_CHECK(cudaMalloc((void**)&sd.d_local[b], (size_t)cap * sizeof(int)));
            CU_CHECK(cudaMalloc((void**)&sd.d_all[b], (size_t)g->world * cap * sizeof(int)));
            CU_CHECK(cudaMallocHost((void**)&sd.h_all[b], (size_t)g->world * cap * sizeof(int)));
            if (!sd.sampled[b]) CU_CHECK(cudaEventCreateWithFlags(&sd.sampled[b], cudaEventDisableTiming));
            if (!sd.gathered[b]) CU_CHECK(cudaEventCreateWithFlags(&sd.gathered[b], cudaEventDisableTiming));
        }
    }
    g->side_cap = cap;
    return PA_OK;
}

/* wait for the all-gather still in flight (if any) and hand out its tokens: size * nseq ints, rank-major */
int pa_group_gather_flush(pa_group* g, int* gathered) {
    if (!g) { pa_set_error("pa_group_gather_flush: NULL group"); return PA_ERR_INVALID; }
    if (g->pending < 0) return 0;
    const int b = g->pending, nseq = g->pending_nseq;
    g->pending = -1;
    for (size_t i = 0; i < g->members.size(); ++i) {
        CU_CHECK(cudaSetDevice(g->members[i]->cfg.device));
        CU_CHECK(cudaEventSynchronize(g->side[i].gathered[b]));
    }
    if (gathered) memcpy(gathered, g->side[0].h_all[b], (size_t)g->world * nseq * sizeof(int));
    return nseq;
}

/* The decode step of the whole group with the gather OFF the critical path.  A rank's next step consumes only its
 * OWN sampled tokens; the tokens of the other ranks are output (detokenising, logging, stop conditions).  So: every
 * local member's model takes its step and hands back its own tokens (next_local[i], nseq ints) after ONE wait on
 * its stream; the all-gather of those tokens runs on a side stream behind the sampler -- beside the NEXT step's
 * kernels -- and its result is handed out by the next call (gathered_prev: size * nseq ints of the PREVIOUS step,
 * rank-major; untouched on the first call) or by pa_group_gather_flush.  Returns the number of sequences per rank
 * that gathered_prev holds (0 on the first call), or a negative pa_status. */
int pa_group_model_step_overlapped(pa_group* g, pa_model* const* models, const int* const* seq_ids, const int* const* tokens,
                                   const float* const* coins, int nseq, int* const* next_local, int* gathered_prev) {
    if (!g || !models || !seq_ids || !tokens || nseq < 1 || !next_local) { pa_set_error("pa_group_model_step_overlapped: bad arguments"); return PA_ERR_INVALID; }
    const int n = (int)g->members.size();
    int rc = side_buffers(g, nseq);
    if (rc != PA_OK) return rc;
    const int b = g->parity;
    std::vector<int> ones(nseq, 1);
    for (int i = 0; i < n; ++i) {
        if (!models[i] || pa_model_handle(models[i]) != g->members[i]) { pa_set_error("pa_group_model_step_overlapped: model %d does not belong to member %d", i, i); return PA_ERR_INVALID; }
        pa_model_want_device_tokens(models[i], 1);
        rc = pa_model_forward_async(models[i], seq_ids[i], ones.data(), tokens[i], coins ? coins[i] : nullptr, nseq);
        if (rc != PA_OK) {
            for (int j = 0; j < i; ++j) pa_model_wait(models[j], nullptr);
            return rc;
        }
        // the member's own tokens leave the model's buffer (the next step's sampler overwrites it) on the main stream ...
        pa_group::Side& sd = g->side[i];
        cudaStream_t ms = (cudaStream_t)g->members[i]->stream;
        if (cudaSetDevice(g->members[i]->cfg.device) != cudaSuccess ||
            cudaMemcpyAsync(sd.d_local[b], pa_model_next_tokens_dev(models[i]), (size_t)nseq * sizeof(int), cudaMemcpyDeviceToDevice, ms) != cudaSuccess ||
            cudaEventRecord(sd.sampled[b], ms) != cudaSuccess || cudaStreamWaitEvent(sd.stream, sd.sampled[b], 0) != cudaSuccess) {
            pa_set_error("pa_group_model_step_overlapped: staging the sampled tokens failed");
            for (int j = 0; j <= i; ++j) pa_model_wait(models[j], nullptr);
            return PA_ERR_CUDA;
        }
    }
    // ... and are gathered on the side streams
    if (g->world == 1) {
        rc = cudaMemcpyAsync(g->side[0].d_all[b], g->side[0].d_local[b], (size_t)nseq * sizeof(int), cudaMemcpyDeviceToDevice, g->side[0].stream) == cudaSuccess ? PA_OK : PA_ERR_CUDA;
    } else {
        if (n > 1) g_nccl.GroupStart();
        for (int i = 0; i < n && rc == PA_OK; ++i) {
            ncclResult_t r = g_nccl.AllGather(g->side[i].d_local[b], g->side[i].d_all[b], (size_t)nseq, ncclInt32, g->comms[i], g->side[i].stream);
            if (r != ncclSuccess) { pa_set_error("pa_group_model_step_overlapped: ncclAllGather: %s", g_nccl.GetErrorString(r)); rc = PA_ERR_CUDA; }
            g->members[i]->launches++;
        }
        if (n > 1) g_nccl.GroupEnd();
    }
    for (int i = 0; i < n && rc == PA_OK; ++i) {
        pa_group::Side& sd = g->side[i];
        if (cudaSetDevice(g->members[i]->cfg.device) != cudaSuccess ||
            cudaMemcpyAsync(sd.h_all[b], sd.d_all[b], (size_t)g->world * nseq * sizeof(int), cudaMemcpyDeviceToHost, sd.stream) != cudaSuccess ||
            cudaEventRecord(sd.gathered[b], sd.stream) != cudaSuccess) {
            pa_set_error("pa_group_model_step_overlapped: copy of the gathered tokens failed");
            rc = PA_ERR_CUDA;
        }
    }
    // the PREVIOUS step's gather has had this whole step's queueing time (and its own step) to finish
    int got_prev = 0;
    if (rc == PA_OK) {
        got_prev = pa_group_gather_flush(g, gathered_prev);
        if (got_prev < 0) rc = got_prev;
    }
    for (int i = 0; i < n; ++i) {
        const int rw = pa_model_wait(models[i], next_local[i]);
        if (rc == PA_OK) rc = rw;
    }
    if (rc != PA_OK) return rc;
    g->pending = b;
    g->pending_nseq = nseq;
    g->parity ^= 1;
    return got_prev;
}

}  // extern "C"
