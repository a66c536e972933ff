/* pa_internal.h -- shared between the plain-C host side (pa_block_manager.c, pa_step.c,
 * pa_compat.c) and the CUDA side (pa_cuda.cu, pa_kernels.cu).  Not installed. */
#ifndef PA_INTERNAL_H
#define PA_INTERNAL_H

#include "paged_attn.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* step tables: ONE pinned int32 buffer mirrored with ONE cudaMemcpyAsync (pa_step_upload).
 * Offsets are in ints from the start of the buffer. */
typedef struct pa_step_layout {
    int nseq;          /* batch rows */
    int ntok;          /* new tokens in this step (rows of slot_mapping) */
    int tstride;       /* ints per block-table row in the mirror (pages actually used, padded) */
    int total_pages;   /* cum_pages[nseq] */
    int max_q;         /* max new tokens of one sequence */
    int off_kv_end;    /* [nseq]   cached tokens visible to the LAST query row (= context length) */
    int off_kv_start;  /* [nseq]   first visible token (sliding window; 0 by default) */
    int off_cum_pages; /* [nseq+1] prefix sum of pages in [kv_start/bs, ceil(kv_end/bs)) */
    int off_q_row0;    /* [nseq+1] prefix sum of n_new (first query row of each sequence) */
    int off_slot;      /* [ntok]   slot_mapping */
    int off_table;     /* [nseq][tstride] block-table rows */
    int total_ints;
    int uploaded;      /* device copy is current */
} pa_step_layout;

struct pa_handle {
    pa_config cfg;
    int C;
    BlockManager* mgr;
    int host_only;
    int compat;                   /* created by create_block_manager(): NH unknown until attention_paged */
    /* device pool: [n_layers][max_blocks][block_size][C] fp32, K and V */
    float* pool_k;
    float* pool_v;
    size_t layer_stride;          /* floats between layers */
    /* step tables: h_step points into a small ring of pinned buffers so the host can build
     * step n+1 while the copy of step n is still in flight (plain malloc when host_only) */
    int* h_step;
    int* d_step;
    size_t d_step_cap_ints;
    void* step_ring;              /* opaque, owned by pa_cuda.cu */
    pa_step_layout step;
    unsigned long step_uploads;   /* counts pa_step_upload calls: what was derived from the tables of one step is reused until it moves */
    int* step_seq_ids;            /* host [max_seqs] */
    int* step_n_new;              /* host [max_seqs] */
    int* step_pre_len;            /* host [max_seqs] context length of each sequence before the step being built */
    int* slot_scratch;            /* host [max_batch_tokens] */
    /* decode split workspace */
    float* d_ws;
    size_t ws_floats;
    int* d_counters;
    size_t n_counters;
    /* staging for the host-buffer entry points */
    float* h_stage;               /* pinned */
    float* d_stage;
    size_t stage_floats;
    void* stream;                 /* handle-owned stream */
    int sm_count;
    int smem_optin;
    int smem_per_sm;
    size_t max_pitch;             /* cudaDeviceProp::memPitch */
    int tune[32];
    long launches;
    void* d_dbg;                  /* optional per-CTA timeline of the last decode launch */
    int dbg_ctas;
    void* decode_attr_fn;         /* kernel whose dynamic-smem attribute has been raised */
    int max_heads;                /* heads the split workspace was sized for */
    void* tc_state;               /* TMA tensor maps of the pool (pa_prefill_tc.cu), lazily built */
    void* tc3_state;              /* the same for the fp32-accurate 3xTF32 prefill (pa_prefill_tc3.cu) */
    void* host_pipe;              /* copy streams + events of pa_decode_step_host_async (pa_kernels.cu) */
    int* compat_ints;             /* attention_paged scratch (page indices + per-row ints), grown on demand, kept */
    size_t compat_ints_cap;
    const int** compat_rows;
    size_t compat_rows_cap;
    int swap_enabled;             /* extension: evicted sequences are swapped out to host memory instead of dropped */
    void* swap_state;             /* pa_sharing.c */
    int swap_failures;            /* evictions whose host copy could not be made (the sequence was dropped, as the reference does) */
};

void pa_set_error(const char* fmt, ...);
/* launch switches of the CALLING THREAD (one host thread drives a handle): set from the handle at every
 * public entry that launches kernels, so two handles on two threads never see each other's settings */
extern __thread int pa_pdl_enabled;
extern __thread int pa_pdl_gate;

/* ---- implemented in pa_block_manager.c (plain C, integer only) ------------------------- */
BlockManager* pa_bm_create(pa_handle* owner, int channels, int block_size, int max_blocks,
                           int max_prompts, int table_stride);
void pa_bm_destroy(BlockManager* m);
/* page choice of add_to_cache (paged_infer.c:518-529); returns page index or -1 */
int pa_bm_choose_page(BlockManager* m, int prompt_id);
int pa_bm_context_len(const BlockManager* m, int prompt_id);
/* drop prompt p's hold on page idx: the page is freed when nobody else holds it (pa_sharing.c keeps holders valid) */
void pa_bm_release_page(BlockManager* m, int p, int idx);

/* ---- implemented in pa_sharing.c ------------------------------------------------ */
#define PA_OWNER_CACHE (-2)      /* KVBlock.prompt_id of a page held only by the prefix cache */
int pa_share_other_holder(BlockManager* m, int p, int idx);     /* another sequence holding idx, PA_OWNER_CACHE, or -1 */
int pa_share_evict_one_cached(BlockManager* m);                 /* frees the LRU cache-only page; 1 if one was freed */
void pa_share_destroy(BlockManager* m);
/* called by the allocator just before it evicts prompt p (block_manager.c:104-113): keeps a host copy when swapping is on */
void pa_swap_on_evict(pa_handle* h, int p);
void pa_swap_destroy(pa_handle* h);
int pa_swap_in_if_needed(pa_handle* h, int seq);

/* ---- implemented in pa_step.c ------------------------------------------------------------- */
int pa_create_compat(const pa_config* cfg, pa_handle** out);
/* step tables from explicit rows (compat attention_paged gets page pointers, not prompt ids):
 * row i has n_pages[i] page indices at tables[i], sees tokens [kv_start[i], kv_end[i]) from its
 * last query row and carries n_q[i] query rows. */
int pa_step_begin_raw(pa_handle* h, int nseq, const int* const* tables, const int* n_pages,
                      const int* kv_start, const int* kv_end, const int* n_q);

/* ---- implemented in pa_cuda.cu (the thin C-ABI layer over the CUDA runtime) ------------- */
int pa_cu_init(pa_handle* h);                 /* select device, query SMs/smem, create stream */
int pa_cu_alloc_pool(pa_handle* h);
void pa_cu_release(pa_handle* h);
int pa_cu_ensure_stage(pa_handle* h, size_t floats);
/* next pinned buffer of the ring with room for `ints` (waits for its previous upload) */
int* pa_cu_step_host_buffer(pa_handle* h, size_t ints);
int pa_cu_step_upload(pa_handle* h, void* stream);
int pa_cu_is_device_ptr(const void* p);
void pa_cu_host_pipe_release(pa_handle* h);
extern unsigned pa_host_free_generation;      /* pa_cuda.cu: counts pa_host_free calls */
/* rows [0, rows) of page src -> page dst, K and V, every layer; stream-ordered on the handle's stream, then synchronised */
int pa_cu_copy_page_rows(pa_handle* h, int src_page, int dst_page, int rows);
/* one page <-> host, K and V, every layer; host layout [layer][block_size*C] for K then the same for V */
int pa_cu_swap_page(pa_handle* h, int page, float* host_k, float* host_v, int to_host);     /* enqueued on the handle's stream */
int pa_cu_swap_sync(pa_handle* h);

/* ---- implemented in pa_prefill.cu / pa_prefill_tc.cu ------------------------------------- */
/* PA_OK = launched; PA_ERR_UNSUPPORTED = shape outside the kernel's domain (use the generic rows kernel) */
int pa_cu_prefill_tiled(pa_handle* h, int layer, const float* q, int q_stride, float* out,
                        int out_stride, int all_new_rows, void* stream);
int pa_cu_prefill_tc(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride,
                     void* stream);
void pa_cu_prefill_tc_release(pa_handle* h);
/* fp32-accurate tensor-core prefill (tcgen05 3xTF32, tolerance 1e-5) */
int pa_cu_prefill_tc3(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride, void* stream);
void pa_cu_prefill_tc3_release(pa_handle* h);

/* ---- implemented in pa_gemm_tc.cu: fp32-accurate (3xTF32) tensor-core GEMM ---------------- */
int pa_cu_gemm_tc(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                  int M, int N, int K, int n_dense, float* pool_k, float* pool_v, const int* slots, int C,
                  int terms, int n_split, const float* residual, int res_stride, int act, void* stream);
void pa_cu_gemm_stream_released(int device, void* stream);      /* frees the split-K workspace kept for that stream */
/* out = act(x.w^T + bias) + residual: tensor cores when the shape allows, else the fp32 SIMT kernel (pa_qkv.cu) */
int pa_cu_linear(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                 int M, int N, int K, const float* residual, int res_stride, int act, int path, void* stream);

/* ---- implemented in pa_layer_fused.cu: everything between two attention launches as ONE resident grid ------- */
struct pa_fused_params;
size_t pa_cu_layer_fused_smem(void);
int pa_cu_layer_fused_max_tiles(void);
int pa_cu_layer_fused_split(int N, int K, int sms, int* tiles_n);
int pa_cu_layer_fused_launch(const struct pa_fused_params* p, int sms, int cooperative, void* stream);
int pa_cu_make_map_2d(void* map_out, const float* ptr, int rows, int K, int row_stride, int box_rows);

/* ---- implemented in pa_model_mega.cu: the whole decode step of a handful of sequences as ONE
 * persistent cooperative kernel (every op of gpt2_forward for one new token per sequence) ------- */
#define PA_MEGA_MAX_SEQS 8
#define PA_MEGA_WARPS_PER_SM 16
#define PA_MEGA_AUTO_SEQS 6      /* chosen by itself up to this many sequences (measured against the chain of per-op kernels: 0.45 vs 0.75 ms at 2, 0.55 vs 0.85 at 4, 0.75 vs 0.78 at 6, 0.83 vs 0.79 at 8) */
typedef struct pa_mega_args {
    /* parameters, checkpoint order (paged_infer.c:441-488) */
    const float *wte, *wpe, *ln1w, *ln1b, *qkvw, *qkvb, *attprojw, *attprojb, *ln2w, *ln2b, *fcw, *fcb, *fcprojw,
        *fcprojb, *lnfw, *lnfb;
    int C, NH, hs, L, V, Vp;
    /* the step: M sequences, one new token each */
    int M;
    int tokens[PA_MEGA_MAX_SEQS], positions[PA_MEGA_MAX_SEQS];     /* by value: no copy to wait for */
    float coins[PA_MEGA_MAX_SEQS];
    int use_coins;                      /* 0: argmax */
    int* next;                          /* [M] sampled tokens: device-visible (mapped pinned host memory is fine) */
    float *x, *q, *atty, *fch, *logits; /* device activations: (M,C) (M,C) (M,C) (M,4C) (M,Vp) */
    /* paged KV cache and the mirrored step tables */
    float *pool_k, *pool_v;
    size_t layer_stride;
    const int *kv_end, *kv_start, *slots, *table;
    int tstride, bs, bs_shift;          /* block size (a power of two) and its log2 */
    float scale;
    /* attention split: tokens per chunk, chunks per sequence at most, partial (o[hs], m, l) workspace */
    int chunk_tokens, max_chunks;
    int local_attn;                     /* 1: the warps of one CTA share a (sequence, head) and merge in shared memory (max_chunks = warps per CTA) */
    float* part;
    unsigned* bar;                      /* grid barrier counter: monotonic, bar_base at launch */
    unsigned bar_base;
    int sm_count;
    unsigned long long* dbg;            /* optional timeline of CTA 0 (PA_MEGA_DEBUG=1), NULL normally */
} pa_mega_args;
/* bytes of dynamic shared memory the kernel needs for this geometry, or 0 when it is outside its domain */
size_t pa_cu_model_mega_smem(int M, int C, int hs, int block_size);
int pa_cu_model_mega_step(const pa_mega_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif
