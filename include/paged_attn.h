/*
 * paged_attn.h -- C ABI of libpaged_attn.so: a B200 (sm_100a) paged KV-cache attention path
 * that drops in behind the C interface of mx60s/llm.c-paged.
 *
 * Plain C (no CUDA headers needed by the including translation unit).  Two groups:
 *
 *   1. COMPAT: the reference's own names, argument order and error behaviour
 *      (block_manager.c:25-201, paged_infer.c:163-240 and :505-573).  The reference has
 *      no FFI/plugin layer -- integration is `#include "block_manager.c"`
 *      (paged_infer.c:10) -- so the boundary is these source-level signatures.
 *      A host translation unit replaces that include with `#include "paged_attn.h"`.
 *
 *   2. EXTENDED (pa_*): what the reference API cannot express -- run-time geometry,
 *      many sequences per step, many layers per manager, device-resident q/out, streams.
 *
 * Error convention: compat calls return NULL / -1 and print to stderr exactly where the
 * reference does (block_manager.c:116-119,138-141,148-151,166-169); pa_* calls return 0 or a
 * negative pa_status and never exit(); pa_last_error() holds the message.
 * There is NO CPU fallback: compute entry points fail loudly when no CUDA device is usable.
 */
#ifndef PAGED_ATTN_H
#define PAGED_ATTN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PA_API __attribute__((visibility("default")))
#else
#define PA_API
#endif

/* ------------------------------------------------------------------------------------------
 * Types.  Field names and meaning follow block_manager.c:9-23; arrays the reference sizes with
 * the MAX_* macros (block_manager.c:4-6) are run-time sized here.
 * ---------------------------------------------------------------------------------------- */
typedef struct pa_handle pa_handle;

typedef struct KVBlock {      /* block_manager.c:9-15 -- one page */
    float* keys;              /* DEVICE pointer: pool_k + index*block_size*C   ([slot][C] fp32, layer 0) */
    float* values;            /* DEVICE pointer: pool_v + index*block_size*C */
    int filled;               /* valid rows */
    int prompt_id;            /* owner, -1 = free */
    int lru_counter;
} KVBlock;

typedef struct BlockManager { /* block_manager.c:17-23 */
    int C;
    KVBlock* blocks;              /* [max_blocks]                      (ref: KVBlock blocks[MAX_BLOCKS]) */
    int** prompt_block_list;      /* [max_prompts] rows of block_table (ref: int [MAX_PROMPTS][MAX_BLOCKS]) */
    int* prompt_block_count;      /* [max_prompts] */
    int lru_epoch;
    /* --- extensions (not in the reference) --- */
    int block_size;               /* BLOCK_SIZE  (block_manager.c:6)  */
    int max_blocks;               /* MAX_BLOCKS  (block_manager.c:5)  */
    int max_prompts;              /* MAX_PROMPTS (block_manager.c:4)  */
    int table_stride;             /* ints per block-table row */
    int* block_table;             /* flat [max_prompts][table_stride]; this is what is mirrored to HBM */
    pa_handle* pa;                /* owning handle (pool, streams); never NULL */
    int* refcount;                /* [max_blocks] holders of a page: sequences (+1 if the prefix cache holds it); 0 = free.
                                     Always 0/1 unless pa_seq_fork / pa_prefix_* are used */
    void* prefix_cache;           /* opaque (pa_sharing.c), NULL until pa_prefix_insert */
    unsigned char* pinned;        /* [max_prompts] 1 while the sequence belongs to the step being built: the LRU never
                                     evicts it (all 0 outside pa_step_begin*, so the reference trace is unchanged) */
} BlockManager;

typedef enum pa_status {
    PA_OK = 0,
    PA_ERR_INVALID = -1,     /* bad argument / geometry */
    PA_ERR_NOMEM = -2,       /* host or device allocation failed */
    PA_ERR_CUDA = -3,        /* CUDA runtime error (message in pa_last_error) */
    PA_ERR_NO_DEVICE = -4,   /* compute call on a host-only handle, or no GPU present */
    PA_ERR_NO_BLOCKS = -5,   /* request_block failed ("No blocks available.") */
    PA_ERR_UNSUPPORTED = -6  /* shape outside every kernel's domain */
} pa_status;

/* ------------------------------------------------------------------------------------------
 * 1. COMPAT API -- same names / arguments as the reference.
 * ---------------------------------------------------------------------------------------- */

/* block_manager.c:38-52.  Geometry = reference macros (32/100/100) unless overridden by
 * pa_set_default_geometry() or the env vars PA_BLOCK_SIZE / PA_MAX_BLOCKS / PA_MAX_PROMPTS.
 * Allocates ONE device pool (K and V, 1 layer) instead of a malloc per page.  Unlike the
 * reference it zero-initialises lru_epoch / filled / lru_counter.  NULL + stderr on failure. */
PA_API BlockManager* create_block_manager(int channels);
/* The reference free()s the manager (block_manager_test.c:53); here that would leak the pool. */
PA_API void destroy_block_manager(BlockManager* manager);

PA_API void print_state(BlockManager* manager, int prompt);                         /* :25-36  */
PA_API int get_next_block_id(BlockManager* manager, int prompt, int block_id);      /* :54-63  */
PA_API KVBlock* get_current_block(BlockManager* manager, int prompt_id);            /* :65-76  (silent) */
PA_API void free_blocks_for_prompt(BlockManager* manager, int prompt_id);           /* :78-90  */
PA_API int find_least_recently_used_block(BlockManager* manager);                   /* :92-102 */
PA_API void page_out_lru_block(BlockManager* manager);                              /* :104-113 */
PA_API KVBlock* request_block(BlockManager* manager, int prompt_id);                /* :115-162 */
/* :165-201.  Returns malloc'd kv[0][i]=keys, kv[1][i]=values (device pointers) of the prompt's
 * pages in table order; caller frees kv[0], kv[1], kv.  NULL if none / bad id. */
PA_API float*** collect_kv_blocks(BlockManager* manager, int prompt_id, int* num_blocks);

/* paged_infer.c:505-573.  qkv is (B,T,3C) fp32, HOST or DEVICE memory; appends K,V of the last
 * n_tail rows of batch row 0 to prompt 0's current page (new page if none/full, else LRU touch),
 * through the KV-append kernel.  Like the reference it does not cross a page boundary. */
PA_API void add_to_cache(BlockManager* manager, float* qkv, int B, int T, int C, int n_tail);

/* paged_infer.c:163-240.  out (B,T,C) and inp (B,T,3C) are HOST or DEVICE memory; key_blocks /
 * value_blocks are the arrays collect_kv_blocks returned (host arrays of device page pointers).
 * Row t attends cached tokens [offset, offset+t].  preatt/att ((B,NH,T,T) scratch, needed only
 * by a backward pass the reference does not have) are NOT materialised and may be NULL. */
PA_API void attention_paged(float* out, float* preatt, float* att, float* inp,
                            float** key_blocks, float** value_blocks,
                            int B, int T, int C, int NH, int offset);

/* The step before the path (paged_infer.c:92-160, call sites :703-706): fp32 GEMM with bias,
 * out (B,T,OC) = inp (B,T,C) . weight (OC,C)^T + bias; matmul_cached computes Q for every window
 * row and K, V only for the last row of each batch entry.  HOST or DEVICE pointers. */
PA_API void matmul_forward(float* out, float* inp, float* weight, float* bias, int B, int T, int C, int OC);
PA_API void matmul_cached(float* out, float* inp, float* weight, float* bias, int B, int T, int C, int OC);

/* geometry used by the next create_block_manager() (<=0 keeps the current value) */
PA_API void pa_set_default_geometry(int block_size, int max_blocks, int max_prompts);
PA_API int pa_default_block_size(void);
#ifdef PA_COMPAT_MACROS          /* for callers that loop over the reference's macros */
#define BLOCK_SIZE (pa_default_block_size())
#endif

/* ------------------------------------------------------------------------------------------
 * 2. EXTENDED API
 * ---------------------------------------------------------------------------------------- */
typedef struct pa_config {
    int block_size;          /* tokens per page                                  (BLOCK_SIZE)  */
    int max_blocks;          /* pages in the pool, shared by all sequences        (MAX_BLOCKS)  */
    int max_seqs;            /* sequence ids 0..max_seqs-1                        (MAX_PROMPTS) */
    int max_blocks_per_seq;  /* block-table row stride; 0 -> max_blocks (as the reference) */
    int n_layers;            /* KV pools behind one block table (reference: 1) */
    int n_heads;             /* NH */
    int head_dim;            /* hs; C = n_heads*head_dim */
    int device;              /* CUDA ordinal; PA_HOST_ONLY = integer tables only (no pool, no compute) */
    int max_batch_tokens;    /* most new tokens in one step over all sequences; 0 -> max_seqs */
} pa_config;
#define PA_HOST_ONLY (-1)

PA_API int pa_create(const pa_config* cfg, pa_handle** out);
PA_API void pa_destroy(pa_handle* h);
PA_API BlockManager* pa_manager(pa_handle* h);
PA_API const char* pa_last_error(void);
PA_API const char* pa_version(void);

/* ---- per-step host scheduling (integer; no device work) ---------------------------------- */
/* Sequence seq_ids[i] receives n_new[i] tokens.  For each token run the page choice of
 * add_to_cache (paged_infer.c:518-529: current page, new page if none/full, else LRU touch),
 * crossing page boundaries when needed (extension), and record slot = page*block_size + row.
 * Builds the step tables (context lengths, page prefix sums, slot mapping, block-table rows).
 * The sequences of the step are pinned while its pages are placed: the whole-prompt LRU eviction may
 * take any OTHER sequence, never one of the step (nor the requester itself, which the single-prompt
 * reference would evict, block_manager.c:157): when only they are left the call fails with
 * PA_ERR_NO_BLOCKS and every sequence of the step is back at the length it had before the call.
 * A sequence id may appear once per step. */
PA_API int pa_step_begin(pa_handle* h, const int* seq_ids, const int* n_new, int nseq);
/* Optional sliding window: row i attends cached tokens [kv_start[i], ctx) (reference `offset`). */
PA_API int pa_step_set_kv_start(pa_handle* h, const int* kv_start);
/* The same tables without appending anything (attention over what is cached). */
PA_API int pa_step_begin_readonly(pa_handle* h, const int* seq_ids, int nseq);
PA_API const int* pa_step_slot_mapping(pa_handle* h, int* n_tokens);
PA_API const int* pa_step_context_lens(pa_handle* h, int* nseq);
PA_API const int* pa_step_block_table(pa_handle* h, int* nseq, int* stride);
/* Mirror the step tables to HBM: ONE cudaMemcpyAsync from pinned memory. */
PA_API int pa_step_upload(pa_handle* h, void* stream);
/* Bounds check of the step tables (every page index inside the pool, every slot inside it, windows and prefix sums
 * consistent): all addresses the kernels form into the KV pool derive from these.  pa_step_upload runs it on every
 * step when the environment has PA_VALIDATE_STEP=1. */
PA_API int pa_step_validate(pa_handle* h);

/* ---- kernels (device pointers, caller's stream; NULL stream = the handle's own) ----------- */
/* KV append: token j of the step (order of pa_step_begin) has K at k+j*row_stride, V likewise. */
PA_API int pa_append(pa_handle* h, int layer, const float* k, const float* v, int row_stride, void* stream);
/* Decode: one query per sequence (row i at q+i*q_stride, head hd at +hd*head_dim) attends
 * cached tokens [kv_start, ctx) of its sequence; out row i at out+i*out_stride. */
PA_API int pa_decode(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride, void* stream);
/* Decode with the KV append fused in (one launch per layer): the step must hold exactly one new
 * token per sequence; row i of q, k, v (same row_stride, e.g. the (B,3C) qkv buffer) is that
 * token.  The new K/V row is read straight from k/v, used, and stored to its page slot. */
PA_API int pa_decode_append(pa_handle* h, int layer, const float* q, const float* k, const float* v, int row_stride,
                            float* out, int out_stride, void* stream);
/* Causal rows: the n_new[i] new tokens of each sequence are the queries (packed in step order);
 * row j of sequence i attends [kv_start, ctx_before + j].  fp32 accuracy (1e-5) on every automatic path: tensor
 * cores with the 3xTF32 split wherever the kernel's domain allows (head_dim 64 / 128, pages of 8..64 resp. 8..32
 * tokens), fp32 SIMT otherwise (PA_TUNE_PREFILL_PATH). */
PA_API int pa_prefill(pa_handle* h, int layer, const float* q, int q_stride, float* out, int out_stride, void* stream);
/* The work list of the persistent tensor-core prefill kernel for the step being built (host only, no GPU needed;
 * what pa_prefill uploads once per step): one CTA per column, `n_ctas` columns (at most `max_ctas`, 0 = the device's
 * SM count, 148 without a device), `rows` units per column; units[row * n_ctas + cta] = { sequence index in the step
 * (-1: none), q tile | head << 16 }; a q tile is 128 query rows, key_tile = 64 keys at head_dim 64, 32 at 128.
 * `units` may be NULL to ask for the sizes only; cap = its capacity in entries.  Exposed for tests and tools: every
 * (sequence, q tile, head) must appear exactly once, and the columns' work (key tiles) must balance. */
PA_API int pa_prefill_schedule(pa_handle* h, int key_tile, int max_ctas, int* units, size_t cap, int* n_ctas, int* rows);

/* QKV projection of the step's new tokens with the KV append fused into its epilogue
 * (matmul_cached + add_to_cache in one kernel): x (ntok, C) rows in step order, w (3C, C), bias (3C)
 * or NULL.  Q goes to q_out (ntok, C); K and V go straight to each token's page slot. */
PA_API int pa_qkv_append(pa_handle* h, int layer, const float* x, int x_stride, const float* w, const float* bias,
                         float* q_out, int q_stride, void* stream);
/* plain fp32 GEMM with bias on device pointers: out (M,N) = x (M,K) . w (N,K)^T + bias */
PA_API int pa_matmul_bias(const float* x, int x_stride, const float* w, const float* bias, float* out, int out_stride,
                          int M, int N, int K, void* stream);

/* ---- the whole decode step of the model around the path ----------------------- */
/* gpt2_forward (paged_infer.c:646-728) for one new token per sequence, over ALL layers (the
 * reference fork stops at layer 0), on a handle created with n_layers KV pools.  Parameters: the
 * checkpoint's 16 tensors in file order (paged_infer.c:441-488) in one buffer. */
typedef struct pa_model pa_model;
typedef struct pa_model_config {
    int max_seq_len;   /* maxT: rows of wpe */
    int vocab_size;    /* V */
    int n_layers;      /* L  (== handle n_layers) */
    int n_heads;       /* NH (== handle n_heads) */
    int channels;      /* C  (== n_heads * head_dim of the handle) */
} pa_model_config;
PA_API size_t pa_model_param_count(const pa_model_config* cfg);
/* params_host: pa_model_param_count floats, or NULL for synthetic random-init weights (seeded).
 * max_batch: the most tokens (and sequences) one step may carry. */
PA_API int pa_model_create(pa_handle* h, const pa_model_config* cfg, const float* params_host, unsigned long long seed,
                           int max_batch, pa_model** out);
PA_API void pa_model_destroy(pa_model* m);
/* Sequence seq_ids[i] receives tokens[i] at its next position; next_tokens[i] is sampled from its
 * logits with coins[i] in [0,1) as sample_mult does (paged_infer.c:838-848), or argmax if coins is NULL.
 * Up to 6 sequences the whole step runs as ONE persistent cooperative kernel (weights streamed once,
 * grid barriers between the ops), beyond that as a chain of per-op kernels (PA_TUNE_MODEL_PATH). */
PA_API int pa_model_decode_step(pa_model* m, const int* seq_ids, const int* tokens, const float* coins, int nseq,
                                int* next_tokens);
/* General step (prompt prefill, chunked prefill, decode, or a mix): sequence seq_ids[i] receives n_new[i] >= 1
 * tokens, packed in step order in `tokens`; next_tokens[i] comes from the logits of its last new position.
 * max_batch of pa_model_create bounds the tokens of one step. */
PA_API int pa_model_forward(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins,
                            int nseq, int* next_tokens);
/* The same step in two halves: _async enqueues everything on the handle's stream and returns without a host
 * synchronisation; pa_model_wait synchronises and hands out the sampled tokens (next_tokens may be NULL).
 * A host driving several GPUs queues every GPU's step before it waits for any (pa_group_model_step). */
PA_API int pa_model_forward_async(pa_model* m, const int* seq_ids, const int* n_new, const int* tokens, const float* coins,
                                  int nseq);
PA_API int pa_model_wait(pa_model* m, int* next_tokens);
PA_API int* pa_model_next_tokens_dev(pa_model* m);        /* device: sampled tokens of the step in flight / last step, [nseq] */
PA_API void pa_model_want_device_tokens(pa_model* m, int on); /* keep them in device memory too (the persistent kernel otherwise writes mapped host memory only) */
PA_API pa_handle* pa_model_handle(pa_model* m);
PA_API float* pa_model_params(pa_model* m);               /* device */
PA_API float* pa_model_logits(pa_model* m, int* stride);  /* device, (nseq, stride) of the last step */

/* ---- multi-GPU: sequences sharded over the GPUs of one box, each GPU with its own block manager,
 * page pool and tables; the attention needs no collective.  The ONE exchange is the all-gather of the
 * sampled tokens (or of last-position logits) after the sampler: plain ncclAllGather enqueued on each
 * handle's stream behind the sampler kernel, no host synchronisation in between.  NCCL (libnccl.so.2) is
 * opened on first use, not linked.  Only precedent in the reference: Python DDP, train_gpt2.py:400-412. */
typedef struct pa_group pa_group;
#define PA_COMM_ID_BYTES 128
/* (a) ONE process drives n GPUs: a handle + stream per device (devices[i], or 0..n-1 when NULL; cfg->device is
 * ignored) and one communicator per handle (ncclCommInitAll).  pa_group_destroy destroys the handles too. */
PA_API int pa_group_create(const pa_config* cfg, int n_gpus, const int* devices, pa_group** out);
/* (b) one process per GPU (torchrun, MPI): rank 0 makes the id, the launcher distributes its 128 bytes, every
 * rank joins with its own handle (ncclCommInitRank).  The handle stays the caller's. */
PA_API int pa_comm_unique_id(void* id128);
PA_API int pa_group_join(pa_handle* h, const void* id128, int rank, int world, pa_group** out);
PA_API void pa_group_destroy(pa_group* g);
PA_API int pa_group_size(pa_group* g);                    /* ranks in the group */
PA_API int pa_group_local_count(pa_group* g);             /* members driven by this process: n (a) or 1 (b) */
PA_API int pa_group_rank(pa_group* g, int i);             /* rank of local member i */
PA_API pa_handle* pa_group_handle(pa_group* g, int i);    /* handle of local member i */
/* all-gather on the members' streams: send[i] = n_per_rank int32 (device memory of member i), recv[i] =
 * size * n_per_rank int32 there, rank-major.  Stream-ordered; returns without synchronising. */
PA_API int pa_group_gather_tokens(pa_group* g, const int* const* send, int* const* recv, int n_per_rank);
PA_API int pa_group_gather_logits(pa_group* g, const float* const* send, float* const* recv, size_t n_floats_per_rank);
/* One decode step of the whole group: model i (on member i's handle) takes one token per sequence, the sampled
 * tokens of all ranks are gathered behind the samplers, the host waits once per member.
 * all_next: size * nseq ints, rank-major. */
PA_API int pa_group_model_step(pa_group* g, pa_model* const* models, const int* const* seq_ids, const int* const* tokens,
                               const float* const* coins, int nseq, int* all_next);
/* The same step with the gather OFF the critical path: a rank's next step consumes only its OWN sampled tokens
 * (next_local[i]: nseq ints of local member i, handed back after ONE wait on its stream); the all-gather runs on
 * a side stream behind the sampler, beside the next step's kernels, and its result comes out of the NEXT call
 * (gathered_prev: size * nseq ints of the previous step, rank-major; may be NULL) or of pa_group_gather_flush.
 * Returns how many sequences per rank gathered_prev holds (0 on the first call), or a negative pa_status. */
PA_API int pa_group_model_step_overlapped(pa_group* g, pa_model* const* models, const int* const* seq_ids,
                                          const int* const* tokens, const float* const* coins, int nseq,
                                          int* const* next_local, int* gathered_prev);
PA_API int pa_group_gather_flush(pa_group* g, int* gathered);   /* waits for the gather in flight; returns its nseq (0: none) */
PA_API int pa_nccl_version(void);                         /* e.g. 22809; 0 when NCCL cannot be opened */

/* ---- the reference's on-disk formats (host only, no device needed) ---------------- */
/* checkpoint gpt2_124M.bin (reader paged_infer.c:436-502): 256 x int32 header + 16 fp32 tensors */
PA_API int pa_checkpoint_read_config(const char* path, pa_model_config* cfg);
PA_API int pa_checkpoint_read_params(const char* path, float* params, size_t n_floats);
PA_API int pa_checkpoint_write(const char* path, const pa_model_config* cfg, const float* params);
/* version 2 of the same file (train_gpt2.py:266-320): weights and biases as bf16 (round to nearest even), the
 * layernorm tensors in fp32 at the end.  pa_checkpoint_read_config / _read_params read both versions. */
PA_API int pa_checkpoint_write_bf16(const char* path, const pa_model_config* cfg, const float* params);
PA_API int pa_model_create_from_checkpoint(pa_handle* h, const char* path, int max_batch, pa_model** out);
/* token stream (dataloader_*, paged_infer.c:769-818): raw int32 ids, batches of B*T (+1 target) */
typedef struct pa_dataloader pa_dataloader;
PA_API int pa_dataloader_open(const char* path, int B, int T, pa_dataloader** out);
PA_API void pa_dataloader_reset(pa_dataloader* d);
PA_API int pa_dataloader_num_batches(const pa_dataloader* d);
PA_API int pa_dataloader_next_batch(pa_dataloader* d, const int** inputs, const int** targets);
PA_API void pa_dataloader_close(pa_dataloader* d);
PA_API int pa_tokens_write(const char* path, const int* ids, size_t n);
/* tokenizer gpt2_tokenizer.bin (tokenizer_init / tokenizer_decode, paged_infer.c:875-915) */
typedef struct pa_tokenizer pa_tokenizer;
PA_API int pa_tokenizer_open(const char* path, pa_tokenizer** out);
PA_API unsigned pa_tokenizer_vocab_size(const pa_tokenizer* t);
PA_API const char* pa_tokenizer_decode(const pa_tokenizer* t, unsigned token_id);
PA_API void pa_tokenizer_close(pa_tokenizer* t);
PA_API int pa_tokenizer_write(const char* path, const char* const* pieces, const unsigned char* lens, unsigned vocab_size);

/* ---- whole step with HOST buffers (the end-to-end entry: H2D, append, decode, D2H, sync) -- */
/* qkv_host: (nseq, 3C) rows of the step's sequences [q | k | v]; out_host: (nseq, C).  Pinned
 * buffers (pa_host_alloc) are read and written by the kernel directly over PCIe (zero-copy);
 * pageable buffers are staged through pinned memory with cudaMemcpyAsync. */
PA_API int pa_decode_step_host(pa_handle* h, int layer, const float* qkv_host, float* out_host);
/* The same without the final synchronisation (PINNED buffers only), so a host can queue the layers of a
 * step, or several steps, and call pa_decode_step_host_sync once before it reads out_host or reuses
 * qkv_host.  Zero-copy (default) or, with PA_TUNE_NO_ZEROCOPY, staged copies on their own streams that
 * overlap the neighbouring layers' kernels. */
PA_API int pa_decode_step_host_async(pa_handle* h, int layer, const float* qkv_host, float* out_host);
PA_API int pa_decode_step_host_sync(pa_handle* h);
/* every layer of the step in one call: layer l reads qkv_host + l * qkv_layer_stride floats (0: the same rows for
 * every layer) and writes out_host + l * out_layer_stride.  With pinned buffers the step is staged as a whole: ONE
 * host-to-device copy of its inputs and ONE device-to-host copy of its outputs on copy streams of their own, double
 * buffered, so the copies of neighbouring steps run beside the kernels and the kernels never wait on PCIe. */
PA_API int pa_decode_step_host_layers_async(pa_handle* h, const float* qkv_host, size_t qkv_layer_stride, float* out_host,
                                            size_t out_layer_stride);
/* Completion tickets: pa_decode_step_host_mark returns a ticket (> 0) for everything queued so far on this handle;
 * pa_decode_step_host_wait blocks until that work -- kernels and output copies -- is done, and only that.  The host
 * queues step n+1 (pa_step_begin builds its tables in the next pinned buffer of the ring; the layers follow on the
 * stream) BEFORE it waits for step n's outputs, so the device never idles while the host turns a step around.
 * The attention path has no dependency of a step on its predecessor's outputs: its inputs are the host's q|k|v rows.
 * A ticket stays valid until 8 newer ones were made. */
PA_API int pa_decode_step_host_mark(pa_handle* h);
PA_API int pa_decode_step_host_wait(pa_handle* h, int ticket);

/* ---- sequence bookkeeping ----------------------------------------------------------------- */
PA_API int pa_seq_len(pa_handle* h, int seq_id);                 /* cached tokens */
PA_API int pa_seq_truncate(pa_handle* h, int seq_id, int new_len); /* roll back (frees emptied pages) */
PA_API int pa_seq_free(pa_handle* h, int seq_id);                /* = free_blocks_for_prompt */
/* Undo the appends of the last pa_step_begin (speculative-decoding style roll back): every
 * sequence of the step loses the n_new tokens it received; emptied pages return to the pool. */
PA_API int pa_step_rollback(pa_handle* h);
/* Install an externally built block table (e.g. a shuffled / fragmented layout for benchmarks):
 * the sequence must be empty and the pages free. */
PA_API int pa_seq_adopt(pa_handle* h, int seq_id, const int* blocks, int n_blocks, int n_tokens);

/* ---- allocator extensions (all OFF unless called: the reference trace is unchanged) ---- */
/* Parallel sampling (the reference's unused `P`, paged_infer.c:958): dst (empty) becomes a copy of
 * src.  Full pages are SHARED (reference-counted, read-only), the partial last page is copied on
 * the device (all layers), so both sequences can append independently afterwards. */
PA_API int pa_seq_fork(pa_handle* h, int src_seq, int dst_seq);
/* Prefix sharing by hashing ("if you do hashing you need to change up this policy",
 * block_manager.c:109-111): register the FULL pages of a prefilled sequence under the chained hash
 * of their token ids; the cache keeps them alive after the sequence is freed (evicted LRU-first
 * before any live sequence is). */
PA_API int pa_prefix_insert(pa_handle* h, int seq_id, const int* tokens, int n_tokens);
/* seq_id must be empty: adopts the longest cached page-aligned prefix of tokens (never the whole
 * prompt: at least one token is left to compute) and returns the number of tokens matched. */
PA_API int pa_prefix_match(pa_handle* h, int seq_id, const int* tokens, int n_tokens);
PA_API int pa_prefix_cached_pages(pa_handle* h);
/* Swap-out instead of drop-on-evict: with swapping on, a sequence evicted by the allocator (LRU, whole
 * prompt, as block_manager.c:104-113) keeps a host copy of its pages (all layers) and is brought back
 * -- into whatever pages are free then -- the next time pa_step_begin* names it. */
PA_API int pa_set_evict_swap(pa_handle* h, int enable);
PA_API int pa_seq_swap_out(pa_handle* h, int seq_id);       /* explicit */
PA_API int pa_seq_swap_in(pa_handle* h, int seq_id);
PA_API int pa_swap_failures(pa_handle* h);                  /* evictions whose host copy could not be made (sequence dropped; also on stderr) */
PA_API int pa_seq_swapped_tokens(pa_handle* h, int seq_id);  /* tokens held in the host copy, 0 if resident or unknown */
PA_API int pa_page_refcount(pa_handle* h, int page);

/* ---- pool access (tests, benchmarks, checkpointing) --------------------------------------- */
PA_API float* pa_pool_k(pa_handle* h, int layer);                /* device, [max_blocks][bs][C] */
PA_API float* pa_pool_v(pa_handle* h, int layer);
PA_API size_t pa_pool_bytes(pa_handle* h);                       /* K+V, all layers */
PA_API int pa_device(pa_handle* h);
PA_API void* pa_stream_of(pa_handle* h);                         /* the handle-owned stream (host-buffer entries run on it) */
PA_API int pa_sm_count(pa_handle* h);

/* ---- tuning knobs (benchmarks/tests select a kernel or a tile shape) ----------------------- */
typedef enum pa_tune_key {
    PA_TUNE_DECODE_PATH = 0,   /* 0 auto, 1 stream (TMA/mbarrier) kernel, 2 generic SIMT kernel, 3 small-batch kernel (one CTA per sequence and head; auto below a measured amount of KV) */
    PA_TUNE_HEADS_PER_TILE = 1,/* 0 auto */
    PA_TUNE_STAGES = 2,        /* 0 auto */
    PA_TUNE_GRID = 3,          /* 0 auto (CTAs) */
    PA_TUNE_COUNT_LAUNCHES = 4,/* read-only counter of kernels launched by this handle */
    PA_TUNE_STATIC_PCT = 5,    /* 0 auto: share of the page stream split statically (rest is claimed dynamically) */
    PA_TUNE_DYN_UNITS = 6,     /* 0 auto: pages per dynamically claimed range */
    PA_TUNE_DEBUG_TIMELINE = 7,/* 1: the stream decode kernel records a per-CTA timeline (pa_debug_timeline) */
    PA_TUNE_NO_PDL = 8,        /* 1: launch the decode kernel without programmatic dependent launch */
    PA_TUNE_NO_ZEROCOPY = 9,   /* host-buffer entries, pinned buffers: 0 auto (one layer per call: zero-copy, the kernel reads / writes the mapped host rows itself; pa_decode_step_host_layers_async: whole-step staging, one input and one output copy per step on copy streams beside the kernels), 1 per-layer staged copies, 2 zero-copy always, 3 whole-step staging or fail */
    PA_TUNE_LAST_HPG = 10,     /* read-only: heads per tile, ring stages and CTAs of the last stream-decode launch */
    PA_TUNE_LAST_STAGES = 11,  /*            (0 when the last decode ran on the generic kernel) */
    PA_TUNE_LAST_GRID = 12,
    PA_TUNE_PREFILL_PATH = 13, /* 0 auto (tcgen05 3xTF32 wherever its domain allows -- head_dim 64/128, pages of 8..64 resp. 8..32 tokens -- else tiled fp32 SIMT: both within 1e-5), 1 tiled fp32 SIMT, 2 generic rows kernel, 3 tcgen05 plain TF32 (opt-in, own tolerance 5e-3), 4 tcgen05 3xTF32 (fp32-accurate; fails outside its domain) */
    PA_TUNE_TC_WARPGROUPS = 14,/* tcgen05 prefill: softmax warpgroups per CTA, 0 auto, 1 or 2 */
    PA_TUNE_TC_KEY_TILE = 15,  /* tcgen05 prefill, head_dim 64: keys per tile, 0 auto (64), 64 or 128 */
    PA_TUNE_GEMM_PATH = 16,    /* projections: 0 auto (<= 4 rows: weight-streaming GEMV, else tcgen05 3xTF32, fp32-accurate), 1 fp32 SIMT, 2 tcgen05 3xTF32, 3 tcgen05 plain TF32 (reduced precision, own tolerance), 4 GEMV */
    PA_TUNE_GEMM_SPLIT_K = 17, /* tensor-core projections: CTAs per output tile splitting K. 0 auto; n > 0: n (<= 16, clamped to what is co-resident), partial tiles reduced through an L2 workspace; -2 / -4: a 2- / 4-CTA cluster reducing through distributed shared memory */
    PA_TUNE_MODEL_PATH = 18,   /* pa_model_forward: 0 auto (<= 6 sequences of one new token: ONE persistent kernel for the whole step, else the chain of per-op kernels), 1 chain, 2 persistent kernel, 3 one resident grid per layer between the attention launches (steps of <= 128 tokens; opt-in: measured slower than the chain); 2 and 3 fail when the step is outside their domain */
    PA_TUNE_MAX
} pa_tune_key;
PA_API int pa_tune_set(pa_handle* h, int key, int value);
PA_API int pa_tune_get(pa_handle* h, int key);
/* Per-CTA timeline of the last stream-decode launch, 8 words per CTA: [0] entry ns, [1] first
 * TMA issue ns, [2] first tile landed ns, [3] last tile consumed ns (globaltimer), [4] producer
 * cycles waiting for a free slot, [5] consumer cycles waiting for data, [6] consumer cycles in
 * segment ends (partials, merges), [7] tiles.  Returns the number of CTAs written. */
PA_API int pa_debug_timeline(pa_handle* h, unsigned long long* out, int max_ctas);

/* ---- thin CUDA plumbing for plain-C hosts (no cuda_runtime.h needed) ----------------------- */
PA_API int pa_device_count(void);
PA_API int pa_set_device(int device);                            /* current device of the calls below (handles switch by themselves) */
PA_API void* pa_dev_alloc(size_t bytes);
PA_API void pa_dev_free(void* p);
PA_API void* pa_host_alloc(size_t bytes);                        /* pinned */
PA_API void pa_host_free(void* p);
PA_API int pa_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
PA_API int pa_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
PA_API int pa_memset(void* dst, int value, size_t bytes, void* stream);
/* N(mean, stdv) written by the device from a counter-based hash of (seed, index): synthetic pools of any size
 * without a host copy; element i is a pure function of (seed, i) */
PA_API int pa_fill_normal(float* dev, size_t n, float stdv, float mean, unsigned long long seed, void* stream);
PA_API void* pa_stream_create(void);
PA_API void pa_stream_destroy(void* stream);
PA_API int pa_stream_sync(void* stream);
PA_API int pa_device_sync(void);
PA_API void* pa_event_create(void);
PA_API void pa_event_destroy(void* ev);
PA_API int pa_event_record(void* ev, void* stream);
PA_API float pa_event_elapsed_ms(void* start, void* stop);       /* syncs on stop */
/* write `bytes` of device scratch (> L2) to evict the L2 between timed iterations */
PA_API int pa_flush_l2(void* scratch, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAGED_ATTN_H */
