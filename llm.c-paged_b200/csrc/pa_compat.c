/*
 * pa_compat.c -- host side, plain C: the reference-named entry points that move data
 * (create_block_manager, add_to_cache, attention_paged) on top of the pa_* layer, so that the
 * call site paged_infer.c:710-715 compiles against paged_attn.h unchanged.
 *
 * The integer decisions (which page, LRU stamps, `filled`) are taken on the host exactly as
 * paged_infer.c:518-529,570 takes them; the bytes move through the CUDA kernels.  Buffers may be
 * host memory (as in the reference, staged through pinned memory inside the call) or device
 * memory.  Calls are synchronous like the reference's.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_internal.h"

/* ---- default geometry: the reference's macros (block_manager.c:4-6) ------------------------ */
static int g_block_size = 32, g_max_blocks = 100, g_max_prompts = 100;

static int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    if (!v || !*v) return fallback;
    int x = atoi(v);
    return x > 0 ? x : fallback;
}
void pa_set_default_geometry(int block_size, int max_blocks, int max_prompts) {
    if (block_size > 0) g_block_size = block_size;
    if (max_blocks > 0) g_max_blocks = max_blocks;
    if (max_prompts > 0) g_max_prompts = max_prompts;
}
int pa_default_block_size(void) { return env_int("PA_BLOCK_SIZE", g_block_size); }

/* ---- registry: attention_paged receives page pointers, not a manager -----------------------
 * A growable array behind a mutex: managers may be created and destroyed from different host threads
 * (each manager itself is driven by one thread, like the reference's). */
static pthread_mutex_t g_live_lock = PTHREAD_MUTEX_INITIALIZER;
static pa_handle** g_live = NULL;
static int g_live_n = 0, g_live_cap = 0;

static int registry_add(pa_handle* h) {
    int ok = 1;
    pthread_mutex_lock(&g_live_lock);
    if (g_live_n == g_live_cap) {
        int cap = g_live_cap ? 2 * g_live_cap : 16;
        pa_handle** p = (pa_handle**)realloc(g_live, (size_t)cap * sizeof(*p));
        if (p) { g_live = p; g_live_cap = cap; } else ok = 0;
    }
    if (ok) g_live[g_live_n++] = h;
    pthread_mutex_unlock(&g_live_lock);
    return ok;
}
static void registry_remove(pa_handle* h) {
    pthread_mutex_lock(&g_live_lock);
    for (int i = 0; i < g_live_n; i++)
        if (g_live[i] == h) { g_live[i] = g_live[--g_live_n]; break; }
    pthread_mutex_unlock(&g_live_lock);
}
static pa_handle* registry_find_by_page(const float* key_page) {
    pa_handle* found = NULL;
    pthread_mutex_lock(&g_live_lock);
    for (int i = 0; i < g_live_n && !found; i++) {
        pa_handle* h = g_live[i];
        if (h->pool_k && key_page >= h->pool_k && key_page < h->pool_k + h->layer_stride * h->cfg.n_layers) found = h;
    }
    pthread_mutex_unlock(&g_live_lock);
    return found;
}

BlockManager* create_block_manager(int channels) {
    pa_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.block_size = env_int("PA_BLOCK_SIZE", g_block_size);
    cfg.max_blocks = env_int("PA_MAX_BLOCKS", g_max_blocks);
    cfg.max_seqs = env_int("PA_MAX_PROMPTS", g_max_prompts);
    cfg.n_layers = 1;                 /* one manager = one layer's KV (paged_infer.c:433) */
    cfg.n_heads = 1;                  /* NH arrives with attention_paged */
    cfg.head_dim = channels;
    cfg.device = getenv("PA_DEVICE") ? atoi(getenv("PA_DEVICE")) : 0;
    cfg.max_batch_tokens = cfg.block_size * 4 > 1024 ? cfg.block_size * 4 : 1024;
    if (channels < 1) { fprintf(stderr, "create_block_manager: invalid channels %d\n", channels); return NULL; }
    /* compat handles size the split workspace for any NH dividing C */
    pa_handle* h = NULL;
    if (pa_create_compat(&cfg, &h) != PA_OK) {
        fprintf(stderr, "create_block_manager: %s\n", pa_last_error());
        return NULL;
    }
    if (!registry_add(h)) {
        fprintf(stderr, "create_block_manager: out of host memory\n");
        pa_destroy(h);
        return NULL;
    }
    return h->mgr;
}

void destroy_block_manager(BlockManager* manager) {
    if (!manager) return;
    registry_remove(manager->pa);
    pa_destroy(manager->pa);
}

/* paged_infer.c:505-573 */
void add_to_cache(BlockManager* manager, float* qkv, int B, int T, int C, int n_tail) {
    if (!manager || !qkv || n_tail < 0 || n_tail > T || C != manager->C || B < 1) {
        fprintf(stderr, "add_to_cache: invalid arguments\n");
        return;
    }
    pa_handle* h = manager->pa;
    if (n_tail == 0) {              /* the reference still performs the page choice */
        pa_bm_choose_page(manager, 0);
        return;
    }
    int seq = 0;                    /* `int b = 0; // placeholder` (paged_infer.c:515) */
    if (pa_step_begin(h, &seq, &n_tail, 1) != PA_OK) {
        fprintf(stderr, "add_to_cache: %s\n", pa_last_error());
        return;
    }
    /* The reference's copy loop runs over b but writes every b to the same rows (:548-566), so
     * the last batch row wins. */
    const float* rows = qkv + ((size_t)(B - 1) * T + (T - n_tail)) * 3 * (size_t)C;
    const float* dev_rows = rows;
    size_t n = (size_t)n_tail * 3 * C;
    if (!pa_cu_is_device_ptr(qkv)) {
        if (pa_cu_ensure_stage(h, n) != PA_OK) { fprintf(stderr, "add_to_cache: %s\n", pa_last_error()); return; }
        memcpy(h->h_stage, rows, n * sizeof(float));
        if (pa_memcpy_h2d(h->d_stage, h->h_stage, n * sizeof(float), h->stream) != PA_OK) {
            fprintf(stderr, "add_to_cache: %s\n", pa_last_error());
            return;
        }
        dev_rows = h->d_stage;
    }
    if (pa_step_upload(h, h->stream) != PA_OK ||
        pa_append(h, 0, dev_rows + C, dev_rows + 2 * C, 3 * C, h->stream) != PA_OK ||
        pa_stream_sync(h->stream) != PA_OK)
        fprintf(stderr, "add_to_cache: %s\n", pa_last_error());
}

/* paged_infer.c:163-240 */
void attention_paged(float* out, float* preatt, float* att, float* inp,
                     float** key_blocks, float** value_blocks,
                     int B, int T, int C, int NH, int offset) {
    (void)preatt; (void)att;        /* backward-only scratch: not materialised */
    if (!out || !inp || !key_blocks || !value_blocks || B < 1 || T < 1 || NH < 1 || C % NH || offset < 0) {
        fprintf(stderr, "attention_paged: invalid arguments\n");
        return;
    }
    pa_handle* h = registry_find_by_page(key_blocks[0]);
    if (!h || h->C != C) {
        fprintf(stderr, "attention_paged: key_blocks do not belong to a live block manager\n");
        return;
    }
    const int bs = h->mgr->block_size;
    const size_t page_floats = (size_t)bs * C;
    const int n_pages = (T - 1 + offset) / bs + 1;        /* pages the reference touches (:190) */
    /* scratch lives in the handle and only ever grows: no allocation per call on the decode loop */
    const size_t need_ints = (size_t)n_pages + (size_t)B * 4;
    if (need_ints > h->compat_ints_cap) {
        int* p = (int*)realloc(h->compat_ints, need_ints * 2 * sizeof(int));
        if (!p) { fprintf(stderr, "attention_paged: out of memory\n"); return; }
        h->compat_ints = p; h->compat_ints_cap = need_ints * 2;
    }
    if ((size_t)B > h->compat_rows_cap) {
        const int** p = (const int**)realloc((void*)h->compat_rows, (size_t)B * 2 * sizeof(int*));
        if (!p) { fprintf(stderr, "attention_paged: out of memory\n"); return; }
        h->compat_rows = p; h->compat_rows_cap = (size_t)B * 2;
    }
    int* table = h->compat_ints;
    int* ints = h->compat_ints + n_pages;
    const int** rows = h->compat_rows;
    for (int i = 0; i < n_pages; i++) {
        size_t koff = (size_t)(key_blocks[i] - h->pool_k), voff = (size_t)(value_blocks[i] - h->pool_v);
        if (koff % page_floats || koff != voff || koff / page_floats >= (size_t)h->cfg.max_blocks) {
            fprintf(stderr, "attention_paged: page %d is not a (keys, values) pair of this pool\n", i);
            return;
        }
        table[i] = (int)(koff / page_floats);
    }
    /* key_blocks/value_blocks are shared by every b (:190 has no b) */
    int* np = ints, *ks = ints + B, *ke = ints + 2 * B, *nq = ints + 3 * B;
    for (int b = 0; b < B; b++) { rows[b] = table; np[b] = n_pages; ks[b] = offset; ke[b] = offset + T; nq[b] = T; }
    /* create_block_manager(channels) cannot know the head count (block_manager.c:38): the manager learns it
     * from the first attention_paged and keeps it (a compat handle sized its workspace for any NH dividing C) */
    if (h->cfg.n_heads != NH) {
        if (!h->compat) { fprintf(stderr, "attention_paged: NH=%d does not match the handle (%d heads)\n", NH, h->cfg.n_heads); return; }
        h->cfg.n_heads = NH;
        h->cfg.head_dim = C / NH;
    }
    if (pa_step_begin_raw(h, B, rows, np, ks, ke, nq) != PA_OK || pa_step_upload(h, h->stream) != PA_OK) {
        fprintf(stderr, "attention_paged: %s\n", pa_last_error());
        return;
    }
    {
        const size_t n_in = (size_t)B * T * 3 * C, n_out = (size_t)B * T * C;
        const int in_dev = pa_cu_is_device_ptr(inp), out_dev = pa_cu_is_device_ptr(out);
        const float* d_in = inp;
        float* d_out = out;
        if (!in_dev || !out_dev) {
            if (pa_cu_ensure_stage(h, n_in + n_out) != PA_OK) { fprintf(stderr, "attention_paged: %s\n", pa_last_error()); return; }
        }
        if (!in_dev) {
            memcpy(h->h_stage, inp, n_in * sizeof(float));
            if (pa_memcpy_h2d(h->d_stage, h->h_stage, n_in * sizeof(float), h->stream) != PA_OK) {
                fprintf(stderr, "attention_paged: %s\n", pa_last_error());
                return;
            }
            d_in = h->d_stage;
        }
        if (!out_dev) d_out = h->d_stage + n_in;
        if (pa_prefill(h, 0, d_in, 3 * C, d_out, C, h->stream) != PA_OK) {
            fprintf(stderr, "attention_paged: %s\n", pa_last_error());
            return;
        }
        if (!out_dev) {
            if (pa_memcpy_d2h(h->h_stage + n_in, d_out, n_out * sizeof(float), h->stream) != PA_OK ||
                pa_stream_sync(h->stream) != PA_OK) {
                fprintf(stderr, "attention_paged: %s\n", pa_last_error());
                return;
            }
            memcpy(out, h->h_stage + n_in, n_out * sizeof(float));
        } else if (pa_stream_sync(h->stream) != PA_OK) {
            fprintf(stderr, "attention_paged: %s\n", pa_last_error());
        }
    }
}
