/*
 * pa_step.c -- host side, plain C: handle life cycle and the per-step scheduling that turns
 * block-manager state into the int32 tables the kernels read (block table rows, context
 * lengths, page prefix sums, slot mapping).  Integer work only; the device is reached through
 * the pa_cu_* layer in pa_cuda.cu.
 *
 * Reference call site this generalises: paged_infer.c:710-715 (add_to_cache ->
 * collect_kv_blocks -> attention_paged, one sequence, one layer).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_internal.h"

static __thread char g_err[512] = "";
__thread int pa_pdl_enabled = 1;      /* programmatic dependent launch for the step's kernel chain (pa_pdl.cuh) */
__thread int pa_pdl_gate = 1;
__thread int pa_launch_cooperative = 0;

void pa_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* pa_last_error(void) { return g_err; }
const char* pa_version(void) { return "paged_attn_b200 0.1 (sm_100a)"; }

static int create_impl(const pa_config* cfg, pa_handle** out, int compat);

int pa_create(const pa_config* cfg, pa_handle** out) { return create_impl(cfg, out, 0); }
/* create_block_manager(): the head count is unknown until attention_paged, so the split
 * workspace is sized for any NH that divides C */
int pa_create_compat(const pa_config* cfg, pa_handle** out) { return create_impl(cfg, out, 1); }

static int create_impl(const pa_config* cfg, pa_handle** out, int compat) {
    if (!cfg || !out) { pa_set_error("pa_create: NULL argument"); return PA_ERR_INVALID; }
    *out = NULL;
    if (cfg->block_size < 1 || cfg->max_blocks < 1 || cfg->max_seqs < 1 || cfg->n_layers < 1 ||
        cfg->n_heads < 1 || cfg->head_dim < 1 || cfg->max_blocks_per_seq < 0) {
        pa_set_error("pa_create: invalid geometry (bs=%d blocks=%d seqs=%d layers=%d NH=%d hs=%d)",
                     cfg->block_size, cfg->max_blocks, cfg->max_seqs, cfg->n_layers, cfg->n_heads, cfg->head_dim);
        return PA_ERR_INVALID;
    }
    if ((long long)cfg->max_blocks * cfg->block_size > 0x7fffffffLL) {
        pa_set_error("pa_create: max_blocks*block_size overflows int32 slots");
        return PA_ERR_INVALID;
    }
    pa_handle* h = (pa_handle*)calloc(1, sizeof(pa_handle));
    if (!h) { pa_set_error("pa_create: out of host memory"); return PA_ERR_NOMEM; }
    h->cfg = *cfg;
    if (h->cfg.max_blocks_per_seq == 0 || h->cfg.max_blocks_per_seq > cfg->max_blocks)
        h->cfg.max_blocks_per_seq = cfg->max_blocks;
    if (h->cfg.max_batch_tokens <= 0) h->cfg.max_batch_tokens = cfg->max_seqs;
    h->C = cfg->n_heads * cfg->head_dim;
    h->host_only = (cfg->device == PA_HOST_ONLY);
    h->compat = compat;
    h->max_heads = compat ? h->C : cfg->n_heads;
    h->layer_stride = (size_t)cfg->max_blocks * cfg->block_size * h->C;
    h->step_seq_ids = (int*)calloc((size_t)cfg->max_seqs, sizeof(int));
    h->step_n_new = (int*)calloc((size_t)cfg->max_seqs, sizeof(int));
    h->step_pre_len = (int*)calloc((size_t)cfg->max_seqs, sizeof(int));
    h->slot_scratch = (int*)calloc((size_t)h->cfg.max_batch_tokens + 1, sizeof(int));
    h->mgr = pa_bm_create(h, h->C, cfg->block_size, cfg->max_blocks, cfg->max_seqs, h->cfg.max_blocks_per_seq);
    if (!h->mgr || !h->step_seq_ids || !h->step_n_new || !h->step_pre_len || !h->slot_scratch) {
        pa_set_error("pa_create: out of host memory");
        pa_destroy(h);
        return PA_ERR_NOMEM;
    }
    if (!h->host_only) {
        int rc = pa_cu_init(h);
        if (rc == PA_OK) rc = pa_cu_alloc_pool(h);
        if (rc != PA_OK) { pa_destroy(h); return rc; }
    }
    *out = h;
    return PA_OK;
}

void pa_destroy(pa_handle* h) {
    if (!h) return;
    pa_swap_destroy(h);
    pa_cu_release(h);
    pa_bm_destroy(h->mgr);
    free(h->step_seq_ids);
    free(h->step_n_new);
    free(h->step_pre_len);
    free(h->slot_scratch);
    free(h->compat_ints);
    free((void*)h->compat_rows);
    free(h);
}

BlockManager* pa_manager(pa_handle* h) { return h ? h->mgr : NULL; }
float* pa_pool_k(pa_handle* h, int layer) { return (h && h->pool_k) ? h->pool_k + (size_t)layer * h->layer_stride : NULL; }
float* pa_pool_v(pa_handle* h, int layer) { return (h && h->pool_v) ? h->pool_v + (size_t)layer * h->layer_stride : NULL; }
size_t pa_pool_bytes(pa_handle* h) { return h ? 2 * (size_t)h->cfg.n_layers * h->layer_stride * sizeof(float) : 0; }
int pa_device(pa_handle* h) { return h ? h->cfg.device : PA_HOST_ONLY; }
void* pa_stream_of(pa_handle* h) { return h ? h->stream : NULL; }
int pa_sm_count(pa_handle* h) { return h ? h->sm_count : 0; }

int pa_tune_set(pa_handle* h, int key, int value) {
    if (!h || key < 0 || key >= PA_TUNE_MAX || key == PA_TUNE_COUNT_LAUNCHES ||
        (key >= PA_TUNE_LAST_HPG && key <= PA_TUNE_LAST_GRID)) return PA_ERR_INVALID;
    h->tune[key] = value;
    return PA_OK;
}
int pa_tune_get(pa_handle* h, int key) {
    if (!h || key < 0 || key >= PA_TUNE_MAX) return PA_ERR_INVALID;
    if (key == PA_TUNE_COUNT_LAUNCHES) return (int)h->launches;
    return h->tune[key];
}

/* ------------------------------------------------------------------------------------------
 * step tables
 * ---------------------------------------------------------------------------------------- */
static int check_batch(pa_handle* h, const int* seq_ids, int nseq, const char* who) {
    if (!h || !seq_ids || nseq < 1 || nseq > h->cfg.max_seqs) {
        pa_set_error("%s: bad batch (nseq=%d, max_seqs=%d)", who, nseq, h ? h->cfg.max_seqs : 0);
        return PA_ERR_INVALID;
    }
    for (int i = 0; i < nseq; i++)
        if (seq_ids[i] < 0 || seq_ids[i] >= h->cfg.max_seqs) {
            pa_set_error("%s: Invalid prompt ID %d", who, seq_ids[i]);
            return PA_ERR_INVALID;
        }
    return PA_OK;
}

static int pages_between(int kv_start, int kv_end, int bs) {
    if (kv_end <= kv_start) return 0;
    return (kv_end + bs - 1) / bs - kv_start / bs;
}

/* (Re)build everything except the slot mapping from the manager's current state. */
static int build_tables(pa_handle* h, int nseq, int ntok, const int* kv_start_in) {
    BlockManager* m = h->mgr;
    const int bs = m->block_size;
    pa_step_layout L;
    memset(&L, 0, sizeof(L));
    L.nseq = nseq;
    L.ntok = ntok;
    int max_pages = 1;
    for (int i = 0; i < nseq; i++) {
        int n = m->prompt_block_count[h->step_seq_ids[i]];
        if (n > max_pages) max_pages = n;
    }
    L.tstride = (max_pages + 3) & ~3;
    int off = 0;
    L.off_kv_end = off;     off += nseq;
    L.off_kv_start = off;   off += nseq;
    L.off_cum_pages = off;  off += nseq + 1;
    L.off_q_row0 = off;     off += nseq + 1;
    L.off_slot = off;       off += ntok;
    off = (off + 3) & ~3;
    L.off_table = off;      off += nseq * L.tstride;
    L.total_ints = off;

    int* buf = pa_cu_step_host_buffer(h, (size_t)L.total_ints);
    if (!buf) return PA_ERR_NOMEM;
    if (ntok > 0) memcpy(buf + L.off_slot, h->slot_scratch, (size_t)ntok * sizeof(int));
    h->h_step = buf;

    int cum = 0, qrow = 0, max_q = 0;
    for (int i = 0; i < nseq; i++) {
        int p = h->step_seq_ids[i];
        int end = pa_bm_context_len(m, p);
        int start = kv_start_in ? kv_start_in[i] : 0;
        if (start < 0) start = 0;
        if (start > end) start = end;
        buf[L.off_kv_end + i] = end;
        buf[L.off_kv_start + i] = start;
        buf[L.off_cum_pages + i] = cum;
        cum += pages_between(start, end, bs);
        buf[L.off_q_row0 + i] = qrow;
        qrow += h->step_n_new[i];
        if (h->step_n_new[i] > max_q) max_q = h->step_n_new[i];
        int n = m->prompt_block_count[p];
        int* row = buf + L.off_table + (size_t)i * L.tstride;
        memcpy(row, m->prompt_block_list[p], (size_t)n * sizeof(int));
        for (int j = n; j < L.tstride; j++) row[j] = 0;
    }
    buf[L.off_cum_pages + nseq] = cum;
    buf[L.off_q_row0 + nseq] = qrow;
    L.total_pages = cum;
    L.max_q = max_q;
    h->step = L;
    return PA_OK;
}

/* The sequences of the step being built are PINNED while pages are placed: the whole-prompt LRU eviction
 * inside request_block (block_manager.c:104-113) must never take a sequence whose slots were already
 * handed out in this step, nor the requester itself.  Returns PA_ERR_INVALID on a duplicate id. */
static int pin_step(pa_handle* h, const int* seq_ids, int nseq, const char* who) {
    unsigned char* pinned = h->mgr->pinned;
    for (int i = 0; i < nseq; i++) {
        if (pinned[seq_ids[i]]) {
            for (int j = 0; j < i; j++) pinned[seq_ids[j]] = 0;
            pa_set_error("%s: sequence %d is named twice in one step", who, seq_ids[i]);
            return PA_ERR_INVALID;
        }
        pinned[seq_ids[i]] = 1;
    }
    return PA_OK;
}
static void unpin_step(pa_handle* h, const int* seq_ids, int nseq) {
    for (int i = 0; i < nseq; i++) h->mgr->pinned[seq_ids[i]] = 0;
}

/* a step that failed part-way leaves no trace in the sequences it had already touched: rows [0, upto]
 * go back to the length they had before the step (pages allocated for tokens that will never be written
 * return to the pool) */
static void undo_partial_step(pa_handle* h, int upto) {
    for (int j = 0; j <= upto; j++) {
        int p = h->step_seq_ids[j];
        if (pa_bm_context_len(h->mgr, p) > h->step_pre_len[j]) pa_seq_truncate(h, p, h->step_pre_len[j]);
    }
}

int pa_step_begin(pa_handle* h, const int* seq_ids, const int* n_new, int nseq) {
    int rc = check_batch(h, seq_ids, nseq, "pa_step_begin");
    if (rc != PA_OK) return rc;
    if (!n_new) { pa_set_error("pa_step_begin: n_new is NULL"); return PA_ERR_INVALID; }
    BlockManager* m = h->mgr;
    const int bs = m->block_size;
    long long ntok = 0;
    for (int i = 0; i < nseq; i++) {
        if (n_new[i] < 0) { pa_set_error("pa_step_begin: n_new[%d] < 0", i); return PA_ERR_INVALID; }
        ntok += n_new[i];
    }
    if (ntok > h->cfg.max_batch_tokens) {
        pa_set_error("pa_step_begin: %lld new tokens > max_batch_tokens %d", ntok, h->cfg.max_batch_tokens);
        return PA_ERR_INVALID;
    }
    rc = pin_step(h, seq_ids, nseq, "pa_step_begin");
    if (rc != PA_OK) return rc;
    /* slot mapping goes to a scratch area first (table size is unknown until pages are placed) */
    int* slots = h->slot_scratch;
    int tok = 0;
    for (int i = 0; i < nseq && rc == PA_OK; i++) {
        int p = seq_ids[i];
        h->step_seq_ids[i] = p;
        h->step_n_new[i] = n_new[i];
        if (h->swap_enabled) {                       /* extension: a swapped-out sequence comes back before it grows */
            rc = pa_swap_in_if_needed(h, p);
            if (rc != PA_OK) { h->step_pre_len[i] = pa_bm_context_len(m, p); undo_partial_step(h, i - 1); break; }
        }
        h->step_pre_len[i] = pa_bm_context_len(m, p);
        int left = n_new[i];
        while (left > 0) {
            int idx = pa_bm_choose_page(m, p);
            if (idx < 0) {
                /* the pool is exhausted and every page left belongs to a sequence of this very step (or the
                 * per-sequence cap is reached): evicting one of them -- or the requester, as the single-prompt
                 * reference would (block_manager.c:157) -- would leave slots of this step pointing into pages
                 * their sequence no longer owns, so this is an error */
                pa_set_error("pa_step_begin: No blocks available (sequence %d; the sequences of a step are never evicted for it)", p);
                undo_partial_step(h, i);
                rc = PA_ERR_NO_BLOCKS;
                break;
            }
            KVBlock* b = &m->blocks[idx];
            int take = bs - b->filled;
            if (take > left) take = left;
            for (int r = 0; r < take; r++) slots[tok++] = idx * bs + b->filled + r;
            b->filled += take;
            left -= take;
        }
    }
    unpin_step(h, seq_ids, nseq);
    if (rc != PA_OK) { h->step.nseq = 0; return rc; }
    return build_tables(h, nseq, tok, NULL);
}

int pa_step_begin_raw(pa_handle* h, int nseq, const int* const* tables, const int* n_pages,
                      const int* kv_start, const int* kv_end, const int* n_q) {
    if (!h || nseq < 1 || !tables || !n_pages || !kv_start || !kv_end || !n_q) {
        pa_set_error("pa_step_begin_raw: bad arguments");
        return PA_ERR_INVALID;
    }
    const int bs = h->mgr->block_size;
    pa_step_layout L;
    memset(&L, 0, sizeof(L));
    L.nseq = nseq;
    int max_pages = 1, ntok = 0;
    for (int i = 0; i < nseq; i++) {
        if (n_pages[i] > max_pages) max_pages = n_pages[i];
        ntok += n_q[i];
    }
    L.ntok = ntok;
    L.tstride = (max_pages + 3) & ~3;
    int off = 0;
    L.off_kv_end = off;     off += nseq;
    L.off_kv_start = off;   off += nseq;
    L.off_cum_pages = off;  off += nseq + 1;
    L.off_q_row0 = off;     off += nseq + 1;
    L.off_slot = off;       off += ntok;
    off = (off + 3) & ~3;
    L.off_table = off;      off += nseq * L.tstride;
    L.total_ints = off;
    int* buf = pa_cu_step_host_buffer(h, (size_t)L.total_ints);
    if (!buf) return PA_ERR_NOMEM;
    h->h_step = buf;
    int cum = 0, qrow = 0;
    for (int i = 0; i < nseq; i++) {
        int end = kv_end[i], start = kv_start[i];
        if (start < 0) start = 0;
        if (start > end) start = end;
        if ((end + bs - 1) / bs > n_pages[i]) {
            pa_set_error("pa_step_begin_raw: row %d needs %d pages, has %d", i, (end + bs - 1) / bs, n_pages[i]);
            return PA_ERR_INVALID;
        }
        buf[L.off_kv_end + i] = end;
        buf[L.off_kv_start + i] = start;
        buf[L.off_cum_pages + i] = cum;
        cum += pages_between(start, end, bs);
        buf[L.off_q_row0 + i] = qrow;
        qrow += n_q[i];
        if (n_q[i] > L.max_q) L.max_q = n_q[i];
        int* row = buf + L.off_table + (size_t)i * L.tstride;
        memcpy(row, tables[i], (size_t)n_pages[i] * sizeof(int));
        for (int j = n_pages[i]; j < L.tstride; j++) row[j] = 0;
    }
    for (int j = 0; j < ntok; j++) buf[L.off_slot + j] = 0;   /* no append in a raw step */
    buf[L.off_cum_pages + nseq] = cum;
    buf[L.off_q_row0 + nseq] = qrow;
    L.total_pages = cum;
    h->step = L;
    return PA_OK;
}

int pa_step_begin_readonly(pa_handle* h, const int* seq_ids, int nseq) {
    int rc = check_batch(h, seq_ids, nseq, "pa_step_begin_readonly");
    if (rc != PA_OK) return rc;
    rc = pin_step(h, seq_ids, nseq, "pa_step_begin_readonly");      /* a swap-in below must not evict another row of this step */
    if (rc != PA_OK) return rc;
    for (int i = 0; i < nseq && rc == PA_OK; i++) {
        h->step_seq_ids[i] = seq_ids[i];
        h->step_n_new[i] = 0;
        if (h->swap_enabled) rc = pa_swap_in_if_needed(h, seq_ids[i]);
        h->step_pre_len[i] = pa_bm_context_len(h->mgr, seq_ids[i]);
    }
    unpin_step(h, seq_ids, nseq);
    if (rc != PA_OK) { h->step.nseq = 0; return rc; }
    return build_tables(h, nseq, 0, NULL);
}

int pa_step_set_kv_start(pa_handle* h, const int* kv_start) {
    if (!h || h->step.nseq < 1 || !h->h_step) { pa_set_error("pa_step_set_kv_start: no step"); return PA_ERR_INVALID; }
    /* prefix sums depend on the window, so rebuild in place (same sizes) */
    pa_step_layout* L = &h->step;
    const int bs = h->mgr->block_size;
    int* buf = h->h_step;
    int cum = 0;
    for (int i = 0; i < L->nseq; i++) {
        int end = buf[L->off_kv_end + i];
        int start = kv_start ? kv_start[i] : 0;
        if (start < 0) start = 0;
        if (start > end) start = end;
        buf[L->off_kv_start + i] = start;
        buf[L->off_cum_pages + i] = cum;
        cum += pages_between(start, end, bs);
    }
    buf[L->off_cum_pages + L->nseq] = cum;
    L->total_pages = cum;
    L->uploaded = 0;
    return PA_OK;
}

const int* pa_step_slot_mapping(pa_handle* h, int* n_tokens) {
    if (!h || !h->h_step) return NULL;
    if (n_tokens) *n_tokens = h->step.ntok;
    return h->h_step + h->step.off_slot;
}
const int* pa_step_context_lens(pa_handle* h, int* nseq) {
    if (!h || !h->h_step) return NULL;
    if (nseq) *nseq = h->step.nseq;
    return h->h_step + h->step.off_kv_end;
}
const int* pa_step_block_table(pa_handle* h, int* nseq, int* stride) {
    if (!h || !h->h_step) return NULL;
    if (nseq) *nseq = h->step.nseq;
    if (stride) *stride = h->step.tstride;
    return h->h_step + h->step.off_table;
}

/* Every address a kernel forms into the KV pool comes from these tables (page index * page size + row, slot * C)
 * inside loops bounded by kv_start / kv_end: checking the tables bounds the kernels' pool accesses.  Run on every
 * upload when PA_VALIDATE_STEP=1 (the test suite sets it; compute-sanitizer is closed on the B200 pool this was
 * developed on), O(batch x pages) integer compares. */
int pa_step_validate(pa_handle* h) {
    if (!h || h->step.nseq < 1 || !h->h_step) { pa_set_error("pa_step_validate: no step"); return PA_ERR_INVALID; }
    const pa_step_layout* L = &h->step;
    const int* buf = h->h_step;
    const int bs = h->mgr->block_size, mb = h->mgr->max_blocks;
    int cum = 0, qrow = 0;
    for (int i = 0; i < L->nseq; i++) {
        const int end = buf[L->off_kv_end + i], start = buf[L->off_kv_start + i];
        if (start < 0 || start > end) { pa_set_error("pa_step_validate: row %d window [%d, %d)", i, start, end); return PA_ERR_INVALID; }
        const int pages = (end + bs - 1) / bs;
        if (pages > L->tstride) { pa_set_error("pa_step_validate: row %d needs %d pages, table rows hold %d", i, pages, L->tstride); return PA_ERR_INVALID; }
        const int* row = buf + L->off_table + (size_t)i * L->tstride;
        for (int j = 0; j < L->tstride; j++)
            if (row[j] < 0 || row[j] >= mb) { pa_set_error("pa_step_validate: row %d page %d = %d outside the pool of %d pages", i, j, row[j], mb); return PA_ERR_INVALID; }
        if (buf[L->off_cum_pages + i] != cum) { pa_set_error("pa_step_validate: page prefix sum of row %d", i); return PA_ERR_INVALID; }
        cum += end > start ? (end + bs - 1) / bs - start / bs : 0;
        if (buf[L->off_q_row0 + i] != qrow) { pa_set_error("pa_step_validate: query-row prefix sum of row %d", i); return PA_ERR_INVALID; }
        const int nq = buf[L->off_q_row0 + i + 1] - qrow;
        if (nq < 0 || nq > end + 1) { pa_set_error("pa_step_validate: row %d has %d query rows for %d cached tokens", i, nq, end); return PA_ERR_INVALID; }
        qrow += nq;
    }
    if (buf[L->off_cum_pages + L->nseq] != cum || cum != L->total_pages) { pa_set_error("pa_step_validate: total pages"); return PA_ERR_INVALID; }
    if (qrow != L->ntok) { pa_set_error("pa_step_validate: %d query rows, %d tokens", qrow, L->ntok); return PA_ERR_INVALID; }
    for (int j = 0; j < L->ntok; j++) {
        const int sl = buf[L->off_slot + j];
        if (sl < 0 || sl >= mb * bs) { pa_set_error("pa_step_validate: slot %d of token %d outside the pool of %d rows", sl, j, mb * bs); return PA_ERR_INVALID; }
    }
    return PA_OK;
}

int pa_step_upload(pa_handle* h, void* stream) {
    if (!h || h->step.nseq < 1) { pa_set_error("pa_step_upload: no step"); return PA_ERR_INVALID; }
    if (h->host_only) { pa_set_error("pa_step_upload: host-only handle has no device"); return PA_ERR_NO_DEVICE; }
    static int validate = -1;
    if (validate < 0) { const char* e = getenv("PA_VALIDATE_STEP"); validate = (e && atoi(e) != 0) ? 1 : 0; }
    if (validate) {
        int rc = pa_step_validate(h);
        if (rc != PA_OK) return rc;
    }
    return pa_cu_step_upload(h, stream);
}

/* ------------------------------------------------------------------------------------------
 * sequence bookkeeping
 * ---------------------------------------------------------------------------------------- */
int pa_seq_len(pa_handle* h, int seq_id) {
    if (!h || seq_id < 0 || seq_id >= h->cfg.max_seqs) return PA_ERR_INVALID;
    return pa_bm_context_len(h->mgr, seq_id);
}

int pa_seq_free(pa_handle* h, int seq_id) {
    if (!h || seq_id < 0 || seq_id >= h->cfg.max_seqs) return PA_ERR_INVALID;
    free_blocks_for_prompt(h->mgr, seq_id);
    return PA_OK;
}

int pa_seq_truncate(pa_handle* h, int seq_id, int new_len) {
    if (!h || seq_id < 0 || seq_id >= h->cfg.max_seqs || new_len < 0) return PA_ERR_INVALID;
    BlockManager* m = h->mgr;
    int len = pa_bm_context_len(m, seq_id);
    if (new_len > len) { pa_set_error("pa_seq_truncate: %d > current length %d", new_len, len); return PA_ERR_INVALID; }
    const int bs = m->block_size;
    int keep = (new_len + bs - 1) / bs;
    int n = m->prompt_block_count[seq_id];
    if (keep > 0 && new_len - (keep - 1) * bs < bs && m->refcount[m->prompt_block_list[seq_id][keep - 1]] > 1) {
        /* `filled` belongs to the page, and a shared page is full for all of its holders */
        pa_set_error("pa_seq_truncate: position %d lies inside a page shared with another sequence", new_len);
        return PA_ERR_UNSUPPORTED;
    }
    m->prompt_block_count[seq_id] = keep;
    for (int i = keep; i < n; i++) pa_bm_release_page(m, seq_id, m->prompt_block_list[seq_id][i]);
    if (keep > 0) m->blocks[m->prompt_block_list[seq_id][keep - 1]].filled = new_len - (keep - 1) * bs;
    return PA_OK;
}

int pa_step_rollback(pa_handle* h) {
    if (!h || h->step.nseq < 1) { pa_set_error("pa_step_rollback: no step"); return PA_ERR_INVALID; }
    for (int i = 0; i < h->step.nseq; i++) {
        int p = h->step_seq_ids[i];
        int len = pa_bm_context_len(h->mgr, p) - h->step_n_new[i];
        int rc = pa_seq_truncate(h, p, len < 0 ? 0 : len);
        if (rc != PA_OK) return rc;
        h->step_n_new[i] = 0;
    }
    return PA_OK;
}

int pa_seq_adopt(pa_handle* h, int seq_id, const int* blocks, int n_blocks, int n_tokens) {
    if (!h || !blocks || seq_id < 0 || seq_id >= h->cfg.max_seqs) return PA_ERR_INVALID;
    BlockManager* m = h->mgr;
    const int bs = m->block_size;
    if (m->prompt_block_count[seq_id] != 0) { pa_set_error("pa_seq_adopt: sequence %d not empty", seq_id); return PA_ERR_INVALID; }
    if (n_blocks < 0 || n_blocks > m->table_stride || n_tokens > n_blocks * bs || n_tokens <= (n_blocks - 1) * bs) {
        pa_set_error("pa_seq_adopt: %d tokens do not fit %d pages exactly", n_tokens, n_blocks);
        return PA_ERR_INVALID;
    }
    for (int i = 0; i < n_blocks; i++) {
        int idx = blocks[i];
        if (idx < 0 || idx >= m->max_blocks || m->blocks[idx].prompt_id != -1) {
            pa_set_error("pa_seq_adopt: page %d invalid or in use", idx);
            for (int j = 0; j < i; j++) { m->blocks[blocks[j]].prompt_id = -1; m->blocks[blocks[j]].filled = 0; }
            return PA_ERR_INVALID;
        }
        m->blocks[idx].prompt_id = seq_id;     /* claim immediately so duplicates are caught */
    }
    for (int i = 0; i < n_blocks; i++) {
        int idx = blocks[i];
        KVBlock* b = &m->blocks[idx];
        size_t off = (size_t)idx * bs * m->C;
        b->keys = h->pool_k ? h->pool_k + off : NULL;
        b->values = h->pool_v ? h->pool_v + off : NULL;
        b->filled = (i + 1 < n_blocks) ? bs : n_tokens - (n_blocks - 1) * bs;
        b->lru_counter = ++m->lru_epoch;
        m->refcount[idx] = 1;
        m->prompt_block_list[seq_id][i] = idx;
    }
    m->prompt_block_count[seq_id] = n_blocks;
    return PA_OK;
}
