/*
 * pa_sharing.c -- allocator extensions the fork's comments point at, host side,
 * plain C:
 *   - parallel sampling (`int P = 5`, unused, paged_infer.c:958): pa_seq_fork shares a sequence's
 *     full pages between copies, reference-counted;
 *   - prefix sharing by hashing ("if you do hashing you need to change up this policy",
 *     block_manager.c:109-111): pa_prefix_insert / pa_prefix_match.
 * Everything here is OFF unless called: with every refcount at 0/1 and no prefix cache the block
 * manager reproduces the reference trace bit for bit (tests/test_block_manager.py keeps checking
 * that against the compiled reference).
 *
 * Invariants: refcount[i] = sequences whose table lists page i, + 1 if the prefix cache holds it;
 * a shared page is FULL and read-only (appends only ever touch a sequence's private last page);
 * KVBlock.prompt_id of a used page is always one of its holders (the LRU evicts whole prompts
 * through it, block_manager.c:104-113) or PA_OWNER_CACHE.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_internal.h"

typedef struct prefix_entry {
    uint64_t key;            /* chained hash of all token ids up to and including this page */
    int page;
    int used;
} prefix_entry;

typedef struct prefix_cache {
    prefix_entry* slots;     /* open addressing, capacity a power of two >= 2 * max_blocks */
    int cap;
    int n_pages;
    int* page_tokens;        /* [max_blocks][block_size] token ids of a cached page (exact match on lookup) */
    uint64_t* page_key;      /* [max_blocks] key under which the page is registered, 0 = not cached */
} prefix_cache;

static uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 33);
}
static uint64_t chain_key(uint64_t parent, const int* tokens, int n) {
    uint64_t h = parent ^ 0x2545F4914F6CDD1Dull;
    for (int i = 0; i < n; i++) h = mix(h, (uint64_t)(uint32_t)tokens[i]);
    return h ? h : 1;
}

static prefix_cache* cache_get(BlockManager* m, int create) {
    if (m->prefix_cache || !create) return (prefix_cache*)m->prefix_cache;
    prefix_cache* c = (prefix_cache*)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->cap = 16;
    while (c->cap < 2 * m->max_blocks) c->cap <<= 1;
    c->slots = (prefix_entry*)calloc((size_t)c->cap, sizeof(prefix_entry));
    c->page_tokens = (int*)calloc((size_t)m->max_blocks * m->block_size, sizeof(int));
    c->page_key = (uint64_t*)calloc((size_t)m->max_blocks, sizeof(uint64_t));
    if (!c->slots || !c->page_tokens || !c->page_key) { free(c->slots); free(c->page_tokens); free(c->page_key); free(c); return NULL; }
    m->prefix_cache = c;
    return c;
}
void pa_share_destroy(BlockManager* m) {
    prefix_cache* c = (prefix_cache*)m->prefix_cache;
    if (!c) return;
    free(c->slots); free(c->page_tokens); free(c->page_key); free(c);
    m->prefix_cache = NULL;
}
static prefix_entry* cache_find(prefix_cache* c, uint64_t key) {
    for (int i = (int)(key & (uint64_t)(c->cap - 1)), n = 0; n < c->cap; i = (i + 1) & (c->cap - 1), n++) {
        if (!c->slots[i].used) return NULL;
        if (c->slots[i].used == 1 && c->slots[i].key == key) return &c->slots[i];
    }
    return NULL;
}
static void cache_put(prefix_cache* c, uint64_t key, int page) {
    for (int i = (int)(key & (uint64_t)(c->cap - 1));; i = (i + 1) & (c->cap - 1)) {
        if (c->slots[i].used != 1) { c->slots[i].key = key; c->slots[i].page = page; c->slots[i].used = 1; return; }
    }
}
static void cache_remove(prefix_cache* c, uint64_t key) {
    prefix_entry* e = cache_find(c, key);
    if (e) e->used = 2;      /* tombstone */
}

int pa_share_other_holder(BlockManager* m, int p, int idx) {
    for (int q = 0; q < m->max_prompts; q++) {
        if (q == p) continue;
        const int* row = m->prompt_block_list[q];
        for (int i = 0; i < m->prompt_block_count[q]; i++)
            if (row[i] == idx) return q;
    }
    prefix_cache* c = (prefix_cache*)m->prefix_cache;
    if (c && c->page_key[idx]) return PA_OWNER_CACHE;
    return -1;
}

/* the least recently used page that only the cache holds goes back to the free list */
int pa_share_evict_one_cached(BlockManager* m) {
    prefix_cache* c = (prefix_cache*)m->prefix_cache;
    if (!c || c->n_pages == 0) return 0;
    int victim = -1;
    for (int i = 0; i < m->max_blocks; i++)
        if (c->page_key[i] && m->refcount[i] == 1 && (victim < 0 || m->blocks[i].lru_counter < m->blocks[victim].lru_counter)) victim = i;
    if (victim < 0) return 0;
    cache_remove(c, c->page_key[victim]);
    c->page_key[victim] = 0;
    c->n_pages--;
    m->refcount[victim] = 0;
    m->blocks[victim].keys = m->blocks[victim].values = NULL;
    m->blocks[victim].filled = 0;
    m->blocks[victim].prompt_id = -1;
    return 1;
}

static int bad_seq(pa_handle* h, int s) { return !h || s < 0 || s >= h->cfg.max_seqs; }

int pa_seq_fork(pa_handle* h, int src, int dst) {
    if (bad_seq(h, src) || bad_seq(h, dst) || src == dst) { pa_set_error("pa_seq_fork: bad sequence ids"); return PA_ERR_INVALID; }
    BlockManager* m = h->mgr;
    if (m->prompt_block_count[dst] != 0) { pa_set_error("pa_seq_fork: sequence %d is not empty", dst); return PA_ERR_INVALID; }
    const int n = m->prompt_block_count[src];
    if (n == 0) return PA_OK;
    const int last = m->prompt_block_list[src][n - 1];
    const int last_rows = m->blocks[last].filled;
    const int shared = last_rows >= m->block_size ? n : n - 1;      /* a partial last page is copied, not shared */
    for (int i = 0; i < shared; i++) {
        const int idx = m->prompt_block_list[src][i];
        m->prompt_block_list[dst][i] = idx;
        m->refcount[idx]++;
    }
    m->prompt_block_count[dst] = shared;
    if (shared < n) {
        /* may evict another sequence -- never src or dst: both are pinned while the page for the copy is found */
        const unsigned char pin_src = m->pinned[src], pin_dst = m->pinned[dst];
        m->pinned[src] = m->pinned[dst] = 1;
        KVBlock* b = request_block(m, dst);
        m->pinned[src] = pin_src; m->pinned[dst] = pin_dst;
        if (!b || m->prompt_block_count[src] != n || m->prompt_block_list[src][n - 1] != last) {
            pa_set_error("pa_seq_fork: no page for the copy of the last page");
            free_blocks_for_prompt(m, dst);
            return PA_ERR_NO_BLOCKS;
        }
        b->filled = last_rows;
        int rc = pa_cu_copy_page_rows(h, last, (int)(b - m->blocks), last_rows);
        if (rc != PA_OK) { free_blocks_for_prompt(m, dst); return rc; }
    }
    return PA_OK;
}

int pa_prefix_insert(pa_handle* h, int seq, const int* tokens, int n_tokens) {
    if (bad_seq(h, seq) || !tokens || n_tokens < 0) { pa_set_error("pa_prefix_insert: bad arguments"); return PA_ERR_INVALID; }
    BlockManager* m = h->mgr;
    const int bs = m->block_size;
    if (n_tokens > pa_bm_context_len(m, seq)) { pa_set_error("pa_prefix_insert: %d tokens, %d cached", n_tokens, pa_bm_context_len(m, seq)); return PA_ERR_INVALID; }
    prefix_cache* c = cache_get(m, 1);
    if (!c) { pa_set_error("pa_prefix_insert: out of host memory"); return PA_ERR_NOMEM; }
    uint64_t key = 0;
    int added = 0;
    for (int i = 0; (i + 1) * bs <= n_tokens; i++) {
        key = chain_key(key, tokens + (size_t)i * bs, bs);
        const int idx = m->prompt_block_list[seq][i];
        prefix_entry* e = cache_find(c, key);
        if (e) continue;                       /* this prefix is cached already (by this or another page) */
        if (c->page_key[idx]) continue;        /* the page is registered under another prefix: leave it */
        cache_put(c, key, idx);
        c->page_key[idx] = key;
        memcpy(c->page_tokens + (size_t)idx * bs, tokens + (size_t)i * bs, (size_t)bs * sizeof(int));
        m->refcount[idx]++;                    /* the cache's own hold */
        c->n_pages++;
        added++;
    }
    return added;
}

int pa_prefix_match(pa_handle* h, int seq, const int* tokens, int n_tokens) {
    if (bad_seq(h, seq) || !tokens || n_tokens < 0) { pa_set_error("pa_prefix_match: bad arguments"); return PA_ERR_INVALID; }
    BlockManager* m = h->mgr;
    if (m->prompt_block_count[seq] != 0) { pa_set_error("pa_prefix_match: sequence %d is not empty", seq); return PA_ERR_INVALID; }
    prefix_cache* c = cache_get(m, 0);
    if (!c) return 0;
    const int bs = m->block_size;
    uint64_t key = 0;
    int n = 0;
    for (int i = 0; (i + 1) * bs < n_tokens && i < m->table_stride; i++) {      /* '<': leave at least one token to compute */
        key = chain_key(key, tokens + (size_t)i * bs, bs);
        prefix_entry* e = cache_find(c, key);
        if (!e || memcmp(c->page_tokens + (size_t)e->page * bs, tokens + (size_t)i * bs, (size_t)bs * sizeof(int)) != 0) break;
        const int idx = e->page;
        m->prompt_block_list[seq][n++] = idx;
        m->refcount[idx]++;
        if (m->blocks[idx].prompt_id < 0) m->blocks[idx].prompt_id = seq;        /* was held by the cache only */
        m->blocks[idx].lru_counter = ++m->lru_epoch;                            /* a hit is a use */
    }
    m->prompt_block_count[seq] = n;
    return n * bs;
}

int pa_prefix_cached_pages(pa_handle* h) {
    prefix_cache* c = h ? (prefix_cache*)h->mgr->prefix_cache : NULL;
    return c ? c->n_pages : 0;
}
int pa_page_refcount(pa_handle* h, int page) {
    if (!h || page < 0 || page >= h->mgr->max_blocks) return PA_ERR_INVALID;
    return h->mgr->refcount[page];
}

/* ---- swap-out instead of drop-on-evict ---------------------------------------------------------
 * The reference drops the victim prompt's KV when the pool is exhausted (page_out_lru_block,
 * block_manager.c:104-113) and the prompt has to be prefilled again.  With swapping on, the
 * allocator's eviction first copies the victim's pages (every layer) to host memory; the sequence
 * is brought back -- into whatever pages are free then -- the next time a step names it. */
typedef struct swap_slot {
    float* buf;          /* [pages][2 (K,V)][n_layers][block_size*C]; pinned when there is a device */
    int n_tokens;
    int n_pages;
} swap_slot;

/* Host copies live in PINNED memory: the page copies are then truly asynchronous DMA transfers on the handle's
 * stream (pageable memory would be staged by the driver, synchronously); all pages of a sequence are enqueued
 * back to back and waited for ONCE (a swap-out before the victim's pages are handed on -- the caller may run its
 * kernels on another stream --, a swap-in before the host copy is released). */
static float* swap_alloc(pa_handle* h, size_t bytes) {
    return (float*)(h->host_only ? malloc(bytes) : pa_host_alloc(bytes));
}
static void swap_free(pa_handle* h, float* p) {
    if (!p) return;
    if (h->host_only) free(p); else pa_host_free(p);          /* (freeing pinned memory waits for the device) */
}

static swap_slot* swap_slots(pa_handle* h, int create) {
    if (!h->swap_state && create) h->swap_state = calloc((size_t)h->cfg.max_seqs, sizeof(swap_slot));
    return (swap_slot*)h->swap_state;
}
static size_t swap_page_floats(const pa_handle* h) { return (size_t)2 * h->cfg.n_layers * h->cfg.block_size * h->C; }

int pa_set_evict_swap(pa_handle* h, int enable) {
    if (!h) return PA_ERR_INVALID;
    if (enable && !swap_slots(h, 1)) { pa_set_error("pa_set_evict_swap: out of host memory"); return PA_ERR_NOMEM; }
    h->swap_enabled = enable ? 1 : 0;
    return PA_OK;
}
int pa_seq_swapped_tokens(pa_handle* h, int seq) {
    if (bad_seq(h, seq)) return PA_ERR_INVALID;
    swap_slot* sl = swap_slots(h, 0);
    return sl && sl[seq].buf ? sl[seq].n_tokens : 0;
}
void pa_swap_destroy(pa_handle* h) {
    swap_slot* sl = swap_slots(h, 0);
    if (!sl) return;
    for (int i = 0; i < h->cfg.max_seqs; i++) swap_free(h, sl[i].buf);
    free(sl);
    h->swap_state = NULL;
}

int pa_seq_swap_out(pa_handle* h, int seq) {
    if (bad_seq(h, seq)) { pa_set_error("pa_seq_swap_out: bad sequence id"); return PA_ERR_INVALID; }
    BlockManager* m = h->mgr;
    const int n = m->prompt_block_count[seq];
    if (n == 0) return PA_OK;
    swap_slot* sl = swap_slots(h, 1);
    if (!sl) { pa_set_error("pa_seq_swap_out: out of host memory"); return PA_ERR_NOMEM; }
    const size_t pf = swap_page_floats(h);
    swap_free(h, sl[seq].buf);
    sl[seq].buf = swap_alloc(h, (size_t)n * pf * sizeof(float));
    if (!sl[seq].buf) { pa_set_error("pa_seq_swap_out: out of host memory (%d pages)", n); return PA_ERR_NOMEM; }
    sl[seq].n_tokens = pa_bm_context_len(m, seq);
    sl[seq].n_pages = n;
    for (int i = 0; i < n; i++) {
        float* k = sl[seq].buf + (size_t)i * pf;
        int rc = pa_cu_swap_page(h, m->prompt_block_list[seq][i], k, k + pf / 2, 1);
        if (rc != PA_OK) { swap_free(h, sl[seq].buf); sl[seq].buf = NULL; return rc; }
    }
    {   /* one wait for the whole sequence: its pages may be handed to work on ANOTHER stream right away */
        int rc = pa_cu_swap_sync(h);
        if (rc != PA_OK) { swap_free(h, sl[seq].buf); sl[seq].buf = NULL; return rc; }
    }
    free_blocks_for_prompt(m, seq);
    return PA_OK;
}

/* allocator hook: the victim is about to lose its pages; keep a host copy.  A copy that cannot be made
 * degrades to the reference's behaviour (the sequence is dropped and must be prefilled again) -- loudly:
 * stderr like the reference's allocator messages, and counted in pa_swap_failures(). */
static void swap_failed(pa_handle* h, int p, swap_slot* sl, const char* why) {
    if (sl && sl[p].buf) { swap_free(h, sl[p].buf); sl[p].buf = NULL; }
    h->swap_failures++;
    fprintf(stderr, "Swap-out of prompt %d failed (%s); its KV is dropped.\n", p, why);
}
void pa_swap_on_evict(pa_handle* h, int p) {
    if (bad_seq(h, p)) return;
    BlockManager* m = h->mgr;
    const int n = m->prompt_block_count[p];
    if (n == 0) return;
    swap_slot* sl = swap_slots(h, 1);
    if (!sl) { swap_failed(h, p, NULL, "out of host memory"); return; }
    const size_t pf = swap_page_floats(h);
    swap_free(h, sl[p].buf);
    sl[p].buf = swap_alloc(h, (size_t)n * pf * sizeof(float));
    if (!sl[p].buf) { swap_failed(h, p, sl, "out of host memory"); return; }
    sl[p].n_tokens = pa_bm_context_len(m, p);
    sl[p].n_pages = n;
    for (int i = 0; i < n; i++) {
        float* k = sl[p].buf + (size_t)i * pf;
        if (pa_cu_swap_page(h, m->prompt_block_list[p][i], k, k + pf / 2, 1) != PA_OK) { swap_failed(h, p, sl, pa_last_error()); return; }
    }
    if (pa_cu_swap_sync(h) != PA_OK) swap_failed(h, p, sl, pa_last_error());
}
int pa_swap_failures(pa_handle* h) { return h ? h->swap_failures : PA_ERR_INVALID; }

int pa_seq_swap_in(pa_handle* h, int seq) {
    if (bad_seq(h, seq)) { pa_set_error("pa_seq_swap_in: bad sequence id"); return PA_ERR_INVALID; }
    swap_slot* sl = swap_slots(h, 0);
    if (!sl || !sl[seq].buf) return PA_OK;
    BlockManager* m = h->mgr;
    if (m->prompt_block_count[seq] != 0) { pa_set_error("pa_seq_swap_in: sequence %d holds pages and a swap copy", seq); return PA_ERR_INVALID; }
    float* buf = sl[seq].buf;          /* detach first: the allocations below may evict (and swap out) other sequences */
    const int n = sl[seq].n_pages, n_tokens = sl[seq].n_tokens;
    sl[seq].buf = NULL;
    const size_t pf = swap_page_floats(h);
    int rc = PA_OK;
    for (int i = 0; i < n && rc == PA_OK; i++) {
        KVBlock* b = request_block(m, seq);
        if (!b || m->prompt_block_count[seq] != i + 1) { pa_set_error("pa_seq_swap_in: No blocks available (sequence %d)", seq); rc = PA_ERR_NO_BLOCKS; break; }
        b->filled = (i + 1 < n) ? m->block_size : n_tokens - (n - 1) * m->block_size;
        float* k = buf + (size_t)i * pf;
        rc = pa_cu_swap_page(h, (int)(b - m->blocks), k, k + pf / 2, 0);
    }
    if (rc == PA_OK) rc = pa_cu_swap_sync(h);       /* the pages are in place before the host copy goes away */
    if (rc != PA_OK) {                 /* put the copy back so nothing is lost */
        free_blocks_for_prompt(m, seq);
        swap_free(h, sl[seq].buf);
        sl[seq].buf = buf; sl[seq].n_pages = n; sl[seq].n_tokens = n_tokens;
        return rc;
    }
    swap_free(h, buf);
    return PA_OK;
}
int pa_swap_in_if_needed(pa_handle* h, int seq) {
    swap_slot* sl = swap_slots(h, 0);
    if (!sl || bad_seq(h, seq) || !sl[seq].buf) return PA_OK;
    return pa_seq_swap_in(h, seq);
}
