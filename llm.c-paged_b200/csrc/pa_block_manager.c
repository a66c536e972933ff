/*
 * pa_block_manager.c -- host side, plain C: the page allocator + per-sequence block table + LRU
 * behind the reference's block_manager.c interface (block_manager.c:25-201), with run-time
 * geometry and pages that are offsets into ONE device pool instead of a malloc per page.
 *
 * Observable integer behaviour is the reference's, bit for bit (tests/test_block_manager.py
 * drives this file, the compiled reference and the oracle through the same traces):
 *   - first-fit allocation of the lowest free page index          (block_manager.c:121-128)
 *   - on exhaustion the WHOLE prompt owning the page with the smallest lru_counter strictly
 *     below lru_epoch is evicted                                   (:92-113,130-142)
 *   - lru_counter = ++lru_epoch on allocation                      (:153-155)
 *   - free does not reset lru_counter                              (:78-90)
 * Deliberate differences (documented in DESIGN.md): state is zero-initialised (the reference
 * leaves lru_epoch/filled/lru_counter as malloc garbage, :38-52); get_current_block is silent
 * (the reference printf's twice per call, :67,:73); an optional per-sequence page cap.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_internal.h"

BlockManager* pa_bm_create(pa_handle* owner, int channels, int block_size, int max_blocks,
                           int max_prompts, int table_stride) {
    BlockManager* m = (BlockManager*)calloc(1, sizeof(BlockManager));
    if (!m) return NULL;
    m->C = channels;
    m->block_size = block_size;
    m->max_blocks = max_blocks;
    m->max_prompts = max_prompts;
    m->table_stride = table_stride;
    m->pa = owner;
    m->blocks = (KVBlock*)calloc((size_t)max_blocks, sizeof(KVBlock));
    m->block_table = (int*)calloc((size_t)max_prompts * table_stride, sizeof(int));
    m->prompt_block_list = (int**)calloc((size_t)max_prompts, sizeof(int*));
    m->prompt_block_count = (int*)calloc((size_t)max_prompts, sizeof(int));
    m->refcount = (int*)calloc((size_t)max_blocks, sizeof(int));
    m->pinned = (unsigned char*)calloc((size_t)max_prompts, 1);
    if (!m->blocks || !m->block_table || !m->prompt_block_list || !m->prompt_block_count || !m->refcount || !m->pinned) {
        pa_bm_destroy(m);
        return NULL;
    }
    for (int p = 0; p < max_prompts; p++) m->prompt_block_list[p] = m->block_table + (size_t)p * table_stride;
    for (int i = 0; i < max_blocks; i++) m->blocks[i].prompt_id = -1;
    return m;
}

void pa_bm_destroy(BlockManager* m) {
    if (!m) return;
    pa_share_destroy(m);
    free(m->refcount);
    free(m->pinned);
    free(m->blocks);
    free(m->block_table);
    free(m->prompt_block_list);
    free(m->prompt_block_count);
    free(m);
}

static int valid_prompt(const BlockManager* m, int p) { return p >= 0 && p < m->max_prompts; }

/* page i of layer 0 inside the pool (NULL pool on a host-only handle) */
static void bind_page(BlockManager* m, int idx) {
    pa_handle* h = m->pa;
    size_t off = (size_t)idx * m->block_size * m->C;
    m->blocks[idx].keys = (h && h->pool_k) ? h->pool_k + off : NULL;
    m->blocks[idx].values = (h && h->pool_v) ? h->pool_v + off : NULL;
}

void print_state(BlockManager* m, int prompt) {
    if (!m || !valid_prompt(m, prompt)) return;
    printf("Block manager llru %d\n", m->lru_epoch);
    int n = m->prompt_block_count[prompt];
    printf("Prompt %d block count: %d\n", prompt, n);
    for (int i = 0; i < n; i++) {
        int id = m->prompt_block_list[prompt][i];
        printf("Block %d: filled %d, llru %d\n", id, m->blocks[id].filled, m->blocks[id].lru_counter);
    }
}

int get_next_block_id(BlockManager* m, int prompt, int block_id) {
    if (!m || !valid_prompt(m, prompt)) return -1;
    const int* row = m->prompt_block_list[prompt];
    /* the reference scans the whole row, stale entries included (block_manager.c:57-61) */
    for (int i = 0; i + 1 < m->table_stride; i++)
        if (row[i] == block_id) return row[i + 1];
    return -1;
}

KVBlock* get_current_block(BlockManager* m, int prompt_id) {
    if (!m || !valid_prompt(m, prompt_id)) return NULL;
    int n = m->prompt_block_count[prompt_id];
    if (n == 0) return NULL;
    return &m->blocks[m->prompt_block_list[prompt_id][n - 1]];
}

/* Prompt p lets go of page idx.  Without sharing (refcount 1) this is the reference's free
 * (block_manager.c:82-87); a shared page (pa_seq_fork / pa_prefix_*) stays with its other holders. */
void pa_bm_release_page(BlockManager* m, int p, int idx) {
    KVBlock* b = &m->blocks[idx];
    if (m->refcount[idx] > 1) {
        m->refcount[idx]--;
        if (b->prompt_id == p) b->prompt_id = pa_share_other_holder(m, p, idx);   /* keep a valid owner for the LRU */
        return;
    }
    m->refcount[idx] = 0;
    b->keys = NULL;
    b->values = NULL;
    b->filled = 0;
    b->prompt_id = -1;          /* lru_counter keeps its value, as in the reference */
}

void free_blocks_for_prompt(BlockManager* m, int prompt_id) {
    if (!m || !valid_prompt(m, prompt_id)) return;
    int n = m->prompt_block_count[prompt_id];
    m->prompt_block_count[prompt_id] = 0;      /* first: the holder search must not find this prompt */
    for (int i = 0; i < n; i++) pa_bm_release_page(m, prompt_id, m->prompt_block_list[prompt_id][i]);
}

int find_least_recently_used_block(BlockManager* m) {
    int victim = -1;
    int lowest = m->lru_epoch;
    for (int i = 0; i < m->max_blocks; i++) {
        const KVBlock* b = &m->blocks[i];
        /* (pages held only by the prefix cache are not a prompt's; a sequence of the step being built is pinned) */
        if (b->prompt_id >= 0 && b->lru_counter < lowest && !m->pinned[b->prompt_id]) {
            lowest = b->lru_counter;
            victim = i;
        }
    }
    return victim;
}

void page_out_lru_block(BlockManager* m) {
    int victim = find_least_recently_used_block(m);
    if (victim != -1) free_blocks_for_prompt(m, m->blocks[victim].prompt_id);
}

static int lowest_free_page(const BlockManager* m) {
    for (int i = 0; i < m->max_blocks; i++)
        if (m->blocks[i].prompt_id == -1) return i;
    return -1;
}

KVBlock* request_block(BlockManager* m, int prompt_id) {
    if (!m || !valid_prompt(m, prompt_id)) {
        fprintf(stderr, "Invalid prompt ID.\n");
        return NULL;
    }
    if (m->prompt_block_count[prompt_id] >= m->table_stride) {   /* extension: per-sequence cap */
        fprintf(stderr, "No blocks available.\n");
        return NULL;
    }
    int idx = lowest_free_page(m);
    if (idx == -1 && pa_share_evict_one_cached(m)) idx = lowest_free_page(m);   /* extension: cached prefixes go first */
    if (idx == -1) {
        /* the reference pages out ONE prompt (block_manager.c:130-133), which always frees a page;
         * with shared pages (extension) a victim may free nothing, so keep going until one does */
        for (int tries = 0; idx == -1 && tries < m->max_prompts; tries++) {
            int victim = find_least_recently_used_block(m);
            if (victim == -1) break;
            if (m->pa && m->pa->swap_enabled) pa_swap_on_evict(m->pa, m->blocks[victim].prompt_id);   /* extension */
            free_blocks_for_prompt(m, m->blocks[victim].prompt_id);
            idx = lowest_free_page(m);
        }
        if (idx == -1) {
            fprintf(stderr, "No blocks available.\n");
            return NULL;
        }
    }
    KVBlock* b = &m->blocks[idx];
    bind_page(m, idx);
    b->prompt_id = prompt_id;
    b->filled = 0;
    m->refcount[idx] = 1;
    b->lru_counter = ++m->lru_epoch;
    /* note: if the eviction above hit prompt_id itself its count is 0 again here, exactly as
     * in the reference, which re-reads the count after paging out (block_manager.c:157) */
    m->prompt_block_list[prompt_id][m->prompt_block_count[prompt_id]++] = idx;
    return b;
}

float*** collect_kv_blocks(BlockManager* m, int prompt_id, int* num_blocks) {
    if (!m || !valid_prompt(m, prompt_id)) {
        fprintf(stderr, "Invalid prompt ID.\n");
        return NULL;
    }
    int n = m->prompt_block_count[prompt_id];
    *num_blocks = n;
    if (n == 0) return NULL;
    float*** kv = (float***)malloc(2 * sizeof(float**));
    if (!kv) { fprintf(stderr, "Memory allocation failed for kv_pointers.\n"); return NULL; }
    kv[0] = (float**)malloc((size_t)n * sizeof(float*));
    kv[1] = (float**)malloc((size_t)n * sizeof(float*));
    if (!kv[0] || !kv[1]) {
        fprintf(stderr, "Memory allocation failed for key/value pointers.\n");
        free(kv[0]); free(kv[1]); free(kv);
        return NULL;
    }
    for (int i = 0; i < n; i++) {
        const KVBlock* b = &m->blocks[m->prompt_block_list[prompt_id][i]];
        kv[0][i] = b->keys;
        kv[1][i] = b->values;
    }
    return kv;
}

/* ---- helpers shared with pa_step.c / pa_compat.c ------------------------------------------- */
int pa_bm_choose_page(BlockManager* m, int prompt_id) {
    KVBlock* cur = get_current_block(m, prompt_id);
    if (cur) {
        if (cur->filled >= m->block_size) cur = request_block(m, prompt_id);
        else cur->lru_counter = ++m->lru_epoch;
    } else {
        cur = request_block(m, prompt_id);
    }
    return cur ? (int)(cur - m->blocks) : -1;
}

/* Cached tokens of a sequence.  attention_paged addresses token g at page g/BLOCK_SIZE
 * (paged_infer.c:190), i.e. every page but the last is full. */
int pa_bm_context_len(const BlockManager* m, int prompt_id) {
    int n = m->prompt_block_count[prompt_id];
    if (n == 0) return 0;
    return (n - 1) * m->block_size + m->blocks[m->prompt_block_list[prompt_id][n - 1]].filled;
}
