/*
 * ref_wrap.c -- TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * This text is appended (by oracle/build_ref.sh) to the *unmodified* reference
 * sources block_manager.c + paged_infer.c as they lie under /root/reference; the
 * only edit the recipe applies is a sed on the three geometry #defines
 * (block_manager.c:4-6, paged_infer.c:16-18), which cannot be overridden with -D.
 * Everything below only CALLS the reference functions; it restates none of them.
 * The translation unit is built with -fvisibility=hidden, so only the ref_*
 * entry points below are exported and nothing collides with libpaged_attn.so,
 * which exports the reference's own names.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load the resulting oracle/_ref/libref_*.so.
 */
#include <fcntl.h>
#include <time.h>

#define REF_API __attribute__((visibility("default")))

/* ---- geometry / environment ------------------------------------------------ */
REF_API void ref_geometry(int* block_size, int* max_blocks, int* max_prompts) {
    *block_size = BLOCK_SIZE; *max_blocks = MAX_BLOCKS; *max_prompts = MAX_PROMPTS;
}
REF_API int ref_omp_threads(void) {
#ifdef OMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
REF_API unsigned long ref_sizeof_manager(void) { return sizeof(BlockManager); }

/* The reference's hot functions printf per call/token (block_manager.c:67,73;
 * paged_infer.c:519-572).  Tests silence fd 1 around them instead of editing them. */
static int g_saved_stdout = -1;
REF_API void ref_silence(int on) {
    fflush(stdout);
    if (on && g_saved_stdout < 0) {
        int devnull = open("/dev/null", O_WRONLY);
        g_saved_stdout = dup(1);
        dup2(devnull, 1);
        close(devnull);
    } else if (!on && g_saved_stdout >= 0) {
        dup2(g_saved_stdout, 1);
        close(g_saved_stdout);
        g_saved_stdout = -1;
    }
}

/* ---- block manager (block_manager.c:25-201) -------------------------------- */
REF_API BlockManager* ref_create(int channels) {
    BlockManager* m = create_block_manager(channels);
    /* create_block_manager (block_manager.c:38-52) leaves lru_epoch, filled and
     * lru_counter uninitialised (malloc).  Zero them so traces are deterministic. */
    m->lru_epoch = 0;
    for (int i = 0; i < MAX_BLOCKS; i++) { m->blocks[i].filled = 0; m->blocks[i].lru_counter = 0; }
    /* prompt_block_list is malloc garbage too, and get_next_block_id (:54-63) scans all of it */
    memset(m->prompt_block_list, 0, sizeof(m->prompt_block_list));
    return m;
}
REF_API void ref_destroy(BlockManager* m) {
    for (int p = 0; p < MAX_PROMPTS; p++) free_blocks_for_prompt(m, p);
    free(m);
}
static int ref_index_of(BlockManager* m, KVBlock* b) { return b ? (int)(b - m->blocks) : -1; }
REF_API int ref_request_block(BlockManager* m, int p) { return ref_index_of(m, request_block(m, p)); }
REF_API int ref_get_current_block(BlockManager* m, int p) { return ref_index_of(m, get_current_block(m, p)); }
REF_API void ref_free_blocks_for_prompt(BlockManager* m, int p) { free_blocks_for_prompt(m, p); }
REF_API int ref_find_lru(BlockManager* m) { return find_least_recently_used_block(m); }
REF_API void ref_page_out_lru(BlockManager* m) { page_out_lru_block(m); }
REF_API int ref_get_next_block_id(BlockManager* m, int p, int id) { return get_next_block_id(m, p, id); }
REF_API void ref_print_state(BlockManager* m, int p) { print_state(m, p); }
REF_API int ref_lru_epoch(BlockManager* m) { return m->lru_epoch; }
REF_API int ref_block_count(BlockManager* m, int p) { return m->prompt_block_count[p]; }
REF_API int ref_block_table(BlockManager* m, int p, int* out, int cap) {
    int n = m->prompt_block_count[p];
    for (int i = 0; i < n && i < cap; i++) out[i] = m->prompt_block_list[p][i];
    return n;
}
REF_API void ref_block_info(BlockManager* m, int idx, int* filled, int* prompt_id, int* lru_counter) {
    *filled = m->blocks[idx].filled; *prompt_id = m->blocks[idx].prompt_id; *lru_counter = m->blocks[idx].lru_counter;
}
REF_API void ref_block_ptrs(BlockManager* m, int idx, float** keys, float** values) {
    *keys = m->blocks[idx].keys; *values = m->blocks[idx].values;
}
/* field pokes a caller of the reference does directly (paged_infer.c:524,570;
 * block_manager_test.c:16,24,30) */
REF_API void ref_touch(BlockManager* m, int idx) { m->blocks[idx].lru_counter = ++m->lru_epoch; }
REF_API void ref_set_filled(BlockManager* m, int idx, int filled) { m->blocks[idx].filled = filled; }

/* ---- KV append + paged attention (paged_infer.c:505-573, 163-240) ---------- */
REF_API void ref_add_to_cache(BlockManager* m, float* qkv, int B, int T, int C, int n_tail) {
    add_to_cache(m, qkv, B, T, C, n_tail);
}
REF_API void ref_attention_paged(float* out, float* preatt, float* att, float* inp,
                                 float** key_blocks, float** value_blocks,
                                 int B, int T, int C, int NH, int offset) {
    attention_paged(out, preatt, att, inp, key_blocks, value_blocks, B, T, C, NH, offset);
}
/* call-site order of paged_infer.c:713-715: collect_kv_blocks -> attention_paged.
 * preatt/att may be NULL, then scratch is allocated here.  Returns -1 if the
 * prompt has no blocks (collect_kv_blocks returns NULL, block_manager.c:172-174). */
REF_API int ref_attend_prompt(BlockManager* m, int prompt, float* out, float* preatt, float* att,
                              float* inp, int B, int T, int C, int NH, int offset) {
    int nb = 0;
    float*** kv = collect_kv_blocks(m, prompt, &nb);
    if (!kv) return -1;
    int own = 0;
    if (!preatt || !att) {
        own = 1;
        preatt = (float*)malloc((size_t)B * NH * T * T * sizeof(float));
        att = (float*)malloc((size_t)B * NH * T * T * sizeof(float));
    }
    attention_paged(out, preatt, att, inp, kv[0], kv[1], B, T, C, NH, offset);
    if (own) { free(preatt); free(att); }
    free(kv[0]); free(kv[1]); free(kv);   /* the reference leaks these (paged_infer.c:713) */
    return nb;
}
/* wall clock exactly as the reference takes it (paged_infer.c:1019-1020,1085-1087).  The
 * (B,NH,T,T) scratch is allocated once and reused, as gpt2_forward does (:598-637). */
static float* g_scratch = NULL;
static size_t g_scratch_floats = 0;
REF_API double ref_time_attend_prompt(BlockManager* m, int prompt, float* out, float* inp,
                                      int B, int T, int C, int NH, int offset, int reps) {
    size_t need = (size_t)B * NH * T * T;
    if (g_scratch_floats < 2 * need) {
        free(g_scratch);
        g_scratch = (float*)malloc(2 * need * sizeof(float));
        g_scratch_floats = 2 * need;
        memset(g_scratch, 0, 2 * need * sizeof(float));   /* fault the pages in outside the clock */
    }
    float* preatt = g_scratch;
    float* att = g_scratch + need;
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int nb = 0;
        float*** kv = collect_kv_blocks(m, prompt, &nb);
        if (!kv) { best = -1.0; break; }
        attention_paged(out, preatt, att, inp, kv[0], kv[1], B, T, C, NH, offset);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        free(kv[0]); free(kv[1]); free(kv);
        double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
        if (dt < best) best = dt;
    }
    return best;
}

/* ---- the step before the path (next row): paged_infer.c:92-160 - */
REF_API void ref_matmul_forward(float* out, float* inp, float* w, float* bias, int B, int T, int C, int OC) {
    matmul_forward(out, inp, w, bias, B, T, C, OC);
}
REF_API void ref_matmul_cached(float* out, float* inp, float* w, float* bias, int B, int T, int C, int OC) {
    matmul_cached(out, inp, w, bias, B, T, C, OC);
}

/* ---- the reference RNG (paged_infer.c:826-835) ------------------------------ */
/* the rest of the decode layer: paged_infer.c:24-89, :243-286, :838-848 */
REF_API void ref_encoder_forward(float* out, int* inp, float* wte, float* wpe, int B, int T, int C) {
    encoder_forward(out, inp, wte, wpe, B, T, C);
}
REF_API void ref_layernorm_forward(float* out, float* mean, float* rstd, float* inp, float* w, float* b, int B, int T, int C) {
    layernorm_forward(out, mean, rstd, inp, w, b, B, T, C);
}
REF_API void ref_gelu_forward(float* out, float* inp, int N) { gelu_forward(out, inp, N); }
REF_API void ref_residual_forward(float* out, float* a, float* b, int N) { residual_forward(out, a, b, N); }
REF_API void ref_softmax_forward(float* probs, float* logits, int B, int T, int V) { softmax_forward(probs, logits, B, T, V); }
REF_API int ref_sample_mult(float* probs, int n, float coin) { return sample_mult(probs, n, coin); }
/* on-disk formats: the reference's own readers */
REF_API long ref_checkpoint_load(const char* path, int* cfg5, float* params_out, long cap) {
    GPT2 model;
    gpt2_build_from_checkpoint(&model, path);                      /* prints the hyperparameters; exit(1) on a bad file */
    cfg5[0] = model.config.max_seq_len; cfg5[1] = model.config.vocab_size; cfg5[2] = model.config.num_layers;
    cfg5[3] = model.config.num_heads; cfg5[4] = model.config.channels;
    long n = (long)model.num_parameters;
    if (params_out && n <= cap) memcpy(params_out, model.params_memory, (size_t)n * sizeof(float));
    free(model.params_memory);
    return n;
}
REF_API void* ref_dataloader_open(const char* path, int B, int T) {
    DataLoader* d = (DataLoader*)malloc(sizeof(DataLoader));
    dataloader_init(d, path, B, T);
    return d;
}
REF_API int ref_dataloader_num_batches(void* d) { return ((DataLoader*)d)->num_batches; }
REF_API void ref_dataloader_next(void* dv, int* out) {
    DataLoader* d = (DataLoader*)dv;
    dataloader_next_batch(d);
    memcpy(out, d->batch, ((size_t)d->B * d->T + 1) * sizeof(int));
}
REF_API void ref_dataloader_reset(void* d) { dataloader_reset((DataLoader*)d); }
REF_API void ref_dataloader_free(void* d) { dataloader_free((DataLoader*)d); free(d); }
REF_API void* ref_tokenizer_open(const char* path) {
    Tokenizer* t = (Tokenizer*)malloc(sizeof(Tokenizer));
    tokenizer_init(t, path);
    return t;
}
REF_API int ref_tokenizer_vocab(void* t) { return ((Tokenizer*)t)->init_ok ? (int)((Tokenizer*)t)->vocab_size : -1; }
REF_API const char* ref_tokenizer_decode(void* t, unsigned id) { return tokenizer_decode((Tokenizer*)t, id); }
REF_API unsigned int ref_random_u32(unsigned long long* s) { return random_u32(s); }
REF_API float ref_random_f32(unsigned long long* s) { return random_f32(s); }
