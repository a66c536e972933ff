/*
 * paged_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * A CPU restatement, in plain C with run-time geometry, of the reference's paged
 * KV-cache path.  Every function cites the reference lines it follows
 * (/root/reference/...).  It exists so the CUDA path can be checked on machines
 * where /root/reference is absent and at geometries the reference's compile-time
 * macros (block_manager.c:4-6) do not cover.
 *
 * PINNING: tests/test_oracle_pinned.py checks this file against the reference's own
 * compiled code (oracle/_ref/libref_*.so, built by oracle/build_ref.sh from
 * /root/reference) -- integer state bit-exact over randomized allocator traces, and
 * (strict flavour) attention outputs bit-for-bit -- and against the committed golden
 * vectors in tests/golden/ that were generated from the compiled reference by
 * tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load liboracle_*.so.  libpaged_attn.so never links or calls it.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- page + manager state: block_manager.c:9-23 with run-time sizes --------- */
typedef struct {
    float* keys;        /* [bs][C] */
    float* values;      /* [bs][C] */
    int filled;
    int prompt_id;      /* -1 = free */
    int lru_counter;
} orc_page;

typedef struct {
    int C, bs, max_blocks, max_prompts;
    orc_page* pages;    /* [max_blocks] */
    int* table;         /* [max_prompts][max_blocks]  == prompt_block_list */
    int* count;         /* [max_prompts]              == prompt_block_count */
    int lru_epoch;
    int alloc_data;     /* 0: integer-only manager (no K/V storage) */
} orc_manager;

/* create_block_manager, block_manager.c:38-52 (plus the zeroing the reference omits) */
ORC_API orc_manager* orc_create(int C, int bs, int max_blocks, int max_prompts, int alloc_data) {
    orc_manager* m = (orc_manager*)calloc(1, sizeof(orc_manager));
    m->C = C; m->bs = bs; m->max_blocks = max_blocks; m->max_prompts = max_prompts;
    m->alloc_data = alloc_data;
    m->pages = (orc_page*)calloc((size_t)max_blocks, sizeof(orc_page));
    m->table = (int*)calloc((size_t)max_prompts * max_blocks, sizeof(int));
    m->count = (int*)calloc((size_t)max_prompts, sizeof(int));
    for (int i = 0; i < max_blocks; i++) m->pages[i].prompt_id = -1;
    return m;
}

/* free_blocks_for_prompt, block_manager.c:78-90 (lru_counter is NOT reset there) */
ORC_API void orc_free_blocks_for_prompt(orc_manager* m, int prompt) {
    int n = m->count[prompt];
    const int* row = m->table + (size_t)prompt * m->max_blocks;
    for (int i = 0; i < n; i++) {
        orc_page* pg = &m->pages[row[i]];
        free(pg->keys); free(pg->values);
        pg->keys = pg->values = NULL;
        pg->filled = 0;
        pg->prompt_id = -1;
    }
    m->count[prompt] = 0;
}

ORC_API void orc_destroy(orc_manager* m) {
    for (int p = 0; p < m->max_prompts; p++) orc_free_blocks_for_prompt(m, p);
    free(m->pages); free(m->table); free(m->count); free(m);
}

/* get_next_block_id, block_manager.c:54-63 */
ORC_API int orc_get_next_block_id(orc_manager* m, int prompt, int block_id) {
    const int* row = m->table + (size_t)prompt * m->max_blocks;
    for (int i = 0; i + 1 < m->max_blocks; i++)
        if (row[i] == block_id) return row[i + 1];
    return -1;
}

/* get_current_block, block_manager.c:65-76: last page of the prompt, or none */
ORC_API int orc_get_current_block(orc_manager* m, int prompt) {
    int n = m->count[prompt];
    return n == 0 ? -1 : m->table[(size_t)prompt * m->max_blocks + n - 1];
}

/* find_least_recently_used_block, block_manager.c:92-102: strict minimum below the epoch */
ORC_API int orc_find_lru(orc_manager* m) {
    int best = -1, best_counter = m->lru_epoch;
    for (int i = 0; i < m->max_blocks; i++) {
        if (m->pages[i].prompt_id == -1) continue;
        if (m->pages[i].lru_counter < best_counter) { best_counter = m->pages[i].lru_counter; best = i; }
    }
    return best;
}

/* page_out_lru_block, block_manager.c:104-113: the WHOLE owning prompt is evicted */
ORC_API void orc_page_out_lru(orc_manager* m) {
    int victim = orc_find_lru(m);
    if (victim != -1) orc_free_blocks_for_prompt(m, m->pages[victim].prompt_id);
}

static int orc_first_free(const orc_manager* m) {
    for (int i = 0; i < m->max_blocks; i++) if (m->pages[i].prompt_id == -1) return i;
    return -1;
}

/* request_block, block_manager.c:115-162.  Returns the page index or -1. */
ORC_API int orc_request_block(orc_manager* m, int prompt) {
    if (prompt < 0 || prompt >= m->max_prompts) return -1;           /* :116-119 */
    int idx = orc_first_free(m);                                     /* :121-128 */
    if (idx == -1) {                                                 /* :130-142 */
        orc_page_out_lru(m);
        idx = orc_first_free(m);
        if (idx == -1) return -1;
    }
    orc_page* pg = &m->pages[idx];
    if (m->alloc_data) {                                             /* :145-151 */
        pg->keys = (float*)malloc((size_t)m->bs * m->C * sizeof(float));
        pg->values = (float*)malloc((size_t)m->bs * m->C * sizeof(float));
    }
    pg->prompt_id = prompt;                                          /* :153-155 */
    pg->filled = 0;
    pg->lru_counter = ++m->lru_epoch;
    m->table[(size_t)prompt * m->max_blocks + m->count[prompt]] = idx;   /* :157-159 */
    m->count[prompt]++;
    return idx;
}

/* ---- accessors ---------------------------------------------------------------- */
ORC_API int orc_lru_epoch(orc_manager* m) { return m->lru_epoch; }
ORC_API int orc_block_count(orc_manager* m, int prompt) { return m->count[prompt]; }
ORC_API int orc_block_table(orc_manager* m, int prompt, int* out, int cap) {
    int n = m->count[prompt];
    for (int i = 0; i < n && i < cap; i++) out[i] = m->table[(size_t)prompt * m->max_blocks + i];
    return n;
}
ORC_API void orc_block_info(orc_manager* m, int idx, int* filled, int* prompt_id, int* lru_counter) {
    *filled = m->pages[idx].filled; *prompt_id = m->pages[idx].prompt_id; *lru_counter = m->pages[idx].lru_counter;
}
ORC_API void orc_block_ptrs(orc_manager* m, int idx, float** k, float** v) {
    *k = m->pages[idx].keys; *v = m->pages[idx].values;
}
ORC_API void orc_touch(orc_manager* m, int idx) { m->pages[idx].lru_counter = ++m->lru_epoch; }
ORC_API void orc_set_filled(orc_manager* m, int idx, int f) { m->pages[idx].filled = f; }
/* cached tokens of a prompt = sum of `filled` over its table */
ORC_API int orc_context_len(orc_manager* m, int prompt) {
    int n = 0;
    for (int i = 0; i < m->count[prompt]; i++) n += m->pages[m->table[(size_t)prompt * m->max_blocks + i]].filled;
    return n;
}
/* slot of logical position pos: table[pos/bs]*bs + pos%bs  (implied by paged_infer.c:190,548-566) */
ORC_API int orc_slot(orc_manager* m, int prompt, int pos) {
    return m->table[(size_t)prompt * m->max_blocks + pos / m->bs] * m->bs + pos % m->bs;
}

/* ---- KV append: add_to_cache, paged_infer.c:505-573 ---------------------------
 * The reference hard-codes prompt 0 (:515); `prompt` generalises it.  The page
 * choice (:518-529) is returned so callers can check the integer side on its own.
 * Returns the page index written, or -1 when no page could be had. */
ORC_API int orc_choose_page(orc_manager* m, int prompt) {
    int cur = orc_get_current_block(m, prompt);
    if (cur >= 0) {
        if (m->pages[cur].filled >= m->bs) cur = orc_request_block(m, prompt);     /* :520-522 */
        else m->pages[cur].lru_counter = ++m->lru_epoch;                           /* :524 */
    } else {
        cur = orc_request_block(m, prompt);                                        /* :528 */
    }
    return cur;
}
ORC_API int orc_add_to_cache(orc_manager* m, int prompt, const float* qkv, int B, int T, int C, int n_tail) {
    int cur = orc_choose_page(m, prompt);
    if (cur < 0) return -1;
    orc_page* pg = &m->pages[cur];
    if (m->alloc_data) {
        size_t base = (size_t)pg->filled * C;                                      /* :536 */
        for (int b = 0; b < B; b++) {                                              /* :548-566 */
            int pos = 0;
            for (int t = T - n_tail; t < T; t++, pos++) {
                const float* row = qkv + ((size_t)b * T + t) * 3 * C;
                memcpy(pg->keys + base + (size_t)pos * C, row + C, (size_t)C * sizeof(float));
                memcpy(pg->values + base + (size_t)pos * C, row + 2 * C, (size_t)C * sizeof(float));
            }
        }
    }
    pg->filled += n_tail;                                                          /* :570 */
    return cur;
}

/* ---- attention: attention_paged, paged_infer.c:163-240 -------------------------
 * One (query row, head): the four passes in the reference's evaluation order.
 * `scores`/`probs` hold at least nkeys floats.  Key/value row j of the window is
 * global token g = j + first_key, living in page g/bs at row g%bs (:190,:231). */
static void orc_attend_row(float* out_h, float* scores, float* probs, const float* q,
                           float* const* kpages, float* const* vpages,
                           int first_key, int nkeys, int bs, int C, int col0, int hs, float scale) {
    float mx = -10000.0f;                                                          /* :187 */
    for (int j = 0; j < nkeys; j++) {                                              /* :188-203 */
        int g = j + first_key;
        const float* k = kpages[g / bs] + (size_t)(g % bs) * C + col0;
        float dot = 0.0f;
        for (int i = 0; i < hs; i++) dot += q[i] * k[i];
        dot *= scale;
        if (dot > mx) mx = dot;
        scores[j] = dot;
    }
    float denom = 0.0f;                                                            /* :207-212 */
    for (int j = 0; j < nkeys; j++) {
        float e = expf(scores[j] - mx);
        denom += e;
        probs[j] = e;
    }
    float inv = denom == 0.0f ? 0.0f : 1.0f / denom;                               /* :213 */
    for (int j = 0; j < nkeys; j++) probs[j] *= inv;                               /* :216-224 */
    for (int i = 0; i < hs; i++) out_h[i] = 0.0f;                                  /* :228 */
    for (int j = 0; j < nkeys; j++) {                                              /* :229-236 */
        int g = j + first_key;
        const float* v = vpages[g / bs] + (size_t)(g % bs) * C + col0;
        float w = probs[j];
        for (int i = 0; i < hs; i++) out_h[i] += w * v[i];
    }
}

/* Full window, same signature as the reference plus bs.  preatt/att (B,NH,T,T)
 * receive the same side outputs (scores; normalised probabilities, zeros above
 * the diagonal, :216-224). */
ORC_API void orc_attention_paged(float* out, float* preatt, float* att, const float* inp,
                                 float* const* key_blocks, float* const* value_blocks,
                                 int B, int T, int C, int NH, int offset, int bs) {
    int hs = C / NH;
    float scale = 1.0 / sqrtf(hs);                                                 /* :174, double divide */
    #pragma omp parallel for collapse(3)
    for (int b = 0; b < B; b++)
        for (int t = 0; t < T; t++)
            for (int h = 0; h < NH; h++) {
                const float* q = inp + ((size_t)b * T + t) * 3 * C + h * hs;
                float* sc = preatt + (((size_t)b * NH + h) * T + t) * T;
                float* pr = att + (((size_t)b * NH + h) * T + t) * T;
                orc_attend_row(out + ((size_t)b * T + t) * C + h * hs, sc, pr, q,
                               key_blocks, value_blocks, offset, t + 1, bs, C, h * hs, hs, scale);
                for (int j = t + 1; j < T; j++) pr[j] = 0.0f;
            }
}

/* Decode = row t=T-1 of the window: one query per sequence over its
 * cached tokens [kv_start, ctx).  q: (nseq, q_stride) with head h at h*hs; out:
 * (nseq, out_stride).  Sequences are prompts seq_ids[i] of one manager. */
ORC_API int orc_decode_batch(orc_manager* m, const int* seq_ids, const int* kv_start, int nseq, int NH,
                             const float* q, int q_stride, float* out, int out_stride) {
    int C = m->C, bs = m->bs, hs = C / NH;
    float scale = 1.0 / sqrtf(hs);
    int rc = 0;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nseq; i++) {
        int p = seq_ids[i];
        int nb = m->count[p];
        int ctx = orc_context_len(m, p);
        int first = kv_start ? kv_start[i] : 0;
        int nkeys = ctx - first;
        if (nb == 0 || nkeys <= 0) {
            for (int c = 0; c < C; c++) out[(size_t)i * out_stride + c] = 0.0f;
            continue;
        }
        float** kp = (float**)malloc(sizeof(float*) * nb);
        float** vp = (float**)malloc(sizeof(float*) * nb);
        for (int j = 0; j < nb; j++) {
            kp[j] = m->pages[m->table[(size_t)p * m->max_blocks + j]].keys;
            vp[j] = m->pages[m->table[(size_t)p * m->max_blocks + j]].values;
        }
        float* sc = (float*)malloc(sizeof(float) * 2 * (size_t)nkeys);
        for (int h = 0; h < NH; h++)
            orc_attend_row(out + (size_t)i * out_stride + h * hs, sc, sc + nkeys,
                           q + (size_t)i * q_stride + h * hs, kp, vp, first, nkeys, bs, C, h * hs, hs, scale);
        free(sc); free(kp); free(vp);
    }
    return rc;
}

/* fp64 tie-breaker "truth" for long contexts (parity arithmetic):
 * same mathematics, double accumulation, exp in double. */
ORC_API int orc_decode_batch_f64(orc_manager* m, const int* seq_ids, const int* kv_start, int nseq, int NH,
                                 const float* q, int q_stride, double* out, int out_stride) {
    int C = m->C, bs = m->bs, hs = C / NH;
    double scale = (double)(float)(1.0 / sqrtf(hs));
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nseq; i++) {
        int p = seq_ids[i];
        int ctx = orc_context_len(m, p);
        int first = kv_start ? kv_start[i] : 0;
        int nkeys = ctx - first;
        for (int c = 0; c < C; c++) out[(size_t)i * out_stride + c] = 0.0;
        if (m->count[p] == 0 || nkeys <= 0) continue;
        double* sc = (double*)malloc(sizeof(double) * (size_t)nkeys);
        for (int h = 0; h < NH; h++) {
            const float* qh = q + (size_t)i * q_stride + h * hs;
            double mx = -10000.0;
            for (int j = 0; j < nkeys; j++) {
                int g = j + first;
                const float* k = m->pages[m->table[(size_t)p * m->max_blocks + g / bs]].keys + (size_t)(g % bs) * C + h * hs;
                double dot = 0.0;
                for (int d = 0; d < hs; d++) dot += (double)qh[d] * (double)k[d];
                dot *= scale;
                if (dot > mx) mx = dot;
                sc[j] = dot;
            }
            double denom = 0.0;
            for (int j = 0; j < nkeys; j++) { sc[j] = exp(sc[j] - mx); denom += sc[j]; }
            double inv = denom == 0.0 ? 0.0 : 1.0 / denom;
            double* o = out + (size_t)i * out_stride + h * hs;
            for (int j = 0; j < nkeys; j++) {
                int g = j + first;
                const float* v = m->pages[m->table[(size_t)p * m->max_blocks + g / bs]].values + (size_t)(g % bs) * C + h * hs;
                double w = sc[j] * inv;
                for (int d = 0; d < hs; d++) o[d] += w * (double)v[d];
            }
        }
        free(sc);
    }
    return 0;
}

/* General causal rows (prefill / chunked prefill): sequence i has n_q[i] query rows
 * packed in q at row q_row0[i]; row j sees cached tokens [kv_start, base_len[i]+j).
 * The reference's full window is kv_start=offset, base_len=offset+1, n_q=T. */
ORC_API int orc_attend_rows(orc_manager* m, const int* seq_ids, const int* kv_start, const int* base_len,
                            const int* n_q, const int* q_row0, int nseq, int NH,
                            const float* q, int q_stride, float* out, int out_stride) {
    int C = m->C, bs = m->bs, hs = C / NH;
    float scale = 1.0 / sqrtf(hs);
    for (int i = 0; i < nseq; i++) {
        int p = seq_ids[i];
        int nb = m->count[p];
        float** kp = (float**)malloc(sizeof(float*) * (nb ? nb : 1));
        float** vp = (float**)malloc(sizeof(float*) * (nb ? nb : 1));
        for (int j = 0; j < nb; j++) {
            kp[j] = m->pages[m->table[(size_t)p * m->max_blocks + j]].keys;
            vp[j] = m->pages[m->table[(size_t)p * m->max_blocks + j]].values;
        }
        int first = kv_start ? kv_start[i] : 0;
        #pragma omp parallel for collapse(2) schedule(dynamic, 1)
        for (int j = 0; j < n_q[i]; j++)
            for (int h = 0; h < NH; h++) {
                int nkeys = base_len[i] + j - first;
                size_t row = (size_t)q_row0[i] + j;
                float* o = out + row * out_stride + h * hs;
                if (nkeys <= 0 || nb == 0) { for (int d = 0; d < hs; d++) o[d] = 0.0f; continue; }
                float* sc = (float*)malloc(sizeof(float) * 2 * (size_t)nkeys);
                orc_attend_row(o, sc, sc + nkeys, q + row * q_stride + h * hs, kp, vp, first, nkeys, bs, C, h * hs, hs, scale);
                free(sc);
            }
        free(kp); free(vp);
    }
    return 0;
}

/* ---- the step before the path: matmul_forward / matmul_cached,
 * paged_infer.c:92-114 and :117-160 ------------------------------------------------ */
static float orc_dot_bias(const float* x, const float* w, float bias, int n) {
    float acc = bias;
    for (int i = 0; i < n; i++) acc += x[i] * w[i];
    return acc;
}
ORC_API void orc_matmul_forward(float* out, const float* inp, const float* weight, const float* bias,
                                int B, int T, int C, int OC) {
    #pragma omp parallel for
    for (int r = 0; r < B * T; r++)
        for (int o = 0; o < OC; o++)
            out[(size_t)r * OC + o] = orc_dot_bias(inp + (size_t)r * C, weight + (size_t)o * C, bias ? bias[o] : 0.0f, C);
}
/* Q for every window row, K and V for the last row only (:117-160) */
ORC_API void orc_matmul_cached(float* out, const float* inp, const float* weight, const float* bias,
                               int B, int T, int C, int OC) {
    #pragma omp parallel for
    for (int b = 0; b < B; b++) {
        for (int t = 0; t < T; t++)
            for (int o = 0; o < C; o++)
                out[((size_t)b * T + t) * OC + o] =
                    orc_dot_bias(inp + ((size_t)b * T + t) * C, weight + (size_t)o * C, bias ? bias[o] : 0.0f, C);
        const float* x = inp + ((size_t)b * T + T - 1) * C;
        float* dst = out + ((size_t)b * T + T - 1) * OC;
        for (int o = C; o < 3 * C; o++)
            dst[o] = orc_dot_bias(x, weight + (size_t)o * C, bias ? bias[o] : 0.0f, C);
    }
}

/* ---- the rest of the decode layer ------------------------------------------------
 * encoder_forward :24-46, layernorm_forward :49-89, gelu_forward :243-251, residual_forward
 * :253-257, softmax_forward :259-286, sample_mult :838-848 (all paged_infer.c) */
ORC_API void orc_encoder_forward(float* out, const int* inp, const float* wte, const float* wpe, int B, int T, int C) {
    for (int b = 0; b < B; b++)
        for (int t = 0; t < T; t++) {
            const float* e = wte + (size_t)inp[b * T + t] * C;
            const float* ps = wpe + (size_t)t * C;
            float* o = out + ((size_t)b * T + t) * C;
            for (int i = 0; i < C; i++) o[i] = e[i] + ps[i];
        }
}
ORC_API void orc_layernorm_forward(float* out, float* mean, float* rstd, const float* inp, const float* weight,
                                   const float* bias, int B, int T, int C) {
    const float eps = 1e-5f;                                   /* :56 */
    for (int r = 0; r < B * T; r++) {
        const float* x = inp + (size_t)r * C;
        float m = 0.0f;
        for (int i = 0; i < C; i++) m += x[i];
        m = m / C;
        float v = 0.0f;
        for (int i = 0; i < C; i++) { float d = x[i] - m; v += d * d; }
        v = v / C;
        float s = 1.0f / sqrtf(v + eps);
        float* o = out + (size_t)r * C;
        for (int i = 0; i < C; i++) o[i] = (s * (x[i] - m)) * weight[i] + bias[i];
        if (mean) mean[r] = m;
        if (rstd) rstd[r] = s;
    }
}
ORC_API void orc_gelu_forward(float* out, const float* inp, int N) {
    const float k = sqrtf(2.0f / M_PI);                        /* GELU_SCALING_FACTOR :243 */
    for (int i = 0; i < N; i++) {
        float x = inp[i];
        float cube = 0.044715f * x * x * x;
        out[i] = 0.5f * x * (1.0f + tanhf(k * (x + cube)));
    }
}
ORC_API void orc_residual_forward(float* out, const float* a, const float* b, int N) {
    for (int i = 0; i < N; i++) out[i] = a[i] + b[i];
}
ORC_API void orc_softmax_forward(float* probs, const float* logits, int B, int T, int V) {
    #pragma omp parallel for
    for (int r = 0; r < B * T; r++) {
        const float* l = logits + (size_t)r * V;
        float* p = probs + (size_t)r * V;
        float maxval = -10000.0f;                              /* :270 */
        for (int i = 0; i < V; i++) if (l[i] > maxval) maxval = l[i];
        float sum = 0.0f;
        for (int i = 0; i < V; i++) { p[i] = expf(l[i] - maxval); sum += p[i]; }
        for (int i = 0; i < V; i++) p[i] /= sum;
    }
}
ORC_API int orc_sample_mult(const float* probabilities, int n, float coin) {
    float cdf = 0.0f;
    for (int i = 0; i < n; i++) {
        cdf += probabilities[i];
        if (coin < cdf) return i;
    }
    return n - 1;
}

/* One decode step of the whole model for a batch of sequences, the reference's gpt2_forward
 * (:646-728) with T = 1 rows, the layer loop run over all L layers (the fork stops at l < 1, :659)
 * and one KV manager per layer.  params: the checkpoint's 16 tensors in file order (:441-488).
 * Per layer: ln1 -> matmul (QKV) -> add_to_cache -> last-row attention -> attproj -> residual ->
 * ln2 -> fc -> gelu -> fcproj -> residual; then lnf -> logits (wte^T) .  Returns logits (nseq, V). */
ORC_API int orc_model_decode_step(orc_manager** layer_mgrs, int L, int NH, int C, int V, int maxT,
                                  const float* params, const int* seq_ids, const int* tokens, const int* positions,
                                  int nseq, float* logits) {
    const float* wte = params;
    const float* wpe = wte + (size_t)V * C;
    const float* ln1w = wpe + (size_t)maxT * C;
    const float* ln1b = ln1w + (size_t)L * C;
    const float* qkvw = ln1b + (size_t)L * C;
    const float* qkvb = qkvw + (size_t)L * 3 * C * C;
    const float* attprojw = qkvb + (size_t)L * 3 * C;
    const float* attprojb = attprojw + (size_t)L * C * C;
    const float* ln2w = attprojb + (size_t)L * C;
    const float* ln2b = ln2w + (size_t)L * C;
    const float* fcw = ln2b + (size_t)L * C;
    const float* fcb = fcw + (size_t)L * 4 * C * C;
    const float* fcprojw = fcb + (size_t)L * 4 * C;
    const float* fcprojb = fcprojw + (size_t)L * C * 4 * C;
    const float* lnfw = fcprojb + (size_t)L * C;
    const float* lnfb = lnfw + C;
    float* x = (float*)malloc((size_t)nseq * C * sizeof(float));
    float* ln = (float*)malloc((size_t)nseq * C * sizeof(float));
    float* qkv = (float*)malloc((size_t)nseq * 3 * C * sizeof(float));
    float* atty = (float*)malloc((size_t)nseq * C * sizeof(float));
    float* proj = (float*)malloc((size_t)nseq * C * sizeof(float));
    float* fch = (float*)malloc((size_t)nseq * 4 * C * sizeof(float));
    if (!x || !ln || !qkv || !atty || !proj || !fch) return -1;
    for (int s = 0; s < nseq; s++)                                  /* encoder_forward with the token's own position */
        for (int i = 0; i < C; i++) x[(size_t)s * C + i] = wte[(size_t)tokens[s] * C + i] + wpe[(size_t)positions[s] * C + i];
    for (int l = 0; l < L; l++) {
        orc_layernorm_forward(ln, NULL, NULL, x, ln1w + (size_t)l * C, ln1b + (size_t)l * C, nseq, 1, C);
        orc_matmul_forward(qkv, ln, qkvw + (size_t)l * 3 * C * C, qkvb + (size_t)l * 3 * C, nseq, 1, C, 3 * C);
        for (int s = 0; s < nseq; s++)
            if (orc_add_to_cache(layer_mgrs[l], seq_ids[s], qkv + (size_t)s * 3 * C, 1, 1, C, 1) < 0) return -2;
        if (orc_decode_batch(layer_mgrs[l], seq_ids, NULL, nseq, NH, qkv, 3 * C, atty, C) != 0) return -3;
        orc_matmul_forward(proj, atty, attprojw + (size_t)l * C * C, attprojb + (size_t)l * C, nseq, 1, C, C);
        orc_residual_forward(x, x, proj, nseq * C);
        orc_layernorm_forward(ln, NULL, NULL, x, ln2w + (size_t)l * C, ln2b + (size_t)l * C, nseq, 1, C);
        orc_matmul_forward(fch, ln, fcw + (size_t)l * 4 * C * C, fcb + (size_t)l * 4 * C, nseq, 1, C, 4 * C);
        orc_gelu_forward(fch, fch, nseq * 4 * C);
        orc_matmul_forward(proj, fch, fcprojw + (size_t)l * C * 4 * C, fcprojb + (size_t)l * C, nseq, 1, 4 * C, C);
        orc_residual_forward(x, x, proj, nseq * C);
    }
    orc_layernorm_forward(ln, NULL, NULL, x, lnfw, lnfb, nseq, 1, C);
    orc_matmul_forward(logits, ln, wte, NULL, nseq, 1, C, V);
    free(x); free(ln); free(qkv); free(atty); free(proj); free(fch);
    return 0;
}

/* ---- deterministic inputs: the reference's xorshift64* (paged_infer.c:826-835)
 * with Box-Muller on top --------------------------------------------- */
ORC_API unsigned int orc_random_u32(unsigned long long* state) {
    *state ^= *state >> 12;
    *state ^= *state << 25;
    *state ^= *state >> 27;
    return (unsigned int)((*state * 0x2545F4914F6CDD1Dull) >> 32);
}
ORC_API float orc_random_f32(unsigned long long* state) {
    return (orc_random_u32(state) >> 8) / 16777216.0f;
}
ORC_API void orc_fill_normal(float* dst, size_t n, unsigned long long seed) {
    unsigned long long s = seed;
    for (size_t i = 0; i < n; i += 2) {
        float u1 = orc_random_f32(&s), u2 = orc_random_f32(&s);
        if (u1 < 1e-7f) u1 = 1e-7f;
        float r = sqrtf(-2.0f * logf(u1));
        dst[i] = r * cosf(6.28318530717958647692f * u2);
        if (i + 1 < n) dst[i + 1] = r * sinf(6.28318530717958647692f * u2);
    }
}
ORC_API void orc_fill_uniform(float* dst, size_t n, float lo, float hi, unsigned long long seed) {
    unsigned long long s = seed;
    for (size_t i = 0; i < n; i++) dst[i] = lo + (hi - lo) * orc_random_f32(&s);
}

ORC_API int orc_omp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* wall clock as the reference takes it (paged_infer.c:1019-1020) */
ORC_API double orc_time_decode_batch(orc_manager* m, const int* seq_ids, int nseq, int NH,
                                     const float* q, int q_stride, float* out, int out_stride, int reps) {
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        orc_decode_batch(m, seq_ids, NULL, nseq, NH, q, q_stride, out, out_stride);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
        if (dt < best) best = dt;
    }
    return best;
}
