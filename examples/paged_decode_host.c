/*
 * examples/paged_decode_host.c -- a plain-C host (gcc, no CUDA headers) driving the paged
 * attention path through include/paged_attn.h, twice:
 *
 *   1. the reference's own call site, paged_infer.c:710-715, unchanged:
 *          add_to_cache(manager, l_qkv, B, T, C, n);
 *          kv_blocks = collect_kv_blocks(manager, 0, &num_blocks);
 *          attention_paged(l_atty, l_preatt, l_att, l_qkv, kv_blocks[0], kv_blocks[1], B, T, C, NH, offset);
 *      with main's sliding window (paged_infer.c:1055-1057: T=32, offset = t-T);
 *   2. the extended batch API: 8 sequences, 2 layers, one decode step per iteration.
 *
 * Build:  gcc -O2 -Iinclude examples/paged_decode_host.c -L llm.c-paged_b200 -lpaged_attn \
 *             -Wl,-rpath,$PWD/llm.c-paged_b200 -lm -o examples/paged_decode_host
 * Prints a checksum per part; tests/test_gpu_example.py compares them with the CPU oracle.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "paged_attn.h"

/* the reference's RNG (paged_infer.c:826-835), for reproducible inputs */
static unsigned int random_u32(unsigned long long* state) {
    *state ^= *state >> 12;
    *state ^= *state << 25;
    *state ^= *state >> 27;
    return (unsigned int)((*state * 0x2545F4914F6CDD1Dull) >> 32);
}
static float random_f32(unsigned long long* state) { return (random_u32(state) >> 8) / 16777216.0f; }

static double checksum(const float* x, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += (double)x[i] * (double)((i % 7) + 1);
    return s;
}

static int part1_reference_call_site(void) {
    const int B = 1, T = 32, C = 768, NH = 12, steps = 19;
    BlockManager* manager = create_block_manager(C);          /* block 32, 100 pages: the reference's macros */
    if (!manager) return 1;
    unsigned long long rng = 1337;
    float* stream = (float*)malloc((size_t)(T + steps) * 3 * C * sizeof(float));
    for (size_t i = 0; i < (size_t)(T + steps) * 3 * C; i++) stream[i] = random_f32(&rng) * 2.0f - 1.0f;
    float* l_atty = (float*)malloc((size_t)B * T * C * sizeof(float));
    double sum = 0.0;
    for (int step = 0; step < steps; step++) {
        float* l_qkv = stream + (size_t)step * 3 * C;          /* window slides by one token */
        int n = step == 0 ? T : 1;                             /* first_pass ? T : 1  (paged_infer.c:696-708) */
        int offset = step;
        add_to_cache(manager, l_qkv, B, T, C, n);
        int num_blocks;
        float*** kv_blocks = collect_kv_blocks(manager, 0, &num_blocks);
        attention_paged(l_atty, NULL, NULL, l_qkv, kv_blocks[0], kv_blocks[1], B, T, C, NH, offset);
        free(kv_blocks[0]); free(kv_blocks[1]); free(kv_blocks);
        sum += checksum(l_atty, (size_t)B * T * C);
    }
    printf("part1 blocks=%d filled0=%d filled1=%d lru_epoch=%d checksum=%.6f\n", manager->prompt_block_count[0],
           manager->blocks[manager->prompt_block_list[0][0]].filled,
           manager->blocks[manager->prompt_block_list[0][1]].filled, manager->lru_epoch, sum);
    free(stream); free(l_atty);
    destroy_block_manager(manager);
    return 0;
}

static int part2_batched_decode(void) {
    pa_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.block_size = 16; cfg.max_blocks = 256; cfg.max_seqs = 8; cfg.n_layers = 2;
    cfg.n_heads = 12; cfg.head_dim = 64; cfg.device = 0; cfg.max_batch_tokens = 64;
    const int C = 768, B = 8, L = 2;
    pa_handle* h = NULL;
    if (pa_create(&cfg, &h) != PA_OK) { fprintf(stderr, "pa_create: %s\n", pa_last_error()); return 1; }
    float* qkv = (float*)pa_host_alloc((size_t)B * 3 * C * sizeof(float));   /* pinned: the kernel reads it in place */
    float* out = (float*)pa_host_alloc((size_t)B * C * sizeof(float));
    int seq_ids[8], n_new[8];
    for (int i = 0; i < B; i++) { seq_ids[i] = i; n_new[i] = 1; }
    unsigned long long rng = 42;
    double sum = 0.0;
    for (int step = 0; step < 40; step++) {                    /* crosses two page boundaries */
        if (pa_step_begin(h, seq_ids, n_new, B) != PA_OK) { fprintf(stderr, "%s\n", pa_last_error()); return 1; }
        for (int layer = 0; layer < L; layer++) {
            for (int i = 0; i < B * 3 * C; i++) qkv[i] = random_f32(&rng) * 2.0f - 1.0f;
            if (pa_decode_step_host(h, layer, qkv, out) != PA_OK) { fprintf(stderr, "%s\n", pa_last_error()); return 1; }
            sum += checksum(out, (size_t)B * C);
        }
    }
    BlockManager* m = pa_manager(h);
    printf("part2 ctx=%d pages=%d table0=[%d %d %d] checksum=%.6f\n", pa_seq_len(h, 0), m->prompt_block_count[0],
           m->prompt_block_list[0][0], m->prompt_block_list[0][1], m->prompt_block_list[0][2], sum);
    pa_host_free(qkv); pa_host_free(out);
    pa_destroy(h);
    return 0;
}

int main(void) {
    if (pa_device_count() < 1) { fprintf(stderr, "no CUDA device: libpaged_attn has no CPU fallback\n"); return 2; }
    if (part1_reference_call_site()) return 1;
    if (part2_batched_decode()) return 1;
    return 0;
}
