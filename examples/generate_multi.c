/*
 * generate_multi.c -- plain C (gcc, no CUDA / NCCL headers): the reference's sampling loop
 * (paged_infer.c:1028-1063: prefill the prompt, then one token per step) for sequences SHARDED over the
 * GPUs of one box, ONE process driving all of them through a pa_group: every GPU has its own block
 * manager, page pool, tables and a replica of the weights; the only exchange is the all-gather of the
 * sampled tokens behind the sampler (NCCL, enqueued on each GPU's stream -- pa_group_model_step).
 *
 *   ./generate_multi [n_gpus [first_rank]]      (default: every visible GPU, first_rank 0)
 *
 * Rank r takes batch r of the token file as its prompts and draws its coins from its own xorshift64*
 * stream (seed 1337 + r), so what a rank generates does not depend on how many GPUs take part:
 * `generate_multi 1 1` (one GPU playing rank 1) prints the rank-1 lines of `generate_multi 2`, and rank 0
 * generates what examples/generate.c generates.  Tokens are printed FROM THE GATHERED BUFFER.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "paged_attn.h"

static unsigned int random_u32(unsigned long long* state) {      /* paged_infer.c:826-832 */
    *state ^= *state >> 12;
    *state ^= *state << 25;
    *state ^= *state >> 27;
    return (unsigned int)((*state * 0x2545F4914F6CDD1Dull) >> 32);
}
static float random_f32(unsigned long long* state) { return (random_u32(state) >> 8) / 16777216.0f; }

#define CHECK(call) do { int rc_ = (call); if (rc_ < 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, pa_last_error()); return 1; } } while (0)
#define MAX_GPUS 16
enum { B = 4, PROMPT_SIZE = 32, TOTAL = 50 };

static int write_synthetic(const char* ckpt, const char* toks) {      /* the files examples/generate.c writes */
    pa_model_config cfg = { .max_seq_len = 128, .vocab_size = 512, .n_layers = 2, .n_heads = 4, .channels = 256 };
    size_t n = pa_model_param_count(&cfg);
    float* p = (float*)malloc(n * sizeof(float));
    if (!p) return -1;
    unsigned long long s = 42;
    for (size_t i = 0; i < n; i++) p[i] = (random_f32(&s) - 0.5f) * 0.2f;
    int rc = pa_checkpoint_write(ckpt, &cfg, p);
    free(p);
    if (rc < 0) return rc;
    int ids[4096];
    for (int i = 0; i < 4096; i++) ids[i] = (int)(random_u32(&s) % 512);
    return pa_tokens_write(toks, ids, 4096);
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : pa_device_count();
    const int first_rank = argc > 2 ? atoi(argv[2]) : 0;
    if (n < 1 || n > MAX_GPUS) { fprintf(stderr, "n_gpus must be 1..%d (visible: %d)\n", MAX_GPUS, pa_device_count()); return 1; }
    const char* ckpt = "/tmp/pa_synth_gpt2_multi.bin";
    const char* toks = "/tmp/pa_synth_tokens_multi.bin";
    CHECK(write_synthetic(ckpt, toks));
    pa_model_config mc;
    CHECK(pa_checkpoint_read_config(ckpt, &mc));

    pa_config cfg = { .block_size = 16, .max_blocks = 64, .max_seqs = B, .max_blocks_per_seq = 0, .n_layers = mc.n_layers,
                      .n_heads = mc.n_heads, .head_dim = mc.channels / mc.n_heads, .device = 0, .max_batch_tokens = B * PROMPT_SIZE };
    pa_group* group;
    CHECK(pa_group_create(&cfg, n, NULL, &group));
    printf("group: %d GPU(s) in one process, NCCL %d, ranks %d..%d\n", pa_group_size(group), pa_nccl_version(), first_rank,
           first_rank + n - 1);

    pa_model* models[MAX_GPUS];
    static int prompts[MAX_GPUS][B * PROMPT_SIZE], next[MAX_GPUS][B], seq_ids[MAX_GPUS][B], all_next[MAX_GPUS * B];
    static float coins[MAX_GPUS][B];
    const int* seq_ptr[MAX_GPUS]; const int* tok_ptr[MAX_GPUS]; const float* coin_ptr[MAX_GPUS];
    unsigned long long rng[MAX_GPUS];
    pa_dataloader* loader;
    CHECK(pa_dataloader_open(toks, B, PROMPT_SIZE, &loader));
    for (int r = 0; r < first_rank; r++) { const int* skip; CHECK(pa_dataloader_next_batch(loader, &skip, NULL)); }
    for (int i = 0; i < n; i++) {
        CHECK(pa_model_create_from_checkpoint(pa_group_handle(group, i), ckpt, B * PROMPT_SIZE, &models[i]));
        const int* inputs;
        CHECK(pa_dataloader_next_batch(loader, &inputs, NULL));
        memcpy(prompts[i], inputs, sizeof(prompts[i]));
        rng[i] = 1337ull + (unsigned long long)(first_rank + i);
        for (int b = 0; b < B; b++) seq_ids[i][b] = b;
        seq_ptr[i] = seq_ids[i]; tok_ptr[i] = next[i]; coin_ptr[i] = coins[i];
    }
    static int gen[MAX_GPUS][B][TOTAL];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    /* prefill: every GPU's prompts in one step each, queued on all GPUs before any is waited for */
    int n_new[B];
    for (int b = 0; b < B; b++) n_new[b] = PROMPT_SIZE;
    for (int i = 0; i < n; i++) {
        for (int b = 0; b < B; b++) coins[i][b] = random_f32(&rng[i]);
        CHECK(pa_model_forward_async(models[i], seq_ids[i], n_new, prompts[i], coins[i], B));
    }
    for (int i = 0; i < n; i++) CHECK(pa_model_wait(models[i], next[i]));
    for (int i = 0; i < n; i++) for (int b = 0; b < B; b++) gen[i][b][PROMPT_SIZE] = next[i][b];      /* the prefill's token: local, no gather */
    /* decode: one group step per token.  A rank's next step needs only its OWN sampled tokens (next[i], handed back
     * after one wait); the all-gather of a step's tokens runs on a side stream beside the NEXT step and comes out of
     * the next call -- gen[][][t] for t > PROMPT_SIZE is filled FROM THE GATHERED BUFFER, one step late. */
    int* next_ptr[MAX_GPUS];
    for (int i = 0; i < n; i++) next_ptr[i] = next[i];
    for (int t = PROMPT_SIZE + 1; t < TOTAL; t++) {
        for (int i = 0; i < n; i++) for (int b = 0; b < B; b++) coins[i][b] = random_f32(&rng[i]);
        int have = pa_group_model_step_overlapped(group, models, seq_ptr, tok_ptr, coin_ptr, B, next_ptr, all_next);
        CHECK(have);
        if (have > 0)       /* the tokens of position t - 1, every rank's */
            for (int i = 0; i < n; i++) for (int b = 0; b < B; b++) gen[i][b][t - 1] = all_next[i * B + b];
    }
    CHECK(pa_group_gather_flush(group, all_next));
    for (int i = 0; i < n; i++) for (int b = 0; b < B; b++) gen[i][b][TOTAL - 1] = all_next[i * B + b];
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < n; i++)
        for (int b = 0; b < B; b++) {
            printf("rank %d sequence %d:", first_rank + i, b);
            for (int t = PROMPT_SIZE; t < TOTAL; t++) printf(" %d", gen[i][b][t]);
            printf("\n");
        }
    double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
    printf("generated %d tokens for %d sequences on %d GPU(s) in %.3f ms\n", TOTAL - PROMPT_SIZE, n * B, n, dt * 1e3);
    for (int i = 0; i < n; i++) pa_model_destroy(models[i]);
    pa_dataloader_close(loader);
    pa_group_destroy(group);
    return 0;
}
