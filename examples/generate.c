/*
 * generate.c -- plain C (gcc, no CUDA headers): the reference's `main` (paged_infer.c:953-1090)
 * rebuilt on libpaged_attn.so for a batch of sequences: read a checkpoint in the reference's
 * gpt2_124M.bin layout, take prompts from an int32 token stream, prefill them in one step, then
 * sample autoregressively with the reference's RNG (xorshift64*, :826-835) and sample_mult.
 * Without arguments it first writes a small synthetic checkpoint and token file, so it runs
 * offline:   ./generate [checkpoint.bin tokens.bin [tokenizer.bin]]
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

static int write_synthetic(const char* ckpt, const char* toks) {
    pa_model_config cfg = { .max_seq_len = 128, .vocab_size = 512, .n_layers = 2, .n_heads = 4, .channels = 256 };
    size_t n = pa_model_param_count(&cfg);
    float* p = (float*)malloc(n * sizeof(float));
    if (!p) return -1;
    unsigned long long s = 42;
    for (size_t i = 0; i < n; i++) p[i] = (random_f32(&s) - 0.5f) * 0.2f;
    /* layernorm weights around 1 (tensors 2, 8 and 14 of the file order) are not singled out here: a synthetic
     * model only has to be well conditioned, and weights of +-0.1 are */
    int rc = pa_checkpoint_write(ckpt, &cfg, p);
    free(p);
    if (rc < 0) return rc;
    int ids[4096];
    for (int i = 0; i < 4096; i++) ids[i] = (int)(random_u32(&s) % 512);
    return pa_tokens_write(toks, ids, 4096);
}

int main(int argc, char** argv) {
    const char* ckpt = argc > 2 ? argv[1] : "/tmp/pa_synth_gpt2.bin";
    const char* toks = argc > 2 ? argv[2] : "/tmp/pa_synth_tokens.bin";
    if (argc <= 2) CHECK(write_synthetic(ckpt, toks));

    pa_model_config mc;
    CHECK(pa_checkpoint_read_config(ckpt, &mc));
    printf("[GPT-2]\nmax_seq_len: %d\nvocab_size: %d\nnum_layers: %d\nnum_heads: %d\nchannels: %d\n", mc.max_seq_len,
           mc.vocab_size, mc.n_layers, mc.n_heads, mc.channels);

    const int B = 4, PROMPT_SIZE = 32, total = 50;              /* the reference: B = 1, PROMPT_SIZE = 32, totalSize = 50 */
    pa_config cfg = { .block_size = 16, .max_blocks = 64, .max_seqs = B, .max_blocks_per_seq = 0, .n_layers = mc.n_layers,
                      .n_heads = mc.n_heads, .head_dim = mc.channels / mc.n_heads, .device = 0, .max_batch_tokens = B * PROMPT_SIZE };
    pa_handle* h;
    CHECK(pa_create(&cfg, &h));
    pa_model* model;
    CHECK(pa_model_create_from_checkpoint(h, ckpt, B * PROMPT_SIZE, &model));

    pa_dataloader* loader;
    CHECK(pa_dataloader_open(toks, B, PROMPT_SIZE, &loader));
    printf("val dataset num_batches: %d\n", pa_dataloader_num_batches(loader));
    const int* inputs;
    CHECK(pa_dataloader_next_batch(loader, &inputs, NULL));
    pa_tokenizer* tok = NULL;
    if (argc > 3 && pa_tokenizer_open(argv[3], &tok) < 0) tok = NULL;

    int seq_ids[4] = {0, 1, 2, 3}, n_new[4], next[4];
    float coins[4];
    int* gen = (int*)malloc((size_t)B * total * sizeof(int));
    unsigned long long rng_state = 1337;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    /* first pass: the whole prompts in one step (the reference's first_pass with n = T) */
    for (int b = 0; b < B; b++) { n_new[b] = PROMPT_SIZE; coins[b] = random_f32(&rng_state); memcpy(gen + b * total, inputs + b * PROMPT_SIZE, PROMPT_SIZE * sizeof(int)); }
    CHECK(pa_model_forward(model, seq_ids, n_new, inputs, coins, B, next));
    for (int t = PROMPT_SIZE; t < total; t++) {
        for (int b = 0; b < B; b++) { gen[b * total + t] = next[b]; coins[b] = random_f32(&rng_state); }
        if (t + 1 < total) CHECK(pa_model_decode_step(model, seq_ids, next, coins, B, next));
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int b = 0; b < B; b++) {
        printf("sequence %d:", b);
        for (int t = PROMPT_SIZE; t < total; t++) {
            const char* piece = tok ? pa_tokenizer_decode(tok, (unsigned)gen[b * total + t]) : NULL;
            if (piece) printf("%s", piece); else printf(" %d", gen[b * total + t]);
        }
        printf("\n");
    }
    double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
    printf("generated %d tokens for %d sequences in %.3f ms; cached tokens per sequence: %d; block table of sequence 0:",
           total - PROMPT_SIZE, B, dt * 1e3, pa_seq_len(h, 0));
    BlockManager* m = pa_manager(h);
    for (int i = 0; i < m->prompt_block_count[0]; i++) printf(" %d", m->prompt_block_list[0][i]);
    printf("\n");
    free(gen);
    if (tok) pa_tokenizer_close(tok);
    pa_dataloader_close(loader);
    pa_model_destroy(model);
    pa_destroy(h);
    return 0;
}
