/*
 * pa_formats.c -- the reference's on-disk formats, plain C, host only:
 *   checkpoint   gpt2_124M.bin: 256 x int32 header [20240326, 1, maxT, V, L, NH, C] + 16 fp32
 *                tensors in write_tensors_fp32 order (reader paged_infer.c:436-502,
 *                writer train_gpt2.py:237-315)
 *   token stream raw int32 ids, batches of B*T+1 with the reference's wrap rule
 *                (paged_infer.c:769-813)
 *   tokenizer    gpt2_tokenizer.bin: 256 x uint32 header [20240328, 1, vocab] + per token one
 *                length byte and the bytes (paged_infer.c:875-915)
 * so that the unmodified reference binary and this library can run on the same synthetic files.
 * Error convention of the extended API: negative pa_status + pa_last_error(), never exit().
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_internal.h"

#define PA_CKPT_MAGIC 20240326
#define PA_TOK_MAGIC 20240328u

static size_t ckpt_param_count(const pa_model_config* c) {
    size_t V = (size_t)c->vocab_size, T = (size_t)c->max_seq_len, L = (size_t)c->n_layers, C = (size_t)c->channels;
    return V * C + T * C + L * (2 * C + 3 * C * C + 3 * C + C * C + C + 2 * C + 4 * C * C + 4 * C + 4 * C * C + C) + 2 * C;
}

/* The Python writer of the reference (train_gpt2.py:298-320) knows two layouts under the same magic: version 1,
 * every tensor fp32 in model order (the one paged_infer.c reads, :441-488), and version 2, written for the
 * bf16 trainer: the ten weight/bias tensors as bf16 in the order wte wpe qkvw qkvb attprojw attprojb fcw fcb
 * fcprojw fcprojb, then the six layernorm tensors in fp32 (train_gpt2.py:266-297).  Both are read here (bf16 is
 * widened exactly); paged_infer.c itself refuses version 2 ("Bad version in model file", :444). */
static int read_header(const char* path, pa_model_config* cfg, int* version);
int pa_checkpoint_read_config(const char* path, pa_model_config* cfg) {
    int version;
    return read_header(path, cfg, &version);
}
static int read_header(const char* path, pa_model_config* cfg, int* version) {
    if (!path || !cfg) { pa_set_error("pa_checkpoint_read_config: NULL argument"); return PA_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { pa_set_error("Error opening model file %s", path); return PA_ERR_INVALID; }
    int hdr[256];
    size_t n = fread(hdr, sizeof(int), 256, f);
    fclose(f);
    if (n != 256) { pa_set_error("model file %s: short header", path); return PA_ERR_INVALID; }
    if (hdr[0] != PA_CKPT_MAGIC) { pa_set_error("Bad magic model file"); return PA_ERR_INVALID; }          /* :443 */
    if (hdr[1] != 1 && hdr[1] != 2) { pa_set_error("Bad version in model file"); return PA_ERR_INVALID; } /* :444 */
    *version = hdr[1];
    cfg->max_seq_len = hdr[2];
    cfg->vocab_size = hdr[3];
    cfg->n_layers = hdr[4];
    cfg->n_heads = hdr[5];
    cfg->channels = hdr[6];
    if (cfg->max_seq_len < 1 || cfg->vocab_size < 1 || cfg->n_layers < 1 || cfg->n_heads < 1 || cfg->channels < 1 ||
        cfg->channels % cfg->n_heads) {
        pa_set_error("model file %s: invalid hyperparameters", path);
        return PA_ERR_INVALID;
    }
    return PA_OK;
}

/* tensor sizes in MODEL order (wte wpe ln1w ln1b qkvw qkvb attprojw attprojb ln2w ln2b fcw fcb fcprojw fcprojb lnfw lnfb) */
static void tensor_sizes(const pa_model_config* c, size_t sz[16]) {
    size_t V = (size_t)c->vocab_size, T = (size_t)c->max_seq_len, L = (size_t)c->n_layers, C = (size_t)c->channels;
    size_t s[16] = {V * C, T * C, L * C, L * C, L * 3 * C * C, L * 3 * C, L * C * C, L * C, L * C, L * C,
                    L * 4 * C * C, L * 4 * C, L * 4 * C * C, L * C, C, C};
    memcpy(sz, s, sizeof(s));
}
/* version 2 file order: model tensor index of each tensor in the file, bf16 ones first */
static const int kV2Order[16] = {0, 1, 4, 5, 6, 7, 10, 11, 12, 13, /* fp32: */ 2, 3, 8, 9, 14, 15};
static const int kV2Bf16 = 10;

int pa_checkpoint_read_params(const char* path, float* params, size_t n_floats) {
    pa_model_config cfg;
    int version;
    int rc = read_header(path, &cfg, &version);
    if (rc != PA_OK) return rc;
    if (!params || n_floats != ckpt_param_count(&cfg)) {
        pa_set_error("pa_checkpoint_read_params: buffer holds %zu floats, the file %zu", n_floats, ckpt_param_count(&cfg));
        return PA_ERR_INVALID;
    }
    FILE* f = fopen(path, "rb");
    if (!f) { pa_set_error("Error opening model file %s", path); return PA_ERR_INVALID; }
    fseek(f, 256 * (long)sizeof(int), SEEK_SET);
    if (version == 1) {
        size_t n = fread(params, sizeof(float), n_floats, f);
        fclose(f);
        if (n != n_floats) { pa_set_error("model file %s: %zu of %zu parameters", path, n, n_floats); return PA_ERR_INVALID; }
        return PA_OK;
    }
    size_t sz[16], off[16], o = 0;
    tensor_sizes(&cfg, sz);
    for (int i = 0; i < 16; i++) { off[i] = o; o += sz[i]; }
    for (int k = 0; k < 16; k++) {
        const int t = kV2Order[k];
        float* dst = params + off[t];
        if (k < kV2Bf16) {
            /* bf16 -> fp32 in place: read the 16-bit words into the upper half of the destination, widen from the end */
            unsigned short* h = (unsigned short*)dst + sz[t];
            if (fread(h, sizeof(unsigned short), sz[t], f) != sz[t]) { fclose(f); pa_set_error("model file %s: truncated bf16 tensor %d", path, t); return PA_ERR_INVALID; }
            for (size_t i = 0; i < sz[t]; i++) {
                unsigned int bits = (unsigned int)h[i] << 16;
                memcpy(dst + i, &bits, sizeof(bits));          /* dst + i never overtakes h + i */
            }
        } else if (fread(dst, sizeof(float), sz[t], f) != sz[t]) {
            fclose(f);
            pa_set_error("model file %s: truncated fp32 tensor %d", path, t);
            return PA_ERR_INVALID;
        }
    }
    fclose(f);
    return PA_OK;
}

/* round-to-nearest-even fp32 -> bf16, as torch's .to(torch.bfloat16) (NaN stays NaN) */
static unsigned short bf16_rne(float x) {
    unsigned int b;
    memcpy(&b, &x, sizeof(b));
    if ((b & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((b >> 16) | 0x40u);
    b += 0x7fffu + ((b >> 16) & 1u);
    return (unsigned short)(b >> 16);
}
/* version 2 (bf16 weights, fp32 layernorms) in the layout of the reference's write_tensors_bf16 */
int pa_checkpoint_write_bf16(const char* path, const pa_model_config* cfg, const float* params) {
    if (!path || !cfg || !params) { pa_set_error("pa_checkpoint_write_bf16: NULL argument"); return PA_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { pa_set_error("cannot create %s", path); return PA_ERR_INVALID; }
    int hdr[256];
    memset(hdr, 0, sizeof(hdr));
    hdr[0] = PA_CKPT_MAGIC; hdr[1] = 2;
    hdr[2] = cfg->max_seq_len; hdr[3] = cfg->vocab_size; hdr[4] = cfg->n_layers; hdr[5] = cfg->n_heads; hdr[6] = cfg->channels;
    int ok = fwrite(hdr, sizeof(int), 256, f) == 256;
    size_t sz[16], off[16], o = 0;
    tensor_sizes(cfg, sz);
    for (int i = 0; i < 16; i++) { off[i] = o; o += sz[i]; }
    unsigned short buf[4096];
    for (int k = 0; k < 16 && ok; k++) {
        const int t = kV2Order[k];
        const float* src = params + off[t];
        if (k >= kV2Bf16) { ok = fwrite(src, sizeof(float), sz[t], f) == sz[t]; continue; }
        for (size_t i0 = 0; i0 < sz[t] && ok; i0 += 4096) {
            const size_t n = sz[t] - i0 < 4096 ? sz[t] - i0 : 4096;
            for (size_t i = 0; i < n; i++) buf[i] = bf16_rne(src[i0 + i]);
            ok = fwrite(buf, sizeof(unsigned short), n, f) == n;
        }
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { pa_set_error("short write to %s", path); return PA_ERR_INVALID; }
    return PA_OK;
}

int pa_checkpoint_write(const char* path, const pa_model_config* cfg, const float* params) {
    if (!path || !cfg || !params) { pa_set_error("pa_checkpoint_write: NULL argument"); return PA_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { pa_set_error("cannot create %s", path); return PA_ERR_INVALID; }
    int hdr[256];
    memset(hdr, 0, sizeof(hdr));
    hdr[0] = PA_CKPT_MAGIC; hdr[1] = 1;
    hdr[2] = cfg->max_seq_len; hdr[3] = cfg->vocab_size; hdr[4] = cfg->n_layers; hdr[5] = cfg->n_heads; hdr[6] = cfg->channels;
    size_t n = ckpt_param_count(cfg);
    int ok = fwrite(hdr, sizeof(int), 256, f) == 256 && fwrite(params, sizeof(float), n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { pa_set_error("short write to %s", path); return PA_ERR_INVALID; }
    return PA_OK;
}

int pa_model_create_from_checkpoint(pa_handle* h, const char* path, int max_batch, pa_model** out) {
    pa_model_config cfg;
    int rc = pa_checkpoint_read_config(path, &cfg);
    if (rc != PA_OK) return rc;
    size_t n = ckpt_param_count(&cfg);
    float* params = (float*)malloc(n * sizeof(float));
    if (!params) { pa_set_error("pa_model_create_from_checkpoint: out of host memory"); return PA_ERR_NOMEM; }
    rc = pa_checkpoint_read_params(path, params, n);
    if (rc == PA_OK) rc = pa_model_create(h, &cfg, params, 0, max_batch, out);
    free(params);
    return rc;
}

/* ---- token stream: dataloader_init / _reset / _next_batch / _free, paged_infer.c:769-818 ------ */
struct pa_dataloader {
    int B, T;
    FILE* tokens_file;
    long file_size;
    long current_position;
    int* batch;          /* B*T+1 ids: inputs = batch, targets = batch + 1 */
    int num_batches;
};

int pa_dataloader_open(const char* path, int B, int T, pa_dataloader** out) {
    if (!path || !out || B < 1 || T < 1) { pa_set_error("pa_dataloader_open: bad arguments"); return PA_ERR_INVALID; }
    *out = NULL;
    FILE* f = fopen(path, "rb");
    if (!f) { pa_set_error("Error opening tokens file %s", path); return PA_ERR_INVALID; }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if ((size_t)size < ((size_t)B * T + 1) * sizeof(int)) {                                              /* :784 */
        fclose(f);
        pa_set_error("Error: file size is too small for the batch size and sequence length");
        return PA_ERR_INVALID;
    }
    pa_dataloader* d = (pa_dataloader*)calloc(1, sizeof(*d));
    int* batch = (int*)malloc(((size_t)B * T + 1) * sizeof(int));
    if (!d || !batch) { fclose(f); free(d); free(batch); pa_set_error("pa_dataloader_open: out of host memory"); return PA_ERR_NOMEM; }
    d->B = B; d->T = T; d->tokens_file = f; d->file_size = size; d->current_position = 0; d->batch = batch;
    d->num_batches = (int)((size_t)size / ((size_t)B * T * sizeof(int)));
    *out = d;
    return PA_OK;
}
void pa_dataloader_reset(pa_dataloader* d) { if (d) d->current_position = 0; }
int pa_dataloader_num_batches(const pa_dataloader* d) { return d ? d->num_batches : 0; }
/* the next B*T+1 ids; *inputs = ids[0..B*T), *targets = ids[1..B*T]; wraps like :802-805 */
int pa_dataloader_next_batch(pa_dataloader* d, const int** inputs, const int** targets) {
    if (!d) { pa_set_error("pa_dataloader_next_batch: NULL loader"); return PA_ERR_INVALID; }
    const size_t n = (size_t)d->B * d->T;
    if ((size_t)d->current_position + (n + 1) * sizeof(int) > (size_t)d->file_size) d->current_position = 0;
    fseek(d->tokens_file, d->current_position, SEEK_SET);
    if (fread(d->batch, sizeof(int), n + 1, d->tokens_file) != n + 1) { pa_set_error("tokens file: short read"); return PA_ERR_INVALID; }
    d->current_position += (long)(n * sizeof(int));
    if (inputs) *inputs = d->batch;
    if (targets) *targets = d->batch + 1;
    return PA_OK;
}
void pa_dataloader_close(pa_dataloader* d) {
    if (!d) return;
    fclose(d->tokens_file);
    free(d->batch);
    free(d);
}
int pa_tokens_write(const char* path, const int* ids, size_t n) {
    if (!path || !ids) { pa_set_error("pa_tokens_write: NULL argument"); return PA_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { pa_set_error("cannot create %s", path); return PA_ERR_INVALID; }
    int ok = fwrite(ids, sizeof(int), n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { pa_set_error("short write to %s", path); return PA_ERR_INVALID; }
    return PA_OK;
}

/* ---- tokenizer: tokenizer_init / tokenizer_decode, paged_infer.c:875-915 ------------------------- */
struct pa_tokenizer {
    unsigned vocab_size;
    char** token_table;
};

int pa_tokenizer_open(const char* path, pa_tokenizer** out) {
    if (!path || !out) { pa_set_error("pa_tokenizer_open: NULL argument"); return PA_ERR_INVALID; }
    *out = NULL;
    FILE* f = fopen(path, "rb");
    if (!f) { pa_set_error("WARNING: Failed to open the tokenizer file %s", path); return PA_ERR_INVALID; }
    unsigned hdr[256];
    if (fread(hdr, sizeof(unsigned), 256, f) != 256 || hdr[0] != PA_TOK_MAGIC || hdr[1] != 1) {
        fclose(f);
        pa_set_error("tokenizer file %s: bad header", path);
        return PA_ERR_INVALID;
    }
    pa_tokenizer* t = (pa_tokenizer*)calloc(1, sizeof(*t));
    if (t) { t->vocab_size = hdr[2]; t->token_table = (char**)calloc(hdr[2] ? hdr[2] : 1, sizeof(char*)); }
    if (!t || !t->token_table) { fclose(f); free(t); pa_set_error("pa_tokenizer_open: out of host memory"); return PA_ERR_NOMEM; }
    for (unsigned i = 0; i < t->vocab_size; i++) {
        unsigned char len;
        if (fread(&len, 1, 1, f) != 1 || len == 0) { fclose(f); pa_tokenizer_close(t); pa_set_error("tokenizer file %s: bad token %u", path, i); return PA_ERR_INVALID; }
        char* piece = (char*)malloc((size_t)len + 1);
        if (!piece || fread(piece, 1, len, f) != len) { free(piece); fclose(f); pa_tokenizer_close(t); pa_set_error("tokenizer file %s: truncated at token %u", path, i); return PA_ERR_INVALID; }
        piece[len] = '\0';
        t->token_table[i] = piece;
    }
    fclose(f);
    *out = t;
    return PA_OK;
}
unsigned pa_tokenizer_vocab_size(const pa_tokenizer* t) { return t ? t->vocab_size : 0; }
/* NULL for an id outside the vocabulary, as tokenizer_decode (:903-915) */
const char* pa_tokenizer_decode(const pa_tokenizer* t, unsigned token_id) {
    if (!t || token_id >= t->vocab_size) return NULL;
    return t->token_table[token_id];
}
void pa_tokenizer_close(pa_tokenizer* t) {
    if (!t) return;
    if (t->token_table) for (unsigned i = 0; i < t->vocab_size; i++) free(t->token_table[i]);
    free(t->token_table);
    free(t);
}
/* pieces[i] is lens[i] (1..255) raw bytes */
int pa_tokenizer_write(const char* path, const char* const* pieces, const unsigned char* lens, unsigned vocab_size) {
    if (!path || !pieces || !lens) { pa_set_error("pa_tokenizer_write: NULL argument"); return PA_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { pa_set_error("cannot create %s", path); return PA_ERR_INVALID; }
    unsigned hdr[256];
    memset(hdr, 0, sizeof(hdr));
    hdr[0] = PA_TOK_MAGIC; hdr[1] = 1; hdr[2] = vocab_size;
    int ok = fwrite(hdr, sizeof(unsigned), 256, f) == 256;
    for (unsigned i = 0; ok && i < vocab_size; i++) {
        if (lens[i] == 0) { ok = 0; break; }
        ok = fwrite(&lens[i], 1, 1, f) == 1 && fwrite(pieces[i], 1, lens[i], f) == lens[i];
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { pa_set_error("short write to %s (or a zero-length token)", path); return PA_ERR_INVALID; }
    return PA_OK;
}
