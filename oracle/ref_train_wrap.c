/*
 * ref_train_wrap.c -- TEST INFRASTRUCTURE ONLY (oracle).
 * Appended by oracle/build_ref.sh to the unmodified reference train_gpt2.c (built
 * with -DTESTING so its main is dropped, train_gpt2.c:961).  Exposes the contiguous
 * attention_forward (train_gpt2.c:220-294) -- the differential property the
 * reference's own test_paged_attn.c checks is "paged == contiguous".
 */
#define REF_API __attribute__((visibility("default")))
REF_API void ref_attention_forward(float* out, float* preatt, float* att, float* inp,
                                   int B, int T, int C, int NH) {
    attention_forward(out, preatt, att, inp, B, T, C, NH);
}
