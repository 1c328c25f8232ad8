/*
 * rqb200 -- C ABI of the B200-native RQ-VAE / RQ-Transformer sampling engine (sm_100a).
 *
 * The reference (kakaobrain/rq-vae-transformer @ 341395e) is pure Python/PyTorch and has no plugin / FFI
 * registry; its boundary for this path is the Python class surface of `rqvae.models`.  Each
 * entry point below replaces the *library calls* behind one reference method and is what a ctypes binding in
 * the reference's own classes would call (INTEGRATION.md shows those stubs).  Conventions:
 *   - plain pointers and sizes only; every `const T*` / `T*` tensor argument is a DEVICE pointer unless the
 *     name ends in `_host`; row-major, contiguous; `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success, a negative RQB200_E* code otherwise; `rqb200_last_error()` gives the message.
 *   - entry points never allocate device memory and never synchronise the device; scratch space is a caller
 *     supplied workspace whose size is reported by the matching `*_workspace_bytes` query.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns RQB200_ENODEV.
 */
#ifndef RQB200_H
#define RQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RQB200_OK 0
#define RQB200_EINVAL (-1)   /* bad argument / unsupported shape            */
#define RQB200_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define RQB200_ENODEV (-3)   /* no CUDA device                               */
#define RQB200_EWORKSPACE (-4) /* workspace too small                        */
#define RQB200_ESTATE (-5)   /* engine not finalised / tensor missing        */

#define RQB200_F32 0
#define RQB200_BF16 1
#define RQB200_F16 2

/* arithmetic modes (DESIGN.md "two modes") */
#define RQB200_MODE_EXACT 0  /* fp32 weights + fp32 FFMA: the bit-exact-indices gate                    */
#define RQB200_MODE_FAST 1   /* fp16 (default) or bf16 operands on tcgen05, fp32 accumulate: throughput  */

/* rqb200_ar_config.flags (fast tier; scheduling only -- none of them changes a result bit, except SEQUENTIAL_PREFILL's
 * summation order) */
#define RQB200_AR_NO_GRAPH 1            /* launch kernel by kernel instead of replaying CUDA graphs                        */
#define RQB200_AR_NO_PDL 2              /* plain stream order instead of programmatic dependent launch                     */
#define RQB200_AR_TRACE 4               /* record 4 globaltimer stamps per launch (rqb200_ar_trace)                        */
#define RQB200_AR_L2_PREFETCH 8         /* GEMMs prefetch into L2 the weight boxes that do not fit their shared-memory ring */
#define RQB200_AR_SHALLOW_RING 16       /* half-depth GEMM rings: two GEMM CTAs of consecutive launches share an SM        */
#define RQB200_AR_SEQUENTIAL_PREFILL 32 /* prefill the prefix token by token with the single-step graph (the prefill oracle) */
#define RQB200_AR_BATCHED_DEEP_RING 64  /* large-M passes (prefill / forward) keep the deep ring: one CTA per SM                  */
#define RQB200_AR_ATTN_ONE_WARP 256    /* body attention: one warp per (b, head) (round-1/2 form) instead of four               */
#define RQB200_AR_TRACE_WEIGHTS 512    /* with TRACE: GEMM stamp 0 = prefetched weight tiles landed (instead of kernel entry)    */
#define RQB200_AR_NO_PARAM_PREFETCH 1024 /* LN1 of block l does not prefetch block l+1's small vectors into L2                  */
#define RQB200_AR_BATCHED_STREAMER 128  /* large-M passes through the weight-streaming GEMM instead of the persistent rows GEMM    */

const char* rqb200_last_error(void);
int rqb200_version(void);
int rqb200_device_count(void);

/* ------------------------------------------------------------------------------------------------ P1
 * RQBottleneck.quantize  (rqvae/models/rqvae/quantizations.py:237-271; VQEmbedding.compute_distances :43-62,
 * find_nearest_embedding :64-69, embed :144-146).  x [N,C] f32, codebook [K,C] f32 (weight[:-1], no padding row).
 * codes [N,D] int64.  quant_list (nullable) [D,N,C] f32 = the D cumulative aggregates (quant_list[i] of the
 * reference).  residual_out (nullable) [N,C] f32 = x - quant_list[D-1].  C must be 256 (quantizations.py:181). */
int rqb200_rq_quantize(const float* x, const float* codebook, int64_t N, int K, int C, int D, int64_t* codes,
                       float* quant_list, float* residual_out, void* stream);

/* One depth of RQBottleneck.get_soft_codes (quantizations.py:371-399): residual [N,C] f32 -> soft_out [N,K] = softmax(-d/temp) with
 * d = VQEmbedding.compute_distances (:43-62); logits_out (nullable) [N,K] = -d/temp (what the stochastic variant samples from). */
int rqb200_rq_soft_codes(const float* residual, const float* codebook, int64_t N, int K, int C, float temp, float* soft_out,
                         float* logits_out, void* stream);

/* RQBottleneck.embed_code (quantizations.py:297-311): out[n,:] = sum_d codebook[codes[n,d],:]  (order d=0..D-1). */
int rqb200_rq_embed_sum(const int64_t* codes, const float* codebook, int64_t N, int D, int K, int C, float* out,
                        void* stream);
/* RQBottleneck.embed_code_with_depth (quantizations.py:313-334): out[n,d,:] = codebook[codes[n,d],:]. */
int rqb200_rq_embed_depth(const int64_t* codes, const float* codebook, int64_t N, int D, int K, int C, float* out,
                          void* stream);

/* ------------------------------------------------------------------------------------------------ sampler
 * sample_from_logits (rqvae/utils/utils.py:82-123; top_k_logits :60-64, top_p_probs :67-79).  logits [B,V] f32.
 * q (nullable) [B,V] f32 Exp(1) noise: torch.multinomial(probs,1) == argmax(probs/q) (SURVEY.md finding 7);
 * NULL means q == 1 (arg-max of the filtered distribution).  top_k <= 0 or >= V disables top-k; top_p >= 1
 * takes the reference's p = 1.0 branch.  out_idx [B] int64.  V <= 16384.  No host sync (the reference's NaN
 * check syncs; here NaN -> -inf is done on the device). */
int rqb200_sample_logits(const float* logits, const float* q, int B, int V, float temperature, int top_k,
                         float top_p, int64_t* out_idx, void* stream);

/* ------------------------------------------------------------------------------------------------ P3
 * RQTransformer (rqvae/models/rqtransformer/transformers.py) -- cached AR sampling. */
typedef struct rqb200_block_weights {
    const void *wqkv, *wproj, *w1, *w2;            /* [3E,E] (rows: query|key|value), [E,E], [4E,E], [E,4E]; weight dtype */
    const float *bqkv, *bproj, *b1, *b2;           /* f32 biases */
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;    /* f32 */
} rqb200_block_weights;

typedef struct rqb200_ar_config {
    int32_t embed_dim, n_head, n_body, n_head_layers;  /* E, heads (E/heads must be 64), body / head depth   */
    int32_t vocab, H, W, D;                             /* V and block_size                                    */
    int32_t vocab_cond, cond_len;                       /* cond_emb rows, block_size_cond (>=1)                */
    int32_t code_dim, codebook_size;                    /* C (=256) and K of the RQ-VAE codebook               */
    int32_t mode;                                       /* RQB200_MODE_*                                       */
    int32_t weight_dtype;                               /* RQB200_F32 (exact); RQB200_F16 or RQB200_BF16 (fast) */
    int32_t flags;                                      /* fast tier: RQB200_AR_* flags                         */
    int32_t split_qkv, split_proj, split_fc1, split_fc2; /* fast tier: split-K factors, 0 = fill the SMs        */
} rqb200_ar_config;

typedef struct rqb200_ar_weights {
    const float *pos_emb_cond, *pos_emb_hw, *pos_emb_d;  /* [cond_len,E], [H*W,E], [D,E] f32                    */
    const float* cond_emb;                                /* [vocab_cond,E] f32                                  */
    const void *w_in, *w_head, *w_cls;                    /* [E,C], [E,C], [V,E]; weight dtype                   */
    const float *b_in, *b_head, *b_cls;
    const float *cls_ln_w, *cls_ln_b;
    const float* codebook;                                /* [K,C] f32 (model_aux.get_code_emb_with_depth)       */
    const rqb200_block_weights* body;                     /* host array [n_body]                                 */
    const rqb200_block_weights* head;                     /* host array [n_head_layers]                          */
    /* optional (cond_len > 1): cond_classifier (transformers.py:100-104) -- only rqb200_ar_forward's cond_logits use it */
    const void* w_ccls;                                   /* [Vc,E], Vc = vocab_cond rounded up to 128 (zero rows); weight dtype; NULL when absent */
    const float *b_ccls, *ccls_ln_w, *ccls_ln_b;
} rqb200_ar_weights;

typedef struct rqb200_ar rqb200_ar;

rqb200_ar* rqb200_ar_create(const rqb200_ar_config* cfg, const rqb200_ar_weights* w);
void rqb200_ar_destroy(rqb200_ar* h);
size_t rqb200_ar_workspace_bytes(const rqb200_ar* h, int B);

/* RQTransformer.sample (transformers.py:294-369) with cached_forward (:190-287) and sample_from_logits fused
 * into one device-side loop (no host sync per token).
 *   partial [B,H,W,D] int64 (prefix used when start_h/start_w > 0), cond [B,cond_len] int64 or NULL (zeros),
 *   top_k_host[D] / top_p_host[D]: per-depth settings (HOST arrays, already clamped like :314-330),
 *   noise (nullable): per-token Exp(1) draws, token t (= the t-th *sampled* (h,w,d) in raster order) at
 *   noise + t*noise_stride, each [B,V] f32; NULL -> q = 1,
 *   logits_out (nullable) [n_tokens,B,V] f32 receives every step's logits (teacher-forcing / parity tests),
 *   force_codes (nullable) [B,H,W,D] int64: teacher forcing -- logits are computed and (optionally) dumped but
 *   the code written back is force_codes' (so the step-parity protocol of SURVEY.md 8c can be run),
 *   out_codes [B,H,W,D] int64. */
int rqb200_ar_sample(rqb200_ar* h, const int64_t* partial, const int64_t* cond, int B, int start_h, int start_w,
                     float temperature, const int32_t* top_k_host, const float* top_p_host, const float* noise,
                     int64_t noise_stride, float* logits_out, const int64_t* force_codes, int64_t* out_codes,
                     void* workspace, size_t workspace_bytes, void* stream);
/* The same loop over the positions [idx_begin, idx_end) of the raster only.  resume == 0: starts like rqb200_ar_sample with
 * start_loc = idx_begin (prefix prefill from `partial`); resume != 0: continues on the KV / context state the previous call left
 * in the SAME workspace (no prefill; `partial` is ignored, out_codes must be the buffer of the previous call).  noise /
 * logits_out are indexed from the first token of THIS span.  Lets a caller draw the per-token noise in bounded chunks. */
int rqb200_ar_sample_span(rqb200_ar* h, const int64_t* partial, const int64_t* cond, int B, int idx_begin, int idx_end, int resume,
                          float temperature, const int32_t* top_k_host, const float* top_p_host, const float* noise,
                          int64_t noise_stride, float* logits_out, const int64_t* force_codes, int64_t* out_codes,
                          void* workspace, size_t workspace_bytes, void* stream);
/* RQTransformer.forward (transformers.py:113-188): teacher-forced logits of complete code maps, all positions at once (fast tier:
 * M = B*T row GEMMs on tcgen05 + causal attention; exact tier: returns RQB200_EINVAL -- use rqb200_ar_sample with force_codes and
 * logits_out, the sequential replay).  codes [B,H,W,D] int64, cond [B,cond_len] or NULL.
 * logits_out [D][H*W][B][V] f32 (token-major: logits of (b, pos, d) at ((d*H*W + pos)*B + b)*V); cond_logits_out (nullable,
 * cond_len > 1 and w_ccls given) [cond_len-1][B][Vc] f32 with Vc = vocab_cond rounded up to a multiple of 128. */
size_t rqb200_ar_forward_workspace_bytes(const rqb200_ar* h, int B);
int rqb200_ar_forward(rqb200_ar* h, const int64_t* codes, const int64_t* cond, int B, float* logits_out, float* cond_logits_out,
                      void* workspace, size_t workspace_bytes, void* stream);
/* fast tier with RQB200_AR_TRACE: copies 4 globaltimer stamps (ns: entry, dependency resolved, accumulator ready / mid, done) per
 * launch slot of the last graph replays to out_host[cap_launches][4] and the slot names ('\n'-separated) to names; returns the
 * number of slots (0 when tracing is off).  Synchronises the device. */
int rqb200_ar_trace(rqb200_ar* h, long long* out_host, int cap_launches, char* names, int names_cap);
/* number of kernels the last rqb200_ar_sample call launched (bench.py's gpu_launches) */
int64_t rqb200_ar_last_launches(const rqb200_ar* h);

/* ------------------------------------------------------------------------------------------------ P2
 * RQVAE encode / decode (rqvae/models/rqvae/rqvae.py:80-109; modules.py:73-98,171-202; layers.py). */
/* OR-ed into rqb200_vae_config.mode (fast tier; diagnostics): GroupNorm statistics by the stand-alone gn_stats pass instead of the
 * producing conv's epilogue */
#define RQB200_VAE_NO_GN_FUSE 0x100
typedef struct rqb200_vae_config {
    int32_t ch, n_levels, ch_mult[8], num_res_blocks;
    int32_t n_attn_res, attn_resolutions[8];
    int32_t resolution, z_channels, embed_dim, in_channels, out_ch;
    int32_t codebook_size, depth;     /* K, D */
    int32_t mode;                     /* RQB200_MODE_*: EXACT = f32 conv weights, FAST = f16 conv weights; | RQB200_VAE_* flags */
} rqb200_vae_config;

typedef struct rqb200_vae rqb200_vae;

rqb200_vae* rqb200_vae_create(const rqb200_vae_config* cfg);
void rqb200_vae_destroy(rqb200_vae* h);
/* register one tensor under its reference state_dict key.  Conv weights must be passed
 * re-laid-out as [Cout,KH,KW,Cin] (OHWI) in the engine's weight dtype; everything else f32 as stored. */
int rqb200_vae_set_tensor(rqb200_vae* h, const char* key, const void* ptr, int dtype, int64_t numel);
/* resolves every layer of encoder+decoder against the registered tensors; fails listing the first missing key */
int rqb200_vae_finalize(rqb200_vae* h);
size_t rqb200_vae_workspace_bytes(const rqb200_vae* h, int B);
/* RQVAE.decode (rqvae.py:85-89): z_q [B,h,w,embed_dim] f32 NHWC -> out [B,out_ch,R,R] f32 NCHW */
int rqb200_vae_decode(rqb200_vae* h, const float* z_q, int B, float* out, void* workspace, size_t workspace_bytes,
                      void* stream);
/* RQVAE.decode_code (rqvae.py:105-109): codes [B,h,w,D] int64 -> out NCHW */
int rqb200_vae_decode_code(rqb200_vae* h, const int64_t* codes, int B, float* out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* RQVAE.encode (rqvae.py:80-83): x [B,in_channels,R,R] f32 NCHW -> z_e [B,h,w,embed_dim] f32 NHWC */
int rqb200_vae_encode(rqb200_vae* h, const float* x, int B, float* z_e, void* workspace, size_t workspace_bytes,
                      void* stream);
int64_t rqb200_vae_last_launches(const rqb200_vae* h);

/* ------------------------------------------------------------------------------------------------ diagnostics
 * Single-kernel entry points used by tests/ and bench.py's roofline leg; not part of the reference-facing surface.
 * rqb200_dbg_gemm_tc: one launch of the tcgen05 weight-streaming GEMM (csrc/gemm_tc.cu):
 *   out[b, n] = act(sum_k W[n,k] X[b,k] + bias[n]) (+ residual[b,n]);  W [N_out,K], X [B,K] both 16-bit: fmt 0 = fp16, 1 = bf16;
 *   partial != NULL: partial [splits,B,N_out] f32 receives the per-split sums instead (no bias / act / residual; B <= 256).
 *   B > 256 (splits == 1) runs as row chunks of 256 (the batched-prefill / teacher-forced-forward shape). */
int rqb200_dbg_gemm_tc(const void* W16, const void* X16, const float* bias, const float* residual, void* out,
                       int out_is_16, int gelu, float* partial, int N_out, int K, int B, int splits, int fmt, void* stream);

/* rqb200_dbg_rq_quantize: rqb200_rq_quantize with the kernel form forced: 1 = csrc/rq_search.cu (2x4 register tile), 2 =
 * csrc/rq_search2.cu (8x8 register tile, 2-CTA clusters splitting the codebook; fails with RQB200_EINVAL for shapes it does not
 * take), 0 = what rqb200_rq_quantize picks.  Both forms are bit-identical (tests/test_gpu_parity.py). */
int rqb200_dbg_rq_quantize(int form, const float* x, const float* codebook, int64_t N, int K, int C, int D, int64_t* codes,
                           float* quant_list, float* residual_out, void* stream);
/* rqb200_dbg_sample_logits: rqb200_sample_logits with the top-k threshold search forced: 0 = 8-pass radix select, 1 = bucket
 * select (the default).  Identical indices. */
int rqb200_dbg_sample_logits(int algo, const float* logits, const float* q, int B, int V, float temperature, int top_k,
                             float top_p, int64_t* out_idx, void* stream);

/* rqb200_dbg_conv_tc: one launch of the tcgen05 implicit-GEMM conv (csrc/conv_tc.cu): X NHWC fp16 [B,H,W,Cin], W OHWI fp16
 * [Cout,ks,ks,Cin], stride 1 "same" padding, out f32 NHWC (+bias, +residual) or NCHW when out_nchw.  X16lo / W16lo
 * (both or neither): the fp16 "lo" halves (value - fp16(value)) -> split-fp16, three products per conv. */
int rqb200_dbg_conv_tc(const void* X16, const void* W16, const void* X16lo, const void* W16lo, const float* bias,
                       const float* residual, float* out, int B, int H, int W, int Cin, int Cout, int ks, int out_nchw,
                       void* stream);

/* rqb200_dbg_chain: micro-benchmark of ONE dependent stage (csrc/dbg_chain.cu), the quantity that bounds the cached AR step.
 * mode 0: PDL chain of empty kernels in a CUDA graph; 1: PDL chain, every CTA reads 16 KB written by other CTAs of the previous
 * kernel and writes 16 KB; 2: one persistent kernel, same data flow, grid-wide barrier between stages; 3: persistent, every CTA
 * waits only for the `fan` producers it reads.  Runs `reps` timed repetitions of an `n_stages` chain on its own stream and
 * returns microseconds per stage.  workspace (device): >= 2*ctas*16 KB + 4 KB + 4*ctas bytes. */
int rqb200_dbg_chain(int mode, int n_stages, int ctas, int threads, int smem_bytes, int fan, int reps, void* workspace,
                     size_t workspace_bytes, float* us_per_stage);
/* rqb200_dbg_chain2: the same PDL chain with the stage's two halves separable.  variant bit 0: every CTA reads `words` floats that
 * another CTA of the previous kernel wrote (all of a thread's loads in flight); bit 1: every CTA writes `words` floats; bit 2: the
 * reads go to lines nobody writes (clean) instead.  workspace (device) >= 3 * ctas * words * 4 bytes. */
int rqb200_dbg_chain2(int variant, int words, int n_stages, int ctas, int threads, int smem_bytes, int reps, void* workspace,
                      size_t workspace_bytes, float* us_per_stage);

/* rqb200_dbg_rows_gemm: the large-M GEMM of the batched prefill / forward passes (csrc/conv_tc.cu launch_rows_gemm_tc: persistent
 * 128 x BN tiles, double-buffered TMEM): out[m,n] = act(sum_k X[m,k] W[n,k] + bias[n]) (+ residual[m,n]).  X [ceil(M/128)*128, K] and
 * W [N_out,K] 16-bit (fmt 0 fp16 / 1 bf16); exactly one of out_f32 / out_16; gelu applies to out_16 only. */
int rqb200_dbg_rows_gemm(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32, void* out_16,
                         int gelu, int fmt, int64_t M, int N_out, int K, void* stream);
/* rqb200_dbg_tma_rate: micro-benchmark of one SM's shared-memory fill rate from L2 (csrc/dbg_tma.cu).  mode 0: tensor-map boxes
 * of `rows` x 128 B (what the GEMM kernels issue); mode 1: 1-D bulk copies of rows*128 contiguous bytes.  `depth` loads in flight,
 * `iters` rounds, every CTA cycling over the same `boxes_total` boxes of `buffer` (>= boxes_total*rows*128 bytes). */
int rqb200_dbg_tma_rate(int mode, int rows, int depth, int iters, int boxes_total, const void* buffer, int ctas,
                        float* bytes_per_clk, float* us_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* RQB200_H */
