// internal launch prototypes (host side) shared by the engine translation units
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rqb {
// rq_search.cu
int launch_rq_quantize(const float* x, const float* cb, int64_t N, int K, int C, int D, int64_t* codes, float* quant_list,
                       float* resid_out, cudaStream_t st, int form = 0);
// rq_search2.cu -- 8x8 register tile, codebook streamed in 32-channel slabs, 2-CTA clusters splitting the codebook (the default form)
bool rq_quantize2_supported(int64_t N, int K, int C);
int launch_rq_quantize2(const float* x, const float* cb, int64_t N, int K, int C, int D, int64_t* codes, float* quant_list,
                        float* resid_out, cudaStream_t st);
int launch_rq_embed(const int64_t* codes, const float* cb, int64_t N, int D, int K, int C, float* out, bool sum,
                    cudaStream_t st);
// sampler.cu
int launch_sample(const float* logits, const float* q, int B, int V, float temperature, int top_k, float top_p,
                  int64_t* out_idx, const int64_t* force, int64_t out_stride, cudaStream_t st, int algo = 1);
// ar_kernels.cu
int launch_linear(const float* X, int64_t ldx, const void* W, int wdtype, const float* bias, const float* R, float* Y,
                  int64_t ldy, int M, int N, int K, int act, cudaStream_t st);
int launch_layernorm(const float* X, int64_t ldx, const float* g, const float* b, float* Y, int64_t ldy, int M, int E,
                     cudaStream_t st);
int launch_attn_cached(const float* qkv, float* kc, float* vc, float* out, int B, int Tn, int T_past, int Tmax, int E,
                       int nh, cudaStream_t st);
int launch_code_emb(const int64_t* codes, const float* cb, int B, int HW, int D, int K, int C, int j0, int J, float* out,
                    cudaStream_t st);
int launch_body_token(const float* lin, const float* pos_hw, int B, int D, int E, int j0, int J, int s0, int Tn, float* X,
                      cudaStream_t st);
int launch_cond_token(const int64_t* cond, const float* cond_emb, const float* pos_cond, int B, int cond_len, int vocab_cond,
                      int E, int Tn, float* X, cudaStream_t st);
int launch_head_cumsum(const int64_t* codes, const float* cb, int B, int HW, int D, int K, int C, int j, int d, float* out,
                       cudaStream_t st);
int launch_row_add(const float* in, int64_t in_row_stride, int64_t in_off, const float* pos, int B, int E, float* out,
                   cudaStream_t st);
// conv_kernels.cu
struct ConvGeom {
    int B, Hi, Wi, Cin;      // input NHWC (before the optional fused nearest x2 upsample)
    int Ho, Wo, Cout;        // output
    int KH, KW, stride;      // 3x3 / 1x1 ; stride 1|2
    int pad;                 // symmetric zero pad of the (possibly upsampled) input; stride-2 convs use pad=0 + implicit
                             // bottom/right zero row/col (F.pad (0,1,0,1), layers.py:52-54)
    int upsample;            // 1: the conv reads nearest-x2-upsampled input (layers.py:31-35)
    int out_nchw;            // 1: write [B,Cout,Ho,Wo] (final conv_out)
    int in_nchw;             // 1: read [B,Cin,Hi,Wi] (encoder conv_in)
};
int launch_conv(const float* X, const void* W, int wdtype, const float* bias, const float* R, float* Y, const ConvGeom& g,
                cudaStream_t st);
int launch_groupnorm_silu(const float* X, const float* gamma, const float* beta, float* Y, double* stats_ws, int B, int HW,
                          int C, int silu, cudaStream_t st);
int launch_vae_attn(const float* qkv, float* out, int B, int HW, int C, cudaStream_t st);
size_t groupnorm_ws_doubles(int B, int HW);
int launch_gn_stats(const float* X, double* stats_ws, int B, int HW, int C, cudaStream_t st);
// conv_tc.cu -- tcgen05 implicit-GEMM conv (fast tier) and its fp16 operand producers
bool conv_tc_supported(int H, int W, int Cin, int Cout, int ks, int stride, int in_nchw);
int launch_conv_tc(const void* X16, const void* W16, const void* X16lo, const void* W16lo, const float* bias,
                   const float* residual, float* out, int B, int H, int W, int Cin, int Cout, int ks, int out_nchw,
                   cudaStream_t st, int stride = 1, double* gn_part = nullptr);
bool conv_tc_gn_fusable(int H, int W, int Cout);
// CTA-pair form (csrc/rows_gemm2.cu: cta_group::2, 256 x 256 tiles); launch_rows_gemm_tc dispatches to it when the shape allows
bool rows_gemm2_supported(int64_t M, int N_out, int K);
int launch_rows_gemm2_tc(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32, void* out_16,
                         int gelu, int fmt, int64_t M, int N_out, int K, cudaStream_t st);
int launch_rows_gemm_tc(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32, void* out_16,
                        int gelu, int fmt, int64_t M, int N_out, int K, cudaStream_t st);
int launch_groupnorm_f16(const float* X, const float* gamma, const float* beta, void* Y16, void* Y16lo, double* stats_ws, int B,
                         int HW, int C, int silu, cudaStream_t st, int fused_chunks = 0);
int launch_cast_f16(const float* X, void* Y16, void* Y16lo, int B, int H, int W, int C, int upsample, cudaStream_t st);
int make_tmap_4d_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_c,
                      uint32_t box_w, uint32_t box_h, uint32_t box_b, uint32_t stride = 1);

// gemm_tc.cu -- tcgen05 weight-streaming GEMM (fast tier)
enum GemmTcMode { GT_F32 = 0, GT_H16 = 1, GT_H16_GELU = 2, GT_PARTIAL = 3 };
struct GemmTcParams {
    int N_out, K, B, splits, mode;   // B = activation rows (batch rows of the cached step, or B*T tokens of a prefill / forward pass)
    int fmt;                      // 16-bit operand / output format: 0 = fp16 (the reference's autocast class), 1 = bf16
    int deep;                     // 1: deepest shared-memory ring (one CTA per SM); 0: half depth (two CTAs of consecutive launches per SM)
    int l2pf;                     // 1: before waiting for the upstream kernel, prefetch into L2 the weight boxes that do not fit the ring
    const float* bias;            // [N_out] (nullable); added as bias * bias_scale
    float bias_scale;
    // GT_F32 only: out = acc + bias + residual[(row0 * res_row_stride) + b * ld_res + n], row0 = res_row_ptr ? *res_row_ptr : 0
    // (ld_res = 0 broadcasts one row -- positional embeddings)
    const float* residual;
    int64_t ld_res, res_row_stride;
    const int* res_row_ptr;
    int res_div;                  // > 1: activation row b reads residual row b / res_div (token-major prefill rows share a positional row)
    void* out;                    // [B, ld_out] f32 / 16-bit
    int64_t ld_out;
    float* partial;               // GT_PARTIAL: [splits][B][N_out] f32 (no bias)
    long long* trace;             // diagnostics, nullable: 4 globaltimer stamps of CTA 0 (entry, dependency resolved, accumulator ready, done)
    int trace_w;                  // diagnostics: stamp 0 = the weight tiles requested ahead of the dependency have landed (instead of entry)
};
inline int gemm_tc_bn(int B) { return B <= 16 ? 16 : B <= 32 ? 32 : B <= 64 ? 64 : B <= 128 ? 128 : 256; }
int make_tmap_weight(CUtensorMap* out, const void* W, int N_out, int K);
int launch_gemm_tc(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p, bool pdl, cudaStream_t st);
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes_log2, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

// ar_fast.cu / sampler.cu -- device-resident per-call state of the fast AR tier
struct StepState {
    int s;          // body sequence index of the token being processed (= cached body keys before it)
    int idx;        // spatial position whose codes are being sampled
    int step;       // tokens sampled so far in this call (indexes noise / logits_out)
    int pad;
    const int64_t* cond;      // [B, cond_len] or null
    int64_t* codes;           // [B, HW, D] working copy (xs)
    const int64_t* force;     // teacher forcing or null
    const float* noise;       // [n_tok][B][V] or null
    float* logits_out;        // [n_tok][B][V] or null
    int64_t noise_stride;
    float temperature;
    int top_k[8];
    float top_p[8];
};
int launch_sample_dyn(const float* logits, const StepState* stt, int d, int B, int V, int HW, int D, cudaStream_t st, bool pdl);

struct ArFast;
ArFast* ar_fast_create(const rqb200_ar_config& cfg, const rqb200_ar_weights& w, const rqb200_block_weights* body,
                       const rqb200_block_weights* head);
void ar_fast_destroy(ArFast* f);
size_t ar_fast_workspace_bytes(const ArFast* f, int B);
// positions [idx_begin, idx_end) of the raster; resume != 0: continue on the KV state the previous call left in this workspace
int ar_fast_sample(ArFast* f, const int64_t* partial, const int64_t* cond, int B, int idx_begin, int idx_end, int resume,
                   float temperature, const int32_t* top_k, const float* top_p, const float* noise, int64_t noise_stride,
                   float* logits_out, const int64_t* force, int64_t* out, void* wsp, size_t ws_bytes, cudaStream_t st);
size_t ar_fast_forward_workspace_bytes(const ArFast* f, int B);
int ar_fast_forward(ArFast* f, const int64_t* codes, const int64_t* cond, int B, float* logits_out, float* cond_logits_out, void* wsp,
                    size_t ws_bytes, cudaStream_t st);
// diagnostics: copies the stage trace of the last replays to the host (RQB200 ar_config.trace); returns the number of launches traced
int ar_fast_trace(ArFast* f, long long* out_host, int cap_launches, char* names, int names_cap);
}  // namespace rqb
