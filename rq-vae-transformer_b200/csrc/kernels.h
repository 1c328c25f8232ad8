// internal launch prototypes (host side) shared by the engine translation units
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rqb {
// rq_search.cu
int launch_rq_quantize(const float* x, const float* cb, int64_t N, int K, int C, int D, int64_t* codes, float* quant_list,
                       float* resid_out, cudaStream_t st);
int launch_rq_embed(const int64_t* codes, const float* cb, int64_t N, int D, int K, int C, float* out, bool sum,
                    cudaStream_t st);
// sampler.cu
int launch_sample(const float* logits, const float* q, int B, int V, float temperature, int top_k, float top_p,
                  int64_t* out_idx, const int64_t* force, int64_t out_stride, cudaStream_t st);
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

// gemm_tc.cu -- tcgen05 weight-streaming GEMM (fast tier)
enum GemmTcMode { GT_F32 = 0, GT_BF16 = 1, GT_BF16_GELU = 2, GT_PARTIAL = 3, GT_QKV = 4 };
struct GemmTcParams {
    int N_out, K, B, splits, mode;
    const float* bias;            // [N_out] (nullable)
    const float* residual;        // GT_F32 only: out = residual + ...   [B, ld_out]
    void* out;                    // [B, ld_out] f32 / bf16
    int64_t ld_out;
    float* partial;               // GT_PARTIAL: [splits][B][N_out] f32
    // GT_QKV: rows [0,E) -> q_out [B,E] bf16 ; [E,2E) -> kc ; [2E,3E) -> vc  (cache [B][nh][Tmax][64] bf16, row *t_ptr)
    __nv_bfloat16 *q_out, *kc, *vc;
    int E, nh, Tmax, t_host;
    const int* t_ptr;
};
inline int gemm_tc_bn(int B) { return B <= 16 ? 16 : B <= 32 ? 32 : B <= 64 ? 64 : B <= 128 ? 128 : 256; }
int launch_gemm_tc(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p, bool pdl, cudaStream_t st);
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes_log2, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);
}  // namespace rqb
