// P3 "fast" tier -- the cached AR step as a PDL-chained, CUDA-graph-replayed sequence of sm_100a kernels.
//
// Same semantics as ar_engine.cu's exact tier (reference: transformers.py:190-369, attentions.py:60-142), different
// arithmetic class: bf16 weights / activations / KV cache on tcgen05 (gemm_tc.cu), fp32 residual stream, fp32
// LayerNorm / softmax / sampler.  What the chain looks like for one transformer block (M = batch rows):
//
//     ln_reduce   x += bias_prev + sum_s partial_prev[s] ; xn = LN(x) (bf16)      <- fused split-K reduction + residual
//     gemm_tc     qkv partials = Wqkv . xn                                         (split-K, 144 CTAs)
//     attn_fast   q,k,v = sum partials + bias ; append k,v to the bf16 cache ; softmax(q k^T/8) v -> att (bf16)
//     gemm_tc     proj partials = Wproj . att
//     ln_reduce   x += bproj + sum partials ; xn = LN2(x)
//     gemm_tc     h = gelu(W1 . xn + b1) (bf16)            (direct epilogue, or split-K + act_reduce)
//     gemm_tc     fc2 partials = W2 . h
//
// Every kernel starts with griddepcontrol.launch_dependents and reads upstream data only after griddepcontrol.wait, so
// the NEXT kernel's prologue -- for the GEMMs: filling the shared-memory ring with weight tiles -- overlaps this one.
// Position-dependent scalars (sequence index, spatial index, token counter) live in a device-side StepState that the
// last kernel of each graph advances, so three captured graphs (cond-token body step, code-token body step, head steps
// + sampling) are replayed for all positions without host involvement.
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace rqb {

template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RQB_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    g_launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------ kernels
// x_out = x_in + bias + sum_s partial[s] (+ extra row) ; xn = LayerNorm(x_out) in bf16.  One CTA per batch row.
__global__ void __launch_bounds__(384)
ln_reduce_kernel(const float* __restrict__ x_in, const float* __restrict__ partial, int S, const float* __restrict__ bias,
                 const float* __restrict__ extra, float* __restrict__ x_out, const float* __restrict__ g,
                 const float* __restrict__ be, __nv_bfloat16* __restrict__ xn, int B, int E) {
    // each thread owns up to 3 float4 chunks of the row (E <= 384*4*3); every load is issued before the first dependent add and
    // the row stays in registers between the statistics and the normalisation
    __shared__ float red[33];
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x;
    const int E4 = E >> 2;
    const int S12 = S < 12 ? S : 12;
    float4 v[3];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int e4 = threadIdx.x + k * 384;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e4 < E4) {
            float4 pr[12];
#pragma unroll
            for (int i = 0; i < 12; i++)
                if (i < S12) pr[i] = reinterpret_cast<const float4*>(partial + ((int64_t)i * B + b) * E)[e4];
            if (x_in) v[k] = reinterpret_cast<const float4*>(x_in + (int64_t)b * E)[e4];
            if (bias) { float4 t = reinterpret_cast<const float4*>(bias)[e4]; v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w; }
#pragma unroll
            for (int i = 0; i < 12; i++)
                if (i < S12) { v[k].x += pr[i].x; v[k].y += pr[i].y; v[k].z += pr[i].z; v[k].w += pr[i].w; }
            for (int i = 12; i < S; i++) {
                float4 p0 = reinterpret_cast<const float4*>(partial + ((int64_t)i * B + b) * E)[e4];
                v[k].x += p0.x; v[k].y += p0.y; v[k].z += p0.z; v[k].w += p0.w;
            }
            if (extra) { float4 t = reinterpret_cast<const float4*>(extra)[e4]; v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w; }
            if (x_out) reinterpret_cast<float4*>(x_out + (int64_t)b * E)[e4] = v[k];
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
    }
    const float mean = block_sum(s, red) / (float)E;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (threadIdx.x + k * 384 < E4) {
            const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
            q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
        }
    const float rstd = rsqrtf(block_sum(q, red) / (float)E + 1e-5f);
    if (xn) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int e4 = threadIdx.x + k * 384;
            if (e4 < E4) {
                const float4 gg = reinterpret_cast<const float4*>(g)[e4], bb = reinterpret_cast<const float4*>(be)[e4];
                __nv_bfloat162 h0 = __floats2bfloat162_rn((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y);
                __nv_bfloat162 h1 = __floats2bfloat162_rn((v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w);
                uint2 pk;
                pk.x = *reinterpret_cast<unsigned*>(&h0);
                pk.y = *reinterpret_cast<unsigned*>(&h1);
                reinterpret_cast<uint2*>(xn + (int64_t)b * E)[e4] = pk;
            }
        }
    }
}

// h = bf16(gelu(sum_s partial[s] + bias))   (only when fc1 runs split-K); 4 elements per thread, all partial loads in flight
__global__ void __launch_bounds__(256)
act_reduce_kernel(const float* __restrict__ partial, int S, const float* __restrict__ bias, __nv_bfloat16* __restrict__ h, int B,
                  int N) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int64_t total4 = (int64_t)B * N / 4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256) {
        const int n = (int)((i * 4) % N);
        float4 pr[4];
#pragma unroll
        for (int s = 0; s < 4; s++)
            if (s < S) pr[s] = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)s * B * N) + i);
        float4 v = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
        for (int s = 0; s < 4; s++)
            if (s < S) { v.x += pr[s].x; v.y += pr[s].y; v.z += pr[s].z; v.w += pr[s].w; }
        for (int s = 4; s < S; s++) {
            float4 p = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)s * B * N) + i);
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = 0.5f * r[k] * (1.0f + erff(r[k] * 0.70710678118654752440f));
        __nv_bfloat162 h0 = __floats2bfloat162_rn(r[0], r[1]), h1 = __floats2bfloat162_rn(r[2], r[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&h0);
        pk.y = *reinterpret_cast<unsigned*>(&h1);
        *reinterpret_cast<uint2*>(h + i * 4) = pk;
    }
}

// one warp per (b, head): reduce the split-K qkv partials (+bias), append k,v at row t of the bf16 cache, attend.
// lane <-> dims (2*lane, 2*lane+1) for q/k/v/out; lane <-> key for the scores (q and the probabilities are
// broadcast through shared memory).  T <= 512.
constexpr int AF_MAXT = 512;
__global__ void __launch_bounds__(128)
attn_fast_kernel(const float* __restrict__ part, int S, const float* __restrict__ bqkv, __nv_bfloat16* __restrict__ kc,
                 __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ att, int B, int E, int nh, int Tmax,
                 const int* __restrict__ t_ptr, int t_host, const float2* __restrict__ stats_in, int nst,
                 const float* __restrict__ cqkv) {
    __shared__ float qs[4][64];
    __shared__ float ps[4][AF_MAXT];
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int bh = blockIdx.x * 4 + wq;
    if (bh >= B * nh) return;
    const int b = bh / nh, h = bh % nh;
    const int t = t_ptr ? *t_ptr : t_host;
    const int c = h * 64 + 2 * lane;
    // stats_in != NULL: the qkv GEMM ran on the raw residual rows with LayerNorm folded into its weights (rqb200_block_weights
    // .cqkv); the row statistics are applied here: q = rstd*(sum - mean*c) + b'
    const bool fold = stats_in != nullptr;
    float2 q = fold ? make_float2(0.f, 0.f) : make_float2(bqkv[c], bqkv[c + 1]);
    float2 k = fold ? make_float2(0.f, 0.f) : make_float2(bqkv[E + c], bqkv[E + c + 1]);
    float2 v = fold ? make_float2(0.f, 0.f) : make_float2(bqkv[2 * E + c], bqkv[2 * E + c + 1]);
#pragma unroll 4
    for (int s = 0; s < S; s++) {
        const float* p = part + ((int64_t)s * B + b) * 3 * E;
        float2 a = *reinterpret_cast<const float2*>(p + c);
        float2 bb = *reinterpret_cast<const float2*>(p + E + c);
        float2 cc = *reinterpret_cast<const float2*>(p + 2 * E + c);
        q.x += a.x; q.y += a.y; k.x += bb.x; k.y += bb.y; v.x += cc.x; v.y += cc.y;
    }
    if (fold) {
        float s1 = 0.f;
        for (int i = lane; i < nst; i += 32) s1 += __ldcg(stats_in + (int64_t)b * nst + i).x;
        const float mean = warp_sum(s1) / (float)E;
        float m2 = 0.f;
        for (int i = lane; i < nst; i += 32) {
            const float2 st = __ldcg(stats_in + (int64_t)b * nst + i);
            const float d = st.x * (1.0f / 128.0f) - mean;
            m2 += st.y + 128.0f * d * d;
        }
        const float rstd = rsqrtf(warp_sum(m2) / (float)E + 1e-5f);
        q.x = rstd * (q.x - mean * cqkv[c]) + bqkv[c];
        q.y = rstd * (q.y - mean * cqkv[c + 1]) + bqkv[c + 1];
        k.x = rstd * (k.x - mean * cqkv[E + c]) + bqkv[E + c];
        k.y = rstd * (k.y - mean * cqkv[E + c + 1]) + bqkv[E + c + 1];
        v.x = rstd * (v.x - mean * cqkv[2 * E + c]) + bqkv[2 * E + c];
        v.y = rstd * (v.y - mean * cqkv[2 * E + c + 1]) + bqkv[2 * E + c + 1];
    }
    __nv_bfloat16* kb = kc + ((int64_t)(b * nh + h) * Tmax) * 64;
    __nv_bfloat16* vb = vc + ((int64_t)(b * nh + h) * Tmax) * 64;
    const __nv_bfloat162 k2 = __floats2bfloat162_rn(k.x, k.y), v2 = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(kb + (int64_t)t * 64 + 2 * lane) = k2;
    *reinterpret_cast<__nv_bfloat162*>(vb + (int64_t)t * 64 + 2 * lane) = v2;
    // use the bf16-rounded q/k/v everywhere (what a later step reads back from the cache)
    const float2 qf = __bfloat1622float2(__floats2bfloat162_rn(q.x, q.y)), kf = __bfloat1622float2(k2), vf = __bfloat1622float2(v2);
    qs[wq][2 * lane] = qf.x;
    qs[wq][2 * lane + 1] = qf.y;
    __syncwarp();
    const float s_new = warp_sum(qf.x * kf.x + qf.y * kf.y) * 0.125f;
    float m = s_new;
    for (int j = lane; j < t; j += 32) {          // scores of the cached rows
        const uint4* kr = reinterpret_cast<const uint4*>(kb + (int64_t)j * 64);
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 w = kr[u];
            const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
            for (int z = 0; z < 4; z++) {
                float2 kk = __bfloat1622float2(kp[z]);
                acc = fmaf(qs[wq][(u * 4 + z) * 2], kk.x, acc);
                acc = fmaf(qs[wq][(u * 4 + z) * 2 + 1], kk.y, acc);
            }
        }
        acc *= 0.125f;
        ps[wq][j] = acc;
        m = fmaxf(m, acc);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < t; j += 32) {
        float e = __expf(ps[wq][j] - m);
        ps[wq][j] = e;
        sum += e;
    }
    const float e_new = __expf(s_new - m);
    sum = warp_sum(sum) + e_new;
    __syncwarp();
    const float inv = 1.0f / sum;
    float2 o = make_float2(e_new * vf.x, e_new * vf.y);
    int j = 0;
    for (; j + 16 <= t; j += 16) {                        // 16 independent V rows in flight
        __nv_bfloat162 raw[16];
#pragma unroll
        for (int u = 0; u < 16; u++) raw[u] = *reinterpret_cast<const __nv_bfloat162*>(vb + (int64_t)(j + u) * 64 + 2 * lane);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const float2 vv = __bfloat1622float2(raw[u]);
            o.x = fmaf(ps[wq][j + u], vv.x, o.x);
            o.y = fmaf(ps[wq][j + u], vv.y, o.y);
        }
    }
    for (; j + 4 <= t; j += 4) {
        __nv_bfloat162 raw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) raw[u] = *reinterpret_cast<const __nv_bfloat162*>(vb + (int64_t)(j + u) * 64 + 2 * lane);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float2 vv = __bfloat1622float2(raw[u]);
            o.x = fmaf(ps[wq][j + u], vv.x, o.x);
            o.y = fmaf(ps[wq][j + u], vv.y, o.y);
        }
    }
    for (; j < t; j++) {
        const float p = ps[wq][j];
        float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + (int64_t)j * 64 + 2 * lane));
        o.x = fmaf(p, vv.x, o.x);
        o.y = fmaf(p, vv.y, o.y);
    }
    *reinterpret_cast<__nv_bfloat162*>(att + (int64_t)b * E + c) = __floats2bfloat162_rn(o.x * inv, o.y * inv);
}

// RQB200_GR chain with folded LayerNorm: x_out = x_in (+ extra row); xq = bf16(x_out); stats[b][tile] = (sum, M2 about the tile
// mean) for every 128-feature tile -- what the GT_GR epilogue of proj / fc2 emits for all later layers.  One CTA per batch row,
// warp <-> tile, lane <-> 4 features.
__global__ void __launch_bounds__(384)
row_prep_kernel(const float* __restrict__ x_in, const float* __restrict__ extra, float* __restrict__ x_out,
                __nv_bfloat16* __restrict__ xq, float2* __restrict__ stats, int B, int E) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, nt = E / 128;
    for (int t = warp; t < nt; t += nw) {
        const int e = t * 128 + 4 * lane;
        float4 v = *reinterpret_cast<const float4*>(x_in + (int64_t)b * E + e);
        if (extra) { const float4 a = *reinterpret_cast<const float4*>(extra + e); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (x_out) *reinterpret_cast<float4*>(x_out + (int64_t)b * E + e) = v;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&h0);
        pk.y = *reinterpret_cast<unsigned*>(&h1);
        *reinterpret_cast<uint2*>(xq + (int64_t)b * E + e) = pk;
        const float s1 = warp_sum((v.x + v.y) + (v.z + v.w));
        const float tm = s1 * (1.0f / 128.0f);
        const float d0 = v.x - tm, d1 = v.y - tm, d2 = v.z - tm, d3 = v.w - tm;
        const float m2 = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
        if (lane == 0) stats[(int64_t)b * nt + t] = make_float2(s1, m2);
    }
}
// arrival counters of the GT_GR GEMMs of one graph: zeroed by the graph's last-but-one kernel for its next replay
__global__ void __launch_bounds__(256) ctr_zero_kernel(unsigned* __restrict__ ctr, int n) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    for (int i = threadIdx.x; i < n; i += 256) ctr[i] = 0u;
}

// token sources --------------------------------------------------------------------------------------------------
// cond token s: x[b,:] = cond_emb[cond[b,s]] + pos_emb_cond[s]                      (transformers.py:224)
__global__ void __launch_bounds__(256)
cond_tok_kernel(const StepState* __restrict__ stt, const float* __restrict__ cond_emb, const float* __restrict__ pos_cond,
                int cond_len, int vocab_cond, int E, float* __restrict__ x) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x, s = stt->s;
    int64_t c = stt->cond ? stt->cond[(int64_t)b * cond_len + s] : 0;
    c = c < 0 ? 0 : (c >= vocab_cond ? vocab_cond - 1 : c);
    for (int e = threadIdx.x; e < E; e += 256) x[(int64_t)b * E + e] = cond_emb[c * E + e] + pos_cond[(int64_t)s * E + e];
}
// summed code embeddings in bf16: mode 0 -> all D codes of position idx-1 (body input), mode d>=1 -> codes 0..d-1 of
// position idx (head input, cumsum)                                              (transformers.py:219-225, 250-255)
__global__ void __launch_bounds__(64)
code_sum_kernel(const StepState* __restrict__ stt, const float* __restrict__ cb, int HW, int D, int K, int C, int mode,
                __nv_bfloat16* __restrict__ out) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x;
    const int pos = mode == 0 ? stt->idx - 1 : stt->idx;
    const int nd = mode == 0 ? D : mode;
    for (int c = threadIdx.x; c < C; c += 64) {
        float a = 0.f;
        for (int i = 0; i < nd; i++) {
            int64_t k = stt->codes[((int64_t)b * HW + pos) * D + i];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            a += cb[k * C + c];
        }
        out[(int64_t)b * C + c] = __float2bfloat16(a);
    }
}
// bookkeeping: which graph just ran decides what advances
__global__ void advance_kernel(StepState* stt, int ds, int didx, int dstep) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    if (threadIdx.x == 0) { stt->s += ds; stt->idx += didx; stt->step += dstep; }
}
__global__ void __launch_bounds__(256) logits_copy_kernel(const StepState* __restrict__ stt, const float* __restrict__ lg, int d,
                                                          int64_t n) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    if (!stt->logits_out) return;
    float* dst = stt->logits_out + (int64_t)(stt->step + d) * n;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) dst[i] = lg[i];
}

__global__ void init_state_kernel(StepState* dst, StepState v) {
    if (threadIdx.x == 0) *dst = v;
}

// ------------------------------------------------------------------------------------------------ engine
struct FastLayer {
    CUtensorMap qkv, proj, fc1, fc2;
};

struct ArFast {
    rqb200_ar_config cfg;
    rqb200_ar_weights w;
    std::vector<rqb200_block_weights> body, head;
    std::vector<FastLayer> lbody, lhead;
    CUtensorMap tm_win, tm_whead, tm_cls;
    // per (workspace, B) state
    void* ws_base = nullptr;
    int B = 0;
    CUtensorMap tx_xn, tx_att, tx_h, tx_s, tx_xq;
    cudaGraphExec_t g_cond = nullptr, g_code = nullptr, g_head = nullptr;
    int64_t n_nodes[3] = {0, 0, 0};      // kernels recorded in each graph (for the launch counter)
    cudaStream_t cap_stream = nullptr;   // capture never happens on the caller's stream (it may be the legacy default stream)
    bool use_graph = true, use_pdl = true;
    int split_qkv = 4, split_proj = 12, split_fc1 = 1, split_fc2 = 12;
    // persistent ("mega") form: phase programs living in the workspace
    bool want_mega = false, use_mega = false;   // opt-in (RQB200_MEGA=1): measured slower than the PDL chain so far, see DESIGN.md
    int mega_split_fc1 = 3;
    // cluster (DSMEM) split-K for proj / fc1 / fc2: the GEMM itself emits x += ..., h = gelu(...) -- no partial round trip
    bool w_tiled = false;        // weights packed tile-major by the host binding (cfg.weight_layout)
    int skip = 0;                // diagnostics only (RQB200_SKIP bitmask): drop a kernel type from the chain to measure its in-situ cost
    bool cluster = false;        // all of proj/fc1/fc2 (measured slower than split-K partials + fused LN reduction: 289 vs 243 ms)
    bool fc1_cluster = false;    // fc1 only (RQB200_FC1_CLUSTER=2|3|4): also slower (311-319 ms) -- cluster launches cost more than they save here
    int cl_proj = 8, cl_fc1 = 2, cl_fc2 = 8;
    // RQB200_GR=1 (experiment for the next round): proj / fc1 / fc2 reduce their split-K partials inside the GEMM (GT_GR), so a block
    // is ln, qkv, attn, proj, ln, fc1, fc2 (7 launches); with LayerNorm folded into the weights (rqb200_block_weights.cqkv / .c1)
    // it is qkv, attn, proj, fc1, fc2 (5 launches).  Needs every GT_GR grid co-resident and the GPU to itself.
    bool gr = false, fold = false;
    mutable int ctr_next = 0;    // next free arrival counter while a chain is being recorded
    int n_sm = 148;
    MegaParams prog_cond = {}, prog_code = {}, prog_head[8] = {};
};

struct FastWs {
    StepState* state;
    MPhase* tables;
    unsigned* bar;
    long long* trace;
    float *XB, *XH, *P, *LOGITS;
    __nv_bfloat16 *XN, *ATT, *Hh, *S;
    __nv_bfloat16 *kc_body, *vc_body, *kc_head, *vc_head;
    // RQB200_GR chain
    unsigned* ctr;               // one arrival counter per output tile of every GT_GR GEMM of a graph
    int ctr_cap;
    __nv_bfloat16* XQ;           // bf16 copy of the residual rows (folded LayerNorm: the qkv / fc1 operand)
    float2* ST;                  // [B][E/128] tile statistics of the residual rows
};

static int pick_split(int n_tiles, int nkb, int want) {
    int s = want > 0 ? want : 148 / n_tiles;
    if (s < 1) s = 1;
    if (s > nkb) s = nkb;
    return s;
}

static size_t fast_layout(const ArFast& f, int B, void* base, size_t cap, FastWs* ws) {
    const rqb200_ar_config& c = f.cfg;
    Arena a(base, cap);
    const int64_t E = c.embed_dim, HW = (int64_t)c.H * c.W, Tb = c.cond_len + HW;
    FastWs w;
    w.state = a.take<StepState>(1);
    w.bar = a.take<unsigned>(64);
    w.trace = a.take<long long>(4096);
    w.tables = a.take<MPhase>((size_t)(8 * c.n_body) * 2 + 2 + (size_t)c.D * (4 + 8 * c.n_head_layers) + 8);
    w.XB = a.take<float>(B * E);
    w.XH = a.take<float>(B * E);
    int maxs = std::max(std::max(f.split_qkv * 3, f.split_proj), std::max(f.split_fc2, std::max(f.split_fc1, f.mega_split_fc1) * 4));
    w.P = a.take<float>((int64_t)maxs * B * E);
    w.LOGITS = a.take<float>((int64_t)B * c.vocab);
    w.XN = a.take<__nv_bfloat16>(B * E);
    w.ATT = a.take<__nv_bfloat16>(B * E);
    w.Hh = a.take<__nv_bfloat16>(B * 4 * E);
    w.S = a.take<__nv_bfloat16>((int64_t)B * c.code_dim);
    w.ctr_cap = (int)(std::max<int64_t>(c.n_body, (int64_t)c.D * c.n_head_layers) * (6 * E / 128));
    w.ctr = a.take<unsigned>(w.ctr_cap);
    w.XQ = a.take<__nv_bfloat16>(B * E);
    w.ST = a.take<float2>(B * (E / 128));
    const int64_t per_body = (int64_t)B * c.n_head * Tb * 64, per_head = (int64_t)B * c.n_head * c.D * 64;
    w.kc_body = a.take<__nv_bfloat16>(per_body * c.n_body);
    w.vc_body = a.take<__nv_bfloat16>(per_body * c.n_body);
    w.kc_head = a.take<__nv_bfloat16>(per_head * c.n_head_layers);
    w.vc_head = a.take<__nv_bfloat16>(per_head * c.n_head_layers);
    if (ws) *ws = w;
    return a.off + 256;
}

static int gemm(const ArFast& f, const CUtensorMap& tw, const CUtensorMap& tx, int N_out, int K, int B, int splits, int mode,
                const float* bias, float bias_scale, void* out, float* partial, const float* residual, int64_t ld_res,
                const int* res_row_ptr, int64_t res_row_stride, cudaStream_t st) {
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = B; p.splits = splits; p.mode = mode;
    p.bias = bias; p.bias_scale = bias_scale; p.out = out; p.ld_out = N_out; p.partial = partial;
    p.residual = residual; p.ld_res = ld_res; p.res_row_ptr = res_row_ptr; p.res_row_stride = res_row_stride;
    p.w_tiled = f.w_tiled ? 1 : 0;
    return launch_gemm_tc(tw, tx, p, f.use_pdl, st);
}

// split-K GEMM with the in-kernel group reduction (GT_GR).  kind 0: out f32 = sum + bias + residual (+ bf16 copy + tile statistics);
// kind 1: out bf16 = gelu(sum + bias), optionally with the folded LayerNorm of the input rows (stats_in, fold_c).
static int gemm_gr(const ArFast& f, FastWs& ws, const CUtensorMap& tw, const CUtensorMap& tx, int N_out, int K, int B, int splits,
                   int kind, const float* bias, void* out, int64_t ld_out, const float* residual, void* out_bf16, float2* stats_out,
                   const float2* stats_in, const float* fold_c, cudaStream_t st) {
    const int tiles = N_out / 128;
    if (f.ctr_next + tiles > ws.ctr_cap) return fail(RQB200_EINVAL, "ar fast tier: out of GT_GR arrival counters");
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = B; p.splits = splits; p.mode = GT_GR;
    p.bias = bias; p.bias_scale = 1.f; p.out = out; p.ld_out = ld_out;
    p.residual = residual; p.ld_res = ld_out;
    p.w_tiled = f.w_tiled ? 1 : 0;
    p.gr_scratch = ws.P; p.gr_counter = ws.ctr + f.ctr_next; p.gr_kind = kind;
    p.gr_out_bf16 = out_bf16; p.gr_stats_out = stats_out; p.gr_stats_in = stats_in; p.gr_fold_c = fold_c;
    f.ctr_next += tiles;
    return launch_gemm_tc(tw, tx, p, f.use_pdl, st);
}

// one transformer stack on the single new token of every batch row; x lives in `x` (fp32), residual additions are
// deferred into the next ln_reduce.  On return the LAST block's fc2 partials (+ its bias) are still pending.
static int fast_stack(const ArFast& f, const std::vector<rqb200_block_weights>& blocks, const std::vector<FastLayer>& maps,
                      FastWs& ws, float* x, bool first_has_pending, const float* pending_bias, const float* pending_extra,
                      const float* x_src, __nv_bfloat16* kc, __nv_bfloat16* vc, int Tmax, const int* t_ptr, int t_host,
                      cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const int E = c.embed_dim, B = f.B;
    const int64_t per = (int64_t)B * c.n_head * Tmax * 64;
    for (size_t l = 0; l < blocks.size(); l++) {
        const rqb200_block_weights& bw = blocks[l];
        if (f.gr) {
            // ---- group-reduce form: x is final after proj / fc2; nothing is ever pending
            const bool first = l == 0;
            const bool need_copy = first && (x_src != x || pending_extra != nullptr);
            const int nt = E / 128;
            if (f.fold) {
                if (first)
                    RQB_TRY(launch_pdl(row_prep_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)x_src,
                                       (const float*)pending_extra, (float*)(need_copy ? x : nullptr), ws.XQ, ws.ST, B, E));
                RQB_TRY(gemm(f, maps[l].qkv, f.tx_xq, 3 * E, E, B, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                             nullptr, 0, st));
                RQB_TRY(launch_pdl(attn_fast_kernel, dim3((unsigned)ceil_div(B * c.n_head, 4)), dim3(128), 0, st, f.use_pdl,
                                   (const float*)ws.P, f.split_qkv, (const float*)bw.bqkv, kc + per * l, vc + per * l, ws.ATT, B, E,
                                   c.n_head, Tmax, t_ptr, t_host, (const float2*)ws.ST, nt, (const float*)bw.cqkv));
            } else {
                RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)(first ? x_src : x),
                                   (const float*)nullptr, 0, (const float*)nullptr, (const float*)(first ? pending_extra : nullptr),
                                   (float*)(need_copy ? x : nullptr), (const float*)bw.ln1_w, (const float*)bw.ln1_b, ws.XN, B, E));
                RQB_TRY(gemm(f, maps[l].qkv, f.tx_xn, 3 * E, E, B, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                             nullptr, 0, st));
                RQB_TRY(launch_pdl(attn_fast_kernel, dim3((unsigned)ceil_div(B * c.n_head, 4)), dim3(128), 0, st, f.use_pdl,
                                   (const float*)ws.P, f.split_qkv, (const float*)bw.bqkv, kc + per * l, vc + per * l, ws.ATT, B, E,
                                   c.n_head, Tmax, t_ptr, t_host, (const float2*)nullptr, 0, (const float*)nullptr));
            }
            RQB_TRY(gemm_gr(f, ws, maps[l].proj, f.tx_att, E, E, B, f.split_proj, 0, bw.bproj, x, E, x, f.fold ? ws.XQ : nullptr,
                            f.fold ? ws.ST : nullptr, nullptr, nullptr, st));
            if (!f.fold)
                RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)x,
                                   (const float*)nullptr, 0, (const float*)nullptr, (const float*)nullptr, (float*)nullptr,
                                   (const float*)bw.ln2_w, (const float*)bw.ln2_b, ws.XN, B, E));
            RQB_TRY(gemm_gr(f, ws, maps[l].fc1, f.fold ? f.tx_xq : f.tx_xn, 4 * E, E, B, f.split_fc1, 1, bw.b1, ws.Hh, 4 * E, nullptr,
                            nullptr, nullptr, f.fold ? ws.ST : nullptr, f.fold ? bw.c1 : nullptr, st));
            RQB_TRY(gemm_gr(f, ws, maps[l].fc2, f.tx_h, E, 4 * E, B, f.split_fc2, 0, bw.b2, x, E, x, f.fold ? ws.XQ : nullptr,
                            f.fold ? ws.ST : nullptr, nullptr, nullptr, st));
            continue;
        }
        if (f.cluster) {
            // ---- cluster split-K form: 6 kernels per block, no partial buffers except for qkv
            const bool first = l == 0;
            const bool need_copy = first && (x_src != x || pending_extra != nullptr);
            RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)(first ? x_src : x),
                               (const float*)nullptr, 0, (const float*)nullptr, (const float*)(first ? pending_extra : nullptr),
                               (float*)(need_copy ? x : nullptr), (const float*)bw.ln1_w, (const float*)bw.ln1_b, ws.XN, B, E));
            RQB_TRY(gemm(f, maps[l].qkv, f.tx_xn, 3 * E, E, B, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                         nullptr, 0, st));
            RQB_TRY(launch_pdl(attn_fast_kernel, dim3((unsigned)ceil_div(B * c.n_head, 4)), dim3(128), 0, st, f.use_pdl,
                               (const float*)ws.P, f.split_qkv, (const float*)bw.bqkv, kc + per * l, vc + per * l, ws.ATT, B, E,
                               c.n_head, Tmax, t_ptr, t_host, (const float2*)nullptr, 0, (const float*)nullptr));
            RQB_TRY(gemm(f, maps[l].proj, f.tx_att, E, E, B, f.cl_proj, GT_F32, bw.bproj, 1.f, x, nullptr, x, E, nullptr, 0, st));
            RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)x,
                               (const float*)nullptr, 0, (const float*)nullptr, (const float*)nullptr, (float*)nullptr,
                               (const float*)bw.ln2_w, (const float*)bw.ln2_b, ws.XN, B, E));
            RQB_TRY(gemm(f, maps[l].fc1, f.tx_xn, 4 * E, E, B, f.cl_fc1, GT_BF16_GELU, bw.b1, 1.f, ws.Hh, nullptr, nullptr, 0, nullptr,
                         0, st));
            RQB_TRY(gemm(f, maps[l].fc2, f.tx_h, E, 4 * E, B, f.cl_fc2, GT_F32, bw.b2, 1.f, x, nullptr, x, E, nullptr, 0, st));
            continue;
        }
        // LN1 (+ pending fc2 reduction of the previous block / previous stack)
        const bool pend = l > 0 || first_has_pending;
        const float* pb = l > 0 ? blocks[l - 1].b2 : pending_bias;
        if (!(f.skip & 1)) RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl,
                           (const float*)(l == 0 ? x_src : x), (const float*)(pend ? ws.P : nullptr), pend ? f.split_fc2 : 0,
                           (const float*)(pend ? pb : nullptr), (const float*)(l == 0 ? pending_extra : nullptr), x,
                           (const float*)bw.ln1_w, (const float*)bw.ln1_b, ws.XN, B, E));
        if (!(f.skip & 2)) RQB_TRY(gemm(f, maps[l].qkv, f.tx_xn, 3 * E, E, B, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                     nullptr, 0, st));
        if (!(f.skip & 4)) RQB_TRY(launch_pdl(attn_fast_kernel, dim3((unsigned)ceil_div(B * c.n_head, 4)), dim3(128), 0, st, f.use_pdl,
                           (const float*)ws.P, f.split_qkv, (const float*)bw.bqkv, kc + per * l, vc + per * l, ws.ATT, B, E,
                           c.n_head, Tmax, t_ptr, t_host, (const float2*)nullptr, 0, (const float*)nullptr));
        if (!(f.skip & 8)) RQB_TRY(gemm(f, maps[l].proj, f.tx_att, E, E, B, f.split_proj, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                     nullptr, 0, st));
        if (!(f.skip & 1)) RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)x,
                           (const float*)ws.P, f.split_proj, (const float*)bw.bproj, (const float*)nullptr, x,
                           (const float*)bw.ln2_w, (const float*)bw.ln2_b, ws.XN, B, E));
        if (f.split_fc1 == 1) {
            // fc1: either one CTA per 128-feature tile (48 CTAs at E=1536) or a 2-CTA cluster per tile with DSMEM reduction
            if (!(f.skip & 16)) RQB_TRY(gemm(f, maps[l].fc1, f.tx_xn, 4 * E, E, B, f.fc1_cluster ? f.cl_fc1 : 1, GT_BF16_GELU, bw.b1, 1.f, ws.Hh, nullptr,
                         nullptr, 0, nullptr, 0, st));
        } else {
            RQB_TRY(gemm(f, maps[l].fc1, f.tx_xn, 4 * E, E, B, f.split_fc1, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                         nullptr, 0, st));
            RQB_TRY(launch_pdl(act_reduce_kernel, dim3((unsigned)std::min<int64_t>(ceil_div((int64_t)B * 4 * E / 4, 256), 1184)), dim3(256), 0, st, f.use_pdl, (const float*)ws.P, f.split_fc1,
                               (const float*)bw.b1, ws.Hh, B, 4 * E));
        }
        if (!(f.skip & 32)) RQB_TRY(gemm(f, maps[l].fc2, f.tx_h, E, 4 * E, B, f.split_fc2, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                     nullptr, 0, st));
    }
    return 0;
}

// ---- persistent form: the same chains as fast_stack/record_*, expressed as phase tables for ar_mega_kernel
static int build_programs(ArFast& f, FastWs& ws) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, Tb = c.cond_len + HW, D = c.D, V = c.vocab;
    CUtensorMap mx_xn, mx_att, mx_h, mx_s;
    RQB_TRY(make_tmap_2d(&mx_xn, ws.XN, 1, E, B, (uint64_t)E * 2, 64, 64));
    RQB_TRY(make_tmap_2d(&mx_att, ws.ATT, 1, E, B, (uint64_t)E * 2, 64, 64));
    RQB_TRY(make_tmap_2d(&mx_h, ws.Hh, 1, 4 * E, B, (uint64_t)E * 8, 64, 64));
    RQB_TRY(make_tmap_2d(&mx_s, ws.S, 1, c.code_dim, B, (uint64_t)c.code_dim * 2, 64, 64));
    std::vector<MPhase> all;
    auto ln = [&](const float* x_in, bool pend, int S, const float* bias, const float* extra, float* x_out, const float* g,
                  const float* be, __nv_bfloat16* xn) {
        MPhase p = {};
        p.type = MP_LN; p.x_in = x_in; p.partial = pend ? ws.P : nullptr; p.S = pend ? S : 0; p.bias = pend ? bias : nullptr;
        p.extra = extra; p.x_out = x_out; p.g = g; p.be = be; p.xn = xn;
        all.push_back(p);
    };
    auto gm = [&](const CUtensorMap& tw, const CUtensorMap& tx, int N_out, int K, int splits, int mode, const float* bias,
                  float bias_scale, void* out, const float* res, int64_t ld_res, const int* res_row_ptr, int64_t res_row_stride) {
        MPhase p = {};
        p.type = MP_GEMM; p.tmW = tw; p.tmX = tx; p.N_out = N_out; p.K = K; p.splits = splits; p.mode = mode; p.gbias = bias;
        p.w_tiled = f.w_tiled ? 1 : 0;
        p.bias_scale = bias_scale; p.out = out; p.gpartial = ws.P; p.res = res; p.ld_res = ld_res; p.res_row_ptr = res_row_ptr;
        p.res_row_stride = res_row_stride;
        all.push_back(p);
    };
    auto stack = [&](const std::vector<rqb200_block_weights>& blocks, const std::vector<FastLayer>& maps, float* x, bool first_pending,
                     const float* pending_bias, const float* pending_extra, const float* x_src, __nv_bfloat16* kc,
                     __nv_bfloat16* vc, int Tmax, const int* t_ptr, int t_host) {
        const int64_t per = (int64_t)B * c.n_head * Tmax * 64;
        for (size_t l = 0; l < blocks.size(); l++) {
            const rqb200_block_weights& bw = blocks[l];
            const bool pend = l > 0 || first_pending;
            ln(l == 0 ? x_src : x, pend, f.split_fc2, l > 0 ? blocks[l - 1].b2 : pending_bias, l == 0 ? pending_extra : nullptr, x,
               bw.ln1_w, bw.ln1_b, ws.XN);
            gm(maps[l].qkv, mx_xn, 3 * E, E, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, nullptr, 0, nullptr, 0);
            MPhase a = {};
            a.type = MP_ATTN; a.apart = ws.P; a.aS = f.split_qkv; a.bqkv = bw.bqkv; a.kc = kc + per * l; a.vc = vc + per * l;
            a.att = ws.ATT; a.Tmax = Tmax; a.t_ptr = t_ptr; a.t_host = t_host;
            all.push_back(a);
            gm(maps[l].proj, mx_att, E, E, f.split_proj, GT_PARTIAL, nullptr, 1.f, nullptr, nullptr, 0, nullptr, 0);
            ln(x, true, f.split_proj, bw.bproj, nullptr, x, bw.ln2_w, bw.ln2_b, ws.XN);
            if (f.mega_split_fc1 <= 1) {
                gm(maps[l].fc1, mx_xn, 4 * E, E, 1, GT_BF16_GELU, bw.b1, 1.f, ws.Hh, nullptr, 0, nullptr, 0);
            } else {
                gm(maps[l].fc1, mx_xn, 4 * E, E, f.mega_split_fc1, GT_PARTIAL, nullptr, 1.f, nullptr, nullptr, 0, nullptr, 0);
                MPhase a2 = {};
                a2.type = MP_ACT; a2.N_out = 4 * E; a2.splits = f.mega_split_fc1; a2.gbias = bw.b1; a2.gpartial = ws.P; a2.out = ws.Hh;
                all.push_back(a2);
            }
            gm(maps[l].fc2, mx_h, E, 4 * E, f.split_fc2, GT_PARTIAL, nullptr, 1.f, nullptr, nullptr, 0, nullptr, 0);
        }
    };
    auto cs = [&](int mode) {
        MPhase p = {};
        p.type = MP_CODESUM; p.cs_mode = mode; p.cs_out = ws.S;
        all.push_back(p);
    };
    std::vector<std::pair<size_t, size_t>> spans;      // [begin, end) per program: cond, code, head 0..D-1
    size_t b0 = all.size();
    stack(f.body, f.lbody, ws.XB, false, nullptr, nullptr, ws.XB, ws.kc_body, ws.vc_body, Tb, &ws.state->s, 0);
    spans.push_back({b0, all.size()});
    b0 = all.size();
    cs(0);
    gm(f.tm_win, mx_s, E, c.code_dim, 1, GT_F32, w.b_in, (float)D, ws.XB, w.pos_emb_hw - E, 0, &ws.state->idx, E);
    stack(f.body, f.lbody, ws.XB, false, nullptr, nullptr, ws.XB, ws.kc_body, ws.vc_body, Tb, &ws.state->s, 0);
    spans.push_back({b0, all.size()});
    for (int d = 0; d < D; d++) {
        b0 = all.size();
        if (d == 0) {
            stack(f.head, f.lhead, ws.XH, true, f.body.back().b2, w.pos_emb_d, ws.XB, ws.kc_head, ws.vc_head, D, nullptr, 0);
        } else {
            cs(d);
            gm(f.tm_whead, mx_s, E, c.code_dim, 1, GT_F32, w.b_head, 1.f, ws.XH, w.pos_emb_d + (int64_t)d * E, 0, nullptr, 0);
            stack(f.head, f.lhead, ws.XH, false, nullptr, nullptr, ws.XH, ws.kc_head, ws.vc_head, D, nullptr, d);
        }
        ln(ws.XH, true, f.split_fc2, f.head.back().b2, nullptr, nullptr, w.cls_ln_w, w.cls_ln_b, ws.XN);
        gm(f.tm_cls, mx_xn, V, E, 1, GT_F32, w.b_cls, 1.f, ws.LOGITS, nullptr, 0, nullptr, 0);
        spans.push_back({b0, all.size()});
    }
    // one synchronous upload per (workspace, batch) binding
    RQB_CUDA(cudaMemcpy(ws.tables, all.data(), all.size() * sizeof(MPhase), cudaMemcpyHostToDevice));
    RQB_CUDA(cudaMemset(ws.bar, 0, 64 * sizeof(unsigned)));
    auto mk = [&](std::pair<size_t, size_t> sp) {
        MegaParams P = {};
        P.phases = ws.tables + sp.first; P.n_phases = (int)(sp.second - sp.first);
        P.B = B; P.E = E; P.nh = c.n_head; P.bar = ws.bar; P.stt = ws.state;
        P.codebook = w.codebook; P.HW = HW; P.D = D; P.Kc = c.codebook_size; P.C = c.code_dim;
        P.trace = (getenv("RQB200_MEGA_TRACE") && P.n_phases < 2000) ? ws.trace : nullptr;
        return P;
    };
    f.prog_cond = mk(spans[0]);
    f.prog_code = mk(spans[1]);
    for (int d = 0; d < D; d++) f.prog_head[d] = mk(spans[2 + d]);
    return 0;
}

static int record_body(ArFast& f, FastWs& ws, bool cond_token, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, Tb = c.cond_len + HW;
    if (f.use_mega) {
        if (cond_token)
            RQB_TRY(launch_pdl(cond_tok_kernel, dim3(B), dim3(256), 0, st, false, (const StepState*)ws.state, w.cond_emb,
                               w.pos_emb_cond, c.cond_len, c.vocab_cond, E, ws.XB));
        RQB_TRY(launch_ar_mega(cond_token ? f.prog_cond : f.prog_code, f.n_sm, st));
        RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, false, ws.state, 1, 0, 0));
        return 0;
    }
    f.ctr_next = 0;
    if (cond_token) {
        RQB_TRY(launch_pdl(cond_tok_kernel, dim3(B), dim3(256), 0, st, f.use_pdl, (const StepState*)ws.state, w.cond_emb,
                           w.pos_emb_cond, c.cond_len, c.vocab_cond, E, ws.XB));
    } else {
        RQB_TRY(launch_pdl(code_sum_kernel, dim3(B), dim3(64), 0, st, f.use_pdl, (const StepState*)ws.state, w.codebook, HW, c.D,
                           c.codebook_size, c.code_dim, 0, ws.S));
        // x = W_in (sum_d e_d) + D b_in + pos_emb_hw[idx-1]       (bias counted D times, transformers.py:220,225)
        RQB_TRY(gemm(f, f.tm_win, f.tx_s, E, c.code_dim, B, 1, GT_F32, w.b_in, (float)c.D, ws.XB, nullptr,
                     w.pos_emb_hw - E /* row idx-1 */, 0, &ws.state->idx, E, st));
    }
    RQB_TRY(fast_stack(f, f.body, f.lbody, ws, ws.XB, false, nullptr, nullptr, ws.XB, ws.kc_body, ws.vc_body, Tb, &ws.state->s,
                       0, st));
    if (f.gr) RQB_TRY(launch_pdl(ctr_zero_kernel, dim3(1), dim3(256), 0, st, f.use_pdl, ws.ctr, f.ctr_next));
    RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, f.use_pdl, ws.state, 1, 0, 0));
    return 0;
}

static int record_head(ArFast& f, FastWs& ws, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, D = c.D, V = c.vocab;
    if (f.use_mega) {
        for (int d = 0; d < D; d++) {
            RQB_TRY(launch_ar_mega(f.prog_head[d], f.n_sm, st));
            RQB_TRY(launch_pdl(logits_copy_kernel, dim3(64), dim3(256), 0, st, false, (const StepState*)ws.state,
                               (const float*)ws.LOGITS, d, (int64_t)B * V));
            RQB_TRY(launch_sample_dyn(ws.LOGITS, ws.state, d, B, V, HW, D, st, false));
        }
        RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, false, ws.state, 0, 1, D));
        return 0;
    }
    f.ctr_next = 0;
    for (int d = 0; d < D; d++) {
        if (d == 0) {
            // spatial ctx = body x + pending fc2 of the last body block ; token = ctx + pos_emb_d[0]  (transformers.py:259-270)
            RQB_TRY(fast_stack(f, f.head, f.lhead, ws, ws.XH, true, f.body.back().b2, w.pos_emb_d, ws.XB, ws.kc_head, ws.vc_head,
                               D, nullptr, 0, st));
        } else {
            RQB_TRY(launch_pdl(code_sum_kernel, dim3(B), dim3(64), 0, st, f.use_pdl, (const StepState*)ws.state, w.codebook, HW, D,
                               c.codebook_size, c.code_dim, d, ws.S));
            RQB_TRY(gemm(f, f.tm_whead, f.tx_s, E, c.code_dim, B, 1, GT_F32, w.b_head, 1.f, ws.XH, nullptr,
                         w.pos_emb_d + (int64_t)d * E, 0, nullptr, 0, st));
            RQB_TRY(fast_stack(f, f.head, f.lhead, ws, ws.XH, false, nullptr, nullptr, ws.XH, ws.kc_head, ws.vc_head, D, nullptr, d,
                               st));
        }
        // classifier: LN(x + pending fc2) -> logits                                              (transformers.py:278-285)
        RQB_TRY(launch_pdl(ln_reduce_kernel, dim3(B), dim3(384), (size_t)0, st, f.use_pdl, (const float*)ws.XH,
                           (const float*)((f.cluster || f.gr) ? nullptr : ws.P), (f.cluster || f.gr) ? 0 : f.split_fc2,
                           (const float*)((f.cluster || f.gr) ? nullptr : f.head.back().b2), (const float*)nullptr, (float*)nullptr, w.cls_ln_w,
                           w.cls_ln_b, ws.XN, B, E));
        RQB_TRY(gemm(f, f.tm_cls, f.tx_xn, V, E, B, 1, GT_F32, w.b_cls, 1.f, ws.LOGITS, nullptr, nullptr, 0, nullptr, 0, st));
        RQB_TRY(launch_pdl(logits_copy_kernel, dim3(64), dim3(256), 0, st, f.use_pdl, (const StepState*)ws.state,
                           (const float*)ws.LOGITS, d, (int64_t)B * V));
        RQB_TRY(launch_sample_dyn(ws.LOGITS, ws.state, d, B, V, HW, D, st, f.use_pdl));
    }
    if (f.gr) RQB_TRY(launch_pdl(ctr_zero_kernel, dim3(1), dim3(256), 0, st, f.use_pdl, ws.ctr, f.ctr_next));
    RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, f.use_pdl, ws.state, 0, 1, D));
    return 0;
}

static int capture(ArFast& f, FastWs& ws, int which, cudaGraphExec_t* out) {
    cudaGraph_t g = nullptr;
    if (!f.cap_stream) RQB_CUDA(cudaStreamCreateWithFlags(&f.cap_stream, cudaStreamNonBlocking));
    cudaStream_t st = f.cap_stream;
    RQB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int64_t before = g_launches;
    int rc = which == 0 ? record_body(f, ws, true, st) : which == 1 ? record_body(f, ws, false, st) : record_head(f, ws, st);
    f.n_nodes[which] = g_launches - before;
    g_launches = before;                   // recording is not launching
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(RQB200_ECUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(RQB200_ECUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(e));
    return 0;
}

static void drop_graphs(ArFast& f) {
    if (f.g_cond) cudaGraphExecDestroy(f.g_cond);
    if (f.g_code) cudaGraphExecDestroy(f.g_code);
    if (f.g_head) cudaGraphExecDestroy(f.g_head);
    f.g_cond = f.g_code = f.g_head = nullptr;
}

ArFast* ar_fast_create(const rqb200_ar_config& cfg, const rqb200_ar_weights& w, const rqb200_block_weights* body_p,
                       const rqb200_block_weights* head_p) {
    std::vector<rqb200_block_weights> body(body_p, body_p + cfg.n_body), head(head_p, head_p + cfg.n_head_layers);
    const int E = cfg.embed_dim;
    if (E % 128 != 0 || cfg.vocab % 128 != 0 || cfg.code_dim % 64 != 0 || cfg.D > 8 || cfg.cond_len + cfg.H * cfg.W > AF_MAXT || E > 4608) {
        set_error("ar fast tier: need E % 128 == 0, V % 128 == 0, code_dim % 64 == 0, D <= 8, cond_len + H*W <= 512, E <= 4608");
        return nullptr;
    }
    ArFast* f = new ArFast();
    f->cfg = cfg; f->w = w; f->body = body; f->head = head;
    f->w_tiled = (cfg.mode & 0x100) != 0;              // bit 8 of `mode`: fast-tier weights are tile-major
    const char* e;
    if ((e = getenv("RQB200_NO_GRAPH")) && e[0] == '1') f->use_graph = false;
    if ((e = getenv("RQB200_NO_PDL")) && e[0] == '1') f->use_pdl = false;
    if ((e = getenv("RQB200_MEGA")) && e[0] == '1') f->want_mega = true;
    if ((e = getenv("RQB200_CLUSTER")) && e[0] == '1') f->cluster = true;
    if ((e = getenv("RQB200_SKIP"))) f->skip = atoi(e);
    if ((e = getenv("RQB200_FC1_CLUSTER"))) { f->cl_fc1 = atoi(e); f->fc1_cluster = f->cl_fc1 > 1; }
    f->cl_proj = std::min(8, E / 64);
    f->cl_fc2 = std::min(8, 4 * E / 64);
    f->cl_fc1 = std::min(f->cl_fc1, E / 64);
    if ((e = getenv("RQB200_MEGA_SPLIT_FC1"))) f->mega_split_fc1 = atoi(e);
    f->mega_split_fc1 = pick_split(4 * E / 128, E / 64, f->mega_split_fc1);
    {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) f->n_sm = n;
    }
    const int nkbE = E / 64;
    f->split_qkv = pick_split(3 * E / 128, nkbE, getenv("RQB200_SPLIT_QKV") ? atoi(getenv("RQB200_SPLIT_QKV")) : 0);
    f->split_proj = pick_split(E / 128, nkbE, getenv("RQB200_SPLIT_PROJ") ? atoi(getenv("RQB200_SPLIT_PROJ")) : 0);
    f->split_fc1 = pick_split(4 * E / 128, nkbE, getenv("RQB200_SPLIT_FC1") ? atoi(getenv("RQB200_SPLIT_FC1")) : 0);
    f->split_fc2 = pick_split(E / 128, 4 * nkbE, getenv("RQB200_SPLIT_FC2") ? atoi(getenv("RQB200_SPLIT_FC2")) : 0);
    {
        bool all_fold = true, any_fold = false;
        for (const auto& b : body) { all_fold &= (b.cqkv && b.c1); any_fold |= (b.cqkv || b.c1); }
        for (const auto& b : head) { all_fold &= (b.cqkv && b.c1); any_fold |= (b.cqkv || b.c1); }
        if ((e = getenv("RQB200_GR")) && e[0] == '1') f->gr = true;
        // every CTA of a GT_GR grid spins for its tile's peers: the whole grid has to be resident (one CTA per SM)
        const int g_max = std::max(std::max(E / 128 * f->split_proj, 4 * E / 128 * f->split_fc1), E / 128 * f->split_fc2);
        if (f->gr && (g_max > f->n_sm || f->cluster || f->want_mega || E % 128 != 0)) {
            fprintf(stderr, "rqb200: RQB200_GR ignored (a GEMM grid of %d CTAs would not be co-resident on %d SMs, or another chain form is selected)\n",
                    g_max, f->n_sm);
            f->gr = false;
        }
        f->fold = f->gr && all_fold;
        if (any_fold && !f->fold) {
            set_error("ar fast tier: LayerNorm-folded weights (rqb200_block_weights.cqkv / .c1) need the RQB200_GR=1 chain and must be given for every block");
            delete f;
            return nullptr;
        }
    }
    auto mk = [&](const std::vector<rqb200_block_weights>& bl, std::vector<FastLayer>& out) -> int {
        out.resize(bl.size());
        for (size_t l = 0; l < bl.size(); l++) {
            RQB_TRY(make_tmap_weight(&out[l].qkv, bl[l].wqkv, 3 * E, E, f->w_tiled));
            RQB_TRY(make_tmap_weight(&out[l].proj, bl[l].wproj, E, E, f->w_tiled));
            RQB_TRY(make_tmap_weight(&out[l].fc1, bl[l].w1, 4 * E, E, f->w_tiled));
            RQB_TRY(make_tmap_weight(&out[l].fc2, bl[l].w2, E, 4 * E, f->w_tiled));
        }
        return 0;
    };
    int rc = mk(body, f->lbody);
    if (!rc) rc = mk(head, f->lhead);
    if (!rc) rc = make_tmap_weight(&f->tm_win, w.w_in, E, cfg.code_dim, f->w_tiled);
    if (!rc) rc = make_tmap_weight(&f->tm_whead, w.w_head, E, cfg.code_dim, f->w_tiled);
    if (!rc) rc = make_tmap_weight(&f->tm_cls, w.w_cls, cfg.vocab, E, f->w_tiled);
    if (rc) { delete f; return nullptr; }
    return f;
}

void ar_fast_destroy(ArFast* f) {
    if (!f) return;
    drop_graphs(*f);
    if (f->cap_stream) cudaStreamDestroy(f->cap_stream);
    delete f;
}

size_t ar_fast_workspace_bytes(const ArFast* f, int B) { return fast_layout(*f, B, nullptr, 0, nullptr); }

int ar_fast_sample(ArFast* f, const int64_t* partial, const int64_t* cond, int B, int start_h, int start_w, float temperature,
                   const int32_t* top_k, const float* top_p, const float* noise, int64_t noise_stride, float* logits_out,
                   const int64_t* force, int64_t* out, void* wsp, size_t ws_bytes, cudaStream_t st) {
    const rqb200_ar_config& c = f->cfg;
    const int E = c.embed_dim, D = c.D, HW = c.H * c.W, cl = c.cond_len;
    if (B < 1 || B > 256) return fail(RQB200_EINVAL, "ar fast tier: batch must be in [1,256] per call");
    if (start_h < 0 || start_w < 0 || start_w >= c.W || start_h > c.H) return fail(RQB200_EINVAL, "ar_sample: bad start_loc");
    FastWs ws;
    size_t need = fast_layout(*f, B, wsp, ws_bytes, &ws);
    if (need > ws_bytes) return fail(RQB200_EWORKSPACE, "ar_sample: workspace too small");
    if (out != partial)
        RQB_CUDA(cudaMemcpyAsync(out, partial, (size_t)B * HW * D * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    const int idx0 = start_h * c.W + start_w;
    if (idx0 >= HW) return 0;
    if (f->ws_base != wsp || f->B != B) {        // (re)bind activation tensor maps + graphs to this workspace
        drop_graphs(*f);
        f->ws_base = wsp;
        f->B = B;
        const int bn = gemm_tc_bn(B);
        RQB_TRY(make_tmap_2d(&f->tx_xn, ws.XN, 1, E, B, (uint64_t)E * 2, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_att, ws.ATT, 1, E, B, (uint64_t)E * 2, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_h, ws.Hh, 1, 4 * E, B, (uint64_t)E * 8, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_s, ws.S, 1, c.code_dim, B, (uint64_t)c.code_dim * 2, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_xq, ws.XQ, 1, E, B, (uint64_t)E * 2, 64, bn));
        f->use_mega = f->want_mega && B <= 64 && E <= 224 * 4 * 3 && mega_smem_bytes(E) <= 227 * 1024;
        if (f->use_mega) RQB_TRY(build_programs(*f, ws));
    }
    StepState h = {};
    h.s = 0; h.idx = 0; h.step = 0;
    h.cond = cond; h.codes = out; h.force = force; h.noise = noise; h.logits_out = logits_out; h.noise_stride = noise_stride;
    h.temperature = temperature;
    for (int d = 0; d < D; d++) { h.top_k[d] = top_k[d]; h.top_p[d] = top_p[d]; }
    RQB_TRY(launch_pdl(init_state_kernel, dim3(1), dim3(32), 0, st, false, ws.state, h));
    if (f->gr) RQB_CUDA(cudaMemsetAsync(ws.ctr, 0, (size_t)ws.ctr_cap * sizeof(unsigned), st));
    auto run = [&](int which, cudaGraphExec_t* g) -> int {
        if (!f->use_graph) return which == 0 ? record_body(*f, ws, true, st) : which == 1 ? record_body(*f, ws, false, st)
                                                                                        : record_head(*f, ws, st);
        if (!*g) RQB_TRY(capture(*f, ws, which, g));
        RQB_CUDA(cudaGraphLaunch(*g, st));
        g_launches += f->n_nodes[which];     // kernels executed by this replay
        return 0;
    };
    // prefill: cond tokens, then (resume) the code tokens of positions < idx0, one cached step each -- causal, so
    // identical to the reference's batched prefill (transformers.py:237-239)
    for (int s = 0; s < cl; s++) RQB_TRY(run(0, &f->g_cond));
    // state.idx must equal (position whose codes feed the body) + 1 while replaying the code-token graph
    for (int j = 1; j <= idx0; j++) {
        RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), 0, st, false, ws.state, 0, 1, 0));
        RQB_TRY(run(1, &f->g_code));
    }
    for (int idx = idx0; idx < HW; idx++) {
        if (idx > idx0) RQB_TRY(run(1, &f->g_code));       // body step on the token of position idx-1 (state.idx == idx)
        RQB_TRY(run(2, &f->g_head));                       // D head steps + sampling; advances idx, step
    }
    if (f->use_mega && getenv("RQB200_MEGA_TRACE")) {      // diagnostics: per-phase time of the LAST launch that wrote the trace
        cudaStreamSynchronize(st);
        std::vector<long long> tr(4096);
        std::vector<MPhase> ph(f->prog_head[D - 1].n_phases);
        cudaMemcpy(tr.data(), ws.trace, 4096 * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(ph.data(), f->prog_head[D - 1].phases, ph.size() * sizeof(MPhase), cudaMemcpyDeviceToHost);
        double sum[5] = {0, 0, 0, 0, 0};
        int cnt[5] = {0, 0, 0, 0, 0};
        double wsum[5] = {0, 0, 0, 0, 0};
        for (size_t i = 0; i < ph.size(); i++) {
            sum[ph[i].type] += (double)(tr[2 * i + 2] - tr[2 * i]);
            wsum[ph[i].type] += (double)(tr[2 * i + 1] - tr[2 * i]);
            cnt[ph[i].type]++;
        }
        const char* nm[5] = {"LN", "GEMM", "ATTN", "CODESUM", "ACT"};
        for (int k = 0; k < 5; k++)
            if (cnt[k]) fprintf(stderr, "[mega trace] %-8s n=%3d avg %.2f us (CTA0 own work %.2f us, barrier wait %.2f us)\n", nm[k], cnt[k],
                                sum[k] / cnt[k] / 1e3, wsum[k] / cnt[k] / 1e3, (sum[k] - wsum[k]) / cnt[k] / 1e3);
        for (size_t i = 0; i < ph.size() && i < 16; i++)
            fprintf(stderr, "[mega trace] phase %2zu type %d N_out %5d K %5d splits %2d : %.2f us (work %.2f)\n", i, ph[i].type, ph[i].N_out,
                    ph[i].K, ph[i].splits, (double)(tr[2 * i + 2] - tr[2 * i]) / 1e3, (double)(tr[2 * i + 1] - tr[2 * i]) / 1e3);
    }
    return 0;
}

}  // namespace rqb
