// P3 "fast" tier -- the cached AR step as a PDL-chained, CUDA-graph-replayed sequence of sm_100a kernels.
//
// Same semantics as ar_engine.cu's exact tier (reference: transformers.py:190-369, attentions.py:60-142), different
// arithmetic class: 16-bit weights / activations / KV cache on tcgen05 (gemm_tc.cu) -- fp16 by default, the reference's own
// autocast class (transformers.py:114,206; main_sampling_fid.py:216), bf16 on request -- fp32 accumulation, fp32 residual
// stream, fp32 LayerNorm / softmax / sampler.
//
// One transformer block on the single new token of every batch row (M = batch rows):
//
//     ln_reduce  x += bias_prev + sum_s partial_prev[s] ; xn = LN(x)     <- fused split-K reduction + residual
//     gemm_tc    qkv partials = Wqkv . xn                                 (split-K, 144 CTAs)
//     attn_fast  q,k,v = sum partials + bias ; append k,v to the cache ; softmax(q k^T/8) v -> att
//     gemm_tc    proj partials = Wproj . att
//     ln_reduce  x += bproj + sum partials ; xn = LN2(x)
//     gemm_tc    fc1 partials ; act_reduce h = gelu(sum + b1)
//     gemm_tc    fc2 partials = W2 . h
//
// Every launch is one all-to-all exchange between the SMs (the step is bound by the latency of these
// dependent exchanges, not by HBM).  Forms with fewer LAUNCHES but the same number of EXCHANGES -- a persistent megakernel with
// grid barriers (round 1), split-K reduced inside the GEMM behind an arrival counter with LayerNorm folded into the weights
// (round 2: 5 launches per block) -- were built, parity-tested and measured slower (258 and 249-256 ms against 194 ms per 64
// images); they are not kept in the tree.
//
// Every kernel starts with griddepcontrol.launch_dependents and reads upstream data only after griddepcontrol.wait, so
// the NEXT kernel's prologue -- for the GEMMs: filling the shared-memory ring with weight tiles -- overlaps this one.
// Position-dependent scalars (sequence index, spatial index, token counter) live in a device-side StepState that the
// last kernel of each graph advances, so a handful of captured graphs (cond-token body step, code-token body step, head
// steps + sampling) are replayed for all positions without host involvement.
//
// The prefix (cond tokens, and on a start_loc resume the code tokens before it) is prefilled in ONE pass of M = B*T row
// GEMMs + a causal attention kernel that writes the KV cache (reference: transformers.py:237-239, attentions.py:60-104 with
// Tnew > 1); the token-by-token replay of the single-step graph remains available (flag) and is the prefill's oracle.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace rqb {

// diagnostic stage trace: 4 globaltimer stamps per launch written by CTA 0 (entry, dependency resolved, mid, done)
#define TR_IN(tr)  do { if ((tr) != nullptr && blockIdx.x == 0 && threadIdx.x == 0) (tr)[0] = tc::gtimer(); } while (0)
#define TR_DEP(tr) do { if ((tr) != nullptr && blockIdx.x == 0 && threadIdx.x == 0) (tr)[1] = tc::gtimer(); } while (0)
#define TR_OUT(tr) do { if ((tr) != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { (tr)[2] = tc::gtimer(); (tr)[3] = (tr)[2]; } } while (0)

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
// The per-block small vectors (biases, LayerNorm parameters: ~80 KB per block, read once per token) are DRAM misses -- 2.8 GB of
// weights and the KV cache pass through L2 between two uses -- and each sits on a stage's critical path (a shared copy for all
// blocks made the step 2 % faster, profiles/dropped_r2/shared_params_r2.txt).  LN1 of block l therefore asks L2 for block l+1's.
struct PrefetchList {
    const void* p[8];
    uint32_t bytes[8];
    int n;
};

// x_out = x_in + bias + sum_s partial[s] (+ extra row) ; xn = LayerNorm(x_out) in 16-bit.  One CTA per row.
// (instantiated as <384, 3> only: 192-thread CTAs with two chunks per thread, padded grids and a 2-CTA-cluster form all measured
//  equal or slower, profiles/dropped_r2/ln_grid_threads_r2.txt)
template <int THREADS, int NCH>
__global__ void __launch_bounds__(THREADS)
ln_reduce_kernel(const float* __restrict__ x_in, const float* __restrict__ partial, int S, const float* __restrict__ bias,
                 const float* __restrict__ extra, float* __restrict__ x_out, const float* __restrict__ g,
                 const float* __restrict__ be, h16* __restrict__ xn, int B, int E, int bf, long long* tr, PrefetchList pf) {
    // each thread owns up to NCH float4 chunks of the row (E <= THREADS*4*NCH); every load is issued before the first dependent add and
    // the row stays in registers between the statistics and the normalisation
    __shared__ float red[33];
    tc::pdl_launch_dependents();
    TR_IN(tr);
    if (blockIdx.x == 0 && threadIdx.x < pf.n) tc::bulk_prefetch_l2(pf.p[threadIdx.x], pf.bytes[threadIdx.x]);   // (parameters: no dependency)
    tc::pdl_wait();
    TR_DEP(tr);
    const int b = blockIdx.x;
    const int E4 = E >> 2;
    const int S12 = S < 12 ? S : 12;
    float4 v[NCH], gg[NCH], bb[NCH];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        const int e4 = threadIdx.x + k * THREADS;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e4 < E4) {
            float4 pr[12];
            if (xn) { gg[k] = reinterpret_cast<const float4*>(g)[e4]; bb[k] = reinterpret_cast<const float4*>(be)[e4]; }
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
    if (xn) {
        const float mean = block_sum(s, red) / (float)E;
        if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[2] = tc::gtimer();      // every load has landed, first reduction done
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; k++)
            if (threadIdx.x + k * THREADS < E4) {
                const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
                q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
            }
        const float rstd = rsqrtf(block_sum(q, red) / (float)E + 1e-5f);
#pragma unroll
        for (int k = 0; k < NCH; k++) {
            const int e4 = threadIdx.x + k * THREADS;
            if (e4 < E4) {
                uint2 pk;
                pk.x = pack_h16x2((v[k].x - mean) * rstd * gg[k].x + bb[k].x, (v[k].y - mean) * rstd * gg[k].y + bb[k].y, bf);
                pk.y = pack_h16x2((v[k].z - mean) * rstd * gg[k].z + bb[k].z, (v[k].w - mean) * rstd * gg[k].w + bb[k].w, bf);
                reinterpret_cast<uint2*>(xn + (int64_t)b * E)[e4] = pk;
            }
        }
        if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[3] = tc::gtimer();
    } else {
        TR_OUT(tr);
    }
}

// The batched passes' LayerNorm (prefill / teacher-forced forward: thousands of rows, no split-K partials): one WARP per row, the
// row in registers between the statistics and the normalisation, rows grid-strided over 8-warp CTAs.  (ln_reduce_kernel's
// one-384-thread-CTA-per-row form is built for 64 rows on 64 SMs; on 4096+ rows it ran at a tenth of the HBM rate.)
// x_out (nullable) = x_in (+ extra row); xn (nullable) = LayerNorm(x) in 16-bit.  NV = float4 chunks per lane (E <= 128 * NV).
template <int NV>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x_in, const float* __restrict__ extra, float* __restrict__ x_out, const float* __restrict__ g,
               const float* __restrict__ be, h16* __restrict__ xn, int64_t M, int E, int bf) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int lane = threadIdx.x & 31;
    const int E4 = E >> 2;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += (int64_t)gridDim.x * 8) {
        const float4* xr = reinterpret_cast<const float4*>(x_in + row * E);
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int e4 = lane + 32 * k;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e4 < E4) v[k] = xr[e4];
        }
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int e4 = lane + 32 * k;
            if (e4 < E4) {
                if (extra) { const float4 t = reinterpret_cast<const float4*>(extra)[e4]; v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w; }
                if (x_out) reinterpret_cast<float4*>(x_out + row * E)[e4] = v[k];
                s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            }
        }
        if (xn) {
            const float mean = warp_sum(s) / (float)E;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < NV; k++)
                if (lane + 32 * k < E4) {
                    const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
                    q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
                }
            const float rstd = rsqrtf(warp_sum(q) / (float)E + 1e-5f);
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const int e4 = lane + 32 * k;
                if (e4 < E4) {
                    const float4 gg = reinterpret_cast<const float4*>(g)[e4], bb = reinterpret_cast<const float4*>(be)[e4];
                    uint2 pk;
                    pk.x = pack_h16x2((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y, bf);
                    pk.y = pack_h16x2((v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w, bf);
                    reinterpret_cast<uint2*>(xn + row * E)[e4] = pk;
                }
            }
        }
    }
}

// h = 16-bit(gelu(sum_s partial[s] + bias))   (only when fc1 runs split-K); 4 elements per thread, all partial loads in flight
__global__ void __launch_bounds__(256)
act_reduce_kernel(const float* __restrict__ partial, int S, const float* __restrict__ bias, h16* __restrict__ h, int B, int N, int bf,
                  long long* tr) {
    tc::pdl_launch_dependents();
    TR_IN(tr);
    tc::pdl_wait();
    TR_DEP(tr);
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
        uint2 pk;
        pk.x = pack_h16x2(r[0], r[1], bf);
        pk.y = pack_h16x2(r[2], r[3], bf);
        *reinterpret_cast<uint2*>(h + i * 4) = pk;
    }
    TR_OUT(tr);
}

// one warp per (b, head): reduce the split-K qkv partials (+bias), append k,v at row t of the 16-bit cache, attend.
// lane <-> dims (2*lane, 2*lane+1) for q/k/v/out everywhere: every cache row is read as ONE coalesced 128 B line per warp
// instruction (a lane-per-key row read costs 8x the L1 wavefronts); the per-key dot products are finished with a 31-shuffle
// transpose-reduce per 32 keys, after which lane j holds the score of key j.  Up to 64 K rows / 64 V rows are in flight at once
// and the V rows are requested before the softmax arithmetic.  T <= 512.
// (Measured and dropped in round 2: pulling the cached rows into L2 ahead of griddepcontrol.wait -- from this kernel or from the
// previous layer's -- changes nothing (the two 64-row phases cost 4.4 us each at T = 64 either way); reading them into registers
// ahead of the wait is a race in the head graph, where the producer of row d-1 is a kernel of the SAME graph that a chain of
// small launches does not keep from still being in flight.)
constexpr int AF_MAXT = 512;

// pv[u] = this lane's partial dot product for key u (u < 32); returns the full dot product of key `lane`
__device__ __forceinline__ float af_transpose_reduce(float (&pv)[32], int lane) {
#pragma unroll
    for (int S = 16; S >= 1; S >>= 1) {
        const bool up = (lane & S) != 0;
#pragma unroll
        for (int i = 0; i < S; i++) {
            const float send = up ? pv[i] : pv[i + S];
            const float keep = up ? pv[i + S] : pv[i];
            pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
        }
    }
    return pv[0];
}

__global__ void __launch_bounds__(128)
attn_fast_kernel(const float* __restrict__ part, int S, const float* __restrict__ bqkv, h16* __restrict__ kc, h16* __restrict__ vc,
                 h16* __restrict__ att, int B, int E, int nh, int Tmax, const int* __restrict__ t_ptr, int t_host, int bf,
                 long long* tr) {
    extern __shared__ float af_smem[];              // ps[4][tp]
    const int tp = (Tmax + 31) & ~31;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    float* ps = af_smem + wq * tp;
    tc::pdl_launch_dependents();
    TR_IN(tr);
    tc::pdl_wait();
    TR_DEP(tr);
    const int bh = blockIdx.x * 4 + wq;
    if (bh < B * nh) {
    const int b = bh / nh, h = bh % nh;
    const int t = t_ptr ? *t_ptr : t_host;
    h16* kb = kc + ((int64_t)(b * nh + h) * Tmax) * 64;
    h16* vb = vc + ((int64_t)(b * nh + h) * Tmax) * 64;
    const int c = h * 64 + 2 * lane;
    float2 q = make_float2(bqkv[c], bqkv[c + 1]);
    float2 k = make_float2(bqkv[E + c], bqkv[E + c + 1]);
    float2 v = make_float2(bqkv[2 * E + c], bqkv[2 * E + c + 1]);
#pragma unroll 4
    for (int s = 0; s < S; s++) {
        const float* p = part + ((int64_t)s * B + b) * 3 * E;
        float2 a = *reinterpret_cast<const float2*>(p + c);
        float2 bb = *reinterpret_cast<const float2*>(p + E + c);
        float2 cc = *reinterpret_cast<const float2*>(p + 2 * E + c);
        q.x += a.x; q.y += a.y; k.x += bb.x; k.y += bb.y; v.x += cc.x; v.y += cc.y;
    }
    const uint32_t k2 = pack_h16x2(k.x, k.y, bf), v2 = pack_h16x2(v.x, v.y, bf);
    *reinterpret_cast<uint32_t*>(kb + (int64_t)t * 64 + 2 * lane) = k2;
    *reinterpret_cast<uint32_t*>(vb + (int64_t)t * 64 + 2 * lane) = v2;
    // use the 16-bit-rounded q/k/v everywhere (what a later step reads back from the cache)
    const float2 qf = unpack_h16x2(pack_h16x2(q.x, q.y, bf), bf), kf = unpack_h16x2(k2, bf), vf = unpack_h16x2(v2, bf);
    const float s_new = warp_sum(qf.x * kf.x + qf.y * kf.y) * 0.125f;
    float m = s_new;
    for (int j0 = 0; j0 < t; j0 += 64) {          // scores of the cached rows: 64 coalesced row reads in flight
        uint32_t kr[64];
#pragma unroll
        for (int u = 0; u < 64; u++)
            kr[u] = (j0 + u < t) ? *reinterpret_cast<const uint32_t*>(kb + (int64_t)(j0 + u) * 64 + 2 * lane) : 0u;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            if (j0 + half * 32 < t) {                                  // (warp-uniform)
                float pv[32];
#pragma unroll
                for (int u = 0; u < 32; u++) {
                    const float2 kk = unpack_h16x2(kr[half * 32 + u], bf);
                    pv[u] = fmaf(qf.y, kk.y, qf.x * kk.x);
                }
                const float sc = af_transpose_reduce(pv, lane) * 0.125f;
                const int j = j0 + half * 32 + lane;
                if (j < t) {
                    ps[j] = sc;
                    m = fmaxf(m, sc);
                }
            }
        }
    }
    // the V rows do not depend on the scores: the first 64 are requested before the softmax arithmetic
    uint32_t raw[64];
#pragma unroll
    for (int u = 0; u < 64; u++)
        raw[u] = (u < t) ? *reinterpret_cast<const uint32_t*>(vb + (int64_t)u * 64 + 2 * lane) : 0u;
    m = warp_max(m);
    if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[2] = tc::gtimer();      // scores done
    float sum = 0.f;
    for (int j = lane; j < t; j += 32) {
        float e = __expf(ps[j] - m);
        ps[j] = e;
        sum += e;
    }
    const float e_new = __expf(s_new - m);
    sum = warp_sum(sum) + e_new;
    __syncwarp();
    const float inv = 1.0f / sum;
    float2 o = make_float2(e_new * vf.x, e_new * vf.y);
    for (int j0 = 0; j0 < t; j0 += 64) {                  // (rows added in cache order)
        if (j0 > 0) {
#pragma unroll
            for (int u = 0; u < 64; u++)
                raw[u] = (j0 + u < t) ? *reinterpret_cast<const uint32_t*>(vb + (int64_t)(j0 + u) * 64 + 2 * lane) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 64; u++)
            if (j0 + u < t) {
                const float2 vv = unpack_h16x2(raw[u], bf);
                o.x = fmaf(ps[j0 + u], vv.x, o.x);
                o.y = fmaf(ps[j0 + u], vv.y, o.y);
            }
    }
    *reinterpret_cast<uint32_t*>(att + (int64_t)b * E + c) = pack_h16x2(o.x * inv, o.y * inv, bf);
    }
    if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[3] = tc::gtimer();
}

// Four warps per (b, head) -- the body stack's form.  The cached K / V rows [0, t) are staged in shared memory by cp.async (16 B
// per request, every request of the CTA in flight at once, issued right after the dependency resolves and overlapped with the
// q/k/v split-K reduction): the whole KV read of a step is ONE memory round trip instead of two (K, then V) serialised per 64 rows
// in one warp's registers.  Rows are 128 B with the 16 B chunks XOR-swizzled by (row & 7): conflict-free for the score pass
// (a lane pair per row, 4 chunks each) and for the output pass (a warp per row, lane <-> dims (2*lane, 2*lane+1)).
// Softmax statistics go through shared memory; warp w adds the rows j = w (mod 4) in cache order and the four partial outputs
// are summed in warp order -> run-to-run deterministic.  q/k/v bits as in attn_fast_kernel (same reduction order).
// 11 CTAs per SM (40 registers, 19.6 KB at 64 rows): the 1536 (b, head) pairs of the 1.4B model at B = 64 are one wave.
constexpr int AF2_MAXROWS = 320;
static size_t attn2_smem(int rows) { return (size_t)rows * 256 + ((size_t)rows + 64 * 3 + 4 * 64 + 8) * sizeof(float); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__global__ void __launch_bounds__(128, 11)
attn_fast2_kernel(const float* __restrict__ part, int S, const float* __restrict__ bqkv, h16* __restrict__ kc, h16* __restrict__ vc,
                  h16* __restrict__ att, int B, int E, int nh, int Tmax, int rows, const int* __restrict__ t_ptr, int t_host, int bf,
                  long long* tr) {
    extern __shared__ __align__(128) uint8_t af2_smem[];
    uint8_t* Ks = af2_smem;                                   // [rows][128 B], chunk c of row j at ((c ^ (j & 7)) << 4)
    uint8_t* Vs = Ks + (size_t)rows * 128;
    float* ps = reinterpret_cast<float*>(Vs + (size_t)rows * 128);   // scores, then exp(score - max)
    float* qs = ps + rows;                // q, k_new, v_new (16-bit-rounded), 64 floats each
    float* kn = qs + 64;
    float* vn = kn + 64;
    float* ov = vn + 64;                  // [4][64] partial outputs
    float* red = ov + 256;                // [0..3] warp maxima, [4..7] warp sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    tc::pdl_launch_dependents();
    TR_IN(tr);
    tc::pdl_wait();
    TR_DEP(tr);
    const int b = blockIdx.x / nh, h = blockIdx.x % nh;
    const int t = t_ptr ? *t_ptr : t_host;
    h16* kb = kc + ((int64_t)(b * nh + h) * Tmax) * 64;
    h16* vb = vc + ((int64_t)(b * nh + h) * Tmax) * 64;
    {
        const uint32_t ks = tc::smem_u32(Ks), vs = tc::smem_u32(Vs);
        for (int i = threadIdx.x; i < t * 8; i += 128) {
            const int j = i >> 3, c = i & 7;
            const uint32_t off = (uint32_t)j * 128u + (uint32_t)((c ^ (j & 7)) << 4);
            cp_async16(ks + off, kb + (int64_t)j * 64 + c * 8);
            cp_async16(vs + off, vb + (int64_t)j * 64 + c * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- q / k / v of the new token: warp 0 -> q, warp 1 -> k, warp 2 -> v (value = bias + p0 + p1 + ..., split order)
    if (w < 3) {
        const int c = h * 64 + 2 * lane;
        float2 a = make_float2(bqkv[w * E + c], bqkv[w * E + c + 1]);
#pragma unroll 4
        for (int s = 0; s < S; s++) {
            const float2 pp = *reinterpret_cast<const float2*>(part + ((int64_t)s * B + b) * 3 * E + w * E + c);
            a.x += pp.x; a.y += pp.y;
        }
        const uint32_t a2 = pack_h16x2(a.x, a.y, bf);
        const float2 af = unpack_h16x2(a2, bf);
        float* dst = w == 0 ? qs : (w == 1 ? kn : vn);
        dst[2 * lane] = af.x;
        dst[2 * lane + 1] = af.y;
        if (w == 1) *reinterpret_cast<uint32_t*>(kb + (int64_t)t * 64 + 2 * lane) = a2;
        if (w == 2) *reinterpret_cast<uint32_t*>(vb + (int64_t)t * 64 + 2 * lane) = a2;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float s_new = warp_sum(qs[2 * lane] * kn[2 * lane] + qs[2 * lane + 1] * kn[2 * lane + 1]) * 0.125f;
    float m = -INFINITY;
    {
        const int half = threadIdx.x & 1;
        for (int j0 = 0; j0 < t; j0 += 64) {
            const int j = j0 + (threadIdx.x >> 1);
            float acc = 0.f;
            if (j < t) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int c = half * 4 + i;
                    const uint4 k4 = *reinterpret_cast<const uint4*>(Ks + (size_t)j * 128 + ((c ^ (j & 7)) << 4));
                    const float4 qa = *reinterpret_cast<const float4*>(qs + c * 8), qb = *reinterpret_cast<const float4*>(qs + c * 8 + 4);
                    float2 kk = unpack_h16x2(k4.x, bf);
                    acc = fmaf(qa.x, kk.x, acc); acc = fmaf(qa.y, kk.y, acc);
                    kk = unpack_h16x2(k4.y, bf);
                    acc = fmaf(qa.z, kk.x, acc); acc = fmaf(qa.w, kk.y, acc);
                    kk = unpack_h16x2(k4.z, bf);
                    acc = fmaf(qb.x, kk.x, acc); acc = fmaf(qb.y, kk.y, acc);
                    kk = unpack_h16x2(k4.w, bf);
                    acc = fmaf(qb.z, kk.x, acc); acc = fmaf(qb.w, kk.y, acc);
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc *= 0.125f;
            if (j < t) {
                if (half == 0) ps[j] = acc;
                m = fmaxf(m, acc);
            }
        }
    }
    m = warp_max(m);
    if (lane == 0) red[w] = m;
    __syncthreads();
    m = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), s_new);
    if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[2] = tc::gtimer();      // scores done
    float sum = 0.f;
    for (int j = threadIdx.x; j < t; j += 128) {
        const float e = __expf(ps[j] - m);
        ps[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + w] = sum;
    __syncthreads();
    const float e_new = __expf(s_new - m);
    const float inv = 1.0f / ((((red[4] + red[5]) + red[6]) + red[7]) + e_new);
    float2 o = w == 0 ? make_float2(e_new * vn[2 * lane], e_new * vn[2 * lane + 1]) : make_float2(0.f, 0.f);
    for (int j = w; j < t; j += 4) {
        const uint32_t v2 = *reinterpret_cast<const uint32_t*>(Vs + (size_t)j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + ((lane & 3) << 2));
        const float2 vv = unpack_h16x2(v2, bf);
        const float pj = ps[j];
        o.x = fmaf(pj, vv.x, o.x);
        o.y = fmaf(pj, vv.y, o.y);
    }
    ov[w * 64 + 2 * lane] = o.x;
    ov[w * 64 + 2 * lane + 1] = o.y;
    __syncthreads();
    if (w == 0) {
        const float ox = ((ov[2 * lane] + ov[64 + 2 * lane]) + ov[128 + 2 * lane]) + ov[192 + 2 * lane];
        const float oy = ((ov[2 * lane + 1] + ov[64 + 2 * lane + 1]) + ov[128 + 2 * lane + 1]) + ov[192 + 2 * lane + 1];
        *reinterpret_cast<uint32_t*>(att + (int64_t)b * E + h * 64 + 2 * lane) = pack_h16x2(ox * inv, oy * inv, bf);
    }
    if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[3] = tc::gtimer();
}

// Causal attention over a whole prefix in one launch (batched prefill / teacher-forced forward).  qkv [M, 3E] 16-bit (bias already
// added by the GEMM epilogue), row of (group g, token t) = t * G + g  (token-major: the rows of one token are contiguous, like the
// single-step buffers).  One CTA per (group, head); the group's K and V rows are staged in shared memory (row stride 66 elements:
// conflict-free for lane <-> key), warp <-> query, lane <-> key for the scores and lane <-> 2 dims for the output -- the same
// arithmetic order as attn_fast_kernel's.  When kc != NULL the K / V rows are also written to the cache [g][head][t][64].
constexpr int PA_MAXT = 512;
static size_t prefill_attn_smem(int T) { return ((size_t)2 * T * 33 + 4 * 64 + (size_t)4 * T) * 4; }
__global__ void __launch_bounds__(128)
prefill_attn_kernel(const h16* __restrict__ qkv, h16* __restrict__ kc, h16* __restrict__ vc, h16* __restrict__ att, int G, int T, int E,
                    int nh, int Tmax, int bf) {
    extern __shared__ uint32_t pa_smem[];
    uint32_t (*ks)[33] = reinterpret_cast<uint32_t (*)[33]>(pa_smem);
    uint32_t (*vs)[33] = reinterpret_cast<uint32_t (*)[33]>(pa_smem + (size_t)T * 33);
    float (*qs)[64] = reinterpret_cast<float (*)[64]>(pa_smem + (size_t)2 * T * 33);
    float* ps_all = reinterpret_cast<float*>(pa_smem + (size_t)2 * T * 33 + 4 * 64);
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int g = blockIdx.x / nh, h = blockIdx.x % nh;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    float* ps = ps_all + (size_t)wq * T;
    for (int i = threadIdx.x; i < T * 32; i += 128) {
        const int t = i >> 5, c2 = i & 31;
        const h16* row = qkv + ((int64_t)t * G + g) * 3 * E + h * 64 + 2 * c2;
        const uint32_t kk = *reinterpret_cast<const uint32_t*>(row + E), vv = *reinterpret_cast<const uint32_t*>(row + 2 * E);
        ks[t][c2] = kk;
        vs[t][c2] = vv;
        if (kc) {
            *reinterpret_cast<uint32_t*>(kc + (((int64_t)g * nh + h) * Tmax + t) * 64 + 2 * c2) = kk;
            *reinterpret_cast<uint32_t*>(vc + (((int64_t)g * nh + h) * Tmax + t) * 64 + 2 * c2) = vv;
        }
    }
    __syncthreads();
    for (int t = wq; t < T; t += 4) {
        const float2 qf = unpack_h16x2(*reinterpret_cast<const uint32_t*>(qkv + ((int64_t)t * G + g) * 3 * E + h * 64 + 2 * lane), bf);
        __syncwarp();
        qs[wq][2 * lane] = qf.x;
        qs[wq][2 * lane + 1] = qf.y;
        __syncwarp();
        float m = -INFINITY;
        for (int j = lane; j <= t; j += 32) {
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < 32; u++) {
                const float2 kk = unpack_h16x2(ks[j][u], bf);
                acc = fmaf(qs[wq][2 * u], kk.x, acc);
                acc = fmaf(qs[wq][2 * u + 1], kk.y, acc);
            }
            acc *= 0.125f;
            ps[j] = acc;
            m = fmaxf(m, acc);
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j <= t; j += 32) {
            const float e = __expf(ps[j] - m);
            ps[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float2 o = make_float2(0.f, 0.f);
        for (int j = 0; j <= t; j++) {
            const float2 vv = unpack_h16x2(vs[j][lane], bf);
            o.x = fmaf(ps[j], vv.x, o.x);
            o.y = fmaf(ps[j], vv.y, o.y);
        }
        const float inv = 1.0f / sum;
        *reinterpret_cast<uint32_t*>(att + ((int64_t)t * G + g) * E + h * 64 + 2 * lane) = pack_h16x2(o.x * inv, o.y * inv, bf);
    }
}

// The same causal attention for prefixes of at most 64 tokens (the 8x8 grids' body pass: T = cond_len + 63) on the warp-level tensor
// cores: Q, K, V of one (group, head) are staged in shared memory (144 B rows: conflict-free ldmatrix), each of the four warps owns
// 16 query rows -- S = Q K^T as 8 n-tiles x 4 k-steps of mma.sync.m16n8k16 (fp32 accumulate), scale, causal mask, row softmax in
// registers (quad shuffles), the probabilities repacked as 16-bit A fragments (the m16n8 accumulator pair of two adjacent key tiles IS
// the m16k16 A fragment), O = P V with V through ldmatrix.trans.  ~100 tensor instructions per warp instead of ~270 k scalar FMAs per
// CTA.  (Not the tcgen05 path: 64 x 64 x 64 per head is two orders of magnitude below a UMMA tile's worth of work.)
template <bool BF>
__device__ __forceinline__ void pa_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (BF)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pa_ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pa_ldsm4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
constexpr int PM_ROW = 144;                           // bytes per staged row (64 x 16-bit + 16 B pad)
template <bool BF>
__global__ void __launch_bounds__(128)
prefill_attn_mma_kernel(const h16* __restrict__ qkv, h16* __restrict__ kc, h16* __restrict__ vc, h16* __restrict__ att, int G, int T, int E,
                        int nh, int Tmax) {
    __shared__ __align__(16) uint8_t sm[3 * 64 * PM_ROW];
    uint8_t* Qs = sm;
    uint8_t* Ks = sm + 64 * PM_ROW;
    uint8_t* Vs = sm + 2 * 64 * PM_ROW;
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int g = blockIdx.x / nh, h = blockIdx.x % nh;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // ---- stage Q, K, V rows [0, T) (rows beyond T: zeros); K / V also go to the cache
    for (int i = threadIdx.x; i < 3 * 64 * 8; i += 128) {
        const int mat = i / 512, t = (i >> 3) & 63, c = i & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t < T) v = *reinterpret_cast<const uint4*>(qkv + ((int64_t)t * G + g) * 3 * E + mat * E + h * 64 + c * 8);
        *reinterpret_cast<uint4*>(sm + mat * 64 * PM_ROW + t * PM_ROW + c * 16) = v;
        if (kc != nullptr && mat > 0 && t < T)
            *reinterpret_cast<uint4*>((mat == 1 ? kc : vc) + (((int64_t)g * nh + h) * Tmax + t) * 64 + c * 8) = v;
    }
    __syncthreads();
    if (16 * w >= T) return;                              // (no query rows for this warp)
    const uint32_t qs = tc::smem_u32(Qs), ks = tc::smem_u32(Ks), vs = tc::smem_u32(Vs);
    // ---- S = Q K^T for this warp's 16 query rows
    float sacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) { sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        uint32_t a[4];
        pa_ldsm4(a, qs + (uint32_t)((16 * w + (lane & 15)) * PM_ROW + kk * 32 + (lane >> 4) * 16));
#pragma unroll
        for (int jp = 0; jp < 4; jp++) {                 // two key tiles per ldmatrix.x4
            uint32_t b[4];
            pa_ldsm4(b, ks + (uint32_t)((16 * jp + (lane & 7) + ((lane >> 4) << 3)) * PM_ROW + kk * 32 + ((lane >> 3) & 1) * 16));
            pa_mma<BF>(sacc[2 * jp], a, b[0], b[1]);
            pa_mma<BF>(sacc[2 * jp + 1], a, b[2], b[3]);
        }
    }
    // ---- scale, causal mask, softmax over the row (a row's 64 scores live in the 4 lanes of a quad: 16 each)
    const int r0 = 16 * w + (lane >> 2), r1 = r0 + 8;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c0 = 8 * j + (lane & 3) * 2;
        sacc[j][0] = (c0 <= r0) ? sacc[j][0] * 0.125f : -INFINITY;
        sacc[j][1] = (c0 + 1 <= r0) ? sacc[j][1] * 0.125f : -INFINITY;
        sacc[j][2] = (c0 <= r1) ? sacc[j][2] * 0.125f : -INFINITY;
        sacc[j][3] = (c0 + 1 <= r1) ? sacc[j][3] * 0.125f : -INFINITY;
        m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
        m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
    uint32_t pa[4][4];                                    // probabilities as A fragments, one per 16-key step
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float e0 = __expf(sacc[j][0] - m0), e1 = __expf(sacc[j][1] - m0), e2 = __expf(sacc[j][2] - m1), e3 = __expf(sacc[j][3] - m1);
        s0 += e0 + e1;
        s1 += e2 + e3;
        pa[j >> 1][(j & 1) * 2] = pack_h16x2(e0, e1, BF ? 1 : 0);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_h16x2(e2, e3, BF ? 1 : 0);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    // ---- O = P V
    float oacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) { oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {                      // 16 keys per step
#pragma unroll
        for (int jp = 0; jp < 4; jp++) {                 // two 8-dim output tiles per ldmatrix.x4.trans
            uint32_t b[4];
            pa_ldsm4_t(b, vs + (uint32_t)((16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * PM_ROW + (2 * jp + (lane >> 4)) * 16));
            pa_mma<BF>(oacc[2 * jp], pa[kk], b[0], b[1]);
            pa_mma<BF>(oacc[2 * jp + 1], pa[kk], b[2], b[3]);
        }
    }
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int d = 8 * j + (lane & 3) * 2;
        if (r0 < T) *reinterpret_cast<uint32_t*>(att + ((int64_t)r0 * G + g) * E + h * 64 + d) = pack_h16x2(oacc[j][0] * i0, oacc[j][1] * i0, BF ? 1 : 0);
        if (r1 < T) *reinterpret_cast<uint32_t*>(att + ((int64_t)r1 * G + g) * E + h * 64 + d) = pack_h16x2(oacc[j][2] * i1, oacc[j][3] * i1, BF ? 1 : 0);
    }
}

// ... and for groups of at most 8 tokens (the head stack of the teacher-forced forward: D tokens per (position, batch row), ~10^5
// (group, head) pairs): one WARP per pair, everything in registers, lane <-> dims (2*lane, 2*lane+1), scores by warp reductions.
template <int TMAXS>
__global__ void __launch_bounds__(128)
prefill_attn_small_kernel(const h16* __restrict__ qkv, h16* __restrict__ kc, h16* __restrict__ vc, h16* __restrict__ att, int G, int T, int E,
                          int nh, int Tmax, int bf) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t pair = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (pair >= (int64_t)G * nh) return;
    const int g = (int)(pair / nh), h = (int)(pair % nh);
    float2 q[TMAXS], k[TMAXS], v[TMAXS];
#pragma unroll
    for (int t = 0; t < TMAXS; t++)
        if (t < T) {
            const h16* row = qkv + ((int64_t)t * G + g) * 3 * E + h * 64 + 2 * lane;
            const uint32_t q2 = *reinterpret_cast<const uint32_t*>(row), k2 = *reinterpret_cast<const uint32_t*>(row + E),
                           v2 = *reinterpret_cast<const uint32_t*>(row + 2 * E);
            q[t] = unpack_h16x2(q2, bf); k[t] = unpack_h16x2(k2, bf); v[t] = unpack_h16x2(v2, bf);
            if (kc) {
                *reinterpret_cast<uint32_t*>(kc + (((int64_t)g * nh + h) * Tmax + t) * 64 + 2 * lane) = k2;
                *reinterpret_cast<uint32_t*>(vc + (((int64_t)g * nh + h) * Tmax + t) * 64 + 2 * lane) = v2;
            }
        }
#pragma unroll
    for (int t = 0; t < TMAXS; t++)
        if (t < T) {
            float sc[TMAXS];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < TMAXS; j++)
                if (j <= t) {
                    sc[j] = warp_sum(fmaf(q[t].y, k[j].y, q[t].x * k[j].x)) * 0.125f;
                    m = fmaxf(m, sc[j]);
                }
            float sum = 0.f;
            float2 o = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < TMAXS; j++)
                if (j <= t) {
                    const float e = __expf(sc[j] - m);
                    sum += e;
                    o.x = fmaf(e, v[j].x, o.x);
                    o.y = fmaf(e, v[j].y, o.y);
                }
            const float inv = 1.0f / sum;
            *reinterpret_cast<uint32_t*>(att + ((int64_t)t * G + g) * E + h * 64 + 2 * lane) = pack_h16x2(o.x * inv, o.y * inv, bf);
        }
}

// token sources --------------------------------------------------------------------------------------------------
// cond token s: x[b,:] = cond_emb[cond[b,s]] + pos_emb_cond[s]                      (transformers.py:224)
// grid (B, n_tokens): token s = stt->s + blockIdx.y, written to row blockIdx.y * B + b
__global__ void __launch_bounds__(256)
cond_tok_kernel(const StepState* __restrict__ stt, const float* __restrict__ cond_emb, const float* __restrict__ pos_cond,
                int cond_len, int vocab_cond, int E, float* __restrict__ x) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x, s = stt->s + blockIdx.y, B = gridDim.x;
    int64_t c = stt->cond ? stt->cond[(int64_t)b * cond_len + s] : 0;
    c = c < 0 ? 0 : (c >= vocab_cond ? vocab_cond - 1 : c);
    float* xr = x + ((int64_t)blockIdx.y * B + b) * E;
    for (int e = threadIdx.x; e < E; e += 256) xr[e] = cond_emb[c * E + e] + pos_cond[(int64_t)s * E + e];
}
// summed code embeddings in 16-bit: mode 0 -> all D codes of position idx-1 (body input), mode d>=1 -> codes 0..d-1 of
// position idx (head input, cumsum)                                              (transformers.py:219-225, 250-255)
// mode < 0 (prefill / forward): grid (B, n_pos); the first -mode codes of position blockIdx.y + pos0, written to row blockIdx.y * B + b
__global__ void __launch_bounds__(64)
code_sum_kernel(const StepState* __restrict__ stt, const float* __restrict__ cb, int HW, int D, int K, int C, int mode, int pos0,
                h16* __restrict__ out, int bf) {
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int b = blockIdx.x, B = gridDim.x;
    const int pos = mode < 0 ? pos0 + blockIdx.y : (mode == 0 ? stt->idx - 1 : stt->idx);
    const int nd = mode == 0 ? D : (mode < 0 ? -mode : mode);
    h16* o = out + ((int64_t)blockIdx.y * B + b) * C;
    for (int c = threadIdx.x; c < C; c += 64) {
        float a = 0.f;
        for (int i = 0; i < nd; i++) {
            int64_t k = stt->codes[((int64_t)b * HW + pos) * D + i];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            a += cb[k * C + c];
        }
        o[c] = pack_h16(a, bf);
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

// keep_pos != 0: a resumed span keeps the position counters the previous span left behind
__global__ void init_state_kernel(StepState* dst, StepState v, int keep_pos) {
    if (threadIdx.x == 0) {
        if (keep_pos) { v.s = dst->s; v.idx = dst->idx; }
        *dst = v;
    }
}

// ------------------------------------------------------------------------------------------------ engine
struct FastLayer {
    CUtensorMap qkv, proj, fc1, fc2;
};

enum { G_COND = 0, G_CODE = 1, G_HEAD = 2, G_HEAD_LOGITS = 3, G_COUNT = 4 };

struct ArFast {
    rqb200_ar_config cfg;
    rqb200_ar_weights w;
    std::vector<rqb200_block_weights> body, head;
    std::vector<FastLayer> lbody, lhead;
    CUtensorMap tm_win, tm_whead, tm_cls, tm_ccls;
    int bf = 0;                          // 16-bit format: 0 fp16, 1 bf16
    // per (workspace, B) state
    void* ws_base = nullptr;
    int B = 0;
    CUtensorMap tx_xn, tx_att, tx_h, tx_s;
    cudaGraphExec_t graphs[G_COUNT] = {nullptr, nullptr, nullptr, nullptr};
    int64_t n_nodes[G_COUNT] = {0, 0, 0, 0};   // kernels recorded in each graph (for the launch counter)
    cudaStream_t cap_stream = nullptr;   // capture never happens on the caller's stream (it may be the legacy default stream)
    bool use_graph = true, use_pdl = true, attn4 = true, param_prefetch = true, deep = true, l2pf = false, batched_prefill = true, batched_deep = false, batched_streamer = false;
    int split_qkv = 4, split_proj = 12, split_fc1 = 1, split_fc2 = 12;
    int n_sm = 148;
    // diagnostic stage trace (cfg.flags & RQB200_AR_TRACE)
    bool trace = false, trace_w = false;
    mutable long long* tr_base = nullptr;
    mutable int tr_next = 0;
    mutable std::vector<std::string> tr_names;
    int tr_graph_base[G_COUNT + 1] = {0, 0, 0, 0, 0};
};

constexpr int TR_CAP = 4096;             // launches per trace buffer

struct FastWs {
    StepState* state;
    long long* trace;
    float *XB, *XH, *P, *LOGITS;
    h16 *XN, *ATT, *Hh, *S;
    h16 *kc_body, *vc_body, *kc_head, *vc_head;
    // batched prefill (M = B * T rows, token-major)
    int64_t Mmax;
    float* PX;                   // [Mmax, E] residual stream
    h16 *PXN, *PQKV, *PATT, *PH, *PS;
};

static int pick_split(int n_tiles, int nkb, int want, int n_sm) {
    int s = want > 0 ? want : n_sm / n_tiles;
    if (s < 1) s = 1;
    if (s > nkb) s = nkb;
    return s;
}

static long long* tr_slot(const ArFast& f, const char* name) {
    if (!f.trace || !f.tr_base || f.tr_next >= TR_CAP) return nullptr;
    if ((int)f.tr_names.size() <= f.tr_next) f.tr_names.resize(f.tr_next + 1);
    f.tr_names[f.tr_next] = name;
    return f.tr_base + 4 * (int64_t)(f.tr_next++);
}

static int prefill_tmax(const rqb200_ar_config& c) { return std::min(PA_MAXT, c.cond_len + c.H * c.W - 1); }

static size_t fast_layout(const ArFast& f, int B, void* base, size_t cap, FastWs* ws) {
    const rqb200_ar_config& c = f.cfg;
    Arena a(base, cap);
    const int64_t E = c.embed_dim, HW = (int64_t)c.H * c.W, Tb = c.cond_len + HW;
    FastWs w;
    w.state = a.take<StepState>(1);
    w.trace = a.take<long long>(4 * TR_CAP);
    w.XB = a.take<float>(B * E);
    w.XH = a.take<float>(B * E);
    int maxs = std::max(std::max(f.split_qkv * 3, f.split_proj), std::max(f.split_fc2, f.split_fc1 * 4));
    w.P = a.take<float>((int64_t)maxs * B * E);
    w.LOGITS = a.take<float>((int64_t)B * c.vocab);
    w.XN = a.take<h16>(B * E);
    w.ATT = a.take<h16>(B * E);
    w.Hh = a.take<h16>(B * 4 * E);
    w.S = a.take<h16>((int64_t)B * c.code_dim);
    const int64_t per_body = (int64_t)B * c.n_head * Tb * 64, per_head = (int64_t)B * c.n_head * c.D * 64;
    w.kc_body = a.take<h16>(per_body * c.n_body);
    w.vc_body = a.take<h16>(per_body * c.n_body);
    w.kc_head = a.take<h16>(per_head * c.n_head_layers);
    w.vc_head = a.take<h16>(per_head * c.n_head_layers);
    w.Mmax = f.batched_prefill ? (int64_t)B * prefill_tmax(c) : 0;
    const int64_t Mp = w.Mmax ? w.Mmax + 128 : 0;          // (the rows GEMM reads whole 128-row tiles)
    w.PX = a.take<float>(Mp * E);
    w.PXN = a.take<h16>(Mp * E);
    w.PQKV = a.take<h16>(Mp * 3 * E);
    w.PATT = a.take<h16>(Mp * E);
    w.PH = a.take<h16>(Mp * 4 * E);
    w.PS = a.take<h16>(w.Mmax * c.code_dim);
    if (ws) *ws = w;
    return a.off + 256;
}

static GemmTcParams gemm_base(const ArFast& f, int N_out, int K, int rows, int splits, int mode) {
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = rows; p.splits = splits; p.mode = mode;
    p.fmt = f.bf; p.deep = f.deep ? 1 : 0; p.l2pf = f.l2pf ? 1 : 0; p.bias_scale = 1.f; p.ld_out = N_out;
    return p;
}

static int gemm(const ArFast& f, const char* name, const CUtensorMap& tw, const CUtensorMap& tx, int N_out, int K, int B, int splits,
                int mode, const float* bias, float bias_scale, void* out, float* partial, const float* residual, int64_t ld_res,
                const int* res_row_ptr, int64_t res_row_stride, cudaStream_t st) {
    GemmTcParams p = gemm_base(f, N_out, K, B, splits, mode);
    p.bias = bias; p.bias_scale = bias_scale; p.out = out; p.partial = partial;
    p.residual = residual; p.ld_res = ld_res; p.res_row_ptr = res_row_ptr; p.res_row_stride = res_row_stride;
    p.trace = tr_slot(f, name);
    p.trace_w = f.trace_w ? 1 : 0;
    return launch_gemm_tc(tw, tx, p, f.use_pdl, st);
}

static int ln(const ArFast& f, const char* name, int rows, const float* x_in, const float* partial, int S, const float* bias,
              const float* extra, float* x_out, const float* g, const float* be, h16* xn, cudaStream_t st,
              const PrefetchList* pf = nullptr) {
    const int E = f.cfg.embed_dim;
    if (rows >= 512 && S == 0 && bias == nullptr && x_in != nullptr) {        // batched passes: warp per row
        const dim3 grid((unsigned)std::min<int64_t>(ceil_div(rows, 8), (int64_t)f.n_sm * 8));
        const int nv = ceil_div(E, 128);
#define RQB_LN_ROWS(NV) launch_pdl(ln_rows_kernel<NV>, grid, dim3(256), (size_t)0, st, true, x_in, extra, x_out, g, be, xn, (int64_t)rows, E, f.bf)
        if (nv <= 8) return RQB_LN_ROWS(8);
        if (nv <= 12) return RQB_LN_ROWS(12);
        if (nv <= 20) return RQB_LN_ROWS(20);
        return RQB_LN_ROWS(36);
#undef RQB_LN_ROWS
    }
    PrefetchList none = {};
    return launch_pdl(ln_reduce_kernel<384, 3>, dim3((unsigned)rows), dim3(384), (size_t)0, st, f.use_pdl, x_in, partial, S, bias, extra, x_out,
                      g, be, xn, rows, E, f.bf, tr_slot(f, name), (pf && f.param_prefetch) ? *pf : none);
}

static int attn(const ArFast& f, FastWs& ws, const float* bqkv, h16* kc, h16* vc, int Tmax, const int* t_ptr, int t_host,
                cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    if (f.attn4 && Tmax >= 16 && Tmax - 1 <= AF2_MAXROWS) {          // the body stack: four warps per (b, head)
        const int rows = (Tmax - 1 + 7) & ~7;                        // cached rows a step can read (row t is the new token); 32 B-aligned float arrays behind them
        RQB_ENSURE_SMEM(attn2_smem(AF2_MAXROWS), attn_fast2_kernel);
        {   // 11 CTAs x 19.6 KB need the largest shared-memory carve-out (L1 is not used by this kernel)
            static std::atomic<uint64_t> carve{0};
            int dev = 0;
            RQB_CUDA(cudaGetDevice(&dev));
            if (!(carve.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
                RQB_CUDA(cudaFuncSetAttribute(attn_fast2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                carve.fetch_or(1ull << (dev & 63), std::memory_order_release);
            }
        }
        return launch_pdl(attn_fast2_kernel, dim3((unsigned)(f.B * c.n_head)), dim3(128), attn2_smem(rows), st, f.use_pdl,
                          (const float*)ws.P, f.split_qkv, bqkv, kc, vc, ws.ATT, f.B, c.embed_dim, c.n_head, Tmax, rows, t_ptr, t_host,
                          f.bf, tr_slot(f, "attn"));
    }
    const size_t smem = (size_t)(4 * ((Tmax + 31) & ~31)) * sizeof(float);
    return launch_pdl(attn_fast_kernel, dim3((unsigned)ceil_div(f.B * c.n_head, 4)), dim3(128), smem, st, f.use_pdl,
                      (const float*)ws.P, f.split_qkv, bqkv, kc, vc, ws.ATT, f.B, c.embed_dim, c.n_head, Tmax, t_ptr, t_host, f.bf,
                      tr_slot(f, "attn"));
}

// one transformer stack on the single new token of every batch row; x lives in `x` (fp32); residual additions are deferred
// into the next ln_reduce.
// fin_g / fin_b (nullable): a LayerNorm applied to the stack's output rows -> ws.XN (the classifier's), fused with whatever
// launch finishes x.
static int fast_stack(const ArFast& f, const std::vector<rqb200_block_weights>& blocks, const std::vector<FastLayer>& maps,
                      FastWs& ws, float* x, const float* pending_extra, const float* x_src, h16* kc, h16* vc, int Tmax,
                      const int* t_ptr, int t_host, const float* fin_g, const float* fin_b, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const int E = c.embed_dim, B = f.B;
    const int64_t per = (int64_t)B * c.n_head * Tmax * 64;
    const float* nof = nullptr;
    for (size_t l = 0; l < blocks.size(); l++) {
        const rqb200_block_weights& bw = blocks[l];
        const bool first = l == 0;
        // LN1 (+ pending fc2 reduction of the previous block)
        const bool pend = l > 0;
        PrefetchList pf = {};
        if (l + 1 < blocks.size()) {
            const rqb200_block_weights& nx = blocks[l + 1];
            const void* ptrs[8] = {nx.ln1_w, nx.ln1_b, nx.bqkv, nx.bproj, nx.ln2_w, nx.ln2_b, nx.b1, bw.b2};
            const uint32_t words[8] = {(uint32_t)E, (uint32_t)E, (uint32_t)(3 * E), (uint32_t)E, (uint32_t)E, (uint32_t)E, (uint32_t)(4 * E),
                                       (uint32_t)E};
            for (int i = 0; i < 8; i++) { pf.p[i] = ptrs[i]; pf.bytes[i] = words[i] * 4u; }
            pf.n = 8;
        }
        RQB_TRY(ln(f, "ln1", B, first ? x_src : x, pend ? ws.P : nof, pend ? f.split_fc2 : 0, pend ? blocks[l - 1].b2 : nof,
                   first ? pending_extra : nof, x, bw.ln1_w, bw.ln1_b, ws.XN, st, &pf));
        RQB_TRY(gemm(f, "qkv", maps[l].qkv, f.tx_xn, 3 * E, E, B, f.split_qkv, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0, nullptr, 0,
                     st));
        RQB_TRY(attn(f, ws, bw.bqkv, kc + per * l, vc + per * l, Tmax, t_ptr, t_host, st));
        RQB_TRY(gemm(f, "proj", maps[l].proj, f.tx_att, E, E, B, f.split_proj, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0, nullptr,
                     0, st));
        RQB_TRY(ln(f, "ln2", B, x, ws.P, f.split_proj, bw.bproj, nof, x, bw.ln2_w, bw.ln2_b, ws.XN, st));
        if (f.split_fc1 == 1) {
            RQB_TRY(gemm(f, "fc1", maps[l].fc1, f.tx_xn, 4 * E, E, B, 1, GT_H16_GELU, bw.b1, 1.f, ws.Hh, nullptr, nullptr, 0, nullptr, 0, st));
        } else {
            RQB_TRY(gemm(f, "fc1", maps[l].fc1, f.tx_xn, 4 * E, E, B, f.split_fc1, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0,
                         nullptr, 0, st));
            RQB_TRY(launch_pdl(act_reduce_kernel, dim3((unsigned)std::min<int64_t>(ceil_div((int64_t)B * 4 * E / 4, 256), 1184)), dim3(256),
                               (size_t)0, st, f.use_pdl, (const float*)ws.P, f.split_fc1, bw.b1, ws.Hh, B, 4 * E, f.bf,
                               tr_slot(f, "act_reduce")));
        }
        RQB_TRY(gemm(f, "fc2", maps[l].fc2, f.tx_h, E, 4 * E, B, f.split_fc2, GT_PARTIAL, nullptr, 1.f, nullptr, ws.P, nullptr, 0, nullptr, 0,
                     st));
    }
    // fold the last block's pending fc2 reduction into x (x is final on return) -- and the caller's LayerNorm, if any
    RQB_TRY(ln(f, "finalize", B, x, ws.P, f.split_fc2, blocks.back().b2, nof, x, fin_g, fin_b, fin_g ? ws.XN : nullptr, st));
    return 0;
}

static int record_body(ArFast& f, FastWs& ws, bool cond_token, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, Tb = c.cond_len + HW;
    if (cond_token) {
        RQB_TRY(launch_pdl(cond_tok_kernel, dim3(B, 1), dim3(256), (size_t)0, st, f.use_pdl, (const StepState*)ws.state, w.cond_emb,
                           w.pos_emb_cond, c.cond_len, c.vocab_cond, E, ws.XB));
    } else {
        RQB_TRY(launch_pdl(code_sum_kernel, dim3(B, 1), dim3(64), (size_t)0, st, f.use_pdl, (const StepState*)ws.state, w.codebook, HW,
                           c.D, c.codebook_size, c.code_dim, 0, 0, ws.S, f.bf));
        // x = W_in (sum_d e_d) + D b_in + pos_emb_hw[idx-1]       (bias counted D times, transformers.py:220,225)
        RQB_TRY(gemm(f, "w_in", f.tm_win, f.tx_s, E, c.code_dim, B, 1, GT_F32, w.b_in, (float)c.D, ws.XB, nullptr,
                     w.pos_emb_hw - E /* row idx-1 */, 0, &ws.state->idx, E, st));
    }
    RQB_TRY(fast_stack(f, f.body, f.lbody, ws, ws.XB, nullptr, ws.XB, ws.kc_body, ws.vc_body, Tb, &ws.state->s, 0, nullptr, nullptr, st));
    RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), (size_t)0, st, f.use_pdl, ws.state, 1, 0, 0));
    return 0;
}

static int record_head(ArFast& f, FastWs& ws, bool with_logits, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, D = c.D, V = c.vocab;
    for (int d = 0; d < D; d++) {
        if (d == 0) {
            // token = spatial ctx (body output) + pos_emb_d[0]                                   (transformers.py:259-270)
            RQB_TRY(fast_stack(f, f.head, f.lhead, ws, ws.XH, w.pos_emb_d, ws.XB, ws.kc_head, ws.vc_head, D, nullptr, 0, w.cls_ln_w,
                               w.cls_ln_b, st));
        } else {
            RQB_TRY(launch_pdl(code_sum_kernel, dim3(B, 1), dim3(64), (size_t)0, st, f.use_pdl, (const StepState*)ws.state, w.codebook,
                               HW, D, c.codebook_size, c.code_dim, d, 0, ws.S, f.bf));
            RQB_TRY(gemm(f, "w_head", f.tm_whead, f.tx_s, E, c.code_dim, B, 1, GT_F32, w.b_head, 1.f, ws.XH, nullptr,
                         w.pos_emb_d + (int64_t)d * E, 0, nullptr, 0, st));
            RQB_TRY(fast_stack(f, f.head, f.lhead, ws, ws.XH, nullptr, ws.XH, ws.kc_head, ws.vc_head, D, nullptr, d, w.cls_ln_w,
                               w.cls_ln_b, st));
        }
        // classifier: LN(x) (fused into the stack's last launch) -> logits                       (transformers.py:278-285)
        RQB_TRY(gemm(f, "cls", f.tm_cls, f.tx_xn, V, E, B, 1, GT_F32, w.b_cls, 1.f, ws.LOGITS, nullptr, nullptr, 0, nullptr, 0, st));
        if (with_logits)
            RQB_TRY(launch_pdl(logits_copy_kernel, dim3(64), dim3(256), (size_t)0, st, f.use_pdl, (const StepState*)ws.state,
                               (const float*)ws.LOGITS, d, (int64_t)B * V));
        RQB_TRY(launch_sample_dyn(ws.LOGITS, ws.state, d, B, V, HW, D, st, f.use_pdl));
    }
    RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), (size_t)0, st, f.use_pdl, ws.state, 0, 1, D));
    return 0;
}

static int record(ArFast& f, FastWs& ws, int which, cudaStream_t st) {
    f.tr_base = ws.trace;
    f.tr_next = f.tr_graph_base[which];
    int rc = which == G_COND ? record_body(f, ws, true, st) : which == G_CODE ? record_body(f, ws, false, st)
                                                                                : record_head(f, ws, which == G_HEAD_LOGITS, st);
    // trace slots: every graph owns a quarter of the buffer
    return rc;
}

static int capture(ArFast& f, FastWs& ws, int which, cudaGraphExec_t* out) {
    cudaGraph_t g = nullptr;
    if (!f.cap_stream) RQB_CUDA(cudaStreamCreateWithFlags(&f.cap_stream, cudaStreamNonBlocking));
    cudaStream_t st = f.cap_stream;
    RQB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int64_t before = g_launches;
    int rc = record(f, ws, which, st);
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
    for (int i = 0; i < G_COUNT; i++) {
        if (f.graphs[i]) cudaGraphExecDestroy(f.graphs[i]);
        f.graphs[i] = nullptr;
    }
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
    f->bf = cfg.weight_dtype == RQB200_BF16 ? 1 : 0;
    f->use_graph = !(cfg.flags & RQB200_AR_NO_GRAPH);
    f->use_pdl = !(cfg.flags & RQB200_AR_NO_PDL);
    f->trace = (cfg.flags & RQB200_AR_TRACE) != 0;
    f->trace_w = (cfg.flags & RQB200_AR_TRACE_WEIGHTS) != 0;
    f->l2pf = (cfg.flags & RQB200_AR_L2_PREFETCH) != 0;
    f->deep = !(cfg.flags & RQB200_AR_SHALLOW_RING);
    f->attn4 = !(cfg.flags & RQB200_AR_ATTN_ONE_WARP);
    f->param_prefetch = !(cfg.flags & RQB200_AR_NO_PARAM_PREFETCH);
    f->batched_prefill = !(cfg.flags & RQB200_AR_SEQUENTIAL_PREFILL);
    f->batched_deep = (cfg.flags & RQB200_AR_BATCHED_DEEP_RING) != 0;
    f->batched_streamer = (cfg.flags & RQB200_AR_BATCHED_STREAMER) != 0;
    {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) f->n_sm = n;
    }
    const int nkbE = E / 64;
    f->split_qkv = pick_split(3 * E / 128, nkbE, cfg.split_qkv, f->n_sm);
    f->split_proj = pick_split(E / 128, nkbE, cfg.split_proj, f->n_sm);
    f->split_fc1 = pick_split(4 * E / 128, nkbE, cfg.split_fc1, f->n_sm);
    f->split_fc2 = pick_split(E / 128, 4 * nkbE, cfg.split_fc2, f->n_sm);
    auto mk = [&](const std::vector<rqb200_block_weights>& bl, std::vector<FastLayer>& out) -> int {
        out.resize(bl.size());
        for (size_t l = 0; l < bl.size(); l++) {
            RQB_TRY(make_tmap_weight(&out[l].qkv, bl[l].wqkv, 3 * E, E));
            RQB_TRY(make_tmap_weight(&out[l].proj, bl[l].wproj, E, E));
            RQB_TRY(make_tmap_weight(&out[l].fc1, bl[l].w1, 4 * E, E));
            RQB_TRY(make_tmap_weight(&out[l].fc2, bl[l].w2, E, 4 * E));
        }
        return 0;
    };
    int rc = mk(body, f->lbody);
    if (!rc) rc = mk(head, f->lhead);
    if (!rc) rc = make_tmap_weight(&f->tm_win, w.w_in, E, cfg.code_dim);
    if (!rc) rc = make_tmap_weight(&f->tm_whead, w.w_head, E, cfg.code_dim);
    if (!rc) rc = make_tmap_weight(&f->tm_cls, w.w_cls, cfg.vocab, E);
    if (!rc && w.w_ccls) rc = make_tmap_weight(&f->tm_ccls, w.w_ccls, (cfg.vocab_cond + 127) / 128 * 128, E);
    if (rc) { delete f; return nullptr; }
    for (int i = 0; i <= G_COUNT; i++) f->tr_graph_base[i] = i * (TR_CAP / G_COUNT);
    return f;
}

void ar_fast_destroy(ArFast* f) {
    if (!f) return;
    drop_graphs(*f);
    if (f->cap_stream) cudaStreamDestroy(f->cap_stream);
    delete f;
}

size_t ar_fast_workspace_bytes(const ArFast* f, int B) { return fast_layout(*f, B, nullptr, 0, nullptr); }

// ---- batched passes (prefill, teacher-forced forward): M = G * T token rows, token-major (row = t * G + g), through one stack.
struct BatchBufs {
    float* X;                    // [M, E] residual stream (in / out)
    h16 *XN, *QKV, *ATT, *H;     // [M,E], [M,3E], [M,E], [M,4E] scratch
};

static int stack_batched(ArFast& f, const std::vector<rqb200_block_weights>& blocks, const std::vector<FastLayer>& maps,
                         const BatchBufs& bb, int G, int T, h16* kc, h16* vc, int64_t kv_per_layer, int Tmax, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const int E = c.embed_dim;
    const int64_t M = (int64_t)G * T;
    if (M > (int64_t)1 << 30 || T > PA_MAXT) return fail(RQB200_EINVAL, "ar fast tier: batched pass too large");
    const bool pdl = true;               // the pass is a PDL chain too: a launch's set-up overlaps its predecessor's tail
    CUtensorMap tx_xn, tx_att, tx_h;
    const int bn = gemm_tc_bn((int)std::min<int64_t>(M, 256));
    RQB_TRY(make_tmap_2d(&tx_xn, bb.XN, 1, E, M, (uint64_t)E * 2, 64, bn));
    RQB_TRY(make_tmap_2d(&tx_att, bb.ATT, 1, E, M, (uint64_t)E * 2, 64, bn));
    RQB_TRY(make_tmap_2d(&tx_h, bb.H, 1, 4 * E, M, (uint64_t)E * 8, 64, bn));
    RQB_ENSURE_SMEM(prefill_attn_smem(PA_MAXT), prefill_attn_kernel);
    const float* nof = nullptr;
    // large-M launches: half-depth rings put two CTAs on an SM, so one tile's epilogue overlaps the other's main loop
    const bool save_deep = f.deep;
    if (M > 256 && !f.batched_deep) f.deep = false;
    struct Restore { ArFast& f; bool d; ~Restore() { f.deep = d; } } restore{f, save_deep};
    // M > 256: the persistent rows GEMM (conv_tc.cu: 128 x 256 tiles, double-buffered TMEM, epilogue overlapped with the next
    // tile); M <= 256: the weight streamer
    const bool rows = M > 256 && !f.batched_streamer;
    for (size_t l = 0; l < blocks.size(); l++) {
        const rqb200_block_weights& bw = blocks[l];
        RQB_TRY(ln(f, "", (int)M, bb.X, nof, 0, nof, nof, nullptr, bw.ln1_w, bw.ln1_b, bb.XN, st));
        if (rows) {
            RQB_TRY(launch_rows_gemm_tc(bb.XN, bw.wqkv, bw.bqkv, nullptr, nullptr, bb.QKV, 0, f.bf, M, 3 * E, E, st));
        } else {
            GemmTcParams p = gemm_base(f, 3 * E, E, (int)M, 1, GT_H16);
            p.bias = bw.bqkv; p.out = bb.QKV;
            RQB_TRY(launch_gemm_tc(maps[l].qkv, tx_xn, p, pdl, st));
        }
        h16* kcl = kc ? kc + kv_per_layer * l : nullptr;
        h16* vcl = vc ? vc + kv_per_layer * l : nullptr;
        if (T <= 4) {                                     // tiny groups (the forward's head stack): a warp per (group, head)
            RQB_TRY(launch_pdl(prefill_attn_small_kernel<4>, dim3((unsigned)ceil_div((int64_t)G * c.n_head, 4)), dim3(128), (size_t)0, st, pdl,
                               (const h16*)bb.QKV, kcl, vcl, bb.ATT, G, T, E, c.n_head, Tmax, f.bf));
        } else if (T <= 8) {
            RQB_TRY(launch_pdl(prefill_attn_small_kernel<8>, dim3((unsigned)ceil_div((int64_t)G * c.n_head, 4)), dim3(128), (size_t)0, st, pdl,
                               (const h16*)bb.QKV, kcl, vcl, bb.ATT, G, T, E, c.n_head, Tmax, f.bf));
        } else if (T >= 16 && T <= 64) {                  // one 64-key tile: the mma.sync form
            if (f.bf) {
                RQB_TRY(launch_pdl(prefill_attn_mma_kernel<true>, dim3((unsigned)(G * c.n_head)), dim3(128), (size_t)0, st, pdl,
                                   (const h16*)bb.QKV, kcl, vcl, bb.ATT, G, T, E, c.n_head, Tmax));
            } else {
                RQB_TRY(launch_pdl(prefill_attn_mma_kernel<false>, dim3((unsigned)(G * c.n_head)), dim3(128), (size_t)0, st, pdl,
                                   (const h16*)bb.QKV, kcl, vcl, bb.ATT, G, T, E, c.n_head, Tmax));
            }
        } else {
            RQB_TRY(launch_pdl(prefill_attn_kernel, dim3((unsigned)(G * c.n_head)), dim3(128), prefill_attn_smem(T), st, pdl,
                               (const h16*)bb.QKV, kcl, vcl, bb.ATT, G, T, E, c.n_head, Tmax, f.bf));
        }
        if (rows) {
            RQB_TRY(launch_rows_gemm_tc(bb.ATT, bw.wproj, bw.bproj, bb.X, bb.X, nullptr, 0, f.bf, M, E, E, st));
        } else {
            GemmTcParams p = gemm_base(f, E, E, (int)M, 1, GT_F32);
            p.bias = bw.bproj; p.out = bb.X; p.residual = bb.X; p.ld_res = E;
            RQB_TRY(launch_gemm_tc(maps[l].proj, tx_att, p, pdl, st));
        }
        RQB_TRY(ln(f, "", (int)M, bb.X, nof, 0, nof, nof, nullptr, bw.ln2_w, bw.ln2_b, bb.XN, st));
        if (rows) {
            RQB_TRY(launch_rows_gemm_tc(bb.XN, bw.w1, bw.b1, nullptr, nullptr, bb.H, 1, f.bf, M, 4 * E, E, st));
            RQB_TRY(launch_rows_gemm_tc(bb.H, bw.w2, bw.b2, bb.X, bb.X, nullptr, 0, f.bf, M, E, 4 * E, st));
        } else {
            {
                GemmTcParams p = gemm_base(f, 4 * E, E, (int)M, 1, GT_H16_GELU);
                p.bias = bw.b1; p.out = bb.H;
                RQB_TRY(launch_gemm_tc(maps[l].fc1, tx_xn, p, pdl, st));
            }
            {
                GemmTcParams p = gemm_base(f, E, 4 * E, (int)M, 1, GT_F32);
                p.bias = bw.b2; p.out = bb.X; p.residual = bb.X; p.ld_res = E;
                RQB_TRY(launch_gemm_tc(maps[l].fc2, tx_h, p, pdl, st));
            }
        }
    }
    return 0;
}

// body input tokens [0, T) of every batch row into X (token-major): token s < cond_len is a cond token (transformers.py:224),
// token s >= cond_len carries the summed input embeddings of the codes of position s - cond_len (:219-225)
static int body_tokens_batched(ArFast& f, const StepState* state, float* X, h16* S, int B, int T, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const rqb200_ar_weights& w = f.w;
    const int E = c.embed_dim, HW = c.H * c.W, cl = c.cond_len;
    RQB_TRY(launch_pdl(cond_tok_kernel, dim3(B, std::min(T, cl)), dim3(256), (size_t)0, st, false, state, w.cond_emb, w.pos_emb_cond, cl,
                       c.vocab_cond, E, X));
    const int n_code = T - cl;       // code tokens of positions 0 .. n_code-1
    if (n_code > 0) {
        const int64_t Mc = (int64_t)B * n_code;
        CUtensorMap tx_s;
        RQB_TRY(make_tmap_2d(&tx_s, S, 1, c.code_dim, Mc, (uint64_t)c.code_dim * 2, 64, gemm_tc_bn((int)std::min<int64_t>(Mc, 256))));
        RQB_TRY(launch_pdl(code_sum_kernel, dim3(B, n_code), dim3(64), (size_t)0, st, false, state, w.codebook, HW, c.D, c.codebook_size,
                           c.code_dim, -c.D, 0, S, f.bf));
        GemmTcParams p = gemm_base(f, E, c.code_dim, (int)Mc, 1, GT_F32);
        p.bias = w.b_in; p.bias_scale = (float)c.D; p.out = X + (int64_t)cl * B * E;
        p.residual = w.pos_emb_hw; p.ld_res = E; p.res_div = B;          // row (j, b) gets pos_emb_hw[j]
        RQB_TRY(launch_gemm_tc(f.tm_win, tx_s, p, false, st));
    }
    return 0;
}

// ---- batched prefill: body tokens [0, T) of every batch row in one pass.  Leaves ws.XB = the last token's output rows, the KV
// cache rows [0, T) written, state.s = T.
static int prefill_batched(ArFast& f, FastWs& ws, int T, cudaStream_t st) {
    const rqb200_ar_config& c = f.cfg;
    const int E = c.embed_dim, B = f.B, HW = c.H * c.W, cl = c.cond_len, Tb = cl + HW;
    const int64_t M = (int64_t)B * T;
    if (M > ws.Mmax || T > PA_MAXT) return fail(RQB200_EINVAL, "ar fast tier: prefix too long for the batched prefill");
    const bool save_pdl = f.use_pdl, save_tr = f.trace;
    f.use_pdl = false;
    f.trace = false;
    int rc = [&]() -> int {
        RQB_TRY(body_tokens_batched(f, ws.state, ws.PX, ws.PS, B, T, st));
        BatchBufs bb = {ws.PX, ws.PXN, ws.PQKV, ws.PATT, ws.PH};
        RQB_TRY(stack_batched(f, f.body, f.lbody, bb, B, T, ws.kc_body, ws.vc_body, (int64_t)B * c.n_head * Tb * 64, Tb, st));
        RQB_CUDA(cudaMemcpyAsync(ws.XB, ws.PX + (int64_t)(T - 1) * B * E, (size_t)B * E * sizeof(float), cudaMemcpyDeviceToDevice, st));
        RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), (size_t)0, st, false, ws.state, T, 0, 0));
        return 0;
    }();
    f.use_pdl = save_pdl;
    f.trace = save_tr;
    return rc;
}

// ---- teacher-forced forward (transformers.py:113-188): all H*W*D logits of given code maps in a handful of large-M GEMM passes.
// Body: T = cond_len + H*W - 1 tokens per batch row (M = B*T rows); head: for every (position, batch row) a group of D tokens
// [spatial ctx + pos_d[0], head_mlp(cumsum_{i<d} e_i) + pos_d[d]] (M = D * H*W * B rows, causal attention inside each group).
// logits_out [D][H*W][B][V] f32 (token-major; the host permutes), cond_logits_out (nullable) [cond_len-1][B][vocab_cond rounded
// up to a multiple of 128] (the cond classifier's weight / bias rows are zero-padded to that size by the caller).
struct FwdWs {
    StepState* state;
    float *BX, *HX;
    h16 *XN, *QKV, *ATT, *H, *S;
};
static size_t forward_layout(const ArFast& f, int B, void* base, size_t cap, FwdWs* out) {
    const rqb200_ar_config& c = f.cfg;
    Arena a(base, cap);
    const int64_t E = c.embed_dim, HW = (int64_t)c.H * c.W, Tb = c.cond_len + HW - 1;
    const int64_t Mb = B * Tb, Mh = (int64_t)c.D * HW * B, Mm = std::max(Mb, Mh) + 128;   // (+128: whole-tile reads of the rows GEMM)
    FwdWs w;
    w.state = a.take<StepState>(1);
    w.BX = a.take<float>(Mb * E);
    w.HX = a.take<float>(Mh * E);
    w.XN = a.take<h16>(Mm * E);
    w.QKV = a.take<h16>(Mm * 3 * E);
    w.ATT = a.take<h16>(Mm * E);
    w.H = a.take<h16>(Mm * 4 * E);
    w.S = a.take<h16>(HW * B * c.code_dim);
    if (out) *out = w;
    return a.off + 256;
}
size_t ar_fast_forward_workspace_bytes(const ArFast* f, int B) { return forward_layout(*f, B, nullptr, 0, nullptr); }

int ar_fast_forward(ArFast* f, const int64_t* codes, const int64_t* cond, int B, float* logits_out, float* cond_logits_out, void* wsp,
                    size_t ws_bytes, cudaStream_t st) {
    const rqb200_ar_config& c = f->cfg;
    const rqb200_ar_weights& w = f->w;
    const int E = c.embed_dim, D = c.D, HW = c.H * c.W, cl = c.cond_len, V = c.vocab, Tb = cl + HW - 1;
    if (B < 1) return fail(RQB200_EINVAL, "ar_forward: B must be > 0");
    if (Tb > PA_MAXT) return fail(RQB200_EINVAL, "ar_forward: sequence too long for the batched attention kernel");
    if (cond_logits_out && (cl < 2 || !w.w_ccls)) return fail(RQB200_EINVAL, "ar_forward: cond logits need cond_len > 1 and a cond classifier");
    FwdWs ws;
    if (forward_layout(*f, B, wsp, ws_bytes, &ws) > ws_bytes) return fail(RQB200_EWORKSPACE, "ar_forward: workspace too small");
    StepState h = {};
    h.cond = cond; h.codes = const_cast<int64_t*>(codes);
    RQB_TRY(launch_pdl(init_state_kernel, dim3(1), dim3(32), (size_t)0, st, false, ws.state, h, 0));
    const bool save_pdl = f->use_pdl, save_tr = f->trace;
    f->use_pdl = false;
    f->trace = false;
    const float* nof = nullptr;
    int rc = [&]() -> int {
        const int G = HW * B;                                  // head groups, g = pos * B + b
        const int64_t Mh = (int64_t)D * G;
        // body
        RQB_TRY(body_tokens_batched(*f, ws.state, ws.BX, ws.S, B, Tb, st));
        BatchBufs bb = {ws.BX, ws.XN, ws.QKV, ws.ATT, ws.H};
        RQB_TRY(stack_batched(*f, f->body, f->lbody, bb, B, Tb, nullptr, nullptr, 0, Tb, st));
        if (cond_logits_out) {                                  // cond_classifier(latents[:, :cond_len-1])        (:153-156)
            const int64_t Mc = (int64_t)(cl - 1) * B;
            CUtensorMap tx;
            RQB_TRY(make_tmap_2d(&tx, ws.XN, 1, E, Mc, (uint64_t)E * 2, 64, gemm_tc_bn((int)std::min<int64_t>(Mc, 256))));
            RQB_TRY(ln(*f, "", (int)Mc, ws.BX, nof, 0, nof, nof, nullptr, w.ccls_ln_w, w.ccls_ln_b, ws.XN, st));
            GemmTcParams p = gemm_base(*f, (c.vocab_cond + 127) / 128 * 128, E, (int)Mc, 1, GT_F32);
            p.bias = w.b_ccls; p.out = cond_logits_out;
            RQB_TRY(launch_gemm_tc(f->tm_ccls, tx, p, false, st));
        }
        // head tokens: d = 0 rows = spatial ctx (body rows of tokens cond_len-1 ..) + pos_emb_d[0]; d >= 1 rows = head_mlp(cumsum)
        RQB_TRY(ln(*f, "", G, ws.BX + (int64_t)(cl - 1) * B * E, nof, 0, nof, w.pos_emb_d, ws.HX, nof, nof, nullptr, st));
        CUtensorMap tx_s;
        RQB_TRY(make_tmap_2d(&tx_s, ws.S, 1, c.code_dim, G, (uint64_t)c.code_dim * 2, 64, gemm_tc_bn(std::min(G, 256))));
        for (int d = 1; d < D; d++) {
            RQB_TRY(launch_pdl(code_sum_kernel, dim3(B, HW), dim3(64), (size_t)0, st, false, (const StepState*)ws.state, w.codebook, HW, D,
                               c.codebook_size, c.code_dim, -d, 0, ws.S, f->bf));
            GemmTcParams p = gemm_base(*f, E, c.code_dim, G, 1, GT_F32);
            p.bias = w.b_head; p.out = ws.HX + (int64_t)d * G * E; p.residual = w.pos_emb_d + (int64_t)d * E; p.ld_res = 0;
            RQB_TRY(launch_gemm_tc(f->tm_whead, tx_s, p, false, st));
        }
        BatchBufs hb = {ws.HX, ws.XN, ws.QKV, ws.ATT, ws.H};
        RQB_TRY(stack_batched(*f, f->head, f->lhead, hb, G, D, nullptr, nullptr, 0, D, st));
        // classifier                                                                                   (:181-183)
        CUtensorMap tx;
        RQB_TRY(make_tmap_2d(&tx, ws.XN, 1, E, Mh, (uint64_t)E * 2, 64, gemm_tc_bn((int)std::min<int64_t>(Mh, 256))));
        RQB_TRY(ln(*f, "", (int)Mh, ws.HX, nof, 0, nof, nof, nullptr, w.cls_ln_w, w.cls_ln_b, ws.XN, st));
        if (Mh > 256 && !f->batched_streamer) {
            RQB_TRY(launch_rows_gemm_tc(ws.XN, w.w_cls, w.b_cls, nullptr, logits_out, nullptr, 0, f->bf, Mh, V, E, st));
        } else {
            GemmTcParams p = gemm_base(*f, V, E, (int)Mh, 1, GT_F32);
            p.bias = w.b_cls; p.out = logits_out;
            RQB_TRY(launch_gemm_tc(f->tm_cls, tx, p, false, st));
        }
        return 0;
    }();
    f->use_pdl = save_pdl;
    f->trace = save_tr;
    return rc;
}

int ar_fast_sample(ArFast* f, const int64_t* partial, const int64_t* cond, int B, int idx_begin, int idx_end, int resume,
                   float temperature, const int32_t* top_k, const float* top_p, const float* noise, int64_t noise_stride,
                   float* logits_out, const int64_t* force, int64_t* out, void* wsp, size_t ws_bytes, cudaStream_t st) {
    const rqb200_ar_config& c = f->cfg;
    const int E = c.embed_dim, D = c.D, HW = c.H * c.W, cl = c.cond_len;
    if (B < 1 || B > 256) return fail(RQB200_EINVAL, "ar fast tier: batch must be in [1,256] per call");
    if (idx_begin < 0 || idx_end > HW || idx_begin > idx_end) return fail(RQB200_EINVAL, "ar_sample: bad position span");
    FastWs ws;
    size_t need = fast_layout(*f, B, wsp, ws_bytes, &ws);
    if (need > ws_bytes) return fail(RQB200_EWORKSPACE, "ar_sample: workspace too small");
    if (!resume && out != partial)
        RQB_CUDA(cudaMemcpyAsync(out, partial, (size_t)B * HW * D * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    if (idx_begin >= idx_end) return 0;
    if (f->ws_base != wsp || f->B != B) {        // (re)bind activation tensor maps + graphs to this workspace
        if (resume) return fail(RQB200_ESTATE, "ar_sample: resume on a workspace / batch the engine is not bound to");
        drop_graphs(*f);
        f->ws_base = wsp;
        f->B = B;
        const int bn = gemm_tc_bn(B);
        RQB_TRY(make_tmap_2d(&f->tx_xn, ws.XN, 1, E, B, (uint64_t)E * 2, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_att, ws.ATT, 1, E, B, (uint64_t)E * 2, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_h, ws.Hh, 1, 4 * E, B, (uint64_t)E * 8, 64, bn));
        RQB_TRY(make_tmap_2d(&f->tx_s, ws.S, 1, c.code_dim, B, (uint64_t)c.code_dim * 2, 64, bn));
    }
    StepState h = {};
    h.s = 0; h.idx = 0; h.step = 0;
    h.cond = cond; h.codes = out; h.force = force; h.noise = noise; h.logits_out = logits_out; h.noise_stride = noise_stride;
    h.temperature = temperature;
    for (int d = 0; d < D; d++) { h.top_k[d] = top_k[d]; h.top_p[d] = top_p[d]; }
    RQB_TRY(launch_pdl(init_state_kernel, dim3(1), dim3(32), (size_t)0, st, false, ws.state, h, resume ? 1 : 0));
    if (f->trace && !resume) RQB_CUDA(cudaMemsetAsync(ws.trace, 0, (size_t)4 * TR_CAP * sizeof(long long), st));
    auto run = [&](int which) -> int {
        if (!f->use_graph) return record(*f, ws, which, st);
        if (!f->graphs[which]) RQB_TRY(capture(*f, ws, which, &f->graphs[which]));
        RQB_CUDA(cudaGraphLaunch(f->graphs[which], st));
        g_launches += f->n_nodes[which];     // kernels executed by this replay
        return 0;
    };
    const int head_graph = logits_out ? G_HEAD_LOGITS : G_HEAD;
    if (!resume) {
        // prefill: cond tokens, then (start_loc resume) the code tokens of positions < idx_begin  (transformers.py:237-239)
        const int T0 = cl + idx_begin;
        if (f->batched_prefill && T0 >= 4 && T0 <= PA_MAXT && (int64_t)B * T0 <= ws.Mmax) {
            RQB_TRY(prefill_batched(*f, ws, T0, st));
            // state.idx must equal idx_begin for the first head graph
            if (idx_begin > 0) RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), (size_t)0, st, false, ws.state, 0, idx_begin, 0));
        } else {
            // one cached step each -- causal, so identical to the batched form
            for (int s = 0; s < cl; s++) RQB_TRY(run(G_COND));
            // state.idx must equal (position whose codes feed the body) + 1 while replaying the code-token graph
            for (int j = 1; j <= idx_begin; j++) {
                RQB_TRY(launch_pdl(advance_kernel, dim3(1), dim3(32), (size_t)0, st, false, ws.state, 0, 1, 0));
                RQB_TRY(run(G_CODE));
            }
        }
    }
    for (int idx = idx_begin; idx < idx_end; idx++) {
        if (idx > idx_begin || resume) RQB_TRY(run(G_CODE));   // body step on the token of position idx-1 (state.idx == idx)
        RQB_TRY(run(head_graph));                              // D head steps + sampling; advances idx, step
    }
    return 0;
}

int ar_fast_trace(ArFast* f, long long* out_host, int cap_launches, char* names, int names_cap) {
    if (!f || !f->trace || !f->tr_base) return 0;
    const int n = std::min(cap_launches, TR_CAP);
    if (cudaMemcpy(out_host, f->tr_base, (size_t)n * 4 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    std::string all;
    for (int i = 0; i < n; i++) {
        all += i < (int)f->tr_names.size() ? f->tr_names[i] : "";
        all += '\n';
    }
    if (names && names_cap > 0) {
        const size_t k = std::min(all.size(), (size_t)names_cap - 1);
        memcpy(names, all.data(), k);
        names[k] = 0;
    }
    return n;
}

}  // namespace rqb
