// P3 "fast" tier, persistent form -- one kernel launch walks a whole transformer stack for the new token.
//
// The PDL-chained per-op kernels of ar_fast.cu are latency-bound: every GEMM of the decode step moves only 5-19 MB, so
// launch + prologue (barrier init, TMEM alloc, tensor-map fetch) + first-byte latency + tail dominate its ~2.5 us of HBM
// time.  This kernel removes all of that:
//   * grid = one CTA per SM, resident for the whole stack (296 phases for the 42-layer body step of the 1.4B model);
//   * warp 0 is a free-running WEIGHT PRODUCER: it walks the phase table ahead of everybody else and keeps a 9-stage,
//     144 KB TMA ring per SM full of the weight tiles this CTA will need next -- weights do not depend on activations,
//     so HBM streaming never stops at a phase boundary (148 SMs x 144 KB = 21 MB in flight: more than a whole GEMM);
//   * phases (LN+split-K-reduce | GEMM | attention | ...) are separated by a grid barrier (one release-add + acquire
//     spin on a monotonic counter); the MMA warp only ever waits for ACTIVATIONS, never for weights;
//   * GEMM phases are the same swap-AB tcgen05 tiles as gemm_tc.cu (128 output features x 64 batch columns x 16, bf16,
//     fp32 accumulate in TMEM), split-K partials are reduced in fixed order by the next phase (deterministic).
// Reference semantics: attentions.py:134-142 per block, transformers.py:190-287 per token (see ar_fast.cu).
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace rqb {

constexpr int MG_THREADS = 256;
constexpr int MG_WSTAGES = 9;
constexpr int MG_XSTAGES = 4;
constexpr int MG_BN = 64;
constexpr int MG_A_BYTES = 128 * 64 * 2;
constexpr int MG_X_BYTES = MG_BN * 64 * 2;
constexpr int MG_WORK = 224;              // threads of warps 1..7

__device__ __forceinline__ void work_sync() { asm volatile("bar.sync 1, 224;" ::: "memory"); }

// Grid barrier for the worker warps.  bar.sync gives CTA-scope happens-before from every worker's stores to thread 0; its
// gpu-scope release-add / acquire-spin then carries them (cumulativity) to every other CTA; the trailing bar.sync hands the
// acquired view to the CTA's other threads.  Generic -> async-proxy ordering for TMA reads of freshly written operands is
// established by a .global proxy fence on both sides (writers before arriving, the TMA-issuing thread after the barrier).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, int wt) {
    __syncwarp();                                              // bar.sync is warp-aligned: reconverge single-lane roles first
    asm volatile("fence.proxy.async.global;" ::: "memory");
    work_sync();
    if (wt == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while ((int)(v - target) < 0);
    }
    work_sync();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// block reductions over the 224 worker threads (warps 1..7), scratch >= 8 floats
__device__ __forceinline__ float work_sum(float v, float* scratch, int wt) {
    v = warp_sum(v);
    if ((wt & 31) == 0) scratch[wt >> 5] = v;     // callers alternate between two scratch halves; a grid barrier separates reuse
    work_sync();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) r += scratch[i];
    return r;
}

// NOTE on all phase_* functions: every field of the (global-memory) phase descriptor is copied into registers first --
// the descriptor may alias the stores below as far as the compiler knows, and re-loading fields after every store was
// the dominant cost of the first version of this kernel.
__device__ __forceinline__ void f4add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// Each worker thread owns up to LN_SLOTS float4 chunks of the row (E <= 224*4*LN_SLOTS); all of a thread's loads (x, bias,
// <= 12 split-K partials per chunk) are issued before the first dependent add, and the row never leaves registers.
constexpr int LN_SLOTS = 3;
__device__ void phase_ln(const MPhase& phr, int B, int E, float* row, float* scratch, int wt) {
    const float* x_in = phr.x_in; const float* partial = phr.partial; const int S = phr.S; const float* bias = phr.bias;
    const float* extra = phr.extra; float* x_out = phr.x_out; const float* g = phr.g; const float* be = phr.be;
    __nv_bfloat16* xn = phr.xn;
    const int E4 = E >> 2;
    const int S12 = S < 12 ? S : 12;
    (void)row;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        float4 v[LN_SLOTS];
        float4 pr[LN_SLOTS][12];
#pragma unroll
        for (int k = 0; k < LN_SLOTS; k++) {
            const int e4 = wt + k * MG_WORK;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e4 < E4) {
                if (x_in) v[k] = ldcg4(x_in + (int64_t)b * E + e4 * 4);
#pragma unroll
                for (int i = 0; i < 12; i++)
                    if (i < S12) pr[k][i] = ldcg4(partial + ((int64_t)i * B + b) * E + e4 * 4);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < LN_SLOTS; k++) {
            const int e4 = wt + k * MG_WORK;
            if (e4 < E4) {
                if (bias) f4add(v[k], *reinterpret_cast<const float4*>(bias + e4 * 4));
#pragma unroll
                for (int i = 0; i < 12; i++)
                    if (i < S12) f4add(v[k], pr[k][i]);
                for (int i = 12; i < S; i++) f4add(v[k], ldcg4(partial + ((int64_t)i * B + b) * E + e4 * 4));
                if (extra) f4add(v[k], *reinterpret_cast<const float4*>(extra + e4 * 4));
                if (x_out) *reinterpret_cast<float4*>(x_out + (int64_t)b * E + e4 * 4) = v[k];
                s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            }
        }
        const float mean = work_sum(s, scratch, wt) / (float)E;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < LN_SLOTS; k++)
            if (wt + k * MG_WORK < E4) {
                float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
                q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
            }
        const float rstd = rsqrtf(work_sum(q, scratch + 8, wt) / (float)E + 1e-5f);
        if (xn) {
#pragma unroll
            for (int k = 0; k < LN_SLOTS; k++) {
                const int e4 = wt + k * MG_WORK;
                if (e4 < E4) {
                    const float4 gg = *reinterpret_cast<const float4*>(g + e4 * 4), bb = *reinterpret_cast<const float4*>(be + e4 * 4);
                    __nv_bfloat162 h0 = __floats2bfloat162_rn((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn((v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<unsigned*>(&h0);
                    pk.y = *reinterpret_cast<unsigned*>(&h1);
                    *reinterpret_cast<uint2*>(xn + (int64_t)b * E + e4 * 4) = pk;
                }
            }
        }
    }
}

// one warp per (b, head); identical arithmetic to attn_fast_kernel (ar_fast.cu)
struct AttnArgs {
    const float* apart; int aS; const float* bqkv; __nv_bfloat16 *kc, *vc, *att; int Tmax; const int* t_ptr; int t_host;
};
__device__ void phase_attn(const MPhase& phr, int B, int E, int nh, float* qs_all, float* ps_all, int wt) {
    const AttnArgs ph = {phr.apart, phr.aS, phr.bqkv, phr.kc, phr.vc, phr.att, phr.Tmax, phr.t_ptr, phr.t_host};
    const int lane = wt & 31, w = wt >> 5;                    // w in 0..6
    float* qs = qs_all + w * 64;
    float* ps = ps_all + w * 512;
    const int t = ph.t_ptr ? *ph.t_ptr : ph.t_host;
    const int total = B * nh;
    for (int bh = blockIdx.x * 7 + w; bh < total; bh += gridDim.x * 7) {
        const int b = bh / nh, h = bh % nh;
        const int c = h * 64 + 2 * lane;
        float2 q = make_float2(ph.bqkv[c], ph.bqkv[c + 1]);
        float2 k = make_float2(ph.bqkv[E + c], ph.bqkv[E + c + 1]);
        float2 v = make_float2(ph.bqkv[2 * E + c], ph.bqkv[2 * E + c + 1]);
#pragma unroll 4
        for (int s = 0; s < ph.aS; s++) {
            const float* p = ph.apart + ((int64_t)s * B + b) * 3 * E;
            float2 a = __ldcg(reinterpret_cast<const float2*>(p + c));
            float2 bb = __ldcg(reinterpret_cast<const float2*>(p + E + c));
            float2 cc = __ldcg(reinterpret_cast<const float2*>(p + 2 * E + c));
            q.x += a.x; q.y += a.y; k.x += bb.x; k.y += bb.y; v.x += cc.x; v.y += cc.y;
        }
        __nv_bfloat16* kb = ph.kc + ((int64_t)(b * nh + h) * ph.Tmax) * 64;
        __nv_bfloat16* vb = ph.vc + ((int64_t)(b * nh + h) * ph.Tmax) * 64;
        const __nv_bfloat162 k2 = __floats2bfloat162_rn(k.x, k.y), v2 = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(kb + (int64_t)t * 64 + 2 * lane) = k2;
        *reinterpret_cast<__nv_bfloat162*>(vb + (int64_t)t * 64 + 2 * lane) = v2;
        const float2 qf = __bfloat1622float2(__floats2bfloat162_rn(q.x, q.y)), kf = __bfloat1622float2(k2), vf = __bfloat1622float2(v2);
        __syncwarp();
        qs[2 * lane] = qf.x;
        qs[2 * lane + 1] = qf.y;
        __syncwarp();
        const float s_new = warp_sum(qf.x * kf.x + qf.y * kf.y) * 0.125f;
        float m = s_new;
        for (int j = lane; j < t; j += 32) {
            const uint4* kr = reinterpret_cast<const uint4*>(kb + (int64_t)j * 64);
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                uint4 wv = __ldcg(kr + u);
                const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    float2 kk = __bfloat1622float2(kp[z]);
                    acc = fmaf(qs[(u * 4 + z) * 2], kk.x, acc);
                    acc = fmaf(qs[(u * 4 + z) * 2 + 1], kk.y, acc);
                }
            }
            acc *= 0.125f;
            ps[j] = acc;
            m = fmaxf(m, acc);
        }
        m = warp_max(m);
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
        int j = 0;
        for (; j + 8 <= t; j += 8) {
            unsigned raw[8];
#pragma unroll
            for (int u = 0; u < 8; u++) raw[u] = __ldcg(reinterpret_cast<const unsigned*>(vb + (int64_t)(j + u) * 64 + 2 * lane));
#pragma unroll
            for (int u = 0; u < 8; u++) {
                float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u]));
                o.x = fmaf(ps[j + u], vv.x, o.x);
                o.y = fmaf(ps[j + u], vv.y, o.y);
            }
        }
        for (; j < t; j++) {
            const unsigned raw = __ldcg(reinterpret_cast<const unsigned*>(vb + (int64_t)j * 64 + 2 * lane));
            float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            o.x = fmaf(ps[j], vv.x, o.x);
            o.y = fmaf(ps[j], vv.y, o.y);
        }
        *reinterpret_cast<__nv_bfloat162*>(ph.att + (int64_t)b * E + c) = __floats2bfloat162_rn(o.x * inv, o.y * inv);
    }
}

// h = bf16(gelu(sum_s partial[s] + bias)) -- only when fc1 runs split-K
struct ActArgs { int N_out, splits; const float* gbias; const float* gpartial; void* out; };
__device__ void phase_act(const MPhase& phr, int B, int wt) {
    const ActArgs ph = {phr.N_out, phr.splits, phr.gbias, phr.gpartial, phr.out};
    const int N = ph.N_out;
    const int64_t total4 = (int64_t)B * N / 4;
    for (int64_t i = (int64_t)blockIdx.x * MG_WORK + wt; i < total4; i += (int64_t)gridDim.x * MG_WORK) {
        const int n = (int)((i * 4) % N);
        float4 v = *reinterpret_cast<const float4*>(ph.gbias + n);
#pragma unroll 3
        for (int s = 0; s < ph.splits; s++) {
            float4 p = ldcg4(ph.gpartial + (int64_t)s * B * N + i * 4);
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = 0.5f * r[k] * (1.0f + erff(r[k] * 0.70710678118654752440f));
        __nv_bfloat162 h0 = __floats2bfloat162_rn(r[0], r[1]), h1 = __floats2bfloat162_rn(r[2], r[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&h0);
        pk.y = *reinterpret_cast<unsigned*>(&h1);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ph.out) + i * 4) = pk;
    }
}

__device__ void phase_codesum(const MPhase& phr, const MegaParams& P, int wt) {
    struct { int cs_mode; __nv_bfloat16* cs_out; } ph = {phr.cs_mode, phr.cs_out};
    const StepState* stt = P.stt;
    const int pos = ph.cs_mode == 0 ? stt->idx - 1 : stt->idx;
    const int nd = ph.cs_mode == 0 ? P.D : ph.cs_mode;
    const int64_t* codes = stt->codes;
    for (int b = blockIdx.x; b < P.B; b += gridDim.x)
        for (int c = wt; c < P.C; c += MG_WORK) {
            float a = 0.f;
            for (int i = 0; i < nd; i++) {
                long long k = codes[((int64_t)b * P.HW + pos) * P.D + i];
                k = k < 0 ? 0 : (k >= P.Kc ? P.Kc - 1 : k);
                a += P.codebook[k * P.C + c];
            }
            ph.cs_out[(int64_t)b * P.C + c] = __float2bfloat16(a);
        }
}

__global__ void __launch_bounds__(MG_THREADS, 1) ar_mega_kernel(MegaParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wring = smem;
    uint8_t* xring = wring + MG_WSTAGES * MG_A_BYTES;
    float* row = reinterpret_cast<float*>(xring + MG_XSTAGES * MG_X_BYTES);       // [E]
    float* qs = row + P.E;                                                        // [7*64]
    float* ps = qs + 7 * 64;                                                      // [7*512]
    float* scratch = ps + 7 * 512;                                                // [8]
    MPhase* dsm = reinterpret_cast<MPhase*>(scratch + 8 + 8);                     // [2] staged phase descriptors (64 B aligned)
    uint64_t* full_w = reinterpret_cast<uint64_t*>(dsm + 2);
    uint64_t* empty_w = full_w + MG_WSTAGES;
    uint64_t* full_x = empty_w + MG_WSTAGES;
    uint64_t* empty_x = full_x + MG_XSTAGES;
    uint64_t* tmem_full = empty_x + MG_XSTAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MG_WSTAGES; s++) { tc::mbar_init(&full_w[s], 1); tc::mbar_init(&empty_w[s], 1); }
        for (int s = 0; s < MG_XSTAGES; s++) { tc::mbar_init(&full_x[s], 1); tc::mbar_init(&empty_x[s], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::mbar_init(tmem_empty, 4);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= weight producer: runs ahead through the whole phase table =================
        // lane 0 streams weight tiles; the other lanes warm L2 with the small vectors (biases, LayerNorm affine) and the
        // descriptor lines of upcoming phases so that the workers' dependent loads hit L2 instead of HBM
        uint32_t wc = 0;
        for (int pi = 0; pi < P.n_phases; pi++) {
            const MPhase& ph = P.phases[pi];
            const int type = ph.type;
            if (lane >= 1) {
                auto pf = [&](const void* base_ptr, int bytes) {
                    if (base_ptr == nullptr) return;
                    for (int o = (lane - 1) * 128; o < bytes; o += 31 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base_ptr) + o));
                };
                if (pi + 4 < P.n_phases) pf(&P.phases[pi + 4], (int)sizeof(MPhase));
                if (type == MP_LN) { pf(ph.bias, P.E * 4); pf(ph.g, P.E * 4); pf(ph.be, P.E * 4); pf(ph.extra, P.E * 4); }
                else if (type == MP_ATTN) pf(ph.bqkv, 3 * P.E * 4);
                else if (type == MP_ACT || type == MP_GEMM) pf(ph.gbias, ph.N_out * 4);
            }
            if (type != MP_GEMM) continue;
            if (lane == 0) {
                tc::prefetch_tmap(&ph.tmX);
                tc::prefetch_tmap(&ph.tmW);
                const int splits = ph.splits;
                const int n_units = (ph.N_out / 128) * splits, nkb_total = ph.K / 64;
                for (int u = cta; u < n_units; u += G) {
                    const int tile = u / splits, split = u % splits;
                    const int kb0 = (int)((int64_t)nkb_total * split / splits), kb1 = (int)((int64_t)nkb_total * (split + 1) / splits);
                    for (int kb = kb0; kb < kb1; kb++, wc++) {
                        const int s = wc % MG_WSTAGES;
                        tc::mbar_wait(&empty_w[s], ((wc / MG_WSTAGES) & 1) ^ 1);
                        tc::mbar_expect_tx(&full_w[s], MG_A_BYTES);
                        tc::tma_load_2d(wring + s * MG_A_BYTES, &ph.tmW, &full_w[s], ph.w_tiled ? 0 : kb * 64,
                                        ph.w_tiled ? (tile * nkb_total + kb) * 128 : tile * 128, tc::L2_EVICT_FIRST);
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // ================= workers (warps 1..7): walk the phases in lock step =================
        const int wt = threadIdx.x - 32;
        unsigned base = 0;
        if (wt == 0) base = *reinterpret_cast<volatile unsigned*>(&P.bar[1]);
        uint32_t wc = 0, xc = 0, uc = 0;
        constexpr uint32_t idesc = tc::umma_idesc(128, MG_BN, 1);
        auto stamp = [&](int i) {
            if (P.trace && cta == 0 && wt == 0) {
                long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                P.trace[i] = t;
            }
        };
        stamp(0);
        auto stage_desc = [&](int pi) {       // copy the scalar part of descriptor pi into its SMEM slot (tensor maps stay in global)
            if (pi < P.n_phases) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.phases[pi]) + 64;        // skip 2 x 128 B tensor maps
                uint32_t* dst = reinterpret_cast<uint32_t*>(&dsm[pi & 1]) + 64;
                constexpr int NW = (int)(sizeof(MPhase) / 4) - 64;
                if (wt < NW) dst[wt] = __ldcg(src + wt);
            }
        };
        stage_desc(0);
        work_sync();
        for (int pi = 0; pi < P.n_phases; pi++) {
            const MPhase& ph = dsm[pi & 1];
            const MPhase& phg = P.phases[pi];   // global copy: tensor maps
            stage_desc(pi + 1);                 // loads in flight during this phase; consumed after the next barrier
            if (ph.type == MP_LN) {
                phase_ln(ph, P.B, P.E, row, scratch, wt);
            } else if (ph.type == MP_ATTN) {
                phase_attn(ph, P.B, P.E, P.nh, qs, ps, wt);
            } else if (ph.type == MP_CODESUM) {
                phase_codesum(ph, P, wt);
            } else if (ph.type == MP_ACT) {
                phase_act(ph, P.B, wt);
            } else {
                struct {
                    int N_out, K, splits, mode; const float* gbias; float bias_scale; const float* res; const int* res_row_ptr;
                    long long res_row_stride, ld_res; void* out; float* gpartial;
                } gp = {ph.N_out, ph.K, ph.splits, ph.mode, ph.gbias, ph.bias_scale, ph.res, ph.res_row_ptr, ph.res_row_stride,
                        ph.ld_res, ph.out, ph.gpartial};
                const CUtensorMap* tmXp = &phg.tmX;
                const int n_units = (gp.N_out / 128) * gp.splits, nkb_total = gp.K / 64;
                const int PB = P.B;
                if (warp == 6) {
                    // ---- activation producer
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.global;" ::: "memory");   // this thread's TMA reads data other CTAs just wrote
                        for (int u = cta; u < n_units; u += G) {
                            const int split = u % gp.splits;
                            const int kb0 = (int)((int64_t)nkb_total * split / gp.splits), kb1 = (int)((int64_t)nkb_total * (split + 1) / gp.splits);
                            for (int kb = kb0; kb < kb1; kb++, xc++) {
                                const int s = xc % MG_XSTAGES;
                                tc::mbar_wait(&empty_x[s], ((xc / MG_XSTAGES) & 1) ^ 1);
                                tc::mbar_expect_tx(&full_x[s], MG_X_BYTES);
                                tc::tma_load_2d(xring + s * MG_X_BYTES, tmXp, &full_x[s], kb * 64, 0, tc::L2_EVICT_LAST);
                            }
                        }
                    }
                } else if (warp == 1) {
                    // ---- MMA issuer
                    for (int u = cta; u < n_units; u += G, uc++) {
                        const int split = u % gp.splits;
                        const int kb0 = (int)((int64_t)nkb_total * split / gp.splits), kb1 = (int)((int64_t)nkb_total * (split + 1) / gp.splits);
                        tc::mbar_wait(tmem_empty, (uc & 1) ^ 1);
                        tc::tc_fence_after();
                        for (int kb = kb0; kb < kb1; kb++, wc++, xc++) {
                            const int sw = wc % MG_WSTAGES, sx = xc % MG_XSTAGES;
                            tc::mbar_wait(&full_w[sw], (wc / MG_WSTAGES) & 1);
                            tc::mbar_wait(&full_x[sx], (xc / MG_XSTAGES) & 1);
                            tc::tc_fence_after();
                            if (lane == 0) {
                                const uint32_t a = tc::smem_u32(wring + sw * MG_A_BYTES), b = tc::smem_u32(xring + sx * MG_X_BYTES);
#pragma unroll
                                for (int j = 0; j < 4; j++)
                                    tc::umma_f16(tmem_base, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + j * 32), idesc,
                                                 (kb > kb0 || j > 0) ? 1u : 0u);
                                tc::umma_commit(&empty_w[sw]);
                                tc::umma_commit(&empty_x[sx]);
                                if (kb == kb1 - 1) tc::umma_commit(tmem_full);
                            }
                            __syncwarp();
                        }
                    }
                } else if (warp >= 2 && warp <= 5) {
                    // ---- epilogue
                    const int q = warp & 3;
                    for (int u = cta; u < n_units; u += G, uc++) {
                        const int tile = u / gp.splits, split = u % gp.splits;
                        const int n = tile * 128 + q * 32 + lane;
                        tc::mbar_wait(tmem_full, uc & 1);
                        tc::tc_fence_after();
                        const float bias = (gp.gbias != nullptr && gp.mode != GT_PARTIAL) ? gp.gbias[n] * gp.bias_scale : 0.f;
                        const float* res = nullptr;
                        if (gp.mode == GT_F32 && gp.res != nullptr)
                            res = gp.res + (gp.res_row_ptr ? (int64_t)(*gp.res_row_ptr) * gp.res_row_stride : 0);
                        uint32_t r4[4][16];
#pragma unroll
                        for (int c = 0; c < 4; c++) tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r4[c]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const int c0 = c * 16;
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                const int b = c0 + i;
                                if (b >= PB) break;
                                float v = __uint_as_float(r4[c][i]) + bias;
                                if (gp.mode == GT_PARTIAL) {
                                    gp.gpartial[((int64_t)split * PB + b) * gp.N_out + n] = v;
                                } else if (gp.mode == GT_BF16_GELU) {
                                    reinterpret_cast<__nv_bfloat16*>(gp.out)[(int64_t)b * gp.N_out + n] =
                                        __float2bfloat16(0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)));
                                } else {
                                    if (res) v += res[(int64_t)b * gp.ld_res + n];
                                    reinterpret_cast<float*>(gp.out)[(int64_t)b * gp.N_out + n] = v;
                                }
                            }
                        }
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(tmem_empty);
                    }
                }
                // keep the per-thread ring counters of every worker warp in step (each role advanced only its own)
                {
                    uint32_t kbs = 0, us = 0;
                    for (int u = cta; u < n_units; u += G) {
                        const int split = u % gp.splits;
                        kbs += (uint32_t)((int64_t)nkb_total * (split + 1) / gp.splits - (int64_t)nkb_total * split / gp.splits);
                        us++;
                    }
                    if (warp != 1) wc += kbs;
                    if (warp != 1 && warp != 6) xc += kbs;
                    if (warp == 6 || warp == 7) uc += us;
                }
            }
            stamp(2 * pi + 1);                                                       // own work done
            grid_barrier(&P.bar[0], base + (unsigned)(pi + 1) * (unsigned)G, wt);   // `base` is only meaningful in wt == 0
            stamp(2 * pi + 2);
        }
        if (cta == 0 && wt == 0) P.bar[1] = base + (unsigned)P.n_phases * (unsigned)G;
    }
    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------ host side
size_t mega_smem_bytes(int E) {
    return (size_t)MG_WSTAGES * MG_A_BYTES + (size_t)MG_XSTAGES * MG_X_BYTES + (size_t)(E + 7 * 64 + 7 * 512 + 16) * 4 +
           2 * sizeof(MPhase) + (2 * MG_WSTAGES + 2 * MG_XSTAGES + 2) * 8 + 16 + 1024 + 64;
}

int launch_ar_mega(const MegaParams& P, int n_sm, cudaStream_t st) {
    const size_t smem = mega_smem_bytes(P.E);
    if (smem > 227 * 1024) return fail(RQB200_EINVAL, "ar_mega: shared memory budget exceeded (embed_dim too large)");
    static size_t attr = 0;
    if (smem > attr) {
        RQB_CUDA(cudaFuncSetAttribute(ar_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    ar_mega_kernel<<<n_sm, MG_THREADS, smem, st>>>(P);
    return check_launch("ar_mega");
}

}  // namespace rqb
