// P3 "fast" tier workhorse -- weight-streaming skinny GEMM on tcgen05 tensor cores.
//
//   D[n, b] = sum_k W[n, k] * X[b, k]            W: [N_out, K] bf16 (nn.Linear layout), X: [B, K] bf16, D fp32 in TMEM
//
// Replaces every nn.Linear of the cached AR step (reference: attentions.py:69-71,99,117-122; transformers.py:94) at
// M = batch rows.  At B <= 256 these GEMMs are HBM-bound on the *weights* (SURVEY.md finding 5), so the kernel is laid
// out as a weight streamer with the operands SWAPPED: the weight tile is the UMMA "A" operand (M = 128 output features
// per CTA, K-major -- exactly the [out,in] row-major layout checkpoints already have, no transpose), the activations are
// the "B" operand (N = batch padded to 16).  One elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 TMEM);
// weights and activations arrive through TMA (SWIZZLE_128B, 64-element K slabs) into a STAGES-deep mbarrier ring.
//
// Programmatic dependent launch: weight tiles do not depend on the previous kernel, so the producer warp fills the ring
// with weights BEFORE griddepcontrol.wait; only the activation loads, the epilogue's residual reads and all stores wait
// for the upstream kernel.  Back-to-back kernels of the per-token chain thereby keep HBM busy across kernel boundaries.
//
// Split-K (blockIdx.x = tile * splits + split) spreads the N_out/128 tiles of the narrow GEMMs (proj, fc2: 12 tiles at
// E = 1536) over all 148 SMs; partial tiles go to an fp32 workspace [split][B][N_out] and are summed in a FIXED order by
// the consumer kernel (ln_reduce) -> deterministic, no atomics.
#include "kernels.h"
#include "tc_common.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>

namespace rqb {

static int get_encode_fn(PFN_cuTensorMapEncodeTiled_v12000* fn) {
    static PFN_cuTensorMapEncodeTiled_v12000 cached = nullptr;
    if (!cached) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return fail(RQB200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
        cached = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    *fn = cached;
    return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes_log2, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
    PFN_cuTensorMapEncodeTiled_v12000 enc;
    RQB_TRY(get_encode_fn(&enc));
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = elem_bytes_log2 == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RQB200_ECUDA, "cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
    return 0;
}

int make_tmap_4d_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_c,
                      uint32_t box_w, uint32_t box_h, uint32_t box_b) {
    PFN_cuTensorMapEncodeTiled_v12000 enc;
    RQB_TRY(get_encode_fn(&enc));
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {box_c, box_w, box_h, box_b};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RQB200_ECUDA, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
    return 0;
}

constexpr int GT_THREADS = 192;
constexpr int GT_A_BYTES = 128 * 64 * 2;     // 128 output features x 64 k, bf16

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// CL == true: the `splits` CTAs of one output tile form a thread-block cluster; instead of writing fp32 partials to global
// memory they stage them in their own shared memory, and after a cluster barrier CTA r sums rows [r*rp, (r+1)*rp) of the
// tile over all peers through distributed shared memory (fixed peer order -> deterministic) and applies the epilogue.
// Removes the partial round trip through L2 and the reduction work from the consumer kernel.
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_map(uint32_t local_addr, uint32_t rank) {
    uint32_t ra;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    return ra;
}
// not volatile / no memory clobber on purpose: the S loads of one element are independent and must be allowed to overlap;
// ordering against the producers' stores is provided by the cluster barriers around the reduction
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr));
    return v;
}

// ---- GT_GR tail, executed by the 128 epilogue threads (warps 2..5; `ew` = 0..3) after they have stored their partial tile.
__device__ __forceinline__ unsigned gr_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gr_bar_epilogue() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float gr_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __noinline__ void gr_reduce_tail(const GemmTcParams& p, int tile, int split, int ew, int lane) {
    const int S = p.splits;
    // 1. publish this CTA's partial tile, wait for the S - 1 peers of the tile.  One gpu-scope release by one thread AFTER the
    //    CTA barrier covers every epilogue thread's stores (release is cumulative over what the barrier ordered before it) --
    //    the pattern of a cooperative-groups grid sync; no per-thread fence.
    gr_bar_epilogue();
    if (ew == 0 && lane == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.gr_counter + tile), "r"(1u) : "memory");
        const long long t0 = clock64();
        while (gr_ld_acquire(p.gr_counter + tile) < (unsigned)S) {
            if (clock64() - t0 > (1ll << 32)) __trap();      // > 2 s: the grid is not co-resident (see GemmTcParams) -- fail, do not hang
        }
    }
    gr_bar_epilogue();
    // 2. reduce rows [r0, r1) of the tile: warp <-> row (stride 4), lane <-> 4 consecutive output features
    const int rp = (p.B + S - 1) / S;
    const int r0 = split * rp, r1 = (r0 + rp) < p.B ? (r0 + rp) : p.B;
    const float* sc = p.gr_scratch + (int64_t)tile * S * p.B * 128;
    const int n0 = tile * 128 + 4 * lane;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
        bias4 = *reinterpret_cast<const float4*>(p.bias + n0);
        bias4.x *= p.bias_scale; bias4.y *= p.bias_scale; bias4.z *= p.bias_scale; bias4.w *= p.bias_scale;
    }
    float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.gr_stats_in) c4 = *reinterpret_cast<const float4*>(p.gr_fold_c + n0);
    const float* res = nullptr;
    if (p.gr_kind == 0 && p.residual != nullptr) res = p.residual + (p.res_row_ptr ? (int64_t)(*p.res_row_ptr) * p.res_row_stride : 0);
    const int nst = p.K / 128;
    for (int r = r0 + ew; r < r1; r += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < S; s0 += 8) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (s0 + i < S) v[i] = __ldcg(reinterpret_cast<const float4*>(sc + ((int64_t)(s0 + i) * p.B + r) * 128) + lane);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (s0 + i < S) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }   // fixed order
        }
        if (p.gr_stats_in) {
            // LayerNorm statistics of input row r from its per-tile (sum, M2) pairs (Chan's parallel combination)
            float s1 = 0.f;
            for (int t = lane; t < nst; t += 32) s1 += __ldcg(p.gr_stats_in + (int64_t)r * nst + t).x;
            const float mean = gr_warp_sum(s1) / (float)p.K;
            float m2 = 0.f;
            for (int t = lane; t < nst; t += 32) {
                const float2 st = __ldcg(p.gr_stats_in + (int64_t)r * nst + t);
                const float d = st.x * (1.0f / 128.0f) - mean;
                m2 += st.y + 128.0f * d * d;
            }
            const float rstd = rsqrtf(gr_warp_sum(m2) / (float)p.K + 1e-5f);
            acc.x = rstd * (acc.x - mean * c4.x); acc.y = rstd * (acc.y - mean * c4.y);
            acc.z = rstd * (acc.z - mean * c4.z); acc.w = rstd * (acc.w - mean * c4.w);
        }
        acc.x += bias4.x; acc.y += bias4.y; acc.z += bias4.z; acc.w += bias4.w;
        if (p.gr_kind == 1) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(gelu_erf_f(acc.x), gelu_erf_f(acc.y));
            __nv_bfloat162 h1 = __floats2bfloat162_rn(gelu_erf_f(acc.z), gelu_erf_f(acc.w));
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&h0);
            pk.y = *reinterpret_cast<unsigned*>(&h1);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)r * p.ld_out + n0) = pk;
        } else {
            if (res) {
                const float4 rr = __ldcg(reinterpret_cast<const float4*>(res + (int64_t)r * p.ld_res + n0));
                acc.x += rr.x; acc.y += rr.y; acc.z += rr.z; acc.w += rr.w;
            }
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (int64_t)r * p.ld_out + n0) = acc;
            if (p.gr_out_bf16) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
                uint2 pk;
                pk.x = *reinterpret_cast<unsigned*>(&h0);
                pk.y = *reinterpret_cast<unsigned*>(&h1);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.gr_out_bf16) + (int64_t)r * p.ld_out + n0) = pk;
            }
            if (p.gr_stats_out) {
                const float s1 = gr_warp_sum((acc.x + acc.y) + (acc.z + acc.w));
                const float tm = s1 * (1.0f / 128.0f);
                const float d0 = acc.x - tm, d1 = acc.y - tm, d2 = acc.z - tm, d3 = acc.w - tm;
                const float m2 = gr_warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
                if (lane == 0) p.gr_stats_out[(int64_t)r * (p.N_out / 128) + tile] = make_float2(s1, m2);
            }
        }
    }
}

template <int BN, int STAGES, bool CL, bool GR = false>
__global__ void __launch_bounds__(GT_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, GemmTcParams p) {
    constexpr int B_BYTES = BN * 64 * 2;
    constexpr int STAGE_BYTES = GT_A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int nkb_total = p.K / 64;
    const int kb0 = (int)((int64_t)nkb_total * split / p.splits), kb1 = (int)((int64_t)nkb_total * (split + 1) / p.splits);
    const int nkb = kb1 - kb0;

    tc::pdl_launch_dependents();             // let the next kernel of the chain start its own weight prefetch
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmW);
        tc::prefetch_tmap(&tmX);
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        if (p.relinq) tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer.  Weights first (independent of the upstream kernel), then wait, then activations.
            const int pre = nkb < STAGES ? nkb : STAGES;
            for (int i = 0; i < pre; i++) {
                tc::mbar_expect_tx(&full[i], STAGE_BYTES);
                tc::tma_load_2d(smem + i * STAGE_BYTES, &tmW, &full[i], p.w_tiled ? 0 : (kb0 + i) * 64,
                                p.w_tiled ? (tile * nkb_total + kb0 + i) * 128 : tile * 128, tc::L2_EVICT_FIRST);
            }
            if (p.l2pf)
                for (int i = pre; i < nkb; i++)
                    tc::tma_prefetch_2d(&tmW, p.w_tiled ? 0 : (kb0 + i) * 64,
                                        p.w_tiled ? (tile * nkb_total + kb0 + i) * 128 : tile * 128);
            tc::pdl_wait();
            for (int i = 0; i < pre; i++)
                tc::tma_load_2d(smem + i * STAGE_BYTES + GT_A_BYTES, &tmX, &full[i], (kb0 + i) * 64, 0, tc::L2_EVICT_LAST);
            for (int i = pre; i < nkb; i++) {
                const int s = i % STAGES;
                tc::mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                tc::mbar_expect_tx(&full[s], STAGE_BYTES);
                tc::tma_load_2d(smem + s * STAGE_BYTES, &tmW, &full[s], p.w_tiled ? 0 : (kb0 + i) * 64,
                                p.w_tiled ? (tile * nkb_total + kb0 + i) * 128 : tile * 128, tc::L2_EVICT_FIRST);
                tc::tma_load_2d(smem + s * STAGE_BYTES + GT_A_BYTES, &tmX, &full[s], (kb0 + i) * 64, 0, tc::L2_EVICT_LAST);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        constexpr uint32_t idesc = tc::umma_idesc(128, BN, 1 /*bf16*/);
        for (int i = 0; i < nkb; i++) {
            const int s = i % STAGES;
            tc::mbar_wait(&full[s], (i / STAGES) & 1);
            tc::tc_fence_after();
            if (lane == 0) {
                const uint32_t a = tc::smem_u32(smem + s * STAGE_BYTES), b = a + GT_A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    tc::umma_f16(tmem_base, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + j * 32), idesc,
                                 (i > 0 || j > 0) ? 1u : 0u);
                tc::umma_commit(&empty[s]);                    // frees the ring slot once these MMAs have read it
                if (i == nkb - 1) tc::umma_commit(tmem_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4, thread <-> one output feature n, all batch columns
        tc::pdl_wait();
        const int q = warp & 3;
        const int n = tile * 128 + q * 32 + lane;
        tc::mbar_wait(tmem_full, 0);
        tc::tc_fence_after();
        const bool nvalid = n < p.N_out;
        if (CL) {
            float* stg = reinterpret_cast<float*>(smem);                    // [128][BN + 1] fp32, over the drained ring
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) stg[(q * 32 + lane) * (BN + 1) + c0 + i] = __uint_as_float(r[i]);
            }
        } else {
        const float bias = (p.bias != nullptr && nvalid && p.mode != GT_PARTIAL && !GR) ? p.bias[n] * p.bias_scale : 0.f;
        const float* res = nullptr;
        if (p.mode == GT_F32 && p.residual != nullptr)
            res = p.residual + (p.res_row_ptr ? (int64_t)(*p.res_row_ptr) * p.res_row_stride : 0);
        // issue the TMEM loads of up to 64 columns back to back, wait once
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 64) {
            uint32_t r4[4][16];
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (cb + c * 16 < BN) tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb + c * 16), r4[c]);
            tc::tmem_ld_wait();
            if (!nvalid) continue;
#pragma unroll
            for (int c = 0; c < 4; c++) {
            const int c0 = cb + c * 16;
            if (c0 >= BN) break;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int b = c0 + i;
                if (b >= p.B) break;
                float v = __uint_as_float(r4[c][i]) + bias;
                if constexpr (GR) {      // group-reduce instantiation: the partial tile goes to the tile-major L2 scratch
                    __stcg(p.gr_scratch + (((int64_t)tile * p.splits + split) * p.B + b) * 128 + (q * 32 + lane), v);
                    continue;
                }
                switch (p.mode) {
                    case GT_F32:
                        if (res) v += res[(int64_t)b * p.ld_res + n];
                        reinterpret_cast<float*>(p.out)[(int64_t)b * p.ld_out + n] = v;
                        break;
                    case GT_BF16:
                        reinterpret_cast<__nv_bfloat16*>(p.out)[(int64_t)b * p.ld_out + n] = __float2bfloat16(v);
                        break;
                    case GT_BF16_GELU:
                        reinterpret_cast<__nv_bfloat16*>(p.out)[(int64_t)b * p.ld_out + n] = __float2bfloat16(gelu_erf_f(v));
                        break;
                    case GT_PARTIAL:
                        p.partial[((int64_t)split * p.B + b) * p.N_out + n] = v;
                        break;
                    default:
                        break;
                }
            }
            }
        }
        if constexpr (GR) gr_reduce_tail(p, tile, split, warp - 2, lane);
        }
    }
    if (CL) {
        __syncwarp();
        cluster_sync_all();                                                 // every CTA of the tile has staged its partial
        tc::pdl_wait();
        const int S = p.splits, rp = (128 + S - 1) / S;
        const int r0 = split * rp, r1 = (r0 + rp) < 128 ? (r0 + rp) : 128;
        const uint32_t stg_u32 = tc::smem_u32(smem);
        const float* res = nullptr;
        if (p.mode == GT_F32 && p.residual != nullptr)
            res = p.residual + (p.res_row_ptr ? (int64_t)(*p.res_row_ptr) * p.res_row_stride : 0);
        const int total = (r1 - r0) * p.B;
        uint32_t peer[8];
#pragma unroll
        for (int pr = 0; pr < 8; pr++) peer[pr] = dsmem_map(stg_u32, (uint32_t)(pr < S ? pr : 0));
        for (int idx = threadIdx.x; idx < total; idx += GT_THREADS) {
            // consecutive threads -> consecutive output features (coalesced stores); b is the slow index
            const int row = r0 + idx % (r1 - r0), b = idx / (r1 - r0);
            const int n = tile * 128 + row;
            const uint32_t off = (uint32_t)(row * (BN + 1) + b) * 4u;
            float part[8];
#pragma unroll
            for (int pr = 0; pr < 8; pr++) part[pr] = pr < S ? ld_dsmem_f32(peer[pr] + off) : 0.f;
            float v = 0.f;
#pragma unroll
            for (int pr = 0; pr < 8; pr++) v += part[pr];                       // fixed order: deterministic
            if (p.bias) v += p.bias[n] * p.bias_scale;
            if (p.mode == GT_F32) {
                if (res) v += res[(int64_t)b * p.ld_res + n];
                reinterpret_cast<float*>(p.out)[(int64_t)b * p.ld_out + n] = v;
            } else if (p.mode == GT_BF16_GELU) {
                reinterpret_cast<__nv_bfloat16*>(p.out)[(int64_t)b * p.ld_out + n] = __float2bfloat16(gelu_erf_f(v));
            } else {
                reinterpret_cast<__nv_bfloat16*>(p.out)[(int64_t)b * p.ld_out + n] = __float2bfloat16(v);
            }
        }
        __syncwarp();
        cluster_sync_all();                                                 // peers may still be reading this CTA's staging area
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, bool CL, bool GR = false>
static int launch_gemm_tc_t(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p, bool pdl, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (GT_A_BYTES + BN * 128) + 1024 + 256;
    static_assert(!CL || (size_t)STAGES * (GT_A_BYTES + BN * 128) >= (size_t)128 * (BN + 1) * 4, "staging area must fit in the ring");
    RQB_ENSURE_SMEM(smem, gemm_tc_kernel<BN, STAGES, CL, GR>);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ceil_div(p.N_out, 128) * p.splits));
    cfg.blockDim = dim3(GT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (CL) {
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = (unsigned)p.splits;
        at[1].val.clusterDim.y = 1;
        at[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    RQB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, CL, GR>, tmW, tmX, p));
    g_launches++;
    return 0;
}

int launch_gemm_tc(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p_in, bool pdl, cudaStream_t st) {
    GemmTcParams p = p_in;
    if (const char* e = getenv("RQB200_GEMM_L2PF")) p.l2pf = atoi(e);
    if (const char* e = getenv("RQB200_GEMM_RELINQ")) p.relinq = atoi(e);
    if (p.K % 64 != 0 || p.N_out % 128 != 0) return fail(RQB200_EINVAL, "gemm_tc: need K % 64 == 0 and N_out % 128 == 0");
    if (p.B < 1 || p.B > 256) return fail(RQB200_EINVAL, "gemm_tc: batch rows must be in [1,256]");
    if (p.splits < 1 || p.splits > p.K / 64) return fail(RQB200_EINVAL, "gemm_tc: bad split count");
    const int bn = gemm_tc_bn(p.B);
    if (p.mode == GT_GR && (p.gr_scratch == nullptr || p.gr_counter == nullptr || p.ld_out % 4 != 0 || (p.gr_stats_in && p.K % 128 != 0)))
        return fail(RQB200_EINVAL, "gemm_tc: GT_GR needs scratch, counters, ld_out % 4 == 0 (and K % 128 == 0 with a folded LayerNorm)");
    if (p.mode == GT_GR) {
        switch (bn) {
            case 16: return launch_gemm_tc_t<16, 8, false, true>(tmW, tmX, p, pdl, st);
            case 32: return launch_gemm_tc_t<32, 8, false, true>(tmW, tmX, p, pdl, st);
            case 64: return launch_gemm_tc_t<64, 8, false, true>(tmW, tmX, p, pdl, st);
            case 128: return launch_gemm_tc_t<128, 6, false, true>(tmW, tmX, p, pdl, st);
            default: return launch_gemm_tc_t<256, 4, false, true>(tmW, tmX, p, pdl, st);
        }
    }
    if (p.splits > 1 && p.mode != GT_PARTIAL) {          // split-K with an in-kernel (cluster / DSMEM) reduction
        if (p.splits > 8) return fail(RQB200_EINVAL, "gemm_tc: cluster split-K supports at most 8 splits");
        switch (bn) {
            case 16: return launch_gemm_tc_t<16, 8, true>(tmW, tmX, p, pdl, st);
            case 32: return launch_gemm_tc_t<32, 8, true>(tmW, tmX, p, pdl, st);
            case 64: return launch_gemm_tc_t<64, 8, true>(tmW, tmX, p, pdl, st);
            case 128: return launch_gemm_tc_t<128, 6, true>(tmW, tmX, p, pdl, st);
            default: return launch_gemm_tc_t<256, 4, true>(tmW, tmX, p, pdl, st);
        }
    }
    // RQB200_GEMM_STAGES (experiment): a shallower ring lets 2-3 GEMM CTAs of consecutive launches share an SM, so the
    // next GEMM of the PDL chain prefetches its weights while the current one still runs.  Same k order -> same bits.
    int stages = 8;
    if (const char* e = getenv("RQB200_GEMM_STAGES")) stages = atoi(e);
    if (bn <= 64 && stages != 8) {
        if (bn == 64) {
            if (stages == 3) return launch_gemm_tc_t<64, 3, false>(tmW, tmX, p, pdl, st);
            if (stages == 4) return launch_gemm_tc_t<64, 4, false>(tmW, tmX, p, pdl, st);
            if (stages == 6) return launch_gemm_tc_t<64, 6, false>(tmW, tmX, p, pdl, st);
        } else if (bn == 32) {
            if (stages == 3 || stages == 4) return launch_gemm_tc_t<32, 4, false>(tmW, tmX, p, pdl, st);
        } else {
            if (stages == 3 || stages == 4) return launch_gemm_tc_t<16, 4, false>(tmW, tmX, p, pdl, st);
        }
    }
    switch (bn) {
        case 16: return launch_gemm_tc_t<16, 8, false>(tmW, tmX, p, pdl, st);
        case 32: return launch_gemm_tc_t<32, 8, false>(tmW, tmX, p, pdl, st);
        case 64: return launch_gemm_tc_t<64, 8, false>(tmW, tmX, p, pdl, st);
        case 128: return launch_gemm_tc_t<128, 6, false>(tmW, tmX, p, pdl, st);
        default: return launch_gemm_tc_t<256, 4, false>(tmW, tmX, p, pdl, st);
    }
}

// weight tensor map: row-major [N_out, K], or tile-major [N_out/128][K/64][128][64] viewed as a 64-wide 2-D tensor
int make_tmap_weight(CUtensorMap* out, const void* W, int N_out, int K, bool tiled) {
    if (tiled) return make_tmap_2d(out, W, 1, 64, (uint64_t)(N_out / 128) * (K / 64) * 128, 128, 64, 128);
    return make_tmap_2d(out, W, 1, (uint64_t)K, (uint64_t)N_out, (uint64_t)K * 2, 64, 128);
}

}  // namespace rqb

// ---- diagnostic entry point (tests/test_gpu_tc.py): one GEMM through the tcgen05 kernel

extern "C" int rqb200_dbg_gemm_tc(const void* W_bf16, const void* X_bf16, const float* bias, const float* residual, void* out,
                                  int out_is_bf16, int gelu, float* partial, int N_out, int K, int B, int splits,
                                  void* stream) {
    using namespace rqb;
    CUtensorMap tw, tx;
    const int bn = gemm_tc_bn(B);
    // splits < 0: W is tile-major ([N_out/128][K/64][128][64]) and |splits| is the split count
    const bool tiled = splits < 0;
    if (tiled) splits = -splits;
    RQB_TRY(make_tmap_weight(&tw, W_bf16, N_out, K, tiled));
    RQB_TRY(make_tmap_2d(&tx, X_bf16, 1, (uint64_t)K, (uint64_t)B, (uint64_t)K * 2, 64, (uint32_t)bn));
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = B; p.splits = splits;
    p.bias = bias; p.bias_scale = 1.f; p.residual = residual; p.ld_res = N_out; p.out = out; p.ld_out = N_out; p.partial = partial;
    p.w_tiled = tiled ? 1 : 0;
    p.mode = (splits > 1 && partial != nullptr) ? GT_PARTIAL : (out_is_bf16 ? (gelu ? GT_BF16_GELU : GT_BF16) : GT_F32);
    return launch_gemm_tc(tw, tx, p, false, (cudaStream_t)stream);
}

// one GT_GR launch (split-K with the in-kernel group reduction); scratch >= splits*B*N_out floats, counters >= N_out/128 (zeroed here)
extern "C" int rqb200_dbg_gemm_gr(const void* W_bf16, const void* X_bf16, const float* bias, const float* residual, void* out, int kind,
                                  float* scratch, unsigned* counters, void* out_bf16, float* stats_out, const float* stats_in,
                                  const float* fold_c, int N_out, int K, int B, int splits, void* stream) {
    using namespace rqb;
    CUtensorMap tw, tx;
    const int bn = gemm_tc_bn(B);
    RQB_TRY(make_tmap_weight(&tw, W_bf16, N_out, K, false));
    RQB_TRY(make_tmap_2d(&tx, X_bf16, 1, (uint64_t)K, (uint64_t)B, (uint64_t)K * 2, 64, (uint32_t)bn));
    int dev = 0, n_sm = 0;
    RQB_CUDA(cudaGetDevice(&dev));
    RQB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if ((N_out / 128) * splits > n_sm) return fail(RQB200_EINVAL, "dbg_gemm_gr: the grid would not be co-resident");
    RQB_CUDA(cudaMemsetAsync(counters, 0, (size_t)(N_out / 128) * sizeof(unsigned), (cudaStream_t)stream));
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = B; p.splits = splits; p.mode = GT_GR;
    p.bias = bias; p.bias_scale = 1.f; p.residual = kind == 0 ? residual : nullptr; p.ld_res = N_out; p.out = out; p.ld_out = N_out;
    p.gr_scratch = scratch; p.gr_counter = counters; p.gr_kind = kind; p.gr_out_bf16 = out_bf16;
    p.gr_stats_out = reinterpret_cast<float2*>(stats_out); p.gr_stats_in = reinterpret_cast<const float2*>(stats_in);
    p.gr_fold_c = fold_c;
    return launch_gemm_tc(tw, tx, p, false, (cudaStream_t)stream);
}
