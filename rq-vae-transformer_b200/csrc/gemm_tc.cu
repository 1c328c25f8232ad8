// P3 "fast" tier workhorse -- weight-streaming skinny GEMM on tcgen05 tensor cores.
//
//   D[n, b] = sum_k W[n, k] * X[b, k]            W: [N_out, K] (nn.Linear layout), X: [B, K], both fp16 or both bf16
//                                                (GemmTcParams.fmt; the reference's amp class is fp16), D fp32 in TMEM
//
// Replaces every nn.Linear of the cached AR step (reference: attentions.py:69-71,99,117-122; transformers.py:94) at
// M = batch rows.  At B <= 256 these GEMMs are HBM-bound on the *weights* (SURVEY.md finding 5), so the kernel is laid
// out as a weight streamer with the operands SWAPPED: the weight tile is the UMMA "A" operand (M = 128 output features
// per CTA, K-major -- exactly the [out,in] row-major layout checkpoints already have, no transpose), the activations are
// the "B" operand (N = batch padded to 16).  One elected thread issues tcgen05.mma (128 x BN x 16, kind::f16 -> fp32 TMEM);
// weights and activations arrive through TMA (SWIZZLE_128B, 64-element K slabs) into a STAGES-deep mbarrier ring.
//
// Programmatic dependent launch: weight tiles do not depend on the previous kernel, so the producer warp fills the ring
// with weights BEFORE griddepcontrol.wait; only the activation loads, the epilogue's residual reads and all stores wait
// for the upstream kernel.  Back-to-back kernels of the per-token chain thereby keep HBM busy across kernel boundaries.
//
// Split-K (blockIdx.x = tile * splits + split) spreads the N_out/128 tiles of the narrow GEMMs (proj, fc2: 12 tiles at
// E = 1536) over all 148 SMs; partial tiles go to an fp32 workspace [split][B][N_out] and are summed in a FIXED order by
// the consumer kernel (ln_reduce / attn_fast / act_reduce) -> deterministic, no atomics.  (A form that reduced inside this kernel --
// partial tiles to L2, one arrival counter per tile, each split CTA reducing its slice of the rows -- was built and measured in
// round 2: an in-kernel barrier costs what a kernel boundary costs and GEMM-after-GEMM loses the weight prefetch; 249 vs 194 ms.)
//
// More activation rows than one UMMA N (256) -- batched prefill, teacher-forced forward -- run as gridDim.y row chunks of BN.
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

// stride > 1: the box samples every `stride`-th pixel along W and H (a strided conv's input pixels for one filter tap): boxDim is
// the traversed extent, the unit loads boxDim / elementStride elements
int make_tmap_4d_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_c,
                      uint32_t box_w, uint32_t box_h, uint32_t box_b, uint32_t stride) {
    PFN_cuTensorMapEncodeTiled_v12000 enc;
    RQB_TRY(get_encode_fn(&enc));
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {box_c, box_w * stride, box_h * stride, box_b};
    cuuint32_t estr[4] = {1, stride, stride, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RQB200_ECUDA, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
    return 0;
}

constexpr int GT_THREADS = 192;
constexpr int GT_A_BYTES = 128 * 64 * 2;     // 128 output features x 64 k, 16-bit

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN, int MODE, bool FULL>
__device__ __forceinline__ void gt_epilogue_cols(const GemmTcParams& p, uint32_t taddr, bool nvalid, float bias, const float* res,
                                                 int64_t res_ld, int rdiv, float* out_f, h16* out_h, int64_t ld, int m0, int nb) {
    // issue the TMEM loads of up to 64 columns back to back, wait once
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 64) {
        uint32_t r4[4][16];
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (cb + c * 16 < BN) tc::tmem_ld16(taddr + (uint32_t)(cb + c * 16), r4[c]);
        tc::tmem_ld_wait();
        if (nvalid) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (cb + c * 16 < BN) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int col = cb + c * 16 + i;                    // activation row m0 + col
                        if (FULL || col < nb) {
                            float v = __uint_as_float(r4[c][i]);
                            if (MODE != GT_PARTIAL) v += bias;
                            if (MODE == GT_F32 && res != nullptr)
                                v += res[(rdiv ? (int64_t)((m0 + col) / rdiv) : (int64_t)(m0 + col)) * res_ld];
                            if (MODE == GT_PARTIAL || MODE == GT_F32) out_f[(int64_t)col * ld] = v;
                            else if (MODE == GT_H16) out_h[(int64_t)col * ld] = pack_h16(v, p.fmt);
                            else out_h[(int64_t)col * ld] = pack_h16(gelu_erf_f(v), p.fmt);
                        }
                    }
                }
            }
        }
    }
}

// Epilogue of one thread: accumulator columns [0, BN) of TMEM lane `taddr` = output feature n for the activation rows
// m0 .. m0 + nb - 1.  One warp per scheduler here, so the instruction stream is kept short and branch-free: the mode is a
// template parameter, row addresses are base + constant * stride, and the row bound is checked only for ragged chunks.
template <int BN, int MODE>
__device__ __forceinline__ void gt_epilogue(const GemmTcParams& p, uint32_t taddr, int n, int split, int m0, int nb) {
    const bool nvalid = n < p.N_out;
    const float bias = (MODE != GT_PARTIAL && p.bias != nullptr && nvalid) ? p.bias[n] * p.bias_scale : 0.f;
    const float* res = nullptr;
    int64_t res_ld = 0;
    if (MODE == GT_F32 && p.residual != nullptr && nvalid) {
        res = p.residual + (p.res_row_ptr ? (int64_t)(*p.res_row_ptr) * p.res_row_stride : 0) + n;
        res_ld = p.ld_res;
    }
    const int rdiv = p.res_div > 1 ? p.res_div : 0;                 // 0: residual row == activation row
    float* out_f = nullptr;
    h16* out_h = nullptr;
    int64_t ld = 0;
    if (MODE == GT_PARTIAL) {
        out_f = p.partial + ((int64_t)split * p.B + m0) * p.N_out + n;
        ld = p.N_out;
    } else if (MODE == GT_F32) {
        out_f = reinterpret_cast<float*>(p.out) + (int64_t)m0 * p.ld_out + n;
        ld = p.ld_out;
    } else {
        out_h = reinterpret_cast<h16*>(p.out) + (int64_t)m0 * p.ld_out + n;
        ld = p.ld_out;
    }
    if (nb == BN) gt_epilogue_cols<BN, MODE, true>(p, taddr, nvalid, bias, res, res_ld, rdiv, out_f, out_h, ld, m0, nb);
    else gt_epilogue_cols<BN, MODE, false>(p, taddr, nvalid, bias, res, res_ld, rdiv, out_f, out_h, ld, m0, nb);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, GemmTcParams p) {
    constexpr int B_BYTES = BN * 64 * 2;
    constexpr int STAGE_BYTES = GT_A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* xfull = full + STAGES;              // activations land on their own barrier (weights: full[])
    uint64_t* empty = xfull + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int m0 = blockIdx.y * BN;                  // first activation row of this CTA's chunk
    const int nkb_total = p.K / 64;
    const int kb0 = (int)((int64_t)nkb_total * split / p.splits), kb1 = (int)((int64_t)nkb_total * (split + 1) / p.splits);
    const int nkb = kb1 - kb0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;

    tc::pdl_launch_dependents();             // let the next kernel of the chain start its own weight prefetch
    if (warp == 0 && lane == 0) {
        if (tr && !p.trace_w) p.trace[0] = tc::gtimer();
        tc::prefetch_tmap(&tmW);
        tc::prefetch_tmap(&tmX);
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&xfull[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);     // (also relinquishes the allocation permit: co-resident CTAs do not wait)
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer.  Weights first (independent of the upstream kernel), then wait, then activations.
            const int pre = nkb < STAGES ? nkb : STAGES;
            // streamed once (M <= 256: evict first) or shared by every row chunk of a large-M launch (keep in L2)
            const uint64_t w_hint = gridDim.y > 1 ? tc::L2_EVICT_LAST : tc::L2_EVICT_FIRST;
            for (int i = 0; i < pre; i++) {
                tc::mbar_expect_tx(&full[i], GT_A_BYTES);
                tc::tma_load_2d(smem + i * STAGE_BYTES, &tmW, &full[i], (kb0 + i) * 64, tile * 128, w_hint);
            }
            if (p.l2pf)
                for (int i = pre; i < nkb; i++) tc::tma_prefetch_2d(&tmW, (kb0 + i) * 64, tile * 128);
            tc::pdl_wait();
            if (tr) p.trace[1] = tc::gtimer();
            for (int i = 0; i < pre; i++) {
                tc::mbar_expect_tx(&xfull[i], B_BYTES);
                tc::tma_load_2d(smem + i * STAGE_BYTES + GT_A_BYTES, &tmX, &xfull[i], (kb0 + i) * 64, m0, tc::L2_EVICT_LAST);
            }
            for (int i = pre; i < nkb; i++) {
                const int s = i % STAGES;
                tc::mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                tc::mbar_expect_tx(&full[s], GT_A_BYTES);
                tc::tma_load_2d(smem + s * STAGE_BYTES, &tmW, &full[s], (kb0 + i) * 64, tile * 128, w_hint);
                tc::mbar_expect_tx(&xfull[s], B_BYTES);
                tc::tma_load_2d(smem + s * STAGE_BYTES + GT_A_BYTES, &tmX, &xfull[s], (kb0 + i) * 64, m0, tc::L2_EVICT_LAST);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        // one polling lane: 31 idle lanes spinning on the same mbarrier only add shared-memory traffic next to the TMA writes
        const uint32_t idesc = tc::umma_idesc(128, BN, p.fmt);
        if (lane == 0) {
            if (tr && p.trace_w) {
                // diagnostic: when did the weight tiles requested ahead of the dependency land?  (replaces the entry stamp)
                for (int i = 0; i < (nkb < STAGES ? nkb : STAGES); i++) tc::mbar_wait(&full[i], 0);
                p.trace[0] = tc::gtimer();
            }
            for (int i = 0; i < nkb; i++) {
                const int s = i % STAGES;
                tc::mbar_wait(&full[s], (i / STAGES) & 1);
                tc::mbar_wait(&xfull[s], (i / STAGES) & 1);
                tc::tc_fence_after();
                const uint32_t a = tc::smem_u32(smem + s * STAGE_BYTES), b = a + GT_A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    tc::umma_f16(tmem_base, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + j * 32), idesc,
                                 (i > 0 || j > 0) ? 1u : 0u);
                tc::umma_commit(&empty[s]);                    // frees the ring slot once these MMAs have read it
                if (i == nkb - 1) tc::umma_commit(tmem_full);  // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4, thread <-> one output feature n, all batch columns
        tc::pdl_wait();
        const int q = warp & 3;
        const int n = tile * 128 + q * 32 + lane;
        if (lane == 0) tc::mbar_wait(tmem_full, 0);              // (one polling lane per warp)
        __syncwarp();
        tc::tc_fence_after();
        if (tr && warp == 2 && lane == 0) p.trace[2] = tc::gtimer();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int nb = (p.B - m0) < BN ? (p.B - m0) : BN;          // valid activation rows of this chunk
        switch (p.mode) {
            case GT_PARTIAL: gt_epilogue<BN, GT_PARTIAL>(p, taddr, n, split, m0, nb); break;
            case GT_F32: gt_epilogue<BN, GT_F32>(p, taddr, n, split, m0, nb); break;
            case GT_H16: gt_epilogue<BN, GT_H16>(p, taddr, n, split, m0, nb); break;
            default: gt_epilogue<BN, GT_H16_GELU>(p, taddr, n, split, m0, nb); break;
        }
        if (tr && warp == 2 && lane == 0) p.trace[3] = tc::gtimer();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES>
static int launch_gemm_tc_t(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p, bool pdl, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (GT_A_BYTES + BN * 128) + 1024 + 512;
    RQB_ENSURE_SMEM(smem, gemm_tc_kernel<BN, STAGES>);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ceil_div(p.N_out, 128) * p.splits), (unsigned)ceil_div(p.B, BN));
    cfg.blockDim = dim3(GT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RQB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES>, tmW, tmX, p));
    g_launches++;
    return 0;
}

// ring depth: `deep` = the deepest ring that fits (one CTA per SM, the whole K slice of a split prefetched ahead of the upstream
// kernel); otherwise half of it, so that two CTAs of consecutive launches share an SM.  Same k order -> same bits either way.
int launch_gemm_tc(const CUtensorMap& tmW, const CUtensorMap& tmX, const GemmTcParams& p, bool pdl, cudaStream_t st) {
    if (p.K % 64 != 0 || p.N_out % 128 != 0) return fail(RQB200_EINVAL, "gemm_tc: need K % 64 == 0 and N_out % 128 == 0");
    if (p.B < 1) return fail(RQB200_EINVAL, "gemm_tc: no activation rows");
    if (p.splits < 1 || p.splits > p.K / 64) return fail(RQB200_EINVAL, "gemm_tc: bad split count");
    if (p.fmt != 0 && p.fmt != 1) return fail(RQB200_EINVAL, "gemm_tc: fmt must be 0 (fp16) or 1 (bf16)");
    const int bn = gemm_tc_bn(p.B);
    if (p.B > 256 && p.mode == GT_PARTIAL) return fail(RQB200_EINVAL, "gemm_tc: split-K takes at most 256 activation rows");
    if (p.mode != GT_PARTIAL && p.splits != 1) return fail(RQB200_EINVAL, "gemm_tc: direct epilogues need splits == 1");
    const bool deep = p.deep != 0;
    switch (bn) {
        case 16: return deep ? launch_gemm_tc_t<16, 8>(tmW, tmX, p, pdl, st) : launch_gemm_tc_t<16, 4>(tmW, tmX, p, pdl, st);
        case 32: return deep ? launch_gemm_tc_t<32, 8>(tmW, tmX, p, pdl, st) : launch_gemm_tc_t<32, 4>(tmW, tmX, p, pdl, st);
        case 64: return deep ? launch_gemm_tc_t<64, 8>(tmW, tmX, p, pdl, st) : launch_gemm_tc_t<64, 4>(tmW, tmX, p, pdl, st);
        case 128: return deep ? launch_gemm_tc_t<128, 6>(tmW, tmX, p, pdl, st) : launch_gemm_tc_t<128, 3>(tmW, tmX, p, pdl, st);
        default: return deep ? launch_gemm_tc_t<256, 4>(tmW, tmX, p, pdl, st) : launch_gemm_tc_t<256, 2>(tmW, tmX, p, pdl, st);
    }
}

// weight tensor map: row-major [N_out, K] 16-bit, box = 64 k x 128 rows
int make_tmap_weight(CUtensorMap* out, const void* W, int N_out, int K) {
    return make_tmap_2d(out, W, 1, (uint64_t)K, (uint64_t)N_out, (uint64_t)K * 2, 64, 128);
}

}  // namespace rqb

// ---- diagnostic entry points (tests/test_gpu_tc.py, bench.py's roofline leg): one GEMM through the tcgen05 kernel

extern "C" int rqb200_dbg_gemm_tc(const void* W16, const void* X16, const float* bias, const float* residual, void* out,
                                  int out_is_16, int gelu, float* partial, int N_out, int K, int B, int splits, int fmt,
                                  void* stream) {
    using namespace rqb;
    CUtensorMap tw, tx;
    const int bn = gemm_tc_bn(B);
    RQB_TRY(make_tmap_weight(&tw, W16, N_out, K));
    RQB_TRY(make_tmap_2d(&tx, X16, 1, (uint64_t)K, (uint64_t)B, (uint64_t)K * 2, 64, (uint32_t)bn));
    GemmTcParams p = {};
    p.N_out = N_out; p.K = K; p.B = B; p.splits = splits; p.fmt = fmt; p.deep = 1;
    p.bias = bias; p.bias_scale = 1.f; p.residual = residual; p.ld_res = N_out; p.out = out; p.ld_out = N_out; p.partial = partial;
    p.mode = (splits > 1 || partial != nullptr) ? GT_PARTIAL : (out_is_16 ? (gelu ? GT_H16_GELU : GT_H16) : GT_F32);
    if (p.mode == GT_PARTIAL && partial == nullptr) return fail(RQB200_EINVAL, "dbg_gemm_tc: splits > 1 needs a partial buffer");
    return launch_gemm_tc(tw, tx, p, false, (cudaStream_t)stream);
}
