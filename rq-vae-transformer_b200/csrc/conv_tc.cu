// P2 "fast" tier -- NHWC implicit-GEMM convolution on tcgen05 tensor cores (fp16 operands, fp32 accumulate in TMEM).
//
// Replaces cuDNN's conv behind ResnetBlock / AttnBlock / Upsample / conv_in / conv_out of the decoder (reference:
// rqvae/models/rqvae/layers.py:100-120,158-182,31-35; modules.py:171-202).  The reference's own GPU path runs these convs
// with TF32 allowed (10-bit mantissa); fp16 operands carry the same mantissa width.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[cout, tap, cin]
//
//   M = 128 output pixels per tile: a TW x TH x NB box of the NHWC activation tensor (NB > 1 images per tile when the
//       feature map is smaller than 128 pixels), fetched per filter tap by ONE 4-D TMA load whose coordinates are shifted
//       by the tap offset -- out-of-range rows/columns are zero-filled by the TMA unit, which IS the conv's zero padding;
//       no im2col buffer, no halo logic in the kernel.
//   N = BN output channels (16 | 128 | 256), K = taps x Cin walked in 64-channel slabs (one 128 B swizzled row per pixel).
//
// Persistent CTAs (grid = #SMs) loop over (pixel tile, cout tile) pairs; warp 0 = TMA producer, warp 1 = single-thread
// tcgen05.mma issuer, warps 2-5 = epilogue.  Two TMEM accumulators (2 x BN columns) are ping-ponged so the epilogue of
// tile i (TMEM -> registers -> +bias (+residual) -> fp32 NHWC / NCHW stores) overlaps the main loop of tile i+1.
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace rqb {

struct ConvTcParams {
    int B, H, W, Cin, Cout;        // H,W: OUTPUT extent; the input is H*stride x W*stride
    int ks;                        // 1 or 3
    int stride;                    // 1: "same" zero padding; 2: no left/top pad, one zero column/row on the right/bottom (F.pad
                                   // (0,1,0,1) + stride-2 conv, layers.py:50-57) -- both are the tensor map's out-of-bounds fill
    int TW, TH, NB;                // tile box, TW*TH*NB == 128
    int tiles_x, tiles_y, tiles_b, n_tiles_n;
    const float* bias;
    const float* residual;         // [B,H,W,Cout] f32 or null
    float* out;                    // NHWC f32, or NCHW f32 when out_nchw
    int out_nchw;
    // "rows GEMM" use of the same kernel (launch_rows_gemm_tc: a 1x1 conv over M token rows viewed as 16x8-pixel images):
    void* out16;                   // non-null: 16-bit NHWC output (fmt) instead of `out`, optionally through GELU
    int gelu, fmt;                 // fmt: 16-bit operand / output format, 0 fp16 (all convs), 1 bf16
    int64_t m_rows;                // > 0: only pixels (rows) below m_rows are stored
    // GroupNorm(32) statistics of the OUTPUT (bias / residual included), for the GroupNorm that consumes it next (layers.py:16-17,
    // 100-120): every epilogue warp (32 pixels of one image) writes (sum, sum of squares) per group as fp64 to
    // gn_part[((b * gn_chunks + chunk) * 32 + group) * 2] -- the partial layout gn_finalize_kernel reduces.  NULL: off.
    double* gn_part;
    int gn_chunks;                 // H * W / 32
};

constexpr int CT_THREADS = 192;
constexpr int CT_A_BYTES = 128 * 64 * 2;

// PASSES == 1: single fp16 product.  PASSES == 3: split-fp16 ("fp16x3") -- both operands are carried as hi + lo fp16 pairs and
// the accumulator receives A_hi W_hi + A_lo W_hi + A_hi W_lo (the dropped A_lo W_lo term is ~2^-22 relative): fp32-class
// products on the fp16 tensor pipe, which is what keeps 60 chained convs inside the 1e-3 pixel tolerance.
// GroupNorm partial statistics of one 16-channel chunk of a warp's 32 pixels: NV = 2 * (16 / CG) values (the sums, then the sums
// of squares, of the chunk's 16 / CG groups) are formed per thread and reduced over the 32 lanes with a transpose-reduce
// (log2(NV) halving steps, then plain butterfly steps): NV + 1 shuffles of depth 5 instead of 5 * NV.  On return `tot` is the total
// of value index `vidx` and `writer` marks the one lane per value that stores it.
template <int CG>
__device__ __forceinline__ void gn_chunk_stats(const float (&wv)[16], bool valid, int lane, float& tot, int& vidx, bool& writer) {
    constexpr int NG = 16 / CG, NV = 2 * NG;
    float vals[NV];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < CG; i++) {
            const float x = valid ? wv[g * CG + i] : 0.f;
            s1 += x;
            s2 = fmaf(x, x, s2);
        }
        vals[g] = s1;
        vals[NG + g] = s2;
    }
    int n = NV;
    vidx = 0;
    int keep_mask = 31;                         // lanes that end up holding identical totals differ only in these bits
#pragma unroll
    for (int S = 16; S >= 1; S >>= 1) {
        if (n > 1) {
            const int half = n >> 1;
            const bool up = (lane & S) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; i++)
                if (i < half) {
                    const float send = up ? vals[i] : vals[i + half];
                    const float keep = up ? vals[i + half] : vals[i];
                    vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
                }
            vidx += up ? half : 0;
            keep_mask &= ~S;
            n = half;
        } else {
            vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], S);
        }
    }
    tot = vals[0];
    writer = (lane & keep_mask) == 0;
}

template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, ConvTcParams p) {
    constexpr int B_BYTES = BN * 64 * 2;
    constexpr int NOPS = PASSES == 3 ? 2 : 1;
    constexpr int STAGE_BYTES = NOPS * (CT_A_BYTES + B_BYTES);
    constexpr int OFF_B = NOPS * CT_A_BYTES;                 // [A_hi | A_lo | B_hi | B_lo]
    constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;      // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cslabs = p.Cin / 64;
    const int nkb = p.ks * p.ks * cslabs;
    const int pad = (p.ks == 3 && p.stride == 1) ? 1 : 0;
    const int sx = p.stride;
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    const int total = m_tiles * p.n_tiles_n;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                   // running k-block counter across tiles
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
                const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.TW, y0 = ty * p.TH, b0 = tb * p.NB;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    tc::mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    const int tap = kb / cslabs, c0 = (kb % cslabs) * 64;
                    const int ky = tap / p.ks, kx = tap % p.ks;
                    tc::mbar_expect_tx(&full[s], STAGE_BYTES);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    tc::tma_load_4d(st, &tmA, &full[s], c0, x0 * sx + kx - pad, y0 * sx + ky - pad, b0, tc::L2_EVICT_NORMAL);
                    tc::tma_load_2d(st + OFF_B, &tmB, &full[s], tap * p.Cin + c0, nt * BN, tc::L2_EVICT_LAST);
                    if (PASSES == 3) {
                        tc::tma_load_4d(st + CT_A_BYTES, &tmAlo, &full[s], c0, x0 * sx + kx - pad, y0 * sx + ky - pad, b0, tc::L2_EVICT_NORMAL);
                        tc::tma_load_2d(st + OFF_B + B_BYTES, &tmBlo, &full[s], tap * p.Cin + c0, nt * BN, tc::L2_EVICT_LAST);
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::umma_idesc(128, BN, p.fmt);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tcount++) {
            const uint32_t as = tcount & 1;
            tc::mbar_wait(&tempty[as], ((tcount >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
            tc::tc_fence_after();
            for (int kb = 0; kb < nkb; kb++, it++) {
                const int s = it % STAGES;
                tc::mbar_wait(&full[s], (it / STAGES) & 1);
                tc::tc_fence_after();
                if (lane == 0) {
                    const uint32_t a = tc::smem_u32(smem + s * STAGE_BYTES), b = a + OFF_B;
                    if (PASSES == 3) {                              // small terms first, the dominant product last
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc::umma_f16(tmem_base + as * BN, tc::umma_desc_k128(a + CT_A_BYTES + j * 32), tc::umma_desc_k128(b + j * 32),
                                         idesc, (kb > 0 || j > 0) ? 1u : 0u);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc::umma_f16(tmem_base + as * BN, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + B_BYTES + j * 32),
                                         idesc, 1u);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        tc::umma_f16(tmem_base + as * BN, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + j * 32), idesc,
                                     (PASSES == 3 || kb > 0 || j > 0) ? 1u : 0u);
                    tc::umma_commit(&empty[s]);
                    if (kb == nkb - 1) tc::umma_commit(&tfull[as]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;                                   // row of the tile == TMEM lane
        const int rx = r % p.TW, ry = (r / p.TW) % p.TH, rb = r / (p.TW * p.TH);
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tcount++) {
            const uint32_t as = tcount & 1;
            const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
            const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
            const int x = tx * p.TW + rx, y = ty * p.TH + ry, b = tb * p.NB + rb;
            const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
            const bool valid = (x < p.W) && (y < p.H) && (b < p.B) && (p.m_rows == 0 || pix < p.m_rows);
            tc::mbar_wait(&tfull[as], (tcount >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c0, v);
                tc::tmem_ld_wait();
                const int n0 = nt * BN + c0;
                if (n0 >= p.Cout) continue;                                 // (warp-uniform)
                if (p.out_nchw) {
                    if (!valid) continue;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int n = n0 + i;
                        if (n < p.Cout)
                            p.out[(((int64_t)b * p.Cout + n) * p.H + y) * p.W + x] = __uint_as_float(v[i]) + p.bias[n];
                    }
                } else if (p.out16 != nullptr) {
                    if (!valid) continue;
                    float w[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + i);
                        w[i] = __uint_as_float(v[i]) + bb.x; w[i + 1] = __uint_as_float(v[i + 1]) + bb.y;
                        w[i + 2] = __uint_as_float(v[i + 2]) + bb.z; w[i + 3] = __uint_as_float(v[i + 3]) + bb.w;
                    }
                    if (p.gelu) {
#pragma unroll
                        for (int i = 0; i < 16; i++) w[i] = 0.5f * w[i] * (1.0f + erff(w[i] * 0.70710678118654752440f));
                    }
                    uint4 pk[2];
                    uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
                    for (int i = 0; i < 8; i++) pw[i] = pack_h16x2(w[2 * i], w[2 * i + 1], p.fmt);
                    uint4* o16 = reinterpret_cast<uint4*>(reinterpret_cast<h16*>(p.out16) + pix * p.Cout + n0);
                    o16[0] = pk[0];
                    o16[1] = pk[1];
                } else {
                    float* o = p.out + pix * p.Cout + n0;
                    const float* rs = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
                    float wv[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + i);
                        float4 w = make_float4(__uint_as_float(v[i]) + bb.x, __uint_as_float(v[i + 1]) + bb.y,
                                               __uint_as_float(v[i + 2]) + bb.z, __uint_as_float(v[i + 3]) + bb.w);
                        if (rs && valid) {
                            float4 rr = *reinterpret_cast<const float4*>(rs + i);
                            w.x += rr.x; w.y += rr.y; w.z += rr.z; w.w += rr.w;
                        }
                        if (valid) *reinterpret_cast<float4*>(o + i) = w;
                        wv[i] = w.x; wv[i + 1] = w.y; wv[i + 2] = w.z; wv[i + 3] = w.w;
                    }
                    if (p.gn_part != nullptr) {
                        // this warp's 32 pixels belong to one image (TW*TH >= 32, checked by the host); cg = 4 | 8 | 16 channels
                        const int cg = p.Cout >> 5;
                        const int wpi = (p.TW * p.TH) >> 5;                      // warps (32-pixel chunks) per image within a tile
                        const int chunk = (ty * p.tiles_x + tx) * wpi + (q % wpi);
                        const int bw = tb * p.NB + (q * 32) / (p.TW * p.TH);      // image of this warp
                        float tot;
                        int vidx;
                        bool writer;
                        if (cg == 4) gn_chunk_stats<4>(wv, valid, lane, tot, vidx, writer);
                        else if (cg == 8) gn_chunk_stats<8>(wv, valid, lane, tot, vidx, writer);
                        else gn_chunk_stats<16>(wv, valid, lane, tot, vidx, writer);
                        const int ngr = 16 / cg;
                        if (writer && bw < p.B)                                    // vidx < ngr: sum of group vidx; else sum of squares
                            p.gn_part[(((int64_t)bw * p.gn_chunks + chunk) * 32 + (n0 / cg + (vidx % ngr))) * 2 + (vidx / ngr)] = (double)tot;
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[as]);               // 4 epilogue warps -> accumulator free
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, int PASSES>
static int launch_conv_tc_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmAlo, const CUtensorMap& tmBlo,
                            const ConvTcParams& p, int n_sm, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (PASSES == 3 ? 2 : 1) * (CT_A_BYTES + BN * 128) + 1024 + 256;
    static_assert(smem <= 227 * 1024, "conv_tc: shared memory budget");
    RQB_ENSURE_SMEM(smem, conv_tc_kernel<BN, STAGES, PASSES>);
    const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles_n;
    const int grid = total < n_sm ? total : n_sm;
    conv_tc_kernel<BN, STAGES, PASSES><<<grid, CT_THREADS, smem, st>>>(tmA, tmB, tmAlo, tmBlo, p);
    return check_launch("conv_tc");
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

bool conv_tc_supported(int H, int W, int Cin, int Cout, int ks, int stride, int in_nchw) {
    if ((stride != 1 && !(stride == 2 && ks == 3)) || in_nchw || (ks != 1 && ks != 3)) return false;
    if (Cin % 64 != 0) return false;
    if (Cout != 3 && Cout % 128 != 0 && Cout != 64) return false;     // bias/residual float4 path needs Cout % 16 == 0
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    return pow2(H) && pow2(W);
}

// the epilogue can emit the output's GroupNorm(32) partial statistics when every epilogue warp's 32 pixels lie in one image and
// a 16-channel chunk holds whole groups
bool conv_tc_gn_fusable(int H, int W, int Cout) {
    const int TW = W < 16 ? W : 16, TH = (128 / TW) < H ? (128 / TW) : H;
    const int cg = Cout / 32;
    return Cout % 32 == 0 && (cg == 4 || cg == 8 || cg == 16) && TW * TH >= 32 && (H * W) % 32 == 0;
}

// X: NHWC fp16 [B,H*stride,W*stride,Cin]; Wt: [Cout, ks, ks, Cin] fp16; out fp32 [B,H,W,Cout].  X16lo/W16lo non-null -> split-fp16
// (3 products).  H, W are the OUTPUT extent.
int launch_conv_tc(const void* X16, const void* W16, const void* X16lo, const void* W16lo, const float* bias,
                   const float* residual, float* out, int B, int H, int W, int Cin, int Cout, int ks, int out_nchw,
                   cudaStream_t st, int stride, double* gn_part) {
    ConvTcParams p = {};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ks = ks; p.stride = stride;
    p.TW = W < 16 ? W : 16;
    p.TH = (128 / p.TW) < H ? (128 / p.TW) : H;
    p.NB = 128 / (p.TW * p.TH);
    if (p.TW * p.TH * p.NB != 128) return fail(RQB200_EINVAL, "conv_tc: feature map extent must be a power of two");
    p.tiles_x = W / p.TW; p.tiles_y = H / p.TH; p.tiles_b = (int)ceil_div(B, p.NB);
    const int BN = Cout <= 16 ? 16 : (Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64));
    p.n_tiles_n = (int)ceil_div(Cout, BN);
    p.bias = bias; p.residual = residual; p.out = out; p.out_nchw = out_nchw;
    if (gn_part != nullptr) {
        if (!conv_tc_gn_fusable(H, W, Cout) || out_nchw) return fail(RQB200_EINVAL, "conv_tc: GroupNorm statistics cannot be fused for this shape");
        p.gn_part = gn_part;
        p.gn_chunks = H * W / 32;
    }
    CUtensorMap tmA, tmB;
    RQB_TRY(make_tmap_4d_nhwc(&tmA, X16, (uint64_t)Cin, (uint64_t)W * stride, (uint64_t)H * stride, (uint64_t)B, 64, (uint32_t)p.TW,
                              (uint32_t)p.TH, (uint32_t)p.NB, (uint32_t)stride));
    RQB_TRY(make_tmap_2d(&tmB, W16, 1, (uint64_t)ks * ks * Cin, (uint64_t)Cout, (uint64_t)ks * ks * Cin * 2, 64, (uint32_t)BN));
    const int n_sm = sm_count();
    if (X16lo != nullptr && W16lo != nullptr) {
        CUtensorMap tmAlo, tmBlo;
        RQB_TRY(make_tmap_4d_nhwc(&tmAlo, X16lo, (uint64_t)Cin, (uint64_t)W * stride, (uint64_t)H * stride, (uint64_t)B, 64,
                                  (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.NB, (uint32_t)stride));
        RQB_TRY(make_tmap_2d(&tmBlo, W16lo, 1, (uint64_t)ks * ks * Cin, (uint64_t)Cout, (uint64_t)ks * ks * Cin * 2, 64, (uint32_t)BN));
        switch (BN) {
            case 16: return launch_conv_tc_t<16, 5, 3>(tmA, tmB, tmAlo, tmBlo, p, n_sm, st);
            case 64: return launch_conv_tc_t<64, 4, 3>(tmA, tmB, tmAlo, tmBlo, p, n_sm, st);
            case 128: return launch_conv_tc_t<128, 3, 3>(tmA, tmB, tmAlo, tmBlo, p, n_sm, st);
            default: return launch_conv_tc_t<256, 2, 3>(tmA, tmB, tmAlo, tmBlo, p, n_sm, st);
        }
    }
    switch (BN) {
        case 16: return launch_conv_tc_t<16, 8, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
        case 64: return launch_conv_tc_t<64, 8, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
        case 128: return launch_conv_tc_t<128, 6, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
        default: return launch_conv_tc_t<256, 4, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
    }
}

// Rows GEMM through the persistent conv kernel: out[m, n] = act(sum_k X[m,k] W[n,k] + bias[n]) (+ residual[m,n]) for M token rows
// -- a 1x1 "conv" over ceil(M/128) images of 16x8 pixels.  The batched prefill / teacher-forced forward passes of the AR tier
// (csrc/ar_fast.cu) use it for M > 256: persistent CTAs, 128 x BN tiles, double-buffered TMEM accumulators whose epilogue
// overlaps the next tile's main loop -- what gemm_tc_kernel (a weight streamer built for M <= 256) does not have.
// X [M_alloc, K] 16-bit with M_alloc >= ceil(M/128)*128 rows readable; exactly one of out_f32 / out_16 non-null;
// residual (f32, may alias out_f32) only with out_f32.  N_out % 128 == 0, K % 64 == 0.
int launch_rows_gemm_tc(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32, void* out_16,
                        int gelu, int fmt, int64_t M, int N_out, int K, cudaStream_t st) {
    if (N_out % 128 != 0 || K % 64 != 0 || M < 1 || (out_f32 == nullptr) == (out_16 == nullptr) || bias == nullptr)
        return fail(RQB200_EINVAL, "rows_gemm_tc: need N_out % 128 == 0, K % 64 == 0, a bias and exactly one output");
    // 256 x 256 tiles on CTA pairs when the shape allows (RQB200_ROWS_GEMM_1CTA=1, read once: this kernel for every shape)
    static const bool one_cta = [] { const char* e = std::getenv("RQB200_ROWS_GEMM_1CTA"); return e && e[0] == '1'; }();
    if (!one_cta && rows_gemm2_supported(M, N_out, K))
        return launch_rows_gemm2_tc(X16, W16, bias, residual, out_f32, out_16, gelu, fmt, M, N_out, K, st);
    ConvTcParams p = {};
    p.B = (int)ceil_div(M, 128); p.H = 8; p.W = 16; p.Cin = K; p.Cout = N_out; p.ks = 1; p.stride = 1;
    p.TW = 16; p.TH = 8; p.NB = 1;
    p.tiles_x = 1; p.tiles_y = 1; p.tiles_b = p.B;
    const int BN = N_out % 256 == 0 ? 256 : 128;
    p.n_tiles_n = N_out / BN;
    p.bias = bias; p.residual = residual; p.out = out_f32; p.out_nchw = 0;
    p.out16 = out_16; p.gelu = gelu; p.fmt = fmt; p.m_rows = M;
    CUtensorMap tmA, tmB;
    RQB_TRY(make_tmap_4d_nhwc(&tmA, X16, (uint64_t)K, 16, 8, (uint64_t)p.B, 64, 16, 8, 1, 1));
    RQB_TRY(make_tmap_2d(&tmB, W16, 1, (uint64_t)K, (uint64_t)N_out, (uint64_t)K * 2, 64, (uint32_t)BN));
    const int n_sm = sm_count();
    if (BN == 256) return launch_conv_tc_t<256, 4, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
    return launch_conv_tc_t<128, 6, 1>(tmA, tmB, tmA, tmB, p, n_sm, st);
}

// ------------------------------------------------------------------------------------------------ fp16 operand producers
// 4 fp32 values -> fp16 hi (+ optional fp16 lo = value - hi)
__device__ __forceinline__ void store_split4(__half* hi, __half* lo, const float (&v)[4]) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(hi) = pk;
    if (lo) {
        float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
        pk.x = *reinterpret_cast<uint32_t*>(&l0);
        pk.y = *reinterpret_cast<uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(lo) = pk;
    }
}

// one warp per (image, group): reduce the per-chunk fp64 partial sums once (instead of once per apply CTA); fixed order
__global__ void gn_finalize_kernel(double* __restrict__ part, int B, int nchunks, double n, double eps) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= B * 32) return;
    const int b = i / 32, g = i % 32;
    double ts = 0.0, tss = 0.0;
    for (int c = lane; c < nchunks; c += 32) {
        const double* o = part + (((int64_t)b * nchunks + c) * 32 + g) * 2;
        ts += o[0];
        tss += o[1];
    }
    ts = warp_sum_d(ts);
    tss = warp_sum_d(tss);
    if (lane != 0) return;
    double mean = ts / n, var = tss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    double* fin = part + (int64_t)B * nchunks * 64 + (int64_t)i * 2;
    fin[0] = mean;
    fin[1] = 1.0 / sqrt(var + eps);
}

// GroupNorm apply (+SiLU) from the fp64 partial statistics of gn_stats_kernel, writing the fp16 NHWC conv operand
__global__ void __launch_bounds__(256) gn_apply_f16_kernel(const float* __restrict__ X, const double* __restrict__ part,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           __half* __restrict__ Y, __half* __restrict__ Ylo, int HW, int C,
                                                           float eps, int silu, int nchunks) {
    __shared__ float s_mean[32], s_rstd[32];
    const int b = blockIdx.y;
    const int cg = C / 32;
    if (threadIdx.x < 32) {                                        // finalised by gn_finalize_kernel: (mean, rstd) per (b, group)
        const double* fin = part + (int64_t)gridDim.y * nchunks * 64 + ((int64_t)b * 32 + threadIdx.x) * 2;
        s_mean[threadIdx.x] = (float)fin[0];
        s_rstd[threadIdx.x] = (float)fin[1];
    }
    __syncthreads();
    (void)eps;
    // thread <-> 8 consecutive channels (two float4 loads, one 16 B store per output tensor); 32-bit indexing per image
    const int c8n = C / 8;
    const unsigned total8 = (unsigned)HW * (unsigned)c8n;
    const float* Xb = X + (int64_t)b * HW * C;
    __half* Yb = Y + (int64_t)b * HW * C;
    __half* Lb = Ylo ? Ylo + (int64_t)b * HW * C : nullptr;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total8; i += gridDim.x * 256u) {
        const int c = (int)(i % (unsigned)c8n) * 8;
        const float4 x0 = __ldcs(reinterpret_cast<const float4*>(Xb + (size_t)i * 8));
        const float4 x1 = __ldcs(reinterpret_cast<const float4*>(Xb + (size_t)i * 8 + 4));
        const float4 ga0 = *reinterpret_cast<const float4*>(gamma + c), ga1 = *reinterpret_cast<const float4*>(gamma + c + 4);
        const float4 be0 = *reinterpret_cast<const float4*>(beta + c), be1 = *reinterpret_cast<const float4*>(beta + c + 4);
        const int g0 = c / cg, g1 = (c + 4) / cg;                 // cg % 4 == 0 -> each half shares a group
        const float m0 = s_mean[g0], r0 = s_rstd[g0], m1 = s_mean[g1], r1 = s_rstd[g1];
        float v[8] = {(x0.x - m0) * r0 * ga0.x + be0.x, (x0.y - m0) * r0 * ga0.y + be0.y, (x0.z - m0) * r0 * ga0.z + be0.z,
                      (x0.w - m0) * r0 * ga0.w + be0.w, (x1.x - m1) * r1 * ga1.x + be1.x, (x1.y - m1) * r1 * ga1.y + be1.y,
                      (x1.z - m1) * r1 * ga1.z + be1.z, (x1.w - m1) * r1 * ga1.w + be1.w};
        if (silu) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = v[k] / (1.0f + __expf(-v[k]));
        }
        __half2 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
            const float2 f = __half22float2(h[k]);
            l[k] = __floats2half2_rn(v[2 * k] - f.x, v[2 * k + 1] - f.y);
        }
        *reinterpret_cast<uint4*>(Yb + (size_t)i * 8) = *reinterpret_cast<uint4*>(h);
        if (Lb) *reinterpret_cast<uint4*>(Lb + (size_t)i * 8) = *reinterpret_cast<uint4*>(l);
    }
}

// fp32 NHWC -> fp16 NHWC, optionally nearest x2 upsampled (layers.py:31-35 folded into the operand producer)
__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ X, __half* __restrict__ Y, __half* __restrict__ Ylo,
                                                       int B, int H, int W, int C, int upsample) {
    const int c4n = C / 4;
    const int Ho = upsample ? 2 * H : H, Wo = upsample ? 2 * W : W;
    const int64_t total4 = (int64_t)B * Ho * Wo * c4n;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % c4n);
        int64_t pix = i / c4n;
        const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((int64_t)Wo * Ho));
        const int ix = upsample ? ox >> 1 : ox, iy = upsample ? oy >> 1 : oy;
        const float4 x = *reinterpret_cast<const float4*>(X + (((int64_t)b * H + iy) * W + ix) * C + c4 * 4);
        const float v[4] = {x.x, x.y, x.z, x.w};
        store_split4(Y + i * 4, Ylo ? Ylo + i * 4 : nullptr, v);
    }
}

// fused_chunks > 0: the partial statistics were already written by the producing conv's epilogue (fused_chunks = HW / 32 partials
// per image and group); otherwise gn_stats_kernel computes them (HW / 256 partials)
int launch_groupnorm_f16(const float* X, const float* gamma, const float* beta, void* Y16, void* Y16lo, double* stats_ws, int B,
                         int HW, int C, int silu, cudaStream_t st, int fused_chunks) {
    if (C % 128 != 0) return fail(RQB200_EINVAL, "groupnorm_f16: C % 128 != 0");
    const int nchunks = fused_chunks > 0 ? fused_chunks : (int)ceil_div(HW, 256);
    if (fused_chunks <= 0) RQB_TRY(launch_gn_stats(X, stats_ws, B, HW, C, st));
    gn_finalize_kernel<<<(unsigned)ceil_div(B * 32 * 32, 128), 128, 0, st>>>(stats_ws, B, nchunks, (double)HW * (C / 32), 1e-6);
    RQB_TRY(check_launch("gn_finalize"));
    int gx = (int)std::min<int64_t>(ceil_div((int64_t)HW * C / 8, 256), 1024);
    gn_apply_f16_kernel<<<dim3(gx, B), 256, 0, st>>>(X, stats_ws, gamma, beta, (__half*)Y16, (__half*)Y16lo, HW, C, 1e-6f, silu, nchunks);
    return check_launch("gn_apply_f16");
}

int launch_cast_f16(const float* X, void* Y16, void* Y16lo, int B, int H, int W, int C, int upsample, cudaStream_t st) {
    if (C % 4 != 0) return fail(RQB200_EINVAL, "cast_f16: C % 4 != 0");
    int64_t total4 = (int64_t)B * H * W * C / 4 * (upsample ? 4 : 1);
    int gx = (int)std::min<int64_t>(ceil_div(total4, 256), 148 * 16);
    cast_f16_kernel<<<gx, 256, 0, st>>>(X, (__half*)Y16, (__half*)Y16lo, B, H, W, C, upsample);
    return check_launch("cast_f16");
}

}  // namespace rqb

// diagnostic entry point: one conv through the tcgen05 path (tests/test_gpu_tc.py)
// out_nchw bit 0: NCHW output; bits 8.. : stride (0/1 -> 1, 2 -> the Downsample conv; then H, W are the OUTPUT extent)
extern "C" int rqb200_dbg_conv_tc(const void* X16, const void* W16, const void* X16lo, const void* W16lo, const float* bias,
                                  const float* residual, float* out, int B, int H, int W, int Cin, int Cout, int ks, int out_nchw,
                                  void* stream) {
    const int stride = (out_nchw >> 8) > 1 ? (out_nchw >> 8) : 1;
    if (!rqb::conv_tc_supported(H, W, Cin, Cout, ks, stride, 0)) return rqb::fail(RQB200_EINVAL, "conv_tc: unsupported shape");
    return rqb::launch_conv_tc(X16, W16, X16lo, W16lo, bias, residual, out, B, H, W, Cin, Cout, ks, out_nchw & 1, (cudaStream_t)stream,
                               stride, nullptr);
}

// diagnostic entry point: the rows GEMM (tests/test_gpu_tc.py).  X must have ceil(M/128)*128 readable rows.
extern "C" int rqb200_dbg_rows_gemm(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32,
                                    void* out_16, int gelu, int fmt, int64_t M, int N_out, int K, void* stream) {
    return rqb::launch_rows_gemm_tc(X16, W16, bias, residual, out_f32, out_16, gelu, fmt, M, N_out, K, (cudaStream_t)stream);
}
