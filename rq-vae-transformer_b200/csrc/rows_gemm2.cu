// Large-M GEMM of the batched prefill / teacher-forced forward passes (reference: every nn.Linear of
// transformers.py:113-188 / attentions.py:60-104 at M = B*T token rows) on CTA PAIRS:
//
//   out[m, n] = act(sum_k X[m,k] W[n,k] + bias[n]) (+ residual[m,n])        X [M,K], W [N,K] 16-bit (fp16 / bf16), fp32 accumulate
//
// tcgen05.mma.cta_group::2, one 256 x 256 x 16 instruction per K step for two SMs: CTA r of the pair stages rows [128 r, 128 r + 128)
// of the 256-row X tile and rows [128 r, +128) of the 256-row W tile (its half of the N extent), the leader CTA's single thread
// issues the MMAs for both, and each CTA's TMEM receives the accumulator rows of its own 128 X rows.  Per 64-wide K block an SM
// therefore pulls 32 KB out of L2 for 512 tensor-pipe cycles (62.5 B/clk at full rate); the single-CTA 128 x 256 tile of
// conv_tc_kernel needs 48 KB for the same 512 cycles (94 B/clk).  What limits both is the L2 -> SM fabric with all 148 SMs loading:
// ncu reads 453 MB of l1tex__m_xbar2l1tex_read_bytes for the 52 us qkv launch of this kernel = 8.7 TB/s = ~45 B/clk per SM, at 66 %
// tensor-pipe activity -- the pair form needs a third fewer bytes per flop and is that much closer to the tensor rate.
//
// Persistent: one CTA pair per two SMs, tiles strided over the pairs; a 6-deep TMA ring; TWO accumulators (2 x 256 TMEM columns,
// the whole 512-column TMEM) so that the eight epilogue warps of each CTA drain tile i while the MMAs of tile i+1 run.
// Barriers: full[s] lives in the leader (both CTAs' TMA loads complete_tx on it); the MMA thread's tcgen05.commit multicasts the
// "slot free" / "accumulator ready" arrivals to both CTAs; the sixteen epilogue warps of the pair arrive on the leader's tempty[a].
#include "kernels.h"
#include "tc_common.cuh"

namespace rqb {

constexpr int G2_THREADS = 320;                        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int G2_STAGES = 6;
constexpr int G2_TILE_BYTES = 128 * 64 * 2;            // one 128-row K-major operand tile of one K block
constexpr int G2_STAGE_BYTES = 2 * G2_TILE_BYTES;      // this CTA's X rows + this CTA's W rows

struct RowsGemm2Params {
    int64_t M;
    int N, K;
    int m_tiles, n_tiles;          // 256-row / 256-column tiles
    const float* bias;             // [N]
    const float* residual;         // [M,N] f32 or null (f32 output only; may alias out)
    float* out;                    // [M,N] f32, or
    void* out16;                   // [M,N] 16-bit (fmt), optionally through GELU
    int gelu, fmt;
};

namespace g2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is counted on the LEADER CTA's mbarrier (same shared-memory offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t hint) {
    const uint32_t mbar = tc::smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {      // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}
// arrive on the LEADER's copy of a barrier (from either CTA of the pair), release at cluster scope
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(tc::smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(tc::smem_u32(bar)),
        "r"(parity)
        : "memory");
}
}  // namespace g2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
rows_gemm2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, RowsGemm2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t* empty = full + G2_STAGES;
    uint64_t* tfull = empty + G2_STAGES;     // [2]
    uint64_t* tempty = tfull + 2;            // [2]  (the leader's copy is the one that counts)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    tc::pdl_launch_dependents();            // the next launch of the pass may set itself up while this one runs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = g2::cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nkb = p.K / 64;
    const int total = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmX);
        tc::prefetch_tmap(&tmW);
        for (int s = 0; s < G2_STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 16); }
        tc::fence_barrier_init();
    }
    if (warp == 1) g2::tmem_alloc_pair(tmem_slot, 512);
    tc::tc_fence_before();
    g2::cluster_sync();                      // barriers initialised and TMEM allocated in both CTAs before anyone signals the peer
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tc::pdl_wait();                          // barriers, TMEM and the cluster hand-shake are done ahead of the upstream kernel's end

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer (both CTAs): this CTA's 128 X rows and 128 W rows of every K block
            uint32_t it = 0;
            for (int tile = pair; tile < total; tile += npairs) {
                const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
                const int row0 = mt * 256 + (int)rank * 128, col0 = nt * 256 + (int)rank * 128;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % G2_STAGES;
                    tc::mbar_wait(&empty[s], ((it / G2_STAGES) & 1) ^ 1);
                    if (leader) tc::mbar_expect_tx(&full[s], 2 * G2_STAGE_BYTES);       // four tiles: two from each CTA
                    uint8_t* st = smem + s * G2_STAGE_BYTES;
                    g2::tma_load_2d_pair(st, &tmX, &full[s], kb * 64, row0, tc::L2_EVICT_NORMAL);
                    g2::tma_load_2d_pair(st + G2_TILE_BYTES, &tmW, &full[s], kb * 64, col0, tc::L2_EVICT_LAST);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // ---- MMA issuer (leader CTA, one thread): 256 x 256 x 16 per instruction over both SMs
            const uint32_t idesc = tc::umma_idesc(256, 256, p.fmt);
            uint32_t it = 0, tcount = 0;
            for (int tile = pair; tile < total; tile += npairs, tcount++) {
                const uint32_t as = tcount & 1;
                g2::mbar_wait_cluster(&tempty[as], ((tcount >> 1) & 1) ^ 1);        // both CTAs' epilogues have drained this accumulator
                tc::tc_fence_after();
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % G2_STAGES;
                    tc::mbar_wait(&full[s], (it / G2_STAGES) & 1);
                    tc::tc_fence_after();
                    const uint32_t a = tc::smem_u32(smem + s * G2_STAGE_BYTES), b = a + G2_TILE_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        g2::umma_f16_pair(tmem_base + as * 256, tc::umma_desc_k128(a + j * 32), tc::umma_desc_k128(b + j * 32), idesc,
                                          (kb > 0 || j > 0) ? 1u : 0u);
                    g2::umma_commit_pair(&empty[s]);                               // slot free in both CTAs
                    if (kb == nkb - 1) g2::umma_commit_pair(&tfull[as]);           // accumulator ready in both CTAs
                }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps 2..9 (both CTAs): TMEM lane quarter = warp % 4 (two warps per quarter, 128 of the 256 columns each),
        //      thread <-> one X row.  The fp32 + residual form takes 64 columns per step so that 16 residual loads of a thread are in
        //      flight at once (at 16 columns per step their latency, ~16 dependent round trips per tile, was the kernel's bound).
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        uint32_t tcount = 0;
        for (int tile = pair; tile < total; tile += npairs, tcount++) {
            const uint32_t as = tcount & 1;
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int64_t m = (int64_t)mt * 256 + (int64_t)rank * 128 + r;
            const bool valid = m < p.M;
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + (uint32_t)(half * 128);
            const int nbase = nt * 256 + half * 128;
            if (lane == 0) tc::mbar_wait(&tfull[as], (tcount >> 1) & 1);
            __syncwarp();
            tc::tc_fence_after();
            if (p.out16 != nullptr) {
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 16) {
                    uint32_t v[16];
                    tc::tmem_ld16(tacc + (uint32_t)c0, v);
                    tc::tmem_ld_wait();
                    if (!valid) continue;
                    const int n0 = nbase + c0;
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
                    uint4* o16 = reinterpret_cast<uint4*>(reinterpret_cast<h16*>(p.out16) + m * p.N + n0);
                    o16[0] = pk[0];
                    o16[1] = pk[1];
                }
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 64) {
                    const int n0 = nbase + c0;
                    float4 rr[16];
                    if (p.residual != nullptr && valid) {
                        const float4* rs = reinterpret_cast<const float4*>(p.residual + m * p.N + n0);
#pragma unroll
                        for (int i = 0; i < 16; i++) rr[i] = rs[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) tc::tmem_ld16(tacc + (uint32_t)(c0 + g * 16), v[g]);
                    tc::tmem_ld_wait();
                    if (!valid) continue;
                    float4* o = reinterpret_cast<float4*>(p.out + m * p.N + n0);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + 4 * i);
                        const uint32_t* vv = &v[i >> 2][(i & 3) * 4];
                        o[i] = make_float4(__uint_as_float(vv[0]) + bb.x + rr[i].x, __uint_as_float(vv[1]) + bb.y + rr[i].y,
                                           __uint_as_float(vv[2]) + bb.z + rr[i].z, __uint_as_float(vv[3]) + bb.w + rr[i].w);
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) g2::mbar_arrive_leader(&tempty[as]);        // 16 epilogue warps of the pair -> accumulator free
        }
    }
    tc::tc_fence_before();
    g2::cluster_sync();                      // no CTA leaves (or frees TMEM) while its peer may still signal it
    if (warp == 1) g2::tmem_dealloc_pair(tmem_base, 512);
}

// true when the pair kernel takes the shape (launch_rows_gemm_tc falls back to the single-CTA persistent kernel otherwise)
bool rows_gemm2_supported(int64_t M, int N_out, int K) { return M >= 512 && N_out % 256 == 0 && K % 64 == 0; }

int launch_rows_gemm2_tc(const void* X16, const void* W16, const float* bias, const float* residual, float* out_f32, void* out_16,
                         int gelu, int fmt, int64_t M, int N_out, int K, cudaStream_t st) {
    if (!rows_gemm2_supported(M, N_out, K) || (out_f32 == nullptr) == (out_16 == nullptr) || bias == nullptr)
        return fail(RQB200_EINVAL, "rows_gemm2: need M >= 512, N_out % 256 == 0, K % 64 == 0, a bias and exactly one output");
    RowsGemm2Params p = {};
    p.M = M; p.N = N_out; p.K = K;
    p.m_tiles = (int)ceil_div(M, 256); p.n_tiles = N_out / 256;
    p.bias = bias; p.residual = residual; p.out = out_f32; p.out16 = out_16; p.gelu = gelu; p.fmt = fmt;
    CUtensorMap tmX, tmW;
    // rows beyond M are the tensor map's out-of-bounds zero fill (never stored: the epilogue checks m < M)
    RQB_TRY(make_tmap_2d(&tmX, X16, 1, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, 64, 128));
    RQB_TRY(make_tmap_2d(&tmW, W16, 1, (uint64_t)K, (uint64_t)N_out, (uint64_t)K * 2, 64, 128));
    constexpr size_t smem = (size_t)G2_STAGES * G2_STAGE_BYTES + 1024 + 256;
    RQB_ENSURE_SMEM(smem, rows_gemm2_kernel);
    int dev = 0, n_sm = 0;
    RQB_CUDA(cudaGetDevice(&dev));
    RQB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int total = p.m_tiles * p.n_tiles;
    const int pairs = std::max(1, std::min(total, n_sm / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RQB_CUDA(cudaLaunchKernelEx(&cfg, rows_gemm2_kernel, tmX, tmW, p));
    g_launches++;
    return 0;
}

}  // namespace rqb
