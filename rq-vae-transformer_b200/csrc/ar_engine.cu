// P3 host orchestration -- RQTransformer.sample as one device-side loop (no host sync, no per-token Python).
//
// Mirrors rqvae/models/rqtransformer/transformers.py: sample :294-369 (raster (h,w,d) order, start_loc resume),
// cached_forward :190-287 (body step once per spatial position, head step per depth, classifier), init_cache :289-292.
// What is deliberately different from the reference's execution (not from its arithmetic):
//   * KV caches are pre-allocated [layer][B][nh][Tmax][64] and appended in place (reference: torch.cat per step/layer);
//   * only the NEW position's code embeddings are computed each step (reference re-embeds the whole prefix twice per
//     token, transformers.py:217-220,249-255 -- row-wise identical values);
//   * the sampled code is written on the device and consumed by the next step's kernels; nothing returns to the host.
#include <algorithm>
#include <vector>

#include "kernels.h"

struct rqb200_ar {
    rqb200_ar_config cfg;
    rqb200_ar_weights w;
    std::vector<rqb200_block_weights> body, head;
    int64_t last_launches = 0;
    rqb::ArFast* fast = nullptr;
};

namespace rqb {

struct ArWs {
    float *X, *XN, *QKV, *ATT, *H, *CTX, *TOK, *EMB, *LIN, *LOGITS;
    float *kc_body, *vc_body, *kc_head, *vc_head;
};

static size_t ar_layout(const rqb200_ar_config& c, int B, void* base, size_t cap, ArWs* ws) {
    Arena a(base, cap);
    const int64_t E = c.embed_dim, HW = (int64_t)c.H * c.W, Tb = c.cond_len + HW, D = c.D;
    const int64_t Mmax = (int64_t)B * Tb;                 // worst-case prefill (start_loc resume)
    float* X = a.take<float>(Mmax * E);
    float* XN = a.take<float>(Mmax * E);
    float* QKV = a.take<float>(Mmax * 3 * E);
    float* ATT = a.take<float>(Mmax * E);
    float* H = a.take<float>(Mmax * 4 * E);
    float* CTX = a.take<float>((int64_t)B * E);
    float* TOK = a.take<float>((int64_t)B * E);
    float* EMB = a.take<float>((int64_t)B * HW * D * c.code_dim);
    float* LIN = a.take<float>((int64_t)B * HW * D * E);
    float* LOGITS = a.take<float>((int64_t)B * c.vocab);
    const int64_t per_body = (int64_t)B * c.n_head * Tb * 64, per_head = (int64_t)B * c.n_head * D * 64;
    float* kcb = a.take<float>(per_body * c.n_body);
    float* vcb = a.take<float>(per_body * c.n_body);
    float* kch = a.take<float>(per_head * c.n_head_layers);
    float* vch = a.take<float>(per_head * c.n_head_layers);
    if (ws) *ws = ArWs{X, XN, QKV, ATT, H, CTX, TOK, EMB, LIN, LOGITS, kcb, vcb, kch, vch};
    return a.off + 256;
}

// one transformer stack over M = B*Tn rows held in ws.X (in place).  attentions.py:134-142 per block.
static int run_stack(const rqb200_ar* h, const std::vector<rqb200_block_weights>& blocks, ArWs& ws, int B, int Tn, int T_past,
                     int Tmax, float* kc, float* vc, cudaStream_t st) {
    const rqb200_ar_config& c = h->cfg;
    const int E = c.embed_dim, M = B * Tn, wd = c.weight_dtype;
    const int64_t per = (int64_t)B * c.n_head * Tmax * 64;
    for (size_t l = 0; l < blocks.size(); l++) {
        const rqb200_block_weights& bw = blocks[l];
        RQB_TRY(launch_layernorm(ws.X, E, bw.ln1_w, bw.ln1_b, ws.XN, E, M, E, st));
        RQB_TRY(launch_linear(ws.XN, E, bw.wqkv, wd, bw.bqkv, nullptr, ws.QKV, 3 * E, M, 3 * E, E, 0, st));
        RQB_TRY(launch_attn_cached(ws.QKV, kc + per * l, vc + per * l, ws.ATT, B, Tn, T_past, Tmax, E, c.n_head, st));
        RQB_TRY(launch_linear(ws.ATT, E, bw.wproj, wd, bw.bproj, ws.X, ws.X, E, M, E, E, 0, st));
        RQB_TRY(launch_layernorm(ws.X, E, bw.ln2_w, bw.ln2_b, ws.XN, E, M, E, st));
        RQB_TRY(launch_linear(ws.XN, E, bw.w1, wd, bw.b1, nullptr, ws.H, 4 * E, M, 4 * E, E, 1, st));
        RQB_TRY(launch_linear(ws.H, 4 * E, bw.w2, wd, bw.b2, ws.X, ws.X, E, M, E, 4 * E, 0, st));
    }
    return 0;
}

// positions [idx0, idx_end) of the raster; resume != 0: no prefill, continue on the caches / context left in this workspace
static int ar_sample_impl(rqb200_ar* h, const int64_t* partial, const int64_t* cond, int B, int idx0, int idx_end, int resume,
                          float temperature, const int32_t* top_k, const float* top_p, const float* noise,
                          int64_t noise_stride, float* logits_out, const int64_t* force, int64_t* out, void* wsp,
                          size_t ws_bytes, cudaStream_t st) {
    const rqb200_ar_config& c = h->cfg;
    const rqb200_ar_weights& w = h->w;
    const int E = c.embed_dim, D = c.D, HW = c.H * c.W, C = c.code_dim, V = c.vocab, K = c.codebook_size;
    const int wd = c.weight_dtype, cl = c.cond_len, Tb = cl + HW;
    if (B <= 0) return fail(RQB200_EINVAL, "ar_sample: B must be > 0");
    if (idx0 < 0 || idx_end > HW || idx0 > idx_end) return fail(RQB200_EINVAL, "ar_sample: bad position span");
    ArWs ws;
    size_t need = ar_layout(c, B, wsp, ws_bytes, &ws);
    if (need > ws_bytes) return fail(RQB200_EWORKSPACE, "ar_sample: workspace too small");
    const int64_t code_bytes = (int64_t)B * HW * D * sizeof(int64_t);
    if (!resume && out != partial) RQB_CUDA(cudaMemcpyAsync(out, partial, code_bytes, cudaMemcpyDeviceToDevice, st));   // xs = partial_sample.clone()
    if (idx0 >= idx_end) return 0;

    if (!resume) {
        // ---- prefill: tokens [cond (cl) | xs_emb[0 .. idx0-1]]  (transformers.py:224-239)
        const int Tn0 = cl + idx0;
        RQB_TRY(launch_cond_token(cond, w.cond_emb, w.pos_emb_cond, B, cl, c.vocab_cond, E, Tn0, ws.X, st));
        if (idx0 > 0) {
            RQB_TRY(launch_code_emb(out, w.codebook, B, HW, D, K, C, 0, idx0, ws.EMB, st));
            RQB_TRY(launch_linear(ws.EMB, C, w.w_in, wd, w.b_in, nullptr, ws.LIN, E, B * idx0 * D, E, C, 0, st));
            RQB_TRY(launch_body_token(ws.LIN, w.pos_emb_hw, B, D, E, 0, idx0, cl, Tn0, ws.X, st));
        }
        RQB_TRY(run_stack(h, h->body, ws, B, Tn0, 0, Tb, ws.kc_body, ws.vc_body, st));
        RQB_TRY(launch_row_add(ws.X, (int64_t)Tn0 * E, (int64_t)(Tn0 - 1) * E, nullptr, B, E, ws.CTX, st));   // latents[:, -1]
    }

    int64_t step = 0;
    for (int idx = idx0; idx < idx_end; idx++) {
        if (idx > idx0 || resume) {   // decode step on the token of position idx-1 (transformers.py:240-242)
            RQB_TRY(launch_code_emb(out, w.codebook, B, HW, D, K, C, idx - 1, 1, ws.EMB, st));
            RQB_TRY(launch_linear(ws.EMB, C, w.w_in, wd, w.b_in, nullptr, ws.LIN, E, B * D, E, C, 0, st));
            RQB_TRY(launch_body_token(ws.LIN, w.pos_emb_hw, B, D, E, idx - 1, 1, 0, 1, ws.X, st));
            RQB_TRY(run_stack(h, h->body, ws, B, 1, cl + idx - 1, Tb, ws.kc_body, ws.vc_body, st));
            RQB_CUDA(cudaMemcpyAsync(ws.CTX, ws.X, (size_t)B * E * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        for (int d = 0; d < D; d++) {
            if (d == 0) {
                RQB_TRY(launch_row_add(ws.CTX, E, 0, w.pos_emb_d, B, E, ws.X, st));                         // ctx + pos_emb_d[0]
            } else {
                RQB_TRY(launch_head_cumsum(out, w.codebook, B, HW, D, K, C, idx, d, ws.EMB, st));         // cumsum_{i<d} e_i
                RQB_TRY(launch_linear(ws.EMB, C, w.w_head, wd, w.b_head, nullptr, ws.TOK, E, B, E, C, 0, st));
                RQB_TRY(launch_row_add(ws.TOK, E, 0, w.pos_emb_d + (int64_t)d * E, B, E, ws.X, st));
            }
            RQB_TRY(run_stack(h, h->head, ws, B, 1, d, D, ws.kc_head, ws.vc_head, st));                   // head cache restarts at d==0
            RQB_TRY(launch_layernorm(ws.X, E, w.cls_ln_w, w.cls_ln_b, ws.XN, E, B, E, st));
            float* lg = logits_out ? logits_out + step * (int64_t)B * V : ws.LOGITS;
            RQB_TRY(launch_linear(ws.XN, E, w.w_cls, wd, w.b_cls, nullptr, lg, V, B, V, E, 0, st));
            const float* q = noise ? noise + step * noise_stride : nullptr;
            const int64_t off = (int64_t)idx * D + d;
            RQB_TRY(launch_sample(lg, q, B, V, temperature, top_k[d], top_p[d], out + off, force ? force + off : nullptr,
                                  (int64_t)HW * D, st));
            step++;
        }
    }
    return 0;
}

}  // namespace rqb

extern "C" {

rqb200_ar* rqb200_ar_create(const rqb200_ar_config* cfg, const rqb200_ar_weights* w) {
    if (!cfg || !w) { rqb::set_error("ar_create: null argument"); return nullptr; }
    if (cfg->embed_dim != cfg->n_head * 64) { rqb::set_error("ar_create: embed_dim must be n_head*64"); return nullptr; }
    if (cfg->embed_dim % 64 || cfg->code_dim % 4 || cfg->vocab > 16384 || cfg->cond_len < 1 || cfg->D < 1) {
        rqb::set_error("ar_create: unsupported shape");
        return nullptr;
    }
    if (cfg->weight_dtype != RQB200_F32 && cfg->weight_dtype != RQB200_BF16 && cfg->weight_dtype != RQB200_F16) {
        rqb::set_error("ar_create: weight dtype");
        return nullptr;
    }
    rqb200_ar* h = new rqb200_ar();
    h->cfg = *cfg;
    h->w = *w;
    h->body.assign(w->body, w->body + cfg->n_body);
    h->head.assign(w->head, w->head + cfg->n_head_layers);
    h->w.body = h->body.data();
    h->w.head = h->head.data();
    if (cfg->mode == RQB200_MODE_FAST) {
        if (cfg->weight_dtype == RQB200_F32) { rqb::set_error("ar_create: fast tier needs fp16 or bf16 weights"); delete h; return nullptr; }
        h->fast = rqb::ar_fast_create(h->cfg, h->w, h->body.data(), h->head.data());
        if (!h->fast) { delete h; return nullptr; }
    }
    return h;
}
void rqb200_ar_destroy(rqb200_ar* h) {
    if (h && h->fast) rqb::ar_fast_destroy(h->fast);
    delete h;
}
size_t rqb200_ar_workspace_bytes(const rqb200_ar* h, int B) {
    if (!h || B <= 0) return 0;
    if (h->fast) return rqb::ar_fast_workspace_bytes(h->fast, B);
    return rqb::ar_layout(h->cfg, B, nullptr, 0, nullptr);
}
int rqb200_ar_sample_span(rqb200_ar* h, const int64_t* partial, const int64_t* cond, int B, int idx_begin, int idx_end, int resume,
                          float temperature, const int32_t* top_k_host, const float* top_p_host, const float* noise,
                          int64_t noise_stride, float* logits_out, const int64_t* force_codes, int64_t* out_codes,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || (!partial && !resume) || !out_codes || !top_k_host || !top_p_host || !workspace)
        return rqb::fail(RQB200_EINVAL, "ar_sample: null argument");
    if (rqb200_device_count() <= 0) return rqb::fail(RQB200_ENODEV, "ar_sample: no CUDA device");
    if (h->cfg.weight_dtype != RQB200_F32 && !h->fast) return rqb::fail(RQB200_EINVAL, "ar_sample: 16-bit weights need the fast tier");
    rqb::g_launches = 0;
    int rc;
    if (h->fast)
        rc = rqb::ar_fast_sample(h->fast, partial, cond, B, idx_begin, idx_end, resume, temperature, top_k_host, top_p_host, noise,
                                 noise_stride, logits_out, force_codes, out_codes, workspace, workspace_bytes, (cudaStream_t)stream);
    else
        rc = rqb::ar_sample_impl(h, partial, cond, B, idx_begin, idx_end, resume, temperature, top_k_host, top_p_host, noise,
                                 noise_stride, logits_out, force_codes, out_codes, workspace, workspace_bytes, (cudaStream_t)stream);
    h->last_launches = rqb::g_launches;
    return rc;
}
int rqb200_ar_sample(rqb200_ar* h, const int64_t* partial, const int64_t* cond, int B, int start_h, int start_w,
                     float temperature, const int32_t* top_k_host, const float* top_p_host, const float* noise,
                     int64_t noise_stride, float* logits_out, const int64_t* force_codes, int64_t* out_codes,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return rqb::fail(RQB200_EINVAL, "ar_sample: null argument");
    if (start_h < 0 || start_w < 0 || start_w >= h->cfg.W || start_h > h->cfg.H) return rqb::fail(RQB200_EINVAL, "ar_sample: bad start_loc");
    const int HW = h->cfg.H * h->cfg.W;
    return rqb200_ar_sample_span(h, partial, cond, B, std::min(start_h * h->cfg.W + start_w, HW), HW, 0, temperature, top_k_host,
                                 top_p_host, noise, noise_stride, logits_out, force_codes, out_codes, workspace, workspace_bytes, stream);
}
size_t rqb200_ar_forward_workspace_bytes(const rqb200_ar* h, int B) {
    if (!h || !h->fast || B <= 0) return 0;
    return rqb::ar_fast_forward_workspace_bytes(h->fast, B);
}
int rqb200_ar_forward(rqb200_ar* h, const int64_t* codes, const int64_t* cond, int B, float* logits_out, float* cond_logits_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !codes || !logits_out || !workspace) return rqb::fail(RQB200_EINVAL, "ar_forward: null argument");
    if (rqb200_device_count() <= 0) return rqb::fail(RQB200_ENODEV, "ar_forward: no CUDA device");
    if (!h->fast) return rqb::fail(RQB200_EINVAL, "ar_forward: the batched forward is a fast-tier path (exact tier: teacher-forced rqb200_ar_sample)");
    rqb::g_launches = 0;
    int rc = rqb::ar_fast_forward(h->fast, codes, cond, B, logits_out, cond_logits_out, workspace, workspace_bytes, (cudaStream_t)stream);
    h->last_launches = rqb::g_launches;
    return rc;
}
int rqb200_ar_trace(rqb200_ar* h, long long* out_host, int cap_launches, char* names, int names_cap) {
    if (!h || !h->fast || !out_host) return 0;
    cudaDeviceSynchronize();
    return rqb::ar_fast_trace(h->fast, out_host, cap_launches, names, names_cap);
}
int64_t rqb200_ar_last_launches(const rqb200_ar* h) { return h ? h->last_launches : 0; }
}
