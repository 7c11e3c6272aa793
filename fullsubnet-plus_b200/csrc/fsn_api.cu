// C ABI of fsnplus_b200 (include/fsnplus_b200.h): model object, parameter store keyed by the reference's
// state_dict names, packing into kernel layouts, and the forward orchestration.
#include "../../include/fsnplus_b200.h"
#include "fsn_common.cuh"
#include "fsn_kernels.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace fsn;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(x)                                                                                               \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess) return fail(FSN_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    // grow-only; a fresh allocation is zero-filled ON THE STREAM THAT WILL USE IT (a legacy-stream cudaMemset is not ordered
    // against kernels on a non-blocking stream)
    int ensure(size_t n, bool zero, cudaStream_t s = nullptr) {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return (int)e;
        bytes = n;
        if (zero) e = cudaMemsetAsync(p, 0, n, s);
        return (int)e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct ParamSpec { std::string key; int64_t numel; };

struct fsn_model {
    fsn_config cfg;
    std::vector<ParamSpec> specs;
    std::map<std::string, std::vector<float>> host;   // raw fp32 parameters (reference layout)
    std::map<std::string, float*> dev;                // same, on device
    DevBuf arena;
    bool finalized = false;
    int Isb = 0;                                      // sub-band LSTM input size
    // packed LSTM weights
    DevBuf sb_frag[4], sb_bias[4], sb_tc5_stream, sb_tc5_bias;
    DevBuf fb_frag[4], fb_bias[4];
    bool tc5_ok = false;
    // layer-wise tcgen05 path (k_lstm_tc5r.cu): per-layer recurrent streams, permuted input-projection matrices, biases
    bool tc5r_ok = false;
    DevBuf r_stream[4], r_wih[4], r_bias[4], r_gin, r_hseq;
    // Front-end workspaces (everything the full-band stage writes and the sub-band LSTM reads), grow-only, keyed by the last
    // (B, T).  TWO lanes: the pipelined entry points (fsn_model_submit / fsn_model_forward_host_async) run the front end of
    // batch i+1 in lane (i+1)&1 on the front stream while the sub-band LSTM of batch i still reads lane i&1 on the LSTM stream.
    struct Lane {
        int wsB = 0, wsT = 0;
        DevBuf fbin, fbout, xa, xb, y1, y2, stats, mu, ximg, magpad, fbx, hseq;
        DevBuf x0, xr;                                 // time-major fb input / relu'd last residual
        DevBuf fbo;                                    // time-major full-band outputs [(branch, b, t), Cp] (fused-unfold path)
        bool fbo_valid = false;                        // the last forward of this lane wrote fbo instead of fbout
        DevBuf tsse_scale, sb_rowsum;
        DevBuf xn, sigma;                              // pre-normalised inputs / sub-band std for the non-default norm types
        alignas(64) unsigned char mapX0[128], mapXa[128], mapXb[128], mapXr[128], mapY2[128];
        cudaEvent_t ev_front = nullptr, ev_lstm = nullptr;   // front end written / LSTM finished reading this lane
        bool used = false;                             // ev_lstm has been recorded (a pipelined batch ran / may still run in this lane)
        void release_all() {
            DevBuf* all[] = {&fbin, &fbout, &xa, &xb, &y1, &y2, &stats, &mu, &ximg, &magpad, &fbx, &hseq, &x0, &xr, &fbo, &tsse_scale, &sb_rowsum, &xn, &sigma};
            for (auto* b : all) b->release();
        }
    } lane[2];
    int last_lane = 0;
    DevBuf cstate, mask_tmp, stage_in[3], stage_out;   // LSTM-side scratch (one LSTM runs at a time), host-entry staging
    // tcgen05 TCN (FullSubNet+): folded / padded weights, per-block tensor maps, time-major activations
    bool tcn5 = false;
    int Cp = 0, tcnNT = 0, tcnNtiles = 0, num_sms = 148;
    DevBuf tW1, tW2, tWfc, tS1, tS2b, tBfc;            // [8][3][512][Cp], [8][3][Cp][512], [3][Cp][Cp], [8][3][Cp] x2, [3][Cp]
    DevBuf ws_h, ws_bar;                               // weight-stationary full-band LSTM: h exchange buffer, grid barrier
    alignas(64) unsigned char mapW1[8][128], mapW2[8][128], mapWfc[128];
    int64_t launches = 0;
    int last_impl = 0;
    // tuning knobs, read ONCE at fsn_model_create (never on the forward path): FSN_LSTM_IMPL overrides cfg.lstm_impl,
    // FSN_NO_WS=1 keeps the full-band LSTM of fullsubnet.Model off the weight-stationary kernel
    int env_impl = 0;
    int env_split = 0;                                 // FSN_TC5_SPLIT: force the small-batch column split (1 / 2 / 4; 0 = auto)
    // FSN_FRONT_OVERLAP=1: the pipelined entry points run the front end of batch i+1 concurrently with the sub-band LSTM of batch i
    // (two lanes, two streams).  Off by default: measured zero-sum on B200 -- the LSTM kernel runs at the board's power cap, and
    // what the front end gains on the 18 idle SMs the LSTM loses in clock (profiles/r02_front_overlap.txt).
    bool env_front_overlap = false;
    int env_pdl = -1;                                  // FSN_PDL: 0 / 1 / -1 (auto: small batches only)
    double ws_cap_bytes = 48e9;                        // FSN_WS_CAP_GB: larger batches are run as sub-batches (plain forward entry points)
    bool env_no_ws = false, env_no_xfuse = false;      // FSN_NO_XFUSE=1: packed sub-band images instead of the fused unfold (A/B only)
    // pipelined execution: front-end stream, LSTM stream (higher priority), copy-in / copy-out streams, per-slot events
    cudaStream_t s_front = nullptr, s_lstm = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in = nullptr, ev_plain = nullptr;   // caller's inputs ready / last plain forward finished
    bool plain_pending = false;
    cudaEvent_t ev_h2d[2] = {}, ev_d2h[2] = {}, ev_done[2] = {};   // per staging slot: inputs copied / mask copied out / forward finished
    bool d2h_used[2] = {false, false};
    DevBuf a_in[2][3], a_out[2];
    int64_t nsub = 0;
    static const int NEV = 32;
    cudaEvent_t ev0[NEV] = {}, ev1[NEV] = {};               // sub-band LSTM start / end
    cudaEvent_t evf0[NEV] = {}, evf1[NEV] = {};             // front end start / end (same ring index)
    int64_t nfwd = 0;                                 // forwards whose LSTM events were recorded
};

// ---------------------------------------------------------------------------------------------
// parameter registry: exactly the reference's state_dict
// ---------------------------------------------------------------------------------------------
static void add(std::vector<ParamSpec>& v, const std::string& k, int64_t n) { v.push_back({k, n}); }

static void lstm_specs(std::vector<ParamSpec>& v, const std::string& pre, int I, int H, int L, int O, int G) {
    for (int l = 0; l < L; ++l) {                                   // G = 4 gate blocks for nn.LSTM, 3 for nn.GRU
        const std::string s = std::to_string(l);
        add(v, pre + ".sequence_model.weight_ih_l" + s, (int64_t)G * H * (l == 0 ? I : H));
        add(v, pre + ".sequence_model.weight_hh_l" + s, (int64_t)G * H * H);
        add(v, pre + ".sequence_model.bias_ih_l" + s, G * H);
        add(v, pre + ".sequence_model.bias_hh_l" + s, G * H);
    }
    add(v, pre + ".fc_output_layer.weight", (int64_t)O * H);
    add(v, pre + ".fc_output_layer.bias", O);
}

static void build_specs(fsn_model* m) {
    const fsn_config& c = m->cfg;
    const int F = c.num_freqs;
    auto& v = m->specs;
    if (c.model_kind == FSN_KIND_PLUS) {
        const char* sfx[3] = {"", "_real", "_imag"};
        const char* cn[3] = {"smallConv1d", "middleConv1d", "largeConv1d"};
        for (int b = 0; b < 3; ++b) {
            const std::string p = std::string("channel_attention") + sfx[b];
            if (c.channel_attention == FSN_ATTN_ECA) { add(v, p + ".conv.weight", 3); continue; }
            if (c.channel_attention == FSN_ATTN_TSSE) {
                for (int i = 0; i < 3; ++i) {
                    add(v, p + "." + cn[i] + ".0.weight", (int64_t)F * c.kersize[i]);
                    add(v, p + "." + cn[i] + ".0.bias", F);
                }
                add(v, p + ".feature_concate_fc.weight", 3);
                add(v, p + ".feature_concate_fc.bias", 1);
            }
            add(v, p + ".fc1.weight", (int64_t)(F / 2) * F);
            add(v, p + ".fc1.bias", F / 2);
            add(v, p + ".fc2.weight", (int64_t)F * (F / 2));
            add(v, p + ".fc2.bias", F);
        }
        for (int b = 0; b < 3; ++b) {
            const std::string p = std::string("fb_model") + sfx[b];
            for (int i = 0; i < 8; ++i) {
                const std::string q = p + ".sequence_model." + std::to_string(i);
                add(v, q + ".conv1x1.weight", (int64_t)512 * F);
                add(v, q + ".conv1x1.bias", 512);
                add(v, q + ".prelu1.weight", 1);
                add(v, q + ".norm1.weight", 512);
                add(v, q + ".norm1.bias", 512);
                add(v, q + ".depthwise_conv.weight", 512 * 3);
                add(v, q + ".depthwise_conv.bias", 512);
                add(v, q + ".prelu2.weight", 1);
                add(v, q + ".norm2.weight", 512);
                add(v, q + ".norm2.bias", 512);
                add(v, q + ".sconv.weight", (int64_t)F * 512);
                add(v, q + ".sconv.bias", F);
            }
            add(v, p + ".fc_output_layer.weight", (int64_t)F * F);
            add(v, p + ".fc_output_layer.bias", F);
        }
        m->Isb = (2 * c.sb_num_neighbors + 1) + 3 * (2 * c.fb_num_neighbors + 1);
    } else {
        lstm_specs(v, "fb_model", F, c.fb_hidden, c.num_layers, F, c.rnn_type == FSN_RNN_GRU ? 3 : 4);
        m->Isb = (2 * c.sb_num_neighbors + 1) + (2 * c.fb_num_neighbors + 1);
    }
    lstm_specs(v, "sb_model", m->Isb, c.sb_hidden, c.num_layers, c.output_size, c.rnn_type == FSN_RNN_GRU ? 3 : 4);
}

// ---------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------
static uint16_t h_bits(float f) { __half h = __float2half_rn(f); uint16_t b; std::memcpy(&b, &h, 2); return b; }

// mma.sync fragment order (k_lstm_mma.cu): [group of 8 units][k16 step][lane][i.b0 i.b1 f.b0 f.b1 | g.b0 g.b1 o.b0 o.b1]
static void pack_mma_layer(const float* w_ih, const float* w_hh, int Kin, int Kin_pad, int H, std::vector<uint32_t>& out) {
    const int K = Kin_pad + H, ksteps = K / 16, groups = H / 8;
    out.assign((size_t)groups * ksteps * 32 * 8, 0u);
    auto wcat = [&](int row, int k) -> float {
        if (k < Kin_pad) return k < Kin ? w_ih[(size_t)row * Kin + k] : 0.f;
        return w_hh[(size_t)row * H + (k - Kin_pad)];
    };
    for (int g = 0; g < groups; ++g)
        for (int ks = 0; ks < ksteps; ++ks)
            for (int lane = 0; lane < 32; ++lane) {
                uint32_t* dst = &out[(((size_t)g * ksteps + ks) * 32 + lane) * 8];
                const int n = g * 8 + lane / 4, k0 = ks * 16 + (lane % 4) * 2;
                for (int q = 0; q < 4; ++q) {
                    const int row = q * H + n;
                    dst[2 * q] = (uint32_t)h_bits(wcat(row, k0)) | ((uint32_t)h_bits(wcat(row, k0 + 1)) << 16);
                    dst[2 * q + 1] = (uint32_t)h_bits(wcat(row, k0 + 8)) | ((uint32_t)h_bits(wcat(row, k0 + 9)) << 16);
                }
            }
}

static int upload(DevBuf& b, const void* src, size_t bytes) {
    int e = b.ensure(bytes, false);
    if (e) return e;
    return (int)cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
}

// nn.GRU (sequence_model.py:39-46) on the LSTM kernels: rewrite the (r, z, n) parameters as four pseudo-gate blocks
//   r: [W_ir | W_hr], z: [W_iz | W_hz], n_x: [W_in | 0], n_h: [0 | W_hn]   (biases alike)
// in the nn.LSTM layout; the expanded arrays replace the originals in m->host, so every packer downstream is unchanged.
static void expand_gru(fsn_model* m, const std::string& pre, int I, int H, int L) {
    for (int l = 0; l < L; ++l) {
        const std::string s = std::to_string(l);
        const int K = (l == 0) ? I : H;
        auto& wi = m->host[pre + ".sequence_model.weight_ih_l" + s];
        auto& wh = m->host[pre + ".sequence_model.weight_hh_l" + s];
        auto& bi = m->host[pre + ".sequence_model.bias_ih_l" + s];
        auto& bh = m->host[pre + ".sequence_model.bias_hh_l" + s];
        // each array is expanded on its own size, so a later fsn_model_set_param of a single key re-expands just that key
        if (wi.size() == (size_t)3 * H * K) {                       // r, z, n_x input blocks; the n_h input block stays zero
            std::vector<float> w4((size_t)4 * H * K, 0.f);
            std::copy(wi.begin(), wi.end(), w4.begin());
            wi.swap(w4);
        }
        if (wh.size() == (size_t)3 * H * H) {                       // r, z recurrent blocks; n_x recurrent block zero; n_h <- W_hn
            std::vector<float> w4((size_t)4 * H * H, 0.f);
            std::copy(wh.begin(), wh.begin() + (size_t)2 * H * H, w4.begin());
            std::copy(wh.begin() + (size_t)2 * H * H, wh.end(), w4.begin() + (size_t)3 * H * H);
            wh.swap(w4);
        }
        if (bi.size() == (size_t)3 * H) { bi.resize((size_t)4 * H, 0.f); }
        if (bh.size() == (size_t)3 * H) {
            std::vector<float> b4((size_t)4 * H, 0.f);
            std::copy(bh.begin(), bh.begin() + 2 * H, b4.begin());
            std::copy(bh.begin() + 2 * H, bh.end(), b4.begin() + 3 * H);
            bh.swap(b4);
        }
    }
}

static int pack_lstm(fsn_model* m, const std::string& pre, int I, int Ipad, int H, int L, DevBuf* frag, DevBuf* bias) {
    for (int l = 0; l < L; ++l) {
        const std::string s = std::to_string(l);
        const auto& wi = m->host[pre + ".sequence_model.weight_ih_l" + s];
        const auto& wh = m->host[pre + ".sequence_model.weight_hh_l" + s];
        const auto& bi = m->host[pre + ".sequence_model.bias_ih_l" + s];
        const auto& bh = m->host[pre + ".sequence_model.bias_hh_l" + s];
        std::vector<uint32_t> f;
        pack_mma_layer(wi.data(), wh.data(), l == 0 ? I : H, l == 0 ? Ipad : H, H, f);
        int e = upload(frag[l], f.data(), f.size() * 4);
        if (e) return e;
        std::vector<float> b(4 * H);
        for (int i = 0; i < 4 * H; ++i) b[i] = bi[i] + bh[i];
        e = upload(bias[l], b.data(), b.size() * 4);
        if (e) return e;
    }
    return 0;
}


// tf32 rounding (round-to-nearest on the 13 dropped mantissa bits) so the tensor core's truncation is exact
static float to_tf32(float x) {
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x00000fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    float y; std::memcpy(&y, &u, 4);
    return y;
}

// FullSubNet+ TCN weights for the tcgen05 GEMMs (k_gemm_tc5.cu): channel dim padded to Cp (multiple of 32),
// the three branches stacked, gLN2 folded into the second 1x1 convolution:
//   conv(W2, gLN(y)) = rstd * (W2 diag(gamma2)) y - mean * rstd * s1 + s2,  s1[n] = sum_c W2'[n,c],  s2[n] = sum_c W2[n,c] beta2[c]
static int pack_tcn5(fsn_model* m) {
    const fsn_config& c = m->cfg;
    const int F = c.num_freqs, Cp = (F + 31) / 32 * 32, Hd = 512;
    m->Cp = Cp;
    int nt = 1;
    for (; nt <= 64; ++nt) if (Cp % nt == 0 && (Cp / nt) % 16 == 0 && Cp / nt <= 256) break;
    if (nt > 64) { m->tcn5 = false; return FSN_OK; }
    m->tcnNtiles = nt; m->tcnNT = Cp / nt;
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) m->num_sms = prop.multiProcessorCount;
    const char* sfx[3] = {"", "_real", "_imag"};
    std::vector<float> W1((size_t)8 * 3 * Hd * Cp, 0.f), Wfc((size_t)3 * Cp * Cp, 0.f);
    std::vector<__half> W2((size_t)8 * 3 * Cp * Hd, __float2half_rn(0.f));      // second 1x1 conv: fp16 operands (hidden activations are fp16)
    std::vector<float> S1((size_t)8 * 3 * Cp, 0.f), S2b((size_t)8 * 3 * Cp, 0.f), Bfc((size_t)3 * Cp, 0.f);
    for (int blk = 0; blk < 8; ++blk)
        for (int b = 0; b < 3; ++b) {
            const std::string q = std::string("fb_model") + sfx[b] + ".sequence_model." + std::to_string(blk) + ".";
            const auto& w1 = m->host[q + "conv1x1.weight"];
            const auto& w2 = m->host[q + "sconv.weight"];
            const auto& b2 = m->host[q + "sconv.bias"];
            const auto& g2 = m->host[q + "norm2.weight"];
            const auto& be2 = m->host[q + "norm2.bias"];
            float* d1 = &W1[((size_t)blk * 3 + b) * Hd * Cp];
            for (int o = 0; o < Hd; ++o)
                for (int k = 0; k < F; ++k) d1[(size_t)o * Cp + k] = to_tf32(w1[(size_t)o * F + k]);
            __half* d2 = &W2[((size_t)blk * 3 + b) * Cp * Hd];
            for (int n = 0; n < F; ++n) {
                double s1 = 0, s2 = 0;
                for (int k = 0; k < Hd; ++k) {
                    const __half wh = __float2half_rn(w2[(size_t)n * Hd + k] * g2[k]);
                    const float wf = __half2float(wh);
                    d2[(size_t)n * Hd + k] = wh;
                    s1 += (double)wf;
                    s2 += (double)w2[(size_t)n * Hd + k] * (double)be2[k];
                }
                S1[((size_t)blk * 3 + b) * Cp + n] = (float)s1;
                S2b[((size_t)blk * 3 + b) * Cp + n] = (float)(s2 + (double)b2[n]);
            }
        }
    for (int b = 0; b < 3; ++b) {
        const auto& w = m->host[std::string("fb_model") + sfx[b] + ".fc_output_layer.weight"];
        const auto& bb = m->host[std::string("fb_model") + sfx[b] + ".fc_output_layer.bias"];
        for (int n = 0; n < F; ++n) {
            for (int k = 0; k < F; ++k) Wfc[((size_t)b * Cp + n) * Cp + k] = to_tf32(w[(size_t)n * F + k]);
            Bfc[(size_t)b * Cp + n] = bb[n];
        }
    }
    if (upload(m->tW1, W1.data(), W1.size() * 4) || upload(m->tW2, W2.data(), W2.size() * sizeof(__half)) || upload(m->tWfc, Wfc.data(), Wfc.size() * 4) ||
        upload(m->tS1, S1.data(), S1.size() * 4) || upload(m->tS2b, S2b.data(), S2b.size() * 4) || upload(m->tBfc, Bfc.data(), Bfc.size() * 4))
        return fail(FSN_ECUDA, "upload of the TCN weights failed");
    for (int blk = 0; blk < 8; ++blk) {
        if (make_tmap_f32_2d(m->mapW1[blk], static_cast<float*>(m->tW1.p) + (size_t)blk * 3 * Hd * Cp, 3 * Hd, Cp, 256) ||
            make_tmap_f16_2d(m->mapW2[blk], static_cast<__half*>(m->tW2.p) + (size_t)blk * 3 * Cp * Hd, 3 * Cp, Hd, m->tcnNT))
            return fail(FSN_ECUDA, "cuTensorMapEncodeTiled failed for the TCN weights");
    }
    if (make_tmap_f32_2d(m->mapWfc, m->tWfc.p, 3 * Cp, Cp, m->tcnNT)) return fail(FSN_ECUDA, "cuTensorMapEncodeTiled failed (fc)");
    m->tcn5 = true;
    return FSN_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int fsn_version(void) { return 100; }
extern "C" const char* fsn_last_error(void) { return g_err; }

extern "C" int fsn_model_create(const fsn_config* cfg, fsn_model** out) {
    if (!cfg || !out) return fail(FSN_EINVAL, "null argument");
    const fsn_config& c = *cfg;
    if (c.model_kind != FSN_KIND_PLUS && c.model_kind != FSN_KIND_FSN) return fail(FSN_EINVAL, "unknown model_kind %d", c.model_kind);
    if (c.num_freqs > 2048) return fail(FSN_EINVAL, "num_freqs > 2048 is not supported");
    if (c.num_freqs < 4 || c.look_ahead < 0 || c.sb_num_neighbors < 0 || c.fb_num_neighbors < 0) return fail(FSN_EINVAL, "bad geometry");
    if (c.sb_num_neighbors >= c.num_freqs || c.fb_num_neighbors >= c.num_freqs) return fail(FSN_EINVAL, "reflect padding needs neighbors < num_freqs");
    if (c.num_layers < 1 || c.num_layers > 4) return fail(FSN_EINVAL, "num_layers must be 1..4");
    if (c.sb_hidden % 16 || c.sb_hidden < 16) return fail(FSN_EINVAL, "sb_model_hidden_size must be a multiple of 16");
    if (c.model_kind == FSN_KIND_FSN && (c.fb_hidden % 16 || c.fb_hidden < 16)) return fail(FSN_EINVAL, "fb_model_hidden_size must be a multiple of 16");
    if (c.output_size < 1 || c.output_size > 8) return fail(FSN_EINVAL, "output_size must be 1..8");
    if (c.rnn_type != FSN_RNN_LSTM && c.rnn_type != FSN_RNN_GRU) return fail(FSN_EINVAL, "unknown rnn_type %d", c.rnn_type);
    if (c.norm_type < FSN_NORM_OFFLINE_LAPLACE || c.norm_type > FSN_NORM_CUMULATIVE_LAYER) return fail(FSN_EINVAL, "unknown norm_type %d", c.norm_type);
    if (c.model_kind == FSN_KIND_PLUS) {
        if (c.channel_attention < FSN_ATTN_TSSE || c.channel_attention > FSN_ATTN_ECA) return fail(FSN_EINVAL, "unknown channel_attention %d", c.channel_attention);
        if (c.subband_num > 1 && c.channel_attention != FSN_ATTN_ECA)
            return fail(FSN_EINVAL, "subband_num > 1 needs the ECA attention (the reference forward raises for the others)");
        if (c.subband_num < 0 || c.subband_num > c.num_freqs / 2) return fail(FSN_EINVAL, "bad subband_num %d", c.subband_num);
        for (int i = 0; i < 3; ++i)
            if (c.channel_attention == FSN_ATTN_TSSE && (c.kersize[i] < 1 || c.kersize[i] > 16)) return fail(FSN_EINVAL, "kersize must be in 1..16");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FSN_ECUDA, "no CUDA device: fsnplus_b200 has no CPU fallback");
    fsn_model* m = new fsn_model();
    m->cfg = c;
    { const char* e = getenv("FSN_LSTM_IMPL"); if (e && *e) m->env_impl = atoi(e); }
    { const char* e = getenv("FSN_NO_WS"); m->env_no_ws = e && atoi(e) != 0; }
    { const char* e = getenv("FSN_NO_XFUSE"); m->env_no_xfuse = e && atoi(e) != 0; }
    { const char* e = getenv("FSN_TC5_SPLIT"); if (e && *e) m->env_split = atoi(e); }
    { const char* e = getenv("FSN_FRONT_OVERLAP"); m->env_front_overlap = e && atoi(e) != 0; }
    // FSN_PDL: programmatic dependent launch of the ~30 kernels of the front-end chain.  0 = never, 1 = always, unset = only for small
    // batches (the column-split regime, where the forward is launch-latency-bound: B = 1 3.21 -> 3.17 ms).  At B = 64 the chain itself gets
    // 0.05 ms shorter but the pipelined step LOSES 0.1-0.4 ms: the denser front end no longer leaves gaps for the side stream's iSTFT
    // kernels, which then run under the power-capped LSTM instead (same-box A/B, scripts/gpu_r2_w.sh).
    { const char* e = getenv("FSN_PDL"); m->env_pdl = e ? (atoi(e) != 0 ? 1 : 0) : -1; }
    { const char* e = getenv("FSN_WS_CAP_GB"); if (e && atof(e) > 0) m->ws_cap_bytes = atof(e) * 1e9; }
    build_specs(m);
    for (int i = 0; i < fsn_model::NEV; ++i) { cudaEventCreate(&m->ev0[i]); cudaEventCreate(&m->ev1[i]); cudaEventCreate(&m->evf0[i]); cudaEventCreate(&m->evf1[i]); }
    { cudaDeviceProp prop; int dev = 0; cudaGetDevice(&dev); if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) m->num_sms = prop.multiProcessorCount; }
    if (m->Isb > 64) { delete m; return fail(FSN_EINVAL, "sub-band input size %d > 64 not supported", m->Isb); }
    *out = m;
    return FSN_OK;
}

extern "C" void fsn_model_destroy(fsn_model* m) {
    if (!m) return;
    cudaDeviceSynchronize();
    DevBuf* all[] = {&m->arena, &m->sb_tc5_stream, &m->sb_tc5_bias, &m->cstate, &m->mask_tmp, &m->stage_in[0], &m->stage_in[1], &m->stage_in[2],
                     &m->stage_out, &m->tW1, &m->tW2, &m->tWfc, &m->tS1, &m->tS2b, &m->tBfc, &m->ws_h, &m->ws_bar};
    for (auto* b : all) b->release();
    for (auto& ln : m->lane) {
        ln.release_all();
        if (ln.ev_front) cudaEventDestroy(ln.ev_front);
        if (ln.ev_lstm) cudaEventDestroy(ln.ev_lstm);
    }
    for (int i = 0; i < 4; ++i) { m->r_stream[i].release(); m->r_wih[i].release(); m->r_bias[i].release(); }
    m->r_gin.release(); m->r_hseq.release();
    if (m->s_front) { cudaStreamDestroy(m->s_front); cudaStreamDestroy(m->s_lstm); cudaEventDestroy(m->ev_in); }
    if (m->ev_plain) cudaEventDestroy(m->ev_plain);
    if (m->s_in) { cudaStreamDestroy(m->s_in); cudaStreamDestroy(m->s_out); for (int i = 0; i < 2; ++i) { cudaEventDestroy(m->ev_h2d[i]); cudaEventDestroy(m->ev_d2h[i]); cudaEventDestroy(m->ev_done[i]); } }
    for (int i = 0; i < 2; ++i) { m->a_out[i].release(); for (int j = 0; j < 3; ++j) m->a_in[i][j].release(); }
    for (int i = 0; i < fsn_model::NEV; ++i) {
        if (m->ev0[i]) cudaEventDestroy(m->ev0[i]);
        if (m->ev1[i]) cudaEventDestroy(m->ev1[i]);
        if (m->evf0[i]) cudaEventDestroy(m->evf0[i]);
        if (m->evf1[i]) cudaEventDestroy(m->evf1[i]);
    }
    for (int i = 0; i < 4; ++i) { m->sb_frag[i].release(); m->sb_bias[i].release(); m->fb_frag[i].release(); m->fb_bias[i].release(); }
    delete m;
}

extern "C" int fsn_model_num_params(const fsn_model* m) { return m ? (int)m->specs.size() : 0; }
extern "C" int fsn_model_param_info(const fsn_model* m, int i, const char** key, int64_t* numel) {
    if (!m || i < 0 || i >= (int)m->specs.size()) return fail(FSN_EINVAL, "bad index");
    if (key) *key = m->specs[i].key.c_str();
    if (numel) *numel = m->specs[i].numel;
    return FSN_OK;
}

extern "C" int fsn_model_set_param(fsn_model* m, const char* key, const float* h, int64_t numel) {
    if (!m || !key || !h) return fail(FSN_EINVAL, "null argument");
    for (const auto& s : m->specs)
        if (s.key == key) {
            if (s.numel != numel) return fail(FSN_EINVAL, "size mismatch for %s: expected %lld, got %lld", key, (long long)s.numel, (long long)numel);
            m->host[key].assign(h, h + numel);
            m->finalized = false;
            return FSN_OK;
        }
    return fail(FSN_EINVAL, "unexpected key %s", key);
}

extern "C" int fsn_model_finalize(fsn_model* m) {
    if (!m) return fail(FSN_EINVAL, "null model");
    const fsn_config& c = m->cfg;
    for (const auto& s : m->specs)
        if (!m->host.count(s.key)) return fail(FSN_ESTATE, "missing parameter %s", s.key.c_str());
    if (c.rnn_type == FSN_RNN_GRU) {
        expand_gru(m, "sb_model", m->Isb, c.sb_hidden, c.num_layers);
        if (c.model_kind == FSN_KIND_FSN) expand_gru(m, "fb_model", c.num_freqs, c.fb_hidden, c.num_layers);
    }
    size_t total = 0;
    for (const auto& s : m->specs) total += (m->host[s.key].size() * 4 + 255) & ~(size_t)255;
    if (m->arena.ensure(total, false)) return fail(FSN_ECUDA, "cudaMalloc of %zu bytes failed", total);
    size_t off = 0;
    for (const auto& s : m->specs) {
        const auto& hv = m->host[s.key];
        float* d = reinterpret_cast<float*>(static_cast<char*>(m->arena.p) + off);
        CK(cudaMemcpy(d, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
        m->dev[s.key] = d;
        off += (hv.size() * 4 + 255) & ~(size_t)255;
    }
    if (pack_lstm(m, "sb_model", m->Isb, 64, c.sb_hidden, c.num_layers, m->sb_frag, m->sb_bias)) return fail(FSN_ECUDA, "packing sb_model failed");
    if (c.model_kind == FSN_KIND_FSN) {
        const int Ipad = (c.num_freqs + 15) / 16 * 16;
        if (pack_lstm(m, "fb_model", c.num_freqs, Ipad, c.fb_hidden, c.num_layers, m->fb_frag, m->fb_bias)) return fail(FSN_ECUDA, "packing fb_model failed");
    }
    m->tc5_ok = lstm_tc5_supported(c.num_layers, c.sb_hidden, m->Isb, c.output_size);
    if (m->tc5_ok) {
        const int H = c.sb_hidden;
        const int64_t bytes = fsn_tc5_weight_stream_bytes(m->Isb, H);
        std::vector<uint16_t> st((size_t)bytes / 2);
        fsn_tc5_pack_weights(m->Isb, H, m->host["sb_model.sequence_model.weight_ih_l0"].data(),
                             m->host["sb_model.sequence_model.weight_hh_l0"].data(),
                             m->host["sb_model.sequence_model.weight_ih_l1"].data(),
                             m->host["sb_model.sequence_model.weight_hh_l1"].data(), st.data());
        if (upload(m->sb_tc5_stream, st.data(), (size_t)bytes)) return fail(FSN_ECUDA, "upload of tcgen05 weight stream failed");
        std::vector<float> bp((size_t)2 * 4 * H);
        for (int l = 0; l < 2; ++l) {
            const auto& bi = m->host["sb_model.sequence_model.bias_ih_l" + std::to_string(l)];
            const auto& bh = m->host["sb_model.sequence_model.bias_hh_l" + std::to_string(l)];
            for (int j = 0; j < H / 32; ++j)
                for (int n = 0; n < 128; ++n) {
                    const int row = fsn_tc5_gate_row(H, j, n);
                    // stored pre-scaled: the kernel evaluates exp2(-log2e * (acc + b)) as one FMA (tanh gate: -2 log2e)
                    const int q = (n % 32) / 8;                        // gate position: i,f,g,o / GRU pseudo-gates r,z,n_x,n_h
                    const float sc = (q == 2 || (q == 3 && c.rnn_type == FSN_RNN_GRU)) ? -2.8853900817779268f : -1.4426950408889634f;
                    bp[(size_t)l * 4 * H + j * 128 + n] = sc * (bi[row] + bh[row]);
                }
        }
        if (upload(m->sb_tc5_bias, bp.data(), bp.size() * 4)) return fail(FSN_ECUDA, "upload failed");
    }
    m->tc5r_ok = !m->tc5_ok && lstm_tc5r_supported(c.sb_hidden, c.output_size) && m->Isb <= 64;
    if (m->tc5r_ok) {
        const int H = c.sb_hidden;
        for (int l = 0; l < c.num_layers; ++l) {
            const std::string q = "sb_model.sequence_model.", sl = std::to_string(l);
            const int Kin = (l == 0) ? m->Isb : H, Kpad = (l == 0) ? 64 : H;
            std::vector<uint16_t> st, wp;
            std::vector<float> bp;
            lstm_tc5r_pack_layer(H, Kin, Kpad, m->host[q + "weight_ih_l" + sl].data(), m->host[q + "weight_hh_l" + sl].data(),
                                 m->host[q + "bias_ih_l" + sl].data(), m->host[q + "bias_hh_l" + sl].data(), c.rnn_type == FSN_RNN_GRU, st, wp, bp);
            if (upload(m->r_stream[l], st.data(), st.size() * 2) || upload(m->r_wih[l], wp.data(), wp.size() * 2) ||
                upload(m->r_bias[l], bp.data(), bp.size() * 4))
                return fail(FSN_ECUDA, "upload of the layer-wise tcgen05 weights failed");
        }
    }
    if (c.model_kind == FSN_KIND_PLUS) {
        int rc = pack_tcn5(m);
        if (rc) return rc;
    }
    m->finalized = true;
    return FSN_OK;
}

static const float* P(fsn_model* m, const std::string& k) { return m->dev.at(k); }


static void fill_ws(fsn_model* m, LstmWsLaunch& w) {
    const fsn_config& c = m->cfg;
    for (int l = 0; l < c.num_layers; ++l) {
        const std::string s = std::to_string(l);
        w.w_ih[l] = P(m, "fb_model.sequence_model.weight_ih_l" + s); w.w_hh[l] = P(m, "fb_model.sequence_model.weight_hh_l" + s);
        w.b_ih[l] = P(m, "fb_model.sequence_model.bias_ih_l" + s); w.b_hh[l] = P(m, "fb_model.sequence_model.bias_hh_l" + s);
    }
    w.L = c.num_layers; w.H = c.fb_hidden; w.I = c.num_freqs; w.Ipad = (c.num_freqs + 15) / 16 * 16; w.fast = c.fast_math; w.gru = c.rnn_type == FSN_RNN_GRU;
}

// lstm_impl = auto: fused two-layer tcgen05 kernel when the geometry fits, else the layer-wise tcgen05 path (k_lstm_tc5r.cu), else
// the generic mma.sync kernel; lstm_impl = mma forces the generic kernel.
static int pick_impl(const fsn_model* m) {
    int impl = m->env_impl ? m->env_impl : m->cfg.lstm_impl;
    if (impl == FSN_LSTM_AUTO) impl = (m->tc5_ok || m->tc5r_ok) ? FSN_LSTM_TCGEN05 : FSN_LSTM_MMA;
    return impl;
}
// Fused unfold: the two-layer tcgen05 kernel builds its input tiles itself (k_lstm_tc5d.cu, x-tile builders) from the window source
// and the full-band outputs -- for the per-utterance norms; the per-frame cumulative norms keep the packed images.
static bool use_xfuse(const fsn_model* m) {
    return pick_impl(m) == FSN_LSTM_TCGEN05 && m->tc5_ok && !m->env_no_xfuse &&
           (m->cfg.norm_type == FSN_NORM_OFFLINE_LAPLACE || m->cfg.norm_type == FSN_NORM_OFFLINE_GAUSSIAN);
}
// tcgen05 requested, but the geometry is outside the fused two-layer kernel: run the layer-wise kernel (k_lstm_tc5r.cu)
static bool use_layerwise(const fsn_model* m) { return pick_impl(m) == FSN_LSTM_TCGEN05 && !m->tc5_ok && m->tc5r_ok; }


static int ensure_ws(fsn_model* m, fsn_model::Lane& ln, int B, int T, cudaStream_t s, cudaStream_t sl) {
    const fsn_config& c = m->cfg;
    const int F = c.num_freqs, Tp = T + c.look_ahead, Pp = (Tp + 3) & ~3;
    const int nbr = (c.model_kind == FSN_KIND_PLUS) ? 3 : 1;
    const size_t act = (size_t)nbr * B * F * Pp * 4;
    const int rows = B * F, ntiles = ((rows + 127) / 128 + 1) / 2 * 2;      // whole CTA pairs (k_lstm_tc5d.cu)
    const bool regeo = (ln.wsB != B || ln.wsT != T);
    int e = 0;
    e |= ln.fbin.ensure(act, true, s);
    e |= ln.fbout.ensure(act, true, s);
    e |= ln.mu.ensure((size_t)B * 4, true, s);
    e |= ln.sigma.ensure((size_t)B * 4, true, s);
    e |= ln.tsse_scale.ensure((size_t)nbr * B * F * 4 * (1 + tsse_row_floats()), true, s);      // per-row scale | row statistics
    e |= ln.sb_rowsum.ensure((size_t)B * 4 * F * 2 * 4, true, s);
    if (c.norm_type != FSN_NORM_OFFLINE_LAPLACE) e |= ln.xn.ensure((size_t)nbr * B * F * Tp * 4, true, s);
    // images are re-zeroed whenever the geometry changes (rows beyond B*F and k >= I must stay zero)
    const size_t img_bytes = (size_t)ntiles * Tp * 16384;
    if (regeo) { ln.ximg.release(); }
    if (!use_xfuse(m)) e |= ln.ximg.ensure(img_bytes, true, s);
    // LSTM-side scratch: shared by both lanes (LSTM launches are serialised on the LSTM stream)
    int ra = 0;
    size_t cs = lstm_mma_cstate_bytes(c.num_layers, rows, c.sb_hidden, &ra);
    size_t cs5 = lstm_tc5_cstate_bytes(ntiles, c.sb_hidden);
    if (use_layerwise(m)) {
        const size_t M = (size_t)ntiles * Tp * 128, csr = lstm_tc5r_cstate_bytes(ntiles, c.sb_hidden);
        if (csr > cs5) cs5 = csr;
        e |= m->r_gin.ensure(M * 4 * c.sb_hidden * 2, false, sl);
        if (c.num_layers > 1) e |= m->r_hseq.ensure(M * c.sb_hidden * 2, false, sl);
    }
    if (c.model_kind == FSN_KIND_FSN) {
        int ra2 = 0;
        const size_t csf = lstm_mma_cstate_bytes(c.num_layers, B, c.fb_hidden, &ra2);
        if (csf > cs) cs = csf;
    }
    e |= m->cstate.ensure(cs > cs5 ? cs : cs5, true, sl);
    if (c.model_kind == FSN_KIND_PLUS) {
        if (!m->tcn5) return fail(FSN_ESTATE, "the TCN weights were not packed");
        const size_t trows = (size_t)3 * B * Tp;
        if (regeo) { ln.x0.release(); ln.xa.release(); ln.xb.release(); ln.xr.release(); ln.y1.release(); ln.y2.release(); }
        e |= ln.x0.ensure(trows * m->Cp * 4, true, s);
        e |= ln.xa.ensure(trows * m->Cp * 4, true, s);
        e |= ln.xb.ensure(trows * m->Cp * 4, true, s);
        e |= ln.xr.ensure(trows * m->Cp * 4, true, s);
        e |= ln.y1.ensure(trows * 512 * sizeof(__half), true, s);
        e |= ln.y2.ensure(trows * 512 * sizeof(__half), true, s);
        e |= ln.stats.ensure((size_t)8 * 2 * 3 * B * 2 * sizeof(double) + (size_t)8 * 3 * B * sizeof(float), true, s);   // gLN sums | per-block stream maxima
        if (use_xfuse(m)) e |= ln.fbo.ensure(trows * m->Cp * 4, true, s);
        if (!e && regeo) {
            if (make_tmap_f32_2d(ln.mapX0, ln.x0.p, trows, m->Cp, 128) || make_tmap_f32_2d(ln.mapXa, ln.xa.p, trows, m->Cp, 128) ||
                make_tmap_f32_2d(ln.mapXb, ln.xb.p, trows, m->Cp, 128) || make_tmap_f32_2d(ln.mapXr, ln.xr.p, trows, m->Cp, 128) ||
                make_tmap_f16_2d(ln.mapY2, ln.y2.p, trows, 512, 128))
                return fail(FSN_ECUDA, "cuTensorMapEncodeTiled failed for the activations");
        }
    } else {
        const int Ipad = (F + 15) / 16 * 16, rows_pad = (B + 63) / 64 * 64;
        e |= ln.magpad.ensure((size_t)B * F * Pp * 4, true, s);
        e |= ln.fbx.ensure((size_t)Tp * rows_pad * Ipad * 2, true, s);
        e |= ln.hseq.ensure((size_t)B * c.fb_hidden * Pp * 4, true, s);
    }
    if (e) return fail(FSN_ECUDA, "workspace allocation failed for B=%d T=%d", B, T);
    ln.wsB = B; ln.wsT = T;
    return FSN_OK;
}

// d_enh != null: write the enhanced spectrum (decompress_cIRM x noisy) instead of the mask -- fused into the epilogue of the tcgen05
// kernels, a separate pass over a scratch mask for the generic kernel
static int run_sb_lstm(fsn_model* m, fsn_model::Lane& ln, int B, int T, float* d_out, const XSrc& xs, const float* d_real, const float* d_imag,
                       float* d_enh, cudaStream_t s) {
    const fsn_config& c = m->cfg;
    const int F = c.num_freqs, Tp = T + c.look_ahead, rows = B * F, ntiles = (rows + 127) / 128;
    const int impl = pick_impl(m);
    m->last_impl = impl;
    if (use_layerwise(m)) {
        // one GEMM (input projection of every row and time step, k_gemm_f16.cu) + one recurrent launch per layer
        const int H = c.sb_hidden, ntp = (ntiles + 1) / 2 * 2;
        const long long M = (long long)ntp * Tp * 128;
        for (int l = 0; l < c.num_layers; ++l) {
            const int K = (l == 0) ? 64 : H;
            const void* X = (l == 0) ? ln.ximg.p : m->r_hseq.p;
            // row-major Gin[M, 4H] = X[M, K] * Wp[4H, K]^T, both operands K-major
            GemmF16Launch g{M, 4 * H, K, static_cast<__half*>(m->r_gin.p), 4LL * H};
            int ge = launch_gemm_f16(X, m->r_wih[l].p, g, m->num_sms, s);
            if (ge) return fail(FSN_ECUDA, "input-projection GEMM launch failed (layer %d): %s", l, cudaGetErrorString((cudaError_t)ge));
            m->launches++;
            LstmTc5rLaunch a{};
            a.wstream = static_cast<const __half*>(m->r_stream[l].p);
            a.bias = static_cast<const float*>(m->r_bias[l].p);
            a.fc_w = P(m, "sb_model.fc_output_layer.weight");
            a.fc_b = P(m, "sb_model.fc_output_layer.bias");
            a.H = H; a.rows = rows; a.Tp = Tp; a.ntiles = ntiles;
            a.gin = static_cast<const __half*>(m->r_gin.p);
            a.hseq = static_cast<__half*>(m->r_hseq.p);
            a.cstate = static_cast<float*>(m->cstate.p);
            a.out = d_out; a.F = F; a.la = c.look_ahead;
            a.act = c.sb_act; a.fast = c.fast_math; a.gru = c.rnn_type == FSN_RNN_GRU; a.last = (l == c.num_layers - 1);
            a.nreal = d_real; a.nimag = d_imag; a.enh = reinterpret_cast<float2*>(d_enh);
            int e = launch_lstm_tc5r(a, s);
            if (e) return fail(FSN_ECUDA, "layer-wise tcgen05 LSTM launch failed (layer %d): %s", l, cudaGetErrorString((cudaError_t)e));
            m->launches++;
        }
    } else if (impl == FSN_LSTM_TCGEN05) {
        if (!m->tc5_ok) return fail(FSN_EINVAL, "tcgen05 LSTM needs hidden %% 64 == 0 and <= 512 (<= 384 for the fused two-layer kernel), input <= 64, output_size 2");
        LstmTc5Launch a{};
        a.wstream = static_cast<const __half*>(m->sb_tc5_stream.p);
        a.bias = static_cast<const float*>(m->sb_tc5_bias.p);
        a.fc_w = P(m, "sb_model.fc_output_layer.weight");
        a.fc_b = P(m, "sb_model.fc_output_layer.bias");
        a.H = c.sb_hidden; a.I = m->Isb; a.rows = rows; a.Tp = Tp;
        a.img = static_cast<const __half*>(ln.ximg.p); a.ntiles = ntiles;
        a.xs = xs;
        a.split = m->env_split;
        a.nreal = d_real; a.nimag = d_imag; a.enh = reinterpret_cast<float2*>(d_enh);
        a.cstate = static_cast<float*>(m->cstate.p);
        a.out = d_out; a.F = F; a.la = c.look_ahead; a.act = c.sb_act; a.fast = c.fast_math; a.gru = c.rnn_type == FSN_RNN_GRU;
        int e = launch_lstm_tc5_dbuf(a, s);
        if (e) return fail(FSN_ECUDA, "tcgen05 LSTM launch failed: %s", cudaGetErrorString((cudaError_t)e));
    } else {
        LstmMmaLaunch a{};
        for (int l = 0; l < c.num_layers; ++l) {
            a.w.wfrag[l] = static_cast<const uint4*>(m->sb_frag[l].p);
            a.w.bias[l] = static_cast<const float*>(m->sb_bias[l].p);
        }
        a.w.fc_w = P(m, "sb_model.fc_output_layer.weight");
        a.w.fc_b = P(m, "sb_model.fc_output_layer.bias");
        a.L = c.num_layers; a.H = c.sb_hidden; a.I = m->Isb; a.Ipad = 64;
        a.rows = rows; a.Tp = Tp;
        a.img = static_cast<const __half*>(ln.ximg.p); a.ntiles = ntiles;
        a.cstate = static_cast<float*>(m->cstate.p);
        lstm_mma_cstate_bytes(c.num_layers, rows, c.sb_hidden, &a.rows_alloc);
        a.out = d_out; a.O = c.output_size; a.F = F; a.la = c.look_ahead; a.act = c.sb_act;
        a.fast = c.fast_math; a.gru = c.rnn_type == FSN_RNN_GRU;
        if (d_enh) {                                                    // generic kernel: mask into a scratch buffer, then one fused post-processing pass
            if (m->mask_tmp.ensure((size_t)B * 2 * F * T * 4, false, s)) return fail(FSN_ECUDA, "allocation failed");
            a.out = static_cast<float*>(m->mask_tmp.p);
        }
        int e = launch_lstm_mma(a, s);
        if (e) return fail(FSN_ECUDA, "mma LSTM launch failed: %s", cudaGetErrorString((cudaError_t)e));
        if (d_enh) { launch_apply_cirm_planar(a.out, d_real, d_imag, reinterpret_cast<float2*>(d_enh), B, F, T, s); m->launches++; }
    }
    m->launches++;
    return FSN_OK;
}

// The whole forward.  Front end (norm, attention, full-band model, sub-band statistics + packing) on stream `s` into lane `ln`;
// the sub-band LSTM on stream `sl` (== s for the plain entry point; the pipelined entry points pass the LSTM stream, ordered
// after the front end by the lane's ev_front).
static int forward_impl(fsn_model* m, fsn_model::Lane& ln, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                        float* d_out, float* d_enh, cudaStream_t s, cudaStream_t sl) {
    if (!m || !d_mag || (!d_out && !d_enh)) return fail(FSN_EINVAL, "null argument");
    if (d_enh && (!d_real || !d_imag)) return fail(FSN_EINVAL, "the enhanced-spectrum output needs the noisy real and imaginary planes");
    if (d_enh && m->cfg.output_size != 2) return fail(FSN_EINVAL, "the enhanced-spectrum output needs output_size 2 (a complex mask)");
    if (!m->finalized) return fail(FSN_ESTATE, "fsn_model_finalize has not been called");
    const fsn_config& c = m->cfg;
    if (B < 1 || T < 1) return fail(FSN_EINVAL, "bad batch/frames");
    if (c.model_kind == FSN_KIND_PLUS && (!d_real || !d_imag)) return fail(FSN_EINVAL, "FullSubNet_Plus.forward needs mag, real and imag");
    const int F = c.num_freqs, Tp = T + c.look_ahead, Pp = (Tp + 3) & ~3;
    // chained (programmatic dependent) launches of this thread's front-end kernels: see env_pdl; "small" = at most 16 row-tile pairs,
    // the regime of the LSTM kernel's column split
    fsn_chain_launch_set(m->env_pdl == 1 || (m->env_pdl < 0 && ((long long)B * F + 255) / 256 <= 16));
    if (c.model_kind == FSN_KIND_PLUS && c.channel_attention == FSN_ATTN_TSSE)
        for (int i = 0; i < 3; ++i)
            if (Tp < c.kersize[i]) return fail(FSN_EINVAL, "sequence shorter than the TSSE kernel size");
    int rc = ensure_ws(m, ln, B, T, s, sl);
    if (rc) return rc;
    m->launches = 0;
    m->last_lane = (int)(&ln - m->lane);
    const int evi = (int)(m->nfwd % fsn_model::NEV);
    cudaEventRecord(m->evf0[evi], s);

    const bool xfuse = use_xfuse(m);
    ln.fbo_valid = xfuse && c.model_kind == FSN_KIND_PLUS;
    XSrc xs{};
    SbPackLaunch sp{};
    sp.B = B; sp.F = F; sp.Tp = Tp; sp.Ns = c.sb_num_neighbors; sp.Nf = c.fb_num_neighbors; sp.P = Pp;
    sp.mu = static_cast<float*>(ln.mu.p);
    sp.sigma = static_cast<float*>(ln.sigma.p);
    sp.rowsum = static_cast<float*>(ln.sb_rowsum.p);
    sp.norm_type = c.norm_type;
    sp.ximg = static_cast<__half*>(ln.ximg.p);
    sp.plain = use_layerwise(m) ? 1 : 0;                              // the layer-wise path feeds the images to a GEMM as a plain matrix
    sp.ntiles = (B * F + 127) / 128;

    if (c.model_kind == FSN_KIND_PLUS) {
        const char* sfx[3] = {"", "_real", "_imag"};
        const char* cn[3] = {"smallConv1d", "middleConv1d", "largeConv1d"};
        TsseLaunch ta{};
        ta.x[0] = d_mag; ta.x[1] = d_real; ta.x[2] = d_imag;
        ta.nbranch = 3; ta.B = B; ta.F = F; ta.T = T; ta.Tp = Tp; ta.P = Pp; ta.attention = 1 + c.channel_attention; ta.sub = c.subband_num;
        for (int i = 0; i < 3; ++i) ta.ksz[i] = c.kersize[i];
        for (int b = 0; b < 3; ++b) {
            const std::string p = std::string("channel_attention") + sfx[b];
            if (c.channel_attention == FSN_ATTN_ECA) { ta.p[b].eca_w = P(m, p + ".conv.weight"); continue; }
            if (c.channel_attention == FSN_ATTN_TSSE) {
                for (int i = 0; i < 3; ++i) {
                    ta.p[b].conv_w[i] = P(m, p + "." + cn[i] + ".0.weight");
                    ta.p[b].conv_b[i] = P(m, p + "." + cn[i] + ".0.bias");
                }
                ta.p[b].cat_w = P(m, p + ".feature_concate_fc.weight"); ta.p[b].cat_b = P(m, p + ".feature_concate_fc.bias");
            }
            ta.p[b].fc1_w = P(m, p + ".fc1.weight"); ta.p[b].fc1_b = P(m, p + ".fc1.bias");
            ta.p[b].fc2_w = P(m, p + ".fc2.weight"); ta.p[b].fc2_b = P(m, p + ".fc2.bias");
        }
        ta.out = static_cast<float*>(ln.fbin.p);
        ta.scale = static_cast<float*>(ln.tsse_scale.p);
        ta.rows = ta.scale + (size_t)ta.nbranch * B * F;
        ta.out_tm = static_cast<float*>(ln.x0.p); ta.Cp = m->Cp;
        if (c.norm_type != FSN_NORM_OFFLINE_LAPLACE) {
            NormLaunch na{};
            na.x[0] = d_mag; na.x[1] = d_real; na.x[2] = d_imag; na.y = static_cast<float*>(ln.xn.p);
            na.nbranch = 3; na.B = B; na.F = F; na.T = T; na.Tp = Tp; na.type = c.norm_type;
            launch_input_norm(na, s); m->launches++;
            for (int b = 0; b < 3; ++b) ta.x[b] = static_cast<const float*>(ln.xn.p) + (size_t)b * B * F * Tp;
            ta.T = Tp; ta.prenorm = 1;                                  // padded frames are part of the normalised signal
        }
        const int Z = 3 * B;
        double* stats = static_cast<double*>(ln.stats.p);
        float* amax = reinterpret_cast<float*>(stats + (size_t)8 * 2 * Z * 2);            // [8][Z]: max |x| of the stream entering block k, per sample
        CK(cudaMemsetAsync(stats, 0, (size_t)8 * 2 * Z * 2 * sizeof(double) + (size_t)8 * Z * sizeof(float), s));
        ta.amax = amax;                                                                    // block 0: written by the gate kernel
        launch_tsse_norm(ta, s); m->launches += 3;

        static const int dil[8] = {1, 2, 5, 9, 1, 2, 5, 9};     // sequence_model.py:47-58
        {
            const int Cp = m->Cp;
            GemmTc5Launch g{};
            g.rows_per_branch = B * Tp; g.tiles_m = (B * Tp + 127) / 128; g.nbranch = 3; g.Tp = Tp; g.B = B;
            const float* curp = static_cast<const float*>(ln.x0.p);
            const unsigned char* curmap = ln.mapX0;
            for (int blk = 0; blk < 8; ++blk) {
                double* st1 = stats + (size_t)(2 * blk) * Z * 2;
                double* st2 = stats + (size_t)(2 * blk + 1) * Z * 2;
                auto key = [&](int b, const char* leaf) { return std::string("fb_model") + sfx[b] + ".sequence_model." + std::to_string(blk) + "." + leaf; };
                GemmTc5Launch g1 = g;
                g1.epi = EPI5_PRELU_STATS; g1.Kp = Cp; g1.NT = 256; g1.ntiles_n = 2; g1.Npad = 512;
                for (int b = 0; b < 3; ++b) { g1.bias[b] = P(m, key(b, "conv1x1.bias")); g1.prelu[b] = P(m, key(b, "prelu1.weight")); }
                g1.stats_out = st1; g1.Y16 = static_cast<__half*>(ln.y1.p); g1.ldY = 512; g1.amax_in = amax + (size_t)blk * Z;
                int e = launch_gemm_tc5(curmap, m->mapW1[blk], g1, m->num_sms, s);
                if (e) return fail(FSN_ECUDA, "TCN GEMM1 launch failed: %s", cudaGetErrorString((cudaError_t)e));
                m->launches++;

                DwTmLaunch dw{};
                dw.X = static_cast<const __half*>(ln.y1.p); dw.Y = static_cast<__half*>(ln.y2.p);
                dw.Z = Z; dw.B = B; dw.C = 512; dw.Tp = Tp; dw.dilation = dil[blk]; dw.causal = c.tcn_causal ? 1 : 0;
                dw.stats_in = st1; dw.stats_out = st2; dw.amax = amax + (size_t)blk * Z;
                for (int b = 0; b < 3; ++b) {
                    dw.gamma[b] = P(m, key(b, "norm1.weight")); dw.beta[b] = P(m, key(b, "norm1.bias"));
                    dw.w[b] = P(m, key(b, "depthwise_conv.weight")); dw.b[b] = P(m, key(b, "depthwise_conv.bias"));
                    dw.prelu[b] = P(m, key(b, "prelu2.weight"));
                }
                e = launch_dwconv_tm(dw, s);
                if (e) return fail(FSN_ECUDA, "TCN depth-wise conv launch failed: %s", cudaGetErrorString((cudaError_t)e));
                m->launches++;

                float* nxtp = (blk & 1) ? static_cast<float*>(ln.xb.p) : static_cast<float*>(ln.xa.p);
                GemmTc5Launch g2 = g;
                g2.epi = EPI5_GLN_RES; g2.Kp = 512; g2.NT = m->tcnNT; g2.ntiles_n = m->tcnNtiles; g2.Npad = Cp;
                for (int b = 0; b < 3; ++b) {
                    g2.bias[b] = static_cast<const float*>(m->tS2b.p) + ((size_t)blk * 3 + b) * Cp;
                    g2.s1[b] = static_cast<const float*>(m->tS1.p) + ((size_t)blk * 3 + b) * Cp;
                }
                g2.stats_in = st2; g2.count_in = (double)512 * Tp;
                g2.Xold = curp; g2.Y = nxtp; g2.ldY = Cp; g2.amax_out = (blk < 7) ? amax + (size_t)(blk + 1) * Z : nullptr; g2.Xrelu = (blk == 7) ? static_cast<float*>(ln.xr.p) : nullptr;
                e = launch_gemm_tc5(ln.mapY2, m->mapW2[blk], g2, m->num_sms, s);
                if (e) return fail(FSN_ECUDA, "TCN GEMM2 launch failed: %s", cudaGetErrorString((cudaError_t)e));
                m->launches++;
                curp = nxtp;
                curmap = (blk & 1) ? ln.mapXb : ln.mapXa;
            }
            GemmTc5Launch g3 = g;
            g3.epi = EPI5_OUT; g3.Kp = Cp; g3.NT = m->tcnNT; g3.ntiles_n = m->tcnNtiles; g3.Npad = Cp;
            for (int b = 0; b < 3; ++b) g3.bias[b] = static_cast<const float*>(m->tBfc.p) + (size_t)b * Cp;
            g3.out = static_cast<float*>(ln.fbout.p); g3.F = F; g3.P = Pp; g3.act = c.fb_act;
            if (xfuse) { g3.out_tm = static_cast<float*>(ln.fbo.p); g3.ldY = Cp; }
            int e = launch_gemm_tc5(ln.mapXr, m->mapWfc, g3, m->num_sms, s);
            if (e) return fail(FSN_ECUDA, "TCN output GEMM launch failed: %s", cudaGetErrorString((cudaError_t)e));
            m->launches++;
        }

        sp.win = static_cast<const float*>(ln.fbin.p); sp.Pw = Pp;      // post-attention mag branch (fullsubnet_plus.py:182)
        sp.nfb = 3;
        for (int b = 0; b < 3; ++b) sp.fb[b] = static_cast<const float*>(ln.fbout.p) + (size_t)b * B * F * Pp;
        if (xfuse) {                                                        // time-major views: window = x0 (branch 0), outputs = fbo
            const long long rowsB = (long long)Tp * m->Cp;
            for (int b = 0; b < 3; ++b) sp.fb[b] = static_cast<const float*>(ln.fbo.p) + (size_t)b * B * rowsB;
            sp.fb_sb = rowsB; sp.fb_sf = 1; sp.fb_st = m->Cp;
            xs.win = static_cast<const float*>(ln.x0.p); xs.win_sb = rowsB; xs.win_sf = 1; xs.win_st = m->Cp;
            for (int b = 0; b < 3; ++b) xs.fb[b] = sp.fb[b];
            xs.fb_sb = rowsB; xs.fb_sf = 1; xs.fb_st = m->Cp;
        }
    } else {
        TsseLaunch ta{};
        ta.x[0] = d_mag; ta.nbranch = 1; ta.B = B; ta.F = F; ta.T = T; ta.Tp = Tp; ta.P = Pp; ta.attention = 0;
        ta.out = static_cast<float*>(ln.fbin.p);
        ta.scale = static_cast<float*>(ln.tsse_scale.p);
        ta.rows = ta.scale + (size_t)ta.nbranch * B * F;
        if (c.norm_type != FSN_NORM_OFFLINE_LAPLACE) {
            NormLaunch na{};
            na.x[0] = d_mag; na.y = static_cast<float*>(ln.xn.p);
            na.nbranch = 1; na.B = B; na.F = F; na.T = T; na.Tp = Tp; na.type = c.norm_type;
            launch_input_norm(na, s); m->launches++;
            ta.x[0] = static_cast<const float*>(ln.xn.p); ta.T = Tp; ta.prenorm = 1;
        }
        launch_tsse_norm(ta, s); m->launches += 3;
        launch_pad_copy(d_mag, static_cast<float*>(ln.magpad.p), B, F, T, Pp, s); m->launches++;
        const int Ipad = (F + 15) / 16 * 16, rows_pad = (B + 63) / 64 * 64;
        launch_fb_pack(static_cast<const float*>(ln.fbin.p), static_cast<__half*>(ln.fbx.p), B, F, Tp, Pp, rows_pad, Ipad, s); m->launches++;
        if (lstm_ws_supported(c.num_layers, c.fb_hidden, Ipad, B, m->num_sms) && !m->env_no_ws) {
            const size_t hb = (size_t)c.num_layers * 2 * 64 * c.fb_hidden * 2;
            if (m->ws_h.ensure(hb, false, s) || m->ws_bar.ensure(64, true, s)) return fail(FSN_ECUDA, "allocation failed");
            CK(cudaMemsetAsync(m->ws_h.p, 0, hb, s));
            LstmWsLaunch w{};
            fill_ws(m, w);
            w.rows = B; w.rows_pad = rows_pad; w.Tp = Tp;
            w.x = static_cast<const __half*>(ln.fbx.p); w.hbuf = static_cast<__half*>(m->ws_h.p); w.cbuf = nullptr;
            w.barrier = static_cast<unsigned int*>(m->ws_bar.p);
            w.hseq = static_cast<float*>(ln.hseq.p); w.P = Pp; w.resume = 0; w.t0 = 0;
            int e = launch_lstm_ws(w, s);
            if (e) return fail(FSN_ECUDA, "weight-stationary full-band LSTM launch failed: %s", cudaGetErrorString((cudaError_t)e));
            m->launches++;
        } else {
        LstmMmaLaunch a{};
        for (int l = 0; l < c.num_layers; ++l) {
            a.w.wfrag[l] = static_cast<const uint4*>(m->fb_frag[l].p);
            a.w.bias[l] = static_cast<const float*>(m->fb_bias[l].p);
        }
        a.L = c.num_layers; a.H = c.fb_hidden; a.I = F; a.Ipad = Ipad; a.rows = B; a.Tp = Tp;
        a.xplain = static_cast<const __half*>(ln.fbx.p); a.rows_pad = rows_pad;
        a.cstate = static_cast<float*>(m->cstate.p);
        lstm_mma_cstate_bytes(c.num_layers, B, c.fb_hidden, &a.rows_alloc);
        a.hseq = static_cast<float*>(ln.hseq.p); a.P = Pp; a.fast = c.fast_math; a.gru = c.rnn_type == FSN_RNN_GRU;
        int e = launch_lstm_mma(a, s);
        if (e) return fail(FSN_ECUDA, "full-band LSTM launch failed: %s", cudaGetErrorString((cudaError_t)e));
        m->launches++;
        }
        ConvLaunch cf{};
        cf.X = static_cast<const float*>(ln.hseq.p); cf.Y = static_cast<float*>(ln.fbout.p);
        cf.Z = B; cf.M = F; cf.K = c.fb_hidden; cf.Tp = Tp; cf.P = Pp; cf.act = c.fb_act;
        cf.W = P(m, "fb_model.fc_output_layer.weight"); cf.bias = P(m, "fb_model.fc_output_layer.bias");
        launch_conv1x1(cf, s); m->launches++;

        sp.win = static_cast<const float*>(ln.magpad.p); sp.Pw = Pp;    // raw padded magnitude (fullsubnet.py:94)
        sp.nfb = 1;
        sp.fb[0] = static_cast<const float*>(ln.fbout.p);
        if (xfuse) {                                                        // frequency-major views of the padded magnitude / the full-band output
            xs.win = sp.win; xs.win_sb = (long long)F * Pp; xs.win_sf = Pp; xs.win_st = 1;
            xs.fb[0] = sp.fb[0]; xs.fb_sb = (long long)F * Pp; xs.fb_sf = Pp; xs.fb_st = 1;
        }
    }
    launch_sb_stats(sp, s); m->launches += (sp.fb_st > 1) ? 3 : 2;
    if (xfuse) {
        xs.nfb = sp.nfb; xs.Ns = sp.Ns; xs.Nf = sp.Nf; xs.mu = sp.mu; xs.sigma = sp.sigma;
        xs.gauss = (c.norm_type == FSN_NORM_OFFLINE_GAUSSIAN) ? 1 : 0;
    } else {
        launch_sb_pack(sp, s); m->launches++;
    }
    cudaEventRecord(m->evf1[evi], s);
    if (sl != s) {                                                      // pipelined: the LSTM stream picks the lane up when the front end is done
        CK(cudaEventRecord(ln.ev_front, s));
        CK(cudaStreamWaitEvent(sl, ln.ev_front, 0));
    }
    cudaEventRecord(m->ev0[evi], sl);
    rc = run_sb_lstm(m, ln, B, T, d_out, xs, d_real, d_imag, d_enh, sl);
    cudaEventRecord(m->ev1[evi], sl);
    if (rc) return rc;
    m->nfwd++;
    CK(cudaGetLastError());
    return FSN_OK;
}

// ---------------------------------------------------------------------------------------------
// pipelined execution: internal streams, lanes
// ---------------------------------------------------------------------------------------------
static int ensure_pipeline(fsn_model* m) {
    if (m->s_front) return FSN_OK;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));                     // hi = numerically lowest = highest priority
    CK(cudaStreamCreateWithPriority(&m->s_front, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&m->s_lstm, cudaStreamNonBlocking, hi));   // the LSTM's CTA pairs win the SMs when both are pending
    CK(cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming));
    for (auto& ln : m->lane) {
        CK(cudaEventCreateWithFlags(&ln.ev_front, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ln.ev_lstm, cudaEventDisableTiming));
    }
    return FSN_OK;
}

// one pipelined batch: inputs are ready once `ready` (an event, may be null) has fired; returns the lane used
static int submit_impl(fsn_model* m, cudaEvent_t ready, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                       float* d_out, float* d_enh, int* lane_out) {
    int rc = ensure_pipeline(m);
    if (rc) return rc;
    // overlap needs the opt-in (see env_front_overlap) and FullSubNet+ (fullsubnet.Model's full-band LSTM is a cooperative launch and
    // shares the LSTM scratch): otherwise one lane, one stream -- the pipeline still overlaps copies and the caller's post-processing
    const bool overlap = m->env_front_overlap && (m->cfg.model_kind == FSN_KIND_PLUS);
    const int slot = overlap ? (int)(m->nsub & 1) : 0;
    fsn_model::Lane& ln = m->lane[slot];
    cudaStream_t sf = overlap ? m->s_front : m->s_lstm;
    if (ready) CK(cudaStreamWaitEvent(sf, ready, 0));
    if (m->plain_pending) { CK(cudaStreamWaitEvent(sf, m->ev_plain, 0)); CK(cudaStreamWaitEvent(m->s_lstm, m->ev_plain, 0)); }
    if (ln.used) CK(cudaStreamWaitEvent(sf, ln.ev_lstm, 0));              // the LSTM that read this lane has finished
    rc = forward_impl(m, ln, d_mag, d_real, d_imag, B, T, d_out, d_enh, sf, m->s_lstm);
    if (rc) return rc;
    CK(cudaEventRecord(ln.ev_lstm, m->s_lstm));
    ln.used = true;
    m->nsub++;
    if (lane_out) *lane_out = slot;
    return FSN_OK;
}

// Workspace per sample (bytes) of a forward with T frames: the per-sample slices of every grow-only buffer of ensure_ws.  The
// layer-wise path dominates: its time-batched input projection Gin takes 257 x T' x 4H x 2 bytes per sample (0.2 GB at config #5).
static double ws_bytes_per_sample(const fsn_model* m, int T) {
    const fsn_config& c = m->cfg;
    const double F = c.num_freqs, Tp = T + c.look_ahead, rows = F;              // sequences per sample
    double b = 0;
    if (c.model_kind == FSN_KIND_PLUS) b += 3 * F * Tp * 4 * 2 + 3 * Tp * (5.0 * m->Cp + 512) * 4;       // fb_in/out, X0/Xa/Xb/Xr/fbo (fp32), Y1/Y2 (fp16)
    else b += F * Tp * 4 * 4 + Tp * (F + c.fb_hidden) * 4;
    b += rows * Tp * 128;                                                       // packed sub-band images (when used)
    b += rows * c.num_layers * c.sb_hidden * 4;                                  // cell state
    if (use_layerwise(m)) b += rows * Tp * (4.0 * c.sb_hidden + c.sb_hidden) * 2;   // Gin + layer output sequence
    return b;
}

static int forward_plain(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                         float* d_out, float* d_enh, void* stream) {
    if (!m) return fail(FSN_EINVAL, "null argument");
    if (!m->finalized) return fail(FSN_ESTATE, "fsn_model_finalize has not been called");
    if (B < 1 || T < 1) return fail(FSN_EINVAL, "bad batch/frames");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (auto& ln : m->lane)                                             // batches still in flight in the pipeline share the scratch buffers
        if (ln.used) CK(cudaStreamWaitEvent(s, ln.ev_lstm, 0));
    // Samples are independent, so a batch whose workspace would exceed the cap (FSN_WS_CAP_GB, default 48 of the 180 GB) is run as
    // equal sub-batches one after the other on the same stream -- same results, bounded memory (e.g. 256 clips on the layer-wise path).
    const double per = ws_bytes_per_sample(m, T);
    int Bmax = (int)(m->ws_cap_bytes / (per > 1 ? per : 1));
    if (Bmax < 1) Bmax = 1;
    int rc = FSN_OK;
    if (B <= Bmax) {
        rc = forward_impl(m, m->lane[0], d_mag, d_real, d_imag, B, T, d_out, d_enh, s, s);
    } else {
        const int nsplit = (B + Bmax - 1) / Bmax, Bs = (B + nsplit - 1) / nsplit;
        const size_t in_stride = (size_t)m->cfg.num_freqs * T, out_stride = (size_t)m->cfg.output_size * m->cfg.num_freqs * T;
        int64_t launches = 0;
        for (int b0 = 0; b0 < B && rc == FSN_OK; b0 += Bs) {
            const int nb = (B - b0 < Bs) ? B - b0 : Bs;
            rc = forward_impl(m, m->lane[0], d_mag + b0 * in_stride, d_real ? d_real + b0 * in_stride : nullptr, d_imag ? d_imag + b0 * in_stride : nullptr,
                              nb, T, d_out ? d_out + b0 * out_stride : nullptr, d_enh ? d_enh + b0 * in_stride * 2 : nullptr, s, s);
            launches += m->launches;
        }
        m->launches = launches;
    }
    if (rc) return rc;
    if (!m->ev_plain) CK(cudaEventCreateWithFlags(&m->ev_plain, cudaEventDisableTiming));
    CK(cudaEventRecord(m->ev_plain, s));
    m->plain_pending = true;
    return FSN_OK;
}

static int submit_plain(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                        float* d_out, float* d_enh, void* stream) {
    if (!m) return fail(FSN_EINVAL, "null argument");
    int rc = ensure_pipeline(m);
    if (rc) return rc;
    CK(cudaEventRecord(m->ev_in, static_cast<cudaStream_t>(stream)));    // everything enqueued on the caller's stream so far (the inputs)
    return submit_impl(m, m->ev_in, d_mag, d_real, d_imag, B, T, d_out, d_enh, nullptr);
}

extern "C" int fsn_model_forward(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                                 float* d_out, void* stream) {
    if (!d_out) return fail(FSN_EINVAL, "null argument");
    return forward_plain(m, d_mag, d_real, d_imag, B, T, d_out, nullptr, stream);
}
extern "C" int fsn_model_forward_enhance(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                                         float* d_enh, void* stream) {
    if (!d_enh) return fail(FSN_EINVAL, "null argument");
    return forward_plain(m, d_mag, d_real, d_imag, B, T, nullptr, d_enh, stream);
}
extern "C" int fsn_model_submit(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                                float* d_out, void* stream) {
    if (!d_out) return fail(FSN_EINVAL, "null argument");
    return submit_plain(m, d_mag, d_real, d_imag, B, T, d_out, nullptr, stream);
}
extern "C" int fsn_model_submit_enhance(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                                        float* d_enh, void* stream) {
    if (!d_enh) return fail(FSN_EINVAL, "null argument");
    return submit_plain(m, d_mag, d_real, d_imag, B, T, nullptr, d_enh, stream);
}

extern "C" int fsn_model_last_lane(const fsn_model* m) { return m ? m->last_lane : 0; }

extern "C" int fsn_model_wait_lane(fsn_model* m, int32_t lane, void* stream) {
    if (!m || lane < 0 || lane > 1) return fail(FSN_EINVAL, "bad argument");
    if (m->lane[lane].used) CK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), m->lane[lane].ev_lstm, 0));
    return FSN_OK;
}

extern "C" int fsn_model_wait(fsn_model* m, void* stream) {
    if (!m) return fail(FSN_EINVAL, "null model");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (auto& ln : m->lane)
        if (ln.used) CK(cudaStreamWaitEvent(s, ln.ev_lstm, 0));
    return FSN_OK;
}

extern "C" int fsn_model_forward_host(fsn_model* m, const float* h_mag, const float* h_real, const float* h_imag, int32_t B, int32_t T,
                                      float* h_out, void* stream) {
    if (!m || !h_mag || !h_out) return fail(FSN_EINVAL, "null argument");
    const fsn_config& c = m->cfg;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t in_bytes = (size_t)B * c.num_freqs * T * 4, out_bytes = (size_t)B * c.output_size * c.num_freqs * T * 4;
    const float* hin[3] = {h_mag, h_real, h_imag};
    const int nin = (c.model_kind == FSN_KIND_PLUS) ? 3 : 1;
    for (int i = 0; i < nin; ++i) {
        if (!hin[i]) return fail(FSN_EINVAL, "missing input %d", i);
        if (m->stage_in[i].ensure(in_bytes, false)) return fail(FSN_ECUDA, "staging allocation failed");
        CK(cudaMemcpyAsync(m->stage_in[i].p, hin[i], in_bytes, cudaMemcpyHostToDevice, s));
    }
    if (m->stage_out.ensure(out_bytes, false)) return fail(FSN_ECUDA, "staging allocation failed");
    int rc = fsn_model_forward(m, static_cast<const float*>(m->stage_in[0].p), static_cast<const float*>(m->stage_in[1].p),
                               static_cast<const float*>(m->stage_in[2].p), B, T, static_cast<float*>(m->stage_out.p), stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_out, m->stage_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return FSN_OK;
}

extern "C" int fsn_model_forward_host_async(fsn_model* m, const float* h_mag, const float* h_real, const float* h_imag, int32_t B, int32_t T,
                                            float* h_out, void* stream) {
    if (!m || !h_mag || !h_out) return fail(FSN_EINVAL, "null argument");
    (void)stream;                                                        // host buffers carry no stream order; the work runs on internal streams
    const fsn_config& c = m->cfg;
    int rc = ensure_pipeline(m);
    if (rc) return rc;
    if (!m->s_in) {
        CK(cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&m->ev_d2h[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
        }
    }
    const int slot = (int)(m->nsub & 1);                                 // staging slot (the workspace lane is chosen by submit_impl)
    const size_t in_bytes = (size_t)B * c.num_freqs * T * 4, out_bytes = (size_t)B * c.output_size * c.num_freqs * T * 4;
    const float* hin[3] = {h_mag, h_real, h_imag};
    const int nin = (c.model_kind == FSN_KIND_PLUS) ? 3 : 1;
    // the forward that read this input slot (front end: mag / real / imag; fused epilogue: real / imag) has finished
    if (m->d2h_used[slot]) CK(cudaStreamWaitEvent(m->s_in, m->ev_done[slot], 0));
    for (int i = 0; i < nin; ++i) {
        if (!hin[i]) return fail(FSN_EINVAL, "missing input %d", i);
        if (m->a_in[slot][i].bytes < in_bytes) {
            CK(cudaDeviceSynchronize());                                       // (re)allocation only on the first calls / size changes
            if (m->a_in[slot][i].ensure(in_bytes, false)) return fail(FSN_ECUDA, "staging allocation failed");
        }
        CK(cudaMemcpyAsync(m->a_in[slot][i].p, hin[i], in_bytes, cudaMemcpyHostToDevice, m->s_in));
    }
    CK(cudaEventRecord(m->ev_h2d[slot], m->s_in));
    if (m->a_out[slot].bytes < out_bytes) {
        CK(cudaDeviceSynchronize());
        if (m->a_out[slot].ensure(out_bytes, false)) return fail(FSN_ECUDA, "staging allocation failed");
    }
    if (m->d2h_used[slot]) CK(cudaStreamWaitEvent(m->s_lstm, m->ev_d2h[slot], 0));   // the copy-out that read this output slot is done
    int used = 0;
    rc = submit_impl(m, m->ev_h2d[slot], static_cast<const float*>(m->a_in[slot][0].p), static_cast<const float*>(m->a_in[slot][1].p),
                     static_cast<const float*>(m->a_in[slot][2].p), B, T, static_cast<float*>(m->a_out[slot].p), nullptr, &used);
    if (rc) return rc;
    CK(cudaEventRecord(m->ev_done[slot], m->s_lstm));
    CK(cudaStreamWaitEvent(m->s_out, m->ev_done[slot], 0));
    CK(cudaMemcpyAsync(h_out, m->a_out[slot].p, out_bytes, cudaMemcpyDeviceToHost, m->s_out));
    CK(cudaEventRecord(m->ev_d2h[slot], m->s_out));
    m->d2h_used[slot] = true;
    return FSN_OK;
}

extern "C" int fsn_model_sync_host(fsn_model* m) {
    if (!m) return fail(FSN_EINVAL, "null model");
    if (m->s_in) { CK(cudaStreamSynchronize(m->s_in)); }
    if (m->s_lstm) { CK(cudaStreamSynchronize(m->s_front)); CK(cudaStreamSynchronize(m->s_lstm)); }
    if (m->s_out) { CK(cudaStreamSynchronize(m->s_out)); }
    return FSN_OK;
}

extern "C" int fsn_model_get_stage(fsn_model* m, const char* name, float* d_dst, int64_t numel, void* stream) {
    if (!m || !name || !d_dst) return fail(FSN_EINVAL, "null argument");
    fsn_model::Lane& ln = m->lane[m->last_lane];
    if (!ln.wsB) return fail(FSN_ESTATE, "no forward has run yet");
    const fsn_config& c = m->cfg;
    const int F = c.num_freqs, Tp = ln.wsT + c.look_ahead, Pp = (Tp + 3) & ~3, B = ln.wsB;
    const int nbr = (c.model_kind == FSN_KIND_PLUS) ? 3 : 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const std::string n(name);
    if (n == "fb_in" || n == "fb_out") {
        if (numel != (int64_t)nbr * B * F * Tp) return fail(FSN_EINVAL, "%s needs %lld elements", name, (long long)nbr * B * F * Tp);
        if (n == "fb_out" && ln.fbo_valid) {                            // fused-unfold path: the outputs exist time-major only
            launch_tm_to_fm(static_cast<const float*>(ln.fbo.p), d_dst, nbr * B, F, Tp, m->Cp, s);
            CK(cudaGetLastError());
            return FSN_OK;
        }
        const void* src = (n == "fb_in") ? ln.fbin.p : ln.fbout.p;
        CK(cudaMemcpy2DAsync(d_dst, (size_t)Tp * 4, src, (size_t)Pp * 4, (size_t)Tp * 4, (size_t)nbr * B * F, cudaMemcpyDeviceToDevice, s));
        return FSN_OK;
    }
    if (n == "sb_mu") {
        if (numel != B) return fail(FSN_EINVAL, "sb_mu needs %d elements", B);
        CK(cudaMemcpyAsync(d_dst, ln.mu.p, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
        return FSN_OK;
    }
    return fail(FSN_EINVAL, "unknown stage %s", name);
}


// ---------------------------------------------------------------------------------------------
// streaming step API (causal fullsubnet.Model)
// ---------------------------------------------------------------------------------------------
struct fsn_stream {
    fsn_model* m;
    int B, n = 0;
    DevBuf cum_in, cum_sb, fbin, fbx, hseq, fbout, ximg, c_fb, c_sb, h_fb, h_sb, ws_h, ws_c, ws_bar;
    bool use_ws = false;
    int ra_fb = 0, ra_sb = 0, rows_pad = 0, Ipad = 0;
};

extern "C" int fsn_stream_create(fsn_model* m, int32_t B, fsn_stream** out) {
    if (!m || !out || B < 1) return fail(FSN_EINVAL, "bad argument");
    if (!m->finalized) return fail(FSN_ESTATE, "fsn_model_finalize has not been called");
    const fsn_config& c = m->cfg;
    if (c.model_kind != FSN_KIND_FSN) return fail(FSN_EINVAL, "streaming needs fullsubnet.Model (FullSubNet+ is not causal: TSSE pools over all time)");
    if (c.norm_type != FSN_NORM_CUMULATIVE_LAPLACE && c.norm_type != FSN_NORM_CUMULATIVE_LAYER)
        return fail(FSN_EINVAL, "streaming needs a cumulative norm_type");
    fsn_stream* st = new fsn_stream();
    st->m = m; st->B = B;
    const int F = c.num_freqs, rows = B * F, ntiles = (rows + 127) / 128;
    st->Ipad = (F + 15) / 16 * 16; st->rows_pad = (B + 63) / 64 * 64;
    int e = 0;
    e |= st->cum_in.ensure((size_t)B * 2 * sizeof(double), true);
    e |= st->cum_sb.ensure((size_t)rows * 2 * sizeof(double), true);
    e |= st->fbin.ensure((size_t)B * F * 4 * 4, true);
    e |= st->fbx.ensure((size_t)st->rows_pad * st->Ipad * 2, true);
    e |= st->hseq.ensure((size_t)B * c.fb_hidden * 4 * 4, true);
    e |= st->fbout.ensure((size_t)B * F * 4 * 4, true);
    e |= st->ximg.ensure((size_t)ntiles * 16384, true);
    e |= st->c_fb.ensure(lstm_mma_cstate_bytes(c.num_layers, B, c.fb_hidden, &st->ra_fb), true);
    e |= st->c_sb.ensure(lstm_mma_cstate_bytes(c.num_layers, rows, c.sb_hidden, &st->ra_sb), true);
    e |= st->h_fb.ensure((size_t)c.num_layers * st->ra_fb * c.fb_hidden * 2, true);
    e |= st->h_sb.ensure((size_t)c.num_layers * st->ra_sb * c.sb_hidden * 2, true);
    st->use_ws = lstm_ws_supported(c.num_layers, c.fb_hidden, st->Ipad, B, m->num_sms) && !m->env_no_ws;
    if (st->use_ws) {
        e |= st->ws_h.ensure((size_t)c.num_layers * 2 * 64 * c.fb_hidden * 2, true);
        e |= st->ws_c.ensure((size_t)c.num_layers * 64 * c.fb_hidden * 4, true);
        e |= st->ws_bar.ensure(64, true);
    }
    if (e) { delete st; return fail(FSN_ECUDA, "stream state allocation failed"); }
    *out = st;
    return FSN_OK;
}

extern "C" void fsn_stream_destroy(fsn_stream* st) {
    if (!st) return;
    DevBuf* all[] = {&st->cum_in, &st->cum_sb, &st->fbin, &st->fbx, &st->hseq, &st->fbout, &st->ximg, &st->c_fb, &st->c_sb, &st->h_fb, &st->h_sb, &st->ws_h, &st->ws_c, &st->ws_bar};
    for (auto* b : all) b->release();
    delete st;
}

extern "C" int fsn_stream_step(fsn_stream* st, const float* d_mag, float* d_mask, int32_t* h_valid, void* stream) {
    if (!st || !d_mag || !d_mask) return fail(FSN_EINVAL, "null argument");
    fsn_model* m = st->m;
    const fsn_config& c = m->cfg;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int B = st->B, F = c.num_freqs, n = st->n, resume = n > 0;
    StreamNormLaunch na{d_mag, static_cast<float*>(st->fbin.p), static_cast<double*>(st->cum_in.p), B, F, 4, n, c.norm_type};
    launch_stream_norm(na, s);
    launch_fb_pack(static_cast<const float*>(st->fbin.p), static_cast<__half*>(st->fbx.p), B, F, 1, 4, st->rows_pad, st->Ipad, s);
    int e = 0;
    if (st->use_ws) {
        LstmWsLaunch w{};
        fill_ws(m, w);
        w.rows = B; w.rows_pad = st->rows_pad; w.Tp = 1;
        w.x = static_cast<const __half*>(st->fbx.p); w.hbuf = static_cast<__half*>(st->ws_h.p); w.cbuf = static_cast<float*>(st->ws_c.p);
        w.barrier = static_cast<unsigned int*>(st->ws_bar.p);
        w.hseq = static_cast<float*>(st->hseq.p); w.P = 4; w.resume = resume; w.t0 = n;
        e = launch_lstm_ws(w, s);
    } else {
        LstmMmaLaunch a{};
        for (int l = 0; l < c.num_layers; ++l) { a.w.wfrag[l] = static_cast<const uint4*>(m->fb_frag[l].p); a.w.bias[l] = static_cast<const float*>(m->fb_bias[l].p); }
        a.L = c.num_layers; a.H = c.fb_hidden; a.I = F; a.Ipad = st->Ipad; a.rows = B; a.Tp = 1;
        a.xplain = static_cast<const __half*>(st->fbx.p); a.rows_pad = st->rows_pad;
        a.cstate = static_cast<float*>(st->c_fb.p); a.rows_alloc = st->ra_fb;
        a.hseq = static_cast<float*>(st->hseq.p); a.P = 4; a.fast = c.fast_math; a.gru = c.rnn_type == FSN_RNN_GRU;
        a.hstate = static_cast<__half*>(st->h_fb.p); a.resume = resume;
        e = launch_lstm_mma(a, s);
    }
    if (e) return fail(FSN_ECUDA, "full-band LSTM step failed: %s", cudaGetErrorString((cudaError_t)e));
    ConvLaunch cf{};
    cf.X = static_cast<const float*>(st->hseq.p); cf.Y = static_cast<float*>(st->fbout.p);
    cf.Z = B; cf.M = F; cf.K = c.fb_hidden; cf.Tp = 1; cf.P = 4; cf.act = c.fb_act;
    cf.W = P(m, "fb_model.fc_output_layer.weight"); cf.bias = P(m, "fb_model.fc_output_layer.bias");
    launch_conv1x1(cf, s);
    StreamPackLaunch pa{d_mag, static_cast<const float*>(st->fbout.p), 4, static_cast<double*>(st->cum_sb.p), static_cast<__half*>(st->ximg.p),
                        B, F, c.sb_num_neighbors, c.fb_num_neighbors, n, c.norm_type};
    launch_stream_pack(pa, s);
    LstmMmaLaunch b{};
    for (int l = 0; l < c.num_layers; ++l) { b.w.wfrag[l] = static_cast<const uint4*>(m->sb_frag[l].p); b.w.bias[l] = static_cast<const float*>(m->sb_bias[l].p); }
    b.w.fc_w = P(m, "sb_model.fc_output_layer.weight"); b.w.fc_b = P(m, "sb_model.fc_output_layer.bias");
    b.L = c.num_layers; b.H = c.sb_hidden; b.I = m->Isb; b.Ipad = 64; b.rows = B * F; b.Tp = 1;
    b.img = static_cast<const __half*>(st->ximg.p); b.ntiles = (B * F + 127) / 128;
    b.cstate = static_cast<float*>(st->c_sb.p); b.rows_alloc = st->ra_sb;
    b.out = d_mask; b.O = c.output_size; b.F = F; b.la = 0; b.act = c.sb_act; b.fast = c.fast_math; b.gru = c.rnn_type == FSN_RNN_GRU;
    b.hstate = static_cast<__half*>(st->h_sb.p); b.resume = resume;
    e = launch_lstm_mma(b, s);
    if (e) return fail(FSN_ECUDA, "sub-band LSTM step failed: %s", cudaGetErrorString((cudaError_t)e));
    CK(cudaGetLastError());
    if (h_valid) *h_valid = (n >= c.look_ahead) ? 1 : 0;
    st->n++;
    return FSN_OK;
}

extern "C" int fsn_apply_cirm(const float* d_crm, const float* d_noisy, float* d_enh, int32_t B, int32_t F, int32_t T, void* stream) {
    if (!d_crm || !d_noisy || !d_enh || B < 1 || F < 1 || T < 1) return fail(FSN_EINVAL, "bad argument");
    launch_apply_cirm(d_crm, reinterpret_cast<const float2*>(d_noisy), reinterpret_cast<float2*>(d_enh), B, F, T, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return FSN_OK;
}

extern "C" int64_t fsn_model_last_launch_count(const fsn_model* m) { return m ? m->launches : 0; }
extern "C" int fsn_model_lstm_ms_history(fsn_model* m, float* h_ms, int32_t n) {
    if (!m || !h_ms || n < 1) return 0;
    int64_t avail = m->nfwd < fsn_model::NEV ? m->nfwd : fsn_model::NEV;
    if (n > avail) n = (int32_t)avail;
    for (int i = 0; i < n; ++i) {
        const int evi = (int)((m->nfwd - n + i) % fsn_model::NEV);
        float ms = -1.f;
        if (cudaEventSynchronize(m->ev1[evi]) != cudaSuccess || cudaEventElapsedTime(&ms, m->ev0[evi], m->ev1[evi]) != cudaSuccess) ms = -1.f;
        h_ms[i] = ms;
    }
    return n;
}
// [front start, front end, LSTM start, LSTM end] in ms relative to the front start of the oldest reported forward
extern "C" int fsn_model_timeline(fsn_model* m, float* h_ms4, int32_t n) {
    if (!m || !h_ms4 || n < 1) return 0;
    int64_t avail = m->nfwd < fsn_model::NEV ? m->nfwd : fsn_model::NEV;
    if (n > avail) n = (int32_t)avail;
    if (n < 1) return 0;
    const int base = (int)((m->nfwd - n) % fsn_model::NEV);
    for (int i = 0; i < n; ++i) {
        const int evi = (int)((m->nfwd - n + i) % fsn_model::NEV);
        cudaEvent_t ev[4] = {m->evf0[evi], m->evf1[evi], m->ev0[evi], m->ev1[evi]};
        for (int k = 0; k < 4; ++k) {
            float ms = -1.f;
            if (cudaEventSynchronize(ev[k]) != cudaSuccess || cudaEventElapsedTime(&ms, m->evf0[base], ev[k]) != cudaSuccess) ms = -1.f;
            h_ms4[4 * i + k] = ms;
        }
    }
    return n;
}
extern "C" float fsn_model_last_lstm_ms(fsn_model* m) {
    float ms = -1.f;
    return fsn_model_lstm_ms_history(m, &ms, 1) == 1 ? ms : -1.f;
}
extern "C" int fsn_model_last_lstm_impl(const fsn_model* m) { return m ? m->last_impl : 0; }
