// Persistent 2-layer LSTM + Linear for the 257*B independent sub-band sequences (K5 + K6 of SURVEY.md 2a)
// on the 5th-generation tensor cores (tcgen05) of sm_100a.
//
// reference: FullSubNet_Plus.forward -> self.sb_model(sb_input) (fullsubnet_plus.py:205-208), i.e.
//   SequenceModel.forward LSTM branch (audio_zen/model/module/sequence_model.py:113-122):
//   nn.LSTM(I, H, 2, batch_first) + nn.Linear(H, 2), output re-laid out to [B, 2, F, T] and the first
//   look_ahead frames dropped.
//
// Design:
//   * one CTA = 128 sequences (the 128 TMEM lanes) for ALL time steps and BOTH layers;
//   * the recurrent operands h0/h1 live in TENSOR MEMORY as packed fp16 (192 columns each for H=384)
//     and are fed to tcgen05.mma as the A operand (A-from-TMEM form) -- they never touch shared or
//     global memory; one 128-column fp32 accumulator occupies the remaining TMEM columns;
//   * the weights (3.6 MB fp16 for H=384) are streamed from L2 every step as pre-swizzled 16 KB
//     K-major tiles through a ring of shared-memory stages filled by 1-D bulk async copies (TMA engine,
//     mbarrier complete_tx), in exactly the order the MMA issuer consumes them;
//   * each layer-step's gate matrix [128, 4H] is produced in chunks of 128 columns (32 hidden units x
//     i,f,g,o; N=128 because a tcgen05.mma of M=128 costs ~93 cycles for every N <= 128, measured by
//     fsn_probe_tcgen05); all 16 epilogue warps drain the accumulator into registers at once (32 lanes x 32
//     columns each) and release it, so the cell update of chunk j overlaps the MMAs of chunk j+1;
//   * fp32 cell state goes through an L2-resident scratch private to the CTA (coalesced float4);
//   * new hidden values are parked (thread-private shared memory) until the layer-step's last MMA has
//     retired, then written back to TMEM with tcgen05.st;
//   * Linear(H, 2) is accumulated from the fp32 hidden values in the layer-1 epilogue and the mask is
//     written directly in the reference's [B, 2, F, T] layout.
#include "fsn_common.cuh"
#include "fsn_kernels.h"
#include "../../include/fsnplus_b200.h"

#include <cstring>
#include <vector>

namespace fsn {

constexpr int TC5_STAGE = 16384;    // 128 gate columns x 64 k x fp16, SWIZZLE_128B
constexpr int TC5_XIMG = 16384;     // 128 rows x 64 k x fp16, SWIZZLE_128B
constexpr int TC5_EPI_WARPS = 16;   // warps 0-15: epilogue (TMEM lane quarter = warp % 4, column group = warp / 4)
constexpr int TC5_THREADS = (TC5_EPI_WARPS + 2) * 32;   // warp 16: bulk-copy producer, warp 17: MMA issuer + TMEM alloc
constexpr int TC5_MAX_SMEM = 227 * 1024;

struct Tc5Plan { int nstage; size_t fixed, total; };
static inline Tc5Plan tc5_plan(int H) {
    Tc5Plan p;
    p.fixed = TC5_XIMG + (size_t)128 * H * 2 /*park*/ + 4 * 128 * 2 * 4 /*fcpart*/ + 32 * 8 /*barriers*/;
    long avail = TC5_MAX_SMEM - 1024 /*alignment slack*/ - (long)p.fixed;
    p.nstage = (int)(avail / TC5_STAGE);
    if (p.nstage > 12) p.nstage = 12;
    p.total = p.fixed + (size_t)p.nstage * TC5_STAGE + 1024;
    return p;
}

bool lstm_tc5_supported(int L, int H, int I, int O) { return L == 2 && H % 64 == 0 && H >= 64 && H <= 384 && I <= 64 && O == 2; }

size_t lstm_tc5_cstate_bytes(int ntiles, int H) { return (size_t)ntiles * 2 * H * 128 * sizeof(float); }

template <bool FAST>
__global__ void __launch_bounds__(TC5_THREADS, 1) lstm_tc5_kernel(LstmTc5Launch a, int nstage) {
    extern __shared__ uint8_t smem_raw[];
    const int H = a.H, NCH = H / 32, KBH = H / 64, hcols = H / 2, Tp = a.Tp;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stages = smem;
    uint8_t* ximg = stages + (size_t)nstage * TC5_STAGE;
    uint8_t* park = ximg + TC5_XIMG;
    float* fcpart = reinterpret_cast<float*>(park + (size_t)128 * H * 2);   // [4][128][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(fcpart + 4 * 128 * 2);
    uint64_t* full = bars;
    uint64_t* empty = full + nstage;
    uint64_t* xfull = empty + nstage;
    uint64_t* xempty = xfull + 1;
    uint64_t* accfull = xempty + 1;
    uint64_t* accempty = accfull + 1;
    uint64_t* hready = accempty + 1;
    uint64_t* layerdone = hready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layerdone + 1);

    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(xfull, 1); mbar_init(xempty, 1);
        const uint32_t narr = a.elect ? TC5_EPI_WARPS : TC5_EPI_WARPS * 32;   // one elected arrive per epilogue warp, or every thread
        mbar_init(accfull, 1); mbar_init(accempty, narr);
        mbar_init(hready, narr);
        mbar_init(layerdone, 1);
        fence_barrier_init();
    }
    if (warp == TC5_EPI_WARPS + 1) tmem_alloc<512>(tmem_slot);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t acc_col = 2 * hcols;
    const int SPS0 = NCH * (1 + KBH), SPS1 = NCH * 2 * KBH;       // weight stages per layer-step

    if (warp == TC5_EPI_WARPS) {
        // ======================= bulk-copy producer =======================================
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wstream);
            const uint8_t* xsrc = reinterpret_cast<const uint8_t*>(a.img) + (size_t)tile * Tp * TC5_XIMG;
            int slot = 0; uint32_t ph = 0;
            mbar_arrive_expect_tx(xfull, TC5_XIMG);
            bulk_g2s(ximg, xsrc, TC5_XIMG, xfull);
            const int xpoint = SPS0 + SPS1 / 2;                   // x_{t+1} is fetched half way through layer 1 of step t
            for (int t = 0; t < Tp; ++t) {
                for (int s = 0; s < SPS0 + SPS1; ++s) {
                    if (s == xpoint && t + 1 < Tp) {
                        mbar_wait(xempty, t & 1);                 // layer 0 of step t has consumed the single x buffer
                        mbar_arrive_expect_tx(xfull, TC5_XIMG);
                        bulk_g2s(ximg, xsrc + (size_t)(t + 1) * TC5_XIMG, TC5_XIMG, xfull);
                    }
                    mbar_wait(&empty[slot], ph ^ 1);
                    mbar_arrive_expect_tx(&full[slot], TC5_STAGE);
                    bulk_g2s(stages + (size_t)slot * TC5_STAGE, wsrc + (size_t)s * TC5_STAGE, TC5_STAGE, &full[slot]);
                    if (++slot == nstage) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == TC5_EPI_WARPS + 1) {
        // ======================= MMA issuer ================================================
        // The whole warp runs the loop (uniform control flow, operands stay in uniform registers); only the
        // tcgen05 instructions are predicated on one elected lane.
        {
            const uint32_t idesc = umma_idesc_f16(128, 128);
            const uint32_t d = tmem + acc_col;
            int slot = 0; uint32_t ph = 0, accuse = 0, ls = 0;
            for (int t = 0; t < Tp; ++t) {
                for (int layer = 0; layer < 2; ++layer, ++ls) {
                    mbar_wait(hready, ls & 1);                     // h operands of this layer-step are in TMEM
                    uint64_t xdesc = 0;
                    if (layer == 0) {
                        mbar_wait(xfull, t & 1);
                        xdesc = umma_desc_sw128(smem_u32(ximg));
                    }
                    tc5_fence_after();
                    const int nkb = (layer == 0) ? 1 + KBH : 2 * KBH;
                    for (int j = 0; j < NCH; ++j) {
                        if (!(a.debug & 2)) mbar_wait(accempty, (accuse & 1) ^ 1);   // every epilogue warp has drained the accumulator
                        ++accuse;
                        tc5_fence_after();
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(&full[slot], ph);
                            tc5_fence_after();
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(stages + (size_t)slot * TC5_STAGE));
                            const bool xblock = (layer == 0 && kb == 0);
                            const uint32_t acol = (layer == 0) ? (kb - 1) * 32 : (kb < KBH ? kb * 32 : hcols + (kb - KBH) * 32);
                            if (elect_one()) {
                                if (xblock) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) umma_ss(d, xdesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
                                } else {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        umma_ts(d, tmem + acol + kk * 8, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                                }
                                umma_commit(&empty[slot]);         // stage reusable once these MMAs retire
                                if (kb == nkb - 1) {
                                    umma_commit(accfull);
                                    if (j == NCH - 1) {
                                        if (layer == 0) umma_commit(xempty);
                                        umma_commit(layerdone);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++slot == nstage) { slot = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else {
        // ======================= epilogue warps ===========================================
        // Every warp works on every chunk: warp (quarter q, column group cg) owns TMEM lanes [32 q, 32 q + 32) and
        // accumulator columns [32 cg, 32 cg + 32) = gates i,f,g,o of hidden units 32 j + 8 cg + [0, 8).
        const int cg = warp >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;                               // sequence (row) inside the tile
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        {
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = cg; c < 2 * NCH * 2; c += 4) tmem_st8(tl + c * 8, z);   // 2 layers x hcols columns = 4 NCH groups of 8
            tmem_wait_st();
            tc5_fence_before();
            if (a.elect) __syncwarp();
            if (!a.elect || lane == 0) mbar_arrive(hready);
        }
        uint32_t accn = 0, ls = 0;
        float4 cnext[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // t = 0: zero cell state
        float* cbase = a.cstate + (size_t)tile * 2 * H * 128;
        uint8_t* mypark = park + ((size_t)cg * NCH * 128 + r) * 16;
        const int grow = tile * 128 + r;
        const int ob = grow / a.F, of = grow % a.F;
        const int Tout = Tp - a.la;
        const float fcb0 = __ldg(a.fc_b), fcb1 = __ldg(a.fc_b + 1);

        for (int t = 0; t < Tp; ++t) {
            for (int layer = 0; layer < 2; ++layer, ++ls) {
                float fc0 = 0.f, fc1 = 0.f;
                for (int j = 0; j < NCH; ++j) {
                    float4* cp = reinterpret_cast<float4*>(cbase + ((size_t)((layer * NCH + j) * 4 + cg) * 2) * 128 * 4) + r;
                    const float4 c4[2] = {cnext[0], cnext[1]};     // prefetched during the previous chunk
                    const float4* bj = reinterpret_cast<const float4*>(a.bias + (size_t)(layer * NCH + j) * 128 + cg * 32);
                    mbar_wait(accfull, accn & 1);
                    ++accn;
                    tc5_fence_after();
                    uint32_t v[2][16];
                    tmem_ld16(tl + acc_col + cg * 32, v[0]);               // i(8) f(8)
                    tmem_ld16(tl + acc_col + cg * 32 + 16, v[1]);          // g(8) o(8)
                    tmem_wait_ld();
                    tc5_fence_before();
                    if (a.elect) __syncwarp();
                    if (!a.elect || lane == 0) mbar_arrive(accempty);
                    {   // cell state of the NEXT chunk in program order: (layer, j+1), else chunk 0 of the other layer (next step after layer 1)
                        const int nj = (j + 1 < NCH) ? j + 1 : 0;
                        const int nl = (j + 1 < NCH) ? layer : (layer ^ 1);
                        const int nt = (j + 1 < NCH || layer == 0) ? t : t + 1;
                        const float4* np = reinterpret_cast<const float4*>(cbase + ((size_t)((nl * NCH + nj) * 4 + cg) * 2) * 128 * 4) + r;
                        if (nt == 0 || nt >= Tp) { cnext[0] = make_float4(0.f, 0.f, 0.f, 0.f); cnext[1] = cnext[0]; }
                        else { cnext[0] = np[0]; cnext[1] = np[128]; }
                    }

                    if (a.debug & 1) continue;                     // timing experiment: drain only, no cell update
                    const float L2E = 1.4426950408889634f;
                    uint32_t hp[4];
                    float cn[8];
#pragma unroll
                    for (int u4 = 0; u4 < 2; ++u4) {
                        const float4 bi = __ldg(bj + u4), bf = __ldg(bj + 2 + u4), bg = __ldg(bj + 4 + u4), bo = __ldg(bj + 6 + u4);
                        const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, bfa[4] = {bf.x, bf.y, bf.z, bf.w};
                        const float bga[4] = {bg.x, bg.y, bg.z, bg.w}, boa[4] = {bo.x, bo.y, bo.z, bo.w};
                        const float cpv[4] = {c4[u4].x, c4[u4].y, c4[u4].z, c4[u4].w};
                        float hv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int u = u4 * 4 + e;
                            lstm_cell<FAST>(fmaf(__uint_as_float(v[0][u]), -L2E, bia[e]), fmaf(__uint_as_float(v[0][8 + u]), -L2E, bfa[e]),
                                            fmaf(__uint_as_float(v[1][u]), -2.f * L2E, bga[e]), fmaf(__uint_as_float(v[1][8 + u]), -L2E, boa[e]),
                                            cpv[e], cn[u], hv[e]);
                        }
                        if (layer == 1) {
                            const float4 wa = __ldg(reinterpret_cast<const float4*>(a.fc_w + j * 32 + cg * 8) + u4);
                            const float4 wb = __ldg(reinterpret_cast<const float4*>(a.fc_w + H + j * 32 + cg * 8) + u4);
                            fc0 = fmaf(hv[0], wa.x, fmaf(hv[1], wa.y, fmaf(hv[2], wa.z, fmaf(hv[3], wa.w, fc0))));
                            fc1 = fmaf(hv[0], wb.x, fmaf(hv[1], wb.y, fmaf(hv[2], wb.z, fmaf(hv[3], wb.w, fc1))));
                        }
                        hp[2 * u4] = pack_half2(hv[0], hv[1]);
                        hp[2 * u4 + 1] = pack_half2(hv[2], hv[3]);
                    }
                    cp[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    cp[128] = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    *reinterpret_cast<uint4*>(mypark + (size_t)j * 128 * 16) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                }
                // every MMA of this layer-step has retired -> h_{t-1} may be overwritten in TMEM
                mbar_wait(layerdone, ls & 1);
                tc5_fence_after();
                for (int j = 0; j < NCH; ++j) {
                    const uint4 p0 = *reinterpret_cast<const uint4*>(mypark + (size_t)j * 128 * 16);
                    const uint32_t hv[4] = {p0.x, p0.y, p0.z, p0.w};
                    tmem_st4(tl + layer * hcols + j * 16 + cg * 4, hv);
                }
                tmem_wait_st();
                tc5_fence_before();
                if (a.elect) __syncwarp();
                if (!a.elect || lane == 0) mbar_arrive(hready);

                if (layer == 1) {
                    if (cg != 0) { fcpart[(cg * 128 + r) * 2] = fc0; fcpart[(cg * 128 + r) * 2 + 1] = fc1; }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (cg == 0 && t >= a.la && grow < a.rows) {
                        const float o0 = fc0 + fcpart[(128 + r) * 2] + fcpart[(256 + r) * 2] + fcpart[(384 + r) * 2] + fcb0;
                        const float o1 = fc1 + fcpart[(128 + r) * 2 + 1] + fcpart[(256 + r) * 2 + 1] + fcpart[(384 + r) * 2 + 1] + fcb1;
                        a.out[(((size_t)ob * 2 + 0) * a.F + of) * Tout + (t - a.la)] = o0;
                        a.out[(((size_t)ob * 2 + 1) * a.F + of) * Tout + (t - a.la)] = o1;
                    }
                    asm volatile("bar.sync 2, 512;" ::: "memory");    // fcpart is single-buffered
                }
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    if (warp == TC5_EPI_WARPS + 1) tmem_dealloc<512>(tmem);
}

int launch_lstm_tc5(const LstmTc5Launch& a, cudaStream_t s) {
    if (!lstm_tc5_supported(2, a.H, a.I, 2)) return (int)cudaErrorInvalidValue;
    Tc5Plan p = tc5_plan(a.H);
    if (a.nstage_cap > 0 && a.nstage_cap < p.nstage) { p.total -= (size_t)(p.nstage - a.nstage_cap) * TC5_STAGE; p.nstage = a.nstage_cap; }
    if (p.nstage < 2) return (int)cudaErrorInvalidValue;
    cudaError_t e;
    if (a.fast) {
        e = cudaFuncSetAttribute(lstm_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.total);
        if (e != cudaSuccess) return (int)e;
        lstm_tc5_kernel<true><<<a.ntiles, TC5_THREADS, p.total, s>>>(a, p.nstage);
    } else {
        e = cudaFuncSetAttribute(lstm_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.total);
        if (e != cudaSuccess) return (int)e;
        lstm_tc5_kernel<false><<<a.ntiles, TC5_THREADS, p.total, s>>>(a, p.nstage);
    }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Host-side packing of the weight stream (exposed through the C ABI for CPU layout tests).
// Stream order per time step: layer 0, chunk j = 0..H/32-1: [x block][H/64 hidden blocks];
//                             layer 1, chunk j:              [H/64 blocks of W_ih1][H/64 blocks of W_hh1].
// Stage = 128 gate columns (fsn_tc5_gate_row) x 64 k, K-major,
// SWIZZLE_128B.
// ---------------------------------------------------------------------------------------------
static inline uint16_t f2h_bits(float f) {
    __half h = __float2half_rn(f);
    uint16_t b;
    std::memcpy(&b, &h, 2);
    return b;
}

}  // namespace fsn

extern "C" int64_t fsn_tc5_weight_stream_bytes(int32_t I, int32_t H) {
    if (H % 64 || I > 64) return -1;
    const int NCH = H / 32, KBH = H / 64;
    return (int64_t)(NCH * (1 + KBH) + NCH * 2 * KBH) * fsn::TC5_STAGE;
}

// Gate column n (0..127) of chunk j <-> weight row: n = cg*32 + q*8 + u, q in (i,f,g,o), hidden unit 32 j + 8 cg + u.
extern "C" int32_t fsn_tc5_gate_row(int32_t H, int32_t j, int32_t n) { return ((n % 32) / 8) * H + 32 * j + 8 * (n / 32) + (n % 8); }

extern "C" int fsn_tc5_pack_weights(int32_t I, int32_t H, const float* w_ih0, const float* w_hh0, const float* w_ih1,
                                    const float* w_hh1, uint16_t* dst) {
    if (H % 64 || I > 64) return FSN_EINVAL;
    const int NCH = H / 32, KBH = H / 64;
    size_t s = 0;
    auto stage = [&](auto&& getw) {
        uint8_t* img = reinterpret_cast<uint8_t*>(dst) + s * fsn::TC5_STAGE;
        for (int n = 0; n < 128; ++n)
            for (int k = 0; k < 64; ++k) {
                uint16_t b = fsn::f2h_bits(getw(n, k));
                std::memcpy(img + fsn::sw128_offset(n, k), &b, 2);
            }
        ++s;
    };
    for (int j = 0; j < NCH; ++j) {
        auto row = [&](int n) { return fsn_tc5_gate_row(H, j, n); };
        stage([&](int n, int k) { return k < I ? w_ih0[(size_t)row(n) * I + k] : 0.f; });
        for (int kb = 0; kb < KBH; ++kb) stage([&](int n, int k) { return w_hh0[(size_t)row(n) * H + kb * 64 + k]; });
    }
    for (int j = 0; j < NCH; ++j) {
        auto row = [&](int n) { return fsn_tc5_gate_row(H, j, n); };
        for (int kb = 0; kb < KBH; ++kb) stage([&](int n, int k) { return w_ih1[(size_t)row(n) * H + kb * 64 + k]; });
        for (int kb = 0; kb < KBH; ++kb) stage([&](int n, int k) { return w_hh1[(size_t)row(n) * H + kb * 64 + k]; });
    }
    return FSN_OK;
}

extern "C" uint32_t fsn_sw128_offset(uint32_t row, uint32_t k) { return fsn::sw128_offset(row, k); }
