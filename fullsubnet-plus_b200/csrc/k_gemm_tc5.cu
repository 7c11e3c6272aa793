// TCN full-band model (K3 of SURVEY.md 2a) on the 5th-generation tensor cores: the 1x1 convolutions of the eight
// TCNBlocks and the output Linear as persistent TF32 tcgen05 GEMMs over time-major activations.
//
// reference: TCNBlock.forward (audio_zen/model/module/causal_conv.py:96-108) and SequenceModel.forward, TCN branch
// (audio_zen/model/module/sequence_model.py:106-112).
//
// Formulation.  Activations are stored time-major, rows = (branch, sample, frame), columns = channels (padded to a
// multiple of 32 floats = one 128-byte swizzle atom), so every 1x1 convolution is  D[rows, C_out] = X[rows, C_in] *
// W[C_out, C_in]^T  with BOTH operands K-major -- the layout PyTorch already stores W in.  Tiles of 128 rows x 32 k
// (A) and N_TILE x 32 k (B) are fetched by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a shared-memory ring,
// multiplied with tcgen05.mma.kind::tf32 into double-buffered TMEM accumulators, and finished by four epilogue warps:
//   EPI_PRELU_STATS : + bias, PReLU, per-sample sum / sum-of-squares for the following gLN, store (rounded to tf32)
//   EPI_GLN_RES     : gLN folded analytically -- conv(W, gLN(y)) = rstd * (W diag(gamma)) y - mean rstd s1 + s2 --
//                     so the GEMM runs on the raw activation and the per-sample affine is applied here, + residual
//   EPI_OUT         : + bias, output activation, written transposed into the [branch, B, F, T'] layout the sub-band
//                     packer reads.
#include <cuda.h>

#include "fsn_common.cuh"
#include "fsn_kernels.h"
#include "../../include/fsnplus_b200.h"

namespace fsn {

constexpr int G5_EPI_WARPS = 8;       // two warps per TMEM lane quarter, each takes half of the 16-column chunks of a tile
constexpr int G5_THREADS = (2 + G5_EPI_WARPS) * 32;   // warp 0 TMA producer, warp 1 MMA issuer + TMEM alloc, warps 2-9 epilogue
constexpr int G5_A_BYTES = 128 * 128;

__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);   // D=f32, A=B=tf32, K-major
}
__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float(f2tf32(x)); }

template <int EPI>
__global__ void __launch_bounds__(G5_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmTc5Launch a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int NT = a.NT, nstage = a.nstage, stage_bytes = G5_A_BYTES + NT * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nstage * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = full + nstage;
    uint64_t* accfull = empty + nstage;
    uint64_t* accempty = accfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], G5_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nkb = a.Kp / 32;
    const int tiles_per_branch = a.tiles_m * a.ntiles_n;
    const int total = a.nbranch * tiles_per_branch;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int br = tile / tiles_per_branch, rem = tile % tiles_per_branch;
                const int mt = rem / a.ntiles_n, nt = rem % a.ntiles_n;
                const int arow = br * a.rows_per_branch + mt * 128, brow = br * a.Npad + nt * NT;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[slot], ph ^ 1);
                    uint8_t* st = smem + (size_t)slot * stage_bytes;
                    mbar_arrive_expect_tx(&full[slot], stage_bytes);
                    tma_load_2d(st, &mapA, kb * 32, arow, &full[slot]);
                    tma_load_2d(st + G5_A_BYTES, &mapB, kb * 32, brow, &full[slot]);
                    if (++slot == nstage) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {   // warp-uniform issue loop, instructions predicated on one elected lane (see elect_one)
            const uint32_t idesc = umma_idesc_tf32(128, NT);
            int slot = 0; uint32_t ph = 0, use[2] = {0, 0}, it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&accempty[buf], (use[buf] & 1) ^ 1);
                ++use[buf];
                tc5_fence_after();
                const uint32_t d = tmem + buf * 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[slot], ph);
                    tc5_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)slot * stage_bytes);
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + G5_A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_ss_tf32(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                        umma_commit(&empty[slot]);
                        if (kb == nkb - 1) umma_commit(&accfull[buf]);
                    }
                    __syncwarp();
                    if (++slot == nstage) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else {
        const int q = warp & 3, r = q * 32 + lane, half = (warp - 2) >> 2;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const int nchunk = NT / 16, cbeg = half ? (nchunk + 1) / 2 : 0, cend = half ? nchunk : (nchunk + 1) / 2;
        uint32_t use[2] = {0, 0}, it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int br = tile / tiles_per_branch, rem = tile % tiles_per_branch;
            const int mt = rem / a.ntiles_n, nt = rem % a.ntiles_n;
            const int rib = mt * 128 + r;                              // row inside the branch
            const bool valid = rib < a.rows_per_branch;
            const int zb = rib / a.Tp, tt = rib % a.Tp;
            const int z = br * a.B + zb;
            const size_t grow = (size_t)br * a.rows_per_branch + rib;
            const int buf = it & 1;
            const int n0 = nt * NT;
            float mean = 0.f, rstd = 1.f;
            if (EPI == EPI5_GLN_RES && valid) {
                const double su = a.stats_in[2 * z], sq = a.stats_in[2 * z + 1];
                const double mu = su / a.count_in, var = sq / a.count_in - mu * mu;
                mean = (float)mu;
                rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + 1e-8));
            }
            const float slope = (EPI == EPI5_PRELU_STATS) ? __ldg(a.prelu[br]) : 0.f;
            const float* bias = a.bias[br];
            const float* s1 = a.s1[br];
            double lsum = 0.0, lsq = 0.0;
            mbar_wait(&accfull[buf], use[buf] & 1);
            ++use[buf];
            tc5_fence_after();
            float4 xpre[4];
            if (EPI == EPI5_GLN_RES && valid && cbeg < cend) {
                const float4* xo = reinterpret_cast<const float4*>(a.Xold + grow * a.ldY + n0 + cbeg * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) xpre[i] = xo[i];
            }
            for (int c = cbeg; c < cend; ++c) {
                uint32_t v[16];
                tmem_ld16(tl + buf * 256 + c * 16, v);
                tmem_wait_ld();
                const int n = n0 + c * 16;
                float y[16];
                if (EPI == EPI5_PRELU_STATS) {
                    float ls = 0.f, lq = 0.f;
                    float bvec[16];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                        bvec[4 * i4] = b4.x; bvec[4 * i4 + 1] = b4.y; bvec[4 * i4 + 2] = b4.z; bvec[4 * i4 + 3] = b4.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float t = __uint_as_float(v[i]) + bvec[i];
                        t = (t >= 0.f) ? t : slope * t;
                        ls += t; lq = fmaf(t, t, lq);
                        y[i] = round_tf32(t);
                    }
                    lsum += (double)ls; lsq += (double)lq;
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(a.Y + grow * a.ldY + n);
#pragma unroll
                        for (int i = 0; i < 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                    }
                } else if (EPI == EPI5_GLN_RES) {
                    if (valid) {
                        float4 xcur[4] = {xpre[0], xpre[1], xpre[2], xpre[3]};
                        if (c + 1 < cend) {                     // residual of the next chunk: in flight during this chunk's math
                            const float4* xn = reinterpret_cast<const float4*>(a.Xold + grow * a.ldY + n + 16);
#pragma unroll
                            for (int i = 0; i < 4; ++i) xpre[i] = xn[i];
                        }
                        float4* dst = reinterpret_cast<float4*>(a.Y + grow * a.ldY + n);
                        float4* dr = a.Xrelu ? reinterpret_cast<float4*>(a.Xrelu + grow * a.ldY + n) : nullptr;
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 xv = xcur[i4];
                            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
                            const float4 s14 = __ldg(reinterpret_cast<const float4*>(s1 + n) + i4), b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                            const float s1v[4] = {s14.x, s14.y, s14.z, s14.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                            float o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = i4 * 4 + e;
                                const float val = fmaf(rstd, __uint_as_float(v[i]), fmaf(-mean * rstd, s1v[e], bv[e]));
                                o[e] = xa[e] + val;
                            }
                            dst[i4] = make_float4(o[0], o[1], o[2], o[3]);
                            if (dr) dr[i4] = make_float4(fmaxf(o[0], 0.f), fmaxf(o[1], 0.f), fmaxf(o[2], 0.f), fmaxf(o[3], 0.f));
                        }
                    }
                } else if (a.out_tm) {
                    // time-major output [(branch, b, t), Npad]: the layout the sub-band LSTM's x-tile builders read (k_lstm_tc5d.cu);
                    // each thread owns a row -> 64 contiguous bytes per chunk (pad columns: zero weights, zero bias)
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(a.out_tm + grow * a.ldY + n);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                            dst[i4] = make_float4(apply_act(__uint_as_float(v[4 * i4]) + b4.x, a.act), apply_act(__uint_as_float(v[4 * i4 + 1]) + b4.y, a.act),
                                                  apply_act(__uint_as_float(v[4 * i4 + 2]) + b4.z, a.act), apply_act(__uint_as_float(v[4 * i4 + 3]) + b4.w, a.act));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (valid && n + i < a.F)
                            a.out[((size_t)z * a.F + n + i) * a.P + tt] = apply_act(__uint_as_float(v[i]) + __ldg(bias + n + i), a.act);
                    }
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
            if (EPI == EPI5_PRELU_STATS) {
                // rows of a warp almost always belong to one sample: reduce in the warp, one atomic pair per warp
                const int key = valid ? z : -1;
                if (__match_any_sync(0xffffffffu, key) == 0xffffffffu) {
                    lsum = warp_sum_d(lsum); lsq = warp_sum_d(lsq);
                    if (lane == 0 && valid) { atomicAdd(&a.stats_out[2 * z], lsum); atomicAdd(&a.stats_out[2 * z + 1], lsq); }
                } else if (valid) {
                    atomicAdd(&a.stats_out[2 * z], lsum);
                    atomicAdd(&a.stats_out[2 * z + 1], lsq);
                }
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_f32_2d(void* out_map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r;
}

int launch_gemm_tc5(const void* mapA, const void* mapB, GemmTc5Launch a, int num_sms, cudaStream_t s) {
    const int stage_bytes = G5_A_BYTES + a.NT * 128;
    a.nstage = (227 * 1024 - 2048) / stage_bytes;
    if (a.nstage > 8) a.nstage = 8;
    if (a.nstage < 2 || a.NT % 16 || a.NT > 256 || a.Kp % 32) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)a.nstage * stage_bytes + 1024 + 256;
    const int total = a.nbranch * a.tiles_m * a.ntiles_n;
    const int grid = total < num_sms ? total : num_sms;
    const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(mapA);
    const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(mapB);
    cudaError_t e;
#define G5_LAUNCH(E)                                                                                              \
    e = cudaFuncSetAttribute(gemm_tc5_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e != cudaSuccess) return (int)e;                                                                          \
    gemm_tc5_kernel<E><<<grid, G5_THREADS, smem, s>>>(mA, mB, a);
    if (a.epi == EPI5_PRELU_STATS) { G5_LAUNCH(EPI5_PRELU_STATS) }
    else if (a.epi == EPI5_GLN_RES) { G5_LAUNCH(EPI5_GLN_RES) }
    else { G5_LAUNCH(EPI5_OUT) }
#undef G5_LAUNCH
    return (int)cudaGetLastError();
}

// gLN1 apply -> depth-wise conv (k=3, dilation d, zero padding in the normalised domain; causal or symmetric taps) -> PReLU2 + gLN2 statistics,
// on the time-major [rows, C] layout.  reference: causal_conv.py:100-106.  One CTA = one sample x 64 channels x all
// frames: the slab (T' x 64 floats, 48 KB for T' = 190) is staged once in shared memory with the gLN already applied,
// so every input element is read exactly once from HBM/L2 and the three taps come from shared memory.
constexpr int DW_CH = 64;
__global__ void __launch_bounds__(256) dwconv_tm_kernel(DwTmLaunch a) {
    extern __shared__ float slab[];                       // [Tp][DW_CH]
    __shared__ double red[16];
    const int z = blockIdx.y, g = z / a.B, c0 = blockIdx.x * DW_CH;
    const int C = a.C, Tp = a.Tp, d = a.dilation;
    const double cnt = (double)C * (double)Tp;
    const double mu = a.stats_in[2 * z] / cnt;
    const double var = a.stats_in[2 * z + 1] / cnt - mu * mu;
    const float mean = (float)mu, rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + 1e-8));
    const float slope = __ldg(a.prelu[g]);
    const size_t base = (size_t)z * Tp * C + c0;
    const int cq = threadIdx.x & 15, tr = threadIdx.x >> 4;           // 16 float4 columns x 16 frame rows per pass
    float4 ga, be, w0, w1, w2, bb;
    {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma[g] + c0) + cq);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta[g] + c0) + cq);
        ga = make_float4(gm.x * rstd, gm.y * rstd, gm.z * rstd, gm.w * rstd);
        be = make_float4(bt.x - mean * ga.x, bt.y - mean * ga.y, bt.z - mean * ga.z, bt.w - mean * ga.w);
        const float* wp = a.w[g] + (size_t)(c0 + cq * 4) * 3;
        w0 = make_float4(__ldg(wp + 0), __ldg(wp + 3), __ldg(wp + 6), __ldg(wp + 9));
        w1 = make_float4(__ldg(wp + 1), __ldg(wp + 4), __ldg(wp + 7), __ldg(wp + 10));
        w2 = make_float4(__ldg(wp + 2), __ldg(wp + 5), __ldg(wp + 8), __ldg(wp + 11));
        bb = __ldg(reinterpret_cast<const float4*>(a.b[g] + c0) + cq);
    }
    for (int t = tr; t < Tp; t += 16) {
        const float4 v = *reinterpret_cast<const float4*>(a.X + base + (size_t)t * C + cq * 4);
        reinterpret_cast<float4*>(slab + t * DW_CH)[cq] =
            make_float4(fmaf(v.x, ga.x, be.x), fmaf(v.y, ga.y, be.y), fmaf(v.z, ga.z, be.z), fmaf(v.w, ga.w, be.w));
    }
    __syncthreads();
    float ls = 0.f, lq = 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = tr; t < Tp; t += 16) {
        // taps (w0, w1, w2): non-causal (t-d, t, t+d); causal (t-2d, t-d, t) -- padding 2d + chomp, causal_conv.py:74-75,104-105
        const int tl = a.causal ? t - 2 * d : t - d, tm = a.causal ? t - d : t, tr2 = a.causal ? t : t + d;
        const float4 m = (tm >= 0) ? reinterpret_cast<const float4*>(slab + tm * DW_CH)[cq] : zero;
        const float4 l = (tl >= 0) ? reinterpret_cast<const float4*>(slab + tl * DW_CH)[cq] : zero;
        const float4 r = (tr2 < Tp) ? reinterpret_cast<const float4*>(slab + tr2 * DW_CH)[cq] : zero;
        float o[4] = {fmaf(w0.x, l.x, fmaf(w1.x, m.x, fmaf(w2.x, r.x, bb.x))), fmaf(w0.y, l.y, fmaf(w1.y, m.y, fmaf(w2.y, r.y, bb.y))),
                      fmaf(w0.z, l.z, fmaf(w1.z, m.z, fmaf(w2.z, r.z, bb.z))), fmaf(w0.w, l.w, fmaf(w1.w, m.w, fmaf(w2.w, r.w, bb.w)))};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = (o[i] >= 0.f) ? o[i] : slope * o[i];
            ls += o[i]; lq = fmaf(o[i], o[i], lq);
            o[i] = round_tf32(o[i]);
        }
        *reinterpret_cast<float4*>(a.Y + base + (size_t)t * C + cq * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    double s1 = warp_sum_d((double)ls), s2 = warp_sum_d((double)lq);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp] = s1; red[8 + warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x1 = 0, x2 = 0;
        for (int i = 0; i < 8; ++i) { x1 += red[i]; x2 += red[8 + i]; }
        atomicAdd(&a.stats_out[2 * z], x1);
        atomicAdd(&a.stats_out[2 * z + 1], x2);
    }
}

void launch_dwconv_tm(const DwTmLaunch& a, cudaStream_t s) {
    const size_t smem = (size_t)a.Tp * DW_CH * sizeof(float);
    cudaFuncSetAttribute(dwconv_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dwconv_tm_kernel<<<dim3(a.C / DW_CH, a.Z), 256, smem, s>>>(a);
}

}  // namespace fsn
