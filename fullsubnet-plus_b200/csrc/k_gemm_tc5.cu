// TCN full-band model (K3 of SURVEY.md 2a) on the 5th-generation tensor cores: the 1x1 convolutions of the eight
// TCNBlocks and the output Linear as persistent tcgen05 GEMMs over time-major activations, plus the depth-wise convolution between them.
//
// reference: TCNBlock.forward (audio_zen/model/module/causal_conv.py:96-108) and SequenceModel.forward, TCN branch
// (audio_zen/model/module/sequence_model.py:106-112).
//
// Formulation.  Activations are stored time-major, rows = (branch, sample, frame), columns = channels (padded to a
// multiple of 32 floats = one 128-byte swizzle atom), so every 1x1 convolution is  D[rows, C_out] = X[rows, C_in] *
// W[C_out, C_in]^T  with BOTH operands K-major -- the layout PyTorch already stores W in.  Tiles of 128 rows x 128 bytes of k
// (A) and N_TILE x 128 bytes of k (B) are fetched by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a shared-memory ring,
// multiplied with tcgen05.mma (kind::tf32 on the fp32 residual stream, kind::f16 on the fp16 hidden activations) into double-
// buffered TMEM accumulators, and finished by eight epilogue warps whose global loads / stores go through per-warp staging tiles
// (g5_store_rows64: 8 rows x 64 contiguous bytes per instruction instead of 32 rows x 16 bytes):
//   EPI_PRELU_STATS : + bias, PReLU, per-sample sum / sum-of-squares for the following gLN, store as fp16 times a per-sample power of
//                     two (fp16_store_scale: the hidden, 512-channel activations of a block live in fp16 -- the same 10 mantissa
//                     bits the tf32 product consumed, half the bytes; the normalised real / imag streams reach 1e6 and beyond)
//   EPI_GLN_RES     : gLN folded analytically -- conv(W, gLN(y)) = rstd * (W diag(gamma)) y - mean rstd s1 + s2 --
//                     so the GEMM runs on the raw activation and the per-sample affine is applied here, + residual (fp32 stream),
//                     + the running max |x| per sample of the new stream (the next block's store scale).
//                     Its operands (hidden activation, folded weights) are fp16: kind::f16, 64-element k-blocks.
//   EPI_OUT         : + bias, output activation, written time-major (the layout the sub-band LSTM's x-tile builders read) or,
//                     for the packed path, transposed into [branch, B, F, T'].
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

// ---- coalesced epilogue I/O --------------------------------------------------------------------------------------------------
// tcgen05.ld hands every epilogue thread ONE ROW of the tile (32x32b shape), so a direct global store touches 32 different 128-byte
// lines per instruction with 16 bytes each: the L1 tag stage (one line per cycle) -- not HBM -- bounded these kernels (ncu, round 2:
// 2.6 M store + 2.8 M load sector accesses in the second convolution = 20 us of its 48).  Each warp therefore owns a 2 KB staging tile
// (32 rows x 64 bytes, 16-byte units XOR-swizzled): rows go in thread-per-row, come out with lane l <-> (row (l >> 2) + 8 i, unit l & 3),
// i.e. 8 rows x 64 contiguous bytes per instruction (both views are bank-conflict free), and the same in reverse for the residual.
constexpr int G5_STG_BYTES = 32 * 64;
__device__ __forceinline__ uint4* g5_stg(uint8_t* t, int row, int unit) {
    return reinterpret_cast<uint4*>(t + row * 64 + ((unit ^ ((row >> 1) & 3)) << 4));
}
// every lane holds the 64 bytes of its row; row r of the warp lives at gbase + r * pitch (bytes); rows >= nvalid are not written
__device__ __forceinline__ void g5_store_rows64(uint8_t* t, int lane, const uint4 (&v)[4], uint8_t* gbase, size_t pitch, int nvalid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *g5_stg(t, lane, q) = v[q];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = (lane >> 2) + 8 * i;
        const uint4 u = *g5_stg(t, row, lane & 3);
        if (row < nvalid) *reinterpret_cast<uint4*>(gbase + (size_t)row * pitch + (lane & 3) * 16) = u;
    }
    __syncwarp();
}
__device__ __forceinline__ void g5_load_rows64_issue(int lane, const uint8_t* gbase, size_t pitch, int nvalid, uint4 (&g)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = (lane >> 2) + 8 * i;
        g[i] = (row < nvalid) ? *reinterpret_cast<const uint4*>(gbase + (size_t)row * pitch + (lane & 3) * 16) : make_uint4(0, 0, 0, 0);
    }
}
__device__ __forceinline__ void g5_load_rows64_commit(uint8_t* t, int lane, const uint4 (&g)[4], uint4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *g5_stg(t, (lane >> 2) + 8 * i, lane & 3) = g[i];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = *g5_stg(t, lane, q);
    __syncwarp();
}

template <int EPI>
__global__ void __launch_bounds__(G5_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmTc5Launch a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int NT = a.NT, nstage = a.nstage, stage_bytes = G5_A_BYTES + NT * 128;
    uint8_t* stg_all = smem + (size_t)nstage * stage_bytes;                  // [G5_EPI_WARPS][G5_STG_BYTES] epilogue staging tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + G5_EPI_WARPS * G5_STG_BYTES);
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
    // chained launch: barriers and tensor memory are set up while the previous kernel of the chain drains; nothing it wrote is read
    // (and nothing it reads is overwritten) before this point
    pdl_trigger();
    pdl_wait();
    constexpr bool AH = (EPI == EPI5_GLN_RES);             // fp16 operands: a 128-byte swizzle atom holds 64 k-elements instead of 32
    constexpr int KBW = AH ? 64 : 32;
    const int nkb = a.Kp / KBW;
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
                    tma_load_2d(st, &mapA, kb * KBW, arow, &full[slot]);
                    tma_load_2d(st + G5_A_BYTES, &mapB, kb * KBW, brow, &full[slot]);
                    if (++slot == nstage) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {   // warp-uniform issue loop, instructions predicated on one elected lane (see elect_one)
            const uint32_t idesc = AH ? umma_idesc_f16(128, NT) : umma_idesc_tf32(128, NT);
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
                        for (int kk = 0; kk < 4; ++kk) {                  // 32 bytes of K per instruction: 8 tf32 or 16 fp16 elements
                            if (AH) umma_ss(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                            else umma_ss_tf32(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                        }
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
            // fp16 store scale of the hidden activation: a power of two from the absolute maximum of this sample's input stream (the
            // normalised real / imaginary branches reach 1e5 and beyond -- 1 / mean(real) is unbounded); undone exactly by the gLN that follows
            const float ysc = (EPI == EPI5_PRELU_STATS && valid) ? fp16_store_scale(__ldg(a.amax_in + z)) : 1.f;
            float rowmax = 0.f;
            const float* bias = a.bias[br];
            const float* s1 = a.s1[br];
            double lsum = 0.0, lsq = 0.0;
            // coalesced I/O of this warp's 32 rows (see g5_store_rows64): row 0 of the warp, rows valid, staging tile
            uint8_t* stg = stg_all + (warp - 2) * G5_STG_BYTES;
            const size_t wrow0 = (size_t)br * a.rows_per_branch + (size_t)mt * 128 + q * 32;
            int nvalid = a.rows_per_branch - (mt * 128 + q * 32);
            nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
            // GLN_RES: the residual of the first two chunks does not depend on the MMAs -- in flight while this warp waits for the accumulator
            // (the loads miss to HBM; with one chunk ahead the 4-5 chunks of a warp were a chain of dependent DRAM round trips)
            uint4 gA[4], gB[4];
            const uint8_t* xbase = (EPI == EPI5_GLN_RES) ? reinterpret_cast<const uint8_t*>(a.Xold + wrow0 * a.ldY + n0) : nullptr;
            if (EPI == EPI5_GLN_RES) {
                if (cbeg < cend) g5_load_rows64_issue(lane, xbase + (size_t)cbeg * 64, (size_t)a.ldY * sizeof(float), nvalid, gA);
                if (cbeg + 1 < cend) g5_load_rows64_issue(lane, xbase + (size_t)(cbeg + 1) * 64, (size_t)a.ldY * sizeof(float), nvalid, gB);
            }
            mbar_wait(&accfull[buf], use[buf] & 1);
            ++use[buf];
            tc5_fence_after();
            if (EPI == EPI5_PRELU_STATS) {
                // two 16-column chunks per pass: 32 fp16 values = 64 bytes per row
                const size_t pitch = (size_t)a.ldY * sizeof(__half);
                for (int c = cbeg; c < cend; c += 2) {
                    uint32_t v[32];
                    tmem_ld16(tl + buf * 256 + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                    tmem_ld16(tl + buf * 256 + c * 16 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                    tmem_wait_ld();
                    const int n = n0 + c * 16;
                    float ls = 0.f, lq = 0.f;
                    uint32_t hp[16];
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
                        float y[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float t = __uint_as_float(v[4 * i4 + e]) + bv[e];
                            t = (t >= 0.f) ? t : slope * t;
                            ls += t; lq = fmaf(t, t, lq);
                            y[e] = fminf(fmaxf(t * ysc, -65504.f), 65504.f);
                        }
                        hp[2 * i4] = pack_half2(y[0], y[1]);
                        hp[2 * i4 + 1] = pack_half2(y[2], y[3]);
                    }
                    lsum += (double)ls; lsq += (double)lq;
                    const uint4 ov[4] = {make_uint4(hp[0], hp[1], hp[2], hp[3]), make_uint4(hp[4], hp[5], hp[6], hp[7]),
                                         make_uint4(hp[8], hp[9], hp[10], hp[11]), make_uint4(hp[12], hp[13], hp[14], hp[15])};
                    g5_store_rows64(stg, lane, ov, reinterpret_cast<uint8_t*>(a.Y16 + wrow0 * a.ldY + n), pitch, nvalid);
                }
            } else if (EPI == EPI5_GLN_RES) {
                const size_t pitch = (size_t)a.ldY * sizeof(float);
                uint4 xcur[4];
                if (cbeg < cend) g5_load_rows64_commit(stg, lane, gA, xcur);
                auto chunk = [&](int c) {
                    uint32_t v[16];
                    tmem_ld16(tl + buf * 256 + c * 16, v);
                    tmem_wait_ld();
                    const int n = n0 + c * 16;
                    uint4 ov[4], orl[4];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float xa[4] = {__uint_as_float(xcur[i4].x), __uint_as_float(xcur[i4].y), __uint_as_float(xcur[i4].z), __uint_as_float(xcur[i4].w)};
                        const float4 s14 = __ldg(reinterpret_cast<const float4*>(s1 + n) + i4), b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                        const float s1v[4] = {s14.x, s14.y, s14.z, s14.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float val = fmaf(rstd, __uint_as_float(v[i4 * 4 + e]), fmaf(-mean * rstd, s1v[e], bv[e]));
                            o[e] = xa[e] + val;
                            if (valid) rowmax = fmaxf(rowmax, fabsf(o[e]));
                        }
                        ov[i4] = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
                        orl[i4] = make_uint4(__float_as_uint(fmaxf(o[0], 0.f)), __float_as_uint(fmaxf(o[1], 0.f)), __float_as_uint(fmaxf(o[2], 0.f)),
                                             __float_as_uint(fmaxf(o[3], 0.f)));
                    }
                    g5_store_rows64(stg, lane, ov, reinterpret_cast<uint8_t*>(a.Y + wrow0 * a.ldY + n), pitch, nvalid);
                    if (a.Xrelu) g5_store_rows64(stg, lane, orl, reinterpret_cast<uint8_t*>(a.Xrelu + wrow0 * a.ldY + n), pitch, nvalid);
                };
                // chunks in pairs over two register buffers: every residual load is issued two chunk-times before it is needed
                for (int c = cbeg; c < cend; c += 2) {
                    if (c + 2 < cend) g5_load_rows64_issue(lane, xbase + (size_t)(c + 2) * 64, pitch, nvalid, gA);
                    chunk(c);
                    if (c + 1 < cend) {
                        g5_load_rows64_commit(stg, lane, gB, xcur);
                        if (c + 3 < cend) g5_load_rows64_issue(lane, xbase + (size_t)(c + 3) * 64, pitch, nvalid, gB);
                        chunk(c + 1);
                        if (c + 2 < cend) g5_load_rows64_commit(stg, lane, gA, xcur);
                    }
                }
            } else {
                for (int c = cbeg; c < cend; ++c) {
                    uint32_t v[16];
                    tmem_ld16(tl + buf * 256 + c * 16, v);
                    tmem_wait_ld();
                    const int n = n0 + c * 16;
                    if (a.out_tm) {
                        // time-major output [(branch, b, t), Npad]: the layout the sub-band LSTM's x-tile builders read (k_lstm_tc5d.cu)
                        // (pad columns: zero weights, zero bias)
                        uint4 ov[4];
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + i4);
                            ov[i4] = make_uint4(__float_as_uint(apply_act(__uint_as_float(v[4 * i4]) + b4.x, a.act)), __float_as_uint(apply_act(__uint_as_float(v[4 * i4 + 1]) + b4.y, a.act)),
                                                __float_as_uint(apply_act(__uint_as_float(v[4 * i4 + 2]) + b4.z, a.act)), __float_as_uint(apply_act(__uint_as_float(v[4 * i4 + 3]) + b4.w, a.act)));
                        }
                        g5_store_rows64(stg, lane, ov, reinterpret_cast<uint8_t*>(a.out_tm + wrow0 * a.ldY + n), (size_t)a.ldY * sizeof(float), nvalid);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (valid && n + i < a.F)
                                a.out[((size_t)z * a.F + n + i) * a.P + tt] = apply_act(__uint_as_float(v[i]) + __ldg(bias + n + i), a.act);
                        }
                    }
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
            if (EPI == EPI5_GLN_RES && a.amax_out) {      // max |x| of the new stream per sample (max is order-independent: deterministic)
                const int key = valid ? z : -1;
                if (__match_any_sync(0xffffffffu, key) == 0xffffffffu) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) rowmax = fmaxf(rowmax, __shfl_xor_sync(0xffffffffu, rowmax, o));
                    if (lane == 0 && valid) atomicMax(reinterpret_cast<int*>(a.amax_out + z), __float_as_int(rowmax));
                } else if (valid) {
                    atomicMax(reinterpret_cast<int*>(a.amax_out + z), __float_as_int(rowmax));
                }
            }
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
    a.nstage = (227 * 1024 - 2048 - G5_EPI_WARPS * G5_STG_BYTES) / stage_bytes;
    if (a.nstage > 8) a.nstage = 8;
    if (a.nstage < 2 || a.NT % 16 || a.NT > 256 || a.Kp % (a.epi == EPI5_GLN_RES ? 64 : 32)) return (int)cudaErrorInvalidValue;
    if (a.epi == EPI5_PRELU_STATS && a.NT % 64) return (int)cudaErrorInvalidValue;     // its epilogue takes 32 columns per pass and warp half
    const size_t smem = (size_t)a.nstage * stage_bytes + G5_EPI_WARPS * G5_STG_BYTES + 1024 + 256;
    const int total = a.nbranch * a.tiles_m * a.ntiles_n;
    const int grid = total < num_sms ? total : num_sms;
    const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(mapA);
    const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(mapB);
    cudaError_t e;
#define G5_LAUNCH(E)                                                                                              \
    e = cudaFuncSetAttribute(gemm_tc5_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e != cudaSuccess) return (int)e;                                                                          \
    e = launch_chain(gemm_tc5_kernel<E>, dim3(grid), dim3(G5_THREADS), smem, s, mA, mB, a);                       \
    if (e != cudaSuccess) return (int)e;
    if (a.epi == EPI5_PRELU_STATS) { G5_LAUNCH(EPI5_PRELU_STATS) }
    else if (a.epi == EPI5_GLN_RES) { G5_LAUNCH(EPI5_GLN_RES) }
    else { G5_LAUNCH(EPI5_OUT) }
#undef G5_LAUNCH
    return (int)cudaGetLastError();
}

// gLN1 apply -> depth-wise conv (k=3, dilation d, zero padding in the normalised domain; causal or symmetric taps) -> PReLU2 + gLN2 statistics,
// on the time-major [rows, C] fp16 layout.  reference: causal_conv.py:100-106.  One CTA = one sample x 64 channels x DW_TCH frames
// (+ a halo of 2d frames either side): the slab is staged once in shared memory in fp32 with the gLN already applied, so every input
// element is read once from HBM/L2 (halo rows twice) and the three taps come from shared memory.  All arithmetic in fp32; the
// statistics are taken before the result is rounded to fp16.  Clips of any length: the frame axis is chunked.
constexpr int DW_CH = 64, DW_TCH = 256, DW_HALO = 18;    // 2 * max dilation (9) frames: covers both tap geometries
__global__ void __launch_bounds__(256) dwconv_tm_kernel(DwTmLaunch a) {
    extern __shared__ float slab[];                       // [rows <= DW_TCH + 2 DW_HALO][DW_CH]
    __shared__ double red[16];
    pdl_trigger();
    pdl_wait();
    const int z = blockIdx.z, g = z / a.B, c0 = blockIdx.x * DW_CH;
    const int C = a.C, Tp = a.Tp, d = a.dilation;
    const int t0 = blockIdx.y * DW_TCH, t1 = min(t0 + DW_TCH, Tp);
    const int lo = max(t0 - 2 * d, 0), hi = min(t1 + 2 * d, Tp);    // staged frames [lo, hi)
    const double cnt = (double)C * (double)Tp;
    const double mu = a.stats_in[2 * z] / cnt;
    const double var = a.stats_in[2 * z + 1] / cnt - mu * mu;
    const float mean = (float)mu, rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + 1e-8));
    const float slope = __ldg(a.prelu[g]);
    const float unscale = 1.0f / fp16_store_scale(__ldg(a.amax + z));     // exact: a power of two
    const size_t base = (size_t)z * Tp * C + c0;
    const int cq = threadIdx.x & 15, tr = threadIdx.x >> 4;           // 16 columns of 4 channels x 16 frame rows per pass
    // per-channel constants: staged once per CTA (64 channels x 6 values) -- as scalar loads per thread they were 85 % of the kernel's
    // global load requests (ncu, round 2).  Four channels per thread keep the kernel at 4 CTAs per SM (48 registers).
    __shared__ __align__(16) float cst[6][DW_CH];                     // gamma, beta, w0, w1, w2, bias
    if (threadIdx.x < DW_CH) {
        const int ch = c0 + threadIdx.x;
        cst[0][threadIdx.x] = __ldg(a.gamma[g] + ch);
        cst[1][threadIdx.x] = __ldg(a.beta[g] + ch);
        const float* wp = a.w[g] + (size_t)ch * 3;
        cst[2][threadIdx.x] = __ldg(wp); cst[3][threadIdx.x] = __ldg(wp + 1); cst[4][threadIdx.x] = __ldg(wp + 2);
        cst[5][threadIdx.x] = __ldg(a.b[g] + ch);
    }
    __syncthreads();
    float ga[4], be[4], w0[4], w1[4], w2[4], bb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ch = cq * 4 + e;
        const float gr = cst[0][ch] * rstd;
        ga[e] = gr * unscale;                                  // X holds y * 2^-k: (y - mean) rstd gamma + beta = X (2^k rstd gamma) + (beta - mean rstd gamma)
        be[e] = cst[1][ch] - mean * gr;
        w0[e] = cst[2][ch]; w1[e] = cst[3][ch]; w2[e] = cst[4][ch]; bb[e] = cst[5][ch];
    }
    for (int t = lo + tr; t < hi; t += 16) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(a.X + base + (size_t)t * C + cq * 4));
        const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), p1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        *reinterpret_cast<float4*>(slab + (size_t)(t - lo) * DW_CH + cq * 4) =
            make_float4(fmaf(p0.x, ga[0], be[0]), fmaf(p0.y, ga[1], be[1]), fmaf(p1.x, ga[2], be[2]), fmaf(p1.y, ga[3], be[3]));
    }
    __syncthreads();
    float ls = 0.f, lq = 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = t0 + tr; t < t1; t += 16) {
        // taps (w0, w1, w2): non-causal (t-d, t, t+d); causal (t-2d, t-d, t) -- padding 2d + chomp, causal_conv.py:74-75,104-105
        const int tl = a.causal ? t - 2 * d : t - d, tm = a.causal ? t - d : t, tr2 = a.causal ? t : t + d;
        const float4 m = (tm >= 0) ? *reinterpret_cast<const float4*>(slab + (size_t)(tm - lo) * DW_CH + cq * 4) : zero;
        const float4 l = (tl >= 0) ? *reinterpret_cast<const float4*>(slab + (size_t)(tl - lo) * DW_CH + cq * 4) : zero;
        const float4 r = (tr2 < Tp) ? *reinterpret_cast<const float4*>(slab + (size_t)(tr2 - lo) * DW_CH + cq * 4) : zero;
        const float lv[4] = {l.x, l.y, l.z, l.w}, mv[4] = {m.x, m.y, m.z, m.w}, rv[4] = {r.x, r.y, r.z, r.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float y = fmaf(w0[e], lv[e], fmaf(w1[e], mv[e], fmaf(w2[e], rv[e], bb[e])));
            y = (y >= 0.f) ? y : slope * y;
            ls += y; lq = fmaf(y, y, lq);
            o[e] = fminf(fmaxf(y, -65504.f), 65504.f);
        }
        *reinterpret_cast<uint2*>(a.Y + base + (size_t)t * C + cq * 4) = make_uint2(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]));
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

int launch_dwconv_tm(const DwTmLaunch& a, cudaStream_t s) {
    if (a.C % DW_CH || 2 * a.dilation > DW_HALO) return (int)cudaErrorInvalidValue;
    const int rows = (a.Tp < DW_TCH ? a.Tp : DW_TCH + 2 * DW_HALO);
    const size_t smem = (size_t)rows * DW_CH * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(dwconv_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_chain(dwconv_tm_kernel, dim3(a.C / DW_CH, (a.Tp + DW_TCH - 1) / DW_TCH, a.Z), dim3(256), smem, s, a);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

}  // namespace fsn
