// Input projection of the layer-wise LSTM path (k_lstm_tc5r.cu) as a hand-written tcgen05 GEMM:
//   Gin[M, 4H] = X[M, K] * Wp[4H, K]^T      fp16 operands, fp32 accumulation in tensor memory, fp16 result
// reference: the W_ih x_t half of nn.LSTM / nn.GRU (audio_zen/model/module/sequence_model.py:31-46,113-119), batched over all
// rows and time steps because it has no recurrence.  M = row tiles x frames x 128 (a multiple of 128), 4H a multiple of 256,
// K = 64 (layer 0: the packed sub-band input) or H.
//
// Persistent kernel, one CTA per SM, tiles of 128 x 256: both operands K-major and fetched by TMA (cp.async.bulk.tensor.2d,
// SWIZZLE_128B, 64-half k-blocks) into a 4-stage ring, tcgen05.mma.kind::f16 (M = 128, N = 256, K = 16) into two 256-column
// accumulators so the drain of tile i overlaps the MMAs of tile i + 1; eight epilogue warps (two per TMEM lane quarter) convert to
// fp16 and store 32 bytes per thread and 16-column chunk.  Tiles are walked N-fastest: the eight CTAs that share an A tile run at
// the same time (A is read from HBM once, the 2 MB of weights stay in L2).  Bound: the fp16 result (M x 4H x 2 bytes, 6.5 GB per
// layer at BASELINE config #5) has to be written to HBM -- the kernel is write-bound at ~2 ms per layer there.
#include <cuda.h>

#include "fsn_common.cuh"
#include "fsn_kernels.h"

namespace fsn {

constexpr int GF_EPI_WARPS = 8;
constexpr int GF_THREADS = (2 + GF_EPI_WARPS) * 32;     // warp 0 TMA producer, warp 1 MMA issuer + TMEM alloc, warps 2-9 epilogue
constexpr int GF_BM = 128, GF_BN = 256, GF_BK = 64;
constexpr int GF_A_BYTES = GF_BM * GF_BK * 2, GF_B_BYTES = GF_BN * GF_BK * 2, GF_STAGE = GF_A_BYTES + GF_B_BYTES;
constexpr int GF_NSTAGE = 4;

__device__ __forceinline__ void gf_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(GF_THREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmF16Launch a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)GF_NSTAGE * GF_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = full + GF_NSTAGE;
    uint64_t* accfull = empty + GF_NSTAGE;
    uint64_t* accempty = accfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < GF_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], GF_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nkb = a.K / GF_BK, tiles_n = a.N / GF_BN, total = (a.M / GF_BM) * tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int mt = tile / tiles_n, nt = tile % tiles_n;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[slot], ph ^ 1);
                    uint8_t* st = smem + (size_t)slot * GF_STAGE;
                    mbar_arrive_expect_tx(&full[slot], GF_STAGE);
                    gf_tma_load_2d(st, &mapA, kb * GF_BK, mt * GF_BM, &full[slot]);
                    gf_tma_load_2d(st + GF_A_BYTES, &mapB, kb * GF_BK, nt * GF_BN, &full[slot]);
                    if (++slot == GF_NSTAGE) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // warp-uniform issue loop, the tcgen05 instructions predicated on one elected lane (see elect_one in fsn_common.cuh)
        constexpr uint32_t IDESC = umma_idesc_f16(GF_BM, GF_BN);
        int slot = 0; uint32_t ph = 0, use[2] = {0, 0}, it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&accempty[buf], (use[buf] & 1) ^ 1);
            ++use[buf];
            tc5_fence_after();
            const uint32_t d = tmem + buf * GF_BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[slot], ph);
                tc5_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)slot * GF_STAGE);
                const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + GF_A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < GF_BK / 16; ++kk) umma_ss(d, adesc + 2 * kk, bdesc + 2 * kk, IDESC, (kb | kk) != 0);
                    umma_commit(&empty[slot]);
                    if (kb == nkb - 1) umma_commit(&accfull[buf]);
                }
                __syncwarp();
                if (++slot == GF_NSTAGE) { slot = 0; ph ^= 1; }
            }
        }
    } else {
        const int q = warp & 3, r = q * 32 + lane, half = (warp - 2) >> 2;       // TMEM lane quarter, row in the tile, column half
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t use[2] = {0, 0}, it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int mt = tile / tiles_n, nt = tile % tiles_n, buf = it & 1;
            __half* crow = a.C + ((size_t)mt * GF_BM + r) * a.ldc + (size_t)nt * GF_BN + half * (GF_BN / 2);
            mbar_wait(&accfull[buf], use[buf] & 1);
            ++use[buf];
            tc5_fence_after();
#pragma unroll 2
            for (int c = 0; c < GF_BN / 2 / 16; ++c) {
                uint32_t v[16];
                tmem_ld16(tl + buf * GF_BN + half * (GF_BN / 2) + c * 16, v);
                tmem_wait_ld();
                uint4 o0, o1;
                o0.x = pack_half2(__uint_as_float(v[0]), __uint_as_float(v[1]));   o0.y = pack_half2(__uint_as_float(v[2]), __uint_as_float(v[3]));
                o0.z = pack_half2(__uint_as_float(v[4]), __uint_as_float(v[5]));   o0.w = pack_half2(__uint_as_float(v[6]), __uint_as_float(v[7]));
                o1.x = pack_half2(__uint_as_float(v[8]), __uint_as_float(v[9]));   o1.y = pack_half2(__uint_as_float(v[10]), __uint_as_float(v[11]));
                o1.z = pack_half2(__uint_as_float(v[12]), __uint_as_float(v[13])); o1.w = pack_half2(__uint_as_float(v[14]), __uint_as_float(v[15]));
                uint4* dst = reinterpret_cast<uint4*>(crow + c * 16);
                __stcs(dst, o0);                                          // written once, read once by the recurrent kernel: streaming
                __stcs(dst + 1, o1);
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&accempty[buf]);
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
typedef CUresult (*GfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major fp16 matrix [rows, cols] (cols a multiple of 64), boxes of box_rows x 64 halves, SWIZZLE_128B
int make_tmap_f16_2d(void* out_map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static GfEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<GfEncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * sizeof(__half)};
    const cuuint32_t box[2] = {GF_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r;
}

bool gemm_f16_supported(long long M, int N, int K) { return M > 0 && M % GF_BM == 0 && N % GF_BN == 0 && K % GF_BK == 0 && K >= GF_BK; }

int launch_gemm_f16(const void* A, const void* B, const GemmF16Launch& a, int num_sms, cudaStream_t s) {
    if (!gemm_f16_supported(a.M, a.N, a.K)) return (int)cudaErrorInvalidValue;
    alignas(64) CUtensorMap mA, mB;
    if (make_tmap_f16_2d(&mA, A, (uint64_t)a.M, (uint64_t)a.K, GF_BM) || make_tmap_f16_2d(&mB, B, (uint64_t)a.N, (uint64_t)a.K, GF_BN))
        return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)GF_NSTAGE * GF_STAGE + 1024 + 256;
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long total = (a.M / GF_BM) * (long long)(a.N / GF_BN);
    const int grid = total < num_sms ? (int)total : num_sms;
    gemm_f16_kernel<<<grid, GF_THREADS, smem, s>>>(mA, mB, a);
    return (int)cudaGetLastError();
}

}  // namespace fsn
