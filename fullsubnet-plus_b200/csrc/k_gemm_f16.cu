// Input projection of the layer-wise LSTM path (k_lstm_tc5r.cu) as a hand-written tcgen05 GEMM:
//   Gin[M, 4H] = X[M, K] * Wp[4H, K]^T      fp16 operands, fp32 accumulation in tensor memory, fp16 result
// reference: the W_ih x_t half of nn.LSTM / nn.GRU (audio_zen/model/module/sequence_model.py:31-46,113-119), batched over all
// rows and time steps because it has no recurrence.  M = row tiles x frames x 128 (a multiple of 128), 4H a multiple of 256,
// K = 64 (layer 0: the packed sub-band input) or H.
//
// Persistent kernel, one CTA per SM, tiles of 128 x 256: both operands K-major and fetched by TMA (cp.async.bulk.tensor.2d,
// SWIZZLE_128B, 64-half k-blocks) into a 3-stage ring, tcgen05.mma.kind::f16 (M = 128, N = 256, K = 16) into two 256-column
// accumulators so the drain of tile i overlaps the MMAs of tile i + 1; eight epilogue warps (two per TMEM lane quarter) convert to
// fp16, stage the tile in shared memory in the swizzled layout of the result's tensor map and hand it to the TMA engine
// (cp.async.bulk.tensor store: full 128-byte lines, asynchronous -- per-thread 32-byte stores at a 4 KB row pitch reached only 2.3 TB/s
// on the write-bound K = 64 layer).  Tiles are walked N-fastest: the eight CTAs that share an A tile run at
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
constexpr int GF_NSTAGE = 3;
constexpr int GF_CSUB = GF_BM * 64 * 2;                  // one 128-row x 64-column fp16 sub-tile of the result staged for the TMA store (SW128)
constexpr int GF_CSTAGE = (GF_BN / 64) * GF_CSUB;        // 64 KB: the whole 128 x 256 result tile

__device__ __forceinline__ void gf_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// result tile: shared memory -> global through the TMA engine (full 128-byte lines, asynchronous: the epilogue warps go on to the next tile)
__device__ __forceinline__ void gf_tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void gf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gf_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void gf_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void gf_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__global__ void __launch_bounds__(GF_THREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
                GemmF16Launch a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* cstage = smem + (size_t)GF_NSTAGE * GF_STAGE;                 // [4 sub-tiles][128 rows][64 halves], SWIZZLE_128B
    uint64_t* bars = reinterpret_cast<uint64_t*>(cstage + GF_CSTAGE);
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
        // epilogue: TMEM -> fp16 -> shared memory (the swizzled layout of the C tensor map) -> TMA store.  Warp (q, half): rows
        // q*32.., columns half*128.. = two 64-column sub-tiles; the accumulator is released as soon as it is in registers.
        const int q = warp & 3, r = q * 32 + lane, half = (warp - 2) >> 2;
        const bool chief = (warp == 2 && lane == 0);
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t use[2] = {0, 0}, it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int mt = tile / tiles_n, nt = tile % tiles_n, buf = it & 1;
            mbar_wait(&accfull[buf], use[buf] & 1);
            ++use[buf];
            tc5_fence_after();
            if (it > 0) {                                             // the previous tile's TMA stores have read the staging buffer
                if (chief) gf_bulk_wait_read0();
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint8_t* cs = cstage + (size_t)(half * 2 + sub) * GF_CSUB;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    gf_tmem_ld32(tl + buf * GF_BN + half * 128 + sub * 64 + c * 32, v);
                    tmem_wait_ld();
                    if (sub == 1 && c == 1) {                         // last read of this accumulator: hand it back to the MMA issuer
                        tc5_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&accempty[buf]);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        o.x = pack_half2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
                        o.y = pack_half2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
                        o.z = pack_half2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                        o.w = pack_half2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                        *reinterpret_cast<uint4*>(cs + sw128_offset(r, c * 32 + g * 8)) = o;
                    }
                }
            }
            fence_proxy_async();                                      // generic-proxy stores -> visible to the TMA engine
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (chief) {
#pragma unroll
                for (int sub = 0; sub < GF_BN / 64; ++sub) gf_tma_store_2d(&mapC, cstage + (size_t)sub * GF_CSUB, nt * GF_BN + sub * 64, mt * GF_BM);
                gf_bulk_commit();
            }
        }
        if (chief) gf_bulk_wait0();                                   // all result tiles are in global memory before the CTA exits
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
    if (a.ldc != a.N) return (int)cudaErrorInvalidValue;          // the C tensor map describes a dense [M, N] matrix
    alignas(64) CUtensorMap mA, mB, mC;
    if (make_tmap_f16_2d(&mA, A, (uint64_t)a.M, (uint64_t)a.K, GF_BM) || make_tmap_f16_2d(&mB, B, (uint64_t)a.N, (uint64_t)a.K, GF_BN) ||
        make_tmap_f16_2d(&mC, a.C, (uint64_t)a.M, (uint64_t)a.N, GF_BM))
        return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)GF_NSTAGE * GF_STAGE + GF_CSTAGE + 1024 + 256;
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long total = (a.M / GF_BM) * (long long)(a.N / GF_BN);
    const int grid = total < num_sms ? (int)total : num_sms;
    gemm_f16_kernel<<<grid, GF_THREADS, smem, s>>>(mA, mB, mC, a);
    return (int)cudaGetLastError();
}

}  // namespace fsn
