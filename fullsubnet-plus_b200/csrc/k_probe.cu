// tcgen05 / TMEM / bulk-copy self-test: tiny single-CTA GEMMs through every instruction form the persistent
// LSTM kernel relies on (SS and TS tcgen05.mma with SWIZZLE_128B K-major tiles staged by cp.async.bulk,
// tcgen05.st as the A-operand writer, tcgen05.ld 32x32b as the accumulator reader, tcgen05.commit ->
// mbarrier), checked against a host computation, plus issue-rate measurements of the four MMA shapes the
// design discussion in DESIGN.md quotes.  Exposed as fsn_probe_tcgen05() and run by tests/test_gpu_probe.py.
#include "fsn_common.cuh"
#include "fsn_kernels.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fsn {

struct ProbeArgs {
    const uint8_t* a_img;    // 16 KB  [128 x 64] SW128
    const uint8_t* b_img;    // 32 KB  [256 x 64] SW128 (first 64 rows used when N = 64)
    const uint32_t* a_plain; // [128][32] packed half2 (row-major) for the TMEM A operand
    float* d;                // [128][256]
    long long* cycles;       // [1]
    int mode;                // 0 SS, 1 TS, 2 SS+TS accumulate, 3 timing
    int N;                   // 64 or 256
    int ts;                  // timing: 1 = TS, 0 = SS
    int reps;
    int nacc;                // timing: number of accumulators cycled through (1 = dependent chain)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;                 // 16 KB
    uint8_t* sb = smem + 16384;         // 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(tslot);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col = 256;         // TMEM A operand region (32 columns = 64 halves)
    const int N = p.N;

    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], 16384 + N * 128);
        bulk_g2s(sa, p.a_img, 16384, &bars[0]);
        bulk_g2s(sb, p.b_img, N * 128, &bars[0]);
    }
    // A operand into TMEM: thread r holds row r, 32 packed columns
    {
        const uint32_t* src = p.a_plain + (size_t)tid * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = src[c * 8 + i];
            tmem_st8(tl + a_col + c * 8, v);
        }
        tmem_wait_st();
    }
    mbar_wait(&bars[0], 0);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();

    if (tid == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t adesc = umma_desc_sw128(smem_u32(sa)), bdesc = umma_desc_sw128(smem_u32(sb));
        long long t0 = clock64();
        if (p.mode == 0) {
            for (int kk = 0; kk < 4; ++kk) umma_ss(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
        } else if (p.mode == 1) {
            for (int kk = 0; kk < 4; ++kk) umma_ts(tmem, tmem + a_col + kk * 8, bdesc + 2 * kk, idesc, kk != 0);
        } else if (p.mode == 2) {
            for (int kk = 0; kk < 4; ++kk) umma_ss(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
            for (int kk = 0; kk < 4; ++kk) umma_ts(tmem, tmem + a_col + kk * 8, bdesc + 2 * kk, idesc, 1);
        } else {
            // accumulators at columns 0 and (nacc == 2 ? N : 0); the TMEM A operand sits at a_col (>= 2N for N <= 128)
            for (int rpt = 0; rpt < p.reps; ++rpt)
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t d = tmem + ((p.nacc == 2 && (kk & 1)) ? (uint32_t)N : 0u);
                    if (p.ts) umma_ts(d, tmem + a_col + kk * 8, bdesc + 2 * kk, idesc, rpt != 0 || kk > 1);
                    else umma_ss(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, rpt != 0 || kk > 1);
                }
        }
        umma_commit(&bars[1]);
        mbar_wait(&bars[1], 0);
        p.cycles[0] = clock64() - t0;
    }
    __syncthreads();
    mbar_wait(&bars[1], 0);
    tc5_fence_after();
    for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(tl + c * 16, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) p.d[(size_t)tid * 256 + c * 16 + i] = __uint_as_float(v[i]);
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

static uint16_t h_bits(float f) { __half h = __float2half_rn(f); uint16_t b; std::memcpy(&b, &h, 2); return b; }
static float h_val(uint16_t b) { __half h; std::memcpy(&h, &b, 2); return __half2float(h); }

#define PROBE_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = -(int)e_ - 1000; goto done; } } while (0)

int run_probe_tcgen05(float* report, int n) {
    if (n < 16) return -1;
    int rc = 0;
    std::vector<uint16_t> A(128 * 64), B(256 * 64);
    std::vector<uint8_t> aimg(16384), bimg(32768);
    std::vector<uint32_t> aplain(128 * 32);
    std::vector<float> D(128 * 256);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : A) v = h_bits(rnd());
    for (auto& v : B) v = h_bits(rnd());
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) {
            std::memcpy(&aimg[sw128_offset(r, k)], &A[r * 64 + k], 2);
            if (k % 2 == 0) aplain[r * 32 + k / 2] = (uint32_t)A[r * 64 + k] | ((uint32_t)A[r * 64 + k + 1] << 16);
        }
    for (int r = 0; r < 256; ++r)
        for (int k = 0; k < 64; ++k) std::memcpy(&bimg[(size_t)(r / 8) * 1024 + sw128_offset(r % 8, k)], &B[r * 64 + k], 2);

    uint8_t *da = nullptr, *db = nullptr; uint32_t* dp = nullptr; float* dd = nullptr; long long* dc = nullptr;
    const size_t smem = 16384 + 32768 + 64 + 1024;
    ProbeArgs p{};
    PROBE_CK(cudaMalloc(&da, 16384)); PROBE_CK(cudaMalloc(&db, 32768)); PROBE_CK(cudaMalloc(&dp, 128 * 32 * 4));
    PROBE_CK(cudaMalloc(&dd, 128 * 256 * 4)); PROBE_CK(cudaMalloc(&dc, 8));
    PROBE_CK(cudaMemcpy(da, aimg.data(), 16384, cudaMemcpyHostToDevice));
    PROBE_CK(cudaMemcpy(db, bimg.data(), 32768, cudaMemcpyHostToDevice));
    PROBE_CK(cudaMemcpy(dp, aplain.data(), 128 * 32 * 4, cudaMemcpyHostToDevice));
    PROBE_CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.a_img = da; p.b_img = db; p.a_plain = dp; p.d = dd; p.cycles = dc;
    for (int mode = 0; mode < 3; ++mode) {
        p.mode = mode; p.N = 64; p.reps = 1; p.ts = 0; p.nacc = 1;
        PROBE_CK(cudaMemset(dd, 0, 128 * 256 * 4));
        probe_kernel<<<1, 128, smem>>>(p);
        PROBE_CK(cudaGetLastError());
        PROBE_CK(cudaDeviceSynchronize());
        PROBE_CK(cudaMemcpy(D.data(), dd, 128 * 256 * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 64; ++c) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)h_val(A[r * 64 + k]) * (double)h_val(B[c * 64 + k]);
                if (mode == 2) ref *= 2.0;
                maxerr = std::fmax(maxerr, std::fabs(ref - (double)D[r * 256 + c]));
            }
        report[mode] = (float)maxerr;
    }
    {
        int idx = 3;
        // cycles per tcgen05.mma (M=128, K=16) for N in {64,128,192,256} x {TS,SS} x {1,2 accumulators}
        for (int N : {64, 128, 192, 256})
            for (int ts : {1, 0})
                for (int nacc : {1, 2}) {
                    if (nacc == 2 && N > 128) continue;           // two accumulators + A region must fit 512 columns
                    if (idx >= n) break;
                    p.mode = 3; p.N = N; p.ts = ts; p.reps = 256; p.nacc = nacc;
                    probe_kernel<<<1, 128, smem>>>(p);
                    PROBE_CK(cudaGetLastError());
                    PROBE_CK(cudaDeviceSynchronize());
                    long long cyc = 0;
                    PROBE_CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
                    report[idx++] = (float)cyc / (256.0f * 4.0f);
                }
        rc = idx;
    }
done:
    cudaFree(da); cudaFree(db); cudaFree(dp); cudaFree(dd); cudaFree(dc);
    return rc;
}

}  // namespace fsn
