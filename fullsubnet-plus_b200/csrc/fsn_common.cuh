// Common device helpers for the fsnplus_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#ifdef FSN_MBAR_DEBUG
#include <cstdio>
#endif

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "fsnplus_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace fsn {

// ----------------------------------------------------------------------------------------------
// small math
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// sigmoid / tanh from ex2 + rcp (about 2 ulp each): the accurate path.
__device__ __forceinline__ float sigmoid_acc(float x) {
    // 1 / (1 + 2^(-x*log2e)); ex2 saturates to +inf / 0 cleanly, rcp(inf) = 0.
    return rcpf(1.0f + ex2f(-1.4426950408889634f * x));
}
__device__ __forceinline__ float tanh_acc(float x) {
    // 1 - 2 / (1 + 2^(2x*log2e))
    return 1.0f - 2.0f * rcpf(1.0f + ex2f(2.8853900817779268f * x));
}
// MUFU.TANH based variants (max rel. err 2^-11): the fast path, selectable per kernel.
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

template <bool FAST> __device__ __forceinline__ float sigm(float x) { return FAST ? sigmoid_fast(x) : sigmoid_acc(x); }
template <bool FAST> __device__ __forceinline__ float tanh_(float x) { return FAST ? tanh_approx(x) : tanh_acc(x); }

// output activation of SequenceModel (sequence_model.py:84-93, 120-121); codes = FSN_ACT_* of include/fsnplus_b200.h
__device__ __forceinline__ float apply_act(float y, int act) {
    if (act == 1) return fmaxf(y, 0.f);
    if (act == 2) return tanhf(y);
    if (act == 3) return fminf(fmaxf(y, 0.f), 6.f);
    return y;
}

// decompress_cIRM (audio_zen/acoustics/mask.py:60-63, K = 10, limit = 9.9): limit*(m >= limit) - limit*(m <= -limit) + m*(|m| < limit)
// is a clamp; then -K log((K - m) / (K + m)).
__device__ __forceinline__ float decompress_cirm(float v) {
    v = (v >= 9.9f) ? 9.9f : ((v <= -9.9f) ? -9.9f : v);
    return -10.f * logf((10.f - v) / (10.f + v));
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Programmatic dependent launch (front-end chain, ~30 short kernels per forward): a kernel launched with launch_chain() may start
// while its predecessor in the stream is still draining -- its CTAs are placed as SMs free up and run their prologue (barrier
// init, TMEM allocation) -- and MUST execute pdl_wait() before it touches anything the predecessor reads or writes; pdl_wait returns
// when the predecessor grid has completed and its memory is visible (a no-op for a normally launched kernel).  pdl_trigger() lets
// the successor begin launching once every CTA of this grid has executed it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Power-of-two down-scale for storing an activation in fp16 whose input stream has absolute maximum `amax` (per sample): exact
// (no rounding), never amplifies (a bias term must not overflow when the stream is tiny), 1 for amax < 1.
__device__ __forceinline__ float fp16_store_scale(float amax) {
    int e = 0;
    frexpf(fminf(amax, 3.0e38f), &e);                  // amax = m 2^e, m in [0.5, 1)
    e = e < 0 ? 0 : (e > 126 ? 126 : e);
    return __int_as_float((127 - e) << 23);            // 2^-e
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reflect index used by BaseModel.unfold (F.pad mode="reflect", no edge repeat):
// reference audio_zen/model/base_model.py:38.
__host__ __device__ __forceinline__ int reflect_idx(int p, int F) {
    if (p < 0) p = -p;
    if (p > F - 1) p = 2 * (F - 1) - p;
    return p;
}

// Byte offset of element (row r, half k) inside a 128-row x 64-half K-major tile stored with the
// 128-byte swizzle the tcgen05 shared-memory descriptor (SWIZZLE_128B) expects: rows at 128 B pitch,
// 16-byte chunk index XORed with (row & 7).  Tile base must be 1024-byte aligned.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t k) {
    return r * 128u + ((((k >> 3) ^ (r & 7u)) & 7u) << 4) + ((k & 7u) << 1);
}

// ----------------------------------------------------------------------------------------------
// shared-memory / mbarrier / bulk-copy / tcgen05 PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> launch error reported to the host), never hang the box.
#ifndef FSN_MBAR_TIMEOUT_CYCLES
#define FSN_MBAR_TIMEOUT_CYCLES 3000000000ll   // ~2 s at B200 clocks; no legitimate wait is longer than ms
#endif
#ifdef FSN_MBAR_DEBUG
// debug builds (-DFSN_MBAR_DEBUG): a watchdog expiry reports WHICH wait of which role hung before trapping
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
__device__ __noinline__ void mbar_report(int line, uint32_t parity) {
    printf("mbar timeout: line %d block %d thread %d (warp %d) parity %u\n", line, (int)blockIdx.x, (int)threadIdx.x, (int)(threadIdx.x >> 5), parity);
}
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int line) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > FSN_MBAR_TIMEOUT_CYCLES) { mbar_report(line, parity); __nanosleep(2000000); __trap(); }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > FSN_MBAR_TIMEOUT_CYCLES) { __trap(); }
    }
}
#endif

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Same copy with an L2 eviction-priority hint: evict_first for data that is read exactly once (the per-step input images, 392 MB
// per launch at B = 64, which otherwise push the 50 MB cell-state scratch out of L2), evict_last for the weight stream that every
// CTA re-reads every step.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// plain global accesses with an L2 eviction-priority hint (same policies as the bulk copies above)
__device__ __forceinline__ float ldg_hint(const float* p, uint64_t policy) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint4 ldg_u4_hint(const uint4* p, uint64_t policy) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ float4 ld_f4_hint(const float4* p, uint64_t policy) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy) : "memory");
    return v;
}
__device__ __forceinline__ void st_f4_hint(float4* p, float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}

// One lane of a fully converged warp (elect.sync).  The tcgen05 issue loops must stay warp-uniform and predicate only
// the instruction on this: an `if (lane == 0)` region forces every UTCHMMA operand through R2UR moves and serialises the
// mbarrier polls with the issue -- measured 109 vs 64 cycles per M=128,N=128 MMA (fsn_probe_tcgen05, "probe4").
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 64-half (128 B) rows, 8-row groups 1024 B
// apart (field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor; version = 1 on sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);          // start address      [0,14)
    d |= (uint64_t)1u << 16;                              // leading byte off.  [16,30) (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                    // stride byte offset [32,46)
    d |= (uint64_t)1u << 46;                              // version            [46,48)
    d |= (uint64_t)2u << 61;                              // SWIZZLE_128B       [61,64)
    return d;
}
// Instruction descriptor kind::f16: D=f32, A=B=f16, both K-major (InstrDescriptor in the same header).
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)                 // c_format = F32
           | (0u << 7) | (0u << 10)  // a_format = b_format = F16
           | (0u << 15) | (0u << 16) // a_major = b_major = K
           | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Make an mbarrier track completion of all tcgen05 ops issued so far by this thread
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns: thread i of the warp <-> TMEM lane (lane_base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// ---- CTA pairs (cta_group::2): one tcgen05.mma drives both SMs of a 2-CTA cluster; each CTA holds half of B ---------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address valid in every CTA) inside CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
// Remote arrive.  Relaxed on purpose: a cluster-scope RELEASE costs ~1K cycles per arrive (measured: 26 ms vs 12 ms
// kernel), and nothing but the barrier state has to be published -- the payload it guards was written by the async
// proxy (bulk copy -> shared memory) or lives in tensor memory (ordered by tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// cluster-scope variants for data handed over through distributed shared memory (column-split small-batch mode of k_lstm_tc5d.cu):
// the writer stores into the remote CTA, fences at cluster scope and arrives; the reader's wait acquires at cluster scope
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
#ifdef FSN_MBAR_DEBUG
#define mbar_wait_cluster(bar, parity) mbar_wait_cluster_dbg(bar, parity, __LINE__)
__device__ __forceinline__ void mbar_wait_cluster_dbg(uint64_t* bar, uint32_t parity, int line) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > FSN_MBAR_TIMEOUT_CYCLES) { mbar_report(line, parity); __nanosleep(2000000); __trap(); }
    }
}
#else
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > FSN_MBAR_TIMEOUT_CYCLES) { __trap(); }
    }
}
#endif
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2f(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// remote store that carries its own completion: the 16 (8) bytes land in the destination CTA and its mbarrier's transaction count is
// decremented by the same amount -- no fence, no separate arrive; a waiter that observes the phase sees the data
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w), "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_v2f(uint32_t cluster_addr, float a, float b, uint32_t cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr), "f"(a), "f"(b), "r"(cluster_mbar)
                 : "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior cta_group::2 MMAs -> arrive on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// One LSTM cell.  ai/af/ao = -log2(e) * (gate pre-activation), ag = -2 log2(e) * (g pre-activation): the scale and
// the bias are folded into one FMA on the accumulator (biases are stored pre-scaled).
// Accurate path: 5 ex2 + 3 rcp (i*tanh(g) and o*tanh(c) share one reciprocal each); fast path: 5 tanh.approx.
template <bool FAST>
__device__ __forceinline__ void lstm_cell(float ai, float af, float ag, float ao, float cprev, float& c, float& h) {
    if (FAST) {
        const float K = -0.34657359027997264f;                                // -0.5 / log2(e)
        const float si = fmaf(0.5f, tanh_approx(ai * K), 0.5f), sf = fmaf(0.5f, tanh_approx(af * K), 0.5f);
        c = fmaf(sf, cprev, si * tanh_approx(ag * K));
        h = fmaf(0.5f, tanh_approx(ao * K), 0.5f) * tanh_approx(c);
    } else {
        const float ei = ex2f(ai), ef = ex2f(af);
        const float eg = ex2f(fminf(fmaxf(ag, -43.f), 43.f));                  // tanh saturates: |g| <= 15
        const float ig = (1.f - eg) * rcpf((1.f + ei) * (1.f + eg));          // sigmoid(i) * tanh(g)
        c = fmaf(rcpf(1.f + ef), cprev, ig);
        const float eo = ex2f(ao);
        const float ec = ex2f(fminf(fmaxf(c * -2.8853900817779268f, -43.f), 43.f));
        h = (1.f - ec) * rcpf((1.f + eo) * (1.f + ec));                       // sigmoid(o) * tanh(c)
    }
}

// One GRU cell on the same four accumulators: the GRU weights are packed as pseudo-gates (r, z, n_x, n_h) -- n_x holds
// W_in x + b_in (zero recurrent block), n_h holds W_hn h + b_hn (zero input block) -- so every LSTM kernel computes it
// unchanged up to this function.  nn.GRU (sequence_model.py:39-46): n = tanh(n_x + r n_h); h = (1 - z) n + z h_prev.
// ar/az = -log2(e) * pre-activation; anx/anh = -2 log2(e) * pre-activation.  The "cell state" slot carries h in fp32.
template <bool FAST>
__device__ __forceinline__ void gru_cell(float ar, float az, float anx, float anh, float hprev, float& h) {
    float r, z, n;
    if (FAST) {
        const float K = -0.34657359027997264f;                                // -0.5 / log2(e)
        r = fmaf(0.5f, tanh_approx(ar * K), 0.5f);
        z = fmaf(0.5f, tanh_approx(az * K), 0.5f);
        n = tanh_approx(fmaf(r, anh, anx) * K);
    } else {
        r = rcpf(1.f + ex2f(ar));
        z = rcpf(1.f + ex2f(az));
        const float en = ex2f(fminf(fmaxf(fmaf(r, anh, anx), -43.f), 43.f));
        n = (1.f - en) * rcpf(1.f + en);
    }
    h = fmaf(z, hprev - n, n);
}
// plain-argument form for the generic kernels
template <bool FAST>
__device__ __forceinline__ float gru_cell_plain(float gr, float gz, float gnx, float gnh, float hprev) {
    const float r = sigm<FAST>(gr), z = sigm<FAST>(gz);
    const float n = tanh_<FAST>(fmaf(r, gnh, gnx));
    return fmaf(z, hprev - n, n);
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}


// ---- legacy tensor path (mma.sync) used by the generic kernels ------------------------------
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t f2tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr));
}

}  // namespace fsn
