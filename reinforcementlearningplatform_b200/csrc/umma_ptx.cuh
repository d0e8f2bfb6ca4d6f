// umma_ptx.cuh -- inline-PTX wrappers of the Blackwell tensor-core path shared by the K-POLICY kernels (policy_umma.cu:
// 3xTF32, activations in TMEM or shared memory, streamed wide nets; policy_umma16.cu: fp16-split operands, four tiles in
// flight): mbarrier, bulk copy (TMA), tcgen05.alloc / mma / commit / ld / st, descriptors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef UMMA_WAIT_HINT_NS
#define UMMA_WAIT_HINT_NS 20000u
#endif

namespace umma {

constexpr long long WAIT_TIMEOUT = 4000000000ll;       // cycles (~2 s): a wait that long is a protocol bug -> trap

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// non-blocking probe of a phase (mbar_try may suspend the thread for a while)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// try_wait with the default (short, implementation-defined) time limit
__device__ __forceinline__ uint32_t mbar_try_short(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or ~the hint (ns) has passed,
// so a waiting warp costs a handful of issue slots per wait instead of a polling loop (with the default, short time limit
// 16 epilogue warps spent 22 % of all issued instructions spinning -- ncu, profiles/r2/policy_umma.md)
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(UMMA_WAIT_HINT_NS) : "memory");
    return ok;
}
// `parity`: the phase parity whose completion is awaited; on a fresh barrier parity 1 passes at once (free-type barriers).
// A wait that lasts ~2 s is a protocol bug: trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
        // the clock is read once per 1024 probes: inside the probe loop it was a third of the loop's instructions, and
        // the epilogue warps of the policy kernels spend ~18 probes per layer hand-off
#pragma unroll 1
        for (int k = 0; k < 1024; ++k)
            if (mbar_try(bar, parity)) return;
        if (clock64() - t0 > WAIT_TIMEOUT) __trap();
    }
}
// true for exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "elect.sync _|p, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at
// [16,30), SBO >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32 (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 clear), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same two instructions for a warp that runs its issue loop in lock step: every lane executes the asm, the lane with
// `leader` != 0 issues.  Descriptors are passed as (low word, shared high word).
__device__ __forceinline__ void umma_tf32_lohi(uint32_t leader, uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32,
                                               uint32_t desc_hi32, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "setp.ne.b32 q, %6, 0;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// TS form: A operand read from TMEM (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t desc_hi32,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint32_t leader, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\t"
                 "setp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(bar), "r"(leader) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// thread t of the warp receives columns col .. col + W - 1 of TMEM lane (warp_id % 4) * 32 + t
template <int W> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[W]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<1>(uint32_t taddr, uint32_t (&v)[1]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// thread t of the warp writes columns col .. col + 7 of TMEM lane (warp_id % 4) * 32 + t
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}


// instruction descriptor, kind::f16 with fp16 operands: D = F32 (1 << 4), A = B = F16 (format 0), K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SS form, kind::f16 (fp16 operands in shared memory, K = 16 per instruction): same calling convention as umma_tf32_lohi
__device__ __forceinline__ void umma_f16_lohi(uint32_t leader, uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32,
                                              uint32_t desc_hi32, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "setp.ne.b32 q, %6, 0;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// TS form, kind::f16: A = 16-bit elements packed two per 32-bit TMEM column (lane = row), B from shared memory; K = 16
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t desc_hi32,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate) : "memory");
}

// round to TF32's 10-bit mantissa (ties away) on the bit pattern: hi has its low 13 bits clear, so the tensor core's
// own handling of those bits cannot matter; lo = x - hi is exact in fp32 and is at most 2^-11 |x|
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

} // namespace umma
