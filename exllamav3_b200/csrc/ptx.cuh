// Thin inline-PTX wrappers for the sm_100a features this library uses: mbarrier, bulk async copy (TMA engine,
// SASS UBLKCP), tcgen05 (TMEM alloc / st / ld / mma / commit / fences), programmatic dependent launch.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace exl3b { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// One elected lane of a fully converged warp.  Keeping the surrounding control flow warp-uniform lets ptxas keep TMA /
// tcgen05 operands in uniform registers; issuing from inside an `if (lane == 0)` region instead costs a
// R2UR "waterfall" loop of ~16 dependent instructions per UTCHMMA (measured: 56 cycles per MMA issue).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of polling -- polling warps (epilogue / MMA / producer roles) were measured to consume ~40 % of all issue
// slots of the first version of the kernel (profiles/r01_ncu_tc_v1_notes.md).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
// non-blocking probe (diagnostics only)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  The bound is wall-clock
// (%globaltimer, checked every 1024 polls), not an iteration count: try_wait returns early whenever ANY barrier
// activity wakes the warp, so iteration counts say nothing about elapsed time.
// BACKOFF_NS > 0: exponential back-off (BACKOFF_NS, 2x, ... capped at 16x) between polls for roles whose waits can be
// long: polling warps otherwise burn issue slots the decode warps need (measured: ~40 % of all issued instructions of
// the first version of the kernel were polls; profiles/r01_ncu_notes.md).
template <int BACKOFF_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    unsigned long long t0 = 0;
    [[maybe_unused]] uint32_t ns = BACKOFF_NS;
    while (!mbar_try_wait(bar, parity))
    {
        if constexpr (BACKOFF_NS > 0) { __nanosleep(ns); if (ns < 16u * BACKOFF_NS) ns <<= 1; }
        if ((++polls & 1023u) == 0)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull)
            {
                printf("exl3b: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}

// ---- bulk async copy global -> shared (1-D), completion on an mbarrier ----------------------------------------
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// 2-D tiled TMA copy (SASS UTMALDG): box at element coordinates (c0 = innermost, c1 = row) of a CUtensorMap
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3}], [%4], %5;"
                 :: "r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap)
{
    asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}

// ---- device-side tensor-map patching (multi-matrix launches: one template, per-CTA global address) ----------------
// smem_tmap: 128-byte aligned shared-memory copy of the template; gmem_tmap: this CTA's 128-byte slot in global memory.
// Warp-collective (cp_fenceproxy is .sync.aligned).  After this the slot can be used by cp.async.bulk.tensor.
__device__ __forceinline__ void tmap_patch_address(uint32_t smem_tmap, void* gmem_tmap, uint64_t new_addr, int lane)
{
    if (lane == 0)
        asm volatile("tensormap.replace.tile.global_address.shared::cta.b1024.b64 [%0], %1;" :: "r"(smem_tmap), "l"(new_addr) : "memory");
    __syncwarp();
    asm volatile("tensormap.cp_fenceproxy.global.shared::cta.tensormap::generic.release.gpu.sync.aligned [%0], [%1], 128;"
                 :: "l"(gmem_tmap), "r"(smem_tmap) : "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(gmem_tmap) : "memory");
}

// The same for a matrix whose WIDTH differs from the template's: innermost extent (elements) and row pitch (bytes) too.
// (global_stride takes the byte value as of CUDA 12.5; older assemblers wanted it >> 4.)
__device__ __forceinline__ void tmap_patch_address_width(uint32_t smem_tmap, void* gmem_tmap, uint64_t new_addr, uint32_t dim0_elems,
                                                         uint64_t pitch_bytes, int lane)
{
    if (lane == 0)
    {
        asm volatile("tensormap.replace.tile.global_address.shared::cta.b1024.b64 [%0], %1;" :: "r"(smem_tmap), "l"(new_addr) : "memory");
        asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 0, %1;" :: "r"(smem_tmap), "r"(dim0_elems) : "memory");
        asm volatile("tensormap.replace.tile.global_stride.shared::cta.b1024.b64 [%0], 0, %1;" :: "r"(smem_tmap), "l"(pitch_bytes) : "memory");
    }
    __syncwarp();
    asm volatile("tensormap.cp_fenceproxy.global.shared::cta.tensormap::generic.release.gpu.sync.aligned [%0], [%1], 128;"
                 :: "l"(gmem_tmap), "r"(smem_tmap) : "memory");
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(gmem_tmap) : "memory");
}

// ---- relaxed gpu-scope accesses (flag-in-data exchange through L2) ------------------------------------------------------
__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu_u32(uint32_t* p, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// ---- relaxed system-scope vector accesses (flag-in-data exchange between GPUs over NVLink peer memory) -------------------
// A vector access is a set of 32-bit accesses, each of them single-copy atomic: every word is its own flag.
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(uint32_t* p, uint4 v)
{
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 8 consecutive 32-bit columns; thread i of the warp writes TMEM lane (warp%4)*32 + i
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc], kind::f16 (fp16 x fp16 -> fp32), one CTA
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::i8 (8-bit integers -> int32), one CTA
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same, the shared-memory descriptor given as its two 32-bit halves (lets the caller step the 14-bit start-address field
// with one 32-bit add per instruction instead of rebuilding the 64-bit descriptor)
__device__ __forceinline__ void mma_i8_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 bd, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], bd, %4, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// same, stepping the A address and the descriptor start address in place after the issue (keeps the whole unrolled MMA
// sequence on ONE uniform-register pair instead of one pair + one high-word move per instruction)
template <int A_STEP, int B_STEP>
__device__ __forceinline__ void mma_i8_ts_step(uint32_t d_tmem, uint32_t& a_tmem, uint32_t& b_desc_lo, uint32_t b_desc_hi, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 bd, {%1, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%2], [%0], bd, %4, p;\n\t"
                 "add.u32 %0, %0, %6;\n\t"
                 "add.u32 %1, %1, %7;\n\t}"
                 : "+r"(a_tmem), "+r"(b_desc_lo) : "r"(d_tmem), "r"(b_desc_hi), "r"(idesc), "r"(accumulate), "n"(A_STEP), "n"(B_STEP) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05 async ops of this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address, leading-dimension and
// stride-dimension byte offsets in 16-byte units, descriptor version 1 (Blackwell), layout type in bits 61..63
// (0 = no swizzle / interleaved 8x16B core matrices, 2 = 128-byte swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type)
{
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) (layout_type & 7) << 61;
    return d;
}

// Instruction descriptor for kind::f16 with fp16 A/B (format 0), fp32 accumulate (c_format 1), both K-major
// (cute::UMMA::InstrDescriptor): n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24.
__device__ __host__ __forceinline__ uint32_t idesc_f16_f32(int M, int N, bool b_mn_major = false)
{
    return (1u << 4) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

// kind::i8: A unsigned 8-bit (format 0), B signed 8-bit (format 1), int32 accumulate (c_format 2), both K-major
__device__ __host__ __forceinline__ uint32_t idesc_u8s8_s32(int M, int N)
{
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

}}  // namespace exl3b::ptx
