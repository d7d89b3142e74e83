// tcgen05 / mbarrier / bulk-copy PTX wrappers and the digit-plane constants shared by the stand-alone product
// (ozaki.cu) and the batched GP phases (gp_ozaki.cuh).  sm_100a only.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace oz {

constexpr int OZ_MAXS = 8;      // 8 accumulators of 64 columns = the 512 TMEM columns of an SM
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64;
constexpr int OZ_BLK = 64 * 64;                 // bytes of one (64 vectors x 64 k) block of a digit plane
constexpr int OZ_THREADS = 192;

__host__ __device__ constexpr int oz_stage_bytes(int S) { return S * (2 * OZ_BLK + OZ_BLK); }
__host__ __device__ constexpr int oz_stages(int S) { return S <= 6 ? 3 : 2; }      // 227 KB of shared memory per CTA
__host__ __device__ constexpr int oz_smem_bytes(int S) { return oz_stages(S) * oz_stage_bytes(S) + 1024 + 256; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 64-byte swizzle: rows of 64 bytes, 8-row atoms of 512 bytes (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::i8: D = S32 (c_format 2), A and B signed 8 bit (format 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);


// one (vector, 16-k chunk) of a staged 64 x 64 block -> 16 bytes of each of the S digit planes, written at the
// 64-byte-swizzled position of the K-major shared-memory layout.  y: the 16 values already scaled to |y| < 64.
template <int S>
__device__ __forceinline__ void emit_digits(double (&y)[16], uint8_t* dst, size_t plane) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
        uint32_t w[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = 4 * p + b;
                const double q = rint(y[j]);               // |q| <= 64
                y[j] = (y[j] - q) * 128.0;                 // exact: the next 7 bits, |.| <= 64
                word |= ((uint32_t)(int)q & 0xffu) << (8 * b);
            }
            w[p] = word;
        }
        *reinterpret_cast<uint4*>(dst + s * plane) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace oz
