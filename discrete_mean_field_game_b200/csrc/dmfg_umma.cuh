// tcgen05 / TMEM plumbing for the reward-net kernels (sm_100a, inline PTX -- no CUTLASS types):
// TMEM allocation, shared-memory matrix descriptors of the un-swizzled ("interleaved") canonical layouts, the
// kind::tf32 instruction descriptor, single-thread MMA issue, commit to an mbarrier and the TMEM -> register load.
//
// Canonical un-swizzled operand layouts (units: bytes; T = 4 tf32 elements per 16 bytes), as the matrix descriptor
// addresses them (PTX ISA "shared memory matrix layout", restated in CuTe's make_umma_desc):
//   K-major  (operand stored [MN rows][K contiguous]):  core matrix = 8 MN-rows x 16 B, rows 16 B apart;
//            offset(mn, k) = (mn / 8) * SBO + (k / 4) * LBO + (mn % 8) * 16 + (k % 4) * 4
//   MN-major (operand stored [K rows][MN contiguous]):  core matrix = 8 K-rows x 16 B;
//            offset(mn, k) = (mn / 4) * SBO + (k / 8) * LBO + (k % 8) * 16 + (mn % 4) * 4
// One kind::tf32 instruction consumes K = 8 (two 16-byte chunks of a K-major operand).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmfg {
namespace umma {

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor of kind::tf32 with FP32 accumulation: D[M x N] (+)= A[M x 8] * B[8 x N]
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // D format: F32
         | (2u << 7) | (2u << 10)          // A, B format: TF32
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// one warp allocates `ncols` (power of two >= 32) TMEM columns; the base address lands in *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] = (accumulate ? D : 0) + A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed (and has finished
// reading its shared-memory operands); implies tcgen05.fence::before_thread_sync
__device__ __forceinline__ void commit(uint32_t mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}
// 32 TMEM lanes (this warp's quarter: lanes 32 (warp % 4) ..) x 8 consecutive 32-bit columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    // the loaded registers are operands of the wait, so nothing that reads them can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 TMEM lanes x 32 consecutive 32-bit columns <-> 32 registers per thread: a thread's private strip of tensor memory
// (lane = 32 (warp % 4) + lane id), used by the reward-net backward kernel to park its per-thread gradient accumulators
// between the phases that update them (TMEM as spill space: 2 instructions per phase instead of 46 live registers)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr),
                   "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                   "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
                   "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
                   "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// explicit shared-window store (32-bit address): the generic-pointer form makes the compiler rebuild the window base
// (S2UR SR_CgaCtaId + ULEA) in front of every access of an unrolled sequence
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

// x = hi + lo with hi exactly representable in TF32 (top 19 bits): the 3xTF32 split (hi*hi + lo*hi + hi*lo)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// ---------------------------------------------------------------------------------------------------
// dW3 += h^T . dz3 over a tile of 8 KG transitions on the tensor cores (3xTF32, FP32 accumulators resident in TMEM for
// the whole kernel):  D[m (MT x 128), j (16, n3 <= 8 live)] += A[m, n] * B[n, j],  K = n = transitions of the tile.
// Both operands K-MAJOR (measured on B200 with dmfg_umma_probe: K-major un-swizzled operands with LBO = stride of the
// 16-byte K chunks and SBO = stride of the 8-row groups reproduce A.B exactly; with the transpose bits set the same
// instruction returns zeros for kind::tf32, so MN-major operands are not used):
//   A[m][n]  (row m = one activation slot, K = transition):  offA(m, n) = (m / 8) * SBO + (n / 4) * 128 + (m % 8) * 16 + (n % 4) * 4
//   B[j][n]  (row j = fc3 unit,            K = transition):  offB(j, n) = (j / 8) * SBO + (n / 4) * 128 + (j % 8) * 16 + (n % 4) * 4
// with SBO = 256 KG (the 2 KG chunks of a row group are adjacent).  The order of the M rows is free -- the caller maps
// activation index -> m so that its stores need no address arithmetic, and un-permutes when it reads D back.
// ---------------------------------------------------------------------------------------------------
// HS ("half split", 16-row activation slots): the staging stores of a warp go to rows (t, h) of ONE K column per
// transition, and rows h and h + 8 of a slot fall on the same banks (bank = 4 (m % 8) + n % 4) -- a 2-way conflict on every
// store.  With HS the rows of the upper half (h >= 8) live in the upper half of the M tiles and use K column n ^ 2, i.e. the
// two free banks of each quad; their MMAs read a second copy of B with the columns permuted the same way (b2).
template <int MT, int KG, bool HS = false>
struct W3Grad {
    static constexpr uint32_t kLbo = 128, kSbo = 256u * KG;
    static constexpr uint32_t kBytesA = (uint32_t)MT * 16u * kSbo, kBytesB = (HS ? 4u : 2u) * kSbo;  // one of (hi, lo); HS: B then B permuted
    static constexpr bool kHalfSplit = HS;
    static constexpr uint32_t kTmemCols = MT * 16 <= 32 ? 32 : MT * 16 <= 64 ? 64 : MT * 16 <= 128 ? 128 : 256;      // MT M-tiles x N = 16
    static __device__ __forceinline__ uint32_t off(int row, int n) {
        return (uint32_t)(row >> 3) * kSbo + (uint32_t)(n >> 2) * kLbo + (uint32_t)(row & 7) * 16u + (uint32_t)(n & 3) * 4u;
    }
    // ONE thread: the MT x KG x 3 MMAs of a tile (hi*hi + lo*hi + hi*lo per M-tile and K-step), then the commit
    static __device__ __forceinline__ void issue(uint32_t tmem_base, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                 uint32_t b_lo, bool first, uint32_t mbar_addr) {
        constexpr uint32_t idesc = idesc_tf32(128, 16, false, false);
        fence_after_sync();
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int ks = 0; ks < KG; ++ks) {
                const uint32_t ao = (uint32_t)m * 16u * kSbo + (uint32_t)ks * 2u * kLbo, bo = (uint32_t)ks * 2u * kLbo;
                const uint64_t ah = smem_desc(a_hi + ao, kLbo, kSbo), al = smem_desc(a_lo + ao, kLbo, kSbo);
                const uint32_t bsel = (HS && m >= MT / 2) ? 2u * kSbo : 0u;      // upper-half tiles: the permuted copy of B
                const uint64_t bh = smem_desc(b_hi + bsel + bo, kLbo, kSbo), bl = smem_desc(b_lo + bsel + bo, kLbo, kSbo);
                const uint32_t d = tmem_base + 16u * (uint32_t)m;
                mma_tf32(d, ah, bh, idesc, (first && ks == 0) ? 0u : 1u);
                mma_tf32(d, al, bh, idesc, 1u);
                mma_tf32(d, ah, bl, idesc, 1u);
            }
        }
        commit(mbar_addr);
    }
};

}  // namespace umma
}  // namespace dmfg
