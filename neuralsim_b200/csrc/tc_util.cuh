// Blackwell tensor-core plumbing used by the fused MLP kernels: tcgen05.mma with shared-memory operands,
// accumulators in TMEM, completion through an mbarrier.  Inline PTX only (no CUTLASS dependency).
//
// Shared-memory operand layout ("canonical K-major, no swizzle", what cute calls Layout_K_INTER_Atom):
//   the tile is cut into core matrices of 8 rows x 16 bytes (8 fp16 along K); a core matrix is 128 contiguous
//   bytes (row r at r*16).  SBO = byte distance between core matrices that are adjacent along M/N (next 8 rows),
//   LBO = byte distance between core matrices that are adjacent along K (next 8 k).
//   We store a [R x K] fp16 tile chunk-major:   byte(r, k) = (k/8) * (R*16) + r*16 + (k%8)*2
//   => SBO = 128, LBO = R*16.  Thread r owns row r and writes K/8 16-byte vectors with a stride of R*16 bytes:
//   consecutive threads hit consecutive 16-byte slots, i.e. conflict-free st.shared.v4.
//   The very same bytes read as an MN-major operand (rows <-> K) are described by swapping the two offsets
//   (LBO = 128, SBO = R*16) -- used by the weight-gradient MMA that contracts over the points of a tile.
#pragma once
#include "nsb_common.cuh"

namespace nsb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (PTX "tcgen05 matrix descriptor"; fields as in cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) swizzle = 0 (none)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32.
//   [4,6) D format 1 = F32 | [7,10) A format 0 = F16 | [10,13) B format 0 = F16 | [15] A major | [16] B major
//   (0 = K-major, 1 = MN-major) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- TMA (bulk async copy engine): one thread moves `bytes` contiguous bytes global -> shared; completion is counted on an mbarrier.
//      bytes % 16 == 0, both addresses 16-byte aligned.  The issuing thread first announces the byte count (arrive.expect_tx).
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp allocates / frees `cols` TMEM columns (power of two >= 32); the base address lands in *slot (shared)
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane 32*(warp%4)+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM -> registers, 8 consecutive fp32 columns (used inside rolled loops: small code, dynamic column address)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8_f16(const float (&v)[8]) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    uint4 q;
    q.x = *reinterpret_cast<uint32_t *>(&h0); q.y = *reinterpret_cast<uint32_t *>(&h1);
    q.z = *reinterpret_cast<uint32_t *>(&h2); q.w = *reinterpret_cast<uint32_t *>(&h3);
    return q;
}

// write one row of a chunk-major [R x K] fp16 tile: `vals` are K fp32 values rounded to fp16 here
template <int R, int K>
__device__ __forceinline__ void store_row_f16(uint8_t *tile, int r, const float *vals) {
#pragma unroll
    for (int c = 0; c < K / 8; ++c) {
        __half2 h0 = __floats2half2_rn(vals[c * 8 + 0], vals[c * 8 + 1]);
        __half2 h1 = __floats2half2_rn(vals[c * 8 + 2], vals[c * 8 + 3]);
        __half2 h2 = __floats2half2_rn(vals[c * 8 + 4], vals[c * 8 + 5]);
        __half2 h3 = __floats2half2_rn(vals[c * 8 + 6], vals[c * 8 + 7]);
        uint4 q;
        q.x = *reinterpret_cast<uint32_t *>(&h0);
        q.y = *reinterpret_cast<uint32_t *>(&h1);
        q.z = *reinterpret_cast<uint32_t *>(&h2);
        q.w = *reinterpret_cast<uint32_t *>(&h3);
        *reinterpret_cast<uint4 *>(tile + c * (R * 16) + r * 16) = q;
    }
}

}  // namespace tc
}  // namespace nsb
