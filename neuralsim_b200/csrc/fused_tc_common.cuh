// Building blocks shared by the tcgen05 fused kernels (fused_tc.cu, color_tc.cu): tile constants, the per-thread
// LoTD gather that writes straight into a core-matrix A tile, point loading and W1 staging.
#pragma once
#include "lotd_device.cuh"
#include "tc_util.cuh"

namespace nsb {

struct DecoderDevTC {
    const __half *W1, *b1, *W2, *b2;
    int width;
    float beta;
};

constexpr int kTile = 128;
constexpr int NF = 32, HW = 64;           // features, hidden width (zero padded to 64)

// the 16 levels of one point -> row `r` of a chunk-major [R x >=32] fp16 tile (4 bytes per level)
template <int R>
__device__ __forceinline__ void gather_row_to_tile(const PLMeta &m, const __half *__restrict__ grid, const float (&xs)[3],
                                                   int max_level, uint8_t *tile, int r) {
#pragma unroll 1
    for (uint32_t p = 0; p < 16; ++p) {
        uint32_t packed = 0;
        if ((int)m.level[p] <= max_level) {
            uint32_t idx[8];
            float w[8];
            level_corners3(m, p, xs, idx, w);
            packed = level_feat2(grid, idx, w);
        }
        *reinterpret_cast<uint32_t *>(tile + (p >> 2) * (R * 16) + r * 16 + (p & 3) * 4) = packed;
    }
}

__device__ __forceinline__ void load_point(bool from_rays, const float *__restrict__ x, const float *__restrict__ rays_o,
                                           const float *__restrict__ rays_d, const int64_t *__restrict__ ridx,
                                           const float *__restrict__ t, int64_t i, bool valid, float (&xs)[3]) {
    xs[0] = xs[1] = xs[2] = 0.f;
    if (valid) {
        if (from_rays) {
            const int64_t r = ridx ? ridx[i] : i;
            const float tt = t[i];
#pragma unroll
            for (int d = 0; d < 3; ++d) xs[d] = __fmaf_rn(rays_d[r * 3 + d], tt, rays_o[r * 3 + d]);
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) xs[d] = x[i * 3 + d];
        }
    }
    // network space [-1,1] -> table space [0,1] (lotd_encoding.py:165), clamp (lotd.py:60)
#pragma unroll
    for (int d = 0; d < 3; ++d) xs[d] = fminf(fmaxf(__fmaf_rn(xs[d], 0.5f, 0.5f), 1.0e-6f), 1.f - 1.0e-6f);
}

// W1 [width x 32] (fp16, row-major) -> chunk-major [64 x 32] B tile, rows >= width zero
__device__ __forceinline__ void stage_W1(const DecoderDevTC &dec, uint8_t *sB, int tid) {
    for (int e = tid; e < HW * (NF / 8); e += kTile) {
        const int j = e % HW, c = e / HW;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (j < dec.width) q = *reinterpret_cast<const uint4 *>(dec.W1 + j * NF + c * 8);
        *reinterpret_cast<uint4 *>(sB + c * (HW * 16) + j * 16) = q;
    }
}

}  // namespace nsb
