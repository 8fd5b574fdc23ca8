// Device-side building blocks of the LoTD encoding shared by lotd.cu and the fused kernels.
// Index / hash arithmetic follows /root/reference/nr3d_lib/csrc/lotd/include/lotd/lotd_cuda.h:92-143,
// position arithmetic lotd_cuda.h:959-984 (InterpolationType::Linear).
#pragma once
#include <type_traits>

#include "nsb_common.cuh"

namespace nsb {

constexpr int kMaxPseudo = 32;

// Per-pseudo-level view of nsb_lotd_meta, passed to kernels by value (~1.6 KB of kernel parameters).
struct PLMeta {
    uint32_t n_pseudo, n_out, D, F;
    uint32_t level[kMaxPseudo];     // actual level of the pseudo level (for max_level masking)
    uint32_t res[kMaxPseudo][NSB_MAX_DIMS];
    uint32_t base[kMaxPseudo];      // element offset of (level, pseudo slot) inside the flat table
    uint32_t nfeat[kMaxPseudo];     // feature width of the actual level (row stride of a cell)
    uint32_t size[kMaxPseudo];      // cells of the level (hash modulus)
    uint32_t is_hash;               // bit p set -> hashed level
};

inline int make_plmeta(const nsb_lotd_meta *m, PLMeta *o) {
    if (m->n_pseudo_levels > (uint32_t)kMaxPseudo) {
        set_error("LoTD: %u pseudo levels exceed the built maximum %d", m->n_pseudo_levels, kMaxPseudo);
        return 2;
    }
    memset(o, 0, sizeof(*o));
    o->n_pseudo = m->n_pseudo_levels;
    o->n_out = m->n_encoded_dims;
    o->D = m->n_dims_to_encode;
    o->F = m->n_feat_per_pseudo_lvl;
    for (uint32_t p = 0; p < m->n_pseudo_levels; ++p) {
        const uint32_t l = m->map_levels[p];
        o->level[p] = l;
        for (int d = 0; d < NSB_MAX_DIMS; ++d) o->res[p][d] = m->level_res[l][d];
        o->base[p] = m->level_offsets[l] + m->map_cnt[p] * m->n_feat_per_pseudo_lvl;
        o->nfeat[p] = m->level_n_feats[l];
        o->size[p] = m->level_sizes[l];
        if (m->level_types[l] == NSB_LOD_HASH) {
            if (p >= 32) { set_error("LoTD: hashed pseudo level index >= 32 unsupported"); return 2; }
            o->is_hash |= (1u << p);
        }
    }
    return 0;
}

template <bool HALF> struct ValT { using type = float; };
template <> struct ValT<true> { using type = __half; };

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <typename T> __device__ __forceinline__ T add_st(T a, T b);
template <> __device__ __forceinline__ float add_st<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ __half add_st<__half>(__half a, __half b) { return __hadd(a, b); }

// cell = floor(x*scale+0.5), frac = the remainder; scale = res-2 (lotd_hash_only.h:67).  One FMA, as nvcc emits
// for the reference expression `positions[dim] * scale[dim] + 0.5f`.
template <int D>
__device__ __forceinline__ void level_pos(const PLMeta &m, uint32_t p, const float (&xs)[D], uint32_t (&cell)[D],
                                          float (&fr)[D], float (&scale)[D]) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
        scale[d] = (float)(m.res[p][d] - 2u);
        const float v = __fmaf_rn(xs[d], scale[d], 0.5f);
        const float fl = floorf(v);
        cell[d] = (uint32_t)fl;
        fr[d] = v - fl;
    }
}

// weight of corner c: prod_d (bit d of c ? frac_d : 1-frac_d), multiplied in dimension order.
template <int D>
__device__ __forceinline__ float corner_weight(const float (&fr)[D], int c) {
    float w = (c & 1) ? fr[0] : __fsub_rn(1.f, fr[0]);
#pragma unroll
    for (int d = 1; d < D; ++d) w = __fmul_rn(w, (c & (1 << d)) ? fr[d] : __fsub_rn(1.f, fr[d]));
    return w;
}

template <int D>
__device__ __forceinline__ uint32_t corner_index(const PLMeta &m, uint32_t p, const uint32_t (&cell)[D], int c) {
    uint32_t idx;
    if (m.is_hash & (1u << p)) {
        constexpr uint32_t primes[4] = {1u, 2654435761u, 805459861u, 3674653429u};
        idx = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) idx ^= (cell[d] + ((c >> d) & 1)) * primes[d];
        idx %= m.size[p];
    } else {
        idx = 0;
        uint32_t stride = 1;
#pragma unroll
        for (int d = D - 1; d >= 0; --d) {  // last dimension contiguous
            idx += (cell[d] + ((c >> d) & 1)) * stride;
            stride *= m.res[p][d];
        }
    }
    return idx * m.nfeat[p] + m.base[p];
}

template <int D, int F, typename VT>
__device__ __forceinline__ void load_corner(const PLMeta &m, uint32_t p, const VT *__restrict__ grid,
                                            const uint32_t (&cell)[D], int c, VT (&out)[F]) {
    const uint32_t e = corner_index<D>(m, p, cell, c);
    if constexpr (std::is_same<VT, __half>::value) {
#pragma unroll
        for (int f = 0; f < F; f += 2) {
            const uint32_t raw = ld_nc_u32(grid + e + f);
            out[f] = __ushort_as_half((unsigned short)(raw & 0xffffu));
            out[f + 1] = __ushort_as_half((unsigned short)(raw >> 16));
        }
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) out[f] = __ldg(grid + e + f);
    }
}

// grad[dst .. dst+F) += g[f] * w, fp32, fire-and-forget vector reductions (8-byte aligned: F even, offsets even).
template <int F>
__device__ __forceinline__ void red_add(float *dst, const float (&g)[F], float w) {
#pragma unroll
    for (int f = 0; f < F; f += 2) {
        const float a = g[f] * w, b = g[f + 1] * w;
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + f), "f"(a), "f"(b) : "memory");
    }
}

}  // namespace nsb
