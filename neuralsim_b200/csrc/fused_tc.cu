// Fused forward_sdf with the decoder's 32 -> 64 layer on the 5th-generation tensor cores (tcgen05), sm_100a.
//
// One CTA = 128 threads = one tile of 128 points; thread r owns point r for the whole pipeline:
//   gather   thread r walks the 16 LoTD levels of its point (8 corner loads each) and keeps the 32 fp16 features in registers
//   stage    thread r writes its feature row into the A tile in shared memory (core-matrix layout, conflict-free 16-byte stores)
//   MMA      one elected thread issues two tcgen05.mma (M=128, N=64, K=16 each): D[128 x 64] (fp32, TMEM) = A[128 x 32] . W1^T
//   epilogue thread r reads TMEM lane r (its own point's 64 pre-activations) with tcgen05.ld, applies bias + Softplus(beta),
//            rounds, and does the 64 -> 1 layer as a register dot product -> sdf[r]
// so no cross-thread data exchange is needed besides the MMA itself: TMEM lane == point == thread.
// W1 is staged once per (persistent) CTA.  Numerics: same fp16 rounding points as fused.cu / the autocast graph.
#include "lotd_device.cuh"
#include "tc_util.cuh"

namespace nsb {

struct DecoderDevTC {
    const __half *W1, *b1, *W2, *b2;
    int width;
    float beta;
};

constexpr int kTile = 128;

template <int D, int F, int NP>
__device__ __forceinline__ void gather_row_unrolled(const PLMeta &m, const __half *__restrict__ grid, const float (&xs)[D],
                                                    int max_level, float (&h)[NP * F]) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if ((int)m.level[p] > max_level) {
#pragma unroll
            for (int f = 0; f < F; ++f) h[p * F + f] = 0.f;
            continue;
        }
        uint32_t cell[D];
        float fr[D], scale[D];
        level_pos<D>(m, p, xs, cell, fr, scale);
        __half v[1 << D][F];
#pragma unroll
        for (int c = 0; c < (1 << D); ++c) load_corner<D, F, __half>(m, p, grid, cell, c, v[c]);
        __half acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = __float2half_rn(0.f);
#pragma unroll
        for (int c = 0; c < (1 << D); ++c) {
            const float w = corner_weight<D>(fr, c);
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] = __hadd(acc[f], __float2half_rn(__fmul_rn(w, __half2float(v[c][f]))));
        }
#pragma unroll
        for (int f = 0; f < F; ++f) h[p * F + f] = __half2float(acc[f]);
    }
}

__device__ __forceinline__ float softplus_b(float z, float beta) {
    const float zb = z * beta;
    return zb > 20.f ? z : log1pf(expf(zb)) / beta;
}

template <bool FROM_RAYS>
__global__ void __launch_bounds__(kTile)
k_fused_sdf_tc(const PLMeta m, const __half *__restrict__ grid, const DecoderDevTC dec, const float *__restrict__ x,
               const float *__restrict__ rays_o, const float *__restrict__ rays_d, const int64_t *__restrict__ ridx,
               const float *__restrict__ t, int64_t n, int max_level, float *__restrict__ sdf) {
    constexpr int NF = 32, W = 64;
    __shared__ __align__(1024) uint8_t sA[kTile * NF * 2];   // 8 KB : features, chunk-major core-matrix layout
    __shared__ __align__(1024) uint8_t sB[W * NF * 2];       // 4 KB : W1 [64 x 32], same layout
    __shared__ float sb1[W], sW2[W];
    __shared__ float sb2;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- one-time setup: weights -> smem (zero-padded up to 64 hidden units), TMEM allocation, mbarrier
    for (int e = tid; e < W * (NF / 8); e += kTile) {        // e = (row j, chunk c)
        const int j = e % W, c = e / W;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (j < dec.width) q = *reinterpret_cast<const uint4 *>(dec.W1 + j * NF + c * 8);
        *reinterpret_cast<uint4 *>(sB + c * (W * 16) + j * 16) = q;
    }
    if (tid < W) {
        sb1[tid] = tid < dec.width ? __half2float(dec.b1[tid]) : 0.f;
        sW2[tid] = tid < dec.width ? __half2float(dec.W2[tid]) : 0.f;
    }
    if (tid == 0) {
        sb2 = __half2float(dec.b2[0]);
        tc::mbar_init(&mbar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc(kTile, W, 0, 0);
    const uint32_t a_addr = tc::smem_u32(sA), b_addr = tc::smem_u32(sB);
    uint32_t phase = 0;

    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * kTile + tid;
        const bool valid = i < n;
        float xs[3] = {0.f, 0.f, 0.f};
        if (valid) {
            if (FROM_RAYS) {
                const int64_t r = ridx ? ridx[i] : i;
                const float tt = t[i];
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[d] = __fmaf_rn(rays_d[r * 3 + d], tt, rays_o[r * 3 + d]);
            } else {
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[d] = x[i * 3 + d];
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) xs[d] = fminf(fmaxf(__fmaf_rn(xs[d], 0.5f, 0.5f), 1.0e-6f), 1.f - 1.0e-6f);
        float h[NF];
        gather_row_unrolled<3, 2, 16>(m, grid, xs, max_level, h);
        tc::store_row_f16<kTile, NF>(sA, tid, h);
        tc::fence_async_smem();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < NF / 16; ++ks) {
                const uint64_t ad = tc::make_desc(a_addr + ks * 2 * (kTile * 16), kTile * 16, 128);
                const uint64_t bd = tc::make_desc(b_addr + ks * 2 * (W * 16), W * 16, 128);
                tc::mma_f16_ss(tmem, ad, bd, idesc, ks > 0);
            }
            tc::commit(&mbar);
        }
        tc::mbar_wait(&mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        // ---- epilogue: my own row of D
        float out = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float z[32];
            tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, z);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = half * 32 + j;
                const float zz = __half2float(__float2half_rn(z[j] + sb1[col]));
                const float a = __half2float(__float2half_rn(softplus_b(zz, dec.beta)));
                out = fmaf(a, sW2[col], out);
            }
        }
        if (valid) sdf[i] = __half2float(__float2half_rn(out + sb2));
        tc::fence_before_sync();           // TMEM reads done before the next tile's MMA overwrites D
        __syncthreads();
    }
    if (warp == 0) tc::tmem_free<64>(tmem);
}

}  // namespace nsb

using namespace nsb;

extern "C" int nsb_fused_sdf_tc_launch(const nsb_lotd_meta *meta, const void *params_half, const nsb_sdf_decoder *dec, const float *x,
                                       const float *rays_o, const float *rays_d, const int64_t *ridx, const float *t, int64_t n,
                                       int32_t max_level, float *sdf, void *stream, int from_rays) {
    PLMeta m;
    if (make_plmeta(meta, &m)) return 2;
    NSB_REQUIRE(m.n_pseudo == 16 && m.F == 2 && m.D == 3, "nsb_fused_sdf (tensor-core): built for 16 x 2 LoTD features in 3-D");
    NSB_REQUIRE(dec->width >= 1 && dec->width <= 64, "nsb_fused_sdf (tensor-core): decoder width must be <= 64");
    DecoderDevTC d{(const __half *)dec->W1, (const __half *)dec->b1, (const __half *)dec->W2, (const __half *)dec->b2, dec->width,
                   dec->beta};
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t wave = (int64_t)sm_count() * 8;          // <= 8 resident CTAs/SM (TMEM: 8 x 64 columns)
    const unsigned grid = (unsigned)(n_tiles < wave ? n_tiles : wave);
    cudaStream_t s = (cudaStream_t)stream;
    const int ml = max_level < 0 ? -1 : max_level;
    if (from_rays) k_fused_sdf_tc<true><<<grid, kTile, 0, s>>>(m, (const __half *)params_half, d, nullptr, rays_o, rays_d, ridx, t, n, ml, sdf);
    else k_fused_sdf_tc<false><<<grid, kTile, 0, s>>>(m, (const __half *)params_half, d, x, nullptr, nullptr, nullptr, nullptr, n, ml, sdf);
    return check_launch("nsb_fused_sdf(tc)");
}

// =====================================================================================================================
// Backward of forward_sdf wrt. the LoTD table and the decoder weights, one kernel, nothing saved by the forward.
//
// Per tile of 128 points (thread r = point r = TMEM lane r), with d = dL/dsdf[r]:
//   recompute  h (gather) -> A tile;  MMA1: Z = H.W1^T (TMEM cols 0..63);  z = fp16(Z + b1), s = sigmoid(beta z), a = fp16(softplus)
//   dz_j = d * w2_j * s_j                      -> fp16 row of the dZ tile (the reference rounds grad_z to fp16 at the same place)
//   MMA2: dH[128 x 32] = dZ . W1               (A = dZ tile K-major, B = W1^T tile; TMEM cols 0..31, Z is dead by then)
//   scatter dH into the fp32 table gradient    (8 corners x 16 levels, red.global.add.v2.f32)
//   MMA3: dW1x[64 x 40] += dZ^T . [H | 1 | 0]  (both operands are the tiles above read MN-major, K = the 128 points;
//                                               accumulated in TMEM cols 64..103 over all tiles of the persistent CTA;
//                                               column 32 of [H | 1] is a constant one -> db1 for free)
//   dW2_j += sum_r d a_j, db2 += sum_r d       (warp reduce-scatter with shuffles -> shared accumulators)
// The CTA flushes dW1x / dW2 / db2 with one atomicAdd per element when it runs out of tiles.
// =====================================================================================================================
namespace nsb {

// reduce-scatter over the 32 lanes of a warp: on return lane l holds sum over lanes of v[2l], v[2l+1] in (v[0], v[1])
__device__ __forceinline__ void warp_reduce_scatter64(float (&v)[64], int lane) {
#pragma unroll
    for (int half = 32, bit = 16; half >= 2; half >>= 1, bit >>= 1) {
        const bool upper = lane & bit;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k < half) {
                const float send = upper ? v[k] : v[k + half];
                const float keep = upper ? v[k + half] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
}

template <bool FROM_RAYS>
__global__ void __launch_bounds__(kTile)
k_sdf_bwd_tc(const PLMeta m, const __half *__restrict__ grid, const DecoderDevTC dec, const float *__restrict__ x,
             const float *__restrict__ rays_o, const float *__restrict__ rays_d, const int64_t *__restrict__ ridx,
             const float *__restrict__ t, const float *__restrict__ d_sdf, int64_t n, int max_level, float *__restrict__ d_grid,
             float *__restrict__ d_W1, float *__restrict__ d_b1, float *__restrict__ d_W2, float *__restrict__ d_b2) {
    constexpr int NF = 32, W = 64, NX = 40;                   // NX: features + [1,0,..] chunk
    __shared__ __align__(1024) uint8_t sA[kTile * NX * 2];    // 10 KB : [H | 1 | 0..] rows, chunk-major
    __shared__ __align__(1024) uint8_t sDZ[kTile * W * 2];    // 16 KB : dZ rows, chunk-major
    __shared__ __align__(1024) uint8_t sB[W * NF * 2];        //  4 KB : W1   [64 x 32]  (B of MMA1)
    __shared__ __align__(1024) uint8_t sBT[NF * W * 2];       //  4 KB : W1^T [32 x 64]  (B of MMA2)
    __shared__ float sb1[W], sW2[W], sdW2[W];
    __shared__ float sdb2;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < W * (NF / 8); e += kTile) {
        const int j = e % W, c = e / W;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (j < dec.width) q = *reinterpret_cast<const uint4 *>(dec.W1 + j * NF + c * 8);
        *reinterpret_cast<uint4 *>(sB + c * (W * 16) + j * 16) = q;
    }
    for (int e = tid; e < NF * W; e += kTile) {               // W1^T: row = feature k, col = hidden j
        const int k = e % NF, j = e / NF;
        const __half v = j < dec.width ? dec.W1[j * NF + k] : __float2half_rn(0.f);
        *reinterpret_cast<__half *>(sBT + (j / 8) * (NF * 16) + k * 16 + (j % 8) * 2) = v;
    }
    if (tid < W) {
        sb1[tid] = tid < dec.width ? __half2float(dec.b1[tid]) : 0.f;
        sW2[tid] = tid < dec.width ? __half2float(dec.W2[tid]) : 0.f;
        sdW2[tid] = 0.f;
    }
    {   // constant chunk 4 of my A row: [1, 0, 0, 0, 0, 0, 0, 0]
        uint4 q = make_uint4(0x00003C00u, 0, 0, 0);           // fp16 1.0 = 0x3C00
        *reinterpret_cast<uint4 *>(sA + 4 * (kTile * 16) + tid * 16) = q;
    }
    if (tid == 0) {
        sdb2 = 0.f;
        tc::mbar_init(&mbar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_slot);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc1 = tc::make_idesc(kTile, W, 0, 0);    // Z    = H  . W1^T
    const uint32_t idesc2 = tc::make_idesc(kTile, NF, 0, 0);   // dH   = dZ . W1
    const uint32_t idesc3 = tc::make_idesc(W, NX, 1, 1);       // dW1x = dZ^T . [H|1]   (MN-major operands)
    const uint32_t a_addr = tc::smem_u32(sA), dz_addr = tc::smem_u32(sDZ), b_addr = tc::smem_u32(sB), bt_addr = tc::smem_u32(sBT);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;
    bool first_tile = true;

    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * kTile + tid;
        const bool valid = i < n;
        float xs[3] = {0.f, 0.f, 0.f};
        if (valid) {
            if (FROM_RAYS) {
                const int64_t r = ridx ? ridx[i] : i;
                const float tt = t[i];
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[d] = __fmaf_rn(rays_d[r * 3 + d], tt, rays_o[r * 3 + d]);
            } else {
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[d] = x[i * 3 + d];
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) xs[d] = fminf(fmaxf(__fmaf_rn(xs[d], 0.5f, 0.5f), 1.0e-6f), 1.f - 1.0e-6f);
        const float dd = valid ? d_sdf[i] : 0.f;
        {
            float h[NF];
            gather_row_unrolled<3, 2, 16>(m, grid, xs, max_level, h);
            tc::store_row_f16<kTile, NF>(sA, tid, h);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < NF / 16; ++ks)
                tc::mma_f16_ss(tmem, tc::make_desc(a_addr + ks * 2 * (kTile * 16), kTile * 16, 128),
                               tc::make_desc(b_addr + ks * 2 * (W * 16), W * 16, 128), idesc1, ks > 0);
            tc::commit(&mbar);
        }
        tc::mbar_wait(&mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        // ---- activations, dz row, weight-2 partials
        float da[W];                                             // d * a_j, reduced over the warp below
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float z[32], dz[32];
            tc::tmem_ld32(tmem + lane_base + half * 32, z);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = half * 32 + j;
                const float zz = __half2float(__float2half_rn(z[j] + sb1[col]));
                const float zb = zz * dec.beta;
                float a, s;
                if (zb > 20.f) { a = zz; s = 1.f; }
                else { const float e = expf(zb); a = log1pf(e) / dec.beta; s = e / (e + 1.f); }
                a = __half2float(__float2half_rn(a));
                da[col] = dd * a;
                dz[j] = dd * sW2[col] * s;
            }
            // my dZ row, 4 chunks of this half
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                __half2 h0 = __floats2half2_rn(dz[c * 8 + 0], dz[c * 8 + 1]), h1 = __floats2half2_rn(dz[c * 8 + 2], dz[c * 8 + 3]);
                __half2 h2 = __floats2half2_rn(dz[c * 8 + 4], dz[c * 8 + 5]), h3 = __floats2half2_rn(dz[c * 8 + 6], dz[c * 8 + 7]);
                uint4 q;
                q.x = *reinterpret_cast<uint32_t *>(&h0); q.y = *reinterpret_cast<uint32_t *>(&h1);
                q.z = *reinterpret_cast<uint32_t *>(&h2); q.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(sDZ + (half * 4 + c) * (kTile * 16) + tid * 16) = q;
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();                                         // every thread has read Z and written its dZ row
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < W / 16; ++ks)                  // dH = dZ . W1 : K = 64 hidden
                tc::mma_f16_ss(tmem, tc::make_desc(dz_addr + ks * 2 * (kTile * 16), kTile * 16, 128),
                               tc::make_desc(bt_addr + ks * 2 * (NF * 16), NF * 16, 128), idesc2, ks > 0);
#pragma unroll
            for (int ks = 0; ks < kTile / 16; ++ks)              // dW1x += dZ^T . [H|1] : K = 128 points, MN-major operands
                tc::mma_f16_ss(tmem + 64, tc::make_desc(dz_addr + ks * 256, 128, kTile * 16),
                               tc::make_desc(a_addr + ks * 256, 128, kTile * 16), idesc3, (ks > 0) || !first_tile);
            tc::commit(&mbar);
        }
        first_tile = false;
        // overlap: dW2 / db2 partial sums while the tensor core works
        warp_reduce_scatter64(da, lane);
        atomicAdd(&sdW2[2 * lane], da[0]);
        atomicAdd(&sdW2[2 * lane + 1], da[1]);
        const float dsum = warp_sum(dd);
        if (lane == 0) atomicAdd(&sdb2, dsum);
        tc::mbar_wait(&mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        // ---- my dH row -> scatter into the table gradient
        float dh[NF];
        tc::tmem_ld32(tmem + lane_base, dh);
        if (valid && dd != 0.f) {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                if ((int)m.level[p] > max_level) continue;
                uint32_t cell[3];
                float fr[3], sc[3];
                level_pos<3>(m, p, xs, cell, fr, sc);
                const float g[2] = {dh[2 * p], dh[2 * p + 1]};
#pragma unroll
                for (int c = 0; c < 8; ++c) red_add<2>(d_grid + corner_index<3>(m, p, cell, c), g, corner_weight<3>(fr, c));
            }
        }
        tc::fence_before_sync();
        __syncthreads();                                         // tiles + TMEM free for the next iteration
    }
    // ---- flush the CTA's weight-gradient accumulators
    if (!first_tile) {
        tc::fence_after_sync();
        float r0[32], r1[32];
        tc::tmem_ld32(tmem + 64 + lane_base, r0);                // M = 64 accumulator: warp w, lanes 0..15 <-> rows 16w..16w+15
        tc::tmem_ld32(tmem + 96 + lane_base, r1);                // columns 32..39 live in r1[0..7]
        if (lane < 16) {
            const int j = warp * 16 + lane;
            if (j < dec.width) {
#pragma unroll
                for (int k = 0; k < NF; ++k) atomicAdd(d_W1 + j * NF + k, r0[k]);
                atomicAdd(d_b1 + j, r1[0]);
            }
        }
        __syncthreads();
        if (tid < dec.width) atomicAdd(d_W2 + tid, sdW2[tid]);
        if (tid == 0) atomicAdd(d_b2, sdb2);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free<128>(tmem);
}

}  // namespace nsb

extern "C" int nsb_fused_sdf_bwd(const nsb_lotd_meta *meta, const void *params_half, const nsb_sdf_decoder *dec, const float *x,
                                 const float *rays_o, const float *rays_d, const int64_t *ridx, const float *t, const float *d_sdf,
                                 int64_t n, int32_t max_level, float *d_grid, float *d_W1, float *d_b1, float *d_W2, float *d_b2,
                                 void *stream) {
    if (n == 0) return 0;
    NSB_REQUIRE(meta && params_half && dec && d_sdf && d_grid && d_W1 && d_b1 && d_W2 && d_b2, "nsb_fused_sdf_bwd: NULL argument");
    NSB_REQUIRE(x || (rays_o && rays_d && t), "nsb_fused_sdf_bwd: need x or (rays_o, rays_d, t)");
    PLMeta m;
    if (make_plmeta(meta, &m)) return 2;
    NSB_REQUIRE(m.n_pseudo == 16 && m.F == 2 && m.D == 3, "nsb_fused_sdf_bwd: built for 16 x 2 LoTD features in 3-D");
    NSB_REQUIRE(dec->width >= 1 && dec->width <= 64, "nsb_fused_sdf_bwd: decoder width must be <= 64");
    DecoderDevTC d{(const __half *)dec->W1, (const __half *)dec->b1, (const __half *)dec->W2, (const __half *)dec->b2, dec->width,
                   dec->beta};
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t wave = (int64_t)sm_count() * 4;          // TMEM: 4 x 128 columns per SM
    const unsigned grid = (unsigned)(n_tiles < wave ? n_tiles : wave);
    cudaStream_t s = (cudaStream_t)stream;
    const int ml = max_level < 0 ? -1 : max_level;
    if (x == nullptr) k_sdf_bwd_tc<true><<<grid, kTile, 0, s>>>(m, (const __half *)params_half, d, nullptr, rays_o, rays_d, ridx, t, d_sdf, n, ml,
                                                                d_grid, d_W1, d_b1, d_W2, d_b2);
    else k_sdf_bwd_tc<false><<<grid, kTile, 0, s>>>(m, (const __half *)params_half, d, x, nullptr, nullptr, nullptr, nullptr, d_sdf, n, ml,
                                                    d_grid, d_W1, d_b1, d_W2, d_b2);
    return check_launch("nsb_fused_sdf_bwd");
}
