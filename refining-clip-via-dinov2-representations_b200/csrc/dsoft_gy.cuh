// Two-phase backward for sm_100a, second phase: the gradient GEMM dX = G . Y.
//
// The fused backward (dsoft_bwd_kernel) keeps G in shared memory, which limits it to M = 128 x N = 128 S tiles
// next to a 256-column accumulator in TMEM (~40 % tensor activity measured for that MMA shape).  When the
// caller can spare 2 bytes per element of the local [b x cols] blocks (DSOFT_F_GMAT), the backward runs as
//   1. the forward kernels' main loop again (CTA pairs, cta_group::2) with an epilogue that turns each
//      recomputed tile plus the row / column soft-max statistics into the fp16 logit-gradient tile and
//      streams it to HBM (dsoft_fwd_kernel<MODE_CLIP_G / MODE_SOFT_G>; same formulas and rounding as the fused
//      kernel's epilogue), the teacher tile being computed once for the student and the text term;
//   2. this plain tensor-core GEMM dX[b x Dout] = G[b x cols] . Y16[cols x Dout] in the best shape on this
//      part: cta_group::2 MMAs of M = 256 x N = 256 (~1.3 PFLOP/s measured here).
#pragma once

#include "dsoft_kernels.cuh"

namespace dsoft {

// ================================================================================================
// dX = G . Y16      (fp16 x fp16 -> fp32 in TMEM)
// ================================================================================================
//  A = G   blocked [row block][64-column K tile][128 rows][64 cols]: one TMA box (SW128, K-major) = 16 KiB of
//          contiguous memory; addressed as a 2-D tensor of 64 columns x (row blocks * K tiles * 128) rows
//  B = Y16 [cols][Dout], MN-major (features contiguous), box 64 features x 64 rows, SW128
//  AT = true multiplies with G^T instead (dT = G^T . I16 for the CLIP text rows when world == 1, where the text
//  rows' logit-gradient matrix is exactly the transpose of the image rows' one): M runs over G's columns, K over
//  its rows; the A tile of a CTA is two boxes of 64 G-columns x 64 G-rows, MN-major like the B operand.
//  grid (2, n tiles, row pairs * k splits), cluster (2,1,1): CTA `prank` holds 128 of the pair's 256 rows
//  and 128 of the tile's 256 features; the leader issues the MMAs for both.
//  ring: 6 stages x (A 16 KiB | B 2 x 8 KiB); TMEM: 256 fp32 accumulator columns per CTA.
//  warps: 0 = TMA producer, 1 = MMA issuer (leader), 2 = TMEM alloc, 4..11 = drain
constexpr int GY_STAGES = 6;
constexpr int GY_N = 256;
constexpr int GY_SMEM_BYTES = GY_STAGES * SLAB + 1024 + 256;

struct GyParams {
  int b;                // rows of the output (AT: columns of G in scope)
  int dout;             // gradient features
  int ksteps;           // K steps of 64 columns over the whole (padded) column range
  int steps_per_split;  // K steps per split
  int nsplit;
  int ycol0;            // row of Y16 that matches K index 0
  int g_ktiles;         // 64-column tiles per row block of the blocked G (its pitch / 64)
  int tri;              // symmetric G of which only the blocks from each row pair's 256-column diagonal block
                        // onwards exist: K steps below 4 * pair read the transposed blocks instead (AT-style)
  float* acc_part;      // [nsplit][b][dout] fp32
  // ---- symmetric soft matrices across ranks (DSOFT_SYM_W; all zero otherwise)
  int ywrap;            // rows of Y16 wrap around the global batch: row = (ycol0 + 64 k) mod ywrap
  int rb0;              // AT: first G column block (128 columns) of the output; output row 0 = column 128 rb0
  int pair_half, kend_a;  // normal launch: row pairs < pair_half stop at K step kend_a (their G tiles end there)
  int rb_a, klo_b;        // AT: output blocks >= rb_a only sum over the G rows from K step klo_b on
};

template <bool AT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dsoft_gy_kernel(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap gmap64,
                const __grid_constant__ CUtensorMap vmap, const __grid_constant__ GyParams P) {
  // gmap: 128-row boxes (K-major A tiles); gmap64: 64-row boxes of the same matrix (transposed, MN-major A tiles)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GY_STAGES * SLAB);
  uint64_t* ring_full = bars;               // [6] the leader's copy collects both CTAs' bytes
  uint64_t* ring_empty = bars + GY_STAGES;  // [6]
  uint64_t* acc_full = bars + 2 * GY_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int prank = blockIdx.x;  // cluster rank: grid.x == cluster.x == 2
  const bool leader = prank == 0;
  const int nt = blockIdx.y;
  const int pair = blockIdx.z / P.nsplit;
  const int split = blockIdx.z % P.nsplit;
  const int rb = (AT ? P.rb0 : 0) + pair * 2 + prank;  // AT: G column block; else: row block
  int k0 = split * P.steps_per_split;
  int k1 = min(k0 + P.steps_per_split, P.ksteps);
  if (!AT && pair < P.pair_half) k1 = min(k1, P.kend_a);
  if (AT && P.rb0 + pair * 2 >= P.rb_a && P.rb_a > 0) k0 = max(k0, P.klo_b);
  const int f0 = nt * GY_N + prank * (GY_N / 2);  // first feature staged by this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&gmap);
    tma_prefetch_desc(&gmap64);
    tma_prefetch_desc(&vmap);
    for (int i = 0; i < GY_STAGES; ++i) {
      mbar_init(smem_u32(&ring_full[i]), 1);
      mbar_init(smem_u32(&ring_empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<2>(smem_u32(tmem_holder), GY_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int k = k0; k < k1; ++k) {
      mbar_wait(smem_u32(&ring_empty[stage]), phase ^ 1);
      if (elect_one()) {
        const uint32_t full = smem_u32(&ring_full[stage]);
        const uint32_t full_leader = mapa_shared(full, 0);
        const uint32_t dst = smem_u32(smem + stage * SLAB);
        if (leader) mbar_arrive_expect_tx(full, 2 * SLAB);
        if (AT || (P.tri && k < 4 * pair)) {
          // K step k = G rows [64k, 64k + 64) = half `k & 1` of row block `k >> 1`; M = G columns 128 rb ..
          const int grow = ((k >> 1) * P.g_ktiles + 2 * rb) * BM + (k & 1) * 64;
          tma_load_2d_2sm(dst, &gmap64, full_leader, 0, grow);
          tma_load_2d_2sm(dst + TILE_BYTES / 2, &gmap64, full_leader, 0, grow + BM);
        } else {
          tma_load_2d_2sm(dst, &gmap, full_leader, 0, (rb * P.g_ktiles + k) * BM);
        }
        const int yrow = P.ywrap ? (P.ycol0 + k * BK) % P.ywrap : P.ycol0 + k * BK;
        tma_load_2d_2sm(dst + TILE_BYTES, &vmap, full_leader, f0, yrow);
        tma_load_2d_2sm(dst + TILE_BYTES + TILE_BYTES / 2, &vmap, full_leader, f0 + BK, yrow);
      }
      __syncwarp();
      if (++stage == GY_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc_n = make_idesc_bf16(2 * BM, GY_N, 0, 1, 1);  // fp16, A K-major, B MN-major
      const uint32_t idesc_t = make_idesc_bf16(2 * BM, GY_N, 1, 1, 1);  // fp16, A MN-major (transposed blocks)
      for (int k = k0; k < k1; ++k) {
        mbar_wait(smem_u32(&ring_full[stage]), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * SLAB);
          const uint32_t v_addr = a_addr + TILE_BYTES;
          const bool tr = AT || (P.tri && k < 4 * pair);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = tr ? make_smem_desc(a_addr + kk * 2048, TILE_BYTES / 2, 1024)
                                   : make_smem_desc(a_addr + kk * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(v_addr + kk * 2048, TILE_BYTES / 2, 1024);
            umma_cg<2>(tmem_base, ad, bd, tr ? idesc_t : idesc_n, (k == k0 && kk == 0) ? 0u : 1u);
          }
          umma_commit_cg<2>(smem_u32(&ring_empty[stage]));
          if (k == k1 - 1) umma_commit_cg<2>(smem_u32(acc_full));
        }
        __syncwarp();
        if (++stage == GY_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int li = (pair * 2 + prank) * BM + q * 32 + lane;  // output row (AT: relative to column 128 rb0)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float v[32];
    const bool any = k1 > k0;
    if (any) {
      mbar_wait(smem_u32(acc_full), 0);
      tc_fence_after();
    }
    const int fbase = nt * GY_N;
    float* dst = P.acc_part + (static_cast<size_t>(split) * P.b + li) * P.dout + fbase;
    const int nvalid = min(GY_N, P.dout - fbase);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int cf = half * 128 + c * 32;
      if (cf >= nvalid) break;  // warp-uniform
      if (any) {
        tmem_ld32(lane_addr + cf, v);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
      if (li < P.b) {
        if (cf + 32 <= nvalid) {
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4)
            *reinterpret_cast<float4*>(dst + cf + 4 * e4) =
                make_float4(v[4 * e4], v[4 * e4 + 1], v[4 * e4 + 2], v[4 * e4 + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (cf + e < nvalid) dst[cf + e] = v[e];
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg<2>(tmem_base, GY_N);
}

}  // namespace dsoft
