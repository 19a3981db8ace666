// K3 backward: gradient of the spatio-temporal shifted-window attention core.
//
// Autograd twin of winattn_fwd.cu for the op sequence seg18/net/Ours/swin_512.py:119-138 wrapped
// by the gather/scatter of :210-231.  Per (tile, head), with S2 = scale*log2e*QK^T + (bias+mask)*log2e:
//   P  = exp2(S2 - lse2)                    (recomputed, never stored between passes)
//   dP = dO V^T
//   dS = P o (dP - rowsum(P o dP))
//   dV = P^T dO      dQ = scale * dS K      dK = scale * dS^T Q
//   d relative_position_bias_table[idx(i,j), head] += dS[i,j]       (:122-125, summed over windows,
//                                                                     batch and the TxT tiling)
// Inputs  qkv [B,T,H,W,3C] bf16, d_out [B,T,H,W,C] bf16, lse2 from the forward.
// Outputs d_qkv [B,T,H,W,3C] bf16 (scattered to the un-rolled coordinates by TMA),
//         d_table [(2ws-1)^2, nH] fp32 (+=), optional d_qkv_colsum [3C] fp32 (+=, the qkv bias grad).
//
// Streaming: operand chunks [128 rows x 64 ch] flow through a ring of 16 KB slots twice -- once
// for S and dP, once more (from L2) for the three output products.  S, dP, the running sum of dS
// and two 64-column output accumulators live in TMEM.  The bias-table gradient needs the sum of
// dS over every tile this CTA processes; it is accumulated on the tensor core as dS * I (identity
// tile in smem) so no per-element atomics are issued in the main loop.  For that sum to be
// meaningful every tile of a launch uses ONE token order (gm.uniform_quad) and every CTA sees
// ONE head (grid is a multiple of nH).
#include "winattn_common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace stswin {

int fill_geom(WinGeom* gm, int B, int T, int H, int W, int C, int nH, int ws, int shift);
int make_window_tmaps(CUtensorMap* full, CUtensorMap* quad, const void* base, const WinGeom& gm, int channels);

namespace {

constexpr int NR = 6;
constexpr int SLOT_BYTES = 128 * 128;
constexpr int PD_BYTES = 2 * SLOT_BYTES;       // [128 x 128] bf16 as two K-major halves
constexpr int STG_BYTES = 2 * SLOT_BYTES;
constexpr int EYE_BYTES = 64 * 128;            // I64, K-major, 128B swizzle
constexpr int TAB_MAX = 15 * 15;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES =
    1024 + NR * SLOT_BYTES + 2 * PD_BYTES + STG_BYTES + EYE_BYTES + 128 * 4 + 2 * (TAB_MAX + 1) * 4 + 256;
constexpr float kMaskLog2e = -100.0f * 1.4426950408889634f;

template <int L>
__global__ void __launch_bounds__(NUM_THREADS, 1)
winattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv_full, const __grid_constant__ CUtensorMap tm_qkv_quad,
                   const __grid_constant__ CUtensorMap tm_do_full, const __grid_constant__ CUtensorMap tm_do_quad,
                   const __grid_constant__ CUtensorMap tm_dqkv_full, const __grid_constant__ CUtensorMap tm_dqkv_quad,
                   const float* __restrict__ bias_table, const float* __restrict__ lse2, float* __restrict__ d_table,
                   float* __restrict__ d_colsum, const WinGeom gm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ring = smem;
  uint8_t* s_p = s_ring + NR * SLOT_BYTES;
  uint8_t* s_ds = s_p + PD_BYTES;
  uint8_t* s_stg = s_ds + PD_BYTES;
  uint8_t* s_eye = s_stg + STG_BYTES;
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_eye + EYE_BYTES);
  float* s_tab = reinterpret_cast<float*>(s_lut + 128);
  float* s_bacc = s_tab + TAB_MAX + 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_bacc + TAB_MAX + 1) + 7) & ~uintptr_t(7));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NR;
  uint64_t* sdp_full = bars + 2 * NR;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_full + 2;
  uint64_t* pds_free = sdp_full + 3;
  uint64_t* obuf_full = sdp_full + 4;   // [2]
  uint64_t* obuf_free = sdp_full + 6;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = gm.num_tiles * gm.nH;
  const int nc = gm.nc;
  const int head = blockIdx.x % gm.nH;          // constant per CTA: gridDim.x % nH == 0

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv_full);
    tma_prefetch_desc(&tm_qkv_quad);
    tma_prefetch_desc(&tm_do_full);
    tma_prefetch_desc(&tm_do_quad);
    tma_prefetch_desc(&tm_dqkv_full);
    tma_prefetch_desc(&tm_dqkv_quad);
    for (int i = 0; i < NR; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 128);
    mbar_init(pds_full, 128);
    mbar_init(pds_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&obuf_full[i], 1);
      mbar_init(&obuf_free[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // P / dS: entries outside a row's own window stay zero.  Identity tile for the dS running sum.
  for (int i = threadIdx.x; i < 2 * PD_BYTES / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(s_p)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 64 * 8; i += NUM_THREADS) {
    const int n = i >> 3, ch = i & 7;            // row n, 16-byte chunk ch (columns 8*ch .. 8*ch+7)
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((n >> 3) == ch) {
      const uint32_t one = 0x3F80u;              // bf16 1.0
      const int e = n & 7;
      uint32_t w = (e & 1) ? (one << 16) : one;
      if ((e >> 1) == 0) v.x = w; else if ((e >> 1) == 1) v.y = w; else if ((e >> 1) == 2) v.z = w; else v.w = w;
    }
    *reinterpret_cast<uint4*>(s_eye + sw128_offset(n, ch)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_dP = tmem_base + 128;
  const uint32_t tmem_acc = tmem_base + 256;    // running sum of dS over this CTA's tiles
  const uint32_t tmem_out = tmem_base + 384;    // two 64-column output accumulators

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (all lanes issue boxes)
    int slot = 0;
    uint32_t phase = 0;
    auto load = [&](int tile, bool from_do, int ch0) {
      mbar_wait(&empty_bar[slot], phase ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[slot], SLOT_BYTES);
      __syncwarp();
      tile_boxes<true>(gm, tile, ch0, s_ring + slot * SLOT_BYTES, from_do ? &tm_do_full : &tm_qkv_full,
                       from_do ? &tm_do_quad : &tm_qkv_quad, &full_bar[slot], lane);
      if (++slot == NR) { slot = 0; phase ^= 1; }
    };
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / gm.nH;
      const int hq = head * gm.hd;
      for (int c = 0; c < nc; ++c) {
        load(tile, false, 0 * gm.C + hq + c * 64);   // Q_c
        load(tile, false, 1 * gm.C + hq + c * 64);   // K_c
        load(tile, true, hq + c * 64);               // dO_c
        load(tile, false, 2 * gm.C + hq + c * 64);   // V_c
      }
      for (int c = 0; c < nc; ++c) {
        load(tile, true, hq + c * 64);               // dO_c -> dV_c
        load(tile, false, 1 * gm.C + hq + c * 64);   // K_c  -> dQ_c
        load(tile, false, 0 * gm.C + hq + c * 64);   // Q_c  -> dK_c
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 64, 0, 0);   // dS * I64
      constexpr uint32_t idesc_kn = umma_idesc_bf16(128, 64, 0, 1);    // dQ = dS K
      constexpr uint32_t idesc_nn = umma_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q
      int slot = 0;
      uint32_t phase = 0, itp = 0;
      int ob = 0;
      uint32_t ob_phase = 0;
      bool first = true;
      const uint32_t p_addr = smem_u32(s_p), ds_addr = smem_u32(s_ds), eye_addr = smem_u32(s_eye);
      auto next_slot = [&]() { if (++slot == NR) { slot = 0; phase ^= 1; } };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, itp ^= 1) {
        mbar_wait(sdp_free, itp ^ 1);
        tc_fence_after();
        for (int c = 0; c < nc; ++c) {
          for (int pair = 0; pair < 2; ++pair) {     // (Q_c, K_c) -> S ; (dO_c, V_c) -> dP
            const int sa = slot;
            mbar_wait(&full_bar[slot], phase);
            next_slot();
            const int sb = slot;
            mbar_wait(&full_bar[slot], phase);
            next_slot();
            tc_fence_after();
            const uint32_t aa = smem_u32(s_ring + sa * SLOT_BYTES), ba = smem_u32(s_ring + sb * SLOT_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(pair == 0 ? tmem_S : tmem_dP, umma_smem_desc(aa + kk * 32, 16, 1024),
                        umma_smem_desc(ba + kk * 32, 16, 1024), idesc_kk, (c > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&empty_bar[sa]);
            umma_commit(&empty_bar[sb]);
          }
        }
        umma_commit(sdp_full);

        mbar_wait(pds_full, itp);
        tc_fence_after();
        // running sum of dS:  acc[:, 64h + n] += sum_k dS[:, 64h + k] * I64[n, k]
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_acc + h * 64, umma_smem_desc(ds_addr + h * SLOT_BYTES + kk * 32, 16, 1024),
                      umma_smem_desc(eye_addr + kk * 32, 16, 1024), idesc_acc, (!first || kk > 0) ? 1u : 0u);
        first = false;
        for (int c = 0; c < nc; ++c) {
          for (int o = 0; o < 3; ++o) {              // dV_c, dQ_c, dK_c
            mbar_wait(&full_bar[slot], phase);
            mbar_wait(&obuf_free[ob], ob_phase ^ 1);
            tc_fence_after();
            const uint32_t xa = smem_u32(s_ring + slot * SLOT_BYTES);
            const uint32_t dst = tmem_out + ob * 64;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t bdesc = umma_smem_desc(xa + kk * 2048, SLOT_BYTES, 1024);   // [rows x 64ch], MN-major
              if (o == 0)        // dV = P^T dO : A = P viewed MN-major (m = key), k = query
                umma_bf16(dst, umma_smem_desc(p_addr + kk * 2048, SLOT_BYTES, 1024), bdesc, idesc_nn, kk > 0);
              else if (o == 1)   // dQ = dS K   : A = dS K-major, k = key
                umma_bf16(dst, umma_smem_desc(ds_addr + (kk >> 2) * SLOT_BYTES + (kk & 3) * 32, 16, 1024), bdesc,
                          idesc_kn, kk > 0);
              else               // dK = dS^T Q : A = dS viewed MN-major
                umma_bf16(dst, umma_smem_desc(ds_addr + kk * 2048, SLOT_BYTES, 1024), bdesc, idesc_nn, kk > 0);
            }
            umma_commit(&empty_bar[slot]);
            umma_commit(&obuf_full[ob]);
            next_slot();
            if (++ob == 2) { ob = 0; ob_phase ^= 1; }
          }
        }
        umma_commit(pds_free);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax-backward + epilogue
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int sm_tid = threadIdx.x - 64;
    const uint32_t t_lane = uint32_t(wq * 32) << 16;
    const int nbias = (2 * gm.ws - 1) * (2 * gm.ws - 1);
    for (int i = sm_tid; i < nbias; i += 128) {
      s_tab[i] = __ldg(bias_table + i * gm.nH + head) * 1.4426950408889634f;
      s_bacc[i] = 0.f;
    }
    uint32_t itp = 0;
    int ob = 0, stg_sel = 0;
    uint32_t ob_phase = 0;
    constexpr int CH = (L >= 32) ? 32 : 16;
    int key_i = 0, col0 = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, itp ^= 1) {
      const int tile = item / gm.nH;
      const RowGeom rg = row_geom(gm, tile, row);
      s_lut[row] = uint32_t(rg.rr * (2 * gm.ws - 1) + rg.cc) | (uint32_t(rg.id) << 8);
      named_bar_sync(1, 128);
      key_i = (rg.rr + gm.ws - 1) * (2 * gm.ws - 1) + rg.cc + gm.ws - 1;
      col0 = rg.g * L;
      const bool use_mask = rg.wraps;
      const float lse_i = lse2[(size_t)item * 128 + rg.canon];

      mbar_wait(sdp_full, itp);
      mbar_wait(pds_free, itp ^ 1);
      tc_fence_after();
      // pass 1: P = exp2(S2 - lse2) -> smem (bf16), delta = sum_j P dP
      float delta = 0.f;
#pragma unroll
      for (int cb = 0; cb < L / CH; ++cb) {
        uint32_t v[32], w[32];
        tmem_ld_row_chunk<L>(tmem_S, t_lane, col0, cb, wq, lane, v);
        tmem_ld_row_chunk<L>(tmem_dP, t_lane, col0, cb, wq, lane, w);
#pragma unroll
        for (int j8 = 0; j8 < CH / 8; ++j8) {
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float pv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int jj = j8 * 8 + 2 * h + e;
              const uint32_t lj = s_lut[col0 + cb * CH + jj];
              float x = fmaf(__uint_as_float(v[jj]), gm.scale_log2e, s_tab[key_i - int(lj & 0xff)]);
              if (use_mask && (lj >> 8) != uint32_t(rg.id)) x += kMaskLog2e;
              pv[e] = fast_exp2(x - lse_i);
              delta = fmaf(pv[e], __uint_as_float(w[jj]), delta);
            }
            pk[h] = pack_bf16(pv[0], pv[1]);
          }
          const int col = col0 + cb * CH + j8 * 8;
          *reinterpret_cast<uint4*>(s_p + (col >> 6) * SLOT_BYTES + sw128_offset(row, (col & 63) >> 3)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      // pass 2: dS = P o (dP - delta) -> smem (bf16)
#pragma unroll
      for (int cb = 0; cb < L / CH; ++cb) {
        uint32_t w[32];
        tmem_ld_row_chunk<L>(tmem_dP, t_lane, col0, cb, wq, lane, w);
#pragma unroll
        for (int j8 = 0; j8 < CH / 8; ++j8) {
          const int col = col0 + cb * CH + j8 * 8;
          const uint32_t off = (col >> 6) * SLOT_BYTES + sw128_offset(row, (col & 63) >> 3);
          const uint4 pq = *reinterpret_cast<const uint4*>(s_p + off);
          const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
          uint32_t dk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 pf = unpack_bf16(pw[h]);
            // rows of a padding window hold filler data: keep them out of the bias-table sum
            const float d0 = rg.valid ? pf.x * (__uint_as_float(w[j8 * 8 + 2 * h]) - delta) : 0.f;
            const float d1 = rg.valid ? pf.y * (__uint_as_float(w[j8 * 8 + 2 * h + 1]) - delta) : 0.f;
            dk[h] = pack_bf16(d0, d1);
          }
          *reinterpret_cast<uint4*>(s_ds + off) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
        }
      }
      tc_fence_before();
      mbar_arrive(sdp_free);          // S / dP may be overwritten by the next tile
      fence_proxy_async_smem();
      mbar_arrive(pds_full);          // P / dS visible to the tensor core

      // ---- drain dV_c, dQ_c, dK_c : TMEM -> bf16 -> staging -> TMA scatter into d_qkv
      for (int c = 0; c < nc; ++c) {
        for (int o = 0; o < 3; ++o) {
          const int which = (o == 0) ? 2 : (o == 1 ? 0 : 1);
          const float mul = (o == 0) ? 1.0f : gm.scale;
          mbar_wait(&obuf_full[ob], ob_phase);
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tmem_ld32(tmem_out + ob * 64 + t_lane, v0);
          tmem_ld32(tmem_out + ob * 64 + t_lane + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&obuf_free[ob]);
          if (++ob == 2) { ob = 0; ob_phase ^= 1; }
          uint8_t* stg = s_stg + stg_sel * SLOT_BYTES;
          if (sm_tid < 32) tma_wait_group_read<1>();
          named_bar_sync(1, 128);
          const float m2 = rg.valid ? mul : 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16(__uint_as_float(v0[8 * j + 0]) * m2, __uint_as_float(v0[8 * j + 1]) * m2);
            q.y = pack_bf16(__uint_as_float(v0[8 * j + 2]) * m2, __uint_as_float(v0[8 * j + 3]) * m2);
            q.z = pack_bf16(__uint_as_float(v0[8 * j + 4]) * m2, __uint_as_float(v0[8 * j + 5]) * m2);
            q.w = pack_bf16(__uint_as_float(v0[8 * j + 6]) * m2, __uint_as_float(v0[8 * j + 7]) * m2);
            *reinterpret_cast<uint4*>(stg + sw128_offset(row, j)) = q;
            q.x = pack_bf16(__uint_as_float(v1[8 * j + 0]) * m2, __uint_as_float(v1[8 * j + 1]) * m2);
            q.y = pack_bf16(__uint_as_float(v1[8 * j + 2]) * m2, __uint_as_float(v1[8 * j + 3]) * m2);
            q.z = pack_bf16(__uint_as_float(v1[8 * j + 4]) * m2, __uint_as_float(v1[8 * j + 5]) * m2);
            q.w = pack_bf16(__uint_as_float(v1[8 * j + 6]) * m2, __uint_as_float(v1[8 * j + 7]) * m2);
            *reinterpret_cast<uint4*>(stg + sw128_offset(row, 4 + j)) = q;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          const int ch0 = which * gm.C + head * gm.hd + c * 64;
          if (sm_tid < 32) {
            tile_boxes<false>(gm, tile, ch0, stg, &tm_dqkv_full, &tm_dqkv_quad, nullptr, lane);
            tma_commit_group();
          }
          if (d_colsum != nullptr) {
            // thread: column pair (lane), rows 32*wq .. 32*wq+31 of the staged chunk
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const uint32_t wv = *reinterpret_cast<const uint32_t*>(stg + sw128_offset(wq * 32 + r, lane >> 2) + (lane & 3) * 4);
              const float2 f = unpack_bf16(wv);
              s0 += f.x; s1 += f.y;
            }
            atomicAdd(d_colsum + ch0 + 2 * lane, s0);
            atomicAdd(d_colsum + ch0 + 2 * lane + 1, s1);
          }
          stg_sel ^= 1;
        }
      }
    }
    // ---- bias-table gradient: bin the running dS sum by relative position (all MMAs have retired:
    //      the last obuf_full commit covers every earlier tcgen05.mma of the issuing thread)
    named_bar_sync(1, 128);
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < L / CH; ++cb) {
      uint32_t v[32];
      tmem_ld_row_chunk<L>(tmem_acc, t_lane, col0, cb, wq, lane, v);
#pragma unroll
      for (int jj = 0; jj < CH; ++jj) {
        const uint32_t lj = s_lut[col0 + cb * CH + jj];
        atomicAdd(&s_bacc[key_i - int(lj & 0xff)], __uint_as_float(v[jj]));
      }
    }
    named_bar_sync(1, 128);
    for (int i = sm_tid; i < nbias; i += 128) atomicAdd(d_table + i * gm.nH + head, s_bacc[i]);
    if (sm_tid < 32) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <typename K>
int set_smem_bwd(K kern, int bytes) {
  STSWIN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return kOk;
}

}  // namespace

// see include/stswin_b200.h : stswin_winattn_bwd
int winattn_bwd(const void* qkv, const float* bias_table, const float* lse2, const void* d_out, void* d_qkv,
                float* d_table, float* d_qkv_colsum, int B, int T, int H, int W, int C, int nH, int ws, int shift,
                float qk_scale, cudaStream_t stream) {
  STSWIN_CHECK_ARG(qkv && bias_table && lse2 && d_out && d_qkv && d_table, "winattn_bwd: null pointer");
  WinGeom gm;
  int rc = fill_geom(&gm, B, T, H, W, C, nH, ws, shift);
  if (rc != kOk) return rc;
  gm.uniform_quad = 1;
  if (qk_scale > 0.f) { gm.scale = qk_scale; gm.scale_log2e = qk_scale * 1.4426950408889634f; }
  CUtensorMap tq_full, tq_quad, td_full, td_quad, tg_full, tg_quad;
  if ((rc = make_window_tmaps(&tq_full, &tq_quad, qkv, gm, 3 * C)) != kOk) return rc;
  if ((rc = make_window_tmaps(&td_full, &td_quad, d_out, gm, C)) != kOk) return rc;
  if ((rc = make_window_tmaps(&tg_full, &tg_quad, d_qkv, gm, 3 * C)) != kOk) return rc;
  const int items = gm.num_tiles * gm.nH;
  int grid = items < num_sms() ? items : num_sms();
  grid -= grid % nH;                      // one head per CTA (items is a multiple of nH, so grid >= nH)
  if (grid < nH) return set_error(kErrUnsupported, "winattn_bwd: num_heads %d exceeds the SM count", nH);
#define STSWIN_LAUNCH_BWD(LL)                                                                                      \
  case LL: {                                                                                                       \
    if ((rc = set_smem_bwd(winattn_bwd_kernel<LL>, SMEM_BYTES)) != kOk) return rc;                                 \
    winattn_bwd_kernel<LL><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq_full, tq_quad, td_full, td_quad, tg_full, \
                                                                      tg_quad, bias_table, lse2, d_table,         \
                                                                      d_qkv_colsum, gm);                           \
    break;                                                                                                         \
  }
  switch (gm.L) {
    STSWIN_LAUNCH_BWD(16)
    STSWIN_LAUNCH_BWD(32)
    STSWIN_LAUNCH_BWD(64)
    STSWIN_LAUNCH_BWD(128)
    default: return set_error(kErrUnsupported, "winattn_bwd: L=%d", gm.L);
  }
#undef STSWIN_LAUNCH_BWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
