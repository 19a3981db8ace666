// K3 forward: spatio-temporal shifted-window attention core.
//
// Replaces, in one kernel (reference: seg18/net/Ours/swin_512.py)
//   torch.roll(-s) + window_partition + view/permute           :210-218   -> TMA box coordinates
//   q*scale, q @ k^T                                            :119-120   -> tcgen05 (S in TMEM)
//   relative_position_bias gather + repeat(1,T,T), + mask       :122-131   -> registers
//   softmax                                                     :132-134   -> registers (exp2)
//   attn @ v, transpose/reshape                                 :138       -> tcgen05 (O in TMEM)
//   window_reverse + torch.roll(+s)                             :224-231   -> TMA store coordinates
//
// Input  qkv [B, T, H, W, 3C] bf16 (un-rolled token order; channel = which*C + head*hd + d, :116)
// Output out [B, T, H, W, C]  bf16 (same token order), lse2 [num_tiles, nH, 128] fp32
//        (base-2 log-sum-exp of every tile row, consumed by the backward kernel).
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax + epilogue
// (thread <-> tile row <-> TMEM lane).  Operand chunks stream through a ring of 16 KB slots.
#include <cstdlib>
#include <type_traits>

#include "winattn_common.cuh"
#include "host_util.h"

namespace stswin {

namespace {

constexpr int NSLOT = 9;
constexpr int SLOT_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int P_BYTES = 2 * SLOT_BYTES;        // P [128 x 128] bf16 as two K-major halves
constexpr int STG_BYTES = 2 * SLOT_BYTES;      // two output staging chunks
constexpr int TAB_MAX = 15 * 15;               // (2*ws-1)^2 for ws <= 8
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = 1024 + NSLOT * SLOT_BYTES + P_BYTES + STG_BYTES + 128 * 4 + TAB_MAX * 4 + 256;
constexpr float kMaskLog2e = -100.0f * 1.4426950408889634f;
STSWIN_TRACE_DECL(g_trace_fwd)

// L   : window columns a thread walks (16/32/64/128).  GEN (only with L = 128): the tile holds G windows of
//       gm.L tokens with gm.L not a power of two (e.g. 7x7x2 = 98); every thread then walks the whole
//       128-column row and keeps only the columns tagged with its own window.
//
// WS > 0 selects the fast softmax for the shipped geometries (L = T*WS*WS, whole launch in one token
// order, or one order per window: see ORDER below).  The
// column -> (relative-position key, quadrant) map is then a compile-time function of the column,
// so the bias is one LDS at an immediate offset and the shift mask is one additive constant per
// quadrant.  WS == 0 is the generic path (look-up table per column: dense mask, L = 16, GEN).
template <int L, int WS, bool QUAD>
__host__ __device__ constexpr int col_key(int j) {
  constexpr int W1 = WS > 0 ? WS : 1, N = W1 * W1, HW = W1 > 1 ? W1 / 2 : 1, QL = L / 4;
  if (!QUAD) return ((j % N) / W1) * (2 * W1 - 1) + (j % N) % W1;
  const int q = j / QL, p = (j % QL) % (HW * HW);
  return ((q >> 1) * HW + p / HW) * (2 * W1 - 1) + (q & 1) * HW + p % HW;
}

template <int L, int WS, int ORDER, bool GEN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
winattn_fwd_kernel(const __grid_constant__ WinMaps tm_qkv, const __grid_constant__ WinMaps tm_out,
                   const float* __restrict__ bias_table, float* __restrict__ lse2, const WinGeom gm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_ring = smem;
  uint8_t* s_p = s_ring + NSLOT * SLOT_BYTES;
  uint8_t* s_stg = s_p + P_BYTES;
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_stg + STG_BYTES);     // [128] key | id<<8
  float* s_tab = reinterpret_cast<float*>(s_lut + 128);                 // [TAB_MAX] bias * log2e for this head
  static_assert(((128 + TAB_MAX + 1) * 4) % 8 == 0, "mbarriers need 8-byte alignment");
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab + TAB_MAX + 1);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NSLOT;
  uint64_t* s_full = bars + 2 * NSLOT;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = s_full + 3;
  uint64_t* o_free = s_full + 4;
  uint64_t* pv_done = s_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = gm.num_tiles * gm.ngrp;     // work item = (tile, head group)
  const int nc = gm.nc;
  const int SH = gm.SH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv.full);
    tma_prefetch_desc(&tm_out.full);
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // P buffer: entries outside a row's own window stay zero for the whole kernel.  Ring: the padding
  // rows of a tile (general mode) are never written by TMA and must read as zero.
  for (int i = threadIdx.x; i < (NSLOT * SLOT_BYTES + P_BYTES) / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // 128 columns
  const uint32_t tmem_O = tmem_base + 128;    // nc*64 columns (<= 256), or SH*64 when two heads share a chunk

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int slot = 0;
    uint32_t phase = 0;
    int itl = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++itl) {
      const int tile = item / gm.ngrp, hg = item - tile * gm.ngrp;
      for (int step = 0; step < 3 * nc; ++step) {
        // order: Q0 K0 Q1 K1 ... then V0 V1 ...
        int which, c;
        if (step < 2 * nc) { which = step & 1; c = step >> 1; }
        else               { which = 2; c = step - 2 * nc; }
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (lane == 0 && step == 0) WTRACE(g_trace_fwd, itl, 0);
        if (lane == 0 && step == 3 * nc - 1) WTRACE(g_trace_fwd, itl, 1);
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[slot], chunk_tx_bytes(gm));
        __syncwarp();
        tile_boxes<true>(gm, tile, which * gm.C + hg * gm.gch + c * 64, s_ring + slot * SLOT_BYTES, &tm_qkv,
                         &full_bar[slot], lane);
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      int slot = 0;
      uint32_t phase = 0;
      uint32_t it_phase = 0, sub_phase = 0;      // per item / per (item, sub-head)
      const uint32_t p_addr = smem_u32(s_p);
      auto take_slot = [&]() {                   // wait for the next ring slot in program order
        const int sl = slot;
        mbar_wait(&full_bar[slot], phase);
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        return sl;
      };
      int itl = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, it_phase ^= 1, ++itl) {
        int sq[4], sk[4], sv[4];                 // ring slots of this item's chunks (nc <= 4)
        for (int sub = 0; sub < SH; ++sub, sub_phase ^= 1) {
          // S = Q K^T (for SH == 2: over the 32-channel K sub-range of sub-head `sub`)
          mbar_wait(s_free, sub_phase ^ 1);
          tc_fence_after();
          for (int c = 0; c < nc; ++c) {
            if (sub == 0) { sq[c] = take_slot(); sk[c] = take_slot(); }
            if (c == nc - 1) WTRACE(g_trace_fwd, itl, 13);
            tc_fence_after();
            const uint32_t qa = smem_u32(s_ring + sq[c] * SLOT_BYTES), ka = smem_u32(s_ring + sk[c] * SLOT_BYTES);
            const int k0 = (SH == 1) ? 0 : sub * 2, k1 = (SH == 1) ? 4 : sub * 2 + 2;
            for (int kk = k0; kk < k1; ++kk)
              umma_bf16(tmem_S, umma_smem_desc(qa + kk * 32, 16, 1024), umma_smem_desc(ka + kk * 32, 16, 1024), idesc_s,
                        (c > 0 || kk > k0) ? 1u : 0u);
            if (sub == SH - 1) {                 // last reader of these chunks
              umma_commit(&empty_bar[sq[c]]);
              umma_commit(&empty_bar[sk[c]]);
            }
          }
          umma_commit(s_full);
          WTRACE(g_trace_fwd, itl, 2);
          // O = P V  (SH == 2: the whole 64-column chunk; the epilogue keeps the sub-head's own 32 columns)
          mbar_wait(p_full, sub_phase);
          if (sub == 0) mbar_wait(o_free, it_phase ^ 1);
          tc_fence_after();
          WTRACE(g_trace_fwd, itl, 3);
          for (int c = 0; c < nc; ++c) {
            if (sub == 0) sv[c] = take_slot();
            if (c == nc - 1) WTRACE(g_trace_fwd, itl, 14);
            tc_fence_after();
            const uint32_t va = smem_u32(s_ring + sv[c] * SLOT_BYTES);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t adesc = umma_smem_desc(p_addr + (kk >> 2) * SLOT_BYTES + (kk & 3) * 32, 16, 1024);
              const uint64_t bdesc = umma_smem_desc(va + kk * 2048, SLOT_BYTES, 1024);
              umma_bf16(tmem_O + (SH == 1 ? c : sub) * 64, adesc, bdesc, idesc_o, kk > 0 ? 1u : 0u);
            }
            if (sub == SH - 1) umma_commit(&empty_bar[sv[c]]);
          }
          umma_commit(pv_done);                  // P may be overwritten
          WTRACE(g_trace_fwd, itl, 4);
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue
    const int wq = warp & 3;
    const int row = wq * 32 + lane;           // tile row == TMEM lane
    const int sm_tid = threadIdx.x - 64;      // 0..127
    const uint32_t t_lane = uint32_t(wq * 32) << 16;
    const int nbias = (2 * gm.ws - 1) * (2 * gm.ws - 1);
    uint32_t it_phase = 0, sub_phase = 0;
    int stg_sel = 0;
    int itl = 0;
    const bool tr = (threadIdx.x == 64);
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, it_phase ^= 1, ++itl) {
      const int tile = item / gm.ngrp, hg = item - tile * gm.ngrp;
      if (tr) WTRACE(g_trace_fwd, itl, 5);
      const RowGeom rg = row_geom(gm, tile, row);
      float inv_sub[2] = {1.f, 1.f};
     for (int sub = 0; sub < SH; ++sub, sub_phase ^= 1) {
      const int head = hg * SH + sub;
      named_bar_sync(1, 128);                  // everybody is done with the previous LUT / bias table
      // key | region id | spatial position | window tag (g + 1, 0 for a padding row)
      const uint32_t my_tag = rg.inrange ? uint32_t(rg.g + 1) : 0u;
      s_lut[row] = uint32_t(rg.rr * (2 * gm.ws - 1) + rg.cc) | (uint32_t(rg.id) << 8) |
                   (uint32_t(rg.rr * gm.ws + rg.cc) << 16) | (my_tag << 24);
      for (int i = sm_tid; i < nbias; i += 128) s_tab[i] = __ldg(bias_table + i * gm.nH + head) * 1.4426950408889634f;
      named_bar_sync(1, 128);
      const int key_i = (rg.rr + gm.ws - 1) * (2 * gm.ws - 1) + rg.cc + gm.ws - 1;
      const int col0 = GEN ? 0 : rg.g * L;
      const bool use_mask = rg.wraps;
      // dense mask row of this query token (stand-alone WindowAttention with an explicit mask tensor)
      const float* mask_row = gm.mask ? gm.mask + ((size_t)(rg.gw % gm.mask_nw) * gm.N + (rg.rr * gm.ws + rg.cc)) * gm.N : nullptr;

      if (tr) WTRACE(g_trace_fwd, itl, 6);
      mbar_wait(s_full, sub_phase);
      tc_fence_after();
      if (tr) WTRACE(g_trace_fwd, itl, 7);
      float s[L];
      float mx = -INFINITY;
      float sum = 0.f;
      constexpr int CH = (L >= 32) ? 32 : 16;
      if constexpr (WS > 0) {
        // ---- fast path: compile-time column map, instantiated per token order of the row's window
        auto softmax_fast = [&](auto quad_tag) {
          constexpr bool QUAD = decltype(quad_tag)::value;
          constexpr int QL = L / 4, NQ = QUAD ? 4 : 1;
          const float* tp = s_tab + key_i;
          float mq[NQ];
#pragma unroll
          for (int q = 0; q < NQ; ++q) mq[q] = -INFINITY;
#pragma unroll
          for (int cb = 0; cb < L / CH; ++cb) {
            uint32_t v[32];
            tmem_ld_row_chunk<L>(tmem_S, t_lane, col0, cb, wq, lane, v);
#pragma unroll
            for (int jj = 0; jj < CH; ++jj) {
              const int j = cb * CH + jj;
              const float x = fmaf(__uint_as_float(v[jj]), gm.scale_log2e, tp[-col_key<L, WS, QUAD>(j)]);
              s[j] = x;
              mq[QUAD ? j / QL : 0] = fmaxf(mq[QUAD ? j / QL : 0], x);
            }
          }
          // S is in registers: the next item's QK^T may overwrite TMEM now
          tc_fence_before();
          mbar_arrive(s_free);
          if (tr) WTRACE(g_trace_fwd, itl, 8);
          // shift mask (swin_512.py:171-192): a window of the last window row / column holds two bands
          // per wrapping axis; tokens of different bands get -100.  In quadrant order the band pair
          // of a token IS its quadrant, so the mask is one additive constant per quadrant of columns.
          float nq[NQ];
          if constexpr (QUAD) {
            const int q_i = (rg.rr >= WS / 2 ? 2 : 0) | (rg.cc >= WS / 2 ? 1 : 0);
            const int wm = (rg.id >= 3 ? 2 : 0) | (rg.id % 3 != 0 ? 1 : 0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              nq[q] = ((q ^ q_i) & wm) ? kMaskLog2e : 0.f;
              mx = fmaxf(mx, mq[q] + nq[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) nq[q] -= mx;
          } else {
            mx = mq[0];
            nq[0] = -mx;
          }
          mbar_wait(pv_done, sub_phase ^ 1);        // the previous P V product has finished reading P
          if (tr) WTRACE(g_trace_fwd, itl, 9);
#pragma unroll
          for (int j8 = 0; j8 < L / 8; ++j8) {
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int j = j8 * 8 + 2 * h;
              const float p0 = fast_exp2(s[j] + nq[QUAD ? j / QL : 0]);
              const float p1 = fast_exp2(s[j + 1] + nq[QUAD ? (j + 1) / QL : 0]);
              sum += p0 + p1;
              w[h] = pack_bf16(p0, p1);
            }
            const int col = col0 + j8 * 8;
            *reinterpret_cast<uint4*>(s_p + (col >> 6) * SLOT_BYTES + sw128_offset(row, (col & 63) >> 3)) =
                make_uint4(w[0], w[1], w[2], w[3]);
          }
        };
        // ORDER 0: unshifted block, row-major.  1: every window in quadrant order.  2: only the windows
        // that wrap (a warp never straddles two windows for L >= 32, so the branch is warp-uniform).
        if (ORDER == 0 || (ORDER == 2 && !rg.wraps)) softmax_fast(std::false_type{});
        else                                         softmax_fast(std::true_type{});
      } else {
#pragma unroll
      for (int cb = 0; cb < L / CH; ++cb) {
        uint32_t v[32];
        tmem_ld_row_chunk<L>(tmem_S, t_lane, col0, cb, wq, lane, v);
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
          const uint32_t lj = s_lut[col0 + cb * CH + jj];
          float x = fmaf(__uint_as_float(v[jj]), gm.scale_log2e, s_tab[key_i - int(lj & 0xff)]);
          if (use_mask && ((lj >> 8) & 0xffu) != uint32_t(rg.id)) x += kMaskLog2e;
          if (mask_row != nullptr) x = fmaf(__ldg(mask_row + ((lj >> 16) & 0xffu)), 1.4426950408889634f, x);
          if (GEN && (lj >> 24) != my_tag) x = -1.0e30f;        // another window's column, or padding
          s[cb * CH + jj] = x;
          mx = fmaxf(mx, x);
        }
      }
      // S is in registers: the next item's QK^T may overwrite TMEM now
      tc_fence_before();
      mbar_arrive(s_free);

      mbar_wait(pv_done, sub_phase ^ 1);        // the previous P V product has finished reading P
#pragma unroll
      for (int j8 = 0; j8 < L / 8; ++j8) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float p0 = fast_exp2(s[j8 * 8 + 2 * h] - mx);
          float p1 = fast_exp2(s[j8 * 8 + 2 * h + 1] - mx);
          if (GEN && !rg.inrange) { p0 = 0.f; p1 = 0.f; }          // padding row: keep P finite and empty
          sum += p0 + p1;
          w[h] = pack_bf16(p0, p1);
        }
        const int col = col0 + j8 * 8;
        *reinterpret_cast<uint4*>(s_p + (col >> 6) * SLOT_BYTES + sw128_offset(row, (col & 63) >> 3)) =
            make_uint4(w[0], w[1], w[2], w[3]);
      }
      }
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      if (tr) WTRACE(g_trace_fwd, itl, 10);
      inv_sub[sub] = (GEN && !rg.inrange) ? 0.f : 1.0f / sum;
      if (!GEN || rg.inrange) lse2[((size_t)tile * gm.nH + head) * 128 + rg.canon] = mx + log2f(sum);
     }

      // ---- epilogue: O * (1/sum) -> bf16 -> staging -> TMA store at the un-rolled coordinates
      mbar_wait(o_full, it_phase);
      tc_fence_after();
      if (tr) WTRACE(g_trace_fwd, itl, 11);
      for (int c = 0; c < nc; ++c) {
        uint32_t v0[32], v1[32];
        // SH == 1: columns [c*64, +64) of this head.  SH == 2: columns [0,32) of sub-head 0's product
        // and columns [32,64) of sub-head 1's product (each product spans the whole 64-channel chunk).
        const float inv = inv_sub[0], inv1 = inv_sub[SH - 1];
        tmem_ld32(tmem_O + t_lane + (SH == 1 ? c * 64 : 0), v0);
        tmem_ld32(tmem_O + t_lane + (SH == 1 ? c * 64 + 32 : 64 + 32), v1);
        tmem_ld_wait();
        if (c == nc - 1) {
          tc_fence_before();
          mbar_arrive(o_free);
        }
        uint8_t* stg = s_stg + stg_sel * SLOT_BYTES;
        if (sm_tid < 32) tma_wait_group_read<1>();     // the stores that last used this buffer have drained
        named_bar_sync(1, 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16(__uint_as_float(v0[8 * j + 0]) * inv, __uint_as_float(v0[8 * j + 1]) * inv);
          q.y = pack_bf16(__uint_as_float(v0[8 * j + 2]) * inv, __uint_as_float(v0[8 * j + 3]) * inv);
          q.z = pack_bf16(__uint_as_float(v0[8 * j + 4]) * inv, __uint_as_float(v0[8 * j + 5]) * inv);
          q.w = pack_bf16(__uint_as_float(v0[8 * j + 6]) * inv, __uint_as_float(v0[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(stg + sw128_offset(row, j)) = q;
          q.x = pack_bf16(__uint_as_float(v1[8 * j + 0]) * inv1, __uint_as_float(v1[8 * j + 1]) * inv1);
          q.y = pack_bf16(__uint_as_float(v1[8 * j + 2]) * inv1, __uint_as_float(v1[8 * j + 3]) * inv1);
          q.z = pack_bf16(__uint_as_float(v1[8 * j + 4]) * inv1, __uint_as_float(v1[8 * j + 5]) * inv1);
          q.w = pack_bf16(__uint_as_float(v1[8 * j + 6]) * inv1, __uint_as_float(v1[8 * j + 7]) * inv1);
          *reinterpret_cast<uint4*>(stg + sw128_offset(row, 4 + j)) = q;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (sm_tid < 32) {                             // warp 2 issues the scatter, one box per lane
          tile_boxes<false>(gm, tile, hg * gm.gch + c * 64, stg, &tm_out, nullptr, lane);
          tma_commit_group();
        }
        if (tr && c == 0) WTRACE(g_trace_fwd, itl, 15);
        stg_sel ^= 1;
      }
      if (tr) WTRACE(g_trace_fwd, itl, 12);
    }
    if (sm_tid < 32) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int fill_geom(WinGeom* gm, int B, int T, int H, int W, int C, int nH, int ws, int shift) {
  STSWIN_CHECK_ARG(B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && nH > 0 && ws > 0, "winattn: non-positive dimension");
  STSWIN_CHECK_ARG(H % ws == 0 && W % ws == 0, "winattn: H=%d, W=%d must be multiples of the window size %d", H, W, ws);
  STSWIN_CHECK_ARG(C % nH == 0, "winattn: C=%d not divisible by num_heads=%d", C, nH);
  gm->B = B; gm->T = T; gm->H = H; gm->W = W; gm->C = C; gm->nH = nH; gm->ws = ws; gm->shift = shift;
  gm->hd = C / nH;
  gm->N = ws * ws;
  gm->L = T * ws * ws;
  if (!((gm->hd % 64 == 0 && gm->hd <= 256) || (gm->hd == 32 && nH % 2 == 0)))
    return set_error(kErrUnsupported, "winattn: head_dim %d unsupported (need 32 or a multiple of 64, <= 256)", gm->hd);
  if (gm->L > 128 || gm->L < 1)
    return set_error(kErrUnsupported, "winattn: T*ws*ws = %d tokens per window unsupported (need <= 128)", gm->L);
  if (shift < 0 || shift >= ws)
    return set_error(kErrInvalidArg, "winattn: shift %d must be in [0, window_size = %d)", shift, ws);
  if (ws > 8) return set_error(kErrUnsupported, "winattn: window size %d > 8 unsupported", ws);
  gm->general = !(gm->L == 16 || gm->L == 32 || gm->L == 64 || gm->L == 128);
  gm->G = 128 / gm->L;
  gm->nWh = H / ws; gm->nWw = W / ws; gm->nW = gm->nWh * gm->nWw;
  gm->total_windows = B * gm->nW;
  gm->num_tiles = (gm->total_windows + gm->G - 1) / gm->G;
  gm->SH = gm->hd >= 64 ? 1 : 64 / gm->hd;
  gm->ngrp = nH / gm->SH;
  gm->gch = gm->SH * gm->hd;
  gm->nc = gm->gch / 64;
  gm->scale_log2e = 1.4426950408889634f / sqrtf((float)gm->hd);
  gm->scale = 1.0f / sqrtf((float)gm->hd);
  gm->ra = ws - shift; gm->rb = shift;
  {
    int off = 0;
    for (int k = 0; k < 4; ++k) {
      gm->roff[k] = off;
      off += ((k >> 1) ? gm->rb : gm->ra) * ((k & 1) ? gm->rb : gm->ra) * T;
    }
    gm->roff[4] = off;       // == L
  }
  gm->uniform_quad = 0;
  gm->mask = nullptr; gm->mask_nw = 0;
  return kOk;
}

// tensor maps over a [B*T, H, W, channels] bf16 tensor: whole-window box and the four rectangle boxes
int make_window_tmaps(WinMaps* maps, const void* base, const WinGeom& gm, int channels) {
  uint64_t dims[4] = {(uint64_t)channels, (uint64_t)gm.W, (uint64_t)gm.H, (uint64_t)gm.B * gm.T};
  uint64_t str[3] = {(uint64_t)channels * 2, (uint64_t)gm.W * channels * 2, (uint64_t)gm.H * gm.W * channels * 2};
  uint32_t box_full[4] = {64, (uint32_t)gm.ws, (uint32_t)gm.ws, (uint32_t)gm.T};
  int rc = make_tmap(&maps->full, TmapDtype::BF16, 4, base, dims, str, box_full, true);
  if (rc != kOk) return rc;
  for (int k = 0; k < 4; ++k) {
    maps->rect[k] = maps->full;
    if (gm.shift > 0) {
      uint32_t box[4] = {64, (uint32_t)((k & 1) ? gm.rb : gm.ra), (uint32_t)((k >> 1) ? gm.rb : gm.ra), (uint32_t)gm.T};
      if ((rc = make_tmap(&maps->rect[k], TmapDtype::BF16, 4, base, dims, str, box, true)) != kOk) return rc;
    }
  }
  return kOk;
}

template <typename K>
static int set_smem(K kern, int bytes) {
  STSWIN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return kOk;
}

long winattn_lse_elems(int B, int T, int H, int W, int C, int nH, int ws) {
  WinGeom gm;
  if (fill_geom(&gm, B, T, H, W, C, nH, ws, 0) != kOk) return -1;
  return (long)gm.num_tiles * gm.nH * 128;
}

#ifdef STSWIN_TRACE
extern "C" int stswin_debug_trace_fwd(long long* buf) {
  return cudaMemcpyToSymbol(g_trace_fwd, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#endif

// see include/stswin_b200.h : stswin_winattn_fwd
int winattn_fwd(const void* qkv, const float* bias_table, void* out, float* lse2, int B, int T, int H, int W, int C,
                int nH, int ws, int shift, float qk_scale, const float* mask, int mask_windows, cudaStream_t stream) {
  STSWIN_CHECK_ARG(qkv && bias_table && out && lse2, "winattn_fwd: null pointer");
  WinGeom gm;
  int rc = fill_geom(&gm, B, T, H, W, C, nH, ws, shift);
  if (rc != kOk) return rc;
  if (qk_scale > 0.f) { gm.scale = qk_scale; gm.scale_log2e = qk_scale * 1.4426950408889634f; }
  STSWIN_CHECK_ARG(mask == nullptr || mask_windows > 0, "winattn: mask given with mask_windows <= 0");
  gm.mask = mask; gm.mask_nw = mask_windows;
  gm.uniform_quad = (gm.L == 128 && !gm.general) ? 1 : gm.uniform_quad;
  if (const char* e = getenv("STSWIN_FWD_UNIFORM")) gm.uniform_quad = atoi(e);   // TEMP experiment
  WinMaps tq, to;
  if ((rc = make_window_tmaps(&tq, qkv, gm, 3 * C)) != kOk) return rc;
  if ((rc = make_window_tmaps(&to, out, gm, C)) != kOk) return rc;
  const int items = gm.num_tiles * gm.ngrp;
  const int grid = items < num_sms() ? items : num_sms();
#define STSWIN_LAUNCH_FWD(LL, WW, QQ, GG)                                                                      \
  {                                                                                                            \
    if ((rc = set_smem(winattn_fwd_kernel<LL, WW, QQ, GG>, SMEM_BYTES)) != kOk) return rc;                     \
    winattn_fwd_kernel<LL, WW, QQ, GG><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, to, bias_table, lse2, gm); \
  }
  // fast softmax: the shipped geometries (ws 8 or 4, 1 or 2 frames per window, shift 0 or ws/2, no dense mask)
  const bool fast = !gm.general && mask == nullptr && (shift == 0 || 2 * shift == ws) &&
                    ((ws == 8 && (gm.L == 128 || gm.L == 64)) || (ws == 4 && gm.L == 32));
  // token order of a shifted block: small windows keep one box per interior window (16 tiny quadrant
  // boxes per chunk cost more than they save); 128-token windows use one order for the whole launch
  const int order = shift == 0 ? 0 : (gm.uniform_quad ? 1 : 2);
  if (gm.general) STSWIN_LAUNCH_FWD(128, 0, 0, true)
  else if (fast && gm.L == 128 && order == 0) STSWIN_LAUNCH_FWD(128, 8, 0, false)
  else if (fast && gm.L == 128 && order == 1) STSWIN_LAUNCH_FWD(128, 8, 1, false)
  else if (fast && gm.L == 128) STSWIN_LAUNCH_FWD(128, 8, 2, false)
  else if (fast && gm.L == 64 && order == 0) STSWIN_LAUNCH_FWD(64, 8, 0, false)
  else if (fast && gm.L == 64) STSWIN_LAUNCH_FWD(64, 8, 2, false)
  else if (fast && gm.L == 32 && order == 0) STSWIN_LAUNCH_FWD(32, 4, 0, false)
  else if (fast && gm.L == 32) STSWIN_LAUNCH_FWD(32, 4, 2, false)
  else if (gm.L == 16) STSWIN_LAUNCH_FWD(16, 0, 0, false)
  else if (gm.L == 32) STSWIN_LAUNCH_FWD(32, 0, 0, false)
  else if (gm.L == 64) STSWIN_LAUNCH_FWD(64, 0, 0, false)
  else STSWIN_LAUNCH_FWD(128, 0, 0, false)
#undef STSWIN_LAUNCH_FWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
