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
// Streaming: operand chunks [128 rows x 64 ch] flow through 16 KB ring slots twice -- once from
// HBM for S and dP (ring H, producer warp 0), once more from L2 for the three output products
// (ring L, producer warp 6).  Two independent rings let the HBM loads of tile i+1 run while tile i
// is still in its softmax-backward / second pass (a single ring serialised them: 46-60% of HBM
// peak, profiles/r1_attn_bwd_ncu_summary.txt).  S, dP and two 64-column output accumulators live in
// TMEM; outputs go TMEM -> registers -> global directly (each thread owns one token row and writes
// its 128-byte line), so no staging buffer competes with the rings for shared memory.
// The register-level work (softmax backward, output conversion, bias-gradient sums) is done by TWO
// groups of four warps, i.e. two warps per SM sub-partition: with a single warp per sub-partition
// every dependent instruction exposed its full latency and a tile took ~28k cycles.  For L >= 64
// the groups take alternate 32-column chunks of S / dP (row sums are exchanged through shared
// memory); the groups always alternate the 64-column output chunks (group g drains TMEM buffer g).  The bias-table gradient needs the sum of
// dS over every tile this CTA processes: each thread (= query row) keeps N = ws*ws fp32 running
// sums, one per key position (the TxT tiling folds onto the same entry), and bins them by relative
// position once at the end of the kernel -- no per-element atomics in the main loop, and no bf16
// rounding inside a sum that cancels heavily.  For that sum to be meaningful every tile of a
// launch uses ONE token order (gm.uniform_quad) and every CTA sees ONE head (grid % nH == 0).
// (A first version accumulated bf16 dS on the tensor core through an identity tile; its rounding
// noise reached 6% of the gradient for a 6-window batch.)
#include "winattn_common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace stswin {

int fill_geom(WinGeom* gm, int B, int T, int H, int W, int C, int nH, int ws, int shift);
int make_window_tmaps(WinMaps* maps, const void* base, const WinGeom& gm, int channels);

namespace {

constexpr int NRH = 6;                         // ring H: first-pass chunks, streamed from HBM
constexpr int NRL = 3;                         // ring L: second-pass chunks, re-read from L2
constexpr int SLOT_BYTES = 128 * 128;
constexpr int PD_BYTES = 2 * SLOT_BYTES;       // [128 x 128] bf16 as two K-major halves
constexpr int TAB_MAX = 15 * 15;
constexpr int NUM_THREADS = 384;               // warp 0 producer H, 1 MMA, 2 producer L, 3 idle, 4-7 / 8-11 compute groups A / B
constexpr int SMEM_BYTES =
    1024 + (NRH + NRL) * SLOT_BYTES + 2 * PD_BYTES + 128 * 4 + 3 * (TAB_MAX + 1) * 4 + 2 * 128 * 4 + 256;
constexpr float kMaskLog2e = -100.0f * 1.4426950408889634f;

// GEN (only with L = 128): windows of gm.L tokens with gm.L not a power of two, or rectangles of unequal
// size (shift != ws/2): every thread walks the whole 128-column row and keeps the columns tagged with its
// own window; the bias-table gradient then goes through shared-memory atomics instead of the per-column
// register sums (whose column -> position fold relies on the regular power-of-two orders).
template <int L, int SHT, bool GEN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
winattn_bwd_kernel(const __grid_constant__ WinMaps tm_qkv, const __grid_constant__ WinMaps tm_do,
                   __nv_bfloat16* __restrict__ d_qkv, const float* __restrict__ bias_table, const float* __restrict__ lse2, float* __restrict__ d_table,
                   float* __restrict__ d_colsum, const WinGeom gm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_ringh = smem;
  uint8_t* s_ringl = s_ringh + NRH * SLOT_BYTES;
  uint8_t* s_p = s_ringl + NRL * SLOT_BYTES;
  uint8_t* s_ds = s_p + PD_BYTES;
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_ds + PD_BYTES);
  float* s_tab = reinterpret_cast<float*>(s_lut + 128);
  float* s_bacc = s_tab + TAB_MAX + 1;            // [2][TAB_MAX + 1]: one bin array per head of the group
  float* s_delta = s_bacc + 2 * (TAB_MAX + 1);          // [2][128] partial row sums of the two compute groups
  static_assert(((128 + 3 * (TAB_MAX + 1) + 256) * 4) % 8 == 0, "mbarriers need 8-byte alignment");
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + 256);
  uint64_t* fullh = bars;
  uint64_t* emptyh = fullh + NRH;
  uint64_t* fulll = emptyh + NRH;
  uint64_t* emptyl = fulll + NRL;
  uint64_t* sdp_full = emptyl + NRL;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_full + 2;
  uint64_t* pds_free = sdp_full + 3;
  uint64_t* obuf_full = sdp_full + 4;   // [2]
  uint64_t* obuf_free = sdp_full + 6;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = gm.num_tiles * gm.ngrp;     // work item = (tile, head group)
  const int nc = gm.nc;
  const int hg = blockIdx.x % gm.ngrp;          // head group, constant per CTA: gridDim.x % ngrp == 0
  constexpr int SH = SHT;                       // heads per group (2 when head_dim is 32)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv.full);
    tma_prefetch_desc(&tm_do.full);
    for (int i = 0; i < NRH; ++i) {
      mbar_init(&fullh[i], 1);
      mbar_init(&emptyh[i], 1);
    }
    for (int i = 0; i < NRL; ++i) {
      mbar_init(&fulll[i], 1);
      mbar_init(&emptyl[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, (L >= 64) ? 256 : 128);     // every thread that reads S / dP
    mbar_init(pds_full, (L >= 64) ? 256 : 128);     // every thread that writes P / dS
    mbar_init(pds_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&obuf_full[i], 1);
      mbar_init(&obuf_free[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // P / dS: entries outside a row's own window stay zero.  Rings: the padding rows of a tile (general
  // mode) are never written by TMA and must read as zero.
  for (int i = threadIdx.x; i < ((NRH + NRL) * SLOT_BYTES + 2 * PD_BYTES) / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(s_ringh)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_dP = tmem_base + 128;
  const uint32_t tmem_out = tmem_base + 256;    // two 64-column output accumulators

  if (warp == 0 || warp == 2) {
    // ---------------------------------------------------------------- TMA producers (all lanes issue boxes)
    const bool ring_h = (warp == 0);
    uint8_t* ring = ring_h ? s_ringh : s_ringl;
    uint64_t* fullb = ring_h ? fullh : fulll;
    uint64_t* emptyb = ring_h ? emptyh : emptyl;
    const int nslot = ring_h ? NRH : NRL;
    int slot = 0;
    uint32_t phase = 0;
    auto load = [&](int tile, bool from_do, int ch0) {
      mbar_wait(&emptyb[slot], phase ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&fullb[slot], chunk_tx_bytes(gm));
      __syncwarp();
      tile_boxes<true>(gm, tile, ch0, ring + slot * SLOT_BYTES, from_do ? &tm_do : &tm_qkv, &fullb[slot], lane);
      if (++slot == nslot) { slot = 0; phase ^= 1; }
    };
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / gm.ngrp;
      const int hq = hg * gm.gch;
      if (ring_h) {
        for (int c = 0; c < nc; ++c) {
          load(tile, false, 0 * gm.C + hq + c * 64);   // Q_c
          load(tile, false, 1 * gm.C + hq + c * 64);   // K_c
          load(tile, true, hq + c * 64);               // dO_c
          load(tile, false, 2 * gm.C + hq + c * 64);   // V_c
        }
      } else {
        for (int c = 0; c < nc; ++c) {
          load(tile, true, hq + c * 64);               // dO_c -> dV_c
          load(tile, false, 1 * gm.C + hq + c * 64);   // K_c  -> dQ_c
          load(tile, false, 0 * gm.C + hq + c * 64);   // Q_c  -> dK_c
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
      constexpr uint32_t idesc_kn = umma_idesc_bf16(128, 64, 0, 1);    // dQ = dS K
      constexpr uint32_t idesc_nn = umma_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q
      int slot = 0, slotl = 0;
      uint32_t phase = 0, phasel = 0, sub_phase = 0;
      uint32_t obp[2] = {0, 0};           // per output buffer: parity of its next use
      const uint32_t p_addr = smem_u32(s_p), ds_addr = smem_u32(s_ds);
      auto take_h = [&]() { const int sl = slot; mbar_wait(&fullh[slot], phase); if (++slot == NRH) { slot = 0; phase ^= 1; } return sl; };
      auto take_l = [&]() { const int sl = slotl; mbar_wait(&fulll[slotl], phasel); if (++slotl == NRL) { slotl = 0; phasel ^= 1; } return sl; };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int sa[4][2], sb[4][2], sl3[4][3];          // ring slots of this item's chunks (held across sub-heads)
        for (int sub = 0; sub < SH; ++sub, sub_phase ^= 1) {
          mbar_wait(sdp_free, sub_phase ^ 1);
          tc_fence_after();
          // S and dP; with two heads per chunk only the sub-head's 32-channel K range (two k-steps)
          const int k0 = (SH == 1) ? 0 : sub * 2, k1 = (SH == 1) ? 4 : sub * 2 + 2;
          for (int c = 0; c < nc; ++c) {
            for (int pair = 0; pair < 2; ++pair) {     // (Q_c, K_c) -> S ; (dO_c, V_c) -> dP
              if (sub == 0) { sa[c][pair] = take_h(); sb[c][pair] = take_h(); }
              tc_fence_after();
              const uint32_t aa = smem_u32(s_ringh + sa[c][pair] * SLOT_BYTES), ba = smem_u32(s_ringh + sb[c][pair] * SLOT_BYTES);
              for (int kk = k0; kk < k1; ++kk)
                umma_bf16(pair == 0 ? tmem_S : tmem_dP, umma_smem_desc(aa + kk * 32, 16, 1024),
                          umma_smem_desc(ba + kk * 32, 16, 1024), idesc_kk, (c > 0 || kk > k0) ? 1u : 0u);
              if (sub == SH - 1) {
                umma_commit(&emptyh[sa[c][pair]]);
                umma_commit(&emptyh[sb[c][pair]]);
              }
            }
          }
          umma_commit(sdp_full);

          mbar_wait(pds_full, sub_phase);
          tc_fence_after();
          for (int c = 0; c < nc; ++c) {
            for (int o = 0; o < 3; ++o) {              // dV_c, dQ_c, dK_c (whole 64-channel chunk)
              const int ob = ((sub * nc + c) * 3 + o) & 1;   // output chunk k of a tile uses TMEM buffer k % 2
              if (sub == 0) sl3[c][o] = take_l();
              mbar_wait(&obuf_free[ob], obp[ob] ^ 1);
              tc_fence_after();
              const uint32_t xa = smem_u32(s_ringl + sl3[c][o] * SLOT_BYTES);
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
              if (sub == SH - 1) umma_commit(&emptyl[sl3[c][o]]);
              umma_commit(&obuf_full[ob]);
              obp[ob] ^= 1;
            }
          }
          umma_commit(pds_free);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax-backward + epilogue (2 groups)
    const int grp = (warp - 4) >> 2;              // 0: warps 4-7, 1: warps 8-11
    const int wq = warp & 3;                      // TMEM lane quarter
    const int row = wq * 32 + lane;
    const int cm_tid = threadIdx.x - 128;         // 0..255 over both groups
    const uint32_t t_lane = uint32_t(wq * 32) << 16;
    const int nbias = (2 * gm.ws - 1) * (2 * gm.ws - 1);
    for (int i = cm_tid; i < 2 * (TAB_MAX + 1); i += 256) s_bacc[i] = 0.f;
    constexpr bool COLSPLIT = (L >= 64);          // both groups work on S / dP (alternate 32-column chunks)
    constexpr int CH = (L >= 32) ? 32 : 16;
    constexpr int NCHUNK = L / CH;                // chunks of the row's own window columns
    // running sums over tiles of dS[row, key position], one accumulator per key position this
    // thread sees.  L = 128: chunks cb and cb+2 of a thread hold the same positions (row-major order:
    // the other frame; quadrant order folds the two frames inside each 32-column quadrant block).
    constexpr int NACC = COLSPLIT ? 32 : L;
    const bool quad = gm.shift > 0;               // uniform_quad: one token order per launch
    float bacc[SH][NACC];                         // one set per head of the group
#pragma unroll
    for (int h = 0; h < SH; ++h)
#pragma unroll
      for (int k = 0; k < NACC; ++k) bacc[h][k] = 0.f;
    const bool softmax_role = COLSPLIT || grp == 0;
    uint32_t sub_phase = 0, ob_phase = 0;
    int key_i = 0, col0 = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / gm.ngrp;
      const RowGeom rg = row_geom(gm, tile, row);
#pragma unroll
     for (int sub = 0; sub < SH; ++sub, sub_phase ^= 1) {
      const int head = hg * SH + sub;
      named_bar_sync(1, 256);                     // everybody is done with the previous LUT / bias table
      // key | region id | spatial position | window tag (g + 1, 0 for a padding row)
      const uint32_t my_tag = rg.inrange ? uint32_t(rg.g + 1) : 0u;
      if (grp == 0) s_lut[row] = uint32_t(rg.rr * (2 * gm.ws - 1) + rg.cc) | (uint32_t(rg.id) << 8) |
                                 (uint32_t(rg.rr * gm.ws + rg.cc) << 16) | (my_tag << 24);
      if (SH > 1 || item == int(blockIdx.x))      // the head (and so the table) changes only when SH > 1
        for (int i = cm_tid; i < nbias; i += 256) s_tab[i] = __ldg(bias_table + i * gm.nH + head) * 1.4426950408889634f;
      named_bar_sync(1, 256);
      key_i = (rg.rr + gm.ws - 1) * (2 * gm.ws - 1) + rg.cc + gm.ws - 1;
      col0 = GEN ? 0 : rg.g * L;
      const bool use_mask = rg.wraps;
      // dense mask row of this query token (stand-alone WindowAttention with an explicit mask tensor)
      const float* mask_row = gm.mask ? gm.mask + ((size_t)(rg.gw % gm.mask_nw) * gm.N + (rg.rr * gm.ws + rg.cc)) * gm.N : nullptr;

      if (softmax_role) {
        const float lse_i = (GEN && !rg.inrange) ? 0.f : lse2[((size_t)tile * gm.nH + head) * 128 + rg.canon];
        mbar_wait(sdp_full, sub_phase);
        mbar_wait(pds_free, sub_phase ^ 1);
        tc_fence_after();
        // pass 1: P = exp2(S2 - lse2) -> smem (bf16), delta = sum_j P dP
        float delta = 0.f;
#pragma unroll
        for (int ci = 0; ci < (COLSPLIT ? NCHUNK / 2 : NCHUNK); ++ci) {
          const int cb = COLSPLIT ? 2 * ci + grp : ci;
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
                if (use_mask && ((lj >> 8) & 0xffu) != uint32_t(rg.id)) x += kMaskLog2e;
                if (mask_row != nullptr) x = fmaf(__ldg(mask_row + ((lj >> 16) & 0xffu)), 1.4426950408889634f, x);
                pv[e] = fast_exp2(x - lse_i);
                // general geometry: columns of another window (or padding) of this tile do not attend
                if (GEN && ((lj >> 24) != my_tag || my_tag == 0u)) pv[e] = 0.f;
                delta = fmaf(pv[e], __uint_as_float(w[jj]), delta);
              }
              pk[h] = pack_bf16(pv[0], pv[1]);
            }
            const int col = col0 + cb * CH + j8 * 8;
            *reinterpret_cast<uint4*>(s_p + (col >> 6) * SLOT_BYTES + sw128_offset(row, (col & 63) >> 3)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        if (COLSPLIT) {                            // row sum = own half + the other group's half
          s_delta[grp * 128 + row] = delta;
          named_bar_sync(2, 256);
          delta += s_delta[(grp ^ 1) * 128 + row];
        }
        // pass 2: dS = P o (dP - delta) -> smem (bf16)
#pragma unroll
        for (int ci = 0; ci < (COLSPLIT ? NCHUNK / 2 : NCHUNK); ++ci) {
          const int cb = COLSPLIT ? 2 * ci + grp : ci;
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
              const int j = j8 * 8 + 2 * h;          // column inside the chunk, compile-time after unrolling
              if (GEN) {
                // irregular order: no static column -> position map, go through the shared bins
                const uint32_t l0 = s_lut[col0 + cb * CH + j], l1 = s_lut[col0 + cb * CH + j + 1];
                if (d0 != 0.f) atomicAdd(&s_bacc[sub * (TAB_MAX + 1) + key_i - int(l0 & 0xff)], d0);
                if (d1 != 0.f) atomicAdd(&s_bacc[sub * (TAB_MAX + 1) + key_i - int(l1 & 0xff)], d1);
              } else if (COLSPLIT) {
                // row-major order: chunk = (frame, half of the 64 positions) -> position j of that half
                // quadrant order : chunk = quadrant, 16 positions x 2 frames  -> (which quadrant)*16 + j%16
                if (quad && L == 128) { bacc[sub][ci * 16 + (j & 15)] += d0; bacc[sub][ci * 16 + ((j + 1) & 15)] += d1; }
                else                  { bacc[sub][j] += d0; bacc[sub][j + 1] += d1; }
              } else {
                bacc[sub][ci * CH + j] += d0; bacc[sub][ci * CH + j + 1] += d1;
              }
            }
            *reinterpret_cast<uint4*>(s_ds + off) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
          }
        }
        tc_fence_before();
        mbar_arrive(sdp_free);          // S / dP may be overwritten by the next tile
        fence_proxy_async_smem();
        mbar_arrive(pds_full);          // P / dS visible to the tensor core
      }

      // ---- drain dV_c, dQ_c, dK_c : TMEM -> bf16 -> this row's line of d_qkv
      //      (window_reverse + inverse roll = the token index of the row).  Output chunk k of the tile
      //      lands in TMEM buffer k % 2, which is group (k % 2)'s to drain.  With two heads per chunk
      //      (SH == 2) only columns [sub*32, +32) of the product belong to this sub-head.
      __nv_bfloat16* row_out = d_qkv + rg.tok * (3 * gm.C) + hg * gm.gch;
      for (int k = sub * nc * 3 + ((sub * nc * 3 + grp) & 1); k < (sub + 1) * nc * 3; k += 2) {
        if ((k & 1) != grp) continue;
        const int c = (k / 3) - sub * nc, o = k % 3;
        const int which = (o == 0) ? 2 : (o == 1 ? 0 : 1);
        const float m2 = !rg.valid ? 0.f : (o == 0 ? 1.0f : gm.scale);
        mbar_wait(&obuf_full[grp], ob_phase);
        ob_phase ^= 1;
        tc_fence_after();
        constexpr int OW = (SH == 1) ? 64 : 32;         // valid output columns
        float a[64];
        {
          uint32_t v0[32], v1[32];
          tmem_ld32(tmem_out + grp * 64 + t_lane + (SH == 1 ? 0 : sub * 32), v0);
          if (SH == 1) tmem_ld32(tmem_out + grp * 64 + t_lane + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) { a[q] = __uint_as_float(v0[q]) * m2; a[32 + q] = (SH == 1) ? __uint_as_float(v1[q]) * m2 : 0.f; }
        }
        tc_fence_before();
        mbar_arrive(&obuf_free[grp]);
        const int ch0 = which * gm.C + c * 64 + (SH == 1 ? 0 : sub * 32);
        if (rg.valid) {
          uint4* dst = reinterpret_cast<uint4*>(row_out + ch0);
#pragma unroll
          for (int j = 0; j < OW / 8; ++j)
            dst[j] = make_uint4(pack_bf16(a[8 * j], a[8 * j + 1]), pack_bf16(a[8 * j + 2], a[8 * j + 3]),
                                pack_bf16(a[8 * j + 4], a[8 * j + 5]), pack_bf16(a[8 * j + 6], a[8 * j + 7]));
        }
        if (d_colsum != nullptr) {
          if (SH == 1) {
            warp_colsum64(a, lane);          // lane l: columns 2l, 2l+1 summed over this warp's 32 rows
            atomicAdd(d_colsum + ch0 + hg * gm.gch + 2 * lane, a[0]);
            atomicAdd(d_colsum + ch0 + hg * gm.gch + 2 * lane + 1, a[1]);
          } else {
            warp_colsum64(a, lane);          // upper 32 values are zero: lanes 0-15 hold the 32 column sums
            if (lane < 16) {
              atomicAdd(d_colsum + ch0 + hg * gm.gch + 2 * lane, a[0]);
              atomicAdd(d_colsum + ch0 + hg * gm.gch + 2 * lane + 1, a[1]);
            }
          }
        }
      }
     }   // sub
    }
    // ---- bias-table gradient: bin the per-row running sums by relative position, once per CTA and head
    named_bar_sync(1, 256);
    if (!GEN && softmax_role) {
#pragma unroll
      for (int sub = 0; sub < SH; ++sub)
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          // a column of this row's window that accumulator k stands for
          int ck;
          if (COLSPLIT) ck = (quad && L == 128) ? ((2 * (k >> 4) + grp) * 32 + (k & 15)) : (grp * 32 + k);
          else          ck = k;
          const uint32_t lj = s_lut[col0 + ck];
          atomicAdd(&s_bacc[sub * (TAB_MAX + 1) + key_i - int(lj & 0xff)], bacc[sub][k]);
        }
    }
    named_bar_sync(1, 256);
    for (int i = cm_tid; i < SH * nbias; i += 256) {
      const int sub = i / nbias, bin = i - sub * nbias;
      atomicAdd(d_table + bin * gm.nH + hg * SH + sub, s_bacc[sub * (TAB_MAX + 1) + bin]);
    }
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
                float qk_scale, const float* mask, int mask_windows, cudaStream_t stream) {
  STSWIN_CHECK_ARG(qkv && bias_table && lse2 && d_out && d_qkv && d_table, "winattn_bwd: null pointer");
  WinGeom gm;
  int rc = fill_geom(&gm, B, T, H, W, C, nH, ws, shift);
  if (rc != kOk) return rc;
  gm.uniform_quad = 1;
  if (qk_scale > 0.f) { gm.scale = qk_scale; gm.scale_log2e = qk_scale * 1.4426950408889634f; }
  STSWIN_CHECK_ARG(mask == nullptr || mask_windows > 0, "winattn: mask given with mask_windows <= 0");
  gm.mask = mask; gm.mask_nw = mask_windows;
  if (gm.shift > 0 && gm.ra != gm.rb) gm.general = 1;     // unequal rectangles: no static column -> position fold
  WinMaps tq, td;
  if ((rc = make_window_tmaps(&tq, qkv, gm, 3 * C)) != kOk) return rc;
  if ((rc = make_window_tmaps(&td, d_out, gm, C)) != kOk) return rc;
  STSWIN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_qkv) & 15) == 0, "winattn_bwd: d_qkv must be 16-byte aligned");
  const int items = gm.num_tiles * gm.ngrp;
  int grid = items < num_sms() ? items : num_sms();
  grid -= grid % gm.ngrp;                 // one head group per CTA (items is a multiple of ngrp, so grid >= ngrp)
  if (grid < gm.ngrp) return set_error(kErrUnsupported, "winattn_bwd: %d head groups exceed the SM count", gm.ngrp);
#define STSWIN_LAUNCH_BWD(LL, SS, GG)                                                                          \
  {                                                                                                            \
    if ((rc = set_smem_bwd(winattn_bwd_kernel<LL, SS, GG>, SMEM_BYTES)) != kOk) return rc;                     \
    winattn_bwd_kernel<LL, SS, GG><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(                                 \
        tq, td, static_cast<__nv_bfloat16*>(d_qkv), bias_table, lse2, d_table, d_qkv_colsum, gm);              \
  }
#define STSWIN_LAUNCH_BWD_L(LL, GG)                         \
  {                                                         \
    if (gm.SH == 1) STSWIN_LAUNCH_BWD(LL, 1, GG)            \
    else STSWIN_LAUNCH_BWD(LL, 2, GG)                       \
  }
  if (gm.general) STSWIN_LAUNCH_BWD_L(128, true)
  else if (gm.L == 16) STSWIN_LAUNCH_BWD_L(16, false)
  else if (gm.L == 32) STSWIN_LAUNCH_BWD_L(32, false)
  else if (gm.L == 64) STSWIN_LAUNCH_BWD_L(64, false)
  else STSWIN_LAUNCH_BWD_L(128, false)
#undef STSWIN_LAUNCH_BWD_L
#undef STSWIN_LAUNCH_BWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
