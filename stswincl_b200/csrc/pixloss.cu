// K7: label-guided pixel contrastive loss (pixcontrast_18/contrast/models/PixPro_swin_v5.py:48-129).
//
// Reference op sequence per regression_loss call: 5 x bmm(q^T, key) -> five [N,HW,HW] logits,
// 10 one-hot bmm masks (posMask / negMask :48-69), 10 mask*logit products, row sums / divides
// (:119-123), exp / log / mean (:124-128).  Here:
//
//   pix_normalize   F.normalize(dim=1) (:330,362,...) fused with the bf16 cast and the per-channel
//                   key sums the backward needs (one pass over each embedding map)
//   pixloss_fwd     per (sample, 128-query block, key set): similarity tiles q^T k on tcgen05
//                   (both operands channel-major = MN-major, straight from [N,C,HW]); the epilogue
//                   compares labels in registers and keeps four running sums per query pixel
//                   (sum / count of same-label and different-label similarities).  No HWxHW
//                   tensor is ever written.
//   pixloss_finalize  P, N, -log(e^P/(e^P+e^N)+1e-6), mean; and the per-row coefficients of the
//                   backward (dloss/dz takes two values per row and key set)
//   pixloss_bwd     dq = sum_s [ (a - b_s) * M_s K_s^T + b_s * colsum(K_s) ]  with M_s the 0/1
//                   same-label matrix generated tile by tile into shared memory as the bf16 A
//                   operand (exact), K_s streamed by TMA, fp32 TMA add-reduction into dq.
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace stswin {

namespace {

constexpr int MAX_SETS = 8;
constexpr float kEpsCnt = 1e-6f;   // PixPro_swin_v5.py:119-123
constexpr float kEpsLog = 1e-6f;   // PixPro_swin_v5.py:127-128
constexpr float kEpsNorm = 1e-12f; // F.normalize default

struct SetMaps {
  CUtensorMap m[MAX_SETS];
};
struct SetPtrs {
  const uint8_t* lk[MAX_SETS];
};

// ------------------------------------------------------------------------------------------------
// normalise + cast + per-channel sums.  CTA = 32 consecutive pixels of one sample x all channels:
// warp w owns channels w, w+8, ... and keeps eight independent loads in flight (a first version
// walked the channels with one dependent load per iteration and 128 threads per CTA: 170 us per
// launch at N = 32, i.e. 0.5 % of the HBM roofline).
constexpr int PN_WARPS = 8;
template <typename TI>
__global__ void __launch_bounds__(32 * PN_WARPS)
pix_normalize_kernel(const TI* __restrict__ x, __nv_bfloat16* __restrict__ xn, float* __restrict__ inv_norm,
                     float* __restrict__ ksum, int C, int HW, int do_normalize) {
  __shared__ float s_ss[PN_WARPS][32];
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  const bool ok = j < HW;
  const TI* xb = x + (size_t)n * C * HW;
  float inv = 1.f;
  if (do_normalize) {
    float ss = 0.f;
    for (int c0 = warp; c0 < C; c0 += 8 * PN_WARPS) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * PN_WARPS;
        v[u] = (ok && c < C) ? static_cast<float>(xb[(size_t)c * HW + j]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) ss = fmaf(v[u], v[u], ss);
    }
    s_ss[warp][lane] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < PN_WARPS; ++w) tot += s_ss[w][lane];
    inv = 1.0f / fmaxf(sqrtf(tot), kEpsNorm);
  }
  if (warp == 0 && ok && inv_norm != nullptr) inv_norm[(size_t)n * HW + j] = inv;
  for (int c0 = warp; c0 < C; c0 += 8 * PN_WARPS) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u * PN_WARPS;
      v[u] = (ok && c < C) ? static_cast<float>(xb[(size_t)c * HW + j]) : 0.f;     // second pass: L1 / L2 hits
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u * PN_WARPS;
      if (c < C) {
        const __nv_bfloat16 b = __float2bfloat16_rn(v[u] * inv);
        if (ok) xn[((size_t)n * C + c) * HW + j] = b;
        if (ksum != nullptr) {                 // sum what the tensor core will read (bf16-rounded)
          float w = ok ? __bfloat162float(b) : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
          if (lane == 0) atomicAdd(ksum + (size_t)n * C + c, w);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward
constexpr int PF_STAGES = 6;
constexpr int KB_BYTES = 64 * 128 * 2;      // one 64-channel k-block of a 128-pixel tile: 2 chunks x 8 KB
constexpr int PF_THREADS = 192;

struct PixFwdArgs {
  int N, C, HW, n_sets, num_mb, num_tiles, nkb;
  int set_total, set_off;   // this launch handles sets [set_off, set_off + n_sets) of set_total
  const uint8_t* lq;
  float* stats;      // [N, HW, n_sets, 4] : pos_sum, pos_cnt, neg_sum, neg_cnt
};

__global__ void __launch_bounds__(PF_THREADS, 1)
pixloss_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ SetMaps tm_k, const SetPtrs lk,
                   const PixFwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_a = smem;                                   // nkb k-blocks of the query tile (<= 4 x 16 KB)
  uint8_t* s_b = s_a + 4 * KB_BYTES;                     // ring
  uint8_t* s_lk = s_b + PF_STAGES * KB_BYTES;            // labels of the key set, padded to a tile multiple
  const int lk_bytes = p.num_tiles * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lk + ((lk_bytes + 15) & ~15));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + PF_STAGES;
  uint64_t* a_full = bars + 2 * PF_STAGES;
  uint64_t* a_free = a_full + 1;
  uint64_t* acc_full = a_full + 2;    // [2]
  uint64_t* acc_empty = a_full + 4;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.num_mb * p.n_sets;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    for (int i = 0; i < PF_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, itp = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, itp ^= 1) {
        const int s = item % p.n_sets;
        const int mb = (item / p.n_sets) % p.num_mb;
        const int n = item / (p.n_sets * p.num_mb);
        mbar_wait(a_free, itp ^ 1);
        mbar_arrive_expect_tx(a_full, p.nkb * KB_BYTES);
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int c2 = 0; c2 < 2; ++c2)
            tma_load_3d(s_a + kb * KB_BYTES + c2 * 8192, &tm_q, a_full, mb * 128 + c2 * 64, kb * 64, n);
        for (int t = 0; t < p.num_tiles; ++t)
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], KB_BYTES);
            for (int c2 = 0; c2 < 2; ++c2)
              tma_load_3d(s_b + stage * KB_BYTES + c2 * 8192, &tm_k.m[s], &full_bar[stage], t * 128 + c2 * 64, kb * 64, n);
            if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    {   // warp-uniform control flow, one elected lane issues (descriptors stay in uniform registers)
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, itp = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, itp ^= 1) {
        mbar_wait(a_full, itp);
        for (int t = 0; t < p.num_tiles; ++t) {
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t aa = smem_u32(s_a + kb * KB_BYTES), ba = smem_u32(s_b + stage * KB_BYTES);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tmem_base + acc * 128, umma_smem_desc(aa + kk * 2048, 8192, 1024),
                          umma_smem_desc(ba + kk * 2048, 8192, 1024), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader) umma_commit(&acc_full[acc]);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) umma_commit(a_free);
        __syncwarp();
      }
    }
  } else {
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = uint32_t(wq * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int s = item % p.n_sets;
      const int mb = (item / p.n_sets) % p.num_mb;
      const int n = item / (p.n_sets * p.num_mb);
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous item's readers of s_lk are done
      for (int j = tid; j < lk_bytes; j += 128) s_lk[j] = (j < p.HW) ? lk.lk[s][(size_t)n * p.HW + j] : uint8_t(255);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int i = mb * 128 + row;
      const uint32_t li = (i < p.HW) ? p.lq[(size_t)n * p.HW + i] : 254u;
      float pos = 0.f, tot = 0.f;
      int cnt = 0;
      for (int t = 0; t < p.num_tiles; ++t) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + t_lane + acc * 128 + cb * 32, v);
          const uint4 l0 = *reinterpret_cast<const uint4*>(s_lk + t * 128 + cb * 32);
          const uint4 l1 = *reinterpret_cast<const uint4*>(s_lk + t * 128 + cb * 32 + 16);
          const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const uint32_t lj = (lw[jj >> 2] >> ((jj & 3) * 8)) & 0xffu;
            const float z = __uint_as_float(v[jj]);
            tot += z;
            if (lj == li) { pos += z; ++cnt; }
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (i < p.HW) {
        float4 o;
        o.x = pos; o.y = float(cnt); o.z = tot - pos; o.w = float(p.HW - cnt);
        *reinterpret_cast<float4*>(p.stats + (((size_t)n * p.HW + i) * p.set_total + p.set_off + s) * 4) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: loss and backward coefficients.  one thread per query pixel.
__global__ void pixloss_finalize_kernel(const float* __restrict__ stats, int n_sets, long rows, float inv_rows,
                                        float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ float s_part[8];
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float li = 0.f;
  if (r < rows) {
    float psum = 0.f, pcnt = 0.f, nterm = 0.f;
    for (int s = 0; s < n_sets; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(stats + (r * n_sets + s) * 4);
      psum += v.x; pcnt += v.y;
      nterm += v.z / (v.w + kEpsCnt);
    }
    const float P = psum / (pcnt + kEpsCnt);
    const float eP = expf(P), eN = expf(nterm);
    const float ratio = eP / (eP + eN);
    li = -logf(ratio + kEpsLog);
    if (coef != nullptr) {
      // d li / dP = -ratio (1 - ratio) / (ratio + eps) ; d li / dN = + the same
      const float gP = -ratio * (1.f - ratio) / (ratio + kEpsLog) * inv_rows;
      coef[r * (n_sets + 1)] = gP / (pcnt + kEpsCnt);
      for (int s = 0; s < n_sets; ++s) {
        const float ncnt = stats[(r * n_sets + s) * 4 + 3];
        // a key set without any different-label pixel contributes the constant 0 / (0 + eps): its
        // gradient is exactly zero (and 1/eps here would amplify rounding noise a million-fold)
        coef[r * (n_sets + 1) + 1 + s] = ncnt > 0.f ? -gP / (ncnt + kEpsCnt) : 0.f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = li;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_part[w];
    atomicAdd(loss, t * inv_rows);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
constexpr int PB_GEN = 3;                    // generated-mask buffers (128 rows x 64 keys, 16 KB)
constexpr int PB_STAGES = 4;                 // key k-blocks: C rows x 64 keys
constexpr int PB_THREADS = 192;

struct PixBwdArgs {
  int N, C, HW, n_sets, num_mb, nkb;   // nkb = ceil(HW / 64)
  int set_total, set_off;
  const uint8_t* lq;
  const float* coef;     // [N, HW, 1 + n_sets]
  const float* ksum;     // [n_sets, N, C]
  const float* d_loss;   // device scalar (upstream gradient)
};

__global__ void __launch_bounds__(PB_THREADS, 1)
pixloss_bwd_kernel(const __grid_constant__ SetMaps tm_k, const __grid_constant__ CUtensorMap tm_dq, const SetPtrs lk,
                   const PixBwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int b_stage_bytes = p.C * 128;
  uint8_t* s_gen = smem;
  uint8_t* s_b = s_gen + PB_GEN * 16384;
  uint8_t* s_stg = s_b + PB_STAGES * b_stage_bytes;        // 4 warps x 4 KB
  uint8_t* s_lk = s_stg + 4 * 4096;
  const int lk_bytes = p.nkb * 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lk + ((lk_bytes + 15) & ~15));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + PB_STAGES;
  uint64_t* gen_full = bars + 2 * PB_STAGES;
  uint64_t* gen_empty = gen_full + PB_GEN;
  uint64_t* acc_full = gen_empty + PB_GEN;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.num_mb * p.n_sets;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dq);
    for (int i = 0; i < PB_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < PB_GEN; ++i) { mbar_init(&gen_full[i], 128); mbar_init(&gen_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int s = item % p.n_sets;
        const int n = item / (p.n_sets * p.num_mb);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], b_stage_bytes);
          tma_load_3d(s_b + stage * b_stage_bytes, &tm_k.m[s], &full_bar[stage], kb * 64, 0, n);
          if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // warp-uniform control flow, one elected lane issues
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(128, p.C, 0, 0);
      int stage = 0, gb = 0, acc = 0;
      uint32_t phase = 0, gphase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&gen_full[gb], gphase);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t aa = smem_u32(s_gen + gb * 16384), ba = smem_u32(s_b + stage * b_stage_bytes);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem_base + acc * 256, umma_smem_desc(aa + kk * 32, 16, 1024), umma_smem_desc(ba + kk * 32, 16, 1024),
                        idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&gen_empty[gb]);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++gb == PB_GEN) { gb = 0; gphase ^= 1; }
          if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&acc_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = uint32_t(wq * 32) << 16;
    uint8_t* my_stg = s_stg + wq * 4096;
    const float g_up = __ldg(p.d_loss);
    int gb = 0, acc = 0;
    uint32_t gphase = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int s = item % p.n_sets;
      const int mb = (item / p.n_sets) % p.num_mb;
      const int n = item / (p.n_sets * p.num_mb);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = tid; j < lk_bytes; j += 128) s_lk[j] = (j < p.HW) ? lk.lk[s][(size_t)n * p.HW + j] : uint8_t(255);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int i = mb * 128 + row;
      const bool row_ok = i < p.HW;
      const uint32_t li = row_ok ? p.lq[(size_t)n * p.HW + i] : 254u;
      float ca = 0.f, cb_ = 0.f;
      if (row_ok) {
        const float* cf = p.coef + ((size_t)n * p.HW + i) * (p.set_total + 1);
        ca = cf[0] * g_up;
        cb_ = cf[1 + p.set_off + s] * g_up;
      }
      // ---- generate the 0/1 same-label operand, 64 keys at a time
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&gen_empty[gb], gphase ^ 1);
        uint8_t* dst = s_gen + gb * 16384;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint2 l8 = *reinterpret_cast<const uint2*>(s_lk + kb * 64 + c * 8);
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const uint32_t src = (h < 2) ? l8.x : l8.y;
            const uint32_t la = (src >> (((2 * h) & 3) * 8)) & 0xffu, lb = (src >> (((2 * h + 1) & 3) * 8)) & 0xffu;
            w[h] = (la == li ? 0x3F80u : 0u) | (lb == li ? 0x3F800000u : 0u);
          }
          *reinterpret_cast<uint4*>(dst + sw128_offset(row, c)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(&gen_full[gb]);
        if (++gb == PB_GEN) { gb = 0; gphase ^= 1; }
      }
      // ---- epilogue: dq[i, :] += (a - b_s) * D[i, :] + b_s * ksum_s[:]
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const float* ks = p.ksum + ((size_t)(p.set_off + s) * p.N + n) * p.C;
      const float cm = ca - cb_;
#pragma unroll 1
      for (int c32 = 0; c32 < p.C / 32; ++c32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + t_lane + acc * 256 + c32 * 32, v);
        const float ksl = __ldg(ks + c32 * 32 + lane);
        tmem_ld_wait();
        if (lane == 0) tma_wait_group_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = fmaf(cm, __uint_as_float(v[4 * j + e]), cb_ * __shfl_sync(0xffffffffu, ksl, 4 * j + e));
          *reinterpret_cast<float4*>(my_stg + sw128_offset(lane, j)) = make_float4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && mb * 128 + wq * 32 < p.HW) {
          tma_reduce_add_3d(&tm_dq, my_stg, c32 * 32, mb * 128 + wq * 32, n);
          tma_commit_group();
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

constexpr int MAX_TOTAL_SETS = 64;   // launches of MAX_SETS sets each (cross-rank gathered key sets)

int check_pix_shape(int n_sets, int N, int C, int HW) {
  STSWIN_CHECK_ARG(n_sets >= 1 && n_sets <= MAX_TOTAL_SETS, "pixloss: n_sets=%d out of range [1,%d]", n_sets, MAX_TOTAL_SETS);
  STSWIN_CHECK_ARG(N > 0 && HW > 0, "pixloss: empty input");
  if (C % 64 != 0 || C > 256) return set_error(kErrUnsupported, "pixloss: C=%d unsupported (multiple of 64, <= 256)", C);
  if (HW % 8 != 0) return set_error(kErrUnsupported, "pixloss: H*W=%d must be a multiple of 8", HW);
  if (HW > 16384) return set_error(kErrUnsupported, "pixloss: H*W=%d > 16384 unsupported", HW);
  return kOk;
}

int key_tmap(CUtensorMap* tm, const void* base, int N, int C, int HW, uint32_t box_c) {
  uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)N};
  uint64_t str[2] = {(uint64_t)HW * 2, (uint64_t)C * HW * 2};
  uint32_t box[3] = {64, box_c, 1};
  return make_tmap(tm, TmapDtype::BF16, 3, base, dims, str, box, true);
}

}  // namespace

// see include/stswin_b200.h : stswin_pix_normalize
int pix_normalize(const void* x, int x_is_f32, void* xn, float* inv_norm, float* ksum, int N, int C, int HW,
                  int do_normalize, cudaStream_t stream) {
  STSWIN_CHECK_ARG(x && xn && N > 0 && C > 0 && HW > 0, "pix_normalize: bad argument");
  STSWIN_CHECK_ARG(C <= 4096, "pix_normalize: C=%d too large", C);
  if (ksum) STSWIN_CUDA(cudaMemsetAsync(ksum, 0, sizeof(float) * (size_t)N * C, stream));
  dim3 grid((HW + 31) / 32, N);
  if (x_is_f32)
    pix_normalize_kernel<float><<<grid, 32 * PN_WARPS, 0, stream>>>(static_cast<const float*>(x), static_cast<__nv_bfloat16*>(xn),
                                                                    inv_norm, ksum, C, HW, do_normalize);
  else
    pix_normalize_kernel<__nv_bfloat16><<<grid, 32 * PN_WARPS, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                            static_cast<__nv_bfloat16*>(xn), inv_norm, ksum, C,
                                                                            HW, do_normalize);
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_pixloss_fwd
int pixloss_fwd(const void* q, const void* const* keys, const uint8_t* lq, const uint8_t* const* lk, int n_sets, int N,
                int C, int HW, float* row_stats, float* loss, float* coef, cudaStream_t stream) {
  STSWIN_CHECK_ARG(q && keys && lq && lk && row_stats && loss, "pixloss_fwd: null pointer");
  int rc = check_pix_shape(n_sets, N, C, HW);
  if (rc != kOk) return rc;
  CUtensorMap tq;
  if ((rc = key_tmap(&tq, q, N, C, HW, 64)) != kOk) return rc;
  for (int off = 0; off < n_sets; off += MAX_SETS) {
    const int ns = (n_sets - off < MAX_SETS) ? n_sets - off : MAX_SETS;
    SetMaps tk;
    SetPtrs lp;
    for (int s = 0; s < MAX_SETS; ++s) {
      const int src = off + (s < ns ? s : 0);
      STSWIN_CHECK_ARG(keys[src] && lk[src], "pixloss_fwd: null key set %d", src);
      if ((rc = key_tmap(&tk.m[s], keys[src], N, C, HW, 64)) != kOk) return rc;
      lp.lk[s] = lk[src];
    }
    PixFwdArgs a;
    a.N = N; a.C = C; a.HW = HW; a.n_sets = ns; a.set_total = n_sets; a.set_off = off;
    a.num_mb = (HW + 127) / 128; a.num_tiles = (HW + 127) / 128; a.nkb = C / 64;
    a.lq = lq; a.stats = row_stats;
    const int smem = 1024 + 4 * KB_BYTES + PF_STAGES * KB_BYTES + ((a.num_tiles * 128 + 15) & ~15) + 256;
    STSWIN_CUDA(cudaFuncSetAttribute(pixloss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int items = N * a.num_mb * ns;
    const int grid = items < num_sms() ? items : num_sms();
    pixloss_fwd_kernel<<<grid, PF_THREADS, smem, stream>>>(tq, tk, lp, a);
    STSWIN_CUDA(cudaGetLastError());
  }
  STSWIN_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  const long rows = (long)N * HW;
  pixloss_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(row_stats, n_sets, rows, 1.0f / rows, loss, coef);
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_pixloss_bwd
int pixloss_bwd(const void* const* keys, const uint8_t* lq, const uint8_t* const* lk, const float* coef,
                const float* ksum, const float* d_loss, int n_sets, int N, int C, int HW, float* dq32,
                cudaStream_t stream) {
  STSWIN_CHECK_ARG(keys && lq && lk && coef && ksum && d_loss && dq32, "pixloss_bwd: null pointer");
  int rc = check_pix_shape(n_sets, N, C, HW);
  if (rc != kOk) return rc;
  CUtensorMap tdq;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)HW * C * 4};
    uint32_t box[3] = {32, 32, 1};
    if ((rc = make_tmap(&tdq, TmapDtype::F32, 3, dq32, dims, str, box, true)) != kOk) return rc;
  }
  STSWIN_CUDA(cudaMemsetAsync(dq32, 0, sizeof(float) * (size_t)N * HW * C, stream));
  for (int off = 0; off < n_sets; off += MAX_SETS) {
    const int ns = (n_sets - off < MAX_SETS) ? n_sets - off : MAX_SETS;
    SetMaps tk;
    SetPtrs lp;
    for (int s = 0; s < MAX_SETS; ++s) {
      const int src = off + (s < ns ? s : 0);
      STSWIN_CHECK_ARG(keys[src] && lk[src], "pixloss_bwd: null key set %d", src);
      if ((rc = key_tmap(&tk.m[s], keys[src], N, C, HW, (uint32_t)C)) != kOk) return rc;
      lp.lk[s] = lk[src];
    }
    PixBwdArgs a;
    a.N = N; a.C = C; a.HW = HW; a.n_sets = ns; a.set_total = n_sets; a.set_off = off;
    a.num_mb = (HW + 127) / 128; a.nkb = (HW + 63) / 64;
    a.lq = lq; a.coef = coef; a.ksum = ksum; a.d_loss = d_loss;
    const int smem = 1024 + PB_GEN * 16384 + PB_STAGES * C * 128 + 4 * 4096 + ((a.nkb * 64 + 15) & ~15) + 256;
    STSWIN_CUDA(cudaFuncSetAttribute(pixloss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int items = N * a.num_mb * ns;
    const int grid = items < num_sms() ? items : num_sms();
    pixloss_bwd_kernel<<<grid, PB_THREADS, smem, stream>>>(tk, tdq, lp, a);
    STSWIN_CUDA(cudaGetLastError());
  }
  return kOk;
}

}  // namespace stswin
