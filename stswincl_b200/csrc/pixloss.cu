// K7: label-guided pixel contrastive loss (pixcontrast_18/contrast/models/PixPro_swin_v5.py:48-129, :584-597).
//
// Reference op sequence per regression_loss call: 5 x bmm(q^T, key) -> five [N,HW,HW] logits,
// 10 one-hot bmm masks (posMask / negMask :48-69), 10 mask*logit products, row sums / divides
// (:119-123), exp / log / mean (:124-128); ConsistencyLoss.forward calls it twice on label maps it
// first down-samples with F.interpolate(mode='nearest') (:585-590).  Here one step of any number of
// "queries" (the two symmetric regression_loss calls of :592-595 are two queries sharing four key
// sets) is six launches, with no host synchronisation (CUDA-graph capturable):
//
//   pix_labels     nearest down-sampling + .long() + range check of every label map (:54,:585-590), and a
//                  STABLE COUNTING SORT of each map's pixels by label: the key pixels of a set are stored
//                  in label order (sums over keys do not depend on their order), so that a 32-key group
//                  of a similarity tile almost always carries ONE label and the label compare of the
//                  epilogue is per group, not per element
//   pix_prepare    F.normalize(dim=1) (:330,362,...) + bf16 cast + scatter into label order + per-channel
//                  key sums, every embedding map of the step in one launch (16-byte loads, 8-byte stores of label runs)
//   pixloss_fwd    similarity tiles q^T k on tcgen05, CTA pairs (cta_group::2): a pair keeps 256 query
//                  pixels resident and streams 256-key tiles (each CTA loads half of the keys), 256 x 256
//                  accumulators double-buffered in TMEM; the epilogue sums each 32-key group with packed
//                  adds and adds it to the row's positive sum when the group's label is the row's.  No
//                  HW x HW tensor is ever written.
//   pixloss_finalize  P, N, -log(e^P/(e^P+e^N)+1e-6), mean (deterministic two-level sum), and the per-row
//                  coefficients of the backward (dloss/dz takes two values per row and key set)
//   pixloss_bwd    dq = sum_s [ (a - b_s) * M_s K_s^T + b_s * colsum(K_s) ], dense on tcgen05 (CTA pairs,
//                  M = 256 query rows x N = C channels): the 0/1 same-label operand M_s is generated into
//                  shared memory by four warps (one 16-byte pattern per 8 keys of a one-label group), the
//                  keys stream by TMA, fp32 TMA add-reduction into dq
//   pix_dq_finish  chain rule through the fused normalisation and the [N,HW,C] -> [N,C,HW] transpose
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace stswin {

namespace {

constexpr int PX_MAX_Q = 2;       // queries per step (the two symmetric calls of ConsistencyLoss.forward)
constexpr int PX_MAX_SETS = 64;   // key sets per query
constexpr int PX_MAX_PTRS = 16;   // source maps / label maps per prepare launch
constexpr float kEpsCnt = 1e-6f;   // PixPro_swin_v5.py:119-123
constexpr float kEpsLog = 1e-6f;   // PixPro_swin_v5.py:127-128
constexpr float kEpsNorm = 1e-12f; // F.normalize default
constexpr int kMixed = 254;        // group label: the 32 keys of the group carry more than one label
constexpr int kPad = 255;          // label of padding keys / rows and of out-of-range labels (never a class)

__host__ __device__ inline int px_hwp(int HW) { return (HW + 255) & ~255; }                 // key columns, padded
__host__ __device__ inline int px_glp(int HW) { return ((px_hwp(HW) / 32) + 15) & ~15; }    // group labels per row

// which prepared map / label slot every (query, key set) uses
struct PixTable {
  int16_t qmap[PX_MAX_Q], qlab[PX_MAX_Q];
  int16_t kmap[PX_MAX_Q][PX_MAX_SETS], klab[PX_MAX_Q][PX_MAX_SETS];
};

// ------------------------------------------------------------------------------------------------
// labels: one CTA of 8 warps per (label map, sample)
struct LabelArgs {
  const void* src[PX_MAX_PTRS];
  uint8_t dtype[PX_MAX_PTRS];     // 0 u8, 1 f32, 2 i64, 3 i32, 4 bf16, 5 f16
  int n_labels, slot_off, N, Hs, Ws, H, W, class_num;
  uint8_t* lab_nat;      // [slots, N, HWp]  labels in pixel order (query side)
  uint8_t* lab_sorted;   // [slots, N, HWp]  labels in sorted key order
  uint8_t* glab;         // [slots, N, GLp]  label of each 32-key group of the sorted order (kMixed / kPad)
  uint16_t* perm;        // [slots, N, HW]   sorted position of pixel j
  int* hist;             // [slots, N, 256]  pixels per label
  int* err;              // |= 1 when a label is outside [0, class_num)
  float* zero;           // optional buffer cleared by this launch (the channel sums the prepare kernel accumulates into)
  long n_zero;
};

constexpr int LB_WARPS = 8;
__global__ void __launch_bounds__(32 * LB_WARPS) pix_labels_kernel(const LabelArgs p) {
  extern __shared__ uint8_t s_dyn[];     // [HWp] natural order, [HWp] sorted order
  __shared__ int s_cnt[LB_WARPS][256];
  __shared__ int s_wsum[LB_WARPS];
  __shared__ int s_bad;
  const int z = blockIdx.y, n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = p.H * p.W, HWp = px_hwp(HW), GLp = px_glp(HW);
  uint8_t* s_nat = s_dyn;
  uint8_t* s_srt = s_dyn + HWp;
  const size_t row = (size_t)(p.slot_off + z) * p.N + n;
  pdl_launch_dependents();                  // the prepare kernel may start its prologue; it waits for this grid before reading
  if (p.zero != nullptr) {
    const long nb = (long)gridDim.x * gridDim.y, b = (long)blockIdx.y * gridDim.x + blockIdx.x;
    for (long i = b * (32 * LB_WARPS) + tid; i < p.n_zero; i += nb * (32 * LB_WARPS)) p.zero[i] = 0.f;
  }
  for (int c = tid; c < LB_WARPS * 256; c += 32 * LB_WARPS) (&s_cnt[0][0])[c] = 0;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  // (1) F.interpolate(mode='nearest') source index (floor(dst * in/out) in fp32, clamped), .long(), range check
  const float sh = (float)p.Hs / (float)p.H, sw = (float)p.Ws / (float)p.W;
  const void* src = p.src[z];
  const int dt = p.dtype[z];
  bool bad = false;
#pragma unroll 4
  for (int j = tid; j < HWp; j += 32 * LB_WARPS) {
    int l = kPad;
    if (j < HW) {
      const int h = j / p.W, w = j - h * p.W;
      const int ih = min((int)floorf(h * sh), p.Hs - 1), iw = min((int)floorf(w * sw), p.Ws - 1);
      const size_t idx = ((size_t)n * p.Hs + ih) * p.Ws + iw;
      long long v;
      switch (dt) {
        case 0: v = static_cast<const uint8_t*>(src)[idx]; break;
        case 1: v = (long long)static_cast<const float*>(src)[idx]; break;
        case 2: v = static_cast<const long long*>(src)[idx]; break;
        case 3: v = static_cast<const int*>(src)[idx]; break;
        case 4: v = (long long)__bfloat162float(static_cast<const __nv_bfloat16*>(src)[idx]); break;
        default: v = (long long)__half2float(static_cast<const __half*>(src)[idx]); break;
      }
      if (v < 0 || v >= p.class_num) bad = true; else l = (int)v;
    }
    s_nat[j] = (uint8_t)l;
    s_srt[j] = (uint8_t)kPad;
  }
  if (bad) s_bad = 1;
  __syncthreads();
  // (2) per-warp histograms over contiguous pixel ranges (keeps the sort stable across warps)
  const int chunk = (((HW + LB_WARPS - 1) / LB_WARPS) + 31) & ~31;
  const int jbeg = warp * chunk, jend = min(HW, jbeg + chunk);
  for (int j0 = jbeg; j0 < jend; j0 += 32) {
    const int j = j0 + lane;
    const bool on = j < jend;
    const unsigned act = __ballot_sync(0xffffffffu, on);
    if (on) {
      const int l = s_nat[j];
      const unsigned m = __match_any_sync(act, l);
      if ((m & ((1u << lane) - 1u)) == 0u) s_cnt[warp][l] += __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  // (3) first sorted position of every (warp, label): thread t owns label t
  {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < LB_WARPS; ++w) tot += s_cnt[w][tid];
    p.hist[row * 256 + tid] = tot;
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int base = incl - tot;
    for (int w = 0; w < warp; ++w) base += s_wsum[w];
#pragma unroll
    for (int w = 0; w < LB_WARPS; ++w) {
      const int c = s_cnt[w][tid];
      s_cnt[w][tid] = base;
      base += c;
    }
  }
  __syncthreads();
  // (4) stable positions
  for (int j0 = jbeg; j0 < jend; j0 += 32) {
    const int j = j0 + lane;
    const bool on = j < jend;
    const unsigned act = __ballot_sync(0xffffffffu, on);
    int l = 0, r = 0, pos = 0;
    unsigned m = 0;
    if (on) {
      l = s_nat[j];
      m = __match_any_sync(act, l);
      r = __popc(m & ((1u << lane) - 1u));
      pos = s_cnt[warp][l] + r;
    }
    __syncwarp();
    if (on && r == 0) s_cnt[warp][l] += __popc(m);
    __syncwarp();
    if (on) {
      p.perm[row * HW + j] = (uint16_t)pos;
      s_srt[pos] = (uint8_t)l;
    }
  }
  __syncthreads();
  // (5) rows out
  for (int j = tid * 16; j < HWp; j += 32 * LB_WARPS * 16) {
    *reinterpret_cast<uint4*>(p.lab_nat + row * HWp + j) = *reinterpret_cast<const uint4*>(s_nat + j);
    *reinterpret_cast<uint4*>(p.lab_sorted + row * HWp + j) = *reinterpret_cast<const uint4*>(s_srt + j);
  }
  for (int g = tid; g < GLp; g += 32 * LB_WARPS) {
    int gl = kPad;
    if (g * 32 < HWp) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(s_srt + g * 32);
      const uint32_t first = (w[0] & 0xffu) * 0x01010101u;
      bool uni = true;
#pragma unroll
      for (int k = 0; k < 8; ++k) uni = uni && (w[k] == first);
      gl = uni ? (int)(w[0] & 0xffu) : kMixed;
    }
    p.glab[row * GLp + g] = (uint8_t)gl;
  }
  if (tid == 0 && s_bad) atomicOr(p.err, 1);
}

// ------------------------------------------------------------------------------------------------
// normalise + cast + scatter into label order + per-channel sums.  CTA = 64 consecutive pixels of one sample of one
// map x all channels: warp w owns channels w, w+8, ..., a lane two adjacent pixels, the whole [C x 64] tile
// lives in registers between the norm and the scaling (one read of the source).
struct PrepArgs {
  const void* x[PX_MAX_PTRS];
  uint8_t dtype[PX_MAX_PTRS];     // 0 bf16, 1 f32, 2 f16
  int8_t lslot[PX_MAX_PTRS];      // label slot whose sort orders this map's pixels; -1 = pixel order (query maps)
  int n_maps, slot_off, N, C, HW, do_normalize;
  int lo_off;                     // > 0: also store the second bf16 term (x - bf16(x)) in slot (slot + lo_off) -- fp32 mode
  const uint16_t* perm;           // [label slots, N, HW]
  __nv_bfloat16* xn;              // [slots, N, C, HW]
  float* inv_norm;                // [slots, N, HW] (pixel order) or null
  float* ksum;                    // [slots, N, C], zero-filled
  float4* zero;                   // optional: the backward's fp32 gradient accumulator, cleared here (no memset node)
  long n_zero4;
};

template <typename TI>
__device__ __forceinline__ float2 ld_pair(const TI* p);
template <>
__device__ __forceinline__ float2 ld_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <>
__device__ __forceinline__ float2 ld_pair<__half>(const __half* p) {
  return __half22float2(*reinterpret_cast<const __half2*>(p));
}

template <typename TI>
__device__ __forceinline__ float4 ld_quad(const TI* p);
template <>
__device__ __forceinline__ float4 ld_quad<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld_quad<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16(r.x), b = unpack_bf16(r.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 ld_quad<__half>(const __half* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// A lane owns FOUR adjacent pixels (lane & 15) of CPW / 2 channels (warp, lane >> 4, stride 16): 16-byte loads, and in
// label order the four pixels usually stay one aligned run (labels are spatially coherent and the sort is stable), so a
// channel's four values leave as one 8-byte store; otherwise as two pairs or four singles.
template <int CPW, typename TI>
__device__ __forceinline__ void prepare_body(const PrepArgs& p, const TI* __restrict__ xb, int z, int n, float (*s_ss)[64]) {
  constexpr int NU = CPW / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pq = lane & 15, cs = lane >> 4;
  const int j0 = blockIdx.x * 64 + pq * 4;
  const bool ok = j0 < p.HW;                     // HW is a multiple of 8: all four pixels or none
  const int slot = p.slot_off + z;
  const int ls = p.lslot[z];
  uint2 pr = make_uint2(0u, 0u);                 // sorted positions of the four pixels (requested before the tile)
  if (ls >= 0 && ok) pr = *reinterpret_cast<const uint2*>(p.perm + ((size_t)ls * p.N + n) * p.HW + j0);
  float4 v[NU];
  float4 ss = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int c = warp * 2 + cs + 16 * u;
    v[u] = ok ? ld_quad<TI>(xb + (size_t)c * p.HW + j0) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss.x = fmaf(v[u].x, v[u].x, ss.x); ss.y = fmaf(v[u].y, v[u].y, ss.y);
    ss.z = fmaf(v[u].z, v[u].z, ss.z); ss.w = fmaf(v[u].w, v[u].w, ss.w);
  }
  float4 inv = make_float4(1.f, 1.f, 1.f, 1.f);
  if (p.do_normalize) {
    ss.x += __shfl_xor_sync(0xffffffffu, ss.x, 16); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, 16);
    ss.z += __shfl_xor_sync(0xffffffffu, ss.z, 16); ss.w += __shfl_xor_sync(0xffffffffu, ss.w, 16);
    if (cs == 0) *reinterpret_cast<float4*>(&s_ss[warp][pq * 4]) = ss;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 a = *reinterpret_cast<const float4*>(&s_ss[w][pq * 4]);
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    }
    inv = make_float4(1.0f / fmaxf(sqrtf(t.x), kEpsNorm), 1.0f / fmaxf(sqrtf(t.y), kEpsNorm),
                      1.0f / fmaxf(sqrtf(t.z), kEpsNorm), 1.0f / fmaxf(sqrtf(t.w), kEpsNorm));
    if (warp == 0 && cs == 0 && ok && p.inv_norm != nullptr)
      *reinterpret_cast<float4*>(p.inv_norm + ((size_t)slot * p.N + n) * p.HW + j0) = inv;
  }
  int d0 = j0, d1 = j0 + 1, d2 = j0 + 2, d3 = j0 + 3;
  if (ls >= 0 && ok) {
    d0 = pr.x & 0xffffu; d1 = pr.x >> 16; d2 = pr.y & 0xffffu; d3 = pr.y >> 16;
  }
  const bool pair_a = ls < 0 || (d1 == d0 + 1 && (d0 & 1) == 0);
  const bool pair_b = ls < 0 || (d3 == d2 + 1 && (d2 & 1) == 0);
  const bool quad = ls < 0 || (pair_a && pair_b && d2 == d0 + 2 && (d0 & 3) == 0);
  __nv_bfloat16* ob = p.xn + ((size_t)slot * p.N + n) * p.C * p.HW;
  auto store4 = [&](__nv_bfloat16* row, __nv_bfloat162 a, __nv_bfloat162 b) {
    if (quad) {
      *reinterpret_cast<uint2*>(row + d0) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    } else {
      if (pair_a) *reinterpret_cast<__nv_bfloat162*>(row + d0) = a; else { row[d0] = a.x; row[d1] = a.y; }
      if (pair_b) *reinterpret_cast<__nv_bfloat162*>(row + d2) = b; else { row[d2] = b.x; row[d3] = b.y; }
    }
  };
  float cs_sum[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) cs_sum[u] = 0.f;
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int c = warp * 2 + cs + 16 * u;
    const float x0 = v[u].x * inv.x, x1 = v[u].y * inv.y, x2 = v[u].z * inv.z, x3 = v[u].w * inv.w;
    const __nv_bfloat162 a = __floats2bfloat162_rn(x0, x1), b = __floats2bfloat162_rn(x2, x3);
    if (ok) {
      store4(ob + (size_t)c * p.HW, a, b);
      const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);   // sum what the tensor core will read (bf16-rounded)
      cs_sum[u] = (fa.x + fa.y) + (fb.x + fb.y);
      if (p.lo_off > 0) {                           // fp32 mode: second term, and the channel sum of the fp32 values
        store4(ob + (size_t)p.lo_off * p.N * p.C * p.HW + (size_t)c * p.HW, __floats2bfloat162_rn(x0 - fa.x, x1 - fa.y),
               __floats2bfloat162_rn(x2 - fb.x, x3 - fb.y));
        cs_sum[u] = (x0 + x1) + (x2 + x3);
      }
    }
  }
  if (ls >= 0) {                                    // key maps only: the backward's  b_s * sum_j k_j  term
    // butterfly over the 16 pixel-quad lanes: afterwards lane pq holds the 64-pixel sum of value pq
#pragma unroll
    for (int nn = 8, off = 8; off >= 1; nn >>= 1, off >>= 1) {
      const bool hi = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < nn; ++k) {
        const float send = hi ? cs_sum[k] : cs_sum[k + nn];
        const float recv = __shfl_xor_sync(0xffffffffu, send, off);
        cs_sum[k] = (hi ? cs_sum[k + nn] : cs_sum[k]) + recv;
      }
    }
    if (pq < NU) atomicAdd(p.ksum + ((size_t)slot * p.N + n) * p.C + warp * 2 + cs + 16 * pq, cs_sum[0]);
  }
}

template <int CPW>
__global__ void __launch_bounds__(256, 2) pix_prepare_kernel(const PrepArgs p) {
  __shared__ float s_ss[8][64];
  const int z = blockIdx.z, n = blockIdx.y;
  const size_t off = (size_t)n * p.C * p.HW;
  if (p.zero != nullptr) {                  // a fresh buffer nobody reads yet: cleared while the label pass still runs
    const long nb = (long)gridDim.x * gridDim.y * gridDim.z, b = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    for (long i = b * 256 + threadIdx.x; i < p.n_zero4; i += nb * 256) p.zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();                               // the label pass (sort permutation, cleared channel sums) is complete
  pdl_launch_dependents();
  switch (p.dtype[z]) {
    case 1: prepare_body<CPW, float>(p, static_cast<const float*>(p.x[z]) + off, z, n, s_ss); break;
    case 2: prepare_body<CPW, __half>(p, static_cast<const __half*>(p.x[z]) + off, z, n, s_ss); break;
    default: prepare_body<CPW, __nv_bfloat16>(p, static_cast<const __nv_bfloat16*>(p.x[z]) + off, z, n, s_ss); break;
  }
}

// ------------------------------------------------------------------------------------------------
// forward
constexpr int PF_STAGES = 6;
constexpr int PF_BSTAGE = 2 * 8192;      // this CTA's 128 keys x 64 channels of a 256-key tile
constexpr int PF_AKB = 2 * 8192;         // this CTA's 128 query pixels x 64 channels
constexpr int PX_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue / operand generation

struct PixFwdArgs {
  int N, C, HW, HWp, GLp, Q, S, nkb, num_mbp, num_tiles;
  const uint8_t* lab_nat;
  const uint8_t* lab_sorted;
  const uint8_t* glab;
  float* stats;      // [Q, N, HW, S, 2 (column half of a tile), 2 (same-label sum, total sum)]
};

// sum of 32 fp32 values held as 16 register pairs (packed adds)
__device__ __forceinline__ float sum32(const uint32_t (&v)[32]) {
  uint64_t a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    a[k] = f2_add(f2_pack(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])),
                  f2_pack(__uint_as_float(v[2 * k + 16]), __uint_as_float(v[2 * k + 17])));
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = f2_add(a[k], a[k + 4]);
  a[0] = f2_add(f2_add(a[0], a[2]), f2_add(a[1], a[3]));
  float lo, hi;
  f2_unpack(a[0], lo, hi);
  return lo + hi;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PX_THREADS, 1)
pixloss_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const PixTable tab, const PixFwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_a = smem;                                   // nkb k-blocks of this CTA's 128 query pixels (<= 4 x 16 KB)
  uint8_t* s_b = s_a + 4 * PF_AKB;                       // ring
  uint8_t* s_lab = s_b + PF_STAGES * PF_BSTAGE;          // 2 x (sorted labels [HWp], group labels [GLp])
  const int lab_bytes = p.HWp + p.GLp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lab + 2 * lab_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + PF_STAGES;
  uint64_t* a_full = bars + 2 * PF_STAGES;
  uint64_t* a_free = a_full + 1;
  uint64_t* acc_full = a_full + 2;    // [2]
  uint64_t* acc_empty = a_full + 4;   // [2]
  uint64_t* lab_full = a_full + 6;    // [2]
  uint64_t* lab_empty = a_full + 8;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_items = p.Q * p.N * p.num_mbp * p.S;
  // contiguous item ranges: consecutive items of a pair differ in the key set only, so the query tile stays
  const int it0 = (int)((long)cluster_id * num_items / num_clusters);
  const int it1 = (int)((long)(cluster_id + 1) * num_items / num_clusters);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < PF_STAGES; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_full, 2);
    mbar_init(a_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 16);         // leader's copy: the 8 epilogue warps of both CTAs
      mbar_init(&lab_full[i], 1);
      mbar_init(&lab_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                               // everything above overlapped the prepare kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, a_loads = 0, li = 0;
      uint32_t phase = 0;
      for (int item = it0; item < it1; ++item, ++li) {
        const int s = item % p.S;
        const int mbp = (item / p.S) % p.num_mbp;
        const int n = (item / (p.S * p.num_mbp)) % p.N;
        const int q = item / (p.S * p.num_mbp * p.N);
        const int kslot = tab.kmap[q][s], lslot = tab.klab[q][s];
        {   // the key set's sorted labels, for this CTA's epilogue warps
          const int b = li & 1;
          mbar_wait(&lab_empty[b], ((li >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&lab_full[b], lab_bytes);
          bulk_load_1d(s_lab + b * lab_bytes, p.lab_sorted + ((size_t)lslot * p.N + n) * p.HWp, p.HWp, &lab_full[b]);
          bulk_load_1d(s_lab + b * lab_bytes + p.HWp, p.glab + ((size_t)lslot * p.N + n) * p.GLp, p.GLp, &lab_full[b]);
        }
        if (item == it0 || s == 0) {
          if (a_loads > 0) mbar_wait(a_free, (a_loads - 1) & 1);
          const uint32_t la = mapa_u32(a_full, 0);
          mbar_arrive_expect_tx_cluster(la, p.nkb * PF_AKB);
          for (int kb = 0; kb < p.nkb; ++kb)
            for (int c2 = 0; c2 < 2; ++c2)
              tma_load_3d_2sm(s_a + kb * PF_AKB + c2 * 8192, &tm_x, la, mbp * 256 + int(crank) * 128 + c2 * 64, kb * 64,
                              tab.qmap[q] * p.N + n);
          ++a_loads;
        }
        for (int t = 0; t < p.num_tiles; ++t)
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t lf = mapa_u32(&full_bar[stage], 0);
            mbar_arrive_expect_tx_cluster(lf, PF_BSTAGE);
            for (int cc = 0; cc < 2; ++cc)
              tma_load_3d_2sm(s_b + stage * PF_BSTAGE + cc * 8192, &tm_x, lf, t * 256 + (int(crank) * 2 + cc) * 64, kb * 64,
                              kslot * p.N + n);
            if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {   // warp-uniform control flow, one elected lane issues (descriptors stay in uniform registers)
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 1, 1);
      int stage = 0, acc = 0, a_uses = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = it0; item < it1; ++item) {
        if (item == it0 || item % p.S == 0) {
          mbar_wait(a_full, a_uses & 1);
          ++a_uses;
          tc_fence_after();
        }
        for (int t = 0; t < p.num_tiles; ++t) {
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t aa = smem_u32(s_a + kb * PF_AKB), ba = smem_u32(s_b + stage * PF_BSTAGE);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_2sm(tmem_base + acc * 256, umma_smem_desc(aa + kk * 2048, 8192, 1024),
                              umma_smem_desc(ba + kk * 2048, 8192, 1024), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              umma_commit_2sm_mcast(&empty_bar[stage], 0x3);
            }
            __syncwarp();
            if (++stage == PF_STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader) umma_commit_2sm_mcast(&acc_full[acc], 0x3);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if ((item + 1 == it1 || (item + 1) % p.S == 0) && leader) umma_commit_2sm_mcast(a_free, 0x3);
        __syncwarp();
      }
    }
  } else {
    // epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).., the two warps of a lane quarter split the 256 key
    // columns of a tile in halves.  A thread owns one query pixel and 128 columns per tile.
    const int ew = warp - 2, wq = warp & 3, chalf = ew >> 2;
    int acc = 0, li = 0;
    uint32_t acc_phase = 0;
    for (int item = it0; item < it1; ++item, ++li) {
      const int s = item % p.S;
      const int mbp = (item / p.S) % p.num_mbp;
      const int n = (item / (p.S * p.num_mbp)) % p.N;
      const int q = item / (p.S * p.num_mbp * p.N);
      const int i = mbp * 256 + int(crank) * 128 + wq * 32 + lane;
      const bool row_ok = i < p.HW;
      const uint32_t lrow = row_ok ? p.lab_nat[((size_t)tab.qlab[q] * p.N + n) * p.HWp + i] : uint32_t(kPad);
      const int b = li & 1;
      mbar_wait(&lab_full[b], (li >> 1) & 1);
      const uint8_t* lk = s_lab + b * lab_bytes;
      const uint8_t* gl = lk + p.HWp;
      float pos = 0.f, tot = 0.f;
      auto group = [&](const uint32_t (&v)[32], int gcol) {
        const uint32_t g = gl[gcol];
        if (g != uint32_t(kMixed)) {                      // one label for the 32 keys (warp-uniform branch)
          const float sg = sum32(v);
          tot += sg;
          if (g == lrow) pos += sg;
        } else {
          const uint4 l0 = *reinterpret_cast<const uint4*>(lk + gcol * 32);
          const uint4 l1 = *reinterpret_cast<const uint4*>(lk + gcol * 32 + 16);
          const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          float p0 = 0.f, p1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            const uint32_t la = (lw[jj >> 2] >> ((jj & 3) * 8)) & 0xffu, lb = (lw[jj >> 2] >> (((jj + 1) & 3) * 8)) & 0xffu;
            const float za = __uint_as_float(v[jj]), zb = __uint_as_float(v[jj + 1]);
            t0 += za; t1 += zb;
            if (la == lrow) p0 += za;
            if (lb == lrow) p1 += zb;
          }
          tot += t0 + t1;
          pos += p0 + p1;
        }
      };
      for (int t = 0; t < p.num_tiles; ++t) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (uint32_t(wq * 32) << 16) + acc * 256 + chalf * 128;
        const int g0 = t * 8 + chalf * 4;
        uint32_t va[32], vb[32];
        tmem_ld32(t_addr, va);
        tmem_ld_wait();
        tmem_ld32(t_addr + 32, vb);
        group(va, g0);
        tmem_ld_wait();
        tmem_ld32(t_addr + 64, va);
        group(vb, g0 + 1);
        tmem_ld_wait();
        tmem_ld32(t_addr + 96, vb);
        group(va, g0 + 2);
        tmem_ld_wait();
        // this warp's last TMEM read of the tile is in registers: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
        group(vb, g0 + 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (row_ok)
        *reinterpret_cast<float2*>(p.stats + (((((size_t)q * p.N + n) * p.HW + i) * p.S + s) * 2 + chalf) * 2) =
            make_float2(pos, tot);
      __syncwarp();
      if (lane == 0) mbar_arrive(&lab_empty[b]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: per-row loss and backward coefficients, one thread per query pixel; the mean is a two-level sum in
// a fixed order (block partials, summed by the last block to finish), so the loss is bit-reproducible.
struct PixFinArgs {
  int N, HW, HWp, Q, S, n_terms;
  long term_stride;         // floats between the stats of consecutive terms
  const float* stats;       // [Q, N, HW, S, 2, 2]
  const uint8_t* lab_nat;
  const int* hist;          // [label slots, N, 256]
  const int* err;
  float* loss;              // device scalar: sum over the queries of their mean row loss
  float* loss_q;            // [Q] per query, or null
  float* coef;              // [Q, N, HW, 1 + S] or null
  float* partial;           // [Q * blocks_per_q]
  unsigned int* ticket;     // zero before the first call; left zero
};

constexpr int FIN_THREADS = 64;      // small CTAs: the kernel is latency-bound and also clears the backward's accumulator
__global__ void __launch_bounds__(FIN_THREADS) pixloss_finalize_kernel(const PixTable tab, const PixFinArgs p) {
  __shared__ float s_part[8];
  __shared__ bool s_last;
  const int q = blockIdx.y;
  const long rows = (long)p.N * p.HW;
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();                               // the similarity kernel's row sums are complete
  float li = 0.f;
  if (r < rows) {
    const int n = (int)(r / p.HW), i = (int)(r - (long)n * p.HW);
    const int lrow = p.lab_nat[((size_t)tab.qlab[q] * p.N + n) * p.HWp + i];
    const float* st = p.stats + ((size_t)q * rows + r) * p.S * 4;
    float psum = 0.f, pcnt = 0.f, nterm = 0.f;
    for (int s = 0; s < p.S; ++s) {
      float same = 0.f, total = 0.f;
      for (int t = 0; t < p.n_terms; ++t) {                                // fp32 mode: hi*hi + hi*lo + lo*hi
        const float4 v = *reinterpret_cast<const float4*>(st + t * p.term_stride + s * 4);   // (same, total) x 2 column halves
        same += v.x + v.z;
        total += v.y + v.w;
      }
      const float cnt = (float)p.hist[((size_t)tab.klab[q][s] * p.N + n) * 256 + lrow];
      psum += same;
      pcnt += cnt;
      nterm += (total - same) / ((float)p.HW - cnt + kEpsCnt);
    }
    const float P = psum / (pcnt + kEpsCnt);
    const float eP = expf(P), eN = expf(nterm);
    const float ratio = eP / (eP + eN);
    li = -logf(ratio + kEpsLog);
    if (p.coef != nullptr) {
      // d li / dP = -ratio (1 - ratio) / (ratio + eps) ; d li / dN = + the same
      const float gP = -ratio * (1.f - ratio) / (ratio + kEpsLog) / (float)rows;
      float* cf = p.coef + ((size_t)q * rows + r) * (p.S + 1);
      cf[0] = gP / (pcnt + kEpsCnt);
      for (int s = 0; s < p.S; ++s) {
        const float ncnt = (float)p.HW - (float)p.hist[((size_t)tab.klab[q][s] * p.N + n) * 256 + lrow];
        // a key set without any different-label pixel contributes the constant 0 / (0 + eps): its
        // gradient is exactly zero (and 1/eps here would amplify rounding noise a million-fold)
        cf[1 + s] = ncnt > 0.f ? -gP / (ncnt + kEpsCnt) : 0.f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = li;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < FIN_THREADS / 32; ++w) t += s_part[w];
    p.partial[q * gridDim.x + blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(p.ticket, 1u);
    s_last = (done == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    const bool bad = *p.err != 0;
    float total = 0.f;
    for (int qq = 0; qq < p.Q; ++qq) {
      float t = 0.f;                                      // lane-strided partial sums, then a fixed shuffle tree
      for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) t += __ldcg(p.partial + qq * gridDim.x + b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      t = bad ? __int_as_float(0x7fc00000) : t / (float)rows;     // out-of-range label: NaN (the reference raises)
      if (threadIdx.x == 0 && p.loss_q != nullptr) p.loss_q[qq] = t;
      total += t;
    }
    if (threadIdx.x == 0) {
      p.loss[0] = total;
      *p.ticket = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
constexpr int PB_GEN = 4;                    // generated same-label operand buffers (128 rows x 64 keys, 16 KB)
constexpr int PB_STAGES = 6;                 // key k-blocks: this CTA's C/2 channels x 64 keys

struct PixBwdArgs {
  int N, C, HW, HWp, GLp, Q, S, num_mbp, nkb;   // nkb = HWp / 64
  const uint8_t* lab_nat;
  const uint8_t* lab_sorted;
  const uint8_t* glab;
  const float* coef;     // [Q, N, HW, 1 + S]
  const float* ksum;     // [slots, N, C]
  const float* d_loss;   // device scalar (upstream gradient)
  int add_ksum;          // 0 for the second (lo) key term of the fp32 mode: the b_s * colsum(K_s) part is added once
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PX_THREADS, 1)
pixloss_bwd_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_dq,
                   const PixTable tab, const PixBwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int b_stage_bytes = (p.C / 2) * 128;
  uint8_t* s_gen = smem;
  uint8_t* s_b = s_gen + PB_GEN * 16384;
  uint8_t* s_stg = s_b + PB_STAGES * 16384;                // 4 drain warps x 2 x 4 KB
  uint8_t* s_lab = s_stg + 4 * 8192;
  const int lab_bytes = p.HWp + p.GLp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lab + 2 * lab_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + PB_STAGES;
  uint64_t* gen_full = bars + 2 * PB_STAGES;
  uint64_t* gen_empty = gen_full + PB_GEN;
  uint64_t* acc_full = gen_empty + PB_GEN;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* lab_full = acc_full + 4;         // [2]
  uint64_t* lab_empty = acc_full + 6;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_items = p.Q * p.N * p.num_mbp * p.S;
  const int it0 = (int)((long)cluster_id * num_items / num_clusters);
  const int it1 = (int)((long)(cluster_id + 1) * num_items / num_clusters);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_dq);
    for (int i = 0; i < PB_STAGES; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < PB_GEN; ++i) { mbar_init(&gen_full[i], 8); mbar_init(&gen_empty[i], 1); }   // 4 warps x 2 CTAs
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);           // leader's copy: the 4 drain warps of both CTAs
      mbar_init(&lab_full[i], 1);
      mbar_init(&lab_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, li = 0;
      uint32_t phase = 0;
      for (int item = it0; item < it1; ++item, ++li) {
        const int s = item % p.S;
        const int n = (item / (p.S * p.num_mbp)) % p.N;
        const int q = item / (p.S * p.num_mbp * p.N);
        const int kslot = tab.kmap[q][s], lslot = tab.klab[q][s];
        {
          const int b = li & 1;
          mbar_wait(&lab_empty[b], ((li >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&lab_full[b], lab_bytes);
          bulk_load_1d(s_lab + b * lab_bytes, p.lab_sorted + ((size_t)lslot * p.N + n) * p.HWp, p.HWp, &lab_full[b]);
          bulk_load_1d(s_lab + b * lab_bytes + p.HWp, p.glab + ((size_t)lslot * p.N + n) * p.GLp, p.GLp, &lab_full[b]);
        }
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t lf = mapa_u32(&full_bar[stage], 0);
          mbar_arrive_expect_tx_cluster(lf, b_stage_bytes);
          tma_load_3d_2sm(s_b + stage * 16384, &tm_k, lf, kb * 64, int(crank) * (p.C / 2), kslot * p.N + n);
          if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(256, p.C, 0, 0);
      int stage = 0, gb = 0, acc = 0;
      uint32_t phase = 0, gphase = 0, acc_phase = 0;
      for (int item = it0; item < it1; ++item) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&gen_full[gb], gphase);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t aa = smem_u32(s_gen + gb * 16384), ba = smem_u32(s_b + stage * 16384);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_2sm(tmem_base + acc * 256, umma_smem_desc(aa + kk * 32, 16, 1024),
                            umma_smem_desc(ba + kk * 32, 16, 1024), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit_2sm_mcast(&gen_empty[gb], 0x3);
            umma_commit_2sm_mcast(&empty_bar[stage], 0x3);
          }
          __syncwarp();
          if (++gb == PB_GEN) { gb = 0; gphase ^= 1; }
          if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit_2sm_mcast(&acc_full[acc], 0x3);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ---- operand generation: thread = one query pixel of this CTA's 128; per k-block it writes its row of
    // the 0/1 same-label matrix (64 keys = two 32-key groups = 8 x 16 bytes, 128B-swizzled K-major)
    const int row = (warp - 2) * 32 + lane;
    int gb = 0, li = 0;
    uint32_t gphase = 0;
    for (int item = it0; item < it1; ++item, ++li) {
      const int mbp = (item / p.S) % p.num_mbp;
      const int n = (item / (p.S * p.num_mbp)) % p.N;
      const int q = item / (p.S * p.num_mbp * p.N);
      const int i = mbp * 256 + int(crank) * 128 + row;
      const bool row_ok = i < p.HW;
      const uint32_t lrow = row_ok ? p.lab_nat[((size_t)tab.qlab[q] * p.N + n) * p.HWp + i] : uint32_t(kPad);
      const int b = li & 1;
      mbar_wait(&lab_full[b], (li >> 1) & 1);
      const uint8_t* lk = s_lab + b * lab_bytes;
      const uint8_t* gl = lk + p.HWp;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&gen_empty[gb], gphase ^ 1);
        uint8_t* dst = s_gen + gb * 16384;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t g = gl[kb * 2 + h];
          if (g != uint32_t(kMixed)) {
            const uint32_t w = (row_ok && g == lrow) ? 0x3F803F80u : 0u;
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + sw128_offset(row, h * 4 + c)) = make_uint4(w, w, w, w);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint2 l8 = *reinterpret_cast<const uint2*>(lk + kb * 64 + h * 32 + c * 8);
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t src = (e < 2) ? l8.x : l8.y;
                const uint32_t la = (src >> (((2 * e) & 3) * 8)) & 0xffu, lb = (src >> (((2 * e + 1) & 3) * 8)) & 0xffu;
                w[e] = ((row_ok && la == lrow) ? 0x3F80u : 0u) | ((row_ok && lb == lrow) ? 0x3F800000u : 0u);
              }
              *reinterpret_cast<uint4*>(dst + sw128_offset(row, h * 4 + c)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&gen_full[gb], 0));
        if (++gb == PB_GEN) { gb = 0; gphase ^= 1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&lab_empty[b]);
    }
  } else {
    // ---- drain: dq[i, :] += (a - b_s) * D[i, :] + b_s * ksum_s[:]   (fp32 TMA add-reduction)
    const int wq = warp & 3;
    uint8_t* const my_stg = s_stg + (warp - 6) * 8192;
    const float g_up = __ldg(p.d_loss);
    int acc = 0, chunk_no = 0;
    uint32_t acc_phase = 0;
    for (int item = it0; item < it1; ++item) {
      const int s = item % p.S;
      const int mbp = (item / p.S) % p.num_mbp;
      const int n = (item / (p.S * p.num_mbp)) % p.N;
      const int q = item / (p.S * p.num_mbp * p.N);
      const int row0 = mbp * 256 + int(crank) * 128 + wq * 32;
      const int i = row0 + lane;
      float ca = 0.f, cb_ = 0.f;
      if (i < p.HW) {
        const float* cf = p.coef + (((size_t)q * p.N + n) * p.HW + i) * (p.S + 1);
        ca = cf[0] * g_up;
        cb_ = cf[1 + s] * g_up;
      }
      const float* ks = p.ksum + ((size_t)tab.kmap[q][s] * p.N + n) * p.C;
      const float cm = ca - cb_;
      if (!p.add_ksum) cb_ = 0.f;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int nc32 = p.C / 32;
#pragma unroll 1
      for (int c32 = 0; c32 < nc32; ++c32, ++chunk_no) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + acc * 256 + c32 * 32, v);
        const float ksl = __ldg(ks + c32 * 32 + lane);
        uint8_t* buf = my_stg + (chunk_no & 1) * 4096;
        if (lane == 0) tma_wait_group_read<1>();      // this tile's previous add-reduction has been read out
        tmem_ld_wait();
        if (c32 == nc32 - 1) {                        // last TMEM read of the item: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = fmaf(cm, __uint_as_float(v[4 * j + e]), cb_ * __shfl_sync(0xffffffffu, ksl, 4 * j + e));
          *reinterpret_cast<float4*>(buf + sw128_offset(lane, j)) = make_float4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (row0 < p.HW) tma_reduce_add_3d(&tm_dq, buf, c32 * 32, row0, q * p.N + n);
          tma_commit_group();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// dq32 [Q, N, HW, C] fp32 (pixel-major, what the add-reduction produced) -> dq [N, C, HW] in the query's dtype,
// through the Jacobian of x / max(|x|, eps) when the normalisation was fused:  dx = inv * (g - xn (xn . g))
struct FinishArgs {
  void* out[PX_MAX_Q];
  int16_t qmap[PX_MAX_Q];
  int16_t qmap_lo[PX_MAX_Q];         // second bf16 term of the query map (fp32 mode), or -1
  int N, C, HW, out_dtype, chain;      // out_dtype: 0 bf16, 1 f32, 2 f16
  const float* dq32;
  const __nv_bfloat16* xn;
  const float* inv_norm;
};

template <int CPW>
__global__ void __launch_bounds__(256) pix_dq_finish_kernel(const FinishArgs p) {
  extern __shared__ float s_fin[];           // [32][C + 1] gradient tile, then [8][32] partial dots
  const int q = blockIdx.z, n = blockIdx.y, i0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = p.C + 1;
  float* s_dot = s_fin + 32 * ld;
  const int i = i0 + lane;
  const bool ok = i < p.HW;
  // the normalised query values of this thread's pixel (written by the forward's prepare kernel, long complete): requested
  // before the wait, so their latency overlaps the backward kernel's tail and the tile load below
  float xq[CPW];
  float inv = 1.f;
  if (p.chain) {
    const __nv_bfloat16* xb = p.xn + ((size_t)p.qmap[q] * p.N + n) * p.C * p.HW;
    __nv_bfloat16 raw[CPW], raw_lo[CPW];
#pragma unroll
    for (int u = 0; u < CPW; ++u) raw[u] = ok ? xb[(size_t)(warp + 8 * u) * p.HW + i] : __float2bfloat16_rn(0.f);
    if (p.qmap_lo[q] >= 0) {
      const __nv_bfloat16* xl = p.xn + ((size_t)p.qmap_lo[q] * p.N + n) * p.C * p.HW;
#pragma unroll
      for (int u = 0; u < CPW; ++u) raw_lo[u] = ok ? xl[(size_t)(warp + 8 * u) * p.HW + i] : __float2bfloat16_rn(0.f);
    }
    if (ok) inv = p.inv_norm[((size_t)p.qmap[q] * p.N + n) * p.HW + i];
#pragma unroll
    for (int u = 0; u < CPW; ++u) xq[u] = __bfloat162float(raw[u]) + (p.qmap_lo[q] >= 0 ? __bfloat162float(raw_lo[u]) : 0.f);
  }
  pdl_wait();                                // the add-reductions of the backward kernel are complete
  const float* src = p.dq32 + (((size_t)q * p.N + n) * p.HW + i0) * p.C;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    for (int c = lane; c < p.C; c += 32) s_fin[r * ld + c] = (i0 + r < p.HW) ? __ldcg(src + (size_t)r * p.C + c) : 0.f;
  }
  __syncthreads();
  float dot = 0.f;
  if (p.chain) {
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < CPW; ++u) part = fmaf(xq[u], s_fin[lane * ld + warp + 8 * u], part);
    s_dot[warp * 32 + lane] = part;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) dot += s_dot[w * 32 + lane];
  }
  if (!ok) return;
#pragma unroll
  for (int u = 0; u < CPW; ++u) {
    const int c = warp + 8 * u;
    float g = s_fin[lane * ld + c];
    if (p.chain) g = inv * (g - xq[u] * dot);
    const size_t o = ((size_t)n * p.C + c) * p.HW + i;
    if (p.out_dtype == 1) static_cast<float*>(p.out[q])[o] = g;
    else if (p.out_dtype == 0) static_cast<__nv_bfloat16*>(p.out[q])[o] = __float2bfloat16_rn(g);
    else static_cast<__half*>(p.out[q])[o] = __float2half_rn(g);
  }
}

int check_pix_shape(int Q, int S, int N, int C, int HW) {
  STSWIN_CHECK_ARG(Q >= 1 && Q <= PX_MAX_Q, "pixloss: %d queries out of range [1,%d]", Q, PX_MAX_Q);
  STSWIN_CHECK_ARG(S >= 1 && S <= PX_MAX_SETS, "pixloss: n_sets=%d out of range [1,%d]", S, PX_MAX_SETS);
  STSWIN_CHECK_ARG(N > 0 && HW > 0, "pixloss: empty input");
  if (C % 64 != 0 || C > 256) return set_error(kErrUnsupported, "pixloss: C=%d unsupported (multiple of 64, <= 256)", C);
  if (HW % 8 != 0) return set_error(kErrUnsupported, "pixloss: H*W=%d must be a multiple of 8", HW);
  if (HW > 8192) return set_error(kErrUnsupported, "pixloss: H*W=%d > 8192 unsupported", HW);
  return kOk;
}

int fill_table(PixTable* t, const int* qmap, const int* qlab, const int* kmap, const int* klab, int Q, int S,
               int n_slots, int n_lslots) {
  for (int q = 0; q < Q; ++q) {
    STSWIN_CHECK_ARG(qmap[q] >= 0 && qmap[q] < n_slots && qlab[q] >= 0 && qlab[q] < n_lslots, "pixloss: query %d slot out of range", q);
    t->qmap[q] = (int16_t)qmap[q];
    t->qlab[q] = (int16_t)qlab[q];
    for (int s = 0; s < S; ++s) {
      const int km = kmap[q * S + s], kl = klab[q * S + s];
      STSWIN_CHECK_ARG(km >= 0 && km < n_slots && kl >= 0 && kl < n_lslots, "pixloss: key set (%d,%d) slot out of range", q, s);
      t->kmap[q][s] = (int16_t)km;
      t->klab[q][s] = (int16_t)kl;
    }
  }
  return kOk;
}

int map_tmap(CUtensorMap* tm, const void* base, int rows3, int C, int HW, uint32_t box_c) {
  uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)rows3};
  uint64_t str[2] = {(uint64_t)HW * 2, (uint64_t)C * HW * 2};
  uint32_t box[3] = {64, box_c, 1};
  return make_tmap(tm, TmapDtype::BF16, 3, base, dims, str, box, true);
}

}  // namespace

// see include/stswin_b200.h : stswin_pixloss_labels
int pixloss_labels(const void* const* labels, const int* dtypes, int n_labels, int slot_off, int N, int Hs, int Ws, int H,
                   int W, int class_num, uint8_t* lab_nat, uint8_t* lab_sorted, uint8_t* glab, uint16_t* perm, int* hist,
                   int* err_flag, float* ksum_to_clear, long ksum_elems, cudaStream_t stream) {
  STSWIN_CHECK_ARG(labels && dtypes && lab_nat && lab_sorted && glab && perm && hist && err_flag, "pixloss_labels: null pointer");
  STSWIN_CHECK_ARG(n_labels >= 1 && N > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "pixloss_labels: bad shape");
  STSWIN_CHECK_ARG(class_num >= 1 && class_num <= 254, "pixloss_labels: class_num=%d out of range [1,254]", class_num);
  const int HW = H * W;
  STSWIN_CHECK_ARG(HW <= 8192 && HW % 8 == 0, "pixloss_labels: H*W=%d unsupported (multiple of 8, <= 8192)", HW);
  if (slot_off == 0) STSWIN_CUDA(cudaMemsetAsync(err_flag, 0, 2 * sizeof(int), stream));   // bad-label flag + ticket
  for (int off = 0; off < n_labels; off += PX_MAX_PTRS) {
    const int nl = n_labels - off < PX_MAX_PTRS ? n_labels - off : PX_MAX_PTRS;
    LabelArgs a;
    for (int i = 0; i < PX_MAX_PTRS; ++i) {
      const int src = off + (i < nl ? i : 0);
      STSWIN_CHECK_ARG(labels[src] != nullptr && dtypes[src] >= 0 && dtypes[src] <= 5, "pixloss_labels: bad label map %d", src);
      a.src[i] = labels[src];
      a.dtype[i] = (uint8_t)dtypes[src];
    }
    a.n_labels = nl; a.slot_off = slot_off + off; a.N = N; a.Hs = Hs; a.Ws = Ws; a.H = H; a.W = W; a.class_num = class_num;
    a.lab_nat = lab_nat; a.lab_sorted = lab_sorted; a.glab = glab; a.perm = perm; a.hist = hist; a.err = err_flag;
    a.zero = (off == 0 && slot_off == 0) ? ksum_to_clear : nullptr; a.n_zero = ksum_elems;
    pix_labels_kernel<<<dim3(N, nl), 32 * LB_WARPS, 2 * px_hwp(HW), stream>>>(a);
    STSWIN_CUDA(cudaGetLastError());
  }
  return kOk;
}

// see include/stswin_b200.h : stswin_pixloss_prepare
int pixloss_prepare(const void* const* maps, const int* dtypes, const int* label_slots, int n_maps, int slot_off, int N,
                    int C, int HW, int do_normalize, int lo_slot_off, const uint16_t* perm, void* xn, float* inv_norm,
                    float* ksum, float* f32_to_clear, long f32_elems, cudaStream_t stream) {
  STSWIN_CHECK_ARG(f32_to_clear == nullptr || (f32_elems % 4 == 0 && (reinterpret_cast<uintptr_t>(f32_to_clear) & 15) == 0),
                   "pixloss_prepare: the buffer to clear must be 16-byte aligned with a multiple of 4 elements");
  STSWIN_CHECK_ARG(lo_slot_off == 0 || lo_slot_off >= n_maps, "pixloss_prepare: lo_slot_off must be 0 or >= n_maps");
  STSWIN_CHECK_ARG(maps && dtypes && label_slots && xn && ksum && perm, "pixloss_prepare: null pointer");
  STSWIN_CHECK_ARG(n_maps >= 1 && N > 0, "pixloss_prepare: bad shape");
  STSWIN_CHECK_ARG(!do_normalize || inv_norm != nullptr, "pixloss_prepare: normalisation needs inv_norm");
  if (C % 64 != 0 || C > 256) return set_error(kErrUnsupported, "pixloss: C=%d unsupported (multiple of 64, <= 256)", C);
  if (HW % 8 != 0 || HW > 8192) return set_error(kErrUnsupported, "pixloss: H*W=%d unsupported (multiple of 8, <= 8192)", HW);
  for (int off = 0; off < n_maps; off += PX_MAX_PTRS) {
    const int nm = n_maps - off < PX_MAX_PTRS ? n_maps - off : PX_MAX_PTRS;
    PrepArgs a;
    for (int i = 0; i < PX_MAX_PTRS; ++i) {
      const int src = off + (i < nm ? i : 0);
      STSWIN_CHECK_ARG(maps[src] != nullptr && dtypes[src] >= 0 && dtypes[src] <= 2, "pixloss_prepare: bad map %d", src);
      STSWIN_CHECK_ARG(label_slots[src] >= -1 && label_slots[src] < 127, "pixloss_prepare: bad label slot of map %d", src);
      STSWIN_CHECK_ARG((reinterpret_cast<uintptr_t>(maps[src]) & 7) == 0, "pixloss_prepare: map %d is not 8-byte aligned", src);
      a.x[i] = maps[src];
      a.dtype[i] = (uint8_t)dtypes[src];
      a.lslot[i] = (int8_t)label_slots[src];
    }
    a.n_maps = nm; a.slot_off = slot_off + off; a.N = N; a.C = C; a.HW = HW; a.do_normalize = do_normalize;
    a.lo_off = lo_slot_off;
    a.zero = off == 0 ? reinterpret_cast<float4*>(f32_to_clear) : nullptr; a.n_zero4 = f32_elems / 4;
    a.perm = perm; a.xn = static_cast<__nv_bfloat16*>(xn); a.inv_norm = inv_norm; a.ksum = ksum;
    const dim3 grid((HW + 63) / 64, N, nm);
    switch (C / 64) {
      case 1: STSWIN_CUDA(launch_pdl(pix_prepare_kernel<8>, grid, dim3(256), 0, stream, a)); break;
      case 2: STSWIN_CUDA(launch_pdl(pix_prepare_kernel<16>, grid, dim3(256), 0, stream, a)); break;
      case 3: STSWIN_CUDA(launch_pdl(pix_prepare_kernel<24>, grid, dim3(256), 0, stream, a)); break;
      default: STSWIN_CUDA(launch_pdl(pix_prepare_kernel<32>, grid, dim3(256), 0, stream, a)); break;
    }
    STSWIN_CUDA(cudaGetLastError());
  }
  return kOk;
}

// see include/stswin_b200.h : stswin_pixloss_fwd
int pixloss_fwd(const void* xn, int n_slots, int n_label_slots, const uint8_t* lab_nat, const uint8_t* lab_sorted,
                const uint8_t* glab, const int* hist, const int* qmap, const int* qlab, const int* kmap, const int* klab,
                int n_terms, int Q, int S, int N, int C, int HW, float* stats, float* loss, float* loss_per_query, float* coef,
                const int* err_flag, float* partial, unsigned int* ticket, cudaStream_t stream) {
  STSWIN_CHECK_ARG(xn && lab_nat && lab_sorted && glab && hist && qmap && qlab && kmap && klab && stats && loss &&
                       err_flag && partial && ticket, "pixloss_fwd: null pointer");
  STSWIN_CHECK_ARG(n_terms >= 1 && n_terms <= 3, "pixloss_fwd: n_terms=%d out of range [1,3]", n_terms);
  int rc = check_pix_shape(Q, S, N, C, HW);
  if (rc != kOk) return rc;
  CUtensorMap tm;
  if ((rc = map_tmap(&tm, xn, n_slots * N, C, HW, 64)) != kOk) return rc;
  PixFwdArgs a;
  a.N = N; a.C = C; a.HW = HW; a.HWp = px_hwp(HW); a.GLp = px_glp(HW); a.Q = Q; a.S = S; a.nkb = C / 64;
  a.num_mbp = a.HWp / 256; a.num_tiles = a.HWp / 256;
  a.lab_nat = lab_nat; a.lab_sorted = lab_sorted; a.glab = glab;
  const int smem = 1024 + 4 * PF_AKB + PF_STAGES * PF_BSTAGE + 2 * (a.HWp + a.GLp) + 256;
  STSWIN_CUDA(cudaFuncSetAttribute(pixloss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int items = Q * N * a.num_mbp * S;
  const int max_clusters = num_sms() / 2;
  const int grid = 2 * (items < max_clusters ? items : max_clusters);
  const long term_stride = (long)Q * N * HW * S * 4;
  PixTable tab;
  for (int t = 0; t < n_terms; ++t) {      // fp32 mode: one launch per product term (the row sums are linear in the similarities)
    if ((rc = fill_table(&tab, qmap + t * Q, qlab, kmap + t * Q * S, klab, Q, S, n_slots, n_label_slots)) != kOk) return rc;
    a.stats = stats + t * term_stride;
    STSWIN_CUDA(launch_pdl(pixloss_fwd_kernel, dim3(grid), dim3(PX_THREADS), smem, stream, tm, tab, a));
  }
  PixFinArgs f;
  f.N = N; f.HW = HW; f.HWp = a.HWp; f.Q = Q; f.S = S; f.n_terms = n_terms; f.term_stride = term_stride;
  f.stats = stats; f.lab_nat = lab_nat; f.hist = hist; f.err = err_flag;
  f.loss = loss; f.loss_q = loss_per_query; f.coef = coef; f.partial = partial; f.ticket = ticket;
  const long rows = (long)N * HW;
  STSWIN_CUDA(launch_pdl(pixloss_finalize_kernel, dim3((unsigned)((rows + FIN_THREADS - 1) / FIN_THREADS), Q), dim3(FIN_THREADS), 0,
                         stream, tab, f));
  return kOk;
}

// see include/stswin_b200.h : stswin_pixloss_bwd
int pixloss_bwd(const void* xn, int n_slots, int n_label_slots, const uint8_t* lab_nat, const uint8_t* lab_sorted,
                const uint8_t* glab, const int* qmap, const int* qmap_lo, const int* qlab, const int* kmap, const int* klab,
                int n_terms, int Q, int S, int N, int C, int HW, const float* coef, const float* ksum, const float* d_loss,
                float* dq32, int dq32_is_clear, const float* inv_norm, void* const* dq_out, int out_dtype, cudaStream_t stream) {
  STSWIN_CHECK_ARG(xn && lab_nat && lab_sorted && glab && qmap && qlab && kmap && klab && coef && ksum && d_loss && dq32 &&
                       dq_out, "pixloss_bwd: null pointer");
  STSWIN_CHECK_ARG(out_dtype >= 0 && out_dtype <= 2, "pixloss_bwd: bad output dtype %d", out_dtype);
  STSWIN_CHECK_ARG(n_terms >= 1 && n_terms <= 2, "pixloss_bwd: n_terms=%d out of range [1,2]", n_terms);
  int rc = check_pix_shape(Q, S, N, C, HW);
  if (rc != kOk) return rc;
  PixTable tab;
  CUtensorMap tk, tdq;
  if ((rc = map_tmap(&tk, xn, n_slots * N, C, HW, (uint32_t)(C / 2))) != kOk) return rc;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)Q * N};
    uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)HW * C * 4};
    uint32_t box[3] = {32, 32, 1};
    if ((rc = make_tmap(&tdq, TmapDtype::F32, 3, dq32, dims, str, box, true)) != kOk) return rc;
  }
  if (!dq32_is_clear) STSWIN_CUDA(cudaMemsetAsync(dq32, 0, sizeof(float) * (size_t)Q * N * HW * C, stream));
  PixBwdArgs a;
  a.N = N; a.C = C; a.HW = HW; a.HWp = px_hwp(HW); a.GLp = px_glp(HW); a.Q = Q; a.S = S;
  a.num_mbp = a.HWp / 256; a.nkb = a.HWp / 64;
  a.lab_nat = lab_nat; a.lab_sorted = lab_sorted; a.glab = glab; a.coef = coef; a.ksum = ksum; a.d_loss = d_loss;
  const int smem = 1024 + PB_GEN * 16384 + PB_STAGES * 16384 + 4 * 8192 + 2 * (a.HWp + a.GLp) + 256;
  STSWIN_CUDA(cudaFuncSetAttribute(pixloss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int items = Q * N * a.num_mbp * S;
  const int max_clusters = num_sms() / 2;
  const int grid = 2 * (items < max_clusters ? items : max_clusters);
  for (int t = n_terms - 1; t >= 0; --t) {   // fp32 mode: keys = hi + lo, two launches reducing into the same dq (t = 0 last: its table feeds the finish kernel)
    if ((rc = fill_table(&tab, qmap, qlab, kmap + t * Q * S, klab, Q, S, n_slots, n_label_slots)) != kOk) return rc;
    a.add_ksum = (t == 0);
    pixloss_bwd_kernel<<<grid, PX_THREADS, smem, stream>>>(tk, tdq, tab, a);
    STSWIN_CUDA(cudaGetLastError());
  }
  FinishArgs f;
  for (int q = 0; q < PX_MAX_Q; ++q) {
    f.out[q] = dq_out[q < Q ? q : 0];
    f.qmap[q] = tab.qmap[q < Q ? q : 0];
    f.qmap_lo[q] = (int16_t)(qmap_lo != nullptr ? qmap_lo[q < Q ? q : 0] : -1);
    STSWIN_CHECK_ARG(f.out[q] != nullptr, "pixloss_bwd: null gradient output %d", q);
  }
  f.N = N; f.C = C; f.HW = HW; f.out_dtype = out_dtype; f.chain = inv_norm != nullptr; f.dq32 = dq32;
  f.xn = static_cast<const __nv_bfloat16*>(xn); f.inv_norm = inv_norm;
  const dim3 fgrid((HW + 31) / 32, N, Q);
  const int fsmem = (32 * (C + 1) + 8 * 32) * (int)sizeof(float);
  switch (C / 64) {
    case 1: STSWIN_CUDA(launch_pdl(pix_dq_finish_kernel<8>, fgrid, dim3(256), (size_t)fsmem, stream, f)); break;
    case 2: STSWIN_CUDA(launch_pdl(pix_dq_finish_kernel<16>, fgrid, dim3(256), (size_t)fsmem, stream, f)); break;
    case 3: STSWIN_CUDA(launch_pdl(pix_dq_finish_kernel<24>, fgrid, dim3(256), (size_t)fsmem, stream, f)); break;
    default: STSWIN_CUDA(launch_pdl(pix_dq_finish_kernel<32>, fgrid, dim3(256), (size_t)fsmem, stream, f)); break;
  }
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
