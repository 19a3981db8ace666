// Memory-bound glue of the Swin block: LayerNorm forward/backward (with the PatchMerging 2x2
// gather fused in), and the NCHW <-> token-major layout change of the layer container.
//
// Reference sites (seg18/net/Ours/swin_512.py):
//   norm2 / norm1 of the post-norm block      :235   (nn.LayerNorm, eps 1e-5, fp32 statistics)
//   PatchMerging: 2x2 strided gather + cat + LayerNorm(4C)     :266-274
//   permute(0,1,3,4,2).contiguous() / permute(0,1,3,2)         :314, :319, :326
//
// One warp per row, the row held in registers (C <= 2048), 128-bit loads/stores; each kernel
// reads its input once and writes its output once.  Column reductions of the backward (dgamma,
// dbeta and the column sums of dx that are the bias gradient of the preceding Linear) are
// accumulated in registers across the rows a warp owns, reduced through shared memory, and
// flushed with one atomic per column per CTA.
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace stswin {

namespace {

constexpr int LN_THREADS = 256;     // 8 warps = 8 rows in flight per CTA
constexpr int LN_WARPS = LN_THREADS / 32;

struct PmGeom {        // PatchMerging gather geometry (pm == 0: plain rows)
  int pm;
  int H, W, C;         // source token grid and channel count; output row has 4*C channels
};

// pointer to the 8-element group `v` (0 .. Ctot/8) of logical row `row`
template <typename T>
__device__ __forceinline__ T* row_ptr(T* base, long row, int v, int Ctot, const PmGeom& g) {
  if (!g.pm) return base + row * Ctot + v * 8;
  const int H2 = g.H >> 1, W2 = g.W >> 1;
  const int per_img = H2 * W2;
  const long bt = row / per_img;
  const int rem = int(row - bt * per_img);
  const int h2 = rem / W2, w2 = rem - h2 * W2;
  const int ch = v * 8;
  const int seg = ch / g.C;                 // 0:(dh0,dw0) 1:(dh1,dw0) 2:(dh0,dw1) 3:(dh1,dw1)   (:266-270)
  const int dh = seg & 1, dw = seg >> 1;
  const long tok = (bt * g.H + (2 * h2 + dh)) * g.W + (2 * w2 + dw);
  return base + tok * g.C + (ch - seg * g.C);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>   // NV = ceil(Ctot / 256): 16-byte groups per lane
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, long M, int Ctot,
              float eps, PmGeom pg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = Ctot >> 3;
  for (long row = (long)blockIdx.x * LN_WARPS + warp; row < M; row += (long)gridDim.x * LN_WARPS) {
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row_ptr(x, row, vi, Ctot, pg)));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          v[i][2 * k] = f.x; v[i][2 * k + 1] = f.y;
          s += f.x + f.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      }
    }
    const float mu = warp_sum(s) / Ctot;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mu; ss += d * d; }
      }
    const float rs = rsqrtf(warp_sum(ss) / Ctot + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        uint4 q;
        q.x = pack_bf16((v[i][0] - mu) * rs * g0.x + b0.x, (v[i][1] - mu) * rs * g0.y + b0.y);
        q.y = pack_bf16((v[i][2] - mu) * rs * g0.z + b0.z, (v[i][3] - mu) * rs * g0.w + b0.w);
        q.z = pack_bf16((v[i][4] - mu) * rs * g1.x + b1.x, (v[i][5] - mu) * rs * g1.y + b1.y);
        q.w = pack_bf16((v[i][6] - mu) * rs * g1.z + b1.z, (v[i][7] - mu) * rs * g1.w + b1.w);
        *reinterpret_cast<uint4*>(y + row * Ctot + vi * 8) = q;      // output rows are always dense
      }
    }
  }
}

template <int NV, bool COLSUM>
__global__ void __launch_bounds__(LN_THREADS, (NV <= 2 ? 2 : 1))
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dx_colsum, long M, int Ctot, PmGeom pg) {
  extern __shared__ float s_red[];       // [LN_WARPS][Ctot] scratch for the column reductions
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = Ctot >> 3;
  float a_dg[NV][8], a_db[NV][8], a_cs[COLSUM ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a_dg[i][k] = 0.f; a_db[i][k] = 0.f;
      if (COLSUM) a_cs[i][k] = 0.f;
    }
  const bool has_res = dres != nullptr;
  const long stride = (long)gridDim.x * LN_WARPS;
  // raw 16-byte groups of the row being processed, and of the next row (prefetched while this one
  // is reduced: one row per warp is otherwise a single dependent load -> shuffle -> store chain)
  uint4 cd[NV], cx[NV], cr[NV];
  float c_mu = 0.f, c_rs = 0.f;
  auto load_row = [&](long row, uint4 (&d)[NV], uint4 (&xx)[NV], uint4 (&r)[NV], float& mu, float& rs) {
    mu = mean[row]; rs = rstd[row];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        d[i] = __ldg(reinterpret_cast<const uint4*>(dy + row * Ctot + vi * 8));
        xx[i] = __ldg(reinterpret_cast<const uint4*>(row_ptr(x, row, vi, Ctot, pg)));
        if (has_res) r[i] = __ldg(reinterpret_cast<const uint4*>(dres + row * Ctot + vi * 8));
      }
    }
  };
  constexpr bool PREFETCH = NV <= 4;      // NV = 8 (PatchMerging rows of 2048) would spill with two rows live
  long row = (long)blockIdx.x * LN_WARPS + warp;
  if (PREFETCH && row < M) load_row(row, cd, cx, cr, c_mu, c_rs);
  while (row < M) {
    const long nxt = row + stride;
    uint4 nd[PREFETCH ? NV : 1], nx[PREFETCH ? NV : 1], nr[PREFETCH ? NV : 1];
    float n_mu = 0.f, n_rs = 0.f;
    if constexpr (PREFETCH) {
      if (nxt < M) load_row(nxt, nd, nx, nr, n_mu, n_rs);
    } else {
      load_row(row, cd, cx, cr, c_mu, c_rs);
    }
    const float mu = c_mu, rs = c_rs;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const uint32_t wd[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w}, wx[4] = {cx[i].x, cx[i].y, cx[i].z, cx[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fd = unpack_bf16(wd[k]), fx = unpack_bf16(wx[k]);
          const float h0 = (fx.x - mu) * rs, h1 = (fx.y - mu) * rs;
          a_dg[i][2 * k] += fd.x * h0; a_dg[i][2 * k + 1] += fd.y * h1;
          a_db[i][2 * k] += fd.x;      a_db[i][2 * k + 1] += fd.y;
          const float q0 = fd.x * gm[2 * k], q1 = fd.y * gm[2 * k + 1];
          s1 += q0 + q1;
          s2 += q0 * h0 + q1 * h1;
        }
      }
    }
    const float m1 = warp_sum(s1) / Ctot, m2 = warp_sum(s2) / Ctot;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const uint32_t wd[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w}, wx[4] = {cx[i].x, cx[i].y, cx[i].z, cx[i].w};
        const uint32_t wr[4] = {cr[i].x, cr[i].y, cr[i].z, cr[i].w};
        uint32_t wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fd = unpack_bf16(wd[k]), fx = unpack_bf16(wx[k]);
          float o0 = rs * (fd.x * gm[2 * k] - m1 - (fx.x - mu) * rs * m2);
          float o1 = rs * (fd.y * gm[2 * k + 1] - m1 - (fx.y - mu) * rs * m2);
          if (has_res) {
            const float2 fr = unpack_bf16(wr[k]);
            o0 += fr.x; o1 += fr.y;
          }
          wo[k] = pack_bf16(o0, o1);
          if (COLSUM) {                      // sum what the consumer will read (bf16-rounded)
            const float2 fo = unpack_bf16(wo[k]);
            a_cs[i][2 * k] += fo.x; a_cs[i][2 * k + 1] += fo.y;
          }
        }
        *reinterpret_cast<uint4*>(row_ptr(dx, row, vi, Ctot, pg)) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
      }
    }
    if constexpr (PREFETCH) {
#pragma unroll
      for (int i = 0; i < NV; ++i) { cd[i] = nd[i]; cx[i] = nx[i]; cr[i] = nr[i]; }
      c_mu = n_mu; c_rs = n_rs;
    }
    row = nxt;
  }
  // CTA-level column reduction, one quantity at a time through s_red[LN_WARPS][Ctot]
  auto flush = [&](float (&acc)[NV][8], float* out) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[warp * Ctot + vi * 8 + k] = acc[i][k];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Ctot; c += LN_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) t += s_red[w * Ctot + c];
      atomicAdd(out + c, t);
    }
  };
  flush(a_dg, dgamma);
  flush(a_db, dbeta);
  if constexpr (COLSUM) flush(a_cs, dx_colsum);
}

// [batch, R, Cc] -> [batch, Cc, R] with dtype conversion (fp32 or bf16 on either side)
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* __restrict__ in, TO* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const long b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const TI* ib = in + b * (long)R * Cc;
  TO* ob = out + b * (long)R * Cc;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = static_cast<float>(ib[(long)r * Cc + c]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[(long)c * R + r] = static_cast<TO>(tile[threadIdx.x][j]);
  }
}

int ln_grid(long M) {
  long want = (M + LN_WARPS - 1) / LN_WARPS;
  long cap = (long)num_sms() * 8;
  return (int)(want < cap ? want : cap);
}

int check_ln_shape(long M, int Ctot, int pm, int H, int W, int C) {
  STSWIN_CHECK_ARG(M > 0 && Ctot > 0, "layernorm: empty input");
  STSWIN_CHECK_ARG(Ctot % 8 == 0 && Ctot <= 2048, "layernorm: row length %d must be a multiple of 8 and <= 2048", Ctot);
  if (pm) {
    STSWIN_CHECK_ARG(Ctot == 4 * C && C % 8 == 0, "patch-merging layernorm: row length %d != 4*C (C=%d)", Ctot, C);
    STSWIN_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "patch-merging: x size (%d*%d) are not even.", H, W);
    STSWIN_CHECK_ARG(M % ((long)(H / 2) * (W / 2)) == 0, "patch-merging: row count is not a multiple of (H/2)*(W/2)");
  }
  return kOk;
}

}  // namespace

// see include/stswin_b200.h : stswin_layernorm_fwd
int layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, long M,
                  int Ctot, float eps, int pm, int H, int W, int C, cudaStream_t stream) {
  STSWIN_CHECK_ARG(x && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
  int rc = check_ln_shape(M, Ctot, pm, H, W, C);
  if (rc != kOk) return rc;
  PmGeom pg{pm, H, W, C};
  const int nv = (Ctot + 255) / 256;
  const int grid = ln_grid(M);
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto yb = static_cast<__nv_bfloat16*>(y);
#define STSWIN_LN_FWD(NV_) ln_fwd_kernel<NV_><<<grid, LN_THREADS, 0, stream>>>(xb, gamma, beta, yb, mean, rstd, M, Ctot, eps, pg)
  if (nv <= 1) STSWIN_LN_FWD(1);
  else if (nv <= 2) STSWIN_LN_FWD(2);
  else if (nv <= 4) STSWIN_LN_FWD(4);
  else STSWIN_LN_FWD(8);
#undef STSWIN_LN_FWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_layernorm_bwd
int layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                  const void* dres, void* dx, float* dgamma, float* dbeta, float* dx_colsum, long M, int Ctot, int pm,
                  int H, int W, int C, cudaStream_t stream) {
  STSWIN_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
  int rc = check_ln_shape(M, Ctot, pm, H, W, C);
  if (rc != kOk) return rc;
  STSWIN_CHECK_ARG(!(pm && dres), "layernorm_bwd: residual input is not supported together with the patch-merging scatter");
  PmGeom pg{pm, H, W, C};
  const int nv = (Ctot + 255) / 256;
  long want = (M + LN_WARPS - 1) / LN_WARPS;
  const int grid = (int)(want < num_sms() * 2 ? want : num_sms() * 2);   // 2 CTAs / SM when they fit; few column atomics
  const int smem = LN_WARPS * Ctot * 4;
  auto dyb = static_cast<const __nv_bfloat16*>(dy);
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto rb = static_cast<const __nv_bfloat16*>(dres);
  auto dxb = static_cast<__nv_bfloat16*>(dx);
#define STSWIN_LN_BWD(NV_)                                                                                         \
  do {                                                                                                             \
    if (dx_colsum) {                                                                                               \
      STSWIN_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      ln_bwd_kernel<NV_, true><<<grid, LN_THREADS, smem, stream>>>(dyb, xb, mean, rstd, gamma, rb, dxb, dgamma, dbeta, \
                                                                   dx_colsum, M, Ctot, pg);                        \
    } else {                                                                                                       \
      STSWIN_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      ln_bwd_kernel<NV_, false><<<grid, LN_THREADS, smem, stream>>>(dyb, xb, mean, rstd, gamma, rb, dxb, dgamma,   \
                                                                    dbeta, dx_colsum, M, Ctot, pg);                \
    }                                                                                                              \
  } while (0)
  if (nv <= 1) STSWIN_LN_BWD(1);
  else if (nv <= 2) STSWIN_LN_BWD(2);
  else if (nv <= 4) STSWIN_LN_BWD(4);
  else STSWIN_LN_BWD(8);
#undef STSWIN_LN_BWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_transpose
int transpose_cvt(const void* in, int in_f32, void* out, int out_f32, long batch, int R, int Cc, cudaStream_t stream) {
  STSWIN_CHECK_ARG(in && out && batch > 0 && R > 0 && Cc > 0, "transpose: bad argument");
  STSWIN_CHECK_ARG(batch <= 65535, "transpose: batch %ld exceeds gridDim.z", batch);
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, (unsigned)batch), block(32, 8);
  if (in_f32 && out_f32)
    transpose_kernel<float, float><<<grid, block, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), R, Cc);
  else if (in_f32)
    transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), R, Cc);
  else if (out_f32)
    transpose_kernel<__nv_bfloat16, float><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), R, Cc);
  else
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), R, Cc);
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
