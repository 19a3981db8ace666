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

// element (row, 8-element group v) lives at base + row_base(row) + group_off(v): the group offset does not depend on
// the row, so each lane keeps its NV offsets in registers and a row costs two 32-bit divisions, not one per group
template <bool PM>
__device__ __forceinline__ long row_base(long row, int Ctot, const PmGeom& g) {
  if (!PM) return row * Ctot;
  const int H2 = g.H >> 1, W2 = g.W >> 1;
  const unsigned per_img = H2 * W2;
  const unsigned r32 = static_cast<unsigned>(row);      // gathered rows are counted in 32 bits (checked by the launcher)
  const unsigned bt = r32 / per_img;
  const int rem = int(r32 - bt * per_img);
  const int h2 = rem / W2, w2 = rem - h2 * W2;
  return ((long(bt) * g.H + 2 * h2) * g.W + 2 * w2) * g.C;      // the 2x2 neighbourhood's top-left token
}
template <bool PM>
__device__ __forceinline__ int group_off(int v, const PmGeom& g) {
  if (!PM) return v * 8;
  const int ch = v * 8;
  const int seg = ch / g.C;                 // 0:(dh0,dw0) 1:(dh1,dw0) 2:(dh0,dw1) 3:(dh1,dw1)   (:266-270)
  const int dh = seg & 1, dw = seg >> 1;
  return (dh * g.W + dw) * g.C + (ch - seg * g.C);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Both kernels stream rows through a per-warp ring in shared memory filled by cp.async (16 bytes per
// lane and group, no registers held while a row is in flight): NSTG-1 rows ahead per warp keep
// ~100 KB per SM in flight, which one-row-per-warp register kernels could not (they reached 41-55 % of
// the HBM roofline).  A lane only ever reads back the 16-byte groups it copied itself, so
// cp.async.wait_group is the only synchronisation the ring needs.
template <int NV, int NSTG, bool PM>   // NV = ceil(Ctot / 256): 16-byte groups per lane; PM: PatchMerging gather
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, long M, int Ctot,
              float eps, PmGeom pg) {
  extern __shared__ uint4 s_ring4[];     // [LN_WARPS][NSTG][Ctot / 8]
  pdl_wait();                            // programmatic dependent launch: the launch itself overlapped the previous kernel
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = Ctot >> 3;
  const float inv_c = 1.0f / Ctot;        // exact for the power-of-two widths of the model; a multiply per row instead of a division
  uint4* ring = s_ring4 + (size_t)warp * NSTG * nvec;
  const long stride = (long)gridDim.x * LN_WARPS;
  const long row0 = (long)blockIdx.x * LN_WARPS + warp;
  int goff[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) goff[i] = group_off<PM>(lane + 32 * i, pg);
  auto issue = [&](long row, int stg) {
    if (row < M) {
      const __nv_bfloat16* src = x + row_base<PM>(row, Ctot, pg);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) cp_async16(ring + stg * nvec + vi, src + goff[i]);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < NSTG - 1; ++s) issue(row0 + s * stride, s);
  // gamma / beta of this lane's columns stay in registers for every row of the warp, as fp32 pairs: the row
  // arithmetic runs on packed FADD2 / FFMA2 (two columns per issue slot -- these kernels are issue-bound at the
  // power-capped clock of a training step, not HBM-bound)
  uint64_t gr[NV][4], br[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gr[i][k] = vi < nvec ? f2_pack(__ldg(gamma + vi * 8 + 2 * k), __ldg(gamma + vi * 8 + 2 * k + 1)) : f2_pack(0.f, 0.f);
      br[i][k] = vi < nvec ? f2_pack(__ldg(beta + vi * 8 + 2 * k), __ldg(beta + vi * 8 + 2 * k + 1)) : f2_pack(0.f, 0.f);
    }
  }
  int stg = 0;
  for (long row = row0; row < M; row += stride) {
    issue(row + (NSTG - 1) * stride, stg == 0 ? NSTG - 1 : stg - 1);
    cp_async_wait<NSTG - 1>();
    uint64_t v[NV][4];
    uint64_t s2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 q = ring[stg * nvec + vi];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          v[i][k] = f2_pack(f.x, f.y);
          s2 = f2_add(s2, v[i][k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[i][k] = f2_pack(0.f, 0.f);
      }
    }
    float s_lo, s_hi;
    f2_unpack(s2, s_lo, s_hi);
    const float mu = warp_sum(s_lo + s_hi) * inv_c;
    const uint64_t nmu2 = f2_pack(-mu, -mu);
    uint64_t ss2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint64_t d = f2_add(v[i][k], nmu2); ss2 = f2_fma(d, d, ss2); }
      }
    float ss_lo, ss_hi;
    f2_unpack(ss2, ss_lo, ss_hi);
    const float rs = rsqrtf(warp_sum(ss_lo + ss_hi) * inv_c + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    const uint64_t rs2 = f2_pack(rs, rs), nm2 = f2_pack(-mu * rs, -mu * rs);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float y0, y1;
          f2_unpack(f2_fma(f2_fma(v[i][k], rs2, nm2), gr[i][k], br[i][k]), y0, y1);
          w[k] = pack_bf16(y0, y1);
        }
        *reinterpret_cast<uint4*>(y + row * Ctot + vi * 8) = make_uint4(w[0], w[1], w[2], w[3]);   // output rows are always dense
      }
    }
    if (++stg == NSTG) stg = 0;
  }
  cp_async_wait<0>();
}

template <int NV, bool COLSUM, int NSTG, bool PM, bool RES>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dx_colsum, long M, int Ctot, PmGeom pg) {
  pdl_wait();                            // programmatic dependent launch: the launch itself overlapped the previous kernel
  pdl_launch_dependents();
  extern __shared__ uint4 s_ring4[];     // [LN_WARPS][NSTG][NARR][Ctot / 8]; reused as [LN_WARPS][Ctot] floats at the end
  constexpr int NARR = RES ? 3 : 2;      // dy, x (, residual gradient)
  float* s_red = reinterpret_cast<float*>(s_ring4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = Ctot >> 3;
  const float inv_c = 1.0f / Ctot;        // exact for the power-of-two widths of the model; a multiply per row instead of a division
  uint4* ring = s_ring4 + (size_t)warp * NSTG * NARR * nvec;
  // column accumulators and the row arithmetic are fp32 pairs (packed FADD2 / FMUL2 / FFMA2: two columns per issue slot)
  uint64_t a_dg[NV][4], a_db[NV][4], a_cs[COLSUM ? NV : 1][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a_dg[i][k] = f2_pack(0.f, 0.f); a_db[i][k] = f2_pack(0.f, 0.f);
      if (COLSUM) a_cs[i][k] = f2_pack(0.f, 0.f);
    }
  constexpr bool has_res = RES;
  const long stride = (long)gridDim.x * LN_WARPS;
  const long row0 = (long)blockIdx.x * LN_WARPS + warp;
  int goff[PM ? NV : 1];                 // plain rows: the offset is vi * 8
  if constexpr (PM) {
#pragma unroll
    for (int i = 0; i < NV; ++i) goff[i] = group_off<PM>(lane + 32 * i, pg);
  }
  auto issue = [&](long row, int stg) {
    if (row < M) {
      uint4* d = ring + stg * NARR * nvec;
      const __nv_bfloat16* xsrc = x + row_base<PM>(row, Ctot, pg);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
          cp_async16(d + vi, dy + row * Ctot + vi * 8);
          cp_async16(d + nvec + vi, xsrc + (PM ? goff[PM ? i : 0] : vi * 8));
          if (has_res) cp_async16(d + 2 * nvec + vi, dres + row * Ctot + vi * 8);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < NSTG - 1; ++s) issue(row0 + s * stride, s);
  uint64_t gmr[NV][4];                   // gamma of this lane's columns
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      gmr[i][k] = vi < nvec ? f2_pack(__ldg(gamma + vi * 8 + 2 * k), __ldg(gamma + vi * 8 + 2 * k + 1)) : f2_pack(0.f, 0.f);
  }
  int stg = 0;
  float n_mu = 0.f, n_rs = 0.f;
  if (row0 < M) { n_mu = mean[row0]; n_rs = rstd[row0]; }
  for (long row = row0; row < M; row += stride) {
    issue(row + (NSTG - 1) * stride, stg == 0 ? NSTG - 1 : stg - 1);
    const float mu = n_mu, rs = n_rs;
    if (row + stride < M) { n_mu = mean[row + stride]; n_rs = rstd[row + stride]; }
    cp_async_wait<NSTG - 1>();
    const uint4* cur = ring + stg * NARR * nvec;
    const uint64_t nmu2 = f2_pack(-mu, -mu), rs2 = f2_pack(rs, rs);
    uint64_t s1p = f2_pack(0.f, 0.f), s2p = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 cd = cur[vi], cx = cur[nvec + vi];
        const uint32_t wd[4] = {cd.x, cd.y, cd.z, cd.w}, wx[4] = {cx.x, cx.y, cx.z, cx.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fd = unpack_bf16(wd[k]), fx = unpack_bf16(wx[k]);
          const uint64_t d2 = f2_pack(fd.x, fd.y);
          const uint64_t h = f2_mul(f2_add(f2_pack(fx.x, fx.y), nmu2), rs2);      // normalised input
          a_dg[i][k] = f2_fma(d2, h, a_dg[i][k]);
          a_db[i][k] = f2_add(a_db[i][k], d2);
          const uint64_t q = f2_mul(d2, gmr[i][k]);
          s1p = f2_add(s1p, q);
          s2p = f2_fma(q, h, s2p);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    f2_unpack(s1p, s1a, s1b);
    f2_unpack(s2p, s2a, s2b);
    const float m1 = warp_sum(s1a + s1b) * inv_c, m2 = warp_sum(s2a + s2b) * inv_c;
    const uint64_t nm1 = f2_pack(-m1, -m1), nm2 = f2_pack(-m2, -m2);
    __nv_bfloat16* dx_row = dx + row_base<PM>(row, Ctot, pg);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 cd = cur[vi], cx = cur[nvec + vi];
        const uint32_t wd[4] = {cd.x, cd.y, cd.z, cd.w}, wx[4] = {cx.x, cx.y, cx.z, cx.w};
        uint4 cr = make_uint4(0, 0, 0, 0);
        if (has_res) cr = cur[2 * nvec + vi];
        const uint32_t wr[4] = {cr.x, cr.y, cr.z, cr.w};
        uint32_t wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fd = unpack_bf16(wd[k]), fx = unpack_bf16(wx[k]);
          const uint64_t h = f2_mul(f2_add(f2_pack(fx.x, fx.y), nmu2), rs2);
          // rs * (dy * gamma - m1 - h * m2)
          uint64_t o = f2_mul(rs2, f2_fma(h, nm2, f2_fma(f2_pack(fd.x, fd.y), gmr[i][k], nm1)));
          if (has_res) {
            const float2 fr = unpack_bf16(wr[k]);
            o = f2_add(o, f2_pack(fr.x, fr.y));
          }
          float o0, o1;
          f2_unpack(o, o0, o1);
          wo[k] = pack_bf16(o0, o1);
          if (COLSUM) {                      // sum what the consumer will read (bf16-rounded)
            const float2 fo = unpack_bf16(wo[k]);
            a_cs[i][k] = f2_add(a_cs[i][k], f2_pack(fo.x, fo.y));
          }
        }
        *reinterpret_cast<uint4*>(dx_row + (PM ? goff[PM ? i : 0] : vi * 8)) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
      }
    }
    if (++stg == NSTG) stg = 0;
  }
  cp_async_wait<0>();
  // CTA-level column reduction, one quantity at a time through s_red[LN_WARPS][Ctot] (the ring is idle now)
  auto flush = [&](uint64_t (&acc)[NV][4], float* out) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float lo, hi;
          f2_unpack(acc[i][k], lo, hi);
          s_red[warp * Ctot + vi * 8 + 2 * k] = lo;
          s_red[warp * Ctot + vi * 8 + 2 * k + 1] = hi;
        }
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

// dst[b][i] = src[b][i], 16 bytes per access, four loads in flight per thread
__global__ void __launch_bounds__(256)
copy_strided_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long dst_stride16, long src_stride16, long n16) {
  pdl_wait();
  pdl_launch_dependents();
  const uint4* s = src + (long)blockIdx.y * src_stride16;
  uint4* d = dst + (long)blockIdx.y * dst_stride16;
  const long step = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * step < n16; i += 4 * step) {
    const uint4 a = __ldg(s + i), b = __ldg(s + i + step), c = __ldg(s + i + 2 * step), e = __ldg(s + i + 3 * step);
    d[i] = a; d[i + step] = b; d[i + 2 * step] = c; d[i + 3 * step] = e;
  }
  for (; i < n16; i += step) d[i] = __ldg(s + i);
}

// bf16 [batch, R, Cc] -> [batch, Cc, R] without shared memory: a thread loads an 8 x 8 block (eight 16-byte row
// segments, all in flight together), transposes it in registers -- one PRMT per output word pairs the elements of two
// source rows; the rest is register renaming -- and stores eight 16-byte segments of the transposed block.  A warp covers
// 64 rows x 32 columns (lane = 4 * row block + column group): loads are 64-byte row pieces, stores are whole 128-byte lines
// of the output (writes are the scarcer direction of this GPU's HBM).  CTA tile 128 x 128, no barrier.
// (The 64 x 64 shared-memory version with 2-byte scattered stores reached 39-44 % of the HBM roofline inside the step.)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int Cc) {
  pdl_wait();
  pdl_launch_dependents();
  const long b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y * 128 + (warp >> 2) * 64 + (lane >> 2) * 8;      // first of this thread's 8 rows
  const int c = blockIdx.x * 128 + (warp & 3) * 32 + (lane & 3) * 8;        // first of its 8 columns
  if (r >= R || c >= Cc) return;                                            // R, Cc are multiples of 8: whole blocks
  const __nv_bfloat16* ib = in + b * (long)R * Cc + (long)r * Cc + c;
  __nv_bfloat16* ob = out + b * (long)R * Cc + (long)c * R + r;
  uint4 q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ldg(reinterpret_cast<const uint4*>(ib + (long)i * Cc));
#pragma unroll
  for (int k = 0; k < 8; ++k) {                   // output row c + k = source column c + k of rows r .. r + 7
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = reinterpret_cast<const uint32_t*>(&q[2 * j])[k >> 1];        // row 2j,   columns k & ~1, +1
      const uint32_t hi = reinterpret_cast<const uint32_t*>(&q[2 * j + 1])[k >> 1];    // row 2j+1
      w[j] = __byte_perm(lo, hi, (k & 1) ? 0x7632 : 0x5410);
    }
    *reinterpret_cast<uint4*>(ob + (long)k * R) = make_uint4(w[0], w[1], w[2], w[3]);
  }
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
  STSWIN_CHECK_ARG(!pm || M < (1L << 31), "layernorm_fwd: more than 2^31 gathered rows");
  int rc = check_ln_shape(M, Ctot, pm, H, W, C);
  if (rc != kOk) return rc;
  PmGeom pg{pm, H, W, C};
  const int nv = (Ctot + 255) / 256;
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto yb = static_cast<__nv_bfloat16*>(y);
  const long want = (M + LN_WARPS - 1) / LN_WARPS;
  // ring: LN_WARPS x NSTG rows of Ctot bf16; ~64 KB per CTA, as many CTAs per SM as fit
#define STSWIN_LN_FWD(NV_, NSTG_, PM_)                                                                             \
  do {                                                                                                          \
    const int smem = LN_WARPS * NSTG_ * Ctot * 2;                                                               \
    const int per_sm = smem <= 56 * 1024 ? 4 : (smem <= 112 * 1024 ? 2 : 1);                                    \
    const int grid = (int)(want < (long)num_sms() * per_sm ? want : (long)num_sms() * per_sm);                  \
    STSWIN_CUDA(cudaFuncSetAttribute(ln_fwd_kernel<NV_, NSTG_, PM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    STSWIN_CUDA(launch_pdl(ln_fwd_kernel<NV_, NSTG_, PM_>, dim3(grid), dim3(LN_THREADS), (size_t)smem, stream, xb, gamma, beta, yb, \
                           mean, rstd, M, Ctot, eps, pg));                                                       \
  } while (0)
  if (pm) {                      // PatchMerging rows: 4*C channels
    if (nv <= 4) STSWIN_LN_FWD(4, 4, true);
    else STSWIN_LN_FWD(8, 3, true);
  } else if (nv <= 1) STSWIN_LN_FWD(1, 8, false);
  else if (nv <= 2) STSWIN_LN_FWD(2, 6, false);
  else if (nv <= 4) STSWIN_LN_FWD(4, 4, false);
  else STSWIN_LN_FWD(8, 3, false);
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
  STSWIN_CHECK_ARG(!pm || M < (1L << 31), "layernorm_bwd: more than 2^31 gathered rows");
  STSWIN_CHECK_ARG(!(pm && dres), "layernorm_bwd: residual input is not supported together with the patch-merging scatter");
  PmGeom pg{pm, H, W, C};
  const int nv = (Ctot + 255) / 256;
  const long want = (M + LN_WARPS - 1) / LN_WARPS;
  auto dyb = static_cast<const __nv_bfloat16*>(dy);
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto rb = static_cast<const __nv_bfloat16*>(dres);
  auto dxb = static_cast<__nv_bfloat16*>(dx);
  // ring: LN_WARPS x NSTG x 3 rows of Ctot bf16 (dy, x, residual gradient); <= 2 CTAs per SM (few column atomics)
#define STSWIN_LN_BWD_K(NV_, CS_, NSTG_, PM_, RES_)                                                                \
  do {                                                                                                          \
    const int smem = LN_WARPS * NSTG_ * (RES_ ? 3 : 2) * Ctot * 2;                                              \
    const int per_sm = (smem <= 112 * 1024 && NV_ <= 2) ? 2 : 1;                                                \
    const int grid = (int)(want < (long)num_sms() * per_sm ? want : (long)num_sms() * per_sm);                  \
    STSWIN_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV_, CS_, NSTG_, PM_, RES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    STSWIN_CUDA(launch_pdl(ln_bwd_kernel<NV_, CS_, NSTG_, PM_, RES_>, dim3(grid), dim3(LN_THREADS), (size_t)smem, stream, dyb, xb, \
                           mean, rstd, gamma, rb, dxb, dgamma, dbeta, dx_colsum, M, Ctot, pg));                  \
  } while (0)
  // stages: with / without the residual-gradient row (3 / 2 arrays per stage)
#define STSWIN_LN_BWD(NV_, NS_RES_, NS_NORES_, PM_)                                     \
  do {                                                                                  \
    if (dx_colsum && rb) STSWIN_LN_BWD_K(NV_, true, NS_RES_, PM_, true);                \
    else if (dx_colsum) STSWIN_LN_BWD_K(NV_, true, NS_NORES_, PM_, false);              \
    else if (rb) STSWIN_LN_BWD_K(NV_, false, NS_RES_, PM_, true);                       \
    else STSWIN_LN_BWD_K(NV_, false, NS_NORES_, PM_, false);                            \
  } while (0)
  if (pm) {                      // no residual input with the patch-merging scatter (checked above)
    if (nv <= 4) { if (dx_colsum) STSWIN_LN_BWD_K(4, true, 6, true, false); else STSWIN_LN_BWD_K(4, false, 6, true, false); }
    else { if (dx_colsum) STSWIN_LN_BWD_K(8, true, 3, true, false); else STSWIN_LN_BWD_K(8, false, 3, true, false); }
  } else if (nv <= 1) STSWIN_LN_BWD(1, 6, 8, false);
  else if (nv <= 2) STSWIN_LN_BWD(2, 4, 6, false);
  else if (nv <= 4) STSWIN_LN_BWD(4, 4, 6, false);
  else STSWIN_LN_BWD(8, 2, 3, false);
#undef STSWIN_LN_BWD_K
#undef STSWIN_LN_BWD
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_copy_strided
// column sums of a bf16 matrix [R, C] (C a multiple of 8): a CTA owns a slab of rows, a thread one 16-byte column
// group; fp32 partial sums leave through one atomic per column and CTA
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long R,
                                                           int C, int rows_per_cta) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups = C >> 3;                              // 16-byte groups per row
  const int lanes_r = 256 / groups > 0 ? 256 / groups : 1; // rows walked in parallel by the CTA
  __shared__ float s_acc[256][9];
  const int g = threadIdx.x % groups, rsub = threadIdx.x / groups;
  const long r0 = (long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (long r = r0 + rsub; r < r1 && rsub < lanes_r; r += lanes_r) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + r * C + g * 8));
    const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_acc[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (rsub == 0) {                                         // one atomic per column and CTA
    for (int rs = 1; rs < lanes_r; ++rs)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += s_acc[rs * groups + g][k];
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(out + g * 8 + k, acc[k]);
  }
}

int copy_strided(void* dst, long dst_stride, const void* src, long src_stride, long bytes, int batches, cudaStream_t stream) {
  STSWIN_CHECK_ARG(dst && src && bytes > 0 && batches > 0, "copy_strided: bad argument");
  STSWIN_CHECK_ARG(batches <= 65535, "copy_strided: %d batches exceed gridDim.y", batches);
  STSWIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)dst_stride |
                     (uintptr_t)src_stride | (uintptr_t)bytes) & 15) == 0,
                   "copy_strided: pointers, strides and size must be multiples of 16 bytes");
  const long n16 = bytes / 16;
  long per = (n16 + 256 * 4 - 1) / (256 * 4);                       // blocks that give every thread four vectors
  const long cap = ((long)num_sms() * 8 + batches - 1) / batches;   // ~8 blocks per SM in total
  if (per > cap) per = cap;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)batches);
  STSWIN_CUDA(launch_pdl(copy_strided_kernel, dim3(grid), dim3(256), 0, stream, static_cast<uint4*>(dst), static_cast<const uint4*>(src), dst_stride / 16,
                                                src_stride / 16, n16));
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

// see include/stswin_b200.h : stswin_transpose
int transpose_cvt(const void* in, int in_f32, void* out, int out_f32, long batch, int R, int Cc, cudaStream_t stream) {
  STSWIN_CHECK_ARG(in && out && batch > 0 && R > 0 && Cc > 0, "transpose: bad argument");
  STSWIN_CHECK_ARG(batch <= 65535, "transpose: batch %ld exceeds gridDim.z", batch);
  if (!in_f32 && !out_f32 && R % 8 == 0 && Cc % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    dim3 g64((Cc + 127) / 128, (R + 127) / 128, (unsigned)batch);
    STSWIN_CUDA(launch_pdl(transpose_bf16_kernel, g64, dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), R, Cc));
    STSWIN_CUDA(cudaGetLastError());
    return kOk;
  }
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

int colsum_bf16(const void* x, float* out, long R, int C, cudaStream_t stream) {
  STSWIN_CHECK_ARG(x && out && R > 0 && C > 0, "colsum: bad argument");
  STSWIN_CHECK_ARG(C % 8 == 0 && C <= 2048, "colsum: C=%d must be a multiple of 8, <= 2048", C);
  STSWIN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "colsum: x must be 16-byte aligned");
  int ctas = 4 * num_sms();
  long rows_per_cta = (R + ctas - 1) / ctas;
  if (rows_per_cta < 32) rows_per_cta = 32;
  ctas = (int)((R + rows_per_cta - 1) / rows_per_cta);
  STSWIN_CUDA(launch_pdl(colsum_bf16_kernel, dim3(ctas), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(x), out, R, C, (int)rows_per_cta));
  STSWIN_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace stswin
