// Micro-benchmark: what HBM bandwidth does the window-attention TMA access pattern reach on its own?
// 148 persistent CTAs stream [128 rows x 64 ch] bf16 chunks (16 KB) of a [B*T, H, W, 3C] tensor through
// a shared-memory ring with no compute at all.  Build: see tools/micro/build.sh.
//   mode 0: item = (tile, head), items strided over CTAs; per item Q0 K0 Q1 K1 V0 V1 (the kernel's order)
//   mode 1: a CTA takes one tile and walks its 4 heads back to back
//   mode 2: 2-D map over [tokens, 3C], 128 consecutive tokens per chunk (no window shape)
//   mode 3: like 0 but the window is fetched as four quadrant boxes (shifted blocks)
//   mode 4: stage-2 shape (run with argument 2): ws 4, four windows per chunk = four 4 KB boxes, 12 chunks per item
#include <cstdio>
#include <cstdlib>
#include "../../stswincl_b200/csrc/common.cuh"
#include "../../stswincl_b200/csrc/host_util.h"
using namespace stswin;
constexpr int SLOT = 16384;
__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmq,
              const __grid_constant__ CUtensorMap tm2, int mode, int ns, int num_tiles, int C, int nWw, int nW) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ns * SLOT);
  uint64_t* empty = full + ns;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int num_items = num_tiles * 4;
  if (threadIdx.x == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      int tile, head;
      if (mode == 1) { const int per = 4; const int k = it / gridDim.x; tile = (k / per) * gridDim.x + blockIdx.x; head = k % per; if (tile >= num_tiles) break; }
      else { tile = it >> 2; head = it & 3; }
      const int b = tile / nW, win = tile % nW, wh = win / nWw, ww = win % nWw;
      for (int step = 0; step < 6; ++step) {
        int which, c;
        if (step < 4) { which = step & 1; c = step >> 1; } else { which = 2; c = step - 4; }
        const int ch0 = which * C + head * 128 + c * 64;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], SLOT);
        uint8_t* dst = smem + slot * SLOT;
        if (mode == 2) tma_load_2d(dst, &tm2, &full[slot], ch0, tile * 128);
        else if (mode == 3) {
          for (int q = 0; q < 4; ++q)
            tma_load_4d(dst + q * 4096, &tmq, &full[slot], ch0, ww * 8 + (q & 1) * 4, wh * 8 + (q >> 1) * 4, b * 2);
        } else tma_load_4d(dst, &tm, &full[slot], ch0, ww * 8, wh * 8, b * 2);
        if (++slot == ns) { slot = 0; ph ^= 1; }
      }
    }
  } else if (threadIdx.x == 32) {
    int slot = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      if (mode == 1) { const int k = it / gridDim.x; const int tile = (k / 4) * gridDim.x + blockIdx.x; if (tile >= num_tiles) break; }
      for (int step = 0; step < 6; ++step) {
        mbar_wait(&full[slot], ph);
        mbar_arrive(&empty[slot]);
        if (++slot == ns) { slot = 0; ph ^= 1; }
      }
    }
  }
}
__global__ void __launch_bounds__(64, 1)
stream2_kernel(const __grid_constant__ CUtensorMap tm, int ns, int num_tiles, int C, int nWw, int nW, int issue_lanes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ns * SLOT);
  uint64_t* empty = full + ns;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int num_items = num_tiles * 4;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    int slot = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int tile = it >> 2, head = it & 3;
      for (int step = 0; step < 12; ++step) {
        const int which = step / 4, c = step % 4;
        const int ch0 = which * C + head * 256 + c * 64;
        mbar_wait(&empty[slot], ph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], SLOT);
        __syncwarp();
        uint8_t* dst = smem + slot * SLOT;
        for (int g = (issue_lanes == 1 ? 0 : lane); g < 4; g += (issue_lanes == 1 ? 1 : 32)) {
          if (issue_lanes == 1 && lane != 0) break;
          const int gw = tile * 4 + g, b = gw / nW, win = gw % nW, wh = win / nWw, ww = win % nWw;
          tma_load_4d(dst + g * 4096, &tm, &full[slot], ch0, ww * 4, wh * 4, b * 2);
        }
        if (++slot == ns) { slot = 0; ph ^= 1; }
      }
    }
  } else if (threadIdx.x == 32) {
    int slot = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x)
      for (int step = 0; step < 12; ++step) {
        mbar_wait(&full[slot], ph);
        mbar_arrive(&empty[slot]);
        if (++slot == ns) { slot = 0; ph ^= 1; }
      }
  }
}
int main2() {
  const int BT = 32, H = 32, W = 40, C = 1024, C3 = 3 * C;
  const size_t tokens = (size_t)BT * H * W;
  void* buf; cudaMalloc(&buf, tokens * C3 * 2); cudaMemset(buf, 1, tokens * C3 * 2);
  CUtensorMap tm;
  uint64_t dims[4] = {(uint64_t)C3, W, H, BT};
  uint64_t str[3] = {(uint64_t)C3 * 2, (uint64_t)W * C3 * 2, (uint64_t)H * W * C3 * 2};
  uint32_t box[4] = {64, 4, 4, 2};
  if (make_tmap(&tm, TmapDtype::BF16, 4, buf, dims, str, box, true)) { printf("tmap: %s\n", last_error()); return 1; }
  const int nWw = W / 4, nW = (H / 4) * nWw, num_tiles = (BT / 2) * nW / 4;
  cudaFuncSetAttribute(stream2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int il : {1, 32})
    for (int ns : {3, 6, 10}) {
      const int smem = 1024 + ns * SLOT + 256;
      for (int w = 0; w < 2; ++w) stream2_kernel<<<148, 64, smem>>>(tm, ns, num_tiles, C, nWw, nW, il);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) stream2_kernel<<<148, 64, smem>>>(tm, ns, num_tiles, C, nWw, nW, il);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      const double bytes = (double)tokens * C3 * 2;
      printf("stage-2 shape, %2d issuing lane(s), ring %2d slots: %.3f ms  %.0f GB/s  (%s)\n", il, ns, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
int main(int argc, char** argv) {
  if (argc > 1 && atoi(argv[1]) == 2) return main2();
  const int BT = 32, H = 64, W = 80, C = 512, C3 = 3 * C;
  const size_t tokens = (size_t)BT * H * W;
  void* buf; cudaMalloc(&buf, tokens * C3 * 2); cudaMemset(buf, 1, tokens * C3 * 2);
  CUtensorMap tm, tmq, tm2;
  uint64_t dims[4] = {(uint64_t)C3, W, H, BT};
  uint64_t str[3] = {(uint64_t)C3 * 2, (uint64_t)W * C3 * 2, (uint64_t)H * W * C3 * 2};
  uint32_t box[4] = {64, 8, 8, 2}, boxq[4] = {64, 4, 4, 2};
  if (make_tmap(&tm, TmapDtype::BF16, 4, buf, dims, str, box, true) || make_tmap(&tmq, TmapDtype::BF16, 4, buf, dims, str, boxq, true)) { printf("tmap: %s\n", last_error()); return 1; }
  uint64_t d2[2] = {(uint64_t)C3, tokens}; uint64_t s2[1] = {(uint64_t)C3 * 2}; uint32_t b2[2] = {64, 128};
  if (make_tmap(&tm2, TmapDtype::BF16, 2, buf, d2, s2, b2, true)) { printf("tmap2: %s\n", last_error()); return 1; }
  const int nWw = W / 8, nW = (H / 8) * nWw, num_tiles = (BT / 2) * nW;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; ++mode)
    for (int ns : {3, 6, 9, 12}) {
      const int smem = 1024 + ns * SLOT + 256;
      for (int w = 0; w < 2; ++w) stream_kernel<<<148, 64, smem>>>(tm, tmq, tm2, mode, ns, num_tiles, C, nWw, nW);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) stream_kernel<<<148, 64, smem>>>(tm, tmq, tm2, mode, ns, num_tiles, C, nWw, nW);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      const double bytes = (double)tokens * C3 * 2;
      printf("mode %d ring %2d slots: %.3f ms  %.0f GB/s  (%s)\n", mode, ns, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
