// Persistent, warp-specialised tcgen05 GEMM for the dense layers of the Swin block
// (K2 qkv, K4 proj, K6 fc1/fc2 of SURVEY.md section 2c; their dgrad and wgrad twins;
// the PatchMerging reduction).
//
//   D[M,N] = sum_k A[m,k] * B[n,k]          bf16 operands, fp32 accumulation in TMEM
//
// Operand storage ("major"): 0 = K-major  (A stored [M,K] row-major / B stored [N,K] row-major)
//                            1 = MN-major (A stored [K,M] row-major / B stored [K,N] row-major)
// so forward (x W^T) and dgrad (dy Wt^T, with a pre-transposed weight) are (0,0) and the
// weight gradient dW = dy^T x, which reduces over tokens, is (1,1) without any transposed
// copy of an activation.
//
// CTA = 3 warpgroups: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-3 idle, warps 4-11 epilogue;
// setmaxnreg gives the producer group 40 registers and the epilogue groups 232 (the epilogue keeps two 32-column
// accumulator pieces, the bias and the next tile's auxiliary data in registers).
// Tile 128 x 256 x 64, 5-stage TMA->smem ring (128B swizzle), two 256-column TMEM accumulators so
// the epilogue of tile i overlaps the main loop of tile i+1.  The epilogue goes TMEM -> registers
// -> swizzled smem (transpose) -> TMA stores, software-pipelined: the TMEM load of the next 32-column piece is in
// flight while the current piece is processed (tcgen05.wait::ld waits for every outstanding load, so a piece is
// requested right after the wait for its predecessor).  Fused: bias, residual / multiplier tile, erf-GELU together
// with its derivative (stored for the backward, which then only multiplies), column sums (bias
// gradients), and an fp32 TMA add-reduction for split-K weight gradients.
// (A first version stored through TMA from a single staging buffer and serialised on the store's
// read latency: 315 TFLOP/s at K=512; see profiles/.)
#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace stswin {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 5;
constexpr int A_STAGE_BYTES = BM * BK * 2;         // 16 KB: this CTA's 128 rows of the pair's 256-row A tile
constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;   // 16 KB: this CTA's half of the 256-column B tile
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BUF_BYTES = 2 * 32 * 128;  // two 32-row x 128-byte staging tiles per epilogue warp (ping-pong)
constexpr int EPI_WARP0 = 4;                  // warpgroup 0: TMA producer, MMA issuer, two idle warps; warpgroups 1-2: epilogue
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_WARPS * EPI_BUF_BYTES + 256;

enum Mode : int {
  kBias = 0,        // D = acc + bias
  kBiasRes = 1,     // D = acc + bias + aux
  kBiasGelu = 2,    // u = acc + bias ; D = gelu(u) ; D2 = gelu'(u)
  kMulAux = 3,      // D = acc * aux
  kF32Reduce = 4,   // D(fp32) += acc          (split-K, TMA add-reduction)
  kBiasGeluFwd = 5, // D = gelu(acc + bias)    (inference: no derivative output)
  kBiasGeluQ8 = 6,  // as kBiasGelu with D2 = gelu'(u) quantised to one byte (include/stswin_b200.h)
  kMulAuxQ8 = 7,    // D = acc * dequant(aux), aux = the uint8 D2 of kBiasGeluQ8
};

struct GemmArgs {
  int M, N, K;
  int mode;
  int k_splits;
  const float* bias;   // [N] or null
  float* colsum;       // [N] fp32, += column sums of the (bf16-rounded) D, or null
  __nv_bfloat16* D;    // bf16 outputs / aux are accessed with plain coalesced loads and stores
  __nv_bfloat16* D2;
  const __nv_bfloat16* aux;   // (uint8 in the kMulAuxQ8 mode; ld_aux in elements = bytes there)
  long ldd, ld_aux;
};

template <int A_MN, int B_MN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD /* D: bf16 [32 x 64] boxes, or the fp32 reduce target */,
            const __grid_constant__ CUtensorMap tmD2 /* second output of the GELU mode */, const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_stage = smem;
  uint8_t* s_out = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + EPI_WARPS * EPI_BUF_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;
  // A cluster is a pair of CTAs (two SMs of one TPC) that computes one 256 x 256 tile with
  // tcgen05.mma.cta_group::2: CTA r holds rows [r*128, +128) of A and columns [r*128, +128) of B, the
  // leader issues one M = 256 instruction for both tensor cores, each CTA accumulates its 128 rows
  // in its own TMEM and runs its own epilogue.  Per CTA and k-block that is 32 KB of TMA fill and
  // 8 KB of operand reads per MMA instead of 48 KB / 12 KB for two independent 128 x 256 tiles with a
  // multicast B: the single-CTA version sat on the shared-memory bandwidth (fill + operand reads +
  // epilogue staging ~ 220 B/clk against 128 B/clk) at 54 % tensor-pipe activity.
  const int num_mp = (num_m + 1) / 2;
  const int num_items = num_mp * num_n * p.k_splits;        // per cluster
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmD2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);           // leader's copy is the live one: one arrive.expect_tx per CTA of the pair
      mbar_init(&empty_bar[i], 1);          // multicast commit of the leader's MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 2 * EPI_WARPS);    // leader's copy: the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: barrier initialisation, TMEM allocation and the cluster handshake above overlapped the
  // previous kernel's tail; from here on its output is read / buffers it may still read are written
  pdl_wait();
  pdl_launch_dependents();

  if (warp < EPI_WARP0) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int n_blk = item % num_n;
        const int m_blk = 2 * ((item / num_n) % num_mp) + int(crank);   // may be >= num_m (ghost tile: zero fill)
        const int split = item / (num_n * num_mp);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = s_stage + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_STAGE_BYTES;
          const uint32_t lfull = mapa_u32(&full_bar[stage], 0);      // the leader's barrier collects both CTAs' bytes
          mbar_arrive_expect_tx_cluster(lfull, STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(sA + c * (BK * 128), &tmA, lfull, m_blk * BM + c * 64, kb * BK);
          } else {
            tma_load_2d_2sm(sA, &tmA, lfull, kb * BK, m_blk * BM);
          }
          if (B_MN) {               // this CTA's half of B: two of the four 64-column chunks
#pragma unroll
            for (int cc = 0; cc < BN / 128; ++cc)
              tma_load_2d_2sm(sB + cc * (BK * 128), &tmB, lfull, n_blk * BN + (int(crank) * (BN / 128) + cc) * 64, kb * BK);
          } else {                  // rows [crank*128, +128) of the 256-row B tile
            tma_load_2d_2sm(sB, &tmB, lfull, kb * BK, n_blk * BN + int(crank) * (BN / 2));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop, one elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers.  (Issued from inside `if (lane == 0)` every tcgen05.mma was
    // preceded by a serialised chain of R2UR moves, ~100 cycles per instruction: as long as the MMA itself.)
    if (crank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      const uint32_t stage_base = smem_u32(s_stage);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int split = item / (num_n * num_mp);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = stage_base + stage * STAGE_BYTES;
          const uint32_t sB = sA + A_STAGE_BYTES;
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t adesc = A_MN ? umma_smem_desc(sA + kk * 2048, BK * 128, 1024)
                                          : umma_smem_desc(sA + kk * 32, 16, 1024);
              const uint64_t bdesc = B_MN ? umma_smem_desc(sB + kk * 2048, BK * 128, 1024)
                                          : umma_smem_desc(sB + kk * 32, 16, 1024);
              umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            }
            umma_commit_2sm_mcast(&empty_bar[stage], 0x3);   // slot reusable (in both CTAs) once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit_2sm_mcast(&acc_full[acc], 0x3);   // accumulator complete (both CTAs' epilogues)
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------------ epilogue (8 warps)
    // warp w may touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter split the 256
    // accumulator columns in halves (two 64-column chunks each).
    const int ew = warp - EPI_WARP0;                 // 0..7
    const int wq = warp & 3;                         // TMEM lane quarter
    const int chalf = ew >> 2;                       // column half of the tile
    uint8_t* const my_bufs = s_out + ew * EPI_BUF_BYTES;   // two 32-row x 128-byte staging tiles
    uint8_t* my_buf = my_bufs;
    int chunk_no = 0;                                // chunks staged so far (selects the tile)
    constexpr bool has_aux = (MODE == kBiasRes || MODE == kMulAux || MODE == kMulAuxQ8);
    constexpr bool kQ8Aux = (MODE == kMulAuxQ8);           // one-byte aux: a 32 x 64-byte tile per chunk
    constexpr bool kGelu2 = (MODE == kBiasGelu || MODE == kBiasGeluQ8);   // two outputs
    constexpr bool kQ8Out = (MODE == kBiasGeluQ8);
    int acc = 0;
    uint32_t acc_phase = 0;
    // coalesced access pattern of a 32-row x 64-col bf16 chunk: instruction i of lane l touches
    // row 4*i + l/8, 16-byte column group l%8  (8 lanes = one 128-byte line)
    const int crow = lane >> 3, cchunk = lane & 7;
    constexpr int AUXV = kQ8Aux ? 4 : 8;             // 16-byte loads per lane and chunk
    uint4 auxr[has_aux ? 2 : 1][AUXV];               // aux of this warp's two chunks
    auto load_aux = [&](uint4 (&dst)[AUXV], int r0, int cb) {
      if constexpr (kQ8Aux) {                        // rows of 64 bytes: four lanes per row, eight rows per instruction
        const uint8_t* a8 = reinterpret_cast<const uint8_t*>(p.aux);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 8 * i + (lane >> 2), n = cb + (lane & 3) * 16;
          dst[i] = (r < p.M && n < p.N) ? __ldg(reinterpret_cast<const uint4*>(a8 + (size_t)r * p.ld_aux + n)) : make_uint4(0, 0, 0, 0);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 4 * i + crow, n = cb + cchunk * 8;
          dst[i] = (r < p.M && n < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.aux + (size_t)r * p.ld_aux + n))
                                        : make_uint4(0, 0, 0, 0);
        }
      }
    };
    // 16-byte group `chunk` (0..3) of row `row` inside a 64B-swizzled tile of 64-byte rows (the one-byte tiles)
    auto sw64_offset = [](uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); };
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int n_blk = item % num_n;
      const int m_blk = 2 * ((item / num_n) % num_mp) + int(crank);
      const int row0 = m_blk * BM + wq * 32;
      const int col0 = n_blk * BN + chalf * (BN / 2);
      const bool row_ok = (row0 + lane) < p.M;
      const uint32_t t_row = tmem_base + (uint32_t(wq * 32) << 16) + acc * BN + chalf * (BN / 2);

      if constexpr (MODE == kF32Reduce) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<uint4*>(my_buf + sw128_offset(lane, j)) = q;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && col0 + c * 32 < p.N && row0 < p.M) {
            tma_reduce_add_2d(&tmD, my_buf, col0 + c * 32, row0);
            tma_commit_group();
          }
        }
      } else {
        // aux (residual / multiplier) for BOTH 64-column chunks of this warp is requested up front, ahead of the wait
        // for the accumulator: 8 KB per warp, 64 KB per SM in flight -- one chunk ahead left the loads latency-bound.
        // (Requesting the NEXT item's aux as soon as a chunk's registers were staged was measured: x-aux mode
        // 964 -> 921 TFLOP/s at K = 512; these modes sit on the SM's shared-memory bandwidth, not on this latency.)
        if constexpr (has_aux) {
          load_aux(auxr[0], row0, col0);
          load_aux(auxr[1], row0, col0 + 64);
          // ... and the next item's aux lines are pulled into L2 now (one 128-byte line per lane and chunk; no register,
          // no scoreboard): its loads then pay the L2 latency instead of the HBM latency
          const int nitem = item + num_clusters;
          if (nitem < num_items) {
            const int nr = (2 * ((nitem / num_n) % num_mp) + int(crank)) * BM + wq * 32 + lane;
            const int nc = (nitem % num_n) * BN + chalf * (BN / 2);
            if (nr < p.M && nc < p.N) {
              constexpr int esz = kQ8Aux ? 1 : 2;
              const uint8_t* pa = reinterpret_cast<const uint8_t*>(p.aux) + ((size_t)nr * p.ld_aux + nc) * esz;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
              if (!kQ8Aux && nc + 64 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 128));
            }
          }
        }
        // staged 32 x 64 chunk -> global: one TMA box per warp and chunk (clipped at the matrix edge by the
        // tensor map); the warp goes on while the TMA engine reads the tile
        auto store_staged = [&](const CUtensorMap* tm, const uint8_t* buf, int cb) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (cb < p.N && row0 < p.M) tma_store_2d(tm, buf, cb, row0);
            tma_commit_group();
          }
        };
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        // The accumulator is read in four 32-column pieces (two per 64-column chunk) through two register sets: the
        // load of piece i+1 is issued right after the wait for piece i, so it is in flight while piece i is processed.
        // (Straight-line code on purpose: the destination registers of a load in flight must not be touched, which a
        // loop back-edge does not guarantee.)
        uint32_t va[32], vb[32];
        tmem_ld32(t_row, va);
        auto do_chunk = [&](auto c_tag) {
          constexpr int c = decltype(c_tag)::value;     // chunk of this warp's column half; also the aux register set
          constexpr bool last_chunk = (c == BN / 128 - 1);
          const int cbase = col0 + c * 64;
          // this chunk's staging tile; its previous TMA store (two chunks ago; the GELU mode's second
          // output: one chunk ago) must have been read out
          my_buf = my_bufs + (kGelu2 ? 0 : (chunk_no & 1)) * (32 * 128);
          ++chunk_no;
          if (lane == 0) tma_wait_group_read<1>();
          __syncwarp();
          uint32_t auxq[kQ8Aux ? 16 : 1];                 // one-byte aux: this row's 64 bytes
          if constexpr (kQ8Aux) {
            // the byte tile (32 x 64 bytes, 64B swizzle) has another row pitch than the output tile that is written over
            // it: every lane takes its whole row out before any output is staged
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(my_buf + sw64_offset(8 * i + (lane >> 2), lane & 3)) = auxr[c][i];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(my_buf + sw64_offset(lane, j));
              auxq[4 * j] = q.x; auxq[4 * j + 1] = q.y; auxq[4 * j + 2] = q.z; auxq[4 * j + 3] = q.w;
            }
          } else if constexpr (has_aux) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint4*>(my_buf + sw128_offset(4 * i + crow, cchunk)) = auxr[c][i];
          }
          __syncwarp();     // aux staged (each lane then only touches its own row: reads aux, writes output)
          uint32_t out2w[kGelu2 ? (kQ8Out ? 16 : 32) : 1];     // second output of the GELU modes (gelu'(u); bytes when quantised)
          // bias of the chunk's 64 columns, requested before the TMEM waits so that its latency overlaps theirs
          // (the first bias add was the top long-scoreboard stall of the epilogue); N is a multiple of 8, so a
          // group of four columns is inside the matrix or entirely outside
          float bias_r[2][32];
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = cbase + 4 * j;
              const float4 b4 = n < p.N ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
              bias_r[j >> 3][4 * (j & 7)] = b4.x; bias_r[j >> 3][4 * (j & 7) + 1] = b4.y;
              bias_r[j >> 3][4 * (j & 7) + 2] = b4.z; bias_r[j >> 3][4 * (j & 7) + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { bias_r[0][j] = 0.f; bias_r[1][j] = 0.f; }
          }
          auto do_piece = [&](auto half_tag, uint32_t (&v)[32]) {
            constexpr int half = decltype(half_tag)::value;
            const float (&bv)[32] = bias_r[half];        // bias of the 32 columns (same for every row)
            uint32_t auxw[(has_aux && !kQ8Aux) ? 16 : 1];             // this row's aux for the 32 columns of this piece
            if constexpr (has_aux && !kQ8Aux) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(my_buf + sw128_offset(lane, half * 4 + j));
                auxw[4 * j] = q.x; auxw[4 * j + 1] = q.y; auxw[4 * j + 2] = q.z; auxw[4 * j + 3] = q.w;
              }
            }
            uint32_t outw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if constexpr (MODE == kBiasGeluFwd) {
                outw[j] = gelu_x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), bv[2 * j], bv[2 * j + 1]);
              } else if constexpr (MODE == kBiasGelu) {
                gelu_and_grad_x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), bv[2 * j], bv[2 * j + 1], outw[j],
                                 out2w[half * 16 + j]);
              } else if constexpr (kQ8Out) {
                if ((j & 1) == 0) {                        // four columns -> one word of bytes
                  uint32_t q0, q1, q2, q3;
                  gelu_and_grad_q8_x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), bv[2 * j], bv[2 * j + 1], outw[j], q0, q1);
                  gelu_and_grad_q8_x2(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3]), bv[2 * j + 2], bv[2 * j + 3],
                                      outw[j + 1], q2, q3);
                  out2w[half * 8 + (j >> 1)] = __byte_perm(__byte_perm(q0, q1, 0x0040), __byte_perm(q2, q3, 0x0040), 0x5410);
                }
              } else if constexpr (kQ8Aux) {
                const uint64_t g2 = gelu_q8_dequant_x2(auxq[half * 8 + (j >> 1)], (2 * j) & 3, (2 * j + 1) & 3);
                float a0, a1;
                f2_unpack(f2_mul(f2_pack(__uint_as_float(v[2 * j]) + bv[2 * j], __uint_as_float(v[2 * j + 1]) + bv[2 * j + 1]), g2), a0, a1);
                outw[j] = pack_bf16(a0, a1);
              } else {
                float a0 = __uint_as_float(v[2 * j]) + bv[2 * j];
                float a1 = __uint_as_float(v[2 * j + 1]) + bv[2 * j + 1];
                if constexpr (MODE == kBiasRes) {
                  const float2 r = unpack_bf16(auxw[j]);
                  a0 += r.x; a1 += r.y;
                } else if constexpr (MODE == kMulAux) {
                  const float2 u = unpack_bf16(auxw[j]);
                  a0 *= u.x; a1 *= u.y;
                }
                outw[j] = pack_bf16(a0, a1);
              }
            }
            if (!row_ok) {                               // rows past M (last tile only): keep them out of colsum
#pragma unroll
              for (int j = 0; j < 16; ++j) outw[j] = 0u;
              if constexpr (MODE == kBiasGelu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) out2w[half * 16 + j] = 0u;
              }
              if constexpr (kQ8Out) {
#pragma unroll
                for (int j = 0; j < 8; ++j) out2w[half * 8 + j] = 0u;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(my_buf + sw128_offset(lane, half * 4 + j)) =
                  make_uint4(outw[4 * j], outw[4 * j + 1], outw[4 * j + 2], outw[4 * j + 3]);
          };
          // piece 0 of the chunk has landed in va; piece 1 goes into vb while piece 0 is processed
          tmem_ld_wait();
          tmem_ld32(t_row + c * 64 + 32, vb);
          do_piece(std::integral_constant<int, 0>{}, va);
          tmem_ld_wait();
          if constexpr (last_chunk) {
            // the tile's last TMEM read of this warp has landed in registers: hand the accumulator back now (the MMA
            // warp waits for it), not after the arithmetic, staging and stores of this piece
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));   // the leader's MMA warp owns both accumulators
          } else {
            tmem_ld32(t_row + (c + 1) * 64, va);        // the next chunk's piece 0
          }
          do_piece(std::integral_constant<int, 1>{}, vb);
          store_staged(&tmD, my_buf, cbase);
          if (p.colsum != nullptr) {
            // column sums of the staged 32 x 64 chunk: lane (rp = lane / 8, g = lane % 8) sums the 8 columns of 16-byte
            // group g over rows rp, rp + 4, ...; two shuffle steps fold the four row phases; lanes 0-7 then add 8 columns
            // each with two 16-byte vector reductions (16 L2 atomic transactions per warp and chunk instead of 64 scalar
            // ones: every CTA hits the same N addresses, and their serialisation cost 38 us of a 386 us launch)
            float cs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 q = *reinterpret_cast<const uint4*>(my_buf + sw128_offset(4 * i + crow, cchunk));
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16(w[e]);
                cs[2 * e] += f.x; cs[2 * e + 1] += f.y;
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
            }
            const int n = cbase + 8 * cchunk;             // N is a multiple of 8: the group is inside the matrix or outside
            if (lane < 8 && n < p.N) {
              float* dst = p.colsum + n;
              if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(dst + e, cs[e]);
              }
            }
          }
          if constexpr (kGelu2) {
            uint8_t* buf2 = my_bufs + 32 * 128;           // the second output has its own tile
            if (lane == 0) tma_wait_group_read<1>();      // its store of the previous chunk (the D store above may be in flight)
            __syncwarp();
            if constexpr (kQ8Out) {                       // 32 rows x 64 bytes, 64B swizzle
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(buf2 + sw64_offset(lane, j)) =
                    make_uint4(out2w[4 * j], out2w[4 * j + 1], out2w[4 * j + 2], out2w[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(buf2 + sw128_offset(lane, j)) =
                    make_uint4(out2w[4 * j], out2w[4 * j + 1], out2w[4 * j + 2], out2w[4 * j + 3]);
            }
            store_staged(&tmD2, buf2, cbase);
          }
        };
        static_assert(BN / 128 == 2, "two 64-column chunks per epilogue warp");
        do_chunk(std::integral_constant<int, 0>{});
        do_chunk(std::integral_constant<int, 1>{});
      }
      if constexpr (MODE == kF32Reduce) {
        // all TMEM reads of this accumulator are done -> hand it back to the MMA warp (the bf16 modes did so right
        // after their last TMEM load)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still multicast into / arrive on this CTA until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int A_MN, int B_MN, int MODE>
int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmD2,
                const GemmArgs& args, cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, MODE>;
  STSWIN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int num_m = (args.M + BM - 1) / BM, num_n = (args.N + BN - 1) / BN;
  const int items = ((num_m + 1) / 2) * num_n * args.k_splits;        // work items of a CTA pair
  const int max_clusters = num_sms() / 2;
  const int grid = 2 * (items < max_clusters ? items : max_clusters);
  STSWIN_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), (size_t)SMEM_BYTES, stream, tmA, tmB, tmD, tmD2, args));
  return kOk;
}

template <int A_MN, int B_MN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmD2,
           const GemmArgs& args, cudaStream_t stream) {
  switch (args.mode) {
    case kBias: return launch_mode<A_MN, B_MN, kBias>(tmA, tmB, tmD, tmD2, args, stream);
    case kBiasRes: return launch_mode<A_MN, B_MN, kBiasRes>(tmA, tmB, tmD, tmD2, args, stream);
    case kBiasGelu: return launch_mode<A_MN, B_MN, kBiasGelu>(tmA, tmB, tmD, tmD2, args, stream);
    case kBiasGeluFwd: return launch_mode<A_MN, B_MN, kBiasGeluFwd>(tmA, tmB, tmD, tmD2, args, stream);
    case kMulAux: return launch_mode<A_MN, B_MN, kMulAux>(tmA, tmB, tmD, tmD2, args, stream);
    case kBiasGeluQ8: return launch_mode<A_MN, B_MN, kBiasGeluQ8>(tmA, tmB, tmD, tmD2, args, stream);
    case kMulAuxQ8: return launch_mode<A_MN, B_MN, kMulAuxQ8>(tmA, tmB, tmD, tmD2, args, stream);
    default: return launch_mode<A_MN, B_MN, kF32Reduce>(tmA, tmB, tmD, tmD2, args, stream);
  }
}

}  // namespace

// see include/stswin_b200.h : stswin_gemm_bf16
int gemm_bf16(const void* A, int a_major, long lda, const void* B, int b_major, long ldb, void* D, long ldd, void* D2,
              const void* aux, long ld_aux, const float* bias, float* colsum, int M, int N, int K, int mode,
              int k_splits, cudaStream_t stream) {
  STSWIN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  STSWIN_CHECK_ARG(mode >= kBias && mode <= kMulAuxQ8, "gemm: unknown epilogue mode %d", mode);
  STSWIN_CHECK_ARG(a_major == 0 || a_major == 1, "gemm: a_major must be 0 or 1");
  STSWIN_CHECK_ARG(b_major == 0 || b_major == 1, "gemm: b_major must be 0 or 1");
  STSWIN_CHECK_ARG(!(a_major == 1 && b_major == 0), "gemm: (A MN-major, B K-major) is not instantiated");
  STSWIN_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
  STSWIN_CHECK_ARG(A && B && D, "gemm: null operand");
  if (k_splits < 1) k_splits = 1;
  STSWIN_CHECK_ARG(k_splits == 1 || mode == kF32Reduce, "gemm: split-K requires the fp32 reduce epilogue");
  STSWIN_CHECK_ARG((mode != kBiasGelu && mode != kBiasGeluQ8) || D2 != nullptr, "gemm: gelu epilogue needs the derivative output D2");
  STSWIN_CHECK_ARG((mode != kBiasRes && mode != kMulAux && mode != kMulAuxQ8) || aux != nullptr, "gemm: epilogue mode %d needs aux", mode);
  STSWIN_CHECK_ARG((mode != kBiasGeluQ8 && mode != kMulAuxQ8) || (N % 16 == 0 && (mode == kBiasGeluQ8 ? ldd : ld_aux) % 16 == 0),
                   "gemm: the one-byte GELU' modes need N and the leading dimension of the byte matrix to be multiples of 16");
  const int kb_total = (K + BK - 1) / BK;
  if (k_splits > kb_total) k_splits = kb_total;
  // every split must own at least one k-block (an empty split would publish an unwritten accumulator)
  while (k_splits > 1 && (k_splits - 1) * ((kb_total + k_splits - 1) / k_splits) >= kb_total) --k_splits;

  CUtensorMap tmA, tmB, tmD, tmD2;
  int rc;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (a_major == 0) { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = BM; }
    else              { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    str[0] = (uint64_t)lda * 2;
    if ((rc = make_tmap(&tmA, TmapDtype::BF16, 2, A, dims, str, box, true)) != kOk) return rc;
    if (b_major == 0) { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = BN / 2; }   // each CTA of a pair loads half
    else              { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    str[0] = (uint64_t)ldb * 2;
    if ((rc = make_tmap(&tmB, TmapDtype::BF16, 2, B, dims, str, box, true)) != kOk) return rc;
    tmD = tmA;
    tmD2 = tmA;
    if (mode == kF32Reduce) {
      STSWIN_CHECK_ARG(ldd % 4 == 0, "gemm: fp32 ldd must be a multiple of 4");
      dims[0] = N; dims[1] = M;
      str[0] = (uint64_t)ldd * 4; box[0] = 32; box[1] = 32;
      if ((rc = make_tmap(&tmD, TmapDtype::F32, 2, D, dims, str, box, true)) != kOk) return rc;
    } else {
      STSWIN_CHECK_ARG(ldd % 8 == 0 && N % 8 == 0, "gemm: N and ldd must be multiples of 8 elements");
      STSWIN_CHECK_ARG((reinterpret_cast<uintptr_t>(D) & 15) == 0, "gemm: D must be 16-byte aligned");
      STSWIN_CHECK_ARG(D2 == nullptr || (reinterpret_cast<uintptr_t>(D2) & 15) == 0, "gemm: D2 must be 16-byte aligned");
      if (mode == kMulAuxQ8)
        STSWIN_CHECK_ARG((reinterpret_cast<uintptr_t>(aux) & 15) == 0, "gemm: aux must be 16-byte aligned");
      if (mode == kBiasRes || mode == kMulAux)
        STSWIN_CHECK_ARG(ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0,
                         "gemm: aux must be 16-byte aligned with ld_aux a multiple of 8");
      // bf16 outputs leave through TMA: [32 rows x 64 columns] boxes of the 128B-swizzled staging tiles
      dims[0] = N; dims[1] = M;
      str[0] = (uint64_t)ldd * 2; box[0] = 64; box[1] = 32;
      if ((rc = make_tmap(&tmD, TmapDtype::BF16, 2, D, dims, str, box, true)) != kOk) return rc;
      if (D2 != nullptr && mode != kBiasGeluQ8 && (rc = make_tmap(&tmD2, TmapDtype::BF16, 2, D2, dims, str, box, true)) != kOk) return rc;
      if (mode == kBiasGeluQ8) {      // one byte per element: [32 rows x 64 bytes] boxes, 64B swizzle
        str[0] = (uint64_t)ldd;
        if ((rc = make_tmap(&tmD2, TmapDtype::U8, 2, D2, dims, str, box, false, true)) != kOk) return rc;
      }
    }
  }
  GemmArgs args{M, N, K, mode, k_splits, bias, colsum, static_cast<__nv_bfloat16*>(D), static_cast<__nv_bfloat16*>(D2),
                static_cast<const __nv_bfloat16*>(aux), ldd, ld_aux};
  if (a_major == 0 && b_major == 0) return launch<0, 0>(tmA, tmB, tmD, tmD2, args, stream);
  if (a_major == 0 && b_major == 1) return launch<0, 1>(tmA, tmB, tmD, tmD2, args, stream);
  return launch<1, 1>(tmA, tmB, tmD, tmD2, args, stream);
}

}  // namespace stswin
