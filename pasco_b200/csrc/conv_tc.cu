// Sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   out[o, :] = Σ_k  in[nbr[k, o], :] @ W[k]           (output-stationary implicit GEMM)
//
// One persistent CTA per SM walks 128-row output tiles.  Warp roles:
// A CTA owns a GROUP of T = 256/Cout consecutive tiles whose accumulators live side by side in TMEM, and walks
// group → offset k → 64-channel block → tile, so every weight slice W[k] is fetched from L2 once per group
// instead of once per tile.
//   warps 0-3  gather: load the neighbour rows of a tile (coalesced 16-byte loads, 16 lanes per 256-byte row
//              segment; neighbour indices arrive through a cp.async ring one offset ahead), optionally apply
//              the fused BatchNorm affine + activation of the producing layer, split fp32 → bf16 hi (+ lo), and
//              store the 128x64 tile into shared memory in the 128-byte-swizzled K-major UMMA layout.  Loads of
//              slot s+1 are in flight in registers while slot s is converted and stored.
//   warp  9    streams the pre-swizzled weight slices with bulk async copies (UBLKCP) into their own ring.
//   warp  8    one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) into the tile's TMEM accumulator
//              and commits to the A / B "empty" mbarriers; accumulator sets are double-buffered in TMEM so
//              the epilogue of group g overlaps the main loop of group g+1.
//   warps 4-7  epilogue: tcgen05.ld the accumulators (lane = output row), add bias, store fp32 rows.
// No atomics: every output row is written exactly once, results are run-to-run deterministic.
//
// precision 3 (the "fp32" mode): operands split as x = hi + lo (bf16 each) and three MMAs
// hi·hi + lo·hi + hi·lo accumulate in fp32 — ~2^-16 relative error, well inside the 1e-3 parity bound.
#include "common.cuh"
#include "umma.cuh"
#include <cstdlib>

using namespace pasco;
using namespace umma;

namespace {

constexpr int BLOCK_M = 128;
constexpr int KBLK = 64;                    // channels per K-block (= one 128-byte swizzle row of bf16)
constexpr int A_TILE_BYTES = BLOCK_M * 128; // 16 KB
#ifndef PASCO_GATHER_WARPS
#define PASCO_GATHER_WARPS 8
#endif
constexpr int NUM_GATHER_WARPS = PASCO_GATHER_WARPS;            // 8 (16 tile rows each) or 16 (8 rows each)
constexpr int NUM_EPI_WARPS = 4;                                // next 4 warps: warp % 4 = TMEM lane quadrant
constexpr int MMA_WARP = NUM_GATHER_WARPS + NUM_EPI_WARPS;
constexpr int LOAD_WARP = MMA_WARP + 1;
constexpr int NUM_THREADS = (LOAD_WARP + 1) * 32;              // 448 / 704
constexpr int ROWS_PER_WARP = BLOCK_M / NUM_GATHER_WARPS;       // 16 / 8
constexpr int LOADS_PER_SLOT = ROWS_PER_WARP / 2;               // float4 per lane per slot
constexpr int TILES_PER_IDX_STEP = 32 / ROWS_PER_WARP;          // tiles whose indices one warp-wide cp.async covers
constexpr int MAX_STAGES = 8;

struct ConvParams {
  const float* in;
  const int32_t* nbr;  // [K, n_out] or nullptr (identity)
  const uint8_t* wpk;  // packed weights
  const float* bias;
  const float* in_scale;
  const float* in_shift;
  float* out;
  double* stats;       // optional [2*Cout]: column sums / sums of squares of `out` accumulate here (fused BN statistics)
  int64_t n_out;
  int64_t out_pitch;   // row stride of out (floats)
  int64_t in_pitch;    // row stride of in (floats)
  int K, Cin, Cout, in_act;
  int sa, sb, tiles_per_group, tmem_cols;
  int koff_base, koff_step;   // weight slice of table row k = koff_base + koff_step * k
  int ksplit, kchunk;         // split-K: work item v = (tile group v / ksplit, offsets [ks*kchunk, (ks+1)*kchunk)), partial
                              // results go to out + ks * n_out * out_pitch (summed by k_splitk_reduce)
  // k_conv_pl only: the input as pre-split bf16 planes x = hi + lo, [n_in, Cin] each with row pitch pl_pitch (elements)
  const uint16_t* pl_hi;
  const uint16_t* pl_lo;
  int64_t pl_pitch;
  int pl_depth;               // slots whose copies are in flight per producer thread before the oldest is published
};

// ------------------------------------------------------------------------------------------------------------
// weight packing: fp32 [K, Cin, Cout] → per (k, kb): [hi image | lo image], image = N rows x 128 B, swizzled
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pack_weights(const float* __restrict__ W, int K, int Cin, int Cout, int transpose,
                               uint8_t* __restrict__ packed) {
  const int N = transpose ? Cin : Cout;   // B rows
  const int Kc = transpose ? Cout : Cin;  // contraction length
  const int KB = Kc / KBLK;
  int64_t total = (int64_t)K * KB * N * 8;  // one thread per 16-byte chunk (8 bf16)
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t & 7);
    int64_t r = t >> 3;
    int n = (int)(r % N);
    r /= N;
    int kb = (int)(r % KB);
    int k = (int)(r / KB);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int kc = kb * KBLK + c * 8 + j;
      v[j] = transpose ? __ldg(W + ((int64_t)k * Cin + n) * Cout + kc) : __ldg(W + ((int64_t)k * Cin + kc) * Cout + n);
    }
    uint4 hi, lo;
    uint2 h0, l0, h1, l1;
    split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
    split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
    hi = make_uint4(h0.x, h0.y, h1.x, h1.y);
    lo = make_uint4(l0.x, l0.y, l1.x, l1.y);
    int64_t tile = ((int64_t)k * KB + kb) * (int64_t)N * 256;  // hi image then lo image
    int64_t off = (int64_t)n * 128 + ((c ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(packed + tile + off) = hi;
    *reinterpret_cast<uint4*>(packed + tile + (int64_t)N * 128 + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float z, int act) {
  if (act == 1) return fmaxf(z, 0.f);
  if (act == 2) return z > 0.f ? z : 0.01f * z;
  return z;
}

// ------------------------------------------------------------------------------------------------------------
// cp.async helpers for the neighbour-index ring
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int IDX_RING = 6;  // neighbour-index ring depth (entries of T*128 ints, one per (tile group, offset))

// position of a gather warp inside the flat slot sequence  group → k → kb → t
struct SlotIt {
  int64_t v;       // work item handled by this CTA: tile group v / ksplit, offset range number v % ksplit
  int64_t group;   // = v / ksplit
  int gk;          // running (work item, k) counter → index ring slot
  int k, k1, kb, t;  // k walks [k0(v), k1(v))
  int t_eff;       // tiles in this group
  bool valid;
};

// ------------------------------------------------------------------------------------------------------------
// consumer-side roles shared by the register-gather kernel (k_conv_tc) and the plane-gather kernel (k_conv_pl)
// ------------------------------------------------------------------------------------------------------------
struct Pipe {
  uint8_t* a_smem;          // [sa][A_hi | A_lo]
  uint8_t* b_smem;          // [sb][B_hi | B_lo]
  uint64_t *afull, *aempty, *bfull, *bempty, *tfull, *tempty;
  uint32_t tmem_base;
  double* stat_acc;         // [2*Cout] per-CTA accumulators in shared memory (nullptr: no fused statistics)
  int a_stage_bytes, b_stage_bytes, b_tile;
  int KB, T, acc_cols;
  int64_t num_tiles, num_groups;
};

// one thread: streams the pre-swizzled weight slices with bulk async copies (UBLKCP) into the B ring
__device__ __forceinline__ void role_weight_loader(const ConvParams& p, const Pipe& pl) {
  int stage = 0;
  uint32_t phase = 0;
  const uint32_t bytes = (uint32_t)pl.b_stage_bytes;
  for (int64_t v = blockIdx.x; v < pl.num_groups * p.ksplit; v += gridDim.x) {
    const int k0 = (int)(v % p.ksplit) * p.kchunk, k1 = k0 + p.kchunk < p.K ? k0 + p.kchunk : p.K;
    for (int k = k0; k < k1; ++k) {
      for (int kb = 0; kb < pl.KB; ++kb) {
        mbar_wait(smem_u32(pl.bempty + stage), phase ^ 1);
        mbar_arrive_expect_tx(smem_u32(pl.bfull + stage), bytes);
        const uint8_t* src = p.wpk + ((int64_t)(p.koff_base + p.koff_step * k) * pl.KB + kb) * (int64_t)p.Cout * 256;
        bulk_g2s(smem_u32(pl.b_smem + (size_t)stage * pl.b_stage_bytes), src, bytes, smem_u32(pl.bfull + stage));
        if (++stage == p.sb) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
}

// one thread: tcgen05.mma (M=128, N=Cout, K=16) into the tile's TMEM accumulator; commits free the A / B ring slots
// `nm` issuer threads (one per warp) share the work: issuer `mid` owns the tiles t ≡ mid (mod nm) of every group — tiles
// of a group accumulate into different TMEM columns, so their MMAs are independent.  One issuer needs ~100 cycles of
// descriptor / uniform-register work per tcgen05.mma while an M=128, N=64 MMA executes in 32: with a single issuer the
// tensor pipe idled at 30 % (ncu: the issuer thread busy on every instruction of its loop, producers waiting for stages).
// The "B stage free" and "accumulators ready" barriers expect one commit per issuer.
template <int NSPLIT, bool CPASYNC_A = false>
__device__ __forceinline__ void role_mma(const ConvParams& p, const Pipe& pl, int mid = 0, int nm = 1) {
  // Called by a WHOLE warp: waits, counters and descriptor arithmetic are warp-uniform (the compiler keeps them in uniform
  // registers), only tcgen05.mma / tcgen05.commit are issued by the elected lane.  (With the loop inside `if (lane == 0)`
  // every descriptor went through ELECT + R2UR moves: ~13 instructions and ~100 cycles per 32-cycle MMA.)
  const bool lead = elect_one();
  const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.Cout, 0, 0);
  int sa = 0, sb = 0;
  uint32_t pa = 0, pb = 0;
  int it = 0;
  for (int64_t v = blockIdx.x; v < pl.num_groups * p.ksplit; v += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t group = v / p.ksplit;
    const int k0 = (int)(v % p.ksplit) * p.kchunk, k1 = k0 + p.kchunk < p.K ? k0 + p.kchunk : p.K;
    const int64_t rem = pl.num_tiles - group * pl.T;
    const int t_eff = pl.T;        // the last group is padded to T tiles (rows >= n_out: never gathered, never stored)
    (void)rem;
    mbar_wait(smem_u32(pl.tempty + buf), ((it >> 1) & 1) ^ 1);
    tc_fence_after();
    for (int k = k0; k < k1; ++k) {
      for (int kb = 0; kb < pl.KB; ++kb) {
        mbar_wait(smem_u32(pl.bfull + sb), pb);
        const uint32_t b_hi = smem_u32(pl.b_smem + (size_t)sb * pl.b_stage_bytes), b_lo = b_hi + pl.b_tile;
        for (int t = 0; t < t_eff; ++t) {
          if (t % nm == mid) {
          mbar_wait(smem_u32(pl.afull + sa), pa);
          if (CPASYNC_A) fence_proxy_async_smem();   // the A stage was written by cp.async copies (generic proxy)
          tc_fence_after();
          const uint32_t a_hi = smem_u32(pl.a_smem + (size_t)sa * pl.a_stage_bytes), a_lo = a_hi + A_TILE_BYTES;
          const uint32_t d_tmem = pl.tmem_base + (uint32_t)(buf * pl.acc_cols + t * p.Cout);
          // descriptors of the k-step j: the start-address field (bits 0-13, 16-byte units) advances by 2 per 32 bytes
          const uint64_t da_hi0 = make_desc_sw128(a_hi, 16, 1024), db_hi0 = make_desc_sw128(b_hi, 16, 1024);
          const uint64_t da_lo0 = make_desc_sw128(a_lo, 16, 1024), db_lo0 = make_desc_sw128(b_lo, 16, 1024);
          if (lead) {
#pragma unroll
            for (int j = 0; j < KBLK / 16; ++j) {
              const uint32_t accum = (k > k0 || kb > 0 || j > 0) ? 1u : 0u;
              mma_bf16(d_tmem, da_hi0 + 2 * j, db_hi0 + 2 * j, idesc, accum);
              if (NSPLIT == 3) {
                mma_bf16(d_tmem, da_lo0 + 2 * j, db_hi0 + 2 * j, idesc, 1);
                mma_bf16(d_tmem, da_hi0 + 2 * j, db_lo0 + 2 * j, idesc, 1);
              }
            }
            mma_commit(smem_u32(pl.aempty + sa));
          }
          __syncwarp();
          }
          if (++sa == p.sa) {
            sa = 0;
            pa ^= 1;
          }
        }
        if (lead) mma_commit(smem_u32(pl.bempty + sb));
        __syncwarp();
        if (++sb == p.sb) {
          sb = 0;
          pb ^= 1;
        }
      }
    }
    if (lead) mma_commit(smem_u32(pl.tfull + buf));
    __syncwarp();
  }
}

// Column sums over the 32 rows held by a warp (lane = row, v[j] = column j): butterfly transpose-reduce, 31 shuffles;
// lane j returns Σ_rows v[j].  Destroys v.
__device__ __forceinline__ float warp_col_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// one warp per TMEM lane quadrant q: tcgen05.ld the accumulators (lane = output row), add bias, store fp32 rows;
// optionally accumulate the per-channel sum and sum of squares of what was stored (the BatchNorm statistics of the
// next layer) — per warp in fp32 over its 32 rows, then in fp64 in shared memory, one global fp64 atomic per CTA.
__device__ __forceinline__ void role_epilogue(const ConvParams& p, const Pipe& pl, int q, int lane) {
  int it = 0;
  const bool want_stats = pl.stat_acc != nullptr;
  for (int64_t v = blockIdx.x; v < pl.num_groups * p.ksplit; v += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t group = v / p.ksplit;
    float* out_part = p.out + (v % p.ksplit) * p.n_out * p.out_pitch;    // this split's partial output
    const int64_t rem = pl.num_tiles - group * pl.T;
    const int t_eff = pl.T;        // the last group is padded to T tiles (rows >= n_out: never gathered, never stored)
    (void)rem;
    mbar_wait(smem_u32(pl.tfull + buf), (it >> 1) & 1);
    tc_fence_after();
    for (int t = 0; t < t_eff; ++t) {
      const int64_t row = (group * pl.T + t) * BLOCK_M + q * 32 + lane;
      const uint32_t taddr = pl.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * pl.acc_cols + t * p.Cout);
      float* orow = out_part + row * p.out_pitch;
      const bool live = row < p.n_out;
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        const bool full = c0 + 32 <= p.Cout;          // else a 16-column tail (Cout % 32 == 16)
        float v[32];
        if (full) {
          tmem_ld32(taddr + c0, v);
        } else {
          float u[16];
          tmem_ld16(taddr + c0, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = u[j];
            v[j + 16] = 0.f;
          }
        }
        tmem_ld_wait();
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full || j < 16) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
        }
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (full || j < 16) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (want_stats) {
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = live ? v[j] : 0.f;
            sq[j] = v[j] * v[j];
          }
          const float s1 = warp_col_sums32(v, lane);
          const float s2 = warp_col_sums32(sq, lane);
          if (c0 + lane < p.Cout) {
            atomicAdd(pl.stat_acc + c0 + lane, (double)s1);
            atomicAdd(pl.stat_acc + p.Cout + c0 + lane, (double)s2);
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(pl.tempty + buf));
  }
  if (want_stats) {
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps
    for (int i = q * 32 + lane; i < 2 * p.Cout; i += NUM_EPI_WARPS * 32) atomicAdd(p.stats + i, pl.stat_acc[i]);
  }
}

template <int NSPLIT, bool PROLOGUE>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int n_op = (NSPLIT == 3) ? 2 : 1;
  const int b_tile = p.Cout * 128;
  const int a_stage_bytes = n_op * A_TILE_BYTES;
  const int b_stage_bytes = n_op * b_tile;
  const int T = p.tiles_per_group;
  uint8_t* a_smem = smem;                                           // [sa][A_hi | A_lo]
  uint8_t* b_smem = smem + (size_t)p.sa * a_stage_bytes;            // [sb][B_hi | B_lo]
  int* idx_ring = reinterpret_cast<int*>(b_smem + (size_t)p.sb * b_stage_bytes);   // [IDX_RING][T*128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(idx_ring + IDX_RING * T * BLOCK_M);
  uint64_t* afull = bars;                         // [MAX_STAGES]
  uint64_t* aempty = bars + MAX_STAGES;           // [MAX_STAGES]
  uint64_t* bfull = bars + 2 * MAX_STAGES;        // [MAX_STAGES]
  uint64_t* bempty = bars + 3 * MAX_STAGES;       // [MAX_STAGES]
  uint64_t* tfull = bars + 4 * MAX_STAGES;        // [2]
  uint64_t* tempty = bars + 4 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * MAX_STAGES + 4);
  double* stat_acc = p.stats ? reinterpret_cast<double*>(tmem_slot + 4) : nullptr;   // [2*Cout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin / KBLK;
  const int64_t num_tiles = (p.n_out + BLOCK_M - 1) / BLOCK_M;
  const int64_t num_groups = (num_tiles + T - 1) / T;
  const int acc_cols = T * p.Cout;                // TMEM columns of one accumulator set

  if (stat_acc)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += NUM_THREADS) stat_acc[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(smem_u32(afull + s), NUM_GATHER_WARPS);
      mbar_init(smem_u32(aempty + s), 1);
      mbar_init(smem_u32(bfull + s), 1);
      mbar_init(smem_u32(bempty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(tfull + b), 1);
      mbar_init(smem_u32(tempty + b), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  Pipe pl;
  pl.a_smem = a_smem; pl.b_smem = b_smem;
  pl.afull = afull; pl.aempty = aempty; pl.bfull = bfull; pl.bempty = bempty; pl.tfull = tfull; pl.tempty = tempty;
  pl.tmem_base = tmem_base;
  pl.stat_acc = stat_acc;
  pl.a_stage_bytes = a_stage_bytes; pl.b_stage_bytes = b_stage_bytes; pl.b_tile = b_tile;
  pl.KB = KB; pl.T = T; pl.acc_cols = acc_cols; pl.num_tiles = num_tiles; pl.num_groups = num_groups;

  if (warp < NUM_GATHER_WARPS) {
    // ===================================== gather producers =====================================
    const int chunk = lane & 15;  // 16-byte fp32 chunk inside the 64-channel block
    const int rsub = lane >> 4;   // which of the 2 rows this half-warp handles per step
    const bool affine = PROLOGUE && p.in_scale != nullptr;

    // async prefetch of the neighbour indices of (group, k) into ring slot gk % IDX_RING: this warp's 32 rows of
    // each of the group's tiles
    auto prefetch_idx = [&](int64_t group, int k, int gk, int t_eff) {
      int* dst = idx_ring + (gk % IDX_RING) * T * BLOCK_M;
      // lanes 0-15 fetch tile t, lanes 16-31 tile t+1: the warp's 16 rows of each
      for (int t = lane / ROWS_PER_WARP; t < t_eff; t += TILES_PER_IDX_STEP) {
        const int64_t row = (group * T + t) * BLOCK_M + warp * ROWS_PER_WARP + (lane % ROWS_PER_WARP);
        int* d = dst + t * BLOCK_M + warp * ROWS_PER_WARP + (lane % ROWS_PER_WARP);
        if (row < p.n_out) {
          if (p.nbr) cp_async4(smem_u32(d), p.nbr + (int64_t)k * p.n_out + row);
          else *d = (int)row;
        } else {
          *d = -1;
        }
      }
      cp_async_commit();
    };
    const int64_t num_items = num_groups * p.ksplit;
    // position `it` at the first offset of work item v
    auto enter = [&](SlotIt& it, int64_t v) {
      it.v = v;
      it.valid = v < num_items;
      if (it.valid) {
        it.group = v / p.ksplit;
        it.k = (int)(v % p.ksplit) * p.kchunk;
        it.k1 = it.k + p.kchunk < p.K ? it.k + p.kchunk : p.K;
        const int64_t rem = num_tiles - it.group * T;
        it.t_eff = T;              // padded to T tiles (see role_mma)
        (void)rem;
      }
    };
    auto advance = [&](SlotIt& it) {
      if (++it.t < it.t_eff) return;
      it.t = 0;
      if (++it.kb < KB) return;
      it.kb = 0;
      ++it.gk;
      if (++it.k < it.k1) return;
      enter(it, it.v + gridDim.x);
    };
    // the (work item, k) pair that follows `it`'s by `ahead` steps, for index prefetch
    auto prefetch_ahead = [&](const SlotIt& it, int ahead) {
      int64_t v = it.v;
      int k = it.k + ahead, k1 = it.k1;
      const int gk = it.gk + ahead;
      while (v < num_items && k >= k1) {
        const int over = k - k1;
        v += gridDim.x;
        if (v >= num_items) break;
        const int k0 = (int)(v % p.ksplit) * p.kchunk;
        k1 = k0 + p.kchunk < p.K ? k0 + p.kchunk : p.K;
        k = k0 + over;
      }
      if (v < num_items) {
        const int64_t g = v / p.ksplit;
        prefetch_idx(g, k, gk, T);
      } else {
        cp_async_commit();  // keep the group count uniform
      }
    };

    SlotIt ld;
    ld.gk = 0; ld.kb = 0; ld.t = 0; ld.group = 0; ld.k = 0; ld.k1 = 0; ld.t_eff = 0;
    enter(ld, blockIdx.x);
    if (ld.valid) {
      // prime the index ring: entries 0 .. IDX_RING-2
      for (int a = 0; a < IDX_RING - 1; ++a) prefetch_ahead(ld, a);
    }
    SlotIt st = ld;
    int stage = 0;
    uint32_t phase = 0;

    float4 va[LOADS_PER_SLOT], vb[LOADS_PER_SLOT];
    uint32_t valid_a = 0, valid_b = 0;

    // issue the 16 row-segment loads of one slot into registers
    auto issue = [&](SlotIt& it, float4 (&v)[LOADS_PER_SLOT], uint32_t& valid) {
      if (it.kb == 0 && it.t == 0) {
        // first use of ring entry gk: it was prefetched IDX_RING-1 entries ago; keep the ring primed
        cp_async_wait<IDX_RING - 2>();
        __syncwarp();
        prefetch_ahead(it, IDX_RING - 1);
      }
      const int* irow = idx_ring + (it.gk % IDX_RING) * T * BLOCK_M + it.t * BLOCK_M + warp * ROWS_PER_WARP;
      const int cbase = it.kb * KBLK + chunk * 4;
      valid = 0;
      int srcs[LOADS_PER_SLOT];
#pragma unroll
      for (int i = 0; i < LOADS_PER_SLOT; ++i) srcs[i] = irow[i * 2 + rsub];
#pragma unroll
      for (int i = 0; i < LOADS_PER_SLOT; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (srcs[i] >= 0) {
          v[i] = __ldg(reinterpret_cast<const float4*>(p.in + (int64_t)srcs[i] * p.in_pitch + cbase));
          valid |= 1u << i;
        }
      }
    };
    // convert + store one slot into its A stage and publish it
    auto store = [&](const SlotIt& it, float4 (&v)[LOADS_PER_SLOT], uint32_t valid) {
      mbar_wait(smem_u32(aempty + stage), phase ^ 1);
      uint8_t* dst = a_smem + (size_t)stage * a_stage_bytes;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (PROLOGUE && affine) {
        const int cbase = it.kb * KBLK + chunk * 4;
        sc = __ldg(reinterpret_cast<const float4*>(p.in_scale + cbase));
        sh = __ldg(reinterpret_cast<const float4*>(p.in_shift + cbase));
      }
      // byte offset of (row warp*16 + rsub, 16-byte chunk) inside the swizzled tile; row i*2 adds i*256 and flips
      // the chunk by (i*2 & 7)
      const uint32_t row_base = (uint32_t)(warp * ROWS_PER_WARP + rsub);
#pragma unroll
      for (int i = 0; i < LOADS_PER_SLOT; ++i) {
        float4 x = v[i];
        if (PROLOGUE && ((valid >> i) & 1u)) {
          x.x = act_apply(fmaf(x.x, sc.x, sh.x), p.in_act);
          x.y = act_apply(fmaf(x.y, sc.y, sh.y), p.in_act);
          x.z = act_apply(fmaf(x.z, sc.z, sh.z), p.in_act);
          x.w = act_apply(fmaf(x.w, sc.w, sh.w), p.in_act);
        }
        const uint32_t trow = row_base + i * 2;
        const uint32_t off = trow * 128u + ((((uint32_t)chunk >> 1) ^ (trow & 7u)) << 4) + (((uint32_t)chunk & 1u) << 3);
        if (NSPLIT == 3) {
          uint2 hi, lo;
          split4(x, hi, lo);
          *reinterpret_cast<uint2*>(dst + off) = hi;
          *reinterpret_cast<uint2*>(dst + A_TILE_BYTES + off) = lo;
        } else {
          *reinterpret_cast<uint2*>(dst + off) = to_bf16x4(x);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(afull + stage));
      if (++stage == p.sa) {
        stage = 0;
        phase ^= 1;
      }
    };

    if (ld.valid) {
      issue(ld, va, valid_a);
      advance(ld);
      while (st.valid) {
        if (ld.valid) {
          issue(ld, vb, valid_b);
          advance(ld);
        }
        store(st, va, valid_a);
        advance(st);
        if (!st.valid) break;
        if (ld.valid) {
          issue(ld, va, valid_a);
          advance(ld);
        }
        store(st, vb, valid_b);
        advance(st);
      }
    }
    cp_async_wait<0>();
  } else if (warp == LOAD_WARP) {
    if (lane == 0) role_weight_loader(p, pl);
  } else if (warp == MMA_WARP) {
    role_mma<NSPLIT>(p, pl);
  } else {
    role_epilogue(p, pl, warp - NUM_GATHER_WARPS, lane);   // == warp % 4: the TMEM lane quadrant this warp may read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Plane-gather variant: the input arrives PRE-SPLIT as bf16 planes (hi [+ lo]) and the producers only move bytes
// ------------------------------------------------------------------------------------------------------------
// ncu of k_conv_tc (round 1): 3.5 long-scoreboard stall cycles per issued instruction, L2 at 15 % — every gathered
// fp32 value crosses the register file and is converted 27x (once per kernel offset).  Here the split x = hi + lo is
// done ONCE per tensor by the producing pass (k_split_planes / the BatchNorm apply), and a gather warp issues
// 16-byte cp.async (LDGSTS) copies straight from the plane rows into the 128-byte-swizzled UMMA tile:
//   * no registers on the data path, so `pl_depth` whole A stages (32 KB each in fp32 mode) are in flight per SM;
//   * a missing neighbour is a ZERO-FILL copy (src-size 0): no global or L2 traffic at all for the ~38 % of
//     (row, offset) pairs that have no neighbour — the TMA gather4 variant of round 1 had to fetch a zero row for each;
//   * 8 copies per lane per slot instead of ~600 gather/convert/store instructions.
// Completion is tracked by the hardware: after issuing the copies of a slot every producer thread executes
// cp.async.mbarrier.arrive.noinc on the slot's "full" barrier (expected count = 256 producer threads), so the barrier
// completes when the last copy has landed — no wait_group, no software publish step, and the producers run up to `sa`
// slots ahead of the MMA (a first version published slot n-D in software after issuing slot n: ncu showed the MMA
// thread and the producers waiting on each other's hand-off, 1.43 ms).  The MMA thread adds a generic→async proxy
// fence after its barrier wait (CUTLASS's sm100 cp.async collective issues none; it costs one instruction per slot).
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all its cp.async copies issued so far have completed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int PL_MMA_WARPS = 4;                                      // tcgen05.mma issuers (one thread each), see role_mma
constexpr int PL_MMA_WARP0 = LOAD_WARP + 1;                          // issuer 0 = MMA_WARP, issuers 1..3 = warps 14..16
constexpr int NUM_THREADS_PL = (PL_MMA_WARP0 + PL_MMA_WARPS - 1) * 32;   // 544

template <int NSPLIT>
__global__ void __launch_bounds__(NUM_THREADS_PL, 1) k_conv_pl(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int n_op = (NSPLIT == 3) ? 2 : 1;
  const int b_tile = p.Cout * 128;
  const int a_stage_bytes = n_op * A_TILE_BYTES;
  const int b_stage_bytes = n_op * b_tile;
  const int T = p.tiles_per_group;
  uint8_t* a_smem = smem;                                           // [sa][A_hi | A_lo]
  uint8_t* b_smem = smem + (size_t)p.sa * a_stage_bytes;            // [sb][B_hi | B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)p.sb * b_stage_bytes);
  uint64_t* afull = bars;                         // [MAX_STAGES]
  uint64_t* aempty = bars + MAX_STAGES;           // [MAX_STAGES]
  uint64_t* bfull = bars + 2 * MAX_STAGES;        // [MAX_STAGES]
  uint64_t* bempty = bars + 3 * MAX_STAGES;       // [MAX_STAGES]
  uint64_t* tfull = bars + 4 * MAX_STAGES;        // [2]
  uint64_t* tempty = bars + 4 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * MAX_STAGES + 4);
  double* stat_acc = p.stats ? reinterpret_cast<double*>(tmem_slot + 4) : nullptr;   // [2*Cout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin / KBLK;
  const int64_t num_tiles = (p.n_out + BLOCK_M - 1) / BLOCK_M;
  const int64_t num_groups = (num_tiles + T - 1) / T;
  const int acc_cols = T * p.Cout;

  // active issuers.  A parity wait is only sound for a waiter that sees EVERY phase of the barrier, so a stage of the A ring
  // must always be consumed by the same issuer: slot n holds tile n % T (groups are padded to T tiles) and sits in stage
  // n % sa, hence n_mma must divide both T and sa.  (4 issuers on a 6-stage ring gave rare wrong tiles: an issuer came
  // back to a stage two fills later and its parity wait passed one fill early.)
  int n_mma = 1;
  for (int c = 2; c <= PL_MMA_WARPS && c <= p.pl_depth; c <<= 1)
    if (T % c == 0 && p.sa % c == 0) n_mma = c;
  if (stat_acc)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += NUM_THREADS_PL) stat_acc[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(smem_u32(afull + s), 32);   // one cp.async arrival per lane of the warp that fills the slot
      mbar_init(smem_u32(aempty + s), 1);
      mbar_init(smem_u32(bfull + s), 1);
      mbar_init(smem_u32(bempty + s), n_mma);   // one commit per issuer
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(tfull + b), n_mma);
      mbar_init(smem_u32(tempty + b), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  Pipe pl;
  pl.a_smem = a_smem; pl.b_smem = b_smem;
  pl.afull = afull; pl.aempty = aempty; pl.bfull = bfull; pl.bempty = bempty; pl.tfull = tfull; pl.tempty = tempty;
  pl.tmem_base = tmem_base;
  pl.stat_acc = stat_acc;
  pl.a_stage_bytes = a_stage_bytes; pl.b_stage_bytes = b_stage_bytes; pl.b_tile = b_tile;
  pl.KB = KB; pl.T = T; pl.acc_cols = acc_cols; pl.num_tiles = num_tiles; pl.num_groups = num_groups;

  if (warp < NUM_GATHER_WARPS) {
    // ===================================== plane-gather producers =====================================
    // ONE WARP PER SLOT: warp w fills the slots n ≡ w (mod 8) of the CTA's slot sequence (work item → k → kb → tile),
    // all 128 rows of the tile.  (A first version let every warp fill 16 rows of EVERY slot: ncu showed ~300
    // instructions of per-slot bookkeeping per warp for 8 copies, 900 cycles per slot against 384 cycles of MMA.)
    // One copy instruction moves 4 rows: lane = (row sub-index 0..3, 16-byte chunk 0..7 of the 128-byte row); 32
    // instructions per plane per slot.  Lane l holds the neighbour indices of rows l, l+32, l+64, l+96 (four coalesced
    // 128-byte loads, fetched two of the warp's slots = 16 CTA slots ahead), distributed by one shuffle per instruction.
    const int ch = lane & 7, sub = lane >> 3;
    const int64_t num_items = num_groups * p.ksplit;
    struct It {
      int64_t v, group;
      int k, k1, kb, t, t_eff;
      bool valid;
    };
    auto enter = [&](It& it, int64_t v) {
      it.v = v;
      it.valid = v < num_items;
      it.kb = 0; it.t = 0;
      if (it.valid) {
        it.group = v / p.ksplit;
        it.k = (int)(v % p.ksplit) * p.kchunk;
        it.k1 = it.k + p.kchunk < p.K ? it.k + p.kchunk : p.K;
        const int64_t rem = num_tiles - it.group * T;
        it.t_eff = T;              // padded to T tiles (see role_mma)
        (void)rem;
      }
    };
    auto advance = [&](It& it) {
      if (++it.t < it.t_eff) return;
      it.t = 0;
      if (++it.kb < KB) return;
      it.kb = 0;
      if (++it.k < it.k1) return;
      enter(it, it.v + gridDim.x);
    };
    auto advance_n = [&](It& it, int n) {
      for (int i = 0; i < n && it.valid; ++i) advance(it);
    };
    struct Idx4 { int v[4]; };
    auto load_idx = [&](const It& it) -> Idx4 {
      Idx4 r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r.v[j] = -1;
        if (it.valid) {
          const int64_t row = (it.group * T + it.t) * BLOCK_M + j * 32 + lane;
          if (row < p.n_out) r.v[j] = p.nbr ? __ldg(p.nbr + (int64_t)it.k * p.n_out + row) : (int)row;
        }
      }
      return r;
    };
    // Ownership: W = min(8, sa) producer warps are active and warp w fills the slots n = w, w + W, w + 2W, ...  A warp
    // waits only on the "empty" barriers of its own slots; the parity wait is sound because its previous wait (slot
    // n - W) already required slot n - W - sa >= n - 2 sa to be consumed (W <= sa), so it is never two ring revolutions
    // ahead of the barrier, and it cannot be behind it either (the next completion needs this very slot filled).
    // (Two earlier variants deadlocked: a stride of 8 with a ring of 6 lets a warp run two revolutions ahead; letting
    // every warp observe every slot's barrier fails the other way round — a slow observer falls a revolution BEHIND
    // and then waits for a completion that depends on its own slot.  tools/pipeline_sim.py replays both.)
    const int W = NUM_GATHER_WARPS < p.sa ? NUM_GATHER_WARPS : p.sa;
    It it0, it1, it2;
    enter(it0, blockIdx.x);
    if (warp >= W) it0.valid = false;
    advance_n(it0, warp);
    it1 = it0;
    advance_n(it1, W);
    it2 = it1;
    advance_n(it2, W);
    Idx4 q0 = load_idx(it0), q1 = load_idx(it1);
    const uint8_t* base_hi = reinterpret_cast<const uint8_t*>(p.pl_hi);
    const uint8_t* base_lo = reinterpret_cast<const uint8_t*>(p.pl_lo);
    const int64_t pitch_b = p.pl_pitch * 2;
    int stage = warp;                               // ring position of this warp's current slot (warp < W <= sa)
    uint32_t phase = 0;
    auto observe = [&]() {                          // wait until the slot's stage is free
      mbar_wait(smem_u32(aempty + stage), phase ^ 1);
    };
    auto next_stage = [&]() {                       // this warp's next slot is W slots on (at most one wrap: W <= sa)
      stage += W;
      if (stage >= p.sa) {
        stage -= p.sa;
        phase ^= 1;
      }
    };
    while (it0.valid) {
      const Idx4 q2 = load_idx(it2);                // indices of this warp's slot after next
      observe();
      const uint32_t dst0 = smem_u32(a_smem + (size_t)stage * a_stage_bytes) + (uint32_t)sub * 128u;
      const int64_t coff_b = ((int64_t)it0.kb * KBLK + ch * 8) * 2;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        // tile row r = 4i + sub; its index lives in register (i / 8) of lane (4 (i % 8) + sub)
        const int idx = __shfl_sync(0xffffffffu, q0.v[i >> 3], ((i & 7) << 2) + sub);
        const uint32_t nbytes = idx >= 0 ? 16u : 0u;
        const int64_t boff = (int64_t)(idx >= 0 ? idx : 0) * pitch_b + coff_b;
        // byte offset of (row r, chunk ch) in the swizzled tile: r*128 + ((ch ^ (r & 7)) << 4), r & 7 = 4 (i & 1) + sub
        const uint32_t off = (uint32_t)i * 512u + (((uint32_t)ch ^ (((uint32_t)(i & 1) << 2) + (uint32_t)sub)) << 4);
        cp_async16_zfill(dst0 + off, base_hi + boff, nbytes);
        if (NSPLIT == 3) cp_async16_zfill(dst0 + A_TILE_BYTES + off, base_lo + boff, nbytes);
      }
      cp_async_mbar_arrive_noinc(smem_u32(afull + stage));
      next_stage();
      it0 = it1; it1 = it2;
      advance_n(it2, W);
      q0 = q1; q1 = q2;
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else if (warp == LOAD_WARP) {
    if (lane == 0) role_weight_loader(p, pl);
  } else if (warp == MMA_WARP) {
    role_mma<NSPLIT, true>(p, pl, 0, n_mma);
  } else if (warp >= PL_MMA_WARP0) {
    const int mid = warp - PL_MMA_WARP0 + 1;
    if (mid < n_mma) role_mma<NSPLIT, true>(p, pl, mid, n_mma);
  } else {
    role_epilogue(p, pl, warp - NUM_GATHER_WARPS, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

}  // namespace

extern "C" int64_t pasco_conv_packed_bytes(int32_t K, int32_t Cin, int32_t Cout) {
  return (int64_t)K * Cin * Cout * 4;
}

extern "C" int pasco_conv_pack_weights(const float* W, int32_t K, int32_t Cin, int32_t Cout, int32_t transpose,
                                       void* packed, pasco_stream_t s) {
  const int N = transpose ? Cin : Cout, Kc = transpose ? Cout : Cin;
  PASCO_CHECK_ARG(Kc % KBLK == 0, "pasco_conv_pack_weights: contraction channels (%d) must be a multiple of 64", Kc);
  PASCO_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256, "pasco_conv_pack_weights: output channels (%d) must be a multiple of 16 in [16,256]", N);
  int64_t total = (int64_t)K * (Kc / KBLK) * N * 8;
  k_pack_weights<<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>(W, K, Cin, Cout, transpose, (uint8_t*)packed);
  PASCO_CHECK_LAUNCH("pasco_conv_pack_weights");
  return 0;
}

namespace {

// split-K plan: few output tiles and many offsets (the dense bottleneck: 32 tiles x 245 offsets) leave most SMs idle with
// one CTA per tile; cut the offsets into `ksplit` ranges so that ~2 x #SM work items exist.  1 = no split.
int splitk_plan(int K, int64_t n_out) {
  const int64_t tiles = (n_out + BLOCK_M - 1) / BLOCK_M;
  if (K < 18 || tiles * 2 > num_sms()) return 1;
  int ks = (int)((2 * (int64_t)num_sms() + tiles - 1) / tiles);
  if (ks > K / 6) ks = K / 6;                       // at least 6 offsets per work item
  return ks < 2 ? 1 : ks;
}

// out[r, :] = bias + Σ_s part[s, r, :]   (deterministic reduction of the split-K partial outputs)
__global__ void k_splitk_reduce(const float* __restrict__ part, int ksplit, int64_t n, int C, const float* __restrict__ bias,
                                float* __restrict__ out, int64_t out_pitch) {
  const int cv = C >> 2;
  const int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cv;
    const int c = (int)(t - r * cv) << 2;
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + ((int64_t)s * n + r) * C + c));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * out_pitch + c) = acc;
  }
}

int conv_forward_impl(const float* in, int64_t n_in, const int32_t* nbr, int32_t K, int64_t n_out,
                                     int32_t Cin, int32_t Cout, const void* packed_w, const int32_t* koff_map,
                                     const float* bias, const float* in_scale, const float* in_shift, int32_t in_act,
                                     double* stats, float* out, int32_t precision, int64_t in_pitch, int64_t out_pitch,
                                     int ksplit, float* workspace, pasco_stream_t s) {
  PASCO_CHECK_ARG(precision == 1 || precision == 3, "pasco_conv_forward_tc: precision must be 1 (bf16) or 3 (bf16x3)");
  PASCO_CHECK_ARG(Cin % KBLK == 0, "pasco_conv_forward_tc: Cin (%d) must be a multiple of 64", Cin);
  PASCO_CHECK_ARG(Cout % 16 == 0 && Cout >= 16 && Cout <= 256, "pasco_conv_forward_tc: Cout (%d) must be a multiple of 16 in [16,256]", Cout);
  PASCO_CHECK_ARG(K >= 1 && K <= 1024, "pasco_conv_forward_tc: K (%d) out of range", K);
  PASCO_CHECK_ARG((((uintptr_t)in | (uintptr_t)out | (uintptr_t)packed_w) & 15) == 0, "pasco_conv_forward_tc: pointers must be 16-byte aligned");
  if (n_out == 0) return 0;
  int dev = 0, smem_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int n_op = precision == 3 ? 2 : 1;
  const int a_stage = n_op * A_TILE_BYTES, b_stage = n_op * Cout * 128;
  int64_t tiles = (n_out + BLOCK_M - 1) / BLOCK_M;
  int T = 256 / Cout;
  if (T < 1) T = 1;
  if (T > 4) T = 4;
  while (T > 1 && tiles < (int64_t)T * num_sms()) T >>= 1;   // small inputs: spread tiles over more SMs instead
  if (ksplit > 1) T = 1;
  const int idx_bytes = IDX_RING * T * BLOCK_M * 4;
  const int fixed = 1024 /*align slack*/ + idx_bytes + (4 * MAX_STAGES + 4) * 8 + 16 + (stats ? 2 * Cout * 8 : 0);
  int sb = 2;
  int sa = (smem_optin - fixed - sb * b_stage) / a_stage;
  if (sa > 5) {  // room to spare: deepen the weight ring first
    sb = 3;
    sa = (smem_optin - fixed - sb * b_stage) / a_stage;
  }
  if (sa > MAX_STAGES) sa = MAX_STAGES;
  PASCO_CHECK_ARG(sa >= 2, "pasco_conv_forward_tc: not enough shared memory (Cout=%d)", Cout);
  ConvParams p;
  p.in = in; p.nbr = nbr; p.wpk = (const uint8_t*)packed_w; p.bias = bias;
  p.in_scale = in_scale; p.in_shift = in_shift; p.out = out; p.stats = stats;
  p.n_out = n_out;
  p.out_pitch = out_pitch > 0 ? out_pitch : Cout;
  p.ksplit = 1; p.kchunk = K;
  if (ksplit > 1) {       // partial outputs [ksplit, n_out, Cout] into the workspace, bias added by the reduction
    p.ksplit = ksplit; p.kchunk = (K + ksplit - 1) / ksplit;
    p.ksplit = (K + p.kchunk - 1) / p.kchunk;       // no empty ranges
    p.out = workspace; p.out_pitch = Cout; p.bias = nullptr; p.stats = nullptr;
  }
  p.in_pitch = in_pitch > 0 ? in_pitch : Cin;
  PASCO_CHECK_ARG(p.out_pitch % 4 == 0 && p.in_pitch % 4 == 0, "pasco_conv_forward_tc: pitches must be multiples of 4 floats");
  p.K = K; p.Cin = Cin; p.Cout = Cout; p.in_act = in_act;
  p.pl_hi = nullptr; p.pl_lo = nullptr; p.pl_pitch = 0; p.pl_depth = 0;
  p.sa = sa; p.sb = sb; p.tiles_per_group = T; p.tmem_cols = pow2_cols(2 * T * Cout);
  p.koff_base = 0; p.koff_step = 1;
  if (koff_map) {
    bool ident = true, rev = true;
    for (int k = 0; k < K; ++k) {
      ident = ident && koff_map[k] == k;
      rev = rev && koff_map[k] == K - 1 - k;
    }
    PASCO_CHECK_ARG(ident || rev, "pasco_conv_forward_tc: koff_map must be the identity or the reversal");
    if (!ident) { p.koff_base = K - 1; p.koff_step = -1; }
  }
  const size_t smem = (size_t)sa * a_stage + (size_t)sb * b_stage + fixed;
  int64_t groups = (tiles + T - 1) / T * p.ksplit;
  int grid = (int)(groups < num_sms() ? groups : num_sms());
  cudaError_t e;
  const bool prologue = in_scale != nullptr || in_act != 0;
  auto launch = [&](auto kern) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kern<<<grid, NUM_THREADS, smem, (cudaStream_t)s>>>(p);
    return err;
  };
  if (precision == 3) e = prologue ? launch(k_conv_tc<3, true>) : launch(k_conv_tc<3, false>);
  else e = prologue ? launch(k_conv_tc<1, true>) : launch(k_conv_tc<1, false>);
  if (e != cudaSuccess) {
    set_error("pasco_conv_forward_tc: cudaFuncSetAttribute(%zu bytes) failed: %s", smem, cudaGetErrorString(e));
    return -1;
  }
  PASCO_CHECK_LAUNCH("pasco_conv_forward_tc");
  if (p.ksplit > 1) {
    k_splitk_reduce<<<grid_for(n_out * (Cout / 4), 256), 256, 0, (cudaStream_t)s>>>(workspace, p.ksplit, n_out, Cout, bias, out,
                                                                                out_pitch > 0 ? out_pitch : Cout);
    PASCO_CHECK_LAUNCH("pasco_conv_forward_tc(split-K reduce)");
  }
  return 0;
}

}  // namespace

extern "C" int pasco_conv_forward_tc(const float* in, int64_t n_in, const int32_t* nbr, int32_t K, int64_t n_out,
                                     int32_t Cin, int32_t Cout, const void* packed_w, const int32_t* koff_map,
                                     const float* bias, const float* in_scale, const float* in_shift, int32_t in_act,
                                     double* stats, float* out, int32_t precision, int64_t in_pitch, int64_t out_pitch,
                                     pasco_stream_t s) {
  return conv_forward_impl(in, n_in, nbr, K, n_out, Cin, Cout, packed_w, koff_map, bias, in_scale, in_shift, in_act, stats, out,
                           precision, in_pitch, out_pitch, 1, nullptr, s);
}

extern "C" int pasco_conv_forward_planes(const void* hi, const void* lo, int64_t n_in, const int32_t* nbr, int32_t K,
                                         int64_t n_out, int32_t Cin, int32_t Cout, const void* packed_w,
                                         const int32_t* koff_map, const float* bias, double* stats, float* out,
                                         int32_t precision, int64_t plane_pitch, int64_t out_pitch, pasco_stream_t s) {
  PASCO_CHECK_ARG(precision == 1 || precision == 3, "pasco_conv_forward_planes: precision must be 1 or 3");
  PASCO_CHECK_ARG(hi != nullptr && (precision == 1 || lo != nullptr), "pasco_conv_forward_planes: missing plane (precision 3 needs hi and lo)");
  PASCO_CHECK_ARG(Cin % KBLK == 0, "pasco_conv_forward_planes: Cin (%d) must be a multiple of 64", Cin);
  PASCO_CHECK_ARG(Cout % 16 == 0 && Cout >= 16 && Cout <= 256, "pasco_conv_forward_planes: Cout (%d) must be a multiple of 16 in [16,256]", Cout);
  PASCO_CHECK_ARG(K >= 1 && K <= 1024, "pasco_conv_forward_planes: K (%d) out of range", K);
  const int64_t pitch = plane_pitch > 0 ? plane_pitch : Cin;
  PASCO_CHECK_ARG(pitch % 8 == 0 && (((uintptr_t)hi | (uintptr_t)lo | (uintptr_t)out | (uintptr_t)packed_w) & 15) == 0,
                  "pasco_conv_forward_planes: planes must be 16-byte aligned with a pitch multiple of 8 elements");
  (void)n_in;
  if (n_out == 0) return 0;
  int dev = 0, smem_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int n_op = precision == 3 ? 2 : 1;
  const int a_stage = n_op * A_TILE_BYTES, b_stage = n_op * Cout * 128;
  const int64_t tiles = (n_out + BLOCK_M - 1) / BLOCK_M;
  int T = 256 / Cout;
  if (T < 1) T = 1;
  if (T > 4) T = 4;
  while (T > 1 && tiles < (int64_t)T * num_sms()) T >>= 1;
  const int fixed = 1024 + (4 * MAX_STAGES + 4) * 8 + 16 + (stats ? 2 * Cout * 8 : 0);
  int sb = 2;
  int sa = (smem_optin - fixed - sb * b_stage) / a_stage;
  if (sa > MAX_STAGES) sa = MAX_STAGES;
  // an even ring lets two MMA issuers share the tile groups (n_mma must divide T and sa, see k_conv_pl): at C = 128 in
  // fp32 mode 4 stages + 2 issuers beat 5 stages + 1 issuer
  if (T >= 2 && (sa & 1) && sa > 2) sa -= 1;
  static const int sa_env = [] { const char* e = getenv("PASCO_PL_SA"); return e ? atoi(e) : 0; }();
  if (sa_env >= 2 && sa_env < sa) sa = sa_env;
  PASCO_CHECK_ARG(sa >= 2, "pasco_conv_forward_planes: not enough shared memory (Cout=%d)", Cout);
  ConvParams p;
  p.in = nullptr; p.nbr = nbr; p.wpk = (const uint8_t*)packed_w; p.bias = bias;
  p.in_scale = nullptr; p.in_shift = nullptr; p.out = out; p.stats = stats;
  p.n_out = n_out;
  p.out_pitch = out_pitch > 0 ? out_pitch : Cout;
  p.in_pitch = Cin;
  PASCO_CHECK_ARG(p.out_pitch % 4 == 0, "pasco_conv_forward_planes: out_pitch must be a multiple of 4 floats");
  p.K = K; p.Cin = Cin; p.Cout = Cout; p.in_act = 0;
  p.sa = sa; p.sb = sb; p.tiles_per_group = T; p.tmem_cols = pow2_cols(2 * T * Cout);
  p.ksplit = 1; p.kchunk = K;
  p.koff_base = 0; p.koff_step = 1;
  if (koff_map) {
    bool ident = true, rev = true;
    for (int k = 0; k < K; ++k) {
      ident = ident && koff_map[k] == k;
      rev = rev && koff_map[k] == K - 1 - k;
    }
    PASCO_CHECK_ARG(ident || rev, "pasco_conv_forward_planes: koff_map must be the identity or the reversal");
    if (!ident) { p.koff_base = K - 1; p.koff_step = -1; }
  }
  p.pl_hi = (const uint16_t*)hi; p.pl_lo = (const uint16_t*)lo; p.pl_pitch = pitch;
  static const int depth_env = [] { const char* e = getenv("PASCO_PL_DEPTH"); return e ? atoi(e) : 0; }();
  p.pl_depth = PL_MMA_WARPS;                 // MMA issuer warps (PASCO_PL_DEPTH=1..4 overrides: debugging)
  if (depth_env >= 1 && depth_env <= PL_MMA_WARPS) p.pl_depth = depth_env;
  const size_t smem = (size_t)sa * a_stage + (size_t)sb * b_stage + fixed;
  const int64_t groups = (tiles + T - 1) / T;
  const int grid = (int)(groups < num_sms() ? groups : num_sms());
  auto launch = [&](auto kern) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kern<<<grid, NUM_THREADS_PL, smem, (cudaStream_t)s>>>(p);
    return err;
  };
  const cudaError_t e = precision == 3 ? launch(k_conv_pl<3>) : launch(k_conv_pl<1>);
  if (e != cudaSuccess) {
    set_error("pasco_conv_forward_planes: cudaFuncSetAttribute(%zu bytes) failed: %s", smem, cudaGetErrorString(e));
    return -1;
  }
  PASCO_CHECK_LAUNCH("pasco_conv_forward_planes");
  return 0;
}

extern "C" int64_t pasco_conv_splitk_workspace_bytes(int32_t K, int64_t n_out, int32_t Cout) {
  const int ks = splitk_plan(K, n_out);
  return ks > 1 ? (int64_t)ks * n_out * Cout * 4 : 0;
}

extern "C" int pasco_conv_forward_splitk(const float* in, int64_t n_in, const int32_t* nbr, int32_t K, int64_t n_out,
                                         int32_t Cin, int32_t Cout, const void* packed_w, const int32_t* koff_map,
                                         const float* bias, const float* in_scale, const float* in_shift, int32_t in_act,
                                         float* out, int32_t precision, int64_t in_pitch, int64_t out_pitch, void* workspace,
                                         int64_t workspace_bytes, pasco_stream_t s) {
  const int ks = splitk_plan(K, n_out);
  PASCO_CHECK_ARG(ks > 1, "pasco_conv_forward_splitk: this shape is not split (workspace_bytes = 0): call pasco_conv_forward_tc");
  PASCO_CHECK_ARG(workspace != nullptr && workspace_bytes >= (int64_t)ks * n_out * Cout * 4,
                  "pasco_conv_forward_splitk: workspace too small (%lld bytes)", (long long)workspace_bytes);
  PASCO_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "pasco_conv_forward_splitk: workspace must be 16-byte aligned");
  return conv_forward_impl(in, n_in, nbr, K, n_out, Cin, Cout, packed_w, koff_map, bias, in_scale, in_shift, in_act, nullptr, out,
                           precision, in_pitch, out_pitch, ks, (float*)workspace, s);
}
