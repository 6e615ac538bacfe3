// Sparse convolution on tcgen05 with PRE-SPLIT bf16 input planes and an asynchronous-copy gather (sm_100a).
//
// Same output-stationary implicit GEMM, tile groups, weight-slice ring, TMEM double buffering, MMA issue and epilogue as
// conv_tc.cu; what changes is the producer: activations are split x = hi + lo (bf16 each) ONCE per tensor by
// k_split_planes (optionally fused with the BatchNorm affine + activation that produced them) instead of 27x inside
// the gather, and a tile row becomes a pure 128-byte copy per plane: 16-byte cp.async (LDGSTS) straight into the
// 128-byte-swizzled UMMA tile with zero fill for missing neighbours.  The gather warps issue ~1/3 of the instructions
// of the register path and keep DEPTH slots in flight each.
#include "common.cuh"
#include "umma.cuh"

using namespace pasco;
using namespace umma;

namespace {

constexpr int BLOCK_M = 128;
constexpr int KBLK = 64;
constexpr int A_TILE_BYTES = BLOCK_M * 128;
constexpr int NUM_GATHER_WARPS = 8;
constexpr int NUM_EPI_WARPS = 4;
constexpr int MMA_WARP = NUM_GATHER_WARPS + NUM_EPI_WARPS;
constexpr int LOAD_WARP = MMA_WARP + 1;
constexpr int NUM_THREADS = (LOAD_WARP + 1) * 32;
constexpr int ROWS_PER_WARP = BLOCK_M / NUM_GATHER_WARPS;
constexpr int MAX_STAGES = 8;
constexpr int IDX_RING = 8;

struct PlaneParams {
  const __nv_bfloat16* hi;   // [N_in, Cin]
  const __nv_bfloat16* lo;   // [N_in, Cin] (precision 3 only)
  const int32_t* nbr;
  const uint8_t* wpk;
  const float* bias;
  float* out;
  int64_t n_out;
  int64_t out_pitch;
  int K, Cin, Cout;
  int sa, sb, tiles_per_group, tmem_cols;
  int koff_base, koff_step;
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// fp32 [N, C] (row pitch `pitch`) → bf16 planes hi (and lo = bf16(x − hi)); optional y = act(x*scale + shift) first
__global__ void k_split_planes(const float* __restrict__ x, int64_t n, int C, int64_t pitch, const float* __restrict__ scale,
                               const float* __restrict__ shift, int act, __nv_bfloat16* __restrict__ hi,
                               __nv_bfloat16* __restrict__ lo) {
  const int cv = C >> 2;
  const int64_t total = n * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cv;
    const int c = (int)(t - r * cv) << 2;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + r * pitch + c));
    if (scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c)), sh = __ldg(reinterpret_cast<const float4*>(shift + c));
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (act == 1) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if (act == 2) {
      v.x = v.x > 0.f ? v.x : 0.01f * v.x; v.y = v.y > 0.f ? v.y : 0.01f * v.y;
      v.z = v.z > 0.f ? v.z : 0.01f * v.z; v.w = v.w > 0.f ? v.w : 0.01f * v.w;
    }
    uint2 h, l;
    if (lo) {
      split4(v, h, l);
      *reinterpret_cast<uint2*>(lo + r * C + c) = l;
    } else {
      h = to_bf16x4(v);
    }
    *reinterpret_cast<uint2*>(hi + r * C + c) = h;
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_planes(const __grid_constant__ PlaneParams p) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin / KBLK;
  const int64_t num_tiles = (p.n_out + BLOCK_M - 1) / BLOCK_M;
  const int64_t num_groups = (num_tiles + T - 1) / T;
  const int acc_cols = T * p.Cout;                // TMEM columns of one accumulator set

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

  if (warp < NUM_GATHER_WARPS) {
    // ===================================== gather producers (async copies) =====================================
    // The input arrives pre-split into bf16 planes (hi, lo) [N_in, Cin], so a tile row is a pure 128-byte copy per
    // plane and 64-channel block: 16-byte cp.async (LDGSTS) straight into the swizzled UMMA tile, zero-filled
    // (src-size 0) where the neighbour is missing.  No registers, no conversion: DEPTH slots are in flight per warp.
    constexpr int DEPTH = 3;                                   // slots in flight per warp (bounded by the ring: sa - 1)
    const int depth = p.sa - 1 < DEPTH ? p.sa - 1 : DEPTH;
    const int crow = lane >> 3;      // 4 rows per instruction
    const int cchunk = lane & 7;     // 16-byte chunk of the 128-byte row segment
    const int PE = (DEPTH + 1 + T * KB - 1) / (T * KB) + 1;   // index-ring prefetch distance in (group,k) entries

    auto prefetch_idx = [&](int64_t group, int k, int gk, int t_eff) {   // joins the current cp.async group (no commit)
      int* dst = idx_ring + (gk % IDX_RING) * T * BLOCK_M;
      for (int t = lane >> 4; t < t_eff; t += 2) {
        const int64_t row = (group * T + t) * BLOCK_M + warp * ROWS_PER_WARP + (lane & 15);
        int* d = dst + t * BLOCK_M + warp * ROWS_PER_WARP + (lane & 15);
        if (row < p.n_out) {
          if (p.nbr) cp_async4(smem_u32(d), p.nbr + (int64_t)k * p.n_out + row);
          else *d = (int)row;
        } else {
          *d = -1;
        }
      }
    };
    // cursor over (group, k, kb, t)
    int64_t i_group = blockIdx.x;
    int i_k = 0, i_kb = 0, i_t = 0, i_gk = 0;
    int64_t rem0 = num_tiles - i_group * T;
    int i_teff = rem0 < T ? (int)rem0 : T;
    bool i_valid = i_group < num_groups;
    // prefetch cursor over (group, k), PE entries ahead of the issue cursor
    int64_t f_group = blockIdx.x;
    int f_k = 0, f_gk = 0;
    auto prefetch_next_entry = [&]() {
      if (f_group < num_groups) {
        int64_t rem = num_tiles - f_group * T;
        prefetch_idx(f_group, f_k, f_gk, rem < T ? (int)rem : T);
      }
      ++f_gk;
      if (++f_k == p.K) {
        f_k = 0;
        f_group += gridDim.x;
      }
    };
    int st_i = 0, st_p = 0;          // issue / publish stage
    uint32_t ph_i = 0;
    int in_flight = 0;
    auto publish = [&]() {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(afull + st_p));
      if (++st_p == p.sa) st_p = 0;
      --in_flight;
    };
    if (i_valid) {
      for (int a = 0; a < PE; ++a) prefetch_next_entry();
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      while (i_valid) {
        if (i_kb == 0 && i_t == 0) prefetch_next_entry();            // keep the ring PE entries ahead
        mbar_wait(smem_u32(aempty + st_i), ph_i ^ 1);
        const int* irow = idx_ring + (i_gk % IDX_RING) * T * BLOCK_M + i_t * BLOCK_M + warp * ROWS_PER_WARP;
        const uint32_t dst0 = smem_u32(a_smem + (size_t)st_i * a_stage_bytes);
        const int64_t coff = (int64_t)i_kb * KBLK + cchunk * 8;        // bf16 elements
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 4 + crow;                                  // row within this warp's 16
          const int src = irow[r];
          const uint32_t trow = (uint32_t)(warp * ROWS_PER_WARP + r);
          const uint32_t off = trow * 128u + (((uint32_t)cchunk ^ (trow & 7u)) << 4);
          const int64_t e = (int64_t)(src >= 0 ? src : 0) * p.Cin + coff;
          const uint32_t nbytes = src >= 0 ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + off), "l"(p.hi + e), "r"(nbytes) : "memory");
          if (NSPLIT == 3)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + A_TILE_BYTES + off), "l"(p.lo + e), "r"(nbytes) : "memory");
        }
        cp_async_commit();
        ++in_flight;
        if (++st_i == p.sa) {
          st_i = 0;
          ph_i ^= 1;
        }
        // advance the issue cursor
        if (++i_t == i_teff) {
          i_t = 0;
          if (++i_kb == KB) {
            i_kb = 0;
            ++i_gk;
            if (++i_k == p.K) {
              i_k = 0;
              i_group += gridDim.x;
              i_valid = i_group < num_groups;
              if (i_valid) {
                int64_t rem = num_tiles - i_group * T;
                i_teff = rem < T ? (int)rem : T;
              }
            }
          }
        }
        if (in_flight > depth) {
          if (depth >= 3) cp_async_wait<3>();
          else if (depth == 2) cp_async_wait<2>();
          else if (depth == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
          publish();
        }
      }
      // drain
      if (in_flight >= 3) { cp_async_wait<2>(); publish(); }
      if (in_flight >= 2) { cp_async_wait<1>(); publish(); }
      if (in_flight >= 1) { cp_async_wait<0>(); publish(); }
    }
    cp_async_wait<0>();
  } else if (warp == LOAD_WARP) {
    // ===================================== weight-slice loader =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)b_stage_bytes;
      for (int64_t group = blockIdx.x; group < num_groups; group += gridDim.x) {
        for (int k = 0; k < p.K; ++k) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(smem_u32(bempty + stage), phase ^ 1);
            mbar_arrive_expect_tx(smem_u32(bfull + stage), bytes);
            const uint8_t* src = p.wpk + ((int64_t)(p.koff_base + p.koff_step * k) * KB + kb) * (int64_t)p.Cout * 256;
            if (NSPLIT == 3) {
              bulk_g2s(smem_u32(b_smem + (size_t)stage * b_stage_bytes), src, bytes, smem_u32(bfull + stage));
            } else {
              bulk_g2s(smem_u32(b_smem + (size_t)stage * b_stage_bytes), src, bytes, smem_u32(bfull + stage));
            }
            if (++stage == p.sb) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.Cout, 0, 0);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int64_t group = blockIdx.x; group < num_groups; group += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t rem = num_tiles - group * T;
        const int t_eff = rem < T ? (int)rem : T;
        mbar_wait(smem_u32(tempty + buf), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int k = 0; k < p.K; ++k) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(smem_u32(bfull + sb), pb);
            const uint32_t b_hi = smem_u32(b_smem + (size_t)sb * b_stage_bytes), b_lo = b_hi + b_tile;
            for (int t = 0; t < t_eff; ++t) {
              mbar_wait(smem_u32(afull + sa), pa);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(a_smem + (size_t)sa * a_stage_bytes), a_lo = a_hi + A_TILE_BYTES;
              const uint32_t d_tmem = tmem_base + (uint32_t)(buf * acc_cols + t * p.Cout);
#pragma unroll
              for (int j = 0; j < KBLK / 16; ++j) {
                const uint32_t accum = (k > 0 || kb > 0 || j > 0) ? 1u : 0u;
                const uint64_t da_hi = make_desc_sw128(a_hi + j * 32, 16, 1024);
                const uint64_t db_hi = make_desc_sw128(b_hi + j * 32, 16, 1024);
                mma_bf16(d_tmem, da_hi, db_hi, idesc, accum);
                if (NSPLIT == 3) {
                  const uint64_t da_lo = make_desc_sw128(a_lo + j * 32, 16, 1024);
                  const uint64_t db_lo = make_desc_sw128(b_lo + j * 32, 16, 1024);
                  mma_bf16(d_tmem, da_lo, db_hi, idesc, 1);
                  mma_bf16(d_tmem, da_hi, db_lo, idesc, 1);
                }
              }
              mma_commit(smem_u32(aempty + sa));
              if (++sa == p.sa) {
                sa = 0;
                pa ^= 1;
              }
            }
            mma_commit(smem_u32(bempty + sb));
            if (++sb == p.sb) {
              sb = 0;
              pb ^= 1;
            }
          }
        }
        mma_commit(smem_u32(tfull + buf));
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp - NUM_GATHER_WARPS;  // == warp % 4: TMEM lane quadrant this warp may read
    int it = 0;
    for (int64_t group = blockIdx.x; group < num_groups; group += gridDim.x, ++it) {
      const int buf = it & 1;
      const int64_t rem = num_tiles - group * T;
      const int t_eff = rem < T ? (int)rem : T;
      mbar_wait(smem_u32(tfull + buf), (it >> 1) & 1);
      tc_fence_after();
      for (int t = 0; t < t_eff; ++t) {
        const int64_t row = (group * T + t) * BLOCK_M + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + t * p.Cout);
        float* orow = p.out + row * p.out_pitch;
        int c0 = 0;
        for (; c0 + 32 <= p.Cout; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (row < p.n_out) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.bias) {
                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              *reinterpret_cast<float4*>(orow + c0 + j) = o;
            }
          }
        }
        if (c0 < p.Cout) {  // 16-column tail (Cout % 32 == 16)
          float v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (row < p.n_out) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.bias) {
                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              *reinterpret_cast<float4*>(orow + c0 + j) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty + buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int pow2_cols_p(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

}  // namespace

extern "C" int pasco_split_planes(const float* x, int64_t n, int32_t C, int64_t pitch, const float* scale, const float* shift,
                                  int32_t act, void* hi, void* lo, pasco_stream_t s) {
  PASCO_CHECK_ARG(C % 4 == 0 && (pitch == 0 || pitch % 4 == 0), "pasco_split_planes: C and pitch must be multiples of 4");
  if (n == 0) return 0;
  k_split_planes<<<grid_for(n * (C / 4), 256), 256, 0, (cudaStream_t)s>>>(x, n, C, pitch > 0 ? pitch : C, scale, shift, act,
                                                                      (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  PASCO_CHECK_LAUNCH("pasco_split_planes");
  return 0;
}

extern "C" int pasco_conv_forward_planes(const void* hi, const void* lo, int64_t n_in, const int32_t* nbr, int32_t K,
                                         int64_t n_out, int32_t Cin, int32_t Cout, const void* packed_w,
                                         const int32_t* koff_map, const float* bias, float* out, int32_t precision,
                                         int64_t out_pitch, pasco_stream_t s) {
  PASCO_CHECK_ARG(precision == 1 || precision == 3, "pasco_conv_forward_planes: precision must be 1 or 3");
  PASCO_CHECK_ARG(precision == 1 || lo != nullptr, "pasco_conv_forward_planes: precision 3 needs the lo plane");
  PASCO_CHECK_ARG(Cin % KBLK == 0, "pasco_conv_forward_planes: Cin (%d) must be a multiple of 64", Cin);
  PASCO_CHECK_ARG(Cout % 16 == 0 && Cout >= 16 && Cout <= 256, "pasco_conv_forward_planes: Cout (%d) must be a multiple of 16 in [16,256]", Cout);
  PASCO_CHECK_ARG(K >= 1 && K <= 1024, "pasco_conv_forward_planes: K (%d) out of range", K);
  (void)n_in;
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
  while (T > 1 && tiles < (int64_t)T * num_sms()) T >>= 1;
  const int idx_bytes = IDX_RING * T * BLOCK_M * 4;
  const int fixed = 1024 + idx_bytes + (4 * MAX_STAGES + 4) * 8 + 16;
  int sb = 2;
  int sa = (smem_optin - fixed - sb * b_stage) / a_stage;
  if (sa > MAX_STAGES) sa = MAX_STAGES;
  PASCO_CHECK_ARG(sa >= 2, "pasco_conv_forward_planes: not enough shared memory (Cout=%d)", Cout);
  PlaneParams p;
  p.hi = (const __nv_bfloat16*)hi; p.lo = (const __nv_bfloat16*)lo; p.nbr = nbr; p.wpk = (const uint8_t*)packed_w;
  p.bias = bias; p.out = out; p.n_out = n_out; p.out_pitch = out_pitch > 0 ? out_pitch : Cout;
  p.K = K; p.Cin = Cin; p.Cout = Cout;
  p.sa = sa; p.sb = sb; p.tiles_per_group = T; p.tmem_cols = pow2_cols_p(2 * T * Cout);
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
  const size_t smem = (size_t)sa * a_stage + (size_t)sb * b_stage + fixed;
  int64_t groups = (tiles + T - 1) / T;
  int grid = (int)(groups < num_sms() ? groups : num_sms());
  auto launch = [&](auto kern) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kern<<<grid, NUM_THREADS, smem, (cudaStream_t)s>>>(p);
    return err;
  };
  cudaError_t e = precision == 3 ? launch(k_conv_planes<3>) : launch(k_conv_planes<1>);
  if (e != cudaSuccess) {
    set_error("pasco_conv_forward_planes: cudaFuncSetAttribute(%zu bytes) failed: %s", smem, cudaGetErrorString(e));
    return -1;
  }
  PASCO_CHECK_LAUNCH("pasco_conv_forward_planes");
  return 0;
}
