// Sparse convolution on tcgen05 with PRE-SPLIT bf16 input planes gathered by the TMA engine (sm_100a).
//
// Same output-stationary implicit GEMM, tile groups, weight-slice ring, TMEM double buffering, MMA issue and epilogue as
// conv_tc.cu; what changes is the producer: activations are split x = hi + lo (bf16 each) ONCE per tensor by
// k_split_planes (optionally fused with the BatchNorm affine + activation that produced them) instead of 27x inside
// the gather, and a 128-row A tile is then fetched by 32 `cp.async.bulk.tensor.2d.tile::gather4` instructions per plane:
// each names 4 arbitrary row indices of the [N_in, Cin] plane (tensor-map box 64 x 1, SWIZZLE_128B) and the TMA unit
// writes the four 128-byte rows straight into the swizzled UMMA tile.  No LSU instructions, registers or conversions on
// the gather path.  Missing neighbours (-1) are redirected to a pad of PLANE_PAD all-zero rows behind the plane: the
// TMA's own out-of-range zero fill costs ~9 ns per row per SM (4.5x slower than fetching), and ONE shared zero row
// hot-spots an L2 slice (0.6 TB/s), while zero rows spread over the pad run at the full 10 TB/s (all measured).
// Measured (tools/tma_gather_probe.cu): one warp sustains one gather4 per ~46 ns whatever the ring depth, and issuing
// warps scale linearly (6 warps: 34 B/clk/SM, 9.6 TB/s over the GPU), so the 8 producer warps each own one
// (stage, plane) in turn and 4 stages are in flight.
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

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
constexpr int MAX_STAGES = 8;
constexpr int PLANE_PAD = 1024;   // == PASCO_PLANE_PAD_ROWS: zero rows behind every plane (power of two)

struct PlaneParams {
  const __nv_bfloat16* hi;   // [N_in, Cin]
  const __nv_bfloat16* lo;   // [N_in, Cin] (precision 3 only)
  const int32_t* nbr;
  const uint8_t* wpk;
  const float* bias;
  float* out;
  int64_t n_out;
  int64_t out_pitch;
  int zero_row;              // first row of the zero pad (= n_in)
  int K, Cin, Cout;
  int sa, sb, tiles_per_group, tmem_cols;
  int koff_base, koff_step;
};

// [n_rows, C] bf16 plane, box = 64 channels x 1 row, 128-byte swizzle (= the UMMA K-major tile layout)
bool make_plane_map(CUtensorMap* tm, const void* base, int64_t n_rows, int C) {
  return tma::make_row_gather_map(tm, base, n_rows, C, (int64_t)C * 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, KBLK,
                                  CU_TENSOR_MAP_SWIZZLE_128B);
}

// fp32 [N, C] (row pitch `pitch`) → bf16 planes hi (and lo = bf16(x − hi)); optional y = act(x*scale + shift) first
__global__ void k_split_planes(const float* __restrict__ x, int64_t n, int C, int64_t pitch, const float* __restrict__ scale,
                               const float* __restrict__ shift, int act, __nv_bfloat16* __restrict__ hi,
                               __nv_bfloat16* __restrict__ lo) {
  const int cv = C >> 2;
  const int64_t total = (n + PLANE_PAD) * cv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cv;
    const int c = (int)(t - r * cv) << 2;
    if (r >= n) {                                  // the zero pad behind the plane
      *reinterpret_cast<uint2*>(hi + r * C + c) = make_uint2(0u, 0u);
      if (lo) *reinterpret_cast<uint2*>(lo + r * C + c) = make_uint2(0u, 0u);
      continue;
    }
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
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_planes(const __grid_constant__ PlaneParams p, const __grid_constant__ CUtensorMap tm_hi,
                                                                const __grid_constant__ CUtensorMap tm_lo) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin / KBLK;
  const int64_t num_tiles = (p.n_out + BLOCK_M - 1) / BLOCK_M;
  const int64_t num_groups = (num_tiles + T - 1) / T;
  const int acc_cols = T * p.Cout;                // TMEM columns of one accumulator set

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(smem_u32(afull + s), n_op);   // one arrive.expect_tx per plane
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
    // ===================================== gather producers (TMA gather4) =====================================
    // Stage n of the ring is the A tile of (group, k, kb, t) in MMA order.  Warp w owns plane (w % n_op) of the stages
    // n ≡ w / n_op (mod NUM_GATHER_WARPS / n_op); lane l fetches tile rows 4l..4l+3 with one gather4.
    const int my_pl = (n_op == 2) ? (warp & 1) : 0;
    const int my_first = (n_op == 2) ? (warp >> 1) : warp;
    // a waiter may be at most one mbarrier phase ahead: the stride between a warp's stages must not exceed the ring
    const int MOD = (NUM_GATHER_WARPS / n_op) < p.sa ? (NUM_GATHER_WARPS / n_op) : p.sa;
    const CUtensorMap* tm = my_pl ? &tm_lo : &tm_hi;
    // cursor over stages (scalars; advanced MOD stages at a time)
    int64_t c_group = blockIdx.x;
    int c_k = 0, c_kb = 0, c_t = 0;
    int64_t rem0 = num_tiles - c_group * T;
    int c_teff = rem0 < T ? (int)rem0 : T;
    bool c_valid = c_group < num_groups && my_first < MOD;
    auto advance = [&](int steps) {
      for (int s_ = 0; s_ < steps && c_valid; ++s_) {
        if (++c_t == c_teff) {
          c_t = 0;
          if (++c_kb == KB) {
            c_kb = 0;
            if (++c_k == p.K) {
              c_k = 0;
              c_group += gridDim.x;
              c_valid = c_group < num_groups;
              if (c_valid) {
                const int64_t rem = num_tiles - c_group * T;
                c_teff = rem < T ? (int)rem : T;
              }
            }
          }
        }
      }
    };
    auto load_idx = [&]() -> int4 {
      const int64_t row = (c_group * T + c_t) * BLOCK_M + 4 * lane;
      const int z = p.zero_row + ((4 * lane + 128 * (c_k & 7)) & (PLANE_PAD - 1));   // my 4 zero rows of the pad
      int4 v;
      if (p.nbr) {
        const int32_t* src = p.nbr + (int64_t)c_k * p.n_out + row;
        v.x = row + 0 < p.n_out ? __ldg(src + 0) : -1;
        v.y = row + 1 < p.n_out ? __ldg(src + 1) : -1;
        v.z = row + 2 < p.n_out ? __ldg(src + 2) : -1;
        v.w = row + 3 < p.n_out ? __ldg(src + 3) : -1;
      } else {
        v.x = row + 0 < p.n_out ? (int)row + 0 : -1;
        v.y = row + 1 < p.n_out ? (int)row + 1 : -1;
        v.z = row + 2 < p.n_out ? (int)row + 2 : -1;
        v.w = row + 3 < p.n_out ? (int)row + 3 : -1;
      }
      v.x = v.x < 0 ? z + 0 : v.x;
      v.y = v.y < 0 ? z + 1 : v.y;
      v.z = v.z < 0 ? z + 2 : v.z;
      v.w = v.w < 0 ? z + 3 : v.w;
      return v;
    };
    advance(my_first);
    int slot = my_first % p.sa;
    uint32_t phase = (uint32_t)(my_first / p.sa) & 1u;
    int4 idx = make_int4(0, 0, 0, 0);
    if (c_valid) idx = load_idx();
    while (c_valid) {
      const int col = c_kb * KBLK;
      const uint32_t dst = smem_u32(a_smem + (size_t)slot * a_stage_bytes + (size_t)my_pl * A_TILE_BYTES) + (uint32_t)lane * 512u;
      const uint32_t bar = smem_u32(afull + slot);
      const uint32_t ebar = smem_u32(aempty + slot);
      const uint32_t eph = phase ^ 1u;
      const int4 cur = idx;
      advance(MOD);                                  // index prefetch of my next stage overlaps the wait + issue below
      if (c_valid) idx = load_idx();
      mbar_wait(ebar, eph);
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)A_TILE_BYTES);
      __syncwarp();
      tma::gather4(dst, tm, col, cur.x, cur.y, cur.z, cur.w, bar);
      slot += MOD;
      while (slot >= p.sa) {
        slot -= p.sa;
        phase ^= 1u;
      }
    }
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
  k_split_planes<<<grid_for((n + PLANE_PAD) * (C / 4), 256), 256, 0, (cudaStream_t)s>>>(x, n, C, pitch > 0 ? pitch : C, scale, shift, act,
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
  if (n_out == 0) return 0;
  CUtensorMap tm_hi, tm_lo;
  if (!make_plane_map(&tm_hi, hi, n_in + PLANE_PAD, Cin) || !make_plane_map(&tm_lo, lo ? lo : hi, n_in + PLANE_PAD, Cin)) {
    set_error("pasco_conv_forward_planes: cuTensorMapEncodeTiled failed (n_in=%lld, Cin=%d)", (long long)n_in, Cin);
    return -1;
  }
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
  const int fixed = 1024 + (4 * MAX_STAGES + 4) * 8 + 16;
  int sb = 2;
  int sa = (smem_optin - fixed - sb * b_stage) / a_stage;
  if (sa > MAX_STAGES) sa = MAX_STAGES;
  PASCO_CHECK_ARG(sa >= 2, "pasco_conv_forward_planes: not enough shared memory (Cout=%d)", Cout);
  PlaneParams p;
  p.hi = (const __nv_bfloat16*)hi; p.lo = (const __nv_bfloat16*)lo; p.nbr = nbr; p.wpk = (const uint8_t*)packed_w;
  p.bias = bias; p.out = out; p.n_out = n_out; p.zero_row = (int)n_in; p.out_pitch = out_pitch > 0 ? out_pitch : Cout;
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
    if (err == cudaSuccess) kern<<<grid, NUM_THREADS, smem, (cudaStream_t)s>>>(p, tm_hi, tm_lo);
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
