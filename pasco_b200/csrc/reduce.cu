// Column statistics / fused affine+activation (BatchNorm over rows), max-pool, scatter-max.
// All HBM-bound: one coalesced pass over [N, C] float32 per call.
#include <math_constants.h>
#include "common.cuh"

using namespace pasco;

// ------------------------------------------------------------------------------------------------
// activation helpers: 0 none, 1 ReLU, 2 LeakyReLU(0.01)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return fmaxf(z, 0.f);
  if (act == 2) return z > 0.f ? z : 0.01f * z;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return z > 0.f ? 1.f : 0.01f;
  return 1.f;
}

// ------------------------------------------------------------------------------------------------
// column sums: block = 32 (channels) x 8 (row lanes); each warp reads 128 contiguous bytes of a row
// ------------------------------------------------------------------------------------------------
constexpr int kRowLanes = 8;
constexpr int kRowsPerBlock = 512;

// MODE 0: (Σx, Σx²)   MODE 1: (Σdz, Σdz·x) with dz = dy·act'(x·scale+shift)
template <int MODE>
__global__ void k_col_sums(const float* __restrict__ x, const float* __restrict__ dy, int64_t n, int C,
                           const float* __restrict__ scale, const float* __restrict__ shift, int act,
                           double* __restrict__ sums) {
  __shared__ float red[2][kRowLanes][32];
  int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock;
  int64_t row1 = row0 + kRowsPerBlock < n ? row0 + kRowsPerBlock : n;
  for (int c0 = 0; c0 < C; c0 += 32) {
    int c = c0 + threadIdx.x;
    float s0 = 0.f, s1 = 0.f;
    if (c < C) {
      float sc = 1.f, sh = 0.f;
      if (MODE == 1 && scale) {
        sc = __ldg(scale + c);
        sh = __ldg(shift + c);
      }
      for (int64_t r = row0 + threadIdx.y; r < row1; r += kRowLanes) {
        float v = __ldg(x + r * C + c);
        if (MODE == 0) {
          s0 += v;
          s1 += v * v;
        } else {
          float dz = __ldg(dy + r * C + c) * act_grad(v * sc + sh, act);
          s0 += dz;
          s1 += dz * v;
        }
      }
    }
    red[0][threadIdx.y][threadIdx.x] = s0;
    red[1][threadIdx.y][threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.y < 2 && c < C) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < kRowLanes; ++l) acc += (double)red[threadIdx.y][l][threadIdx.x];
      atomicAdd(sums + (int64_t)threadIdx.y * C + c, acc);
    }
    __syncthreads();
  }
}

extern "C" int pasco_bn_stats(const float* x, int64_t n, int32_t C, double* stats, pasco_stream_t s) {
  if (n == 0 || C == 0) return 0;
  int64_t nb = (n + kRowsPerBlock - 1) / kRowsPerBlock;
  k_col_sums<0><<<(unsigned)nb, dim3(32, kRowLanes), 0, (cudaStream_t)s>>>(x, nullptr, n, C, nullptr, nullptr, 0, stats);
  PASCO_CHECK_LAUNCH("pasco_bn_stats");
  return 0;
}

extern "C" int pasco_bn_bwd_reduce(const float* dy, const float* x, int64_t n, int32_t C, const float* scale,
                                   const float* shift, int32_t act, double* sums, pasco_stream_t s) {
  if (n == 0 || C == 0) return 0;
  int64_t nb = (n + kRowsPerBlock - 1) / kRowsPerBlock;
  k_col_sums<1><<<(unsigned)nb, dim3(32, kRowLanes), 0, (cudaStream_t)s>>>(x, dy, n, C, scale, shift, act, sums);
  PASCO_CHECK_LAUNCH("pasco_bn_bwd_reduce");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// y = act(x*scale[c] + shift[c] + residual)
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void k_affine_act(const float* __restrict__ x, int64_t n, int C, const float* __restrict__ scale,
                             const float* __restrict__ shift, int act, const float* __restrict__ residual,
                             float* __restrict__ y) {
  if (VEC) {
    int cv = C >> 2;
    int64_t total = n * cv;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      int c = (int)(t % cv) << 2;
      float4 v = __ldg((const float4*)x + t);
      float4 sc = scale ? __ldg((const float4*)(scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
      float4 sh = shift ? __ldg((const float4*)(shift + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r = residual ? __ldg((const float4*)residual + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = act_fwd(fmaf(v.x, sc.x, sh.x) + r.x, act);
      o.y = act_fwd(fmaf(v.y, sc.y, sh.y) + r.y, act);
      o.z = act_fwd(fmaf(v.z, sc.z, sh.z) + r.z, act);
      o.w = act_fwd(fmaf(v.w, sc.w, sh.w) + r.w, act);
      ((float4*)y)[t] = o;
    }
  } else {
    int64_t total = n * C;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      int c = (int)(t % C);
      float z = fmaf(__ldg(x + t), scale ? __ldg(scale + c) : 1.f, shift ? __ldg(shift + c) : 0.f);
      if (residual) z += __ldg(residual + t);
      y[t] = act_fwd(z, act);
    }
  }
}

extern "C" int pasco_affine_act(const float* x, int64_t n, int32_t C, const float* scale, const float* shift,
                                int32_t act, const float* residual, float* y, pasco_stream_t s) {
  if (n == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)s;
  bool vec = (C % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual | (uintptr_t)scale | (uintptr_t)shift) % 16) == 0);
  if (vec)
    k_affine_act<true><<<grid_for(n * (C / 4), 256), 256, 0, st>>>(x, n, C, scale, shift, act, residual, y);
  else
    k_affine_act<false><<<grid_for(n * C, 256), 256, 0, st>>>(x, n, C, scale, shift, act, residual, y);
  PASCO_CHECK_LAUNCH("pasco_affine_act");
  return 0;
}

// dx = a[c]*dz + b[c]*x + c0[c],  dz = dy * act'(x*scale+shift)
__global__ void k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ x, int64_t n, int C,
                               const float* __restrict__ scale, const float* __restrict__ shift, int act,
                               const float* __restrict__ ca, const float* __restrict__ cb,
                               const float* __restrict__ cc, float* __restrict__ dx) {
  int64_t total = n * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % C);
    float v = __ldg(x + t);
    float z = fmaf(v, scale ? __ldg(scale + c) : 1.f, shift ? __ldg(shift + c) : 0.f);
    float dz = __ldg(dy + t) * act_grad(z, act);
    dx[t] = fmaf(__ldg(ca + c), dz, fmaf(__ldg(cb + c), v, __ldg(cc + c)));
  }
}

extern "C" int pasco_bn_bwd_apply(const float* dy, const float* x, int64_t n, int32_t C, const float* scale,
                                  const float* shift, int32_t act, const float* coef_a, const float* coef_b,
                                  const float* coef_c, float* dx, pasco_stream_t s) {
  if (n == 0 || C == 0) return 0;
  k_bn_bwd_apply<<<grid_for(n * C, 256), 256, 0, (cudaStream_t)s>>>(dy, x, n, C, scale, shift, act, coef_a, coef_b,
                                                                    coef_c, dx);
  PASCO_CHECK_LAUNCH("pasco_bn_bwd_apply");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// atomic float max (works with a -inf initial value)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax((int*)addr, __float_as_int(v));
  else
    atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__global__ void k_maxpool(const float* __restrict__ in, const int32_t* __restrict__ parent, int64_t n, int C,
                          float* __restrict__ out) {
  int64_t total = n * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / C;
    int c = (int)(t - i * C);
    int p = __ldg(parent + i);
    if (p >= 0) atomic_max_float(out + (int64_t)p * C + c, __ldg(in + t));
  }
}

extern "C" int pasco_maxpool_forward(const float* in, const int32_t* parent_of, int64_t n_in, int32_t C, float* out,
                                     pasco_stream_t s) {
  if (n_in == 0 || C == 0) return 0;
  k_maxpool<<<grid_for(n_in * C, 256), 256, 0, (cudaStream_t)s>>>(in, parent_of, n_in, C, out);
  PASCO_CHECK_LAUNCH("pasco_maxpool_forward");
  return 0;
}

__global__ void k_scatter_max(const float* __restrict__ src, const int64_t* __restrict__ index, int64_t n, int C,
                              float* __restrict__ out) {
  int64_t total = n * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / C;
    int c = (int)(t - i * C);
    atomic_max_float(out + __ldg(index + i) * C + c, __ldg(src + t));
  }
}

__global__ void k_scatter_argmax(const float* __restrict__ src, const int64_t* __restrict__ index, int64_t n, int C,
                                 const float* __restrict__ out, unsigned long long* __restrict__ arg) {
  int64_t total = n * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / C;
    int c = (int)(t - i * C);
    int64_t o = __ldg(index + i) * C + c;
    if (__ldg(src + t) == out[o]) atomicMin(arg + o, (unsigned long long)i);
  }
}

__global__ void k_fill_empty(float* __restrict__ out, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    if (out[t] == -CUDART_INF_F) out[t] = 0.f;
}

extern "C" int pasco_scatter_max(const float* src, const int64_t* index, int64_t n, int32_t C, float* out,
                                 int64_t n_seg, int64_t* argmax, pasco_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  if (n > 0 && C > 0) {
    k_scatter_max<<<grid_for(n * C, 256), 256, 0, st>>>(src, index, n, C, out);
    if (argmax) k_scatter_argmax<<<grid_for(n * C, 256), 256, 0, st>>>(src, index, n, C, out, (unsigned long long*)argmax);
  }
  if (n_seg * C > 0) k_fill_empty<<<grid_for(n_seg * C, 256), 256, 0, st>>>(out, n_seg * C);
  PASCO_CHECK_LAUNCH("pasco_scatter_max");
  return 0;
}
