// Probe of the sm_100 TMA row gather (cp.async.bulk.tensor.2d.tile::gather4): (1) which tensor-map box makes it work and
// what lands in shared memory (swizzle, out-of-range rows), (2) its throughput for 128-byte rows from an L2-sized table.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_probe tools/tma_gather_probe.cu && /tmp/tma_probe
// (profiles/r01_tma_gather_probe.txt is the output on the round-1 B200)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// ---- (1) correctness: 4 rows → 512 B of smem, dumped raw
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int4 rows, int col, uint32_t expect_bytes, uint16_t* dump, int* status) {
  __shared__ __align__(1024) uint16_t buf[4 * 64 * 2];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) buf[i] = 0xFFFF;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    mbar_expect(&bar, expect_bytes);
    gather4(buf, &tm, col, rows.x, rows.y, rows.z, rows.w, &bar);
    int spins = 0;
    while (!mbar_try(&bar, 0) && spins < 2000000) ++spins;
    *status = spins;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) dump[i] = buf[i];
}

// ---- (2) throughput: NW issuing warps per CTA, each streaming its own 2-stage ring of 128-row × 128-B tiles (16 KB)
template <int NW>
__global__ void __launch_bounds__(NW * 32) k_stream(const __grid_constant__ CUtensorMap tm_hi, const int* __restrict__ idx, int iters,
                                                    int n_idx_tiles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[NW][2];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) { for (int s = 0; s < NW * 2; ++s) mbar_init(&full[0][0] + s, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long acc = 0;
  uint8_t* mine = base + warp * 32768;
  for (int it = 0; it < iters + 2; ++it) {
    if (it >= 2) {
      const int c = it - 2, s = c & 1;
      while (!mbar_try(&full[warp][s], (c >> 1) & 1)) {}
      acc += *(volatile uint32_t*)(mine + s * 16384 + lane * 128);
      __syncwarp();
    }
    if (it < iters) {
      const int s = it & 1;
      const int tile = ((blockIdx.x * NW + warp) * iters + it) % n_idx_tiles;
      const int4 r = *(const int4*)(idx + (size_t)tile * 128 + lane * 4);
      if (lane == 0) mbar_expect(&full[warp][s], 16384);
      __syncwarp();
      gather4(mine + s * 16384 + lane * 512, &tm_hi, 0, r.x, r.y, r.z, r.w, &full[warp][s]);
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

// ---- (3) issue scaling: NW warps, LANES active lanes each, one buffer per warp (issue → wait → issue ...)
template <int NW, int LANES, int G4B = 512>
__global__ void __launch_bounds__(NW * 32) k_issue(const __grid_constant__ CUtensorMap tm_hi, const int* __restrict__ idx, int iters,
                                                   int n_idx_tiles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[NW];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) { for (int s = 0; s < NW; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* mine = base + warp * (LANES * G4B);
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    const int tile = ((blockIdx.x * NW + warp) * iters + it) % n_idx_tiles;
    const int4 r = *(const int4*)(idx + (size_t)tile * 128 + lane * 4);
    if (lane == 0) mbar_expect(&full[warp], LANES * G4B);
    __syncwarp();
    if (lane < LANES) gather4(mine + lane * G4B, &tm_hi, 0, r.x, r.y, r.z, r.w, &full[warp]);
    while (!mbar_try(&full[warp], it & 1)) {}
    acc += *(volatile uint32_t*)(mine + (lane % LANES) * 64);
    __syncwarp();
  }
  if (acc == 0x1234567ull) *sink = acc;
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int C = 64;
  const long N = 1 << 19;  // 512k rows × 128 B = 64 MB per plane
  std::vector<uint16_t> h((size_t)N * C);
  for (long r = 0; r < N; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(((r & 1023) << 6) | c);  // row tag ≪ 6 | col
  uint16_t *d_hi, *d_lo, *d_dump; int* d_status;
  CK(cudaMalloc(&d_hi, (size_t)N * C * 2)); CK(cudaMalloc(&d_lo, (size_t)N * C * 2));
  CK(cudaMemcpy(d_hi, h.data(), (size_t)N * C * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lo, h.data(), (size_t)N * C * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_dump, 1024)); CK(cudaMalloc(&d_status, 4));

  auto make = [&](void* ptr, int box_rows, CUtensorMap* tm) {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  int good_box = 0;
  for (int box_rows : {1}) {  // box rows 4 → illegal instruction (measured); the gather4 box is one row
    CUtensorMap tm;
    CUresult r = make(d_hi, box_rows, &tm);
    printf("box rows %d: encode rc=%d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    int4 rows = make_int4(5, 17, -1, 1000);
    int st = -1;
    k_probe<<<1, 128>>>(tm, rows, 0, 512, d_dump, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> dump(512);
    CK(cudaMemcpy(dump.data(), d_dump, 1024, cudaMemcpyDeviceToHost));
    printf("  spins=%d (2000000 = timed out)\n", st);
    for (int row = 0; row < 4; ++row) {
      printf("  smem row %d:", row);
      for (int ch = 0; ch < 8; ++ch) printf(" [r%d c%d]", dump[row * 64 + ch * 8] >> 6, dump[row * 64 + ch * 8] & 63);
      printf("\n");
    }
    if (st < 2000000 && good_box == 0 && (dump[0] >> 6) == 5) good_box = box_rows;
  }
  printf("usable box rows: %d\n", good_box);
  if (!good_box) return 0;

  // throughput
  CUtensorMap tm_hi, tm_lo;
  make(d_hi, good_box, &tm_hi); make(d_lo, good_box, &tm_lo);
  const int n_tiles = 8192;
  int* d_idx; CK(cudaMalloc(&d_idx, (size_t)n_tiles * 128 * 4));
  unsigned long long* d_sink; CK(cudaMalloc(&d_sink, 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto fill = [&](int miss_pct, int miss_value, bool sorted) {
    std::vector<int> hidx((size_t)n_tiles * 128);
    uint64_t s = 88172645463325252ull;
    long next = 0;
    for (auto& v : hidx) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      if ((int)(s % 100) < miss_pct) v = miss_value;
      else if (sorted) { next = (next + 1 + (long)((s >> 20) % 3)) % N; v = (int)next; }
      else v = (int)((s >> 8) % N);
    }
    CK(cudaMemcpy(d_idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
  };
  auto run = [&](auto kern, int nw, const char* what) {
    const int smem = nw * 32768 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000, grid = 148;
    kern<<<grid, nw * 32, smem>>>(tm_hi, d_idx, 50, n_tiles, d_sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    kern<<<grid, nw * 32, smem>>>(tm_hi, d_idx, iters, n_tiles, d_sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)grid * nw * iters * 16384;
    printf("%-34s %d issuing warps/SM: %.3f ms, %.1f GB/s aggregate, %.1f B/clk/SM @1.9GHz, %.1f ns per gather4 per SM\n", what, nw, ms,
           bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9, ms * 1e6 / (iters * nw * 32.0));
  };
  printf("-- (2) streaming, 512-B gather4 (bf16 rows of 64 ch, SWIZZLE_128B), 2-stage ring per warp\n");
  fill(10, -1, false);
  run(k_stream<1>, 1, "random rows, 10% missing(-1)");
  run(k_stream<2>, 2, "random rows, 10% missing(-1)");
  run(k_stream<4>, 4, "random rows, 10% missing(-1)");
  run(k_stream<6>, 6, "random rows, 10% missing(-1)");
  printf("-- out-of-range rows are zero-filled by the TMA: cost\n");
  fill(65, -1, false);
  run(k_stream<6>, 6, "random rows, 65% missing(-1)");
  fill(90, -1, false);
  run(k_stream<6>, 6, "random rows, 90% missing(-1)");
  fill(90, (int)N + 5, false);
  run(k_stream<6>, 6, "random rows, 90% missing(N+5)");
  fill(100, -1, false);
  run(k_stream<6>, 6, "all missing(-1)");
  printf("-- missing neighbours redirected to real zero rows instead\n");
  fill(65, (int)N - 1, false);
  run(k_stream<6>, 6, "random, 65% -> ONE real zero row");
  fill(100, (int)N - 1, false);
  run(k_stream<6>, 6, "all -> ONE real row");
  {
    std::vector<int> hidx((size_t)n_tiles * 128);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < hidx.size(); ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      hidx[i] = ((int)(s % 100) < 90) ? (int)(N - 1 - (i % 128)) : (int)((s >> 8) % N);
    }
    CK(cudaMemcpy(d_idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    run(k_stream<6>, 6, "random, 90% -> 128 zero rows");
  }
  fill(0, -1, true);
  run(k_stream<6>, 6, "ascending nearby rows, 0% missing");
  auto run2 = [&](auto kern, int nw, int lanes, int g4b, const CUtensorMap& tmx) {
    const int smem = nw * lanes * g4b + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000, grid = 148;
    kern<<<grid, nw * 32, smem>>>(tmx, d_idx, 50, n_tiles, d_sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    kern<<<grid, nw * 32, smem>>>(tmx, d_idx, iters, n_tiles, d_sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n4 = (double)nw * lanes * iters;
    printf("issue->wait loop, %2d warps x %2d lanes: %.3f ms, %.1f ns per gather4 per SM, %.1f GB/s aggregate, %.2f us per warp round\n", nw,
           lanes, ms, ms * 1e6 / n4, n4 * g4b * 148 / ms / 1e6, ms * 1e3 / iters);
  };
  printf("-- (3) issue scaling, 512-B gather4: every warp issues `lanes` gather4s, waits for them, repeats\n");
  fill(10, -1, false);
  run2(k_issue<1, 32>, 1, 32, 512, tm_hi);
  run2(k_issue<1, 16>, 1, 16, 512, tm_hi);
  run2(k_issue<1, 8>, 1, 8, 512, tm_hi);
  run2(k_issue<1, 1>, 1, 1, 512, tm_hi);
  run2(k_issue<4, 32>, 4, 32, 512, tm_hi);
  run2(k_issue<8, 32>, 8, 32, 512, tm_hi);
  run2(k_issue<12, 32>, 12, 32, 512, tm_hi);
  run2(k_issue<16, 16>, 16, 16, 512, tm_hi);
  run2(k_issue<24, 8>, 24, 8, 512, tm_hi);
  // fp32 rows of 64 channels (256 B), no swizzle: 1 KB per gather4
  float* d_f32; CK(cudaMalloc(&d_f32, (size_t)N * 64 * 4)); CK(cudaMemset(d_f32, 0, (size_t)N * 64 * 4));
  CUtensorMap tm_f32;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)N};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_f32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("-- (4) fp32 rows of 64 channels (256 B, no swizzle), 1 KB per gather4 (map encode rc=%d)\n", (int)r);
  }
  fill(0, -1, false);
  run2(k_issue<1, 32, 1024>, 1, 32, 1024, tm_f32);
  run2(k_issue<2, 32, 1024>, 2, 32, 1024, tm_f32);
  run2(k_issue<4, 32, 1024>, 4, 32, 1024, tm_f32);
  run2(k_issue<6, 32, 1024>, 6, 32, 1024, tm_f32);
  return 0;
}
