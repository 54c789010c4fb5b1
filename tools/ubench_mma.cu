// Micro-benchmark (developer tool, not part of the library): cost of back-to-back tcgen05.mma
// instructions issued by one thread, M=128, as a function of N and operand kind (tf32 K=8, f16 K=16),
// and whether kind::tf32 ignores the low 13 mantissa bits of its operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../deep-turbulence_b200/csrc/tc_ptx.cuh"
using namespace tmg;
namespace tmg { void set_error(const char*, ...) {} std::atomic<int64_t> g_launches{0}; }

__device__ __forceinline__ uint32_t make_idesc(int n, int f16) {
  // D=F32 (bit4), A/B format bits 7-9 / 10-12: tf32 = 2, f16 = 0 (bf16 = 1)
  uint32_t fmt = f16 ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_issue(int N, int f16, int iters, int vary, int nacc, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(N, f16);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 24 * 1024);
    // A: 2 planes x 160 rows x 16 B (LBO = 2560); B: 2 planes x N x 16 B (LBO = N*16)
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t sh = vary ? (uint32_t)((i % 9) * 16) : 0u;
      uint64_t ad = make_desc(a0 + sh, 2560 + 128, 128), bd = make_desc(b0, (uint32_t)N * 16u, 128);
      const uint32_t dcol = tm + (uint32_t)((i % nacc) * N);
      if (f16) mma_f16(dcol, ad, bd, idesc, i >= nacc); else mma_tf32(dcol, ad, bd, idesc, i >= nacc);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}


template <int N, int F16, int NACC>
__global__ void __launch_bounds__(128, 1) k_lean(int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(N, F16);
    const uint64_t ad0 = make_desc(smem_u32(smem), 2560 + 128, 128), bd = make_desc(smem_u32(smem + 24 * 1024), (uint32_t)N * 16u, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint64_t ad = ad0 + (uint64_t)(u % 3);      // shifted start address (16 B units), compile-time
        const uint32_t dcol = tm + (uint32_t)((u % NACC) * N);
        if (F16) mma_f16(dcol, ad, bd, idesc, 1); else mma_tf32(dcol, ad, bd, idesc, 1);
      }
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int N, int F16, int NACC>
void run_lean(long long* d) {
  long long h[2];
  const int iters = 1024;
  cudaFuncSetAttribute(k_lean<N, F16, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k_lean<N, F16, NACC><<<1, 128, 48 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("lean kind=%s nacc=%d N=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", F16 ? "f16 " : "tf32", NACC, N,
         (double)h[0] / iters, (double)h[1] / iters);
}


// Same issue loop with (a) a strided M mapping (SBO = 352 B, as the 16x8-pixel tiles of flow_step_f16.cu), (b) the
// hi/lo x3 descriptor pattern, (c) NOISE other warps streaming shared memory (LDS + STS) meanwhile.
template <int N, int SBO, int X3, int NOISE>
__global__ void __launch_bounds__(32 + 32 * (NOISE > 0 ? NOISE : 1), 1) k_real(int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid < 32) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(N, 1);
      // A: [hl 2][2 planes][512 pos][16 B] = 32 KB; B: [9 taps][hl 2][2 planes][N][16 B]
      const uint64_t a0 = make_desc(smem_u32(smem), 8192, SBO), b0 = make_desc(smem_u32(smem + 64 * 1024), (uint32_t)N * 16u, 128);
      long long t0 = clock64();
      for (int i = 0; i < iters; i += 18) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t ad = a0 + (uint64_t)((3 + tap / 3 - 1) * 22 + 3 + 8 * mt + tap % 3 - 1);
            const uint64_t bd = b0 + (uint64_t)(tap * 2 * 2 * N);
            mma_f16(tm + mt * N, ad, bd, idesc, 1);
            if (X3) { mma_f16(tm + mt * N, ad + 1024, bd, idesc, 1); mma_f16(tm + mt * N, ad, bd + 2 * N, idesc, 1); }
          }
        }
      }
      long long t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
      stop = 1;
    }
  } else if (NOISE > 0) {
    float* f = reinterpret_cast<float*>(smem + 100 * 1024);
    float acc = 0.f;
    int it = 0;
    float c0 = tid, c1 = 1.f, c2 = 2.f, c3 = 3.f, c4 = 4.f, c5 = 5.f, c6 = 6.f, c7 = 7.f;
    while (!stop && it < 200000) {
      if (SBO == 128) {      // ALU-heavy noise: 8 independent FMA chains per thread (scheduler contention, no smem)
#pragma unroll 16
        for (int j = 0; j < 16; ++j) {
          c0 = fmaf(c0, 1.0001f, 0.5f); c1 = fmaf(c1, 1.0001f, 0.5f); c2 = fmaf(c2, 1.0001f, 0.5f); c3 = fmaf(c3, 1.0001f, 0.5f);
          c4 = fmaf(c4, 1.0001f, 0.5f); c5 = fmaf(c5, 1.0001f, 0.5f); c6 = fmaf(c6, 1.0001f, 0.5f); c7 = fmaf(c7, 1.0001f, 0.5f);
        }
      } else {               // shared-memory noise: independent 16-byte loads and stores
        float4* f4 = reinterpret_cast<float4*>(f);
#pragma unroll 8
        for (int j = 0; j < 8; ++j) { float4 t = f4[(tid + j * 67 + it) & 2047]; acc += t.x; f4[(tid + j * 131 + it * 3) & 2047] = t; }
      }
      ++it;
    }
    acc += c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    if (acc == 123.456f) out[2] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int N, int SBO, int X3, int NOISE>
void run_real(long long* d) {
  long long h[2];
  const int iters = 18 * 40;
  cudaFuncSetAttribute(k_real<N, SBO, X3, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_real<N, SBO, X3, NOISE><<<1, 32 + 32 * (NOISE > 0 ? NOISE : 1), 160 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const int n = iters * (X3 ? 3 : 1);
  printf("real N=%3d SBO=%3d x3=%d noise_warps=%2d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", N, SBO, X3, NOISE,
         (double)h[0] / n, (double)h[1] / n);
}

// fold pattern of conv3x3_f16.cu: per tap [A_hi x (B_hi|B_lo), N' = 2N] + [A_lo x B_hi, N] against three N-wide MMAs
template <int N, int FOLD>
__global__ void __launch_bounds__(128, 1) k_fold(int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid < 32 && elect_one()) {
    const uint32_t idesc = make_idesc(N, 1), idesc2 = make_idesc(2 * N, 1);
    const uint64_t a0 = make_desc(smem_u32(smem), 5312, 288), b0 = make_desc(smem_u32(smem + 64 * 1024), (uint32_t)(2 * N) * 16u, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 18) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t ad = a0 + (uint64_t)((tap / 3) * 18 + 8 * mt + tap % 3);
          const uint64_t bd = b0 + (uint64_t)(tap * 2 * 2 * N);
          if (FOLD) { mma_f16(tm + mt * 128, ad, bd, idesc2, 1); mma_f16(tm + mt * 128, ad + 664, bd, idesc, 1); }
          else { mma_f16(tm + mt * 128, ad, bd, idesc, 1); mma_f16(tm + mt * 128, ad + 664, bd, idesc, 1); mma_f16(tm + mt * 128, ad, bd + N, idesc, 1); }
        }
      }
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int N, int FOLD>
void run_fold(long long* d) {
  long long h[2];
  const int iters = 18 * 40;
  cudaFuncSetAttribute(k_fold<N, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_fold<N, FOLD><<<1, 128, 160 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("fold N=%3d fold=%d : %.1f cycles per tap (issue), %.1f (complete)\n", N, FOLD, (double)h[0] / iters, (double)h[1] / iters);
}

// truncation test: D = A*B with A row r = (1 + r*2^-20) in channel 0, B col 0 = 1 at k=0
__global__ void __launch_bounds__(128, 1) k_trunc(float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  float* A = reinterpret_cast<float*>(smem);              // [2 planes][128 rows][4]
  float* B = reinterpret_cast<float*>(smem + 8192);       // [2 planes][16][4]
  for (int i = tid; i < 2 * 128 * 4; i += 128) A[i] = 0.f;
  for (int i = tid; i < 2 * 16 * 4; i += 128) B[i] = 0.f;
  __syncthreads();
  A[tid * 4] = 1.f + (tid & 1) * (1.f / 2048.f + 1.f / 4096.f);
  if (tid == 0) { B[0] = 1.f + 1.f / 2048.f + 1.f / 4096.f; mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 32);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    mma_tf32(tm, make_desc(smem_u32(A), 2048, 128), make_desc(smem_u32(B), 256, 128), make_idesc(16, 0), 0);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  float v[16];
  tmem_ld16(tm + ((uint32_t)((tid >> 5) * 32) << 16), v);
  out[tid] = v[0];
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  run_lean<16, 0, 1>(d);
  run_lean<32, 0, 1>(d);
  run_lean<48, 0, 1>(d);
  run_lean<64, 0, 1>(d);
  run_lean<128, 0, 1>(d);
  run_lean<256, 0, 1>(d);
  run_lean<16, 0, 4>(d);
  run_lean<32, 0, 4>(d);
  run_lean<48, 0, 4>(d);
  run_lean<64, 0, 4>(d);
  run_lean<128, 0, 4>(d);
  run_lean<16, 1, 1>(d);
  run_lean<32, 1, 1>(d);
  run_lean<48, 1, 1>(d);
  run_lean<64, 1, 1>(d);
  run_lean<128, 1, 1>(d);
  run_lean<256, 1, 1>(d);
  run_lean<16, 1, 4>(d);
  run_lean<32, 1, 4>(d);
  run_lean<48, 1, 4>(d);
  run_lean<64, 1, 4>(d);
  run_lean<128, 1, 4>(d);
  run_real<16, 128, 0, 0>(d); run_real<16, 352, 0, 0>(d); run_real<16, 352, 1, 0>(d);
  run_real<16, 352, 1, 8>(d); run_real<16, 352, 1, 20>(d); run_real<32, 352, 1, 20>(d); run_real<16, 128, 1, 20>(d);
  run_lean<96, 1, 1>(d);
  run_fold<48, 0>(d); run_fold<48, 1>(d); run_fold<16, 0>(d); run_fold<16, 1>(d); run_fold<32, 0>(d); run_fold<32, 1>(d); run_fold<64, 0>(d); run_fold<64, 1>(d);
  float* o; cudaMalloc(&o, 512);
  k_trunc<<<1, 128, 16 * 1024>>>(o);
  cudaDeviceSynchronize();
  float ho[128]; cudaMemcpy(ho, o, 512, cudaMemcpyDeviceToHost);
  printf("tf32 operand handling: B = 1+2^-11+2^-12; A[even] = 1, A[odd] = B.  truncation -> D = 1, 1;  round-to-nearest -> D = 1.000977, 1.001954\n");
  for (int r : {0, 1, 2, 3}) printf("  r=%d D=%.9f\n", r, ho[r]);
  return 0;
}
