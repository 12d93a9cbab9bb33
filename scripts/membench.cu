// Layout micro-benchmark for the sweep kernel's streams (no GBP arithmetic): per warp-tile 22 rows of 512 bytes are
// read (14 potential + 7 camera-bound message + 1 edge state) and 8 rows are written back (7 message + 1 state), by
// persistent autonomous warps with a double-buffered stage filled by the copy engine -- the access pattern of
// k_sweep_tma without its gathers.  Two placements of the rows in HBM:
//   soa   row r of tile t at (r * n_tiles + t) * 512      (quad-SoA arrays: 22 streams 16.7 MB apart)
//   tile  row r of tile t at (t * 22 + r) * 512           (tile-major: one contiguous 11 KB record per tile)
// and two ways to write: STG.128 from registers, or one bulk store of the 4 KB block from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench scripts/membench.cu ; run: ./membench [n_tiles]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define ROWS 22
#define WROWS 8
#define WARPS 8

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* d, const void* s, uint32_t n) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s32(s)), "r"(n) : "memory");
}

template <int TILE_MAJOR, int BULK_STORE>
__global__ void __launch_bounds__(WARPS * 32, 1) k_stream(float4* __restrict__ A, uint32_t n_tiles, float* __restrict__ sink) {
  extern __shared__ __align__(1024) float4 smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* st = smem + (size_t)warp * (2 * ROWS * 32 + 8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(st + 2 * ROWS * 32);
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const uint32_t stride = gridDim.x * WARPS;
  uint32_t t = warp * gridDim.x + blockIdx.x;
  if (t >= n_tiles) return;
  auto issue = [&](uint32_t tile, uint32_t b) {
    mbar_expect(bars + b, ROWS * 512);
    if (TILE_MAJOR) {
      bulk_g2s(st + b * ROWS * 32, A + (size_t)tile * ROWS * 32, ROWS * 512, bars + b);
    } else {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) bulk_g2s(st + (b * ROWS + r) * 32, A + ((size_t)r * n_tiles + tile) * 32, 512, bars + b);
    }
  };
  if (lane == 0) issue(t, 0);
  uint32_t buf = 0, phase = 0;
  float acc = 0.f;
  for (;;) {
    const uint32_t tn = t + stride;
    const bool has_next = tn < n_tiles;
    if (has_next && lane == 0) {
      if (BULK_STORE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the other buffer's store has left smem
      issue(tn, buf ^ 1);
    }
    mbar_wait(bars + buf, (phase >> buf) & 1u);
    phase ^= 1u << buf;
    float4* s = st + buf * ROWS * 32;
    float4 v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) v[r] = s[r * 32 + lane];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc += v[r].x + v[r].w;
    if (BULK_STORE) {
#pragma unroll
      for (int r = 0; r < WROWS; ++r) s[(ROWS - WROWS + r) * 32 + lane] = make_float4(v[r].y, v[r].z, acc, 1.f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (TILE_MAJOR) {
          bulk_s2g(A + ((size_t)t * ROWS + (ROWS - WROWS)) * 32, s + (ROWS - WROWS) * 32, WROWS * 512);
        } else {
#pragma unroll
          for (int r = 0; r < WROWS; ++r) bulk_s2g(A + ((size_t)(ROWS - WROWS + r) * n_tiles + t) * 32, s + (ROWS - WROWS + r) * 32, 512);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
#pragma unroll
      for (int r = 0; r < WROWS; ++r) {
        const size_t row = ROWS - WROWS + r;
        float4* dst = TILE_MAJOR ? A + ((size_t)t * ROWS + row) * 32 + lane : A + (row * n_tiles + t) * 32 + lane;
        *dst = make_float4(v[r].y, v[r].z, acc, 1.f);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    }
    if (!has_next) break;
    t = tn;
    buf ^= 1;
  }
  if (BULK_STORE && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678f) sink[0] = acc;
}

template <int TM, int BS>
float run(float4* A, uint32_t n_tiles, float* sink, int iters) {
  const size_t smem = WARPS * (2 * ROWS * 32 + 8) * 16;
  cudaFuncSetAttribute(k_stream<TM, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k_stream<TM, BS><<<148, WARPS * 32, smem>>>(A, n_tiles, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) k_stream<TM, BS><<<148, WARPS * 32, smem>>>(A, n_tiles, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) std::printf("CUDA error: %s\n", cudaGetErrorString(err));
  return ms / iters;
}

int main(int argc, char** argv) {
  const uint32_t n_tiles = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 33144u;
  const size_t bytes = (size_t)n_tiles * ROWS * 512;
  float4* A = nullptr;
  float* sink = nullptr;
  cudaMalloc(&A, bytes);
  cudaMalloc(&sink, 4);
  cudaMemset(A, 0, bytes);
  const double moved = (double)n_tiles * (ROWS + WROWS) * 512;
  std::printf("n_tiles %u: %.1f MB read + %.1f MB written per launch\n", n_tiles, n_tiles * ROWS * 512 / 1e6, n_tiles * WROWS * 512 / 1e6);
  const float a = run<0, 0>(A, n_tiles, sink, 20), b = run<1, 0>(A, n_tiles, sink, 20), c = run<0, 1>(A, n_tiles, sink, 20),
              d = run<1, 1>(A, n_tiles, sink, 20);
  std::printf("soa   + STG        : %7.1f us  %6.0f GB/s\n", a * 1e3, moved / (a * 1e-3) / 1e9);
  std::printf("tile  + STG        : %7.1f us  %6.0f GB/s\n", b * 1e3, moved / (b * 1e-3) / 1e9);
  std::printf("soa   + bulk store : %7.1f us  %6.0f GB/s\n", c * 1e3, moved / (c * 1e-3) / 1e9);
  std::printf("tile  + bulk store : %7.1f us  %6.0f GB/s\n", d * 1e3, moved / (d * 1e-3) / 1e9);
  return 0;
}
