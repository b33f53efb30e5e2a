// probe_tmem_ld.cu — pins the register layout of the tcgen05.ld shapes on this GPU (standalone, not part of the
// library).  Every TMEM cell (lane l, column c) of a 128-lane x 32-column block is written with the value
// (l << 8) | c through tcgen05.st.32x32b (whose layout is known: thread = lane, register i = column i), then read
// back with 16x256b.x1 / 16x128b.x2 / 16x64b.x4 and the (lane, column) found in each register is printed per thread.
// Purpose: the MODE_T epilogue of conv_tc.cu wants a load whose registers pair up as stmatrix(.trans) fragments.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_tmem_ld tools/probe_tmem_ld.cu && tools/probe_tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe_kernel(uint32_t* out /* [3 shapes][128 threads][4 regs] */) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  const uint32_t quarter = base + ((uint32_t)(warp * 32) << 16);     // this warp's 32 TMEM lanes
  // ---- fill: thread = lane (32*warp + lane), register i = column i
  {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)i;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(quarter),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[4];
  // ---- 16x256b.x1: 16 lanes x 8 columns
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(quarter));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int i = 0; i < 4; ++i) out[(0 * 128 + threadIdx.x) * 4 + i] = r[i];
  // ---- 16x128b.x2: 16 lanes x 8 columns
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(quarter));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int i = 0; i < 4; ++i) out[(1 * 128 + threadIdx.x) * 4 + i] = r[i];
  // ---- 16x64b.x4: 16 lanes x 8 columns
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(quarter));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int i = 0; i < 4; ++i) out[(2 * 128 + threadIdx.x) * 4 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base));
  }
}

int main() {
  uint32_t* d = nullptr;
  uint32_t h[3 * 128 * 4];
  if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) { printf("no CUDA device\n"); return 1; }
  cudaMemset(d, 0xFF, sizeof(h));
  probe_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[3] = {"16x256b.x1", "16x128b.x2", "16x64b.x4"};
  for (int s = 0; s < 3; ++s) {
    printf("== tcgen05.ld.%s  (warp 0; register -> (lane,column))\n", names[s]);
    for (int t = 0; t < 32; ++t) {
      printf("t%02d:", t);
      for (int i = 0; i < 4; ++i) {
        const uint32_t v = h[(s * 128 + t) * 4 + i];
        printf("  r%d=(%3u,%2u)", i, v >> 8, v & 0xFF);
      }
      printf("\n");
    }
    // the other warps must show the same pattern shifted by 32 lanes
    int same = 1;
    for (int t = 32; t < 128; ++t)
      for (int i = 0; i < 4; ++i) {
        const uint32_t v = h[(s * 128 + t) * 4 + i], v0 = h[(s * 128 + (t & 31)) * 4 + i];
        if ((v & 0xFF) != (v0 & 0xFF) || (v >> 8) != (v0 >> 8) + 32u * (t >> 5)) same = 0;
      }
    printf("warps 1-3 follow the same pattern (+32 lanes each): %s\n", same ? "yes" : "NO");
  }
  cudaFree(d);
  return 0;
}
