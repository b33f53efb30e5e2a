// hostmem.cu — pinned host staging buffers for frames and label maps (the two ends of the path's PCIe traffic).
//
// The reference stages frames in pageable torch tensors and lets `.cuda()` bounce them (seg_video_new.py,
// semantic_seg.py:440-444).  At 8 GPUs the per-step H2D copies of all ranks share the host's memory system and its
// IOMMU, so how the staging buffer is allocated matters more than the copy call:
//   DRNB200_HOST_PINNED  cudaHostAlloc(portable)                       — what torch's pin_memory() gives
//   DRNB200_HOST_WC      + cudaHostAllocWriteCombined                  — no CPU-cache snooping on the DMA reads;
//                                                                        fast to fill sequentially, very slow to READ
//                                                                        from the CPU: frame (input) buffers only
//   DRNB200_HOST_HUGE    2 MiB-aligned mmap + MADV_HUGEPAGE + cudaHostRegister(portable): the buffer is backed by
//                        transparent huge pages where the kernel grants them (512x fewer IOMMU/page-table entries
//                        for the DMA engine to walk); falls back to 4 KiB pages silently if THP is off.
#include "common.cuh"
#include <sys/mman.h>
#include <mutex>
#include <unordered_map>

namespace drnb200 {
struct HostBlock { size_t bytes; int mode; };
static std::mutex g_host_mu;
static std::unordered_map<void*, HostBlock> g_host_blocks;
}  // namespace drnb200

extern "C" {

int drnb200_host_alloc(void** out, uint64_t bytes, int mode) {
  using namespace drnb200;
  DRN_REQUIRE(out && bytes > 0, "host_alloc: null pointer or zero size");
  DRN_REQUIRE(mode == DRNB200_HOST_PINNED || mode == DRNB200_HOST_WC || mode == DRNB200_HOST_HUGE,
              "host_alloc: unknown mode %d", mode);
  void* p = nullptr;
  size_t len = (size_t)bytes;
  if (mode == DRNB200_HOST_HUGE) {
    const size_t kHuge = 2u << 20;
    len = (len + kHuge - 1) & ~(kHuge - 1);
    // over-allocate by one huge page to be able to align the start
    void* raw = mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw == MAP_FAILED) { set_error("host_alloc: mmap of %zu bytes failed", len + kHuge); return DRNB200_E_NOMEM; }
    uintptr_t a = ((uintptr_t)raw + kHuge - 1) & ~(uintptr_t)(kHuge - 1);
    if (a > (uintptr_t)raw) munmap(raw, a - (uintptr_t)raw);
    const size_t tail = ((uintptr_t)raw + len + kHuge) - (a + len);
    if (tail) munmap((void*)(a + len), tail);
    p = (void*)a;
    madvise(p, len, MADV_HUGEPAGE);                       // advisory: ignored when THP is disabled
    memset(p, 0, len);                                    // fault the pages in (as huge pages) before pinning
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) { munmap(p, len); return cuda_fail(e, "cudaHostRegister"); }
  } else {
    unsigned flags = cudaHostAllocPortable | (mode == DRNB200_HOST_WC ? cudaHostAllocWriteCombined : 0u);
    DRN_CUDA(cudaHostAlloc(&p, len, flags));
  }
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[p] = HostBlock{len, mode};
  }
  *out = p;
  return DRNB200_OK;
}

int drnb200_host_free(void* p) {
  using namespace drnb200;
  if (!p) return DRNB200_OK;
  HostBlock b;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_blocks.find(p);
    DRN_REQUIRE(it != g_host_blocks.end(), "host_free: %p was not returned by drnb200_host_alloc", p);
    b = it->second;
    g_host_blocks.erase(it);
  }
  if (b.mode == DRNB200_HOST_HUGE) {
    cudaError_t e = cudaHostUnregister(p);
    munmap(p, b.bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
  } else {
    DRN_CUDA(cudaFreeHost(p));
  }
  return DRNB200_OK;
}

}  // extern "C"
