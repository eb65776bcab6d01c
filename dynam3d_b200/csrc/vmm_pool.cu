// Growable device arrays with STABLE addresses for the per-episode 3D token memory (the reference grows its numpy / torch arrays by
// concatenation, feature_fields.py:557-570, 643-648, 715-730 -- a re-allocation plus a full copy per append).
//
// B200-first layout: each pool reserves a large virtual-address range once (cuMemAddressReserve: e.g. 4 Mi patches x 1536 B = 6 GiB of VA
// costs nothing out of the 180 GB of HBM3e) and commits physical memory in 2 MiB-granular chunks only when the rollout needs it
// (cuMemCreate + cuMemMap + cuMemSetAccess).  Growth therefore never copies, never moves the base pointer (the pointer tables handed to
// the view runtime stay valid) and does not synchronise the device; new chunks are zero-filled on the caller's stream.
// The driver entry points come from cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda (it must load on the
// GPU-less build box for the ABI checks).
#include <vector>

#include "common.cuh"

namespace {

struct DriverApi {
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

template <typename F>
bool entry(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

DriverApi& api() {
  static DriverApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.ok = entry("cuMemAddressReserve", a.memAddressReserve) && entry("cuMemAddressFree", a.memAddressFree) && entry("cuMemCreate", a.memCreate) &&
           entry("cuMemRelease", a.memRelease) && entry("cuMemMap", a.memMap) && entry("cuMemUnmap", a.memUnmap) &&
           entry("cuMemSetAccess", a.memSetAccess) && entry("cuMemGetAllocationGranularity", a.memGetAllocationGranularity);
  }
  return a;
}

struct VmmPool {
  int device = 0;
  CUdeviceptr base = 0;
  size_t va_bytes = 0, mapped = 0, chunk = 0;
  std::vector<CUmemGenericAllocationHandle> handles;
  CUmemAllocationProp prop{};
};

}  // namespace

#define VP(h) (*reinterpret_cast<VmmPool*>(h))

// Reserves `va_bytes` of virtual address space on `device`; nothing is committed yet.  chunk_bytes = commit granularity (rounded up to the
// driver's minimum, 2 MiB on B200); 0 = 32 MiB.  Returns NULL on failure (message in d3d_last_error()).
extern "C" void* d3d_vmm_create(size_t va_bytes, size_t chunk_bytes, int device) {
  DriverApi& a = api();
  if (!a.ok) { d3d_set_error("CUDA virtual memory management entry points are not available"); return nullptr; }
  VmmPool* p = new VmmPool();
  p->device = device;
  p->prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p->prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p->prop.location.id = device;
  size_t gran = 0;
  if (a.memGetAllocationGranularity(&gran, &p->prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) {
    d3d_set_error("cuMemGetAllocationGranularity failed"); delete p; return nullptr;
  }
  if (chunk_bytes == 0) chunk_bytes = (size_t)32 << 20;
  p->chunk = (chunk_bytes + gran - 1) / gran * gran;
  p->va_bytes = (va_bytes + p->chunk - 1) / p->chunk * p->chunk;
  CUresult r = a.memAddressReserve(&p->base, p->va_bytes, 0, 0, 0);
  if (r != CUDA_SUCCESS) { d3d_set_error("cuMemAddressReserve(%zu bytes) failed (%d)", p->va_bytes, (int)r); delete p; return nullptr; }
  return p;
}

// Commits physical memory until at least `bytes` are mapped (no-op when already there); newly mapped chunks are zero-filled on `stream`.
extern "C" int d3d_vmm_ensure(void* h, size_t bytes, void* stream) {
  VmmPool& p = VP(h);
  DriverApi& a = api();
  D3D_REQUIRE(bytes <= p.va_bytes, "pool grew past its reserved virtual address range");
  const size_t old = p.mapped;
  while (p.mapped < bytes) {
    CUmemGenericAllocationHandle hd;
    CUresult r = a.memCreate(&hd, p.chunk, &p.prop, 0);
    if (r != CUDA_SUCCESS) { d3d_set_error("cuMemCreate(%zu) failed (%d): out of device memory?", p.chunk, (int)r); return D3D_ECUDA; }
    r = a.memMap(p.base + p.mapped, p.chunk, 0, hd, 0);
    if (r != CUDA_SUCCESS) { a.memRelease(hd); d3d_set_error("cuMemMap failed (%d)", (int)r); return D3D_ECUDA; }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = p.device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = a.memSetAccess(p.base + p.mapped, p.chunk, &acc, 1);
    if (r != CUDA_SUCCESS) { a.memUnmap(p.base + p.mapped, p.chunk); a.memRelease(hd); d3d_set_error("cuMemSetAccess failed (%d)", (int)r); return D3D_ECUDA; }
    p.handles.push_back(hd);
    p.mapped += p.chunk;
  }
  if (p.mapped > old) D3D_CHECK_CUDA(cudaMemsetAsync((void*)(p.base + old), 0, p.mapped - old, (cudaStream_t)stream));
  return 0;
}

extern "C" uint64_t d3d_vmm_base(void* h) { return (uint64_t)VP(h).base; }
extern "C" size_t d3d_vmm_mapped(void* h) { return VP(h).mapped; }
extern "C" size_t d3d_vmm_reserved(void* h) { return VP(h).va_bytes; }

// Unmaps and frees everything.  The caller guarantees that no kernel still uses the range (the Python owner synchronises the device).
extern "C" int d3d_vmm_destroy(void* h) {
  if (!h) return 0;
  VmmPool* p = reinterpret_cast<VmmPool*>(h);
  DriverApi& a = api();
  for (size_t i = 0; i < p->handles.size(); ++i) {  // one mapping per chunk: unmap them one by one
    a.memUnmap(p->base + i * p->chunk, p->chunk);
    a.memRelease(p->handles[i]);
  }
  if (p->base) a.memAddressFree(p->base, p->va_bytes);
  delete p;
  return 0;
}
