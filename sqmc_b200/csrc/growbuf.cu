// growbuf.cu -- device arrays that grow in place (CUDA virtual memory management).
//
// The matrix arrays (12 B per stored entry, ~100 GB at 10^7 determinants) are rebuilt or extended in every HCI iteration.
// With cudaMalloc / cudaFree each build paid for unmapping the old and mapping the new arrays (0.1 - 0.9 s of wall time
// that no kernel accounts for, and erratic from call to call), and an incremental build needed the old and the new
// matrix side by side.  A GrowBuf reserves a virtual address range once and maps physical chunks behind it on demand:
// the array keeps its address, grows without a copy, keeps its memory across builds and gives the surplus back on
// request (grow_trim).  Driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): no link-time
// dependency on libcuda.
#include <cuda.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "handle.h"

namespace sqmc {

namespace {
struct Drv {
  bool ok = false, tried = false;
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
} D;

template <typename F>
bool resolve(const char *name, F &fn) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}
bool drv_init() {
  if (D.tried) return D.ok;
  D.tried = true;
  D.ok = resolve("cuMemAddressReserve", D.MemAddressReserve) && resolve("cuMemAddressFree", D.MemAddressFree) && resolve("cuMemCreate", D.MemCreate) &&
         resolve("cuMemRelease", D.MemRelease) && resolve("cuMemMap", D.MemMap) && resolve("cuMemUnmap", D.MemUnmap) &&
         resolve("cuMemSetAccess", D.MemSetAccess) && resolve("cuMemGetAllocationGranularity", D.MemGetAllocationGranularity);
  return D.ok;
}
CUmemAllocationProp device_prop() {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = G.device;
  return p;
}
}  // namespace

// reserve `max_bytes` of address space (no memory yet)
int grow_reserve(GrowBuf &b, size_t max_bytes) {
  if (b.base) return 0;
  if (!drv_init()) { set_error("growbuf: CUDA virtual-memory driver entry points unavailable"); return 1; }
  CUmemAllocationProp prop = device_prop();
  size_t gran = 0;
  if (D.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2ull << 20;
  b.chunk = std::max<size_t>(gran, (size_t)256 << 20) / gran * gran;  // 256 MB chunks: ~400 map calls for a 100 GB array
  b.reserved = (max_bytes + b.chunk - 1) / b.chunk * b.chunk;
  CUdeviceptr p = 0;
  CUresult r = D.MemAddressReserve(&p, b.reserved, 0, 0, 0);
  if (r != CUDA_SUCCESS) { set_error("growbuf: cuMemAddressReserve(%zu) failed (%d)", b.reserved, (int)r); b.reserved = 0; return 1; }
  b.base = (unsigned long long)p;
  b.mapped = 0;
  return 0;
}

// make at least `bytes` usable
int grow_ensure(GrowBuf &b, size_t bytes) {
  if (bytes <= b.mapped) return 0;
  const auto t_begin = std::chrono::steady_clock::now();
  const size_t mapped_before = b.mapped;
  struct Report {
    const std::chrono::steady_clock::time_point t0;
    const size_t before;
    GrowBuf &b;
    ~Report() {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      g_alloc_stall_ms += ms;
      const char *e = getenv("SQMC_ALLOC_TRACE");
      if (!(e && atoi(e) > 0)) return;
      if (ms > 10.0) fprintf(stderr, "[sqmc alloc] mapping %.3f GB of a growable array blocked the host for %.1f ms\n", (b.mapped - before) / 1e9, ms);
    }
  } report{t_begin, mapped_before, b};
  if (!b.base) { set_error("growbuf: not reserved"); return 1; }
  if (bytes > b.reserved) { set_error("growbuf: %zu bytes requested, %zu reserved", bytes, b.reserved); return 1; }
  CUmemAllocationProp prop = device_prop();
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  while (b.mapped < bytes) {
    CUmemGenericAllocationHandle hd = 0;
    CUresult r = D.MemCreate(&hd, b.chunk, &prop, 0);
    if (r != CUDA_SUCCESS) {  // out of device memory: give the temporaries cached by the stream-ordered pool back and retry once
      cudaStreamSynchronize(G.stream);
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, G.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
      r = D.MemCreate(&hd, b.chunk, &prop, 0);
    }
    if (r != CUDA_SUCCESS) { set_error("growbuf: out of device memory mapping %zu bytes (have %zu)", bytes, b.mapped); return 1; }
    r = D.MemMap((CUdeviceptr)(b.base + b.mapped), b.chunk, 0, hd, 0);
    if (r != CUDA_SUCCESS) {
      D.MemRelease(hd);
      set_error("growbuf: cuMemMap failed (%d)", (int)r);
      return 1;
    }
    b.handles.push_back((unsigned long long)hd);
    b.mapped += b.chunk;
  }
  // access rights for the whole new range in ONE call (per-chunk calls were the larger part of the mapping time)
  CUresult r = D.MemSetAccess((CUdeviceptr)(b.base + mapped_before), b.mapped - mapped_before, &acc, 1);
  if (r != CUDA_SUCCESS) { set_error("growbuf: cuMemSetAccess failed (%d)", (int)r); return 1; }
  return 0;
}

// give back the chunks above `bytes` (the contents below stay where they are)
void grow_trim(GrowBuf &b, size_t bytes) {
  if (!b.base) return;
  const size_t keep = (bytes + b.chunk - 1) / b.chunk;
  if (b.handles.size() <= keep) return;
  cudaDeviceSynchronize();
  while (b.handles.size() > keep) {
    b.mapped -= b.chunk;
    D.MemUnmap((CUdeviceptr)(b.base + b.mapped), b.chunk);
    D.MemRelease((CUmemGenericAllocationHandle)b.handles.back());
    b.handles.pop_back();
  }
}

void grow_release(GrowBuf &b) {
  if (!b.base) return;
  grow_trim(b, 0);
  D.MemAddressFree((CUdeviceptr)b.base, b.reserved);
  b = GrowBuf();
}

}  // namespace sqmc
