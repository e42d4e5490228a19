// Symmetric-memory runtime: peer-mappable HBM regions and NVLS multicast objects (N6 replacement).
//
// The reference moves every byte through mpi4py point-to-point / Bcast calls (src/master/baseline_master.py:
// 156-200, src/worker/baseline_worker.py:163-202).  On an NVSwitch box the transport is the memory system:
// each rank allocates its arenas with the CUDA virtual-memory-management API so that the physical allocation
// can be exported as a POSIX file descriptor, handed to the other ranks (SCM_RIGHTS, see parallel/symm.py),
// imported and mapped there; kernels then issue plain ld/st (or multimem.st through a multicast mapping that
// the switch replicates) on those addresses.  This file is the thin C ABI over the driver calls; policy
// (who exports what to whom) lives in Python.
//
// libcuda is resolved at run time through cudaGetDriverEntryPoint so the library links against nothing but
// the CUDA runtime and loads on machines where only the runtime is present.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

namespace {

template <typename T>
T drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<T>(p);
}

#define DRV(fn) static auto p_##fn = drv<decltype(&fn)>(#fn); if (!p_##fn) return -100
#define CK(expr) do { CUresult _r = (expr); if (_r != CUDA_SUCCESS) return (int)_r; } while (0)

CUmemAllocationProp make_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int map_handle(int device, CUmemGenericAllocationHandle h, size_t size, size_t align, void** ptr_out) {
  DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
  CUdeviceptr va = 0;
  CK(p_cuMemAddressReserve(&va, size, align, 0, 0));
  CK(p_cuMemMap(va, size, 0, h, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CK(p_cuMemSetAccess(va, size, &acc, 1));
  *ptr_out = reinterpret_cast<void*>(va);
  return 0;
}

}  // namespace

extern "C" {

int drc_rt_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFree(0);
}

// Allocation granularity for peer-shareable memory on `device`.
int drc_rt_granularity(int device, unsigned long long* out) {
  DRV(cuMemGetAllocationGranularity);
  CUmemAllocationProp prop = make_prop(device);
  size_t g = 0;
  CK(p_cuMemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  *out = g;
  return 0;
}

// Allocate `size` bytes (multiple of the granularity) of exportable device memory, map it locally.
int drc_rt_alloc(int device, unsigned long long size, void** ptr_out, unsigned long long* handle_out, int* fd_out) {
  DRV(cuMemCreate); DRV(cuMemExportToShareableHandle);
  CUmemAllocationProp prop = make_prop(device);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemCreate(&h, size, &prop, 0));
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  int r = map_handle(device, h, size, 0, ptr_out);
  if (r) return r;
  *handle_out = (unsigned long long)h;
  *fd_out = fd;
  return 0;
}

// Import a peer's allocation from its file descriptor and map it for `device`.
int drc_rt_import(int device, int fd, unsigned long long size, void** ptr_out, unsigned long long* handle_out) {
  DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  int r = map_handle(device, h, size, 0, ptr_out);
  if (r) return r;
  *handle_out = (unsigned long long)h;
  return 0;
}

int drc_rt_unmap(void* ptr, unsigned long long size, unsigned long long handle) {
  DRV(cuMemUnmap); DRV(cuMemAddressFree); DRV(cuMemRelease);
  CK(p_cuMemUnmap((CUdeviceptr)ptr, size));
  CK(p_cuMemAddressFree((CUdeviceptr)ptr, size));
  if (handle) CK(p_cuMemRelease((CUmemGenericAllocationHandle)handle));
  return 0;
}

int drc_rt_close_fd(int fd) { return close(fd); }

// ---------------------------------------------------------------- NVLS multicast
int drc_rt_mc_supported(int device, int* out) {
  DRV(cuDeviceGetAttribute);
  int v = 0;
  CK(p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device));
  *out = v;
  return 0;
}

static CUmulticastObjectProp mc_prop(int ndev, unsigned long long size) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned int)ndev;
  p.size = size;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

int drc_rt_mc_granularity(int ndev, unsigned long long size, unsigned long long* out) {
  DRV(cuMulticastGetGranularity);
  CUmulticastObjectProp p = mc_prop(ndev, size);
  size_t g = 0;
  CK(p_cuMulticastGetGranularity(&g, &p, CU_MULTICAST_GRANULARITY_MINIMUM));
  *out = g;
  return 0;
}

int drc_rt_mc_create(int ndev, unsigned long long size, unsigned long long* handle_out, int* fd_out) {
  DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
  CUmulticastObjectProp p = mc_prop(ndev, size);
  CUmemGenericAllocationHandle h;
  CK(p_cuMulticastCreate(&h, &p));
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *handle_out = (unsigned long long)h;
  *fd_out = fd;
  return 0;
}

int drc_rt_mc_import(int fd, unsigned long long* handle_out) {
  DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  *handle_out = (unsigned long long)h;
  return 0;
}

int drc_rt_mc_add_device(unsigned long long mc, int device) {
  DRV(cuMulticastAddDevice);
  CK(p_cuMulticastAddDevice((CUmemGenericAllocationHandle)mc, device));
  return 0;
}

int drc_rt_mc_bind(unsigned long long mc, unsigned long long mc_offset, unsigned long long mem, unsigned long long mem_offset,
                   unsigned long long size) {
  DRV(cuMulticastBindMem);
  CK(p_cuMulticastBindMem((CUmemGenericAllocationHandle)mc, mc_offset, (CUmemGenericAllocationHandle)mem, mem_offset, size, 0));
  return 0;
}

int drc_rt_mc_map(int device, unsigned long long mc, unsigned long long size, void** ptr_out) {
  return map_handle(device, (CUmemGenericAllocationHandle)mc, size, 0, ptr_out);
}

// ---------------------------------------------------------------- misc helpers used by the Python runtime
int drc_rt_peer_access(int device, int peer, int* out) {
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, device, peer);
  *out = can;
  return (int)e;
}

int drc_rt_memset_async(void* ptr, int value, unsigned long long bytes, cudaStream_t stream) {
  return (int)cudaMemsetAsync(ptr, value, bytes, stream);
}

int drc_rt_memcpy_async(void* dst, const void* src, unsigned long long bytes, cudaStream_t stream) {
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream);
}

int drc_rt_sm_count(int device, int* out) {
  return (int)cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, device);
}

}  // extern "C"
