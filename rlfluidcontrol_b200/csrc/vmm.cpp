// vmm.cpp -- see vmm.h
#include "vmm.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/rlfc.h"

namespace rlfc {

namespace {
enum { kCreate, kRelease, kReserve, kFree, kMap, kUnmap, kSetAccess, kGranularity };
const char* kNames[8] = {"cuMemCreate", "cuMemRelease", "cuMemAddressReserve", "cuMemAddressFree",
                         "cuMemMap", "cuMemUnmap", "cuMemSetAccess", "cuMemGetAllocationGranularity"};
using CreateFn = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using ReleaseFn = CUresult (*)(CUmemGenericAllocationHandle);
using ReserveFn = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using FreeFn = CUresult (*)(CUdeviceptr, size_t);
using MapFn = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using UnmapFn = CUresult (*)(CUdeviceptr, size_t);
using SetAccessFn = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using GranFn = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);

CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}
size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

int VmmPool::init(const std::vector<int>& devices, std::string& err) {
  dev_ = devices;
  for (int k = 0; k < 8; k++) {
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(kNames[k], &fn_[k], cudaEnableDefault, &st) != cudaSuccess || !fn_[k]) {
      err = std::string("driver entry point not available: ") + kNames[k];
      return RLFC_ECUDA;
    }
  }
  for (int a : dev_) {
    if (cudaSetDevice(a) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { err = "cannot initialise device " + std::to_string(a); return RLFC_ENODEV; }
    for (int b : dev_) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) {
        err = "device " + std::to_string(a) + " cannot access device " + std::to_string(b) + " (slab mode needs peer access)";
        return RLFC_ENODEV;
      }
    }
  }
  cudaSetDevice(dev_[0]);
  gran_ = 0;
  for (int a : dev_) {
    size_t g = 0;
    CUmemAllocationProp p = prop_for(a);
    if (((GranFn)fn_[kGranularity])(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) { err = "cuMemGetAllocationGranularity failed"; return RLFC_ECUDA; }
    gran_ = g > gran_ ? g : gran_;
  }
  return RLFC_OK;
}

int VmmPool::alloc(void** out, size_t bytes, bool striped, int home, std::string& err) {
  const int n = (int)dev_.size();
  if (bytes == 0) bytes = 1;
  if (bytes < gran_ * (size_t)n) striped = false;       // less than one granule per device: keep it whole
  Block b{};
  std::vector<int> owner;
  if (striped) {
    const size_t chunk = round_up((bytes + n - 1) / n, gran_);
    for (int k = 0; k < n; k++) { b.chunk.push_back(chunk); owner.push_back(dev_[k]); }
  } else {
    b.chunk.push_back(round_up(bytes, gran_));
    owner.push_back(dev_[home]);
  }
  for (size_t c : b.chunk) b.size += c;
  CUdeviceptr va = 0;
  if (((ReserveFn)fn_[kReserve])(&va, b.size, gran_, 0, 0) != CUDA_SUCCESS) { err = "cuMemAddressReserve failed"; return RLFC_ENOMEM; }
  b.va = va;
  size_t off = 0;
  for (size_t k = 0; k < b.chunk.size(); k++) {
    CUmemGenericAllocationHandle h;
    CUmemAllocationProp p = prop_for(owner[k]);
    CUresult r = ((CreateFn)fn_[kCreate])(&h, b.chunk[k], &p, 0);
    if (r != CUDA_SUCCESS) { err = "cuMemCreate failed (" + std::to_string((int)r) + ") for " + std::to_string(b.chunk[k]) + " bytes on device " + std::to_string(owner[k]); blocks_.push_back(b); return RLFC_ENOMEM; }
    b.handles.push_back(h);
    r = ((MapFn)fn_[kMap])(va + off, b.chunk[k], 0, h, 0);
    if (r != CUDA_SUCCESS) { err = "cuMemMap failed (" + std::to_string((int)r) + ")"; blocks_.push_back(b); return RLFC_ECUDA; }
    off += b.chunk[k];
  }
  std::vector<CUmemAccessDesc> acc(n);
  for (int k = 0; k < n; k++) {
    acc[k].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[k].location.id = dev_[k];
    acc[k].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  CUresult r = ((SetAccessFn)fn_[kSetAccess])(va, b.size, acc.data(), acc.size());
  blocks_.push_back(b);
  if (r != CUDA_SUCCESS) { err = "cuMemSetAccess failed (" + std::to_string((int)r) + ")"; return RLFC_ECUDA; }
  mapped_ += b.size;
  *out = (void*)va;
  return RLFC_OK;
}

void VmmPool::release() {
  for (Block& b : blocks_) {
    if (b.va) {
      if (fn_[kUnmap]) ((UnmapFn)fn_[kUnmap])(b.va, b.size);
      for (auto h : b.handles) ((ReleaseFn)fn_[kRelease])(h);
      ((FreeFn)fn_[kFree])(b.va, b.size);
    }
  }
  blocks_.clear();
  mapped_ = 0;
}

}  // namespace rlfc
