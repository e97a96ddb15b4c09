// vmm.h -- one virtual address range backed by physical memory on SEVERAL devices (CUDA virtual memory management),
// mapped read/write on all of them: the "distributed shared memory" the slab-decomposed single domain runs on
// (BASELINE config 5).  A striped allocation puts the k-th n-th of its bytes on device k, so a row-major field is
// row-slab-distributed and a strip-major skewed array is strip-distributed, while every kernel keeps addressing it as
// ONE array: halo reads and the few remote writes travel over NVLink as ordinary loads and stores.
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): the library does not link libcuda,
// so it still loads on a machine without a driver (CPU-side tests).
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace rlfc {

class VmmPool {
 public:
  // devices: CUDA runtime ordinals; every pair must be peer-capable.  Returns 0 or a negative RLFC_E* code.
  int init(const std::vector<int>& devices, std::string& err);
  // `bytes` of device memory visible to all devices at the same address; striped = split evenly over the devices in
  // address order (allocation granularity: 2 MiB per device), otherwise resident on devices[home]
  int alloc(void** out, size_t bytes, bool striped, int home, std::string& err);
  void release();
  size_t granularity() const { return gran_; }
  size_t bytes_mapped() const { return mapped_; }
  ~VmmPool() { release(); }

 private:
  struct Block { unsigned long long va; size_t size; std::vector<unsigned long long> handles; std::vector<size_t> chunk; };
  std::vector<int> dev_;
  std::vector<Block> blocks_;
  size_t gran_ = 0, mapped_ = 0;
  void* fn_[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace rlfc
