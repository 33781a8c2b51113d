// synth_device.cu — the synthetic log generators on the device: config 5's shards (100 / 50 / 25 GB per GPU) are generated
// in HBM, where a host generator plus PCIe would take minutes.  The generator itself is
// synth_gen.h, the same integer code synth.cpp runs on the host; one thread writes one 64 KiB block of the stream.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/matchy_b200.h"
#include "synth_gen.h"

namespace {

// Thread -> block of the stream.  Config 5 changes the line family with every block (family = 1 + (block & 3)), and a warp
// whose lanes write different families runs four generators one after the other: lanes of a warp take blocks with equal
// (k & 3) instead — k = 4 * (32 * (idx / 128) + lane) + (idx / 32) % 4.
__global__ void __launch_bounds__(128) gen_log_kernel(int cfg, sgen::Counts c, uint64_t b0, uint64_t nblocks, uint8_t* out) {
  uint8_t line[sgen::LINE_CAP];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t padded = (nblocks + 127) / 128 * 128;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < padded; idx += stride) {
    const uint64_t k = 4 * (32 * (idx / 128) + (idx & 31)) + ((idx >> 5) & 3);
    if (k < nblocks) sgen::gen_block(cfg, c, b0 + k, out + k * sgen::BLOCK, line);
  }
}

}  // namespace

extern "C" int mgen_log_device(int device, int config, double scale, uint64_t offset, void* dev_out, size_t len) {
  if (config < 1 || config > 5 || !(scale > 0) || offset % sgen::BLOCK != 0 || len % sgen::BLOCK != 0) return MGPU_E_PARAM;
  if (cudaSetDevice(device) != cudaSuccess) return MGPU_E_CUDA;
  if (len == 0) return MGPU_OK;
  const sgen::Counts c = sgen::counts_for(config, scale);
  const uint64_t nblocks = len / sgen::BLOCK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint64_t want = (nblocks + 127) / 128;
  const int grid = (int)(want < (uint64_t)sms * 16 ? want : (uint64_t)sms * 16);
  gen_log_kernel<<<grid, 128>>>(config, c, offset / sgen::BLOCK, nblocks, static_cast<uint8_t*>(dev_out));
  if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return MGPU_E_CUDA;
  return MGPU_OK;
}
