// Gradient exchange over NVLink peer memory for the data-parallel training step (reference: Lightning's DDPPlugin = bucketed NCCL
// all-reduce inside torch DDP, utils/__init__.py:114-119).
//
// Every rank owns an ARENA of [2 buffers][world slots][total] fp32 gradients, allocated with cudaMalloc and shared with the other
// processes of the node through CUDA IPC.  As soon as a bucket of gradients is final, rank r PUSHES it into slot r of every rank's
// arena with cudaMemcpyAsync (copy engines over NVLink / NVSwitch: no SM is taken away from the persistent backward kernels,
// unlike an NCCL all-reduce launched under them, which costs every 148-CTA kernel it overlaps a second wave).  The optimizer
// kernel (elementwise.cu: sgd_kernel / adamw_kernel with n_src > 1) then sums the `world` slots in slot order - the same order
// on every rank, so the replicas stay bit-identical - while it applies the step: the reduction never exists as a separate pass.
// The host side (b200/peer.py) orders the pushes after the producing kernels and closes a step with one tiny barrier.
#include <cstring>

#include "common.cuh"
#include "b200_fe.h"

extern "C" {

int b200_peer_alloc(long long bytes, void** ptr, unsigned char* handle64) {
  B200_REQUIRE(bytes > 0 && ptr != nullptr && handle64 != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  void* p = nullptr;
  B200_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return b200_set_error(B200_ERR_CUDA, "peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return B200_OK;
}

int b200_peer_open(const unsigned char* handle64, void** ptr) {
  B200_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200_OK;
}

int b200_peer_close(void* ptr) {
  if (ptr != nullptr) B200_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return B200_OK;
}

int b200_peer_free(void* ptr) {
  if (ptr != nullptr) B200_CHECK_CUDA(cudaFree(ptr));
  return B200_OK;
}

int b200_peer_copy(void* dst, const void* src, long long bytes, void* stream) {
  B200_REQUIRE(dst != nullptr && src != nullptr && bytes >= 0, "peer_copy: bad arguments");
  if (bytes == 0) return B200_OK;
  B200_CHECK_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return B200_OK;
}

}  // extern "C"
