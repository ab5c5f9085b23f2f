// Probe: throughput of fine-grained REMOTE atomics over NVLink (GPU0 kernels updating GPU1 memory).
// Informs the peer-push design of the sharded frame.  Single process, 2 GPUs.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__global__ void k_u32add(unsigned* p, int C, int n, unsigned seed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  unsigned x = (i * 2654435761u) ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  atomicAdd(p + (x % C), 1u);
}
__global__ void k_u64add(u64* p, int C, int n, unsigned seed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  unsigned x = (i * 2654435761u) ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  atomicAdd(p + (x % C), (u64)i);
}
__global__ void k_u32min(unsigned* p, int C, int n, unsigned seed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  unsigned x = (i * 2654435761u) ^ seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  atomicMin(p + (x % C), x);
}
__global__ void k_local_coherent(unsigned* p, int C, int n, unsigned seed) {   // neighbouring threads -> neighbouring cells
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  atomicAdd(p + ((i / 2 + seed) % C), 1u);
}
int main() {
  int nd = 0; cudaGetDeviceCount(&nd);
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  int can = 0; cudaDeviceCanAccessPeer(&can, 0, 1); printf("peer access 0->1: %d\n", can);
  const int C = 1 << 20;
  unsigned *r32, *l32; u64 *r64, *l64;
  cudaSetDevice(1); cudaMalloc(&r32, C * 4); cudaMalloc(&r64, C * 8); cudaMemset(r32, 0, C * 4); cudaMemset(r64, 0, C * 8);
  cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0);
  cudaMalloc(&l32, C * 4); cudaMalloc(&l64, C * 8); cudaMemset(l32, 0, C * 4); cudaMemset(l64, 0, C * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int n : {200000, 1000000, 8000000}) {
    struct { const char* name; int kind; bool remote; } cases[] = {
      {"u32 add local ", 0, false}, {"u32 add remote", 0, true}, {"u64 add local ", 1, false}, {"u64 add remote", 1, true},
      {"u32 min remote", 2, true}, {"u32 add remote coherent", 3, true}};
    for (auto& cs : cases) {
      float best = 1e9;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        int nb = (n + 255) / 256;
        if (cs.kind == 0) k_u32add<<<nb, 256>>>(cs.remote ? r32 : l32, C, n, rep);
        if (cs.kind == 1) k_u64add<<<nb, 256>>>(cs.remote ? r64 : l64, C, n, rep);
        if (cs.kind == 2) k_u32min<<<nb, 256>>>(r32, C, n, rep);
        if (cs.kind == 3) k_local_coherent<<<nb, 256>>>(r32, C, n, rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("n=%8d %-26s %8.1f us  %8.1f Mops/s\n", n, cs.name, best * 1e3, n / best / 1e3);
    }
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
