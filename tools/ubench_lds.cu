// Shared-memory wavefront cost of the broadcast patterns the team kernel uses for its dense mat-vec operand
// (the profile says the LSU data pipe, not FP64, is what two teams per SM saturate).  8 warps stream loads of
// one pattern; cycles per warp-load per SM ~ wavefronts per request.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_lds ubench_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>   // 8 or 16 bytes
__global__ void __launch_bounds__(256, 1) lds_kernel(long long *cyc, int reps, const int *offs, double *sink) {
  __shared__ __align__(16) double buf[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) buf[i] = i;
  const int off = offs[threadIdx.x & 31];          // byte offset of this lane's address
  __syncthreads();
  unsigned acc = 0;
  const unsigned char *base = reinterpret_cast<const unsigned char *>(buf);
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const unsigned char *p = base + off + 16 * k + ((i & 1) ? 256 : 0);
      if (VEC == 16) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        acc ^= v.x ^ v.w;
      } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        acc ^= v.x ^ v.y;
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x1234567u) *sink = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  long long *cyc, h; int *offs; double *sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&offs, 128); cudaMalloc(&sink, 8);
  struct Pat { const char *name; int off[32]; } pats[8];
  int np = 0;
  auto add = [&](const char *name, auto f) { pats[np].name = name; for (int l = 0; l < 32; ++l) pats[np].off[l] = f(l); ++np; };
  add("all lanes one address", [](int) { return 0; });
  add("lane parity -> 2 addresses 336 B apart (now)", [](int l) { return (l & 1) * 336; });
  add("lane parity -> 2 adjacent 16 B chunks", [](int l) { return (l & 1) * 16; });
  add("lane parity -> 2 addresses 8 B apart", [](int l) { return (l & 1) * 8; });
  add("lane & 3 -> 4 addresses 168 B apart", [](int l) { return (l & 3) * 168; });
  add("lane & 3 -> 4 adjacent 16 B chunks", [](int l) { return (l & 3) * 16; });
  add("half-warps -> 2 addresses 336 B apart", [](int l) { return (l >> 4) * 336; });
  add("32 distinct consecutive", [](int l) { return l * 16; });
  const int reps = 4000;
  for (int p = 0; p < np; ++p) {
    cudaMemcpy(offs, pats[p].off, 128, cudaMemcpyHostToDevice);
    for (int vec : {8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (vec == 8) lds_kernel<8><<<1, 256>>>(cyc, reps, offs, sink); else lds_kernel<16><<<1, 256>>>(cyc, reps, offs, sink);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-48s LDS.%-3d %6.2f cycles per warp-load (8 warps streaming)\n", pats[p].name, vec * 8, (double)h / (reps * 16.0 * 8));
    }
  }
  return 0;
}
