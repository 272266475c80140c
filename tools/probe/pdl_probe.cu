// Diagnostic: does programmatic dependent launch overlap consecutive kernels of a captured graph on this box,
// and does it need the max-shared carveout to make two 110 KB CTAs co-resident?   nvcc -arch=sm_100a pdl_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long ull;
__device__ __forceinline__ ull gtime() { ull t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(192, 1) probe(int id, ull* ts, int work_ns, int use_pdl) {
  extern __shared__ unsigned char sm[];
  ull t0 = gtime();
  if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  ull t1 = gtime();
  while (gtime() - t1 < (ull)work_ns) { sm[threadIdx.x] = (unsigned char)id; }
  ull t2 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { ts[id * 3] = t0; ts[id * 3 + 1] = t1; ts[id * 3 + 2] = t2; }
}
int main(int argc, char** argv) {
  const int N = 12;
  ull* ts; cudaMalloc(&ts, N * 3 * sizeof(ull));
  cudaStream_t s; cudaStreamCreate(&s);
  for (int mode = 0; mode < 6; ++mode) {
    const int pdl = mode & 1, carve = (mode >> 1) & 1, smem = (mode >= 4) ? 1024 : 110 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, carve ? 100 : -1);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, i, ts, 5000, pdl);
      if (e != cudaSuccess) { printf("launch err %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaStreamEndCapture(s, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate err %s\n", cudaGetErrorString(e)); return 1; }
    for (int r = 0; r < 3; ++r) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    ull h[N * 3]; cudaMemcpy(h, ts, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mode pdl=%d carveout=%d smem=%d: total %.1f us; per kernel (start->wait_done->end, us rel. to k0 start):\n", pdl, carve, smem, (h[(N - 1) * 3 + 2] - h[0]) / 1e3);
    for (int i = 0; i < N; ++i) printf("  k%-2d start %7.2f  waited %7.2f  end %7.2f   gap_from_prev_end %6.2f\n", i, (h[i * 3] - h[0]) / 1e3, (h[i * 3 + 1] - h[0]) / 1e3, (h[i * 3 + 2] - h[0]) / 1e3, i ? ((double)h[i * 3 + 1] - (double)h[(i - 1) * 3 + 2]) / 1e3 : 0.0);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
