// Standalone TMA probe (run under gpurun, one mode per process because a fault kills the context):
//   tma_probe2 <mode> <boxW> <boxH> <x0> <y0>
// mode 0: 1-D cp.async.bulk (no descriptor)        mode 1: 2-D tensor tile, descriptor in a __grid_constant__ parameter
// mode 2: descriptor in global memory, no fence     mode 3: descriptor in global memory + fence.proxy.tensormap acquire
// mode 4: descriptor in __constant__ memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__constant__ CUtensorMap c_map;

__global__ void k_probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, const uint32_t *src, int mode, int bw, int bh, int x0, int y0,
                        uint32_t *out, int *status) {
  extern __shared__ __align__(128) uint32_t tile[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned bytes = (unsigned)(bw * bh * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (mode == 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)), "l"(src), "r"(bytes),
                   "r"(smem_u32(&bar))
                   : "memory");
    } else {
      const void *tm = mode == 1 ? (const void *)&pmap : mode == 4 ? (const void *)&c_map : (const void *)gmap;
      if (mode == 3) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)), "l"(tm),
                   "r"(x0), "r"(y0), "r"(smem_u32(&bar))
                   : "memory");
    }
  }
  unsigned done = 0;
  int spin = 0;
  for (; !done && spin < (1 << 22); spin++) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  if (threadIdx.x == 0) { status[0] = (int)done; status[1] = spin; }
  __syncthreads();
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1, BW = argc > 2 ? atoi(argv[2]) : 36, BH = argc > 3 ? atoi(argv[3]) : 12;
  const int x0 = argc > 4 ? atoi(argv[4]) : 30, y0 = argc > 5 ? atoi(argv[5]) : 6;
  const int W = 640, H = 480;
  std::vector<uint32_t> h((size_t)W * H);
  for (int i = 0; i < W * H; i++) h[i] = 1000000u + i;
  uint32_t *d, *dout;
  int *dst;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&dout, BW * BH * 4);
  cudaMalloc(&dst, 8);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  alignas(64) CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}, gstr[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = ((Fn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap *gmap;
  cudaMalloc(&gmap, sizeof(map));
  cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_map, &map, sizeof(map));
  cudaMemset(dout, 0xff, BW * BH * 4);
  cudaMemset(dst, 0, 8);
  cudaDeviceSynchronize();
  k_probe<<<1, 256, BW * BH * 4 + 128>>>(map, gmap, d + x0 + y0 * W, mode, BW, BH, x0, y0, dout, dst);
  cudaError_t le = cudaDeviceSynchronize();
  int st[2] = {0, 0};
  std::vector<uint32_t> o(BW * BH);
  cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(o.data(), dout, BW * BH * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int yy = 0; yy < BH; yy++)
    for (int xx = 0; xx < BW; xx++) {
      const int gx = x0 + xx, gy = y0 + yy;
      uint32_t want = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? 1000000u + gx + gy * W : 0u;
      if (mode == 0) want = 1000000u + x0 + y0 * W + yy * BW + xx;
      bad += o[yy * BW + xx] != want;
    }
  printf("mode %d box %dx%d origin (%d,%d): encode=%d sync=%s done=%d spins=%d mismatches=%d first=%u\n", mode, BW, BH, x0, y0, (int)r, cudaGetErrorString(le), st[0],
         st[1], bad, o[0]);
  return le != cudaSuccess;
}
