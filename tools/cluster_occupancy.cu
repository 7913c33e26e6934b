// How many clusters of the decoder sweeps can be co-resident on this GPU?  (cudaOccupancyMaxActiveClusters)
#include <cstdio>
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3.cuh"
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3_bwd.cuh"
using namespace gscan;
template <typename K>
void query(const char* name, K kernel, size_t smem, int csize) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * 40); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = csize; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
  printf("%s: cluster size %d, smem %zu B -> max active clusters %d (%s)\n", name, csize, smem, n, cudaGetErrorString(e));
}
__global__ void __launch_bounds__(512, 1) probe_kernel(int* p) { extern __shared__ float s[]; if (p) p[0] = (int)s[0]; }
int main() {
  int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  printf("SMs: %d\n", sms);
  for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16}) query("probe 190KB", probe_kernel, 190 * 1024, cs);
  const v3::FwdSmem L = v3::fwd_smem(10, 1, 0);
  printf("fwd smem floats %d\n", L.total);
  return 0;
}
