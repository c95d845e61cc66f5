// Developer tool: compile ONE instantiation of the channel-lane kernel to read ptxas -v (registers / spills) in seconds:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -c tools_dev/ptxas_fbt.cu -o /tmp/ptxas_fbt.o
#include <cstdio>
#include <cstdint>
#include "../spacecraft-pose-estimation-framework_b200/csrc/fused_block_t.cuh"
void ptxas_fbt_launch(const CUtensorMap& a, const spef::fbt::FbtParams& q) {
  spef::fbt::fused_block_t_kernel<1, 6, 2, true, true><<<1, 640, 0, 0>>>(a, a, a, q);
}
