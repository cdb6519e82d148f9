// Link stub for the reference's curope.cpp when only its CPU path (rope_2d_cpu) is built as an oracle:
// curope.cpp forward-declares rope_2d_cuda (curope.cpp:9) but its definition (kernels.cu) does not compile
// against torch 2.11.  Test infrastructure only.
#include <torch/extension.h>
void rope_2d_cuda(torch::Tensor, const torch::Tensor, const float, const float) {
  TORCH_CHECK(false, "oracle/_ref/curope_ref is the reference's CPU path only");
}
