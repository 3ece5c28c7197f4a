// TEST INFRASTRUCTURE (oracle/): pybind entry points over the *reference's own* CUDA
// implementation of the modulated deformable convolution, so that the GPU parity tests can
// run the real DCNv2 kernels on the B200 next to ours.  Nothing here is product code.
//
// The functions declared below are defined in the reference translation unit
//   /root/reference/src/model/deformconv/src/cuda/modulated_deform_conv_cuda.cu
// (compiled where it lies by oracle/build_ref_cuda.py; declarations follow
//  .../src/cuda/modulated_deform_conv_cuda.h).  The Python names and positional order are
// those of the reference's module `DCN` (.../src/vision.cpp:9-10): input, weight, bias,
// offset, mask, [grad_output,] kh, kw, sh, sw, ph, pw, dh, dw, group, deformable_group,
// im2col_step.
#include <torch/extension.h>
#include <vector>

at::Tensor modulated_deform_conv_cuda_forward(
    const at::Tensor &input, const at::Tensor &weight, const at::Tensor &bias,
    const at::Tensor &offset, const at::Tensor &mask,
    const int kernel_h, const int kernel_w, const int stride_h, const int stride_w,
    const int pad_h, const int pad_w, const int dilation_h, const int dilation_w,
    const int group, const int deformable_group, const int im2col_step);

std::vector<at::Tensor> modulated_deform_conv_cuda_backward(
    const at::Tensor &input, const at::Tensor &weight, const at::Tensor &bias,
    const at::Tensor &offset, const at::Tensor &mask, const at::Tensor &grad_output,
    const int kernel_h, const int kernel_w, const int stride_h, const int stride_w,
    const int pad_h, const int pad_w, const int dilation_h, const int dilation_w,
    const int group, const int deformable_group, const int im2col_step);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "reference DCNv2 modulated deformable convolution (patched build, test oracle only)";
  m.def("modulated_deform_conv_forward", &modulated_deform_conv_cuda_forward);
  m.def("modulated_deform_conv_backward", &modulated_deform_conv_cuda_backward);
}
