// TEST INFRASTRUCTURE — declaration of torch_scatter 2.1.2's C++ API entry point used by the reference
// (gaussian_model.cpp:1635); pytorch_scatter is a go-it-alone dependency absent from this image.  The definition
// (a restatement of its published semantics: out[index[i]][j] = max_i src[i][j], arg = the winning i, untouched
// slots = 0 / src.size(dim)) lives in oracle/model_ref_wrap.cpp.
#pragma once
#include <torch/torch.h>
#include <tuple>
std::tuple<torch::Tensor, torch::Tensor> scatter_max(torch::Tensor src, torch::Tensor index, int64_t dim,
                                                     std::optional<torch::Tensor> optional_out = std::nullopt,
                                                     std::optional<int64_t> dim_size = std::nullopt);
