// Stub so the reference's *_cuda_kernel.h headers parse without the torch header tree.
// Those headers only DECLARE host glue taking at::Tensor by value; an incomplete type suffices.
#pragma once
namespace at { class Tensor; }
