/* Build-recipe shim for compiling the UNMODIFIED reference sources against the torch of this image (2.11).
 * ATen dropped AT_DISPATCH_ALL_TYPES_AND_HALF (used by csrc/pack_ops/pack_ops_cuda.cu); the modern spelling is
 * AT_DISPATCH_ALL_TYPES_AND(at::ScalarType::Half, ...).  Force-included by oracle/build_ref.py (-include). */
#pragma once
#include <ATen/Dispatch.h>
#include <ATen/core/DeprecatedTypeProperties.h>
/* AT_DISPATCH_*(tensor.type(), ...) (pack_ops_cuda.cu:2796): the overload for the deprecated type object is gone too. */
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
#ifndef AT_DISPATCH_ALL_TYPES_AND_HALF
#define AT_DISPATCH_ALL_TYPES_AND_HALF(TYPE, NAME, ...) AT_DISPATCH_ALL_TYPES_AND(at::ScalarType::Half, TYPE, NAME, __VA_ARGS__)
#endif
