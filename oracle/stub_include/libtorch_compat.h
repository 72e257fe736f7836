// TEST INFRASTRUCTURE — LibTorch API drift between the reference's pinned 2.0.1 (docker/Dockerfile:109) and the
// 2.11 in this image, force-included (-include) when the reference's gaussian_model.cpp is compiled unmodified:
// the optimizer state map was keyed by c10::guts::to_string(TensorImpl*) (a std::string) in 2.0.1 and is keyed by
// the void* itself in 2.11, and c10::guts::to_string is gone.  Mapping to_string(ptr) -> ptr keeps every
// state.find(key) / state[key] / state.erase(key) in the reference source meaning what it meant.
#pragma once
#include <c10/util/C++17.h>
namespace c10 { namespace guts {
template <class T> inline void* to_string(T* p) { return const_cast<void*>(static_cast<const void*>(p)); }
} }
