#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
enum { LMEDS = 4, RANSAC = 8, RHO = 16 };
template <typename... A>
inline Mat findHomography(A&&...) { shim_abort(); }
}  // namespace cv
