#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
template <typename... A>
inline void cv2eigen(A&&...) { shim_abort(); }
template <typename... A>
inline void eigen2cv(A&&...) { shim_abort(); }
}  // namespace cv
