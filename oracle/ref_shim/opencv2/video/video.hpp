#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
template <typename... A>
inline int buildOpticalFlowPyramid(A&&...) { shim_abort(); }
template <typename... A>
inline void calcOpticalFlowPyrLK(A&&...) { shim_abort(); }
}  // namespace cv
