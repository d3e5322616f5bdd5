#pragma once
#include "opencv2/features2d/features2d.hpp"
namespace cv {
namespace xfeatures2d {
XIVO_SHIM_DETECTOR(SURF)
XIVO_SHIM_DETECTOR(FREAK)
XIVO_SHIM_DETECTOR(BriefDescriptorExtractor)
}  // namespace xfeatures2d
}  // namespace cv
