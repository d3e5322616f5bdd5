// see opencv2/core/core.hpp in this directory: declarations only
#pragma once
#include "opencv2/core/core.hpp"
namespace cv {
class Feature2D {
 public:
  virtual ~Feature2D() {}
  virtual int defaultNorm() const { return 4; }
  virtual int descriptorSize() const { return 0; }
  virtual int descriptorType() const { return 0; }
  virtual void detect(InputArray, std::vector<KeyPoint>&, InputArray = Mat()) { shim_abort(); }
  virtual void compute(InputArray, std::vector<KeyPoint>&, OutputArray) { shim_abort(); }
  virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false) { shim_abort(); }
};
using FeatureDetector = Feature2D;
using DescriptorExtractor = Feature2D;
enum NormTypes2 { NORM_L2 = 4, NORM_HAMMING = 6 };
class BFMatcher {
 public:
  BFMatcher(int = NORM_L2, bool = false) {}
  static Ptr<BFMatcher> create(int = NORM_L2, bool = false) { return std::make_shared<BFMatcher>(); }
  void match(InputArray, InputArray, std::vector<DMatch>&, InputArray = Mat()) const { shim_abort(); }
  void knnMatch(InputArray, InputArray, std::vector<std::vector<DMatch>>&, int, InputArray = Mat(), bool = false) const { shim_abort(); }
};
}  // namespace cv
namespace cv {
#define XIVO_SHIM_DETECTOR(NAME)                          \
  class NAME : public Feature2D {                         \
   public:                                                \
    template <typename... A>                              \
    static Ptr<NAME> create(A&&...) { return std::make_shared<NAME>(); } \
  };
XIVO_SHIM_DETECTOR(FastFeatureDetector)
XIVO_SHIM_DETECTOR(BRISK)
XIVO_SHIM_DETECTOR(ORB)
XIVO_SHIM_DETECTOR(AgastFeatureDetector)
XIVO_SHIM_DETECTOR(GFTTDetector)
XIVO_SHIM_DETECTOR(SIFT)
}  // namespace cv
