// The reference's Canvas (src/visualize.cpp) draws overlays with OpenCV; it is not part of the filter.  The estimator calls
// it unconditionally, so the reference build under oracle/_ref links these no-ops instead of visualize.cpp.
#include "visualize.h"

namespace xivo {
std::unique_ptr<Canvas> Canvas::instance_ = nullptr;
Canvas::Canvas() : save_frames_(false), frame_number_(0) {}
CanvasPtr Canvas::instance() {
  if (!instance_) instance_ = std::unique_ptr<Canvas>(new Canvas());
  return instance_.get();
}
void Canvas::Delete() { instance_.reset(); }
void Canvas::Update(const cv::Mat&) {}
void Canvas::UpdatePointCloud(const MatX2&) {}
void Canvas::Draw(const FeaturePtr) {}
void Canvas::OverlayStateInfo(const State&, const IMUState&, const Vec9&, int, int, int, double) {}
const void Canvas::SaveFrame() {}
}  // namespace xivo
