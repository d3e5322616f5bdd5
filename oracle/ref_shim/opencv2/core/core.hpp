// Type-only stand-in for the OpenCV C++ headers (OpenCV C++ is not installed in this image), just enough for the
// reference's headers to parse so that its EKF-side sources (feature.cpp, oos.cpp, update.cpp, manager.cpp, ...) compile
// from where they lie under /root/reference.  No OpenCV functionality is provided: every function here either is a
// trivial container operation or aborts.  The point-cloud path of the reference (VisualMeasPointCloud) never reaches
// an image routine.  Test infrastructure (oracle/_ref), not product code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6

typedef unsigned char uchar;
namespace cv {
[[noreturn]] inline void shim_abort(const char* what = __builtin_FUNCTION()) {
  std::fprintf(stderr, "[opencv shim] %s called: the OpenCV stand-in has no image routines\n", what);
  std::abort();
}
template <typename T, int N>
struct Vec;

template <typename T>
struct Point_ {
  T x{}, y{};
  Point_() {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
using Point = Point_<int>;
using Point2i = Point_<int>;
using Point2f = Point_<float>;
using Point2d = Point_<double>;
template <typename T>
struct Size_ {
  T width{}, height{};
  Size_() {}
  Size_(T w, T h) : width(w), height(h) {}
};
using Size = Size_<int>;
struct Scalar {
  double v[4]{0, 0, 0, 0};
  Scalar() {}
  Scalar(double a, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
struct Rect {
  int x{}, y{}, width{}, height{};
  Rect() {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Range {
  int start{}, end{};
};

class Mat {
 public:
  int rows = 0, cols = 0, flags = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), flags(type) {}
  Mat(int r, int c, int type, void* d, size_t = 0) : rows(r), cols(c), flags(type), data((unsigned char*)d) {}
  Mat(Size s, int type) : rows(s.height), cols(s.width), flags(type) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return flags; }
  int channels() const { return flags == CV_8UC3 ? 3 : 1; }
  Size size() const { return Size(cols, rows); }
  Mat clone() const { return *this; }
  void copyTo(Mat& o) const { o = *this; }
  void copyTo(Mat&& o) const { o = *this; }
  template <typename T, int N>
  Mat(const Vec<T, N>&, bool = true) {}
  template <typename T>
  operator std::vector<T>() const { shim_abort(); }
  Mat operator*(const Mat&) const { shim_abort(); }
  void release() { data = nullptr; rows = cols = 0; }
  void reserveBuffer(size_t) {}
  void push_back(const Mat&) { shim_abort(); }
  bool isContinuous() const { return true; }
  size_t total() const { return (size_t)rows * cols; }
  size_t elemSize() const { return 1; }
  Mat row(int) const { return *this; }
  Mat col(int) const { return *this; }
  Mat t() const { return *this; }
  void setTo(const Scalar&) {}
  void convertTo(Mat&, int, double = 1, double = 0) const { shim_abort(); }
  template <typename T> T& at(int, int = 0) { shim_abort(); }
  template <typename T> const T& at(int, int = 0) const { shim_abort(); }
  template <typename T> T* ptr(int = 0) { return reinterpret_cast<T*>(data); }
  template <typename T> const T* ptr(int = 0) const { return reinterpret_cast<const T*>(data); }
  unsigned char* ptr(int = 0) { return data; }
  const unsigned char* ptr(int = 0) const { return data; }
  Mat operator()(const Rect&) const { return *this; }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  static Mat ones(int r, int c, int t) { return Mat(r, c, t); }
  static Mat eye(int r, int c, int t) { return Mat(r, c, t); }
};
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, 0) {}
  T& operator()(int, int) { shim_abort(); }
};
class SparseMat {};
using InputArray = const Mat&;
using OutputArray = Mat&;
using InputOutputArray = Mat&;

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
  KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
};

template <typename T>
using Ptr = std::shared_ptr<T>;
template <typename T, typename... A>
Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
  int type = 0, maxCount = 0;
  double epsilon = 0;
  TermCriteria() {}
  TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

class FileNodeIterator;
class FileNode {
 public:
  bool empty() const { return true; }
  size_t size() const { return 0; }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  operator int() const { shim_abort(); }
  operator float() const { shim_abort(); }
  operator double() const { shim_abort(); }
  operator std::string() const { shim_abort(); }
  int type() const { return 0; }
  enum { SEQ = 5, MAP = 6 };
  FileNodeIterator begin() const;
  FileNodeIterator end() const;
};
class FileNodeIterator {
 public:
  FileNode operator*() const { return FileNode(); }
  FileNodeIterator& operator++() { return *this; }
  bool operator!=(const FileNodeIterator&) const { return false; }
};
inline FileNodeIterator FileNode::begin() const { return FileNodeIterator(); }
inline FileNodeIterator FileNode::end() const { return FileNodeIterator(); }
class FileStorage {
 public:
  enum Mode { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode root() const { return FileNode(); }
  template <typename T>
  FileStorage& operator<<(const T&) { return *this; }
};
template <typename T>
inline void operator>>(const FileNode&, T&) { shim_abort(); }

inline int cvRound(double v) { return (int)__builtin_nearbyint(v); }
template <typename T, int N>
struct Vec {
  T val[N]{};
  Vec() {}
  template <typename A, typename B, typename C_>
  Vec(A a, B b, C_ c) { val[0] = (T)a; val[1] = (T)b; val[2] = (T)c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
using Vec3d = Vec<double, 3>;
using Vec3b = Vec<unsigned char, 3>;
enum { NORM_INF = 1, NORM_L1 = 2, NORM_MINMAX = 32 };
inline Mat noArray() { return Mat(); }
template <typename... A>
inline void normalize(A&&...) { shim_abort(); }
template <typename... A>
inline double norm(A&&...) { shim_abort(); }
template <typename... A>
inline void rectangle(A&&...) { shim_abort(); }
template <typename... A>
inline void cvtColor(A&&...) { shim_abort(); }
enum { FILLED = -1 };
}  // namespace cv
