/*
 * oracle/tracker_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement (plain C, scalar) of the third-party arithmetic XIVO's tracker
 * calls on the hot path.  The algorithm lives in OpenCV, which is NOT vendored in
 * /root/reference (find_package(OpenCV), /root/reference/CMakeLists.txt:45, version
 * unpinned, evidence points to 3.4.x); call sites in the reference:
 *   cv::FastFeatureDetector::detect        /root/reference/src/tracker.cpp:39-42, :224
 *   cv::buildOpticalFlowPyramid            /root/reference/src/tracker.cpp:476, :493
 *   cv::calcOpticalFlowPyrLK               /root/reference/src/tracker.cpp:526-528
 * What is restated is the published algorithm of OpenCV's fast.cpp / fast_score.cpp
 * (FAST-9/16 + cornerScore + 3x3 NMS), pyramids.cpp (pyrDown 5-tap, REFLECT_101),
 * lkpyramid.cpp (calcSharrDeriv + LKTrackerInvoker) and color cvtColor BGR2GRAY.
 * PINNING: tests/test_oracle_tracker.py checks every function here against the
 * cv2 4.13 wheel installed in this image (bit-exact for the integer stages,
 * sub-pixel tolerance for LK, whose float summation order is SIMD-path dependent
 * in OpenCV itself).  The reference's own tests hold no golden vectors for the
 * tracker, so cv2 is the only available pin.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * (n - 1) - i;
  }
  return i;
}

/* cvtColor(BGR2GRAY) for 8u, fixed point.  cv2 4.13 (the pin) uses 15-bit
 * coefficients (R 9798, G 19235, B 3735); OpenCV 3.4.x used 14-bit (4899, 9617,
 * 1868) which differs by at most 1 grey level — stated in DESIGN.md. */
void orc_bgr2gray(const uint8_t *bgr, int rows, int cols, uint8_t *gray) {
  for (int i = 0; i < rows * cols; ++i) {
    int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    gray[i] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
  }
}

/* cv::pyrDown for CV_8UC(cn): separable [1 4 6 4 1], BORDER_REFLECT_101,
 * dst = (sum + 128) >> 8, dst size ((cols+1)/2, (rows+1)/2). */
void orc_pyrdown(const uint8_t *src, int rows, int cols, int cn, uint8_t *dst) {
  int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  static const int w[5] = {1, 4, 6, 4, 1};
  int *hbuf = (int *)malloc(sizeof(int) * (size_t)rows * dcols * cn);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < dcols; ++x)
      for (int c = 0; c < cn; ++c) {
        int s = 0;
        for (int k = 0; k < 5; ++k) {
          int sx = reflect101(2 * x + k - 2, cols);
          s += w[k] * src[((size_t)y * cols + sx) * cn + c];
        }
        hbuf[((size_t)y * dcols + x) * cn + c] = s;
      }
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols * cn; ++x) {
      int s = 0;
      for (int k = 0; k < 5; ++k) {
        int sy = reflect101(2 * y + k - 2, rows);
        s += w[k] * hbuf[(size_t)sy * dcols * cn + x];
      }
      dst[(size_t)y * dcols * cn + x] = (uint8_t)((s + 128) >> 8);
    }
  free(hbuf);
}

/* Number of pyramid levels cv::buildOpticalFlowPyramid produces (returns maxLevel
 * actually reached): stops when the next level would be <= winSize. */
int orc_pyramid_levels(int rows, int cols, int win, int max_level) {
  int lvl = 0, r = rows, c = cols;
  for (lvl = 0; lvl <= max_level; ++lvl) {
    r = (r + 1) / 2;
    c = (c + 1) / 2;
    if (c <= win || r <= win) return lvl;
  }
  return max_level;
}

/* calcSharrDeriv: dx = [3 10 3]^T (x) [-1 0 1], dy = [-1 0 1]^T (x) [3 10 3],
 * REFLECT_101 at the image edge; output int16 interleaved (dx,dy) per channel. */
void orc_scharr(const uint8_t *src, int rows, int cols, int cn, int16_t *dst) {
  for (int y = 0; y < rows; ++y) {
    int y0 = reflect101(y - 1, rows), y2 = reflect101(y + 1, rows);
    for (int x = 0; x < cols; ++x) {
      int x0 = reflect101(x - 1, cols), x2 = reflect101(x + 1, cols);
      for (int c = 0; c < cn; ++c) {
#define PX(yy, xx) ((int)src[((size_t)(yy) * cols + (xx)) * cn + c])
        int t0_l = (PX(y0, x0) + PX(y2, x0)) * 3 + PX(y, x0) * 10;
        int t0_r = (PX(y0, x2) + PX(y2, x2)) * 3 + PX(y, x2) * 10;
        int t1_l = PX(y2, x0) - PX(y0, x0);
        int t1_c = PX(y2, x) - PX(y0, x);
        int t1_r = PX(y2, x2) - PX(y0, x2);
#undef PX
        dst[(((size_t)y * cols + x) * cn + c) * 2 + 0] = (int16_t)(t0_r - t0_l);
        dst[(((size_t)y * cols + x) * cn + c) * 2 + 1] = (int16_t)((t1_r + t1_l) * 3 + t1_c * 10);
      }
    }
  }
}

/* ---------------------------------------------------------------- FAST-9/16 */
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* score = OpenCV cornerScore<16>: largest threshold at which the pixel is still a
 * FAST-9 corner; returns 0 when the pixel is not a corner at `threshold`. */
static int fast_score_at(const uint8_t *img, int cols, int x, int y, int threshold) {
  int v = img[(size_t)y * cols + x];
  int d[25];
  for (int k = 0; k < 16; ++k) d[k] = v - img[(size_t)(y + ring_dy[k]) * cols + x + ring_dx[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = -1000;
  for (int s = 0; s < 16; ++s) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; ++k) {
      if (d[s + k] < mn) mn = d[s + k];
      if (d[s + k] > mx) mx = d[s + k];
    }
    if (mn > best) best = mn;    /* all darker-than-centre arc: v - p >= mn */
    if (-mx > best) best = -mx;  /* all brighter arc */
  }
  /* corner at threshold t  <=>  best > t ; score = best - 1 */
  if (best <= threshold) return 0;
  return best - 1;
}

/* Full detect with 3x3 non-max suppression.  Output raster ordered (y major).
 * kp_xy: 2*max_kp ints, kp_score: max_kp ints.  Returns the number found
 * (may exceed max_kp; only the first max_kp are stored).
 * score_map (optional, rows*cols int32) receives the raw score image. */
int orc_fast_detect(const uint8_t *img, int rows, int cols, int threshold, int nonmax,
                    int *kp_xy, int *kp_score, int max_kp, int32_t *score_map) {
  int32_t *sc = score_map ? score_map : (int32_t *)malloc(sizeof(int32_t) * (size_t)rows * cols);
  memset(sc, 0, sizeof(int32_t) * (size_t)rows * cols);
  for (int y = 3; y < rows - 3; ++y)
    for (int x = 3; x < cols - 3; ++x) sc[(size_t)y * cols + x] = fast_score_at(img, cols, x, y, threshold);
  int n = 0;
  for (int y = 3; y < rows - 3; ++y)
    for (int x = 3; x < cols - 3; ++x) {
      int s = sc[(size_t)y * cols + x];
      /* OpenCV keeps scores in a uchar row buffer where a non-corner is 0; a true
       * corner with score 0 (threshold 0, best 1) can never win `score > neighbour`. */
      if (s == 0) continue;
      int keep = 1;
      if (nonmax) {
        for (int dy = -1; dy <= 1 && keep; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            if (sc[(size_t)(y + dy) * cols + x + dx] >= s) { keep = 0; break; }
          }
      }
      if (keep) {
        if (n < max_kp) {
          kp_xy[2 * n] = x;
          kp_xy[2 * n + 1] = y;
          kp_score[n] = s;
        }
        ++n;
      }
    }
  if (!score_map) free(sc);
  return n;
}

/* ------------------------------------------------------- pyramidal Lucas-Kanade */
typedef struct {
  int rows, cols;
  const uint8_t *img; /* rows*cols*cn, no padding: border reads use REFLECT_101 */
  const int16_t *der; /* rows*cols*cn*2 ; reads outside the image return 0 (BORDER_CONSTANT) */
} orc_level;

static inline int pix(const orc_level *L, int cn, int x, int y, int c) {
  return L->img[((size_t)reflect101(y, L->rows) * L->cols + reflect101(x, L->cols)) * cn + c];
}
static inline int der(const orc_level *L, int cn, int x, int y, int c, int which) {
  if (x < 0 || y < 0 || x >= L->cols || y >= L->rows) return 0;
  return L->der[(((size_t)y * L->cols + x) * cn + c) * 2 + which];
}
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

/* LKTrackerInvoker for one level over all points (lkpyramid.cpp). */
static void lk_level(const orc_level *I, const orc_level *J, int cn, const float *prev_pts, float *next_pts,
                     uint8_t *status, float *err, int npts, int win, int max_count, double eps_sq, int level,
                     int max_level, int use_initial_flow, double min_eig_thresh) {
  const float half = (win - 1) * 0.5f;
  const int W_BITS = 14;
  const float FLT_SCALE = 1.f / (1 << 20);
  int16_t *Iw = (int16_t *)malloc(sizeof(int16_t) * (size_t)win * win * cn * 3);
  int16_t *dIw = Iw + (size_t)win * win * cn;
  for (int p = 0; p < npts; ++p) {
    float px = prev_pts[2 * p] * (float)(1. / (1 << level));
    float py = prev_pts[2 * p + 1] * (float)(1. / (1 << level));
    float nx, ny;
    if (level == max_level) {
      if (use_initial_flow) {
        nx = next_pts[2 * p] * (float)(1. / (1 << level));
        ny = next_pts[2 * p + 1] * (float)(1. / (1 << level));
      } else {
        nx = px;
        ny = py;
      }
    } else {
      nx = next_pts[2 * p] * 2.f;
      ny = next_pts[2 * p + 1] * 2.f;
    }
    next_pts[2 * p] = nx;
    next_pts[2 * p + 1] = ny;

    px -= half;
    py -= half;
    int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= I->cols || ipy < -win || ipy >= I->rows) {
      if (level == 0) {
        status[p] = 0;
        if (err) err[p] = 0;
      }
      continue;
    }
    float a = px - ipx, b = py - ipy;
    int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
    int iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
    int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    double sA11 = 0, sA12 = 0, sA22 = 0; /* exact integer sums (OpenCV: float SIMD lanes) */
    for (int y = 0; y < win; ++y)
      for (int x = 0; x < win; ++x)
        for (int c = 0; c < cn; ++c) {
          int X = ipx + x, Y = ipy + y;
          int ival = descale(pix(I, cn, X, Y, c) * iw00 + pix(I, cn, X + 1, Y, c) * iw01 +
                                 pix(I, cn, X, Y + 1, c) * iw10 + pix(I, cn, X + 1, Y + 1, c) * iw11,
                             W_BITS - 5);
          int ix = descale(der(I, cn, X, Y, c, 0) * iw00 + der(I, cn, X + 1, Y, c, 0) * iw01 +
                               der(I, cn, X, Y + 1, c, 0) * iw10 + der(I, cn, X + 1, Y + 1, c, 0) * iw11,
                           W_BITS);
          int iy = descale(der(I, cn, X, Y, c, 1) * iw00 + der(I, cn, X + 1, Y, c, 1) * iw01 +
                               der(I, cn, X, Y + 1, c, 1) * iw10 + der(I, cn, X + 1, Y + 1, c, 1) * iw11,
                           W_BITS);
          size_t o = ((size_t)y * win + x) * cn + c;
          Iw[o] = (int16_t)ival;
          dIw[2 * o] = (int16_t)ix;
          dIw[2 * o + 1] = (int16_t)iy;
          sA11 += (double)ix * ix;
          sA12 += (double)ix * iy;
          sA22 += (double)iy * iy;
        }
    float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
    if ((double)minEig < min_eig_thresh || D < FLT_EPSILON) {
      if (level == 0) status[p] = 0;
      continue;
    }
    D = 1.f / D;
    nx -= half;
    ny -= half;
    float pdx = 0, pdy = 0;
    for (int j = 0; j < max_count; ++j) {
      int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -win || inx >= J->cols || iny < -win || iny >= J->rows) {
        if (level == 0) status[p] = 0;
        break;
      }
      a = nx - inx;
      b = ny - iny;
      iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
      iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
      iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
      iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      double sb1 = 0, sb2 = 0;
      for (int y = 0; y < win; ++y)
        for (int x = 0; x < win; ++x)
          for (int c = 0; c < cn; ++c) {
            int X = inx + x, Y = iny + y;
            size_t o = ((size_t)y * win + x) * cn + c;
            int diff = descale(pix(J, cn, X, Y, c) * iw00 + pix(J, cn, X + 1, Y, c) * iw01 +
                                   pix(J, cn, X, Y + 1, c) * iw10 + pix(J, cn, X + 1, Y + 1, c) * iw11,
                               W_BITS - 5) -
                       Iw[o];
            sb1 += (double)diff * dIw[2 * o];
            sb2 += (double)diff * dIw[2 * o + 1];
          }
      float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
      float dx = (float)((A12 * b2 - A22 * b1) * D);
      float dy = (float)((A12 * b1 - A11 * b2) * D);
      nx += dx;
      ny += dy;
      next_pts[2 * p] = nx + half;
      next_pts[2 * p + 1] = ny + half;
      if ((double)dx * dx + (double)dy * dy <= eps_sq) break;
      if (j > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
        next_pts[2 * p] -= dx * 0.5f;
        next_pts[2 * p + 1] -= dy * 0.5f;
        break;
      }
      pdx = dx;
      pdy = dy;
    }
    if (status[p] && err && level == 0) {
      float fx = next_pts[2 * p] - half, fy = next_pts[2 * p + 1] - half;
      int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -win || inx >= J->cols || iny < -win || iny >= J->rows) {
        status[p] = 0;
        continue;
      }
      float aa = fx - inx, bb = fy - iny;
      iw00 = (int)lrintf((1.f - aa) * (1.f - bb) * (1 << W_BITS));
      iw01 = (int)lrintf(aa * (1.f - bb) * (1 << W_BITS));
      iw10 = (int)lrintf((1.f - aa) * bb * (1 << W_BITS));
      iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      float errval = 0.f;
      for (int y = 0; y < win; ++y)
        for (int x = 0; x < win; ++x)
          for (int c = 0; c < cn; ++c) {
            int X = inx + x, Y = iny + y;
            size_t o = ((size_t)y * win + x) * cn + c;
            int diff = descale(pix(J, cn, X, Y, c) * iw00 + pix(J, cn, X + 1, Y, c) * iw01 +
                                   pix(J, cn, X, Y + 1, c) * iw10 + pix(J, cn, X + 1, Y + 1, c) * iw11,
                               W_BITS - 5) -
                       Iw[o];
            errval += (float)abs(diff);
          }
      err[p] = errval * 1.f / (32 * win * cn * win);
    }
  }
  free(Iw);
}

/* calcOpticalFlowPyrLK(prev, next, prevPts, nextPts(in/out), status, err, (win,win),
 * maxLevel, criteria(COUNT|EPS, max_count, eps), flags = use_initial_flow ?
 * OPTFLOW_USE_INITIAL_FLOW : 0, minEigThreshold).  Images are rows x cols x cn u8. */
int orc_lk_track(const uint8_t *prev, const uint8_t *next, int rows, int cols, int cn, const float *prev_pts,
                 float *next_pts, uint8_t *status, float *err, int npts, int win, int max_level, int max_count,
                 double eps, int use_initial_flow, double min_eig_thresh) {
  int L = orc_pyramid_levels(rows, cols, win, max_level);
  if (max_count < 0) max_count = 0;
  if (max_count > 100) max_count = 100;
  if (eps < 0) eps = 0;
  if (eps > 10) eps = 10;
  double eps_sq = eps * eps;
  uint8_t *pi[16], *pj[16];
  int16_t *di[16];
  int rr[16], cc[16];
  rr[0] = rows;
  cc[0] = cols;
  pi[0] = (uint8_t *)prev;
  pj[0] = (uint8_t *)next;
  for (int l = 1; l <= L; ++l) {
    rr[l] = (rr[l - 1] + 1) / 2;
    cc[l] = (cc[l - 1] + 1) / 2;
    pi[l] = (uint8_t *)malloc((size_t)rr[l] * cc[l] * cn);
    pj[l] = (uint8_t *)malloc((size_t)rr[l] * cc[l] * cn);
    orc_pyrdown(pi[l - 1], rr[l - 1], cc[l - 1], cn, pi[l]);
    orc_pyrdown(pj[l - 1], rr[l - 1], cc[l - 1], cn, pj[l]);
  }
  for (int l = 0; l <= L; ++l) {
    di[l] = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)rr[l] * cc[l] * cn);
    orc_scharr(pi[l], rr[l], cc[l], cn, di[l]);
  }
  for (int p = 0; p < npts; ++p) status[p] = 1;
  for (int l = L; l >= 0; --l) {
    orc_level I = {rr[l], cc[l], pi[l], di[l]}, J = {rr[l], cc[l], pj[l], NULL};
    lk_level(&I, &J, cn, prev_pts, next_pts, status, err, npts, win, max_count, eps_sq, l, L, use_initial_flow,
             min_eig_thresh);
  }
  for (int l = 0; l <= L; ++l) free(di[l]);
  for (int l = 1; l <= L; ++l) {
    free(pi[l]);
    free(pj[l]);
  }
  return L;
}

/* ------------------------------------------------------------------------------------------------------------------
 * BRIEF-32 descriptors, Hamming distance and the cross-checked brute-force matcher of the descriptor path
 * (/root/reference/src/tracker.cpp:231-292 DetectLK rescue, :341-460 UpdateMatch, :530-565 descriptor check in UpdateLK;
 * popcount distance: /root/reference/src/fastbrief.cpp:53-93).  The arithmetic lives in OpenCV (features2d BFMatcher) and
 * opencv_contrib (xfeatures2d BriefDescriptorExtractor), neither vendored in the reference.
 *   BRIEF as in opencv_contrib: grey image, keypoints closer than 28 px (= 48 / 2 + 9 / 2) to the border are dropped, the pixel is
 *   (int)(pt + 0.5), a test compares two 9 x 9 box sums (OpenCV reads them from an integral image: the same integers), test 8 i + k
 *   sets bit 7 - k of byte i.  The 256 test pairs are an OWN table (xivo_b200/csrc/brief_pattern.h, scripts/make_brief_pattern.py):
 *   opencv_contrib's generated table is not available here, so descriptor VALUES are parity-unpinned; the matcher is pinned on cv2.
 *   BFMatcher(NORM_HAMMING, crossCheck = true).knnMatch(query, train, 1, noArray(), compactResult = true): the nearest train
 *   descriptor of every query (first index on ties), kept when that train's nearest query (first index on ties) is the query. */
#include "../xivo_b200/csrc/brief_pattern.h"

static inline int box9(const uint8_t *g, int cols, int x, int y) {
  int s = 0;
  for (int dy = -4; dy <= 4; ++dy)
    for (int dx = -4; dx <= 4; ++dx) s += g[(y + dy) * cols + (x + dx)];
  return s;
}

/* img: grey rows x cols.  kp_xy: n x 2 floats.  desc: n x 32 bytes (zero for dropped keypoints), valid: n flags. */
void orc_brief(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, uint8_t *desc, uint8_t *valid) {
  for (int i = 0; i < n; ++i) {
    const float fx = kp_xy[2 * i], fy = kp_xy[2 * i + 1];
    uint8_t *d = desc + 32 * (size_t)i;
    memset(d, 0, 32);
    /* KeyPointsFilter::runByImageBorder: Rect(border, border, cols - 2 border, rows - 2 border).contains(pt) on the float point */
    valid[i] = (fx >= XB_BRIEF_BORDER && fx < cols - XB_BRIEF_BORDER && fy >= XB_BRIEF_BORDER && fy < rows - XB_BRIEF_BORDER) ? 1 : 0;
    if (!valid[i]) continue;
    const int cx = (int)(fx + 0.5f), cy = (int)(fy + 0.5f);
    for (int t = 0; t < 256; ++t) {
      const signed char *p = kBriefPattern[t];
      const int a = box9(img, cols, cx + p[0], cy + p[1]), b = box9(img, cols, cx + p[2], cy + p[3]);
      if (a < b) d[t >> 3] |= (uint8_t)(1u << (7 - (t & 7)));
    }
  }
}

int orc_hamming(const uint8_t *a, const uint8_t *b, int bytes) {
  int d = 0;
  for (int i = 0; i < bytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

/* Cross-checked 1-nearest-neighbour matches: out (query, train, distance) triples in query order; returns their number. */
int orc_bf_match_crosscheck(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes, int *out3) {
  int m = 0;
  if (nq <= 0 || nt <= 0) return 0;
  int *best_q_of_t = (int *)malloc(sizeof(int) * (size_t)nt), *best_d_of_t = (int *)malloc(sizeof(int) * (size_t)nt);
  for (int j = 0; j < nt; ++j) { best_q_of_t[j] = -1; best_d_of_t[j] = 1 << 30; }
  int *best_t_of_q = (int *)malloc(sizeof(int) * (size_t)nq), *best_d_of_q = (int *)malloc(sizeof(int) * (size_t)nq);
  for (int i = 0; i < nq; ++i) {
    int bt = -1, bd = 1 << 30;
    for (int j = 0; j < nt; ++j) {
      const int d = orc_hamming(q + (size_t)i * bytes, t + (size_t)j * bytes, bytes);
      if (d < bd) { bd = d; bt = j; }
      if (d < best_d_of_t[j]) { best_d_of_t[j] = d; best_q_of_t[j] = i; }
    }
    best_t_of_q[i] = bt; best_d_of_q[i] = bd;
  }
  for (int i = 0; i < nq; ++i)
    if (best_t_of_q[i] >= 0 && best_q_of_t[best_t_of_q[i]] == i) { out3[3 * m] = i; out3[3 * m + 1] = best_t_of_q[i]; out3[3 * m + 2] = best_d_of_q[i]; ++m; }
  free(best_q_of_t); free(best_d_of_t); free(best_t_of_q); free(best_d_of_q);
  return m;
}
