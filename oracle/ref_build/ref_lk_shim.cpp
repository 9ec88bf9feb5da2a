// TEST INFRASTRUCTURE ONLY (oracle/_ref): thin C shim around the *reference's own* vendored OpenCV
// functions so that tests/ and bench.py's cpu_baseline leg can call them through ctypes.
// Wraps cv::buildOpticalFlowPyramid (OCV/video/src/lkpyramid.cpp:726-822) and
// cv::calcOpticalFlowPyrLK (lkpyramid.cpp:1408-1417), called exactly the way HybVIO calls them
// (src/tracker/image_pyramid.cpp:40-48, src/tracker/optical_flow.cpp:33-49).
#include <opencv2/core.hpp>
#include <opencv2/core/utility.hpp>
#include <opencv2/video/tracking.hpp>
#include <cstring>
#include <vector>

struct RefPyr { std::vector<cv::Mat> pyr; int win; };

extern "C" {

int ref_get_num_threads() { return cv::getNumThreads(); }
void ref_set_num_threads(int n) { cv::setNumThreads(n); }

void* ref_pyr_build(const unsigned char* img, int w, int h, int stride, int win, int maxLevel) {
    RefPyr* p = new RefPyr; p->win = win;
    cv::Mat m(h, w, CV_8UC1, const_cast<unsigned char*>(img), (size_t)stride);
    cv::buildOpticalFlowPyramid(m, p->pyr, cv::Size(win, win), maxLevel);
    return p;
}
// rebuild into an existing handle (buffer reuse like util::Allocator recycling)
void ref_pyr_rebuild(void* hp, const unsigned char* img, int w, int h, int stride, int maxLevel) {
    RefPyr* p = (RefPyr*)hp;
    cv::Mat m(h, w, CV_8UC1, const_cast<unsigned char*>(img), (size_t)stride);
    cv::buildOpticalFlowPyramid(m, p->pyr, cv::Size(p->win, p->win), maxLevel);
}
int ref_pyr_levels(void* hp) { return (int)((RefPyr*)hp)->pyr.size() / 2; }
void ref_pyr_level_size(void* hp, int level, int* w, int* h) {
    const cv::Mat& g = ((RefPyr*)hp)->pyr[level * 2]; *w = g.cols; *h = g.rows;
}
// Copies the level *with* its win-pixel padding: gray (h+2win)x(w+2win) u8, deriv same dims x2 s16.
void ref_pyr_get_level_padded(void* hp, int level, unsigned char* gray, short* deriv) {
    RefPyr* p = (RefPyr*)hp; int win = p->win;
    cv::Mat g = p->pyr[level * 2], d = p->pyr[level * 2 + 1];
    int W = g.cols + 2 * win, H = g.rows + 2 * win;
    g.adjustROI(win, win, win, win); d.adjustROI(win, win, win, win);
    CV_Assert(g.cols == W && g.rows == H && d.cols == W && d.rows == H);
    for (int y = 0; y < H; y++) {
        if (gray) std::memcpy(gray + (size_t)y * W, g.ptr(y), (size_t)W);
        if (deriv) std::memcpy(deriv + (size_t)y * W * 2, d.ptr(y), (size_t)W * 4);
    }
}
void ref_pyr_free(void* hp) { delete (RefPyr*)hp; }

// nextPts: in = initial guesses when useInitial, out = result. status: 0/1 (OpenCV's uchar status).
// err is requested (HybVIO passes an err vector) so the level-0 post-iteration bounds check runs.
int ref_lk(void* prevHp, void* nextHp, const float* prevPts, float* nextPts, unsigned char* status,
           int n, int win, int maxLevel, int maxIter, double eps, int useInitial, double minEig) {
    RefPyr* a = (RefPyr*)prevHp; RefPyr* b = (RefPyr*)nextHp;
    if (n == 0) return 0;
    std::vector<cv::Point2f> pp(n), np(n);
    std::memcpy(pp.data(), prevPts, sizeof(float) * 2 * n);
    std::memcpy(np.data(), nextPts, sizeof(float) * 2 * n);
    std::vector<unsigned char> st; std::vector<float> err;
    cv::TermCriteria crit(cv::TermCriteria::COUNT | cv::TermCriteria::EPS, maxIter, eps);
    cv::calcOpticalFlowPyrLK(a->pyr, b->pyr, pp, np, st, err, cv::Size(win, win), maxLevel, crit,
                             useInitial ? cv::OPTFLOW_USE_INITIAL_FLOW : 0, minEig);
    std::memcpy(nextPts, np.data(), sizeof(float) * 2 * n);
    std::memcpy(status, st.data(), n);
    return 0;
}
}
