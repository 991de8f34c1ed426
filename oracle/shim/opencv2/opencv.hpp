// oracle/shim/opencv2/opencv.hpp — a stand-in for the part of OpenCV that the reference's point-selection
// sources touch (see oracle/shim/Eigen/Core for the purpose).  cv::Mat is a plain 8UC1 / 8UC3 / 16UC1
// buffer; cv::cvtColor implements the two conversions pcd_generator::load_image calls (RGB2GRAY, RGB2HSV)
// with OpenCV's 8-bit integer algorithms as pinned against cv2 4.13 in tests/test_oracle_pins.py (the same
// functions the oracle uses: oracle_gray_u8 / oracle_hsv_u8 below are provided by oracle/ref_select.cpp).
// The visualisation calls of pcd_generator.cpp (imshow, applyColorMap ...) are never reached on the
// selection path and are empty here.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_16UC1 2

typedef unsigned char uchar;
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvRound(double v) { return (int)std::lrint(v); }

extern "C" void oracle_gray_u8(const uint8_t *bgr, int n, uint8_t *out);   // RGB2GRAY applied to the stored channel order
extern "C" void oracle_hsv_u8(const uint8_t *bgr, int n, uint8_t *out);    // RGB2HSV, 8-bit, H in [0,180)

namespace cv {
struct Point { int x, y; Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {} };
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Vec3b { uchar val[3]; };
struct KeyPoint { Point2f pt; };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
enum { COLOR_RGB2GRAY = 7, COLOR_RGB2HSV = 41, COLORMAP_JET = 2 };

class Mat {
    std::shared_ptr<std::vector<uchar>> own_;
public:
    int rows = 0, cols = 0, type_ = CV_8UC1;
    uchar *data = nullptr;
    size_t step = 0;   // bytes per row
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext, size_t step_bytes) : rows(r), cols(c), type_(type), data((uchar *)ext), step(step_bytes) {}
    static int elem(int type) { return type == CV_8UC3 ? 3 : type == CV_16UC1 ? 2 : 1; }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step = (size_t)c * elem(type);
        own_ = std::make_shared<std::vector<uchar>>((size_t)r * step);
        data = own_->data();
    }
    void copyTo(Mat &o) const {
        o.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(o.data + (size_t)y * o.step, data + (size_t)y * step, o.step);
    }
    template <class T> T &at(const Point &p) { return *reinterpret_cast<T *>(data + (size_t)p.y * step + (size_t)p.x * sizeof(T)); }
    template <class T> const T &at(const Point &p) const { return *reinterpret_cast<const T *>(data + (size_t)p.y * step + (size_t)p.x * sizeof(T)); }
};

inline void cvtColor(const Mat &src, Mat &dst, int code) {
    if (code == COLOR_RGB2GRAY) {
        dst.create(src.rows, src.cols, CV_8UC1);
        for (int y = 0; y < src.rows; y++) oracle_gray_u8(src.data + (size_t)y * src.step, src.cols, dst.data + (size_t)y * dst.step);
    } else {
        dst.create(src.rows, src.cols, CV_8UC3);
        for (int y = 0; y < src.rows; y++) oracle_hsv_u8(src.data + (size_t)y * src.step, src.cols, dst.data + (size_t)y * dst.step);
    }
}
// "Key: value" scalars of an OpenCV YAML (cvo.cpp:58-64 reads five)
struct FileNode {
    double v = 0;
    operator float() const { return (float)v; }
    operator double() const { return v; }
    operator int() const { return (int)v; }
};
class FileStorage {
    std::string path_;
public:
    enum { READ = 0 };
    FileStorage(const std::string &p, int) : path_(p) {}
    FileNode operator[](const char *key) const {
        FileNode n;
        FILE *f = fopen(path_.c_str(), "r");
        if (!f) return n;
        char line[512];
        const std::string k = std::string(key) + ":";
        while (fgets(line, sizeof(line), f)) {
            const char *p = strstr(line, k.c_str());
            if (p) { n.v = atof(p + k.size()); break; }
        }
        fclose(f);
        return n;
    }
};
// never reached on the selection path (visualize_selected_pixels is commented out of create_pointcloud)
inline void minMaxLoc(const Mat &, double *mn, double *mx) { if (mn) *mn = 0; if (mx) *mx = 0; }
inline void applyColorMap(const Mat &, Mat &, int) {}
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
