// TEST INFRASTRUCTURE — stand-in for the OpenCV types the reference's HEADERS mention (OpenCV C++ is not in this
// image); see ../Eigen/Core for why this exists.  Nothing here computes: every operation aborts if it is called.
#pragma once
#include <cstddef>
#include <cstdlib>
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
[[noreturn]] inline void stub_abort() { std::abort(); }
struct Size { int width = 0, height = 0; Size() = default; template <class A, class B> Size(A w, B h) : width(int(w)), height(int(h)) {} };
struct Vec3f { float v[3]; Vec3f(float a = 0, float b = 0, float c = 0) : v{a, b, c} {} };
struct MatSize { int dims = 2; int p[2] = {0, 0}; int operator[](int i) const { return p[i]; } };
struct Mat {
    int rows = 0, cols = 0; unsigned char* data = nullptr; size_t step = 0; MatSize size;
    Mat() = default;
    template <class... A> Mat(A&&...) {}
    int channels() const { return 1; } int type() const { return 0; } bool empty() const { return true; }
    Mat clone() const { return Mat(); }
    template <class... A> static Mat eye(A&&...) { return Mat(); }
    template <class... A> static Mat zeros(A&&...) { return Mat(); }
    template <class... A> void convertTo(A&&...) const { stub_abort(); }
    template <class T> T* ptr(int = 0) { return nullptr; }
    template <class T> T& at(int, int = 0) { stub_abort(); }
};
template <class T> struct Mat_ : Mat {
    template <class... A> Mat_(A&&...) {}
    struct Init { Mat_* m; Init operator,(T) { return *this; } operator Mat() const { return Mat(); } };
    Init operator<<(T) { return Init{this}; }
};
using InputArray = const Mat&; using OutputArray = Mat&; using InputOutputArray = Mat&;
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
template <class... A> void initUndistortRectifyMap(A&&...) { stub_abort(); }
template <class... A> void remap(A&&...) { stub_abort(); }
template <class... A> void resize(A&&...) { stub_abort(); }
template <class... A> void cvtColor(A&&...) { stub_abort(); }
template <class... A> Mat imread(A&&...) { stub_abort(); }
template <class... A> bool imwrite(A&&...) { stub_abort(); }
namespace cuda {
struct GpuMat {
    int rows = 0, cols = 0; unsigned char* data = nullptr; size_t step = 0;
    GpuMat() = default;
    template <class... A> GpuMat(A&&...) {}
    int channels() const { return 1; } int type() const { return 0; } bool empty() const { return true; }
    GpuMat clone() const { return GpuMat(); }
    template <class... A> void upload(A&&...) { stub_abort(); }
    template <class... A> void download(A&&...) const { stub_abort(); }
    template <class... A> void convertTo(A&&...) const { stub_abort(); }
};
template <class... A> void resize(A&&...) { stub_abort(); }
template <class... A> void remap(A&&...) { stub_abort(); }
template <class... A> void cvtColor(A&&...) { stub_abort(); }
}  // namespace cuda
}  // namespace cv
