// TEST INFRASTRUCTURE — writes / reads an anchor PLY with the reference's OWN PLY library
// (/root/reference/third_party/tinyply/tinyply.h, compiled where it lies; -I in oracle/Makefile), issuing exactly the
// add_properties_to_element sequence of GaussianModel::savePly (src/gaussian_model.cpp:1179-1256) and the
// request_properties_from_element calls of loadPly (:1060-1113, with the names savePly writes).  It pins
// segs_slam_b200/checkpoint.py byte for byte (tests/golden/make_ply_golden.py -> tests/golden/anchors_tinyply.ply).
#define TINYPLY_IMPLEMENTATION
#include "tinyply.h"

#include <cstring>
#include <fstream>
#include <string>
#include <vector>

static std::vector<std::string> numbered(const std::string& p, int n) {
    std::vector<std::string> v(n);
    for (int i = 0; i < n; ++i) v[i] = p + std::to_string(i);
    return v;
}

extern "C" {

// all arrays row-major float32: anchor [A,3], feat [A,F], offset_flat [A,3k] (already transposed/flattened as savePly does),
// opacity [A,1], scale [A,6], rot [A,4]
int ref_save_ply(const char* path, int A, int F, int k3, float* anchor, float* feat, float* offset_flat, float* opacity,
                 float* scale, float* rot)
{
    std::vector<float> normals(size_t(A) * 3, 0.f);
    std::filebuf fb;
    fb.open(path, std::ios::out | std::ios::binary);
    std::ostream os(&fb);
    if (os.fail()) return 1;
    tinyply::PlyFile file;
    auto add = [&](const std::vector<std::string>& names, float* data) {
        file.add_properties_to_element("vertex", names, tinyply::Type::FLOAT32, A, reinterpret_cast<uint8_t*>(data),
                                       tinyply::Type::INVALID, 0);
    };
    add({"x", "y", "z"}, anchor);
    add({"nx", "ny", "nz"}, normals.data());
    add(numbered("anchor_feat_", F), feat);
    add(numbered("offset_", k3), offset_flat);
    add({"opacity"}, opacity);
    add(numbered("scale_", 6), scale);
    add(numbered("rot_", 4), rot);
    file.write(os, true);
    fb.close();
    return 0;
}

// GaussianModel::saveSparsePointsPly's call sequence (src/gaussian_model.cpp:1319-1352): xyz / normals float, rgb uchar
int ref_save_sparse_ply(const char* path, int n, float* xyz, unsigned char* rgb)
{
    std::vector<float> normals(size_t(n) * 3, 0.f);
    std::filebuf fb;
    fb.open(path, std::ios::out | std::ios::binary);
    std::ostream os(&fb);
    if (os.fail()) return 1;
    tinyply::PlyFile file;
    file.add_properties_to_element("vertex", {"x", "y", "z"}, tinyply::Type::FLOAT32, n, reinterpret_cast<uint8_t*>(xyz),
                                   tinyply::Type::INVALID, 0);
    file.add_properties_to_element("vertex", {"nx", "ny", "nz"}, tinyply::Type::FLOAT32, n,
                                   reinterpret_cast<uint8_t*>(normals.data()), tinyply::Type::INVALID, 0);
    file.add_properties_to_element("vertex", {"red", "green", "blue"}, tinyply::Type::UINT8, n, rgb, tinyply::Type::INVALID, 0);
    file.write(os, true);
    fb.close();
    return 0;
}

// reads back with tinyply; returns the anchor count, fills the arrays (same layouts as above)
int ref_load_ply(const char* path, int F, int k3, float* anchor, float* feat, float* offset_flat, float* opacity, float* scale,
                 float* rot)
{
    std::ifstream is(path, std::ios::binary);
    if (is.fail()) return -1;
    tinyply::PlyFile file;
    file.parse_header(is);
    auto a = file.request_properties_from_element("vertex", {"x", "y", "z"});
    auto f = file.request_properties_from_element("vertex", numbered("anchor_feat_", F));
    auto o = file.request_properties_from_element("vertex", numbered("offset_", k3));
    auto op = file.request_properties_from_element("vertex", {"opacity"});
    auto s = file.request_properties_from_element("vertex", numbered("scale_", 6));
    auto r = file.request_properties_from_element("vertex", numbered("rot_", 4));
    file.read(is);
    std::memcpy(anchor, a->buffer.get(), a->buffer.size_bytes());
    std::memcpy(feat, f->buffer.get(), f->buffer.size_bytes());
    std::memcpy(offset_flat, o->buffer.get(), o->buffer.size_bytes());
    std::memcpy(opacity, op->buffer.get(), op->buffer.size_bytes());
    std::memcpy(scale, s->buffer.get(), s->buffer.size_bytes());
    std::memcpy(rot, r->buffer.get(), r->buffer.size_bytes());
    return int(a->count);
}

}  // extern "C"
