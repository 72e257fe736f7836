// A pure C++ consumer of the C ABI (include/segs_raster.h): no Python, no LibTorch — cudaMalloc'd buffers, two lanes,
// segs_raster_views.  This is the shape of the call a C++ host such as SEGS-SLAM's GaussianMapper makes when it renders
// and back-propagates a batch of keyframe views over explicit Gaussians.
//
//   raster_views_capi <scene.bin> <out.bin> [lanes]
//
// scene.bin (little endian; written by tests/test_capi_cpp_gpu.py):
//   int32 P, W, H, n_views; float32 tan_fovx, tan_fovy;
//   float32 means3D[P*3], colors[P*3], opacities[P], scales[P*3], rotations[P*4], background[3];
//   per view: float32 viewmatrix[16], projmatrix[16], campos[3], dL_dout[3*H*W]
// out.bin: per view int32 num_rendered, float32 image[3*H*W]; then float32 grads[P*17] in the order
//   means3D 3 | means2D 3 | colors 3 | opacity 1 | scales 3 | rotations 4 (one array after the other)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/segs_raster.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

template <typename T>
static bool rd(std::FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

static float* to_device(const std::vector<float>& h) {
    float* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(float) + 16) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s scene.bin out.bin [lanes]\n", argv[0]); return 1; }
    const int lanes = argc > 3 ? std::atoi(argv[3]) : 2;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 1; }
    int hdr[4]; float tanf[2];
    if (!rd(f, hdr, 4) || !rd(f, tanf, 2)) return 1;
    const int P = hdr[0], W = hdr[1], H = hdr[2], NV = hdr[3];
    const size_t N = size_t(W) * H;
    std::vector<float> means(size_t(P) * 3), colors(size_t(P) * 3), opac(P), scales(size_t(P) * 3), rots(size_t(P) * 4), bg(3);
    if (!rd(f, means.data(), means.size()) || !rd(f, colors.data(), colors.size()) || !rd(f, opac.data(), opac.size()) ||
        !rd(f, scales.data(), scales.size()) || !rd(f, rots.data(), rots.size()) || !rd(f, bg.data(), 3)) return 1;
    float *d_means = to_device(means), *d_colors = to_device(colors), *d_opac = to_device(opac), *d_scales = to_device(scales),
          *d_rots = to_device(rots), *d_bg = to_device(bg);

    // gradient accumulators: one zeroed block, six slices
    const size_t widths[6] = {3, 3, 3, 1, 3, 4};
    float* d_grads = nullptr;
    CK(cudaMalloc(&d_grads, size_t(P) * 17 * sizeof(float) + 16));
    CK(cudaMemset(d_grads, 0, size_t(P) * 17 * sizeof(float)));
    float* slice[6];
    { size_t off = 0; for (int k = 0; k < 6; ++k) { slice[k] = d_grads + off; off += widths[k] * P; } }

    std::vector<segs_raster_view_args> args(NV);
    std::vector<segs_mapper_view_result> res(NV);
    std::vector<float*> d_img(NV);
    for (int v = 0; v < NV; ++v) {
        std::vector<float> view(16), proj(16), cam(3), dL(3 * N);
        if (!rd(f, view.data(), 16) || !rd(f, proj.data(), 16) || !rd(f, cam.data(), 3) || !rd(f, dL.data(), dL.size())) return 1;
        CK(cudaMalloc(&d_img[v], 3 * N * sizeof(float)));
        segs_raster_view_args a{};
        a.P = P; a.means3D = d_means; a.colors_precomp = d_colors; a.opacities = d_opac; a.scales = d_scales; a.rotations = d_rots;
        a.background = d_bg; a.width = W; a.height = H; a.tan_fovx = tanf[0]; a.tan_fovy = tanf[1];
        a.viewmatrix = to_device(view); a.projmatrix = to_device(proj); a.campos = to_device(cam); a.dL_dout = to_device(dL);
        a.image_out = d_img[v];
        a.grad_means3D = slice[0]; a.grad_means2D = slice[1]; a.grad_colors = slice[2]; a.grad_opacity = slice[3];
        a.grad_scales = slice[4]; a.grad_rotations = slice[5];
        args[v] = a;
    }
    std::fclose(f);

    std::vector<segs_workspace*> ws(lanes);
    std::vector<void*> streams(lanes);
    for (int l = 0; l < lanes; ++l) {
        if (segs_workspace_create(&ws[l]) != SEGS_OK) return 2;
        cudaStream_t s; CK(cudaStreamCreate(&s)); streams[l] = s;
    }
    cudaStream_t main_stream; CK(cudaStreamCreate(&main_stream));
    const int rc = segs_raster_views(NV, args.data(), res.data(), lanes, ws.data(), streams.data(), main_stream);
    if (rc != SEGS_OK) { std::fprintf(stderr, "segs_raster_views: %s\n", segs_last_error()); return 3; }
    CK(cudaStreamSynchronize(main_stream));

    std::FILE* o = std::fopen(argv[2], "wb");
    if (!o) { std::perror(argv[2]); return 1; }
    std::vector<float> img(3 * N);
    for (int v = 0; v < NV; ++v) {
        CK(cudaMemcpy(img.data(), d_img[v], img.size() * sizeof(float), cudaMemcpyDeviceToHost));
        std::fwrite(&res[v].num_rendered, sizeof(int), 1, o);
        std::fwrite(img.data(), sizeof(float), img.size(), o);
    }
    std::vector<float> grads(size_t(P) * 17);
    CK(cudaMemcpy(grads.data(), d_grads, grads.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::fwrite(grads.data(), sizeof(float), grads.size(), o);
    std::fclose(o);
    for (int l = 0; l < lanes; ++l) segs_workspace_destroy(ws[l]);
    std::printf("ok views=%d P=%d lanes=%d launches=%llu workspace0=%zu\n", NV, P, lanes, segs_launch_count(), size_t(0));
    return 0;
}
