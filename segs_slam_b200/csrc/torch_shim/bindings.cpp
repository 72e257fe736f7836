// pybind11 surface of the LibTorch host layer, so the Python test-suite can drive the kernels
// through the SAME C++ entry points a SEGS-SLAM build links (rasterize_points.h) and through the
// C++ autograd function (gaussian_rasterizer.h).  Built in-tree as segs_slam_b200/_segs_torch.so.
#include <torch/extension.h>

#include "densify.h"
#include "fused_mapper.h"
#include "gaussian_rasterizer.h"
#include "keyframe_transforms.h"
#include "loss_utils.h"
#include "rasterize_points.h"

namespace {

// GaussianRasterizer::forward of the L5 twin with the has_* flags derived from numel(), returning
// (color, radii); autograd flows through GaussianRasterizerFunction.
std::tuple<torch::Tensor, torch::Tensor> rasterizer_forward(
    int H, int W, double tanfovx, double tanfovy, torch::Tensor bg, double scale_modifier, torch::Tensor viewmatrix,
    torch::Tensor projmatrix, int sh_degree, torch::Tensor campos, bool prefiltered, torch::Tensor means3D,
    torch::Tensor means2D, torch::Tensor opacities, torch::Tensor shs, torch::Tensor colors_precomp,
    torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp)
{
    GaussianRasterizationSettings settings(H, W, static_cast<float>(tanfovx), static_cast<float>(tanfovy), bg,
                                           static_cast<float>(scale_modifier), viewmatrix, projmatrix, sh_degree,
                                           campos, prefiltered);
    GaussianRasterizer rasterizer(settings);
    return rasterizer.forward(means3D, means2D, opacities, shs.numel() != 0, colors_precomp.numel() != 0,
                              scales.numel() != 0, rotations.numel() != 0, cov3D_precomp.numel() != 0, shs,
                              colors_precomp, scales, rotations, cov3D_precomp);
}

torch::Tensor mark_visible(torch::Tensor means3D, torch::Tensor viewmatrix, torch::Tensor projmatrix) {
    return markVisible(means3D, viewmatrix, projmatrix);
}

torch::Tensor lu_l1_loss(torch::Tensor a, torch::Tensor b) { return loss_utils::l1_loss(a, b); }
torch::Tensor lu_ssim(torch::Tensor a, torch::Tensor b) { return loss_utils::ssim(a, b); }
torch::Tensor lu_psnr(torch::Tensor a, torch::Tensor b) { return loss_utils::psnr(a, b); }
torch::Tensor lu_l1_ssim(torch::Tensor image, torch::Tensor gt, double lambda_dssim, torch::Tensor row_mask) {
    return loss_utils::l1_ssim(image, gt, lambda_dssim, row_mask);
}
void lu_adam_step(std::vector<torch::Tensor> params, std::vector<double> lrs, torch::Tensor grad, torch::Tensor m,
                  torch::Tensor v, int64_t step, double beta1, double beta2, double eps, double weight_decay,
                  double grad_scale, bool zero_grad) {
    loss_utils::adam_step(params, lrs, grad, m, v, step, beta1, beta2, eps, weight_decay, grad_scale, zero_grad);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> kf_transforms(std::vector<double> R, std::vector<double> t,
                                                                                      double FoVx, double FoVy) {
    TORCH_CHECK(R.size() == 9 && t.size() == 3, "R: 9 row-major floats, t: 3 floats");
    std::array<float, 9> Ra; std::array<float, 3> ta;
    for (int i = 0; i < 9; ++i) Ra[i] = static_cast<float>(R[i]);
    for (int i = 0; i < 3; ++i) ta[i] = static_cast<float>(t[i]);
    return keyframe_transforms::computeTransformTensors(Ra, ta, static_cast<float>(FoVx), static_cast<float>(FoVy));
}
std::vector<double> kf_quat_to_rot(double w, double x, double y, double z) {
    auto r = keyframe_transforms::quaternionToRotation(w, x, y, z);
    return std::vector<double>(r.begin(), r.end());
}

// views: list of (world_view_transform, full_proj_transform, camera_center, pose[7], gt_image, row_mask-or-empty)
torch::Tensor fm_render_views(FusedMapper& fm, const std::vector<std::tuple<torch::Tensor, torch::Tensor, torch::Tensor,
                                                                             std::vector<double>, torch::Tensor, torch::Tensor>>& views) {
    std::vector<KeyframeView> kv(views.size());
    for (size_t i = 0; i < views.size(); ++i) {
        kv[i].world_view_transform = std::get<0>(views[i]);
        kv[i].full_proj_transform = std::get<1>(views[i]);
        kv[i].camera_center = std::get<2>(views[i]);
        const auto& p = std::get<3>(views[i]);
        TORCH_CHECK(p.size() == 7, "pose = {t.xyz, q.wxyz}");
        for (int k = 0; k < 7; ++k) kv[i].pose[k] = static_cast<float>(p[k]);
        kv[i].gt_image = std::get<4>(views[i]);
        if (std::get<5>(views[i]).numel()) kv[i].row_mask = std::get<5>(views[i]);
    }
    return fm.render_views(kv);
}

// densify::adjust_anchor over a dict of tensors under the reference's member names (+ "m_<name>" / "v_<name>" moments);
// -> the new dict (+ "_growing_report", "_prune_report")
py::dict dn_adjust_anchor(py::dict st, int check_interval, double success_threshold, double grad_threshold, double min_opacity,
                          int n_offsets, int update_depth, int update_init_factor, int update_hierachy_factor, double voxel_size) {
    static const char* names[6] = {"_anchor", "_offset", "_anchor_feat", "_opacity", "_scaling", "_rotation"};
    densify::AnchorState s;
    auto get = [&](const char* k) { return st[k].cast<torch::Tensor>(); };
    s.anchor = get("_anchor"); s.offset = get("_offset"); s.anchor_feat = get("_anchor_feat"); s.opacity = get("_opacity");
    s.scaling = get("_scaling"); s.rotation = get("_rotation");
    s.opacity_accum = get("opacity_accum"); s.anchor_demon = get("anchor_demon");
    s.offset_gradient_accum = get("offset_gradient_accum"); s.offset_denom = get("offset_denom");
    for (int g = 0; g < 6; ++g) {
        const std::string m = std::string("m_") + names[g], v = std::string("v_") + names[g];
        if (st.contains(m.c_str())) s.exp_avg[g] = st[m.c_str()].cast<torch::Tensor>();
        if (st.contains(v.c_str())) s.exp_avg_sq[g] = st[v.c_str()].cast<torch::Tensor>();
    }
    densify::ModelParams mp;
    mp.n_offsets = n_offsets; mp.update_depth = update_depth; mp.update_init_factor = update_init_factor;
    mp.update_hierachy_factor = update_hierachy_factor; mp.voxel_size = static_cast<float>(voxel_size);
    densify::adjust_anchor(s, mp, check_interval, static_cast<float>(success_threshold), static_cast<float>(grad_threshold),
                           static_cast<float>(min_opacity));
    py::dict out;
    torch::Tensor* ps[6] = {&s.anchor, &s.offset, &s.anchor_feat, &s.opacity, &s.scaling, &s.rotation};
    for (int g = 0; g < 6; ++g) {
        out[names[g]] = *ps[g];
        if (s.exp_avg[g].defined()) out[(std::string("m_") + names[g]).c_str()] = s.exp_avg[g];
        if (s.exp_avg_sq[g].defined()) out[(std::string("v_") + names[g]).c_str()] = s.exp_avg_sq[g];
    }
    out["opacity_accum"] = s.opacity_accum; out["anchor_demon"] = s.anchor_demon;
    out["offset_gradient_accum"] = s.offset_gradient_accum; out["offset_denom"] = s.offset_denom;
    py::list rep;
    for (auto& r : s.growing_report) rep.append(py::make_tuple(r[0], r[1]));
    out["_growing_report"] = rep;
    out["_prune_report"] = py::make_tuple(s.prune_report[0], s.prune_report[1]);
    return out;
}

}  // namespace

PYBIND11_MODULE(_segs_torch, m) {
    m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
    m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
    m.def("markVisible", &mark_visible);
    m.def("RasterizeGaussiansfilterCUDA", &RasterizeGaussiansfilterCUDA);
    m.def("RasterizeGaussiansprojectCUDA", &RasterizeGaussiansprojectCUDA);
    m.def("distCUDA2", &distCUDA2);
    m.def("rasterizer_forward", &rasterizer_forward);
    m.def("l1_loss", &lu_l1_loss);
    m.def("ssim", &lu_ssim);
    m.def("psnr", &lu_psnr);
    m.def("l1_ssim", &lu_l1_ssim);
    m.def("adam_step", &lu_adam_step);
    m.def("adjust_anchor", &dn_adjust_anchor, py::arg("state"), py::arg("check_interval") = 100, py::arg("success_threshold") = 0.8,
          py::arg("grad_threshold") = 0.0002, py::arg("min_opacity") = 0.005, py::arg("n_offsets") = 10, py::arg("update_depth") = 3,
          py::arg("update_init_factor") = 16, py::arg("update_hierachy_factor") = 4, py::arg("voxel_size") = 0.001);
    m.def("computeTransformTensors", &kf_transforms);
    m.def("quaternionToRotation", &kf_quat_to_rot);
    py::class_<FusedMapper>(m, "FusedMapper")
        .def(py::init([](std::vector<torch::Tensor> model, std::vector<c10::optional<torch::Tensor>> weights, std::vector<int> cfg,
                         int H, int W, double tanx, double tany, torch::Tensor bg, double lambda_dssim, double reg_w,
                         std::vector<double> lrs, double eps, int lanes) {
            TORCH_CHECK(cfg.size() == 5, "cfg = {appearance_dim, use_feat_bank, add_opacity_dist, add_cov_dist, add_color_dist}");
            std::vector<torch::Tensor> w;
            for (auto& o : weights) w.push_back(o.has_value() ? *o : torch::Tensor());
            return new FusedMapper(std::move(model), std::move(w), {cfg[0], cfg[1], cfg[2], cfg[3], cfg[4]}, H, W,
                                   static_cast<float>(tanx), static_cast<float>(tany), bg, lambda_dssim, reg_w, std::move(lrs),
                                   eps, lanes);
        }))
        .def("render_views", &fm_render_views)
        .def("grad_flat", &FusedMapper::grad_flat)
        .def("params", &FusedMapper::params)
        .def("adam_step", &FusedMapper::adam_step)
        .def("set_frequency", &FusedMapper::set_frequency)
        .def("workspace_bytes", &FusedMapper::workspace_bytes);
}
