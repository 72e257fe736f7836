// TEST / BASELINE INFRASTRUCTURE — pybind11 veneer over the reference's OWN host layers, compiled unmodified from
// /root/reference/src (oracle/Makefile target `modelref`): GaussianModel (gaussian_model.cpp), GaussianRenderer
// (gaussian_renderer.cpp), GaussianRasterizer (gaussian_rasterizer.cpp), rasterize_points.cu, the CUDA kernels and
// loss_utils.h.  Nothing below restates arithmetic of the hot path; it only
//   * builds a GaussianModel / GaussianKeyframe from tensors handed over by the tests (the reference fills them from
//     a SLAM run: createFromPcd, computeTransformTensors),
//   * calls the reference functions, and
//   * restates the CALL SITE of one mapping iteration (src/gaussian_mapper.cpp:870-1006) so that the reference's own
//     iteration can be timed and compared with the fused step.
// torch_scatter (absent from the image) is the one third-party routine defined here: scatter_max, restating the
// published semantics of pytorch_scatter 2.1.2 (see below).
//
// Only tests/, tests/golden/make_model_golden.py, __graft_entry__.smoke() and bench.py's reference arm load this
// module; the product (segs_slam_b200/) never does.
#include <torch/extension.h>

#include "include/gaussian_model.h"
#include "include/gaussian_renderer.h"
#include "include/loss_utils.h"

// ---- torch_scatter 2.1.2: scatter_max(src, index, dim, out, dim_size) -> (out, arg_out) ---------------------------------
// csrc/scatter.cpp / csrc/cpu/scatter_cpu.cpp of pytorch_scatter: out has size max(index)+1 along `dim` (0 rows when
// index is empty), out[index[i][j]][j] = max over i of src[i][j]; slots no source row maps to hold 0; arg_out holds the
// winning source row, src.size(dim) for untouched slots.  The reference only calls it with dim = 0 on 2-D tensors and
// only reads std::get<0> (gaussian_model.cpp:1635-1637).
std::tuple<torch::Tensor, torch::Tensor> scatter_max(torch::Tensor src, torch::Tensor index, int64_t dim,
                                                     std::optional<torch::Tensor> optional_out, std::optional<int64_t> dim_size)
{
    TORCH_CHECK(dim == 0 && src.dim() == 2 && !optional_out.has_value(), "scatter_max stand-in: dim 0, 2-D, no out");
    int64_t n = dim_size.has_value() ? *dim_size : (index.numel() == 0 ? 0 : index.max().item<int64_t>() + 1);
    auto idx = index.expand_as(src).contiguous();
    auto out = torch::zeros({n, src.size(1)}, src.options());
    if (src.size(0) > 0) out = out.scatter_reduce(0, idx, src, "amax", /*include_self=*/false);
    auto rows = torch::arange(src.size(0), idx.options()).unsqueeze(1).expand_as(src);
    auto hit = src == out.gather(0, idx);
    auto arg = torch::full({n, src.size(1)}, src.size(0), idx.options());
    if (src.size(0) > 0)
        arg = arg.scatter_reduce(0, idx, torch::where(hit, rows, torch::full_like(rows, src.size(0))), "amin", true);
    return std::make_tuple(out, arg);
}

namespace {

using T = torch::Tensor;

struct RefModel {
    std::shared_ptr<GaussianModel> m;
    GaussianOptimizationParams opt;
    GaussianPipelineParams pipe;
    int iteration = 0;
    // frequency regularisation of the call site (src/gaussian_mapper.cpp:930-945); 0 = off.  Replica yamls:
    // lambda_frequency_high 0.01, use_multi_resolution 1, scale_num 3 (cfg/gaussian_mapper/RGB-D/Replica/office0.yaml:140-146)
    double lambda_frequency_high = 0.0;
    bool use_multi_resolution = true;
    int scale_num = 3;
    void set_frequency(double lambda_high, bool multi, int n) { lambda_frequency_high = lambda_high; use_multi_resolution = multi; scale_num = n; }
    T frequency_term(const T& rendered_image, const T& gt_image) const {
        std::vector<float> scales(scale_num, 0.f);
        for (int i = 0; i < scale_num; ++i) scales[i] = float(1.0 / pow(2, i));                 // gaussian_mapper.cpp:514-517
        if (use_multi_resolution) return lambda_frequency_high * loss_utils::multi_scale_loss(rendered_image, gt_image, scales);
        return lambda_frequency_high * loss_utils::high_frequency_loss(rendered_image, gt_image);
    }

    // reference_ctor = true : GaussianModel(const GaussianModelParams&) itself (gaussian_model.cpp:33-180; hard-codes
    //                         torch::kCUDA, so GPU only).
    // reference_ctor = false: GaussianModel(int) (:19-31) + the same module list built on `device` — lets the CPU-only
    //                         build container run generate_neural_gaussians for the committed goldens; a GPU test holds
    //                         both constructions to the same parameter shapes and outputs.
    RefModel(int feat_dim, int n_offsets, double voxel_size, int update_depth, int update_init_factor,
             int update_hierachy_factor, bool use_feat_bank, int appearance_dim, bool add_opacity_dist, bool add_cov_dist,
             bool add_color_dist, bool reference_ctor)
    {
        if (reference_ctor) {
            GaussianModelParams p("", "", "", 3, "images", -1.0f, false, "cuda", false, feat_dim, n_offsets, (float)voxel_size,
                                  update_depth, update_init_factor, update_hierachy_factor, use_feat_bank, appearance_dim,
                                  false, 1, 1.0f, false, add_opacity_dist, add_cov_dist, add_color_dist, 200, false);
            m = std::make_shared<GaussianModel>(p);
        } else {
            m = std::make_shared<GaussianModel>(3);
            auto dev = m->device_type_;
            m->feat_dim = feat_dim; m->n_offsets = n_offsets; m->voxel_size = (float)voxel_size;
            m->update_depth = update_depth; m->update_init_factor = update_init_factor;
            m->update_hierachy_factor = update_hierachy_factor; m->use_feat_bank = use_feat_bank;
            m->appearance_dim = appearance_dim; m->ratio = 1; m->add_opacity_dist = add_opacity_dist;
            m->add_cov_dist = add_cov_dist; m->add_color_dist = add_color_dist; m->embedding_dim = 200;
            m->opacity_dist_dim = add_opacity_dist; m->cov_dist_dim = add_cov_dist; m->color_dist_dim = add_color_dist;
            namespace nn = torch::nn;
            m->mlp_opacity = nn::Sequential(nn::Linear(feat_dim + 3 + m->opacity_dist_dim, feat_dim), nn::ReLU(nn::ReLUOptions().inplace(true)),
                                            nn::Linear(feat_dim, n_offsets), nn::Tanh());
            m->mlp_cov = nn::Sequential(nn::Linear(feat_dim + 3 + m->cov_dist_dim, feat_dim), nn::ReLU(nn::ReLUOptions().inplace(true)),
                                        nn::Linear(feat_dim, 7 * n_offsets));
            m->mlp_color = nn::Sequential(nn::Linear(feat_dim + 3 + m->color_dist_dim + appearance_dim, feat_dim),
                                          nn::ReLU(nn::ReLUOptions().inplace(true)), nn::Linear(feat_dim, 3 * n_offsets), nn::Sigmoid());
            m->mlp_apperance = nn::Sequential(nn::Linear(7, appearance_dim));
            if (use_feat_bank)
                m->mlp_feature_bank = nn::Sequential(nn::Linear(3 + 1, feat_dim), nn::ReLU(nn::ReLUOptions().inplace(true)),
                                                     nn::Linear(feat_dim, 3), nn::Softmax(nn::SoftmaxOptions(1)));
            for (auto* s : {&m->mlp_opacity, &m->mlp_cov, &m->mlp_color, &m->mlp_apperance, &m->mlp_feature_bank}) (*s)->to(dev);
        }
        m->spatial_lr_scale_ = 1.0f;
    }

    std::string device() const { return m->device_type_ == torch::kCUDA ? "cuda" : "cpu"; }

    void set_state(T anchor, T offset, T feat, T scaling, T rotation, T opacity) {
        auto dev = m->device_type_;
        auto mk = [&](T t, bool grad) { return t.detach().clone().to(dev).requires_grad_(grad); };
        m->_anchor = mk(anchor, true); m->_offset = mk(offset, true); m->_anchor_feat = mk(feat, true);
        m->_scaling = mk(scaling, true);
        m->_rotation = mk(rotation, false); m->_opacity = mk(opacity, false);       // gaussian_model.cpp:372-373
        m->Tensor_vec_anchor = {m->_anchor}; m->Tensor_vec_offset = {m->_offset}; m->Tensor_vec_anchor_feat = {m->_anchor_feat};
        m->Tensor_vec_opacity = {m->_opacity}; m->Tensor_vec_scaling = {m->_scaling}; m->Tensor_vec_rotation = {m->_rotation};
    }

    std::vector<torch::nn::Sequential> mlps() const {
        std::vector<torch::nn::Sequential> v = {m->mlp_opacity, m->mlp_cov, m->mlp_color, m->mlp_apperance};
        if (m->use_feat_bank) v.push_back(m->mlp_feature_bank);
        return v;
    }
    // weights in segs_decode_params order: opacity (w1,b1,w2,b2), cov, color, appearance (w,b), bank (w1,b1,w2,b2)
    std::vector<T> mlp_parameters() const {
        std::vector<T> out;
        for (auto& s : mlps()) for (auto& p : s->parameters()) out.push_back(p);
        return out;
    }
    void load_mlp_parameters(std::vector<T> w) {
        torch::NoGradGuard ng;
        auto ps = mlp_parameters();
        TORCH_CHECK(ps.size() == w.size(), "expected ", ps.size(), " MLP tensors, got ", w.size());
        for (size_t i = 0; i < ps.size(); ++i) {
            TORCH_CHECK(ps[i].sizes() == w[i].sizes(), "MLP tensor ", i, " shape mismatch");
            ps[i].copy_(w[i]);
        }
    }

    std::shared_ptr<GaussianKeyframe> keyframe(T view, T proj, T center, std::vector<double> t, std::vector<double> q_wxyz,
                                               double fovx, double fovy, int H, int W) const {
        auto kf = std::make_shared<GaussianKeyframe>(0, 0);
        auto dev = m->device_type_;
        kf->world_view_transform_ = view.to(dev); kf->full_proj_transform_ = proj.to(dev); kf->camera_center_ = center.to(dev);
        kf->t_ = Eigen::Vector3d(t[0], t[1], t[2]);
        kf->R_quaternion_ = Eigen::Quaterniond(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]);
        kf->FoVx_ = (float)fovx; kf->FoVy_ = (float)fovy; kf->image_height_ = H; kf->image_width_ = W;
        return kf;
    }

    // GaussianRenderer::generate_neural_gaussians (gaussian_renderer.cpp:214-334)
    std::vector<T> generate_neural_gaussians(T view, T proj, T center, std::vector<double> t, std::vector<double> q, T visible_mask) {
        auto kf = keyframe(view, proj, center, t, q, 1.0, 1.0, 1, 1);
        T mask = visible_mask.to(m->device_type_);
        auto r = GaussianRenderer::generate_neural_gaussians(kf, 1, 1, m, mask, true);
        return {std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r), std::get<6>(r)};
    }

    // GaussianRenderer::prefilter_voxel (:131-199)
    T prefilter_voxel(T view, T proj, T center, double fovx, double fovy, int H, int W, T bg) {
        auto kf = keyframe(view, proj, center, {0, 0, 0}, {1, 0, 0, 0}, fovx, fovy, H, W);
        T override_color;
        return GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
    }

    // GaussianRenderer::render (:19-127) -> (image, screenspace_points, visibility_filter, radii, mask, neural_opacity, scaling)
    std::vector<T> render(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx, double fovy, int H,
                          int W, T bg, T visible_mask) {
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto r = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, visible_mask, true);
        return {std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r), std::get<6>(r)};
    }

    void training_setup() {
        if (m->appearance_dim > 0 && !m->embedding_appearance) m->setApperance();
        m->trainingSetup(opt);
    }
    void set_learning_rates(std::vector<double> lrs) {
        auto& g = m->optimizer_->param_groups();
        TORCH_CHECK(lrs.size() == g.size(), "expected ", g.size(), " learning rates");
        for (size_t i = 0; i < g.size(); ++i) static_cast<torch::optim::AdamOptions&>(g[i].options()).lr(lrs[i]);
    }
    int n_param_groups() const { return (int)m->optimizer_->param_groups().size(); }

    // Give the four trainable anchor tensors (anchor, offset, feat, scaling) Adam moments: one torch::optim::Adam step on the
    // given gradients.  With every learning rate at 0 (set_learning_rates) the parameters do not move.
    void adam_step_with_grads(std::vector<T> grads) {
        TORCH_CHECK(grads.size() == 4, "expected gradients of anchor, offset, feat, scaling");
        std::vector<T> ps = {m->_anchor, m->_offset, m->_anchor_feat, m->_scaling};
        for (int i = 0; i < 4; ++i) ps[i].mutable_grad() = grads[i].to(ps[i].device()).clone();
        torch::NoGradGuard ng;
        m->optimizer_->step();
        m->optimizer_->zero_grad(true);
    }

    // GaussianModel::training_statis (gaussian_model.cpp:1459-1503); reads viewspace_point_tensor.grad()
    void training_statis(T viewspace_grad, T opacity, T update_filter, T offset_selection_mask, T anchor_visible_mask) {
        T v = torch::zeros_like(viewspace_grad).requires_grad_(true);
        v.mutable_grad() = viewspace_grad;
        torch::NoGradGuard ng;
        T avm = anchor_visible_mask.clone();
        m->training_statis(v, opacity, update_filter, offset_selection_mask, avm);
    }
    // GaussianModel::adjust_anchor (:1705-1762) -> anchor_growing (:1556-1703), prune_anchor (:1505-1555)
    void adjust_anchor(int check_interval, double success_threshold, double grad_threshold, double min_opacity) {
        torch::NoGradGuard ng;
        m->adjust_anchor(check_interval, (float)success_threshold, (float)grad_threshold, (float)min_opacity);
    }
    void set_statistics(T opacity_accum, T anchor_demon, T offset_gradient_accum, T offset_denom) {
        auto dev = m->device_type_;
        m->opacity_accum = opacity_accum.detach().clone().to(dev); m->anchor_demon = anchor_demon.detach().clone().to(dev);
        m->offset_gradient_accum = offset_gradient_accum.detach().clone().to(dev);
        m->offset_denom = offset_denom.detach().clone().to(dev);
    }
    std::vector<T> state() const {
        return {m->_anchor, m->_offset, m->_anchor_feat, m->_scaling, m->_rotation, m->_opacity, m->opacity_accum, m->anchor_demon,
                m->offset_gradient_accum, m->offset_denom};
    }
    // Adam moments of the six anchor groups (anchor, offset, feat, opacity, scaling, rotation): (step, exp_avg, exp_avg_sq)
    std::vector<std::vector<T>> adam_state() const {
        std::vector<std::vector<T>> out;
        auto& groups = m->optimizer_->param_groups();
        auto& st = m->optimizer_->state();
        for (int g = 0; g < 6; ++g) {
            auto& p = groups[g].params()[0];
            auto it = st.find(p.unsafeGetTensorImpl());
            if (it == st.end()) { out.push_back({}); continue; }
            auto& s = static_cast<torch::optim::AdamParamState&>(*it->second);
            out.push_back({torch::tensor((double)s.step()), s.exp_avg(), s.exp_avg_sq()});
        }
        return out;
    }

    // One mapping iteration on one keyframe: the call site src/gaussian_mapper.cpp:870-1006 (learning-rate updates and
    // logging left out), every callee the reference's own.  densify: run training_statis (+ adjust_anchor when due).
    // -> (loss, Ll1, ssim)
    std::vector<double> train_iteration(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx,
                                        double fovy, int H, int W, T bg, T gt_image_in, double lambda_dssim, bool statis,
                                        bool adjust, bool frequency, bool sync) {
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto voxel_visible_mask = GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
        auto pkg = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, voxel_visible_mask, true);
        auto rendered_image = std::get<0>(pkg);
        auto viewspace_point_tensor = std::get<1>(pkg);
        auto visibility_filter = std::get<2>(pkg);
        auto offset_selection_mask = std::get<4>(pkg);
        auto opacity = std::get<5>(pkg);
        auto scaling = std::get<6>(pkg);
        T masked_image = rendered_image;
        T gt_image = gt_image_in;
        T mask_rgb = (gt_image != 0.0f).any(-1);
        mask_rgb = mask_rgb.to(torch::kFloat32).unsqueeze(-1);
        masked_image = masked_image * mask_rgb;
        rendered_image = rendered_image * mask_rgb;
        gt_image = gt_image * mask_rgb;
        auto Ll1 = loss_utils::l1_loss(rendered_image, gt_image);
        auto ssim = loss_utils::ssim(masked_image, gt_image, m->device_type_);
        auto scaling_reg = scaling.prod(1).mean();
        auto loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ssim) + 0.01 * scaling_reg;
        if (frequency && lambda_frequency_high != 0.0) loss = loss + frequency_term(rendered_image, gt_image);     // :930-945
        loss.backward();
        if (sync) torch::cuda::synchronize();
        {
            torch::NoGradGuard ng;
            if (statis) {
                m->training_statis(viewspace_point_tensor, opacity, visibility_filter, offset_selection_mask, voxel_visible_mask);
                if (adjust)
                    m->adjust_anchor(opt.update_interval, opt.success_threshold, opt.densify_grad_threshold, opt.min_opacity);
            }
            m->optimizer_->step();
            m->optimizer_->zero_grad(true);
        }
        ++iteration;
        if (!sync) return {};
        return {loss.item<double>(), Ll1.item<double>(), ssim.item<double>()};
    }

    // Reference baseline "B" (BASELINE.md section 2): the reference's per-view work with gradients ACCUMULATED in .grad
    // (loss scaled by `loss_scale` = 1 / batch) and no optimizer step; optimizer_step() then applies ONE Adam step.
    T backward_view(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx, double fovy, int H, int W,
                    T bg, T gt_image_in, double lambda_dssim, double loss_scale) {
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto voxel_visible_mask = GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
        auto pkg = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, voxel_visible_mask, true);
        auto rendered_image = std::get<0>(pkg);
        auto scaling = std::get<6>(pkg);
        T gt_image = gt_image_in;
        T mask_rgb = (gt_image != 0.0f).any(-1).to(torch::kFloat32).unsqueeze(-1);
        T masked_image = rendered_image * mask_rgb;
        rendered_image = rendered_image * mask_rgb;
        gt_image = gt_image * mask_rgb;
        auto Ll1 = loss_utils::l1_loss(rendered_image, gt_image);
        auto loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - loss_utils::ssim(masked_image, gt_image, m->device_type_)) +
                    0.01 * scaling.prod(1).mean();
        if (lambda_frequency_high != 0.0) loss = loss + frequency_term(rendered_image, gt_image);
        (loss * loss_scale).backward();
        return loss.detach();
    }
    void optimizer_step() {
        torch::NoGradGuard ng;
        m->optimizer_->step();
        m->optimizer_->zero_grad(true);
    }
    // BASELINE config 3: prefilter -> decode -> rasterize and the backward of a given dL_dimage (no loss)
    T render_fwd_bwd(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx, double fovy, int H, int W,
                     T bg, T dL_dimage) {
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto voxel_visible_mask = GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
        auto pkg = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, voxel_visible_mask, true);
        auto image = std::get<0>(pkg);
        image.backward(dL_dimage);
        std::vector<T> params = {m->_anchor, m->_offset, m->_anchor_feat, m->_scaling};
        for (auto& p : mlp_parameters()) params.push_back(p);
        for (auto& p : params) p.mutable_grad() = T();
        return image.detach();
    }
    // forward only (ground-truth images of the synthetic mapping workload)
    T render_image(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx, double fovy, int H, int W, T bg) {
        torch::NoGradGuard ng;
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto voxel_visible_mask = GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
        auto pkg = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, voxel_visible_mask, false);
        return std::get<0>(pkg).detach();
    }

    // backward only (no optimizer step): gradients of the loss of ONE view w.r.t. every trainable tensor, for the parity
    // test of the fused mapping view.  -> (loss, grads of [anchor, offset, feat, scaling] + MLP parameters, image,
    // viewspace grad, radii, offset mask, neural opacity, visible mask)
    std::vector<T> view_gradients(T view, T proj, T center, std::vector<double> t, std::vector<double> q, double fovx, double fovy,
                                  int H, int W, T bg, T gt_image_in, double lambda_dssim) {
        auto kf = keyframe(view, proj, center, t, q, fovx, fovy, H, W);
        T override_color;
        auto voxel_visible_mask = GaussianRenderer::prefilter_voxel(kf, H, W, m, pipe, bg, override_color);
        auto pkg = GaussianRenderer::render(kf, H, W, m, pipe, bg, override_color, voxel_visible_mask, true);
        auto rendered_image = std::get<0>(pkg);
        auto scaling = std::get<6>(pkg);
        T raw_image = rendered_image;
        raw_image.retain_grad();
        T image = rendered_image.detach().clone();
        T gt_image = gt_image_in;
        T mask_rgb = (gt_image != 0.0f).any(-1).to(torch::kFloat32).unsqueeze(-1);
        T masked_image = rendered_image * mask_rgb;
        rendered_image = rendered_image * mask_rgb;
        gt_image = gt_image * mask_rgb;
        auto Ll1 = loss_utils::l1_loss(rendered_image, gt_image);
        auto loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - loss_utils::ssim(masked_image, gt_image, m->device_type_)) +
                    0.01 * scaling.prod(1).mean();
        if (lambda_frequency_high != 0.0) loss = loss + frequency_term(rendered_image, gt_image);
        std::vector<T> params = {m->_anchor, m->_offset, m->_anchor_feat, m->_scaling};
        for (auto& p : mlp_parameters()) params.push_back(p);
        for (auto& p : params) if (p.grad().defined()) p.mutable_grad() = T();
        loss.backward();
        std::vector<T> out = {loss.detach()};
        for (auto& p : params) out.push_back(p.grad().defined() ? p.grad().clone() : torch::zeros_like(p));
        out.push_back(image);
        out.push_back(std::get<1>(pkg).grad().clone());
        out.push_back(std::get<3>(pkg));
        out.push_back(std::get<4>(pkg));
        out.push_back(std::get<5>(pkg).detach());
        out.push_back(voxel_visible_mask);
        out.push_back(raw_image.grad().defined() ? raw_image.grad().clone() : torch::zeros_like(image));    // dL/d(rendered image)
        for (auto& p : params) p.mutable_grad() = T();
        return out;
    }
};

}  // namespace

PYBIND11_MODULE(_model_ref, mod) {
    mod.doc() = "the reference's GaussianModel / GaussianRenderer compiled unmodified (test infrastructure)";
    py::class_<RefModel>(mod, "RefModel")
        .def(py::init<int, int, double, int, int, int, bool, int, bool, bool, bool, bool>(), py::arg("feat_dim") = 32,
             py::arg("n_offsets") = 10, py::arg("voxel_size") = 0.001, py::arg("update_depth") = 3,
             py::arg("update_init_factor") = 16, py::arg("update_hierachy_factor") = 4, py::arg("use_feat_bank") = true,
             py::arg("appearance_dim") = 32, py::arg("add_opacity_dist") = false, py::arg("add_cov_dist") = false,
             py::arg("add_color_dist") = false, py::arg("reference_ctor") = false)
        .def("device", &RefModel::device)
        .def("set_frequency", &RefModel::set_frequency)
        .def("set_state", &RefModel::set_state)
        .def("mlp_parameters", &RefModel::mlp_parameters)
        .def("load_mlp_parameters", &RefModel::load_mlp_parameters)
        .def("generate_neural_gaussians", &RefModel::generate_neural_gaussians)
        .def("prefilter_voxel", &RefModel::prefilter_voxel)
        .def("render", &RefModel::render)
        .def("training_setup", &RefModel::training_setup)
        .def("set_learning_rates", &RefModel::set_learning_rates)
        .def("n_param_groups", &RefModel::n_param_groups)
        .def("adam_step_with_grads", &RefModel::adam_step_with_grads, py::call_guard<py::gil_scoped_release>())
        .def("training_statis", &RefModel::training_statis)
        .def("adjust_anchor", &RefModel::adjust_anchor)
        .def("set_statistics", &RefModel::set_statistics)
        .def("state", &RefModel::state)
        .def("adam_state", &RefModel::adam_state)
        .def("train_iteration", &RefModel::train_iteration, py::call_guard<py::gil_scoped_release>())
        .def("backward_view", &RefModel::backward_view, py::call_guard<py::gil_scoped_release>())
        .def("optimizer_step", &RefModel::optimizer_step, py::call_guard<py::gil_scoped_release>())
        .def("render_fwd_bwd", &RefModel::render_fwd_bwd, py::call_guard<py::gil_scoped_release>())
        .def("render_image", &RefModel::render_image, py::call_guard<py::gil_scoped_release>())
        .def("view_gradients", &RefModel::view_gradients, py::call_guard<py::gil_scoped_release>());
    // the reference's six tensor-level entry points (include/rasterize_points.h:18-102, third_party/simple-knn/spatial.h:14)
    mod.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
    mod.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
    mod.def("RasterizeGaussiansfilterCUDA", &RasterizeGaussiansfilterCUDA);
    mod.def("RasterizeGaussiansprojectCUDA", &RasterizeGaussiansprojectCUDA);
    mod.def("markVisible", &markVisible);
    mod.def("distCUDA2", &distCUDA2);
    // tanfovx exactly as GaussianRenderer::render derives it from the keyframe's FoVx_ (gaussian_renderer.cpp:67-68)
    mod.def("tan_half_fov", [](double fov) { return (double)std::tan((float)fov * 0.5f); });
    // loss_utils::high_frequency_loss / multi_scale_loss of the reference's own header on the tensors' device, with the
    // gradient w.r.t. the first image: -> (value, d value / d img1)
    mod.def("high_frequency_loss", [](T a, T b) {
        T x = a.detach().clone().requires_grad_(true);
        T l = loss_utils::high_frequency_loss(x, b, 0.4f, a.device().type());
        { py::gil_scoped_release nogil; l.backward(); }
        return std::vector<T>{l.detach(), x.grad()};
    });
    mod.def("multi_scale_loss", [](T a, T b, std::vector<float> scales) {
        T x = a.detach().clone().requires_grad_(true);
        T l = loss_utils::multi_scale_loss(x, b, scales, a.device().type());
        { py::gil_scoped_release nogil; l.backward(); }
        return std::vector<T>{l.detach(), x.grad()};
    });
    mod.def("scatter_max", [](T src, T index) { auto r = scatter_max(src, index, 0, std::nullopt, std::nullopt); return std::vector<T>{std::get<0>(r), std::get<1>(r)}; });
}
