"""TEST INFRASTRUCTURE — PyTorch restatement of GaussianModel::training_statis
(/root/reference/src/gaussian_model.cpp:1459-1503), the densification statistics the mapper updates after every
iteration (src/gaussian_mapper.cpp:963), op for op (use_coarse_anchor = false branch).  Only tests/ imports this."""
import torch


def training_statis(stats, viewspace_grad, opacity, update_filter, offset_selection_mask, anchor_visible_mask, n_offsets=10):
    """stats: dict(opacity_accum [A,1], anchor_demon [A,1], offset_gradient_accum [A*k,1], offset_denom [A*k,1]), updated
    in place.  viewspace_grad = viewspace_point_tensor.grad() [P,3]; opacity = neural_opacity [n_vis*k,1];
    update_filter = radii > 0 [P]; offset_selection_mask [n_vis*k] bool; anchor_visible_mask [A] bool."""
    temp_opacity = opacity.clone().view(-1).detach()
    temp_opacity = torch.where(temp_opacity < 0, torch.zeros_like(temp_opacity), temp_opacity)
    temp_opacity = temp_opacity.view(-1, n_offsets)
    stats["opacity_accum"][anchor_visible_mask] = stats["opacity_accum"][anchor_visible_mask] + temp_opacity.sum(1, keepdim=True)
    stats["anchor_demon"][anchor_visible_mask] = stats["anchor_demon"][anchor_visible_mask] + 1
    avm = anchor_visible_mask.unsqueeze(1).repeat(1, n_offsets).view(-1)
    combined_mask = torch.zeros_like(stats["offset_gradient_accum"], dtype=torch.bool).squeeze(1)
    combined_mask[avm] = offset_selection_mask
    temp_mask = combined_mask.clone()
    combined_mask[temp_mask] = update_filter
    grad_norm = torch.linalg.norm(viewspace_grad[update_filter, :2], dim=-1, keepdim=True)
    stats["offset_gradient_accum"][combined_mask] = stats["offset_gradient_accum"][combined_mask] + grad_norm
    stats["offset_denom"][combined_mask] = stats["offset_denom"][combined_mask] + 1
    return stats
