// Fused multi-tensor Adam over the mapper's flat gradient bucket for sm_100a.
//
// Replaces the torch::optim::Adam step the reference takes once per iteration over its 11-23
// parameter groups (/root/reference/src/gaussian_model.cpp:620-872 builds the groups — one tensor
// per group, each with its own learning rate —, src/gaussian_mapper.cpp:1003-1006 steps it).
// LibTorch runs >= 5 elementwise kernels per tensor (mul_, add_, mul_, addcmul_, sqrt, div, add_,
// addcdiv_); here ONE launch updates every tensor: the gradients, first and second moments live in
// three flat FP32 arrays laid out like the all-reduce bucket (mapper.py:GradBucket), the parameters
// stay where the model keeps them.  The 1/B scale of the batch-mean gradient and the clearing of the
// bucket for the next step ride along, so the bucket is read once and written once per step.
//
// Per element (torch/csrc/api/src/optim/adam.cpp of LibTorch 2.0.1, amsgrad off):
//   g   = grad * grad_scale (+ weight_decay * p)
//   m   = m * beta1 + (1 - beta1) * g
//   v   = v * beta2 + (1 - beta2) * g * g
//   p  -= (lr / bias_correction1) * m / (sqrt(v) / sqrt(bias_correction2) + eps)
// HBM-streaming: 16 B read + 12-16 B written per element.
#include "common.cuh"

namespace segs {

namespace {

constexpr int ADAM_MAX_TENSORS = 32;

struct AdamTable {
    float* param[ADAM_MAX_TENSORS];
    unsigned long long begin[ADAM_MAX_TENSORS + 1];   // element offsets into the flat arrays (ascending)
    float step_size[ADAM_MAX_TENSORS];                // lr / bias_correction1
    float sqrt_bc2[ADAM_MAX_TENSORS];
    float beta1[ADAM_MAX_TENSORS], beta2[ADAM_MAX_TENSORS], eps[ADAM_MAX_TENSORS], weight_decay[ADAM_MAX_TENSORS];
    int count;
};

__global__ void __launch_bounds__(256)
adam_step_kernel(const AdamTable t, float* __restrict__ grad, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq, float grad_scale, int zero_grad)
{
    const unsigned long long lo = t.begin[0], hi = t.begin[t.count];
    for (unsigned long long i = lo + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
         i += size_t(gridDim.x) * blockDim.x) {
        int a = 0, b = t.count - 1;          // tensor holding flat element i
        while (a < b) {
            const int mid = (a + b + 1) >> 1;
            if (t.begin[mid] <= i) a = mid; else b = mid - 1;
        }
        float* p = t.param[a] + (i - t.begin[a]);
        float g = grad[i] * grad_scale;
        const float pv = *p;
        if (t.weight_decay[a] != 0.f) g = fmaf(t.weight_decay[a], pv, g);
        const float m = fmaf(1.f - t.beta1[a], g, exp_avg[i] * t.beta1[a]);
        const float v = fmaf(1.f - t.beta2[a], g * g, exp_avg_sq[i] * t.beta2[a]);
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / t.sqrt_bc2[a] + t.eps[a];
        *p = pv - t.step_size[a] * (m / denom);
        if (zero_grad) grad[i] = 0.f;
    }
}

// dst[i] += src[i] over up to 24 (dst, src, n) triples in one launch (gradient accumulation of one view
// into the bucket when the producing kernel cannot accumulate in place)
struct AddList { float* dst[24]; const float* src[24]; unsigned long long n[24]; int count; int atomic; };
__global__ void __launch_bounds__(256)
add_many_kernel(const AddList z)
{
    // blockIdx.y = array: the arrays proceed side by side (18 tiny MLP tensors would otherwise be 18 dependent
    // load-add-store round trips)
    const int a = blockIdx.y;
    float* d = z.dst[a];
    const float* s = z.src[a];
    const unsigned long long n = z.n[a];
    for (unsigned long long i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        if (z.atomic) atomicAdd(d + i, s[i]); else d[i] += s[i];
    }
}

}  // namespace

}  // namespace segs

using namespace segs;

extern "C" {

int segs_adam_step(int n_tensors, const segs_adam_tensor* tensors, float* grad_flat, float* exp_avg_flat,
                   float* exp_avg_sq_flat, float grad_scale, int zero_grad, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_tensors < 0 || (n_tensors > 0 && (!tensors || !grad_flat || !exp_avg_flat || !exp_avg_sq_flat))) {
        set_error("adam: invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    for (int first = 0; first < n_tensors; first += ADAM_MAX_TENSORS) {
        AdamTable t;
        const int n = n_tensors - first < ADAM_MAX_TENSORS ? n_tensors - first : ADAM_MAX_TENSORS;
        unsigned long long total = 0;
        for (int k = 0; k < n; ++k) {
            const segs_adam_tensor& s = tensors[first + k];
            if (!s.param && s.count) { set_error("adam: NULL parameter pointer (tensor %d)", first + k); return SEGS_ERR_INVALID_ARG; }
            if (k > 0 && s.offset != t.begin[k]) { set_error("adam: tensors must tile the flat arrays contiguously (tensor %d)", first + k); return SEGS_ERR_INVALID_ARG; }
            if (s.step < 1) { set_error("adam: step must be >= 1 (tensor %d)", first + k); return SEGS_ERR_INVALID_ARG; }
            t.param[k] = s.param;
            t.begin[k] = s.offset;
            t.begin[k + 1] = s.offset + s.count;
            // bias corrections in double on the host, as LibTorch does (adam.cpp: 1 - std::pow(beta, step))
            const double bc1 = 1.0 - pow((double)s.beta1, (double)s.step);
            const double bc2 = 1.0 - pow((double)s.beta2, (double)s.step);
            t.step_size[k] = (float)((double)s.lr / bc1);
            t.sqrt_bc2[k] = (float)sqrt(bc2);
            t.beta1[k] = s.beta1; t.beta2[k] = s.beta2; t.eps[k] = s.eps; t.weight_decay[k] = s.weight_decay;
            total += s.count;
        }
        t.count = n;
        if (total == 0) continue;
        const unsigned long long want = (total + 255) / 256;
        const int grid = (int)(want < (unsigned long long)(SM_COUNT * 16) ? want : (unsigned long long)(SM_COUNT * 16));
        adam_step_kernel<<<grid, 256, 0, stream>>>(t, grad_flat, exp_avg_flat, exp_avg_sq_flat, grad_scale, zero_grad);
        SEGS_LAUNCH_CHECK();
    }
    return SEGS_OK;
}

int segs_accumulate(int n_arrays, float* const* dst, const float* const* src, const unsigned long long* counts,
                    int atomic, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_arrays < 0 || (n_arrays > 0 && (!dst || !src || !counts))) { set_error("accumulate: invalid argument"); return SEGS_ERR_INVALID_ARG; }
    for (int first = 0; first < n_arrays; first += 24) {
        AddList z;
        z.count = 0;
        z.atomic = atomic;
        unsigned long long largest = 0;
        for (int k = first; k < n_arrays && k < first + 24; ++k) {
            if (!counts[k]) continue;
            if (!dst[k] || !src[k]) { set_error("accumulate: NULL array %d", k); return SEGS_ERR_INVALID_ARG; }
            z.dst[z.count] = dst[k]; z.src[z.count] = src[k]; z.n[z.count] = counts[k]; ++z.count;
            if (counts[k] > largest) largest = counts[k];
        }
        if (!z.count) continue;
        const unsigned long long want = (largest + 255) / 256;
        const int grid = (int)(want < (unsigned long long)(SM_COUNT * 8) ? want : (unsigned long long)(SM_COUNT * 8));
        add_many_kernel<<<dim3(grid, z.count), 256, 0, stream>>>(z);
        SEGS_LAUNCH_CHECK();
    }
    return SEGS_OK;
}

}  // extern "C"
