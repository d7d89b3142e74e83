// Pseudo-label quality metrics (--eval_pslabel of the CLI) on the device:
// get_miou_scene and get_scene_sem_conf of /root/reference/gapro/eval_ps_labels.py:100-172.
// The reference builds N x K one-hot matrices and multiplies them (torch.mm); here ONE pass over the
// points fills the (K+1) x (K'+1) contingency table with integer atomics (exact counts, privatised in
// shared memory when it fits) and the first-point index of every instance id, then one small kernel
// turns the table into the per-instance best IoU with the float32 arithmetic of cal_iou (:36-43).
#include "common.cuh"

namespace {

constexpr int EVAL_SMEM_CELLS = 10240;   // 40 KB table per CTA

__global__ void __launch_bounds__(256)
k_eval_count(const int32_t* __restrict__ gt_inst, const int32_t* __restrict__ ps_inst, int64_t n, int n_gt, int n_ps,
             int32_t* __restrict__ table, int32_t* __restrict__ first_gt, int32_t* __restrict__ first_ps) {
    extern __shared__ int32_t s_tab[];
    const int cols = n_ps + 1, cells = (n_gt + 1) * cols;
    const bool priv = cells <= EVAL_SMEM_CELLS;
    if (priv) {
        for (int c = threadIdx.x; c < cells; c += blockDim.x) s_tab[c] = 0;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int g = gt_inst[i], p = ps_inst[i];
        const int gi = g < 0 ? 0 : g + 1, pi = p < 0 ? 0 : p + 1;     // :120, :127 (negative ids -> row / column 0)
        if (gi > n_gt || pi > n_ps) continue;                         // cannot happen: n_gt, n_ps are max + 1
        if (priv) atomicAdd(&s_tab[gi * cols + pi], 1);
        else atomicAdd(&table[(size_t)gi * cols + pi], 1);
        if (g >= 0) atomicMin(first_gt + g, (int32_t)i);              // idx_[0] of :104-108
        if (p >= 0) atomicMin(first_ps + p, (int32_t)i);
    }
    if (priv) {
        __syncthreads();
        for (int c = threadIdx.x; c < cells; c += blockDim.x) {
            const int v = s_tab[c];
            if (v) atomicAdd(&table[c], v);
        }
    }
}

__global__ void k_eval_fill(int32_t* __restrict__ table, int cells, int32_t* __restrict__ first_gt, int n_gt,
                            int32_t* __restrict__ first_ps, int n_ps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) table[i] = 0;
    if (i < n_gt) first_gt[i] = INT32_MAX;
    if (i < n_ps) first_ps[i] = INT32_MAX;
}

// one warp per GT instance id
__global__ void __launch_bounds__(256)
k_eval_miou(const int32_t* __restrict__ table, const int32_t* __restrict__ first_gt, const int32_t* __restrict__ first_ps,
            const int32_t* __restrict__ gt_sem, const int32_t* __restrict__ ps_sem, int n_gt, int n_ps,
            float* __restrict__ max_iou, int32_t* __restrict__ valid) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_gt) return;
    const int cols = n_ps + 1;
    const int fg = first_gt[g];
    const float gcls = fg == INT32_MAX ? -1.0f : (float)gt_sem[fg];
    int area_g = 0;
    for (int p = lane; p < cols; p += 32) area_g += table[(size_t)(g + 1) * cols + p];
    for (int o = 16; o; o >>= 1) area_g += __shfl_xor_sync(0xffffffffu, area_g, o);
    float best = 0.0f;      // torch.max over a row of ious >= 0 (rows are never empty here: n_ps >= 1)
    for (int p = lane; p < n_ps; p += 32) {
        int area_p = 0;
        for (int r = 0; r <= n_gt; ++r) area_p += table[(size_t)r * cols + p + 1];
        const int fp = first_ps[p];
        const float pcls = fp == INT32_MAX ? -1.0f : (float)ps_sem[fp];
        const float inter = (float)table[(size_t)(g + 1) * cols + p + 1];
        // cal_iou: intersection / (label_pointnum + ps_label_pointnum - intersection + 1e-4), float32
        const float den = __fadd_rn(__fsub_rn(__fadd_rn((float)area_g, (float)area_p), inter), 1e-4f);
        float iou = __fdiv_rn(inter, den);
        iou = iou * (gcls == pcls ? 1.0f : 0.0f);
        best = fmaxf(best, iou);
    }
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        max_iou[g] = best;
        valid[g] = gcls >= 0.0f ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
k_eval_sem_conf(const int32_t* __restrict__ gt_sem, const int32_t* __restrict__ ps_sem, int64_t n, int nc,
                int64_t* __restrict__ conf) {
    extern __shared__ int32_t s_conf[];
    for (int c = threadIdx.x; c < nc * nc; c += blockDim.x) s_conf[c] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int g = gt_sem[i];
        if (g == -100) continue;                                     // :153-156
        int p = ps_sem[i];
        if (p == -100) p = g < 18 ? g + 1 : g - 1;                   // :159-163: unlabelled = a wrong neighbour class
        const int x = p + nc * g;                                    // :165
        if (x >= 0 && x < nc * nc) atomicAdd(&s_conf[x], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nc * nc; c += blockDim.x) {
        const int v = s_conf[c];
        if (v) atomicAdd((unsigned long long*)&conf[c], (unsigned long long)v);
    }
}

}  // namespace

extern "C" size_t gapro_eval_workspace_bytes(int32_t n_gt, int32_t n_ps) {
    if (n_gt < 0 || n_ps < 0) return 0;
    return ((size_t)(n_gt + 1) * (n_ps + 1) + n_gt + n_ps + 16) * 4;
}

extern "C" int gapro_eval_miou_scene(const int32_t* gt_sem, const int32_t* gt_inst, const int32_t* ps_sem,
                                     const int32_t* ps_inst, int64_t n_points, int32_t n_gt, int32_t n_ps,
                                     float* max_iou, int32_t* valid, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GAPRO_REQUIRE(gt_sem && gt_inst && ps_sem && ps_inst && max_iou && valid && ws, "gapro_eval_miou_scene: null pointer");
    GAPRO_REQUIRE(n_points > 0 && n_points < (int64_t)INT32_MAX && n_gt >= 1 && n_ps >= 1,
                  "gapro_eval_miou_scene: need points and at least one id on either side (n_gt=%d n_ps=%d)", n_gt, n_ps);
    if (ws_bytes < gapro_eval_workspace_bytes(n_gt, n_ps)) {
        gapro_set_error("gapro_eval_miou_scene: workspace %zu < %zu bytes", ws_bytes, gapro_eval_workspace_bytes(n_gt, n_ps));
        return GAPRO_ERR_WORKSPACE;
    }
    const int cells = (n_gt + 1) * (n_ps + 1);
    int32_t* table = (int32_t*)ws;
    int32_t* first_gt = table + cells;
    int32_t* first_ps = first_gt + n_gt;
    const int fillN = cells > n_gt ? (cells > n_ps ? cells : n_ps) : (n_gt > n_ps ? n_gt : n_ps);
    k_eval_fill<<<(fillN + 255) / 256, 256, 0, stream>>>(table, cells, first_gt, n_gt, first_ps, n_ps);
    int dev = 0, sms = 148;
    GAPRO_CUDA_TRY(cudaGetDevice(&dev));
    GAPRO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (n_points + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    const size_t smem = cells <= EVAL_SMEM_CELLS ? (size_t)cells * 4 : 0;
    k_eval_count<<<grid, 256, smem, stream>>>(gt_inst, ps_inst, n_points, n_gt, n_ps, table, first_gt, first_ps);
    k_eval_miou<<<(n_gt + 7) / 8, 256, 0, stream>>>(table, first_gt, first_ps, gt_sem, ps_sem, n_gt, n_ps, max_iou, valid);
    GAPRO_KERNEL_CHECK();
    return GAPRO_OK;
}

extern "C" int gapro_eval_sem_conf(const int32_t* gt_sem, const int32_t* ps_sem, int64_t n_points, int32_t num_classes,
                                   int64_t* conf, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GAPRO_REQUIRE(gt_sem && ps_sem && conf, "gapro_eval_sem_conf: null pointer");
    GAPRO_REQUIRE(n_points > 0 && num_classes >= 1 && num_classes <= 64, "gapro_eval_sem_conf: bad sizes");
    GAPRO_CUDA_TRY(cudaMemsetAsync(conf, 0, (size_t)num_classes * num_classes * 8, stream));
    int dev = 0, sms = 148;
    GAPRO_CUDA_TRY(cudaGetDevice(&dev));
    GAPRO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (n_points + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    k_eval_sem_conf<<<grid, 256, (size_t)num_classes * num_classes * 4, stream>>>(gt_sem, ps_sem, n_points, num_classes, conf);
    GAPRO_KERNEL_CHECK();
    return GAPRO_OK;
}

// =============================================================================================
// G - per-instance axis-aligned boxes on the device: getInstanceInfo (gen_ps_utils.py:195-239).
// The reference loops over instance ids with np.where; here ONE pass over the points keeps, per id, the
// min / max of xyz (exact: ordered-uint64 atomics, order-free) and the index of the first point (its semantic
// label is the class, :212-213); a second tiny kernel lists the ids in use in increasing order (:209-210 skips
// empty ids), writes the float64 boxes, the volumes prod(clip(max - min, 0)) (:227) and the classes
// (scannetv2: minus 2 unless -100, :236-237).
// =============================================================================================
namespace {

__device__ __forceinline__ unsigned long long gi_key(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double gi_dbl(unsigned long long u) {
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

__global__ void k_gi_init(unsigned long long* __restrict__ ext, int32_t* __restrict__ first, int n_ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ids * 6) ext[i] = (i % 6) < 3 ? ~0ull : 0ull;
    if (i < n_ids) first[i] = INT32_MAX;
}

__global__ void __launch_bounds__(256)
k_gi_points(const double* __restrict__ xyz, const double* __restrict__ inst, int64_t n, int n_ids,
            unsigned long long* __restrict__ ext, int32_t* __restrict__ first) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const double v = inst[p];
        if (!(v >= 0.0) || v != floor(v) || v >= (double)n_ids) continue;      // np.where(instance_label == i), i >= 0
        const int id = (int)v;
        const double x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
        unsigned long long* e = ext + (size_t)id * 6;
        atomicMin(e + 0, gi_key(x));
        atomicMin(e + 1, gi_key(y));
        atomicMin(e + 2, gi_key(z));
        atomicMax(e + 3, gi_key(x));
        atomicMax(e + 4, gi_key(y));
        atomicMax(e + 5, gi_key(z));
        atomicMin(first + id, (int32_t)p);
    }
}

// one CTA: ids in use in increasing order
__global__ void __launch_bounds__(1024)
k_gi_finish(const unsigned long long* __restrict__ ext, const int32_t* __restrict__ first, const double* __restrict__ sem,
            int n_ids, int scannet, double* __restrict__ boxes, double* __restrict__ vol, double* __restrict__ cls,
            int32_t* __restrict__ n_used) {
    __shared__ int s_cnt[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_ids; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int used = (i < n_ids && first[i] != INT32_MAX) ? 1 : 0;
        s_cnt[threadIdx.x] = used;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {                 // inclusive scan
            const int v = threadIdx.x >= o ? s_cnt[threadIdx.x - o] : 0;
            __syncthreads();
            s_cnt[threadIdx.x] += v;
            __syncthreads();
        }
        if (used) {
            const int k = s_base + s_cnt[threadIdx.x] - 1;
            double lo[3], hi[3];
            double v = 1.0;
            for (int d = 0; d < 3; ++d) {
                lo[d] = gi_dbl(ext[(size_t)i * 6 + d]);
                hi[d] = gi_dbl(ext[(size_t)i * 6 + 3 + d]);
                boxes[6 * k + d] = lo[d];
                boxes[6 * k + 3 + d] = hi[d];
            }
            for (int d = 0; d < 3; ++d) {
                double e = __dsub_rn(hi[d], lo[d]);
                e = e < 0.0 ? 0.0 : e;
                v = d == 0 ? e : __dmul_rn(v, e);
            }
            vol[k] = v;
            double c = sem[first[i]];
            if (scannet && c != -100.0) c -= 2.0;
            cls[k] = c;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_base += s_cnt[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_used = s_base;
}

}  // namespace

extern "C" size_t gapro_instance_info_workspace_bytes(int32_t n_ids) {
    if (n_ids <= 0) return 0;
    return gapro_align_up((size_t)n_ids * 6 * 8, 256) + gapro_align_up((size_t)n_ids * 4, 256);
}

extern "C" int gapro_instance_info(const double* xyz, const double* instance_label, const double* semantic_label,
                                   int64_t n_points, int32_t n_ids, int32_t scannet, double* boxes, double* volumes,
                                   double* classes, int32_t* n_used_dev, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GAPRO_REQUIRE(xyz && instance_label && semantic_label && boxes && volumes && classes && n_used_dev && ws,
                  "gapro_instance_info: null pointer");
    GAPRO_REQUIRE(n_points > 0 && n_points < (int64_t)INT32_MAX && n_ids > 0, "gapro_instance_info: bad sizes");
    if (ws_bytes < gapro_instance_info_workspace_bytes(n_ids)) {
        gapro_set_error("gapro_instance_info: workspace %zu < %zu bytes", ws_bytes, gapro_instance_info_workspace_bytes(n_ids));
        return GAPRO_ERR_WORKSPACE;
    }
    unsigned long long* ext = (unsigned long long*)ws;
    int32_t* first = (int32_t*)((char*)ws + gapro_align_up((size_t)n_ids * 6 * 8, 256));
    k_gi_init<<<(n_ids * 6 + 255) / 256, 256, 0, stream>>>(ext, first, n_ids);
    int dev = 0, sms = 148;
    GAPRO_CUDA_TRY(cudaGetDevice(&dev));
    GAPRO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (n_points + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    k_gi_points<<<grid, 256, 0, stream>>>(xyz, instance_label, n_points, n_ids, ext, first);
    k_gi_finish<<<1, 1024, 0, stream>>>(ext, first, semantic_label, n_ids, scannet, boxes, volumes, classes, n_used_dev);
    GAPRO_KERNEL_CHECK();
    return GAPRO_OK;
}
