// A + A' in POINT order (sm_100a) - containment (gen_ps_utils.py:349-351) and occupancy pooling (:359-363)
// without the by-superpoint gather of k_occupancy:
//
//   k_grid_build    per scene a 32 x 32 x 8 grid over the scene extent; per cell two box bit masks:
//                   `full` = boxes (with the 0.005 margin) that contain the whole cell, `cand` = boxes that
//                   touch it.  Exact: a point of the cell is inside every `full` box and outside every
//                   non-`cand` box, whatever its position in the cell (conjunction of per-axis intervals).
//   k_occ_points    streams xyz (24 B) + dense superpoint id (4 B) in point order - coalesced, no permutation -
//                   looks its cell up (the masks are L2-resident), runs the exact float64 test of
//                   is_within_bb_torch only for the boxes whose boundary crosses the cell, and counts with integer
//                   `red.global.add` into the (superpoint x box) table, which lives in L2.  Integer, order-free,
//                   therefore bit-exact.
//   k_occ_finalize  one warp per superpoint: count / size >= thresh exactly as the reference evaluates it in
//                   float32 (one correctly rounded divide), occupancy bits, n_bbs, and the per-box exclusive
//                   counts / per-scene intersection counts the host state machine needs.
//
// Algorithmic HBM bytes: N * (24 + 4) + 48 * B + 4 * S * words + 4 * S - the same as the gather kernel, but
// read as a stream.
#include "common.cuh"

#define FULL_MASK 0xffffffffu

namespace {

constexpr int GX = 32, GY = 32, GZ = 8, CELLS = GX * GY * GZ;

struct SceneGrid {
    double x0, y0, z0, inv_dx, inv_dy, inv_dz;
};

__device__ __forceinline__ double ordered_to_dbl_(unsigned long long u) {
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

// one thread per cell
template <int WORDS>
__global__ void __launch_bounds__(256)
k_grid_build(const unsigned long long* __restrict__ extent_keys, const int32_t* __restrict__ box_off,
             const double* __restrict__ boxes, double margin, SceneGrid* __restrict__ grids,
             uint32_t* __restrict__ full_mask, uint32_t* __restrict__ cand_mask) {
    const int sc = blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    double lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = ordered_to_dbl_(extent_keys[sc * 6 + d]);
        hi[d] = ordered_to_dbl_(extent_keys[sc * 6 + 3 + d]);
    }
    const int n[3] = {GX, GY, GZ};
    double inv[3], step[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double ext = hi[d] - lo[d];
        step[d] = ext > 0.0 ? ext / n[d] : 1.0;
        inv[d] = 1.0 / step[d];
    }
    if (cell == 0) grids[sc] = SceneGrid{lo[0], lo[1], lo[2], inv[0], inv[1], inv[2]};
    if (cell >= CELLS) return;
    const int ci[3] = {cell % GX, (cell / GX) % GY, cell / (GX * GY)};
    // cell bounds, inflated by 1e-6 of a cell (the index of a point is computed with a few ulps of error) and open
    // at the ends of the grid (indices are clamped)
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double clo[3], chi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        clo[d] = ci[d] == 0 ? -INF : lo[d] + (ci[d] - 1e-6) * step[d];
        chi[d] = ci[d] == n[d] - 1 ? INF : lo[d] + (ci[d] + 1 + 1e-6) * step[d];
    }
    const int b0 = box_off[sc], nb = box_off[sc + 1] - b0;
    uint32_t fm[WORDS], cm[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) fm[w] = cm[w] = 0u;
    for (int b = 0; b < nb; ++b) {
        const double* bx = boxes + 6 * (size_t)(b0 + b);      // warp-uniform: broadcast loads
        bool full = true, touch = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double l = __dsub_rn(bx[d], margin), h = __dadd_rn(bx[3 + d], margin);
            full = full && l <= clo[d] && h >= chi[d];
            touch = touch && !(h < clo[d]) && !(l > chi[d]);
        }
        if (full) fm[b >> 5] |= 1u << (b & 31);
        if (touch) cm[b >> 5] |= 1u << (b & 31);
    }
    uint32_t* f = full_mask + ((size_t)sc * CELLS + cell) * WORDS;
    uint32_t* c = cand_mask + ((size_t)sc * CELLS + cell) * WORDS;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
        f[w] = fm[w];
        c[w] = cm[w] & ~fm[w];      // boxes that need the exact test
    }
}

constexpr int OCC_T = 256, OCC_PER = 4;      // points per thread (their mask loads are issued together)

template <int WORDS>
__global__ void __launch_bounds__(OCC_T, 4)
k_occ_points(const double* __restrict__ xyz, const int32_t* __restrict__ spp_gid, const int64_t* __restrict__ pt_off,
             const int32_t* __restrict__ box_off, const double* __restrict__ boxes, const SceneGrid* __restrict__ grids,
             const uint32_t* __restrict__ full_mask, const uint32_t* __restrict__ cand_mask, int n_scenes, int64_t n,
             double margin, int32_t* __restrict__ cnt_table) {
    __shared__ double s_xyz[OCC_T * OCC_PER * 3];
    const int64_t base = (int64_t)blockIdx.x * (OCC_T * OCC_PER);
    const int64_t left = n - base;
    const int npts = left < OCC_T * OCC_PER ? (int)left : OCC_T * OCC_PER;
    // coalesced 16-byte loads of the chunk's 3 * npts doubles (base * 24 bytes is 16-byte aligned: base % 1024 == 0)
    {
        const double2* src = reinterpret_cast<const double2*>(xyz + 3 * base);
        const int n2 = (3 * npts) >> 1;
        for (int i = threadIdx.x; i < n2; i += OCC_T) {
            const double2 v = __ldcs(src + i);
            s_xyz[2 * i] = v.x;
            s_xyz[2 * i + 1] = v.y;
        }
        if (((3 * npts) & 1) && threadIdx.x == 0) s_xyz[3 * npts - 1] = xyz[3 * base + 3 * npts - 1];
    }
    int gid[OCC_PER];
#pragma unroll
    for (int q = 0; q < OCC_PER; ++q) {
        const int k = q * OCC_T + threadIdx.x;
        gid[q] = k < npts ? __ldcs(spp_gid + base + k) : -1;
    }
    __syncthreads();
    // scene of the chunk's first point; later points only move forward
    const int sc = gapro_find_segment<int64_t>(pt_off, n_scenes, base);
    constexpr int stride = 32 * WORDS;
    // phase 1: cell of every point of this thread and its two masks - all loads in flight together
    int scn[OCC_PER];
    uint32_t fm[OCC_PER][WORDS], cm[OCC_PER][WORDS];
#pragma unroll
    for (int q = 0; q < OCC_PER; ++q) {
        const int k = q * OCC_T + threadIdx.x;
        scn[q] = -1;
        if (k >= npts) continue;
        const int64_t p = base + k;
        int s = sc;
        while (p >= pt_off[s + 1]) ++s;
        scn[q] = s;
        const double x = s_xyz[3 * k], y = s_xyz[3 * k + 1], z = s_xyz[3 * k + 2];
        const SceneGrid G = grids[s];
        int ix = (int)((x - G.x0) * G.inv_dx), iy = (int)((y - G.y0) * G.inv_dy), iz = (int)((z - G.z0) * G.inv_dz);
        ix = ix < 0 ? 0 : (ix > GX - 1 ? GX - 1 : ix);
        iy = iy < 0 ? 0 : (iy > GY - 1 ? GY - 1 : iy);
        iz = iz < 0 ? 0 : (iz > GZ - 1 ? GZ - 1 : iz);
        const size_t cell = ((size_t)s * CELLS + (size_t)(iz * GY + iy) * GX + ix) * WORDS;
#pragma unroll
        for (int w = 0; w < WORDS; ++w) {
            fm[q][w] = __ldg(full_mask + cell + w);
            cm[q][w] = __ldg(cand_mask + cell + w);
        }
    }
    // phase 2: exact float64 test for the boxes whose boundary crosses the cell, then the integer reductions
#pragma unroll
    for (int q = 0; q < OCC_PER; ++q) {
        if (scn[q] < 0) continue;
        const int k = q * OCC_T + threadIdx.x;
        const double x = s_xyz[3 * k], y = s_xyz[3 * k + 1], z = s_xyz[3 * k + 2];
        const int b0 = box_off[scn[q]];
        int32_t* row = cnt_table + (size_t)gid[q] * stride;
#pragma unroll
        for (int w = 0; w < WORDS; ++w) {
            uint32_t m = fm[q][w];
            uint32_t c = cm[q][w];
            while (c) {
                const int bb = __ffs(c) - 1;
                c &= c - 1;
                const double* bx = boxes + 6 * (size_t)(b0 + 32 * w + bb);
                // margins in float64 on the float64 boxes (gen_ps_utils.py:350)
                const bool in = x >= __dsub_rn(__ldg(bx), margin) && y >= __dsub_rn(__ldg(bx + 1), margin) &&
                                z >= __dsub_rn(__ldg(bx + 2), margin) && x <= __dadd_rn(__ldg(bx + 3), margin) &&
                                y <= __dadd_rn(__ldg(bx + 4), margin) && z <= __dadd_rn(__ldg(bx + 5), margin);
                m |= (uint32_t)in << bb;
            }
            while (m) {
                const int bb = __ffs(m) - 1;
                m &= m - 1;
                atomicAdd(row + 32 * w + bb, 1);       // result unused: compiles to RED
            }
        }
    }
}

__device__ __forceinline__ int warp_find_scene_(const int32_t* __restrict__ spp_off, int n_scenes, int g, int lane) {
    int sc = -1;
    for (int base = 0; base <= n_scenes; base += 32) {
        const int i = base + lane;
        const bool le = (i <= n_scenes) && (spp_off[i] <= g);
        sc += __popc(__ballot_sync(FULL_MASK, le));
    }
    return sc;
}

template <int WORDS>
__global__ void __launch_bounds__(256)
k_occ_finalize(const int32_t* __restrict__ cnt_table, const int32_t* __restrict__ seg_off, const int32_t* __restrict__ spp_off,
               const int32_t* __restrict__ box_off, int n_scenes, int s_total, float thresh,
               uint32_t* __restrict__ occ_bits, int32_t* __restrict__ n_bbs, int32_t* __restrict__ excl_cnt,
               int32_t* __restrict__ inter_cnt) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= s_total) return;
    const int cnt = seg_off[g + 1] - seg_off[g];
    const int sc = warp_find_scene_(spp_off, n_scenes, g, lane);
    const int b0 = box_off[sc];
    const int nb = box_off[sc + 1] - b0;
    const float fcnt = (float)cnt;
    uint32_t bits[WORDS];
    int total = 0;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
        const int b = 32 * w + lane;
        const int count = cnt_table[((size_t)g * WORDS + w) * 32 + lane];
        // scatter-mean of 0/1 in float32 (exact sum, one correctly rounded divide), then >= thresh (:359-362)
        const bool occ = (b < nb) && (__fdiv_rn((float)count, fcnt) >= thresh);
        bits[w] = __ballot_sync(FULL_MASK, occ);
        total += __popc(bits[w]);
        if (lane == 0) occ_bits[(size_t)g * WORDS + w] = bits[w];
    }
    if (lane == 0) {
        n_bbs[g] = total;
        if (total == 1) {
#pragma unroll
            for (int w = 0; w < WORDS; ++w)
                if (bits[w]) atomicAdd(excl_cnt + b0 + 32 * w + __ffs(bits[w]) - 1, 1);
        } else if (total >= 2) {
            const int stride = 32 * WORDS;
            int32_t* ic = inter_cnt + (size_t)sc * stride * stride;
#pragma unroll
            for (int w1 = 0; w1 < WORDS; ++w1) {
                uint32_t m1 = bits[w1];
                while (m1) {
                    const int i1 = 32 * w1 + __ffs(m1) - 1;
                    m1 &= m1 - 1;
#pragma unroll
                    for (int w2 = 0; w2 < WORDS; ++w2) {
                        if (w2 < w1) continue;
                        uint32_t m2 = bits[w2];
                        while (m2) {
                            const int i2 = 32 * w2 + __ffs(m2) - 1;
                            m2 &= m2 - 1;
                            if (i2 > i1) atomicAdd(ic + (size_t)i1 * stride + i2, 1);
                        }
                    }
                }
            }
        }
    }
}

struct OccWs {
    size_t grids, full, cand, total;
};
OccWs occ_layout(int n_scenes, int words) {
    OccWs w;
    size_t o = 0;
    w.grids = o;
    o += gapro_align_up((size_t)n_scenes * sizeof(SceneGrid), 256);
    w.full = o;
    o += gapro_align_up((size_t)n_scenes * CELLS * words * 4, 256);
    w.cand = o;
    o += gapro_align_up((size_t)n_scenes * CELLS * words * 4, 256);
    w.total = o;
    return w;
}

}  // namespace

extern "C" size_t gapro_occupancy_points_workspace_bytes(int32_t n_scenes, int32_t words) {
    if (n_scenes <= 0 || words <= 0) return 0;
    return occ_layout(n_scenes, words).total;
}

extern "C" int gapro_occupancy_points(const double* xyz, const int32_t* spp_gid, const int32_t* seg_off,
                                      const int64_t* pt_off_dev, const int32_t* spp_off_dev, const int32_t* box_off_dev,
                                      const double* boxes, const uint64_t* extent_keys, int32_t n_scenes,
                                      int64_t n_points, int32_t s_total, int32_t n_boxes, int32_t words, double margin,
                                      float thresh, uint32_t* occ_bits, int32_t* n_bbs, int32_t* cnt_table,
                                      int32_t* excl_cnt, int32_t* inter_cnt, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GAPRO_REQUIRE(xyz && spp_gid && seg_off && pt_off_dev && spp_off_dev && box_off_dev && boxes && extent_keys &&
                      occ_bits && n_bbs && cnt_table && excl_cnt && inter_cnt && ws,
                  "gapro_occupancy_points: null pointer");
    GAPRO_REQUIRE(n_scenes > 0 && s_total > 0 && n_boxes > 0 && n_points > 0, "gapro_occupancy_points: empty batch");
    GAPRO_REQUIRE(words == 1 || words == 2 || words == 4 || words == 8,
                  "gapro_occupancy_points: words must be 1, 2, 4 or 8 (got %d; at most 256 boxes per scene)", words);
    GAPRO_REQUIRE(((uintptr_t)xyz & 15) == 0, "gapro_occupancy_points: xyz must be 16-byte aligned");
    const OccWs L = occ_layout(n_scenes, words);
    if (ws_bytes < L.total) {
        gapro_set_error("gapro_occupancy_points: workspace %zu < %zu bytes", ws_bytes, L.total);
        return GAPRO_ERR_WORKSPACE;
    }
    SceneGrid* grids = (SceneGrid*)((char*)ws + L.grids);
    uint32_t* full = (uint32_t*)((char*)ws + L.full);
    uint32_t* cand = (uint32_t*)((char*)ws + L.cand);
    GAPRO_CUDA_TRY(cudaMemsetAsync(excl_cnt, 0, (size_t)n_boxes * 4, stream));
    GAPRO_CUDA_TRY(cudaMemsetAsync(inter_cnt, 0, (size_t)n_scenes * 32 * words * 32 * words * 4, stream));
    GAPRO_CUDA_TRY(cudaMemsetAsync(cnt_table, 0, (size_t)s_total * 32 * words * 4, stream));
    const dim3 gg(CELLS / 256, n_scenes);
    const unsigned gp = (unsigned)((n_points + OCC_T * OCC_PER - 1) / (OCC_T * OCC_PER));
    const unsigned gf = (unsigned)((s_total + 7) / 8);
#define LAUNCH_OCC(W)                                                                                                  \
    k_grid_build<W><<<gg, 256, 0, stream>>>((const unsigned long long*)extent_keys, box_off_dev, boxes, margin, grids,  \
                                            full, cand);                                                               \
    k_occ_points<W><<<gp, OCC_T, 0, stream>>>(xyz, spp_gid, pt_off_dev, box_off_dev, boxes, grids, full, cand, n_scenes, \
                                              n_points, margin, cnt_table);                                            \
    k_occ_finalize<W><<<gf, 256, 0, stream>>>(cnt_table, seg_off, spp_off_dev, box_off_dev, n_scenes, s_total, thresh,  \
                                              occ_bits, n_bbs, excl_cnt, inter_cnt)
    switch (words) {
        case 1: LAUNCH_OCC(1); break;
        case 2: LAUNCH_OCC(2); break;
        case 4: LAUNCH_OCC(4); break;
        default: LAUNCH_OCC(8); break;
    }
#undef LAUNCH_OCC
    GAPRO_KERNEL_CHECK();
    return GAPRO_OK;
}
