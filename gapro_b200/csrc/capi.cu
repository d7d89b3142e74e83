// Error plumbing and the HOST-side pair state machine of the gapro_b200 C ABI.
#include <math.h>
#include <stdarg.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void gapro_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int gapro_version(void) { return 100; }
extern "C" const char* gapro_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// IoU of axis-aligned boxes in float64 — the formula of batch_giou_cross
// (/root/reference/gapro/gen_ps_utils.py:39-50): clamp(min(hi)-max(lo), 0) product over the
// three axes, union + 1e-6 in the denominator.  Diagonal zeroed (:386).
// Compiled without FMA contraction of the products (see build flags: -fmad=false only affects
// device code; host code is plain IEEE double and gcc does not contract across statements here).
static inline double clamp0(double x) { return x > 0.0 ? x : 0.0; }

static double box_iou(const double* a, const double* b) {
    double ex = clamp0(fmin(a[3], b[3]) - fmax(a[0], b[0]));
    double ey = clamp0(fmin(a[4], b[4]) - fmax(a[1], b[1]));
    double ez = clamp0(fmin(a[5], b[5]) - fmax(a[2], b[2]));
    volatile double inter = ex * ey;
    inter = inter * ez;
    volatile double va = clamp0(a[3] - a[0]) * clamp0(a[4] - a[1]);
    va = va * clamp0(a[5] - a[2]);
    volatile double vb = clamp0(b[3] - b[0]) * clamp0(b[4] - b[1]);
    vb = vb * clamp0(b[5] - b[2]);
    volatile double uni = va + vb;
    uni = uni - inter;
    return inter / (uni + 1e-6);
}

extern "C" int gapro_box_iou(const double* boxes, int32_t B, double* iou) {
    GAPRO_REQUIRE(boxes && iou && B >= 0, "gapro_box_iou: bad arguments");
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < B; ++j) iou[(size_t)i * B + j] = (i == j) ? 0.0 : box_iou(boxes + 6 * i, boxes + 6 * j);
    return GAPRO_OK;
}

// is_box1_in_box2 (/root/reference/gapro/gen_ps_utils.py:75-76), float64
static bool box_in_box(const double* b1, const double* b2, double offset) {
    for (int d = 0; d < 3; ++d) {
        volatile double lo = b1[d] + offset;
        volatile double hi = b1[3 + d] - offset;
        if (!(lo >= b2[d])) return false;
        if (!(hi <= b2[3 + d])) return false;
    }
    return true;
}

// The (b1, b2) walk of /root/reference/gapro/gen_ps_utils.py:388-448 without the GP calls.
extern "C" int gapro_enumerate_events(const double* boxes, int32_t B, const int32_t* excl_cnt,
                                      const int32_t* inter_cnt, int32_t stride, int32_t* ev_kind, int32_t* ev_b1,
                                      int32_t* ev_b2, int32_t capacity) {
    GAPRO_REQUIRE(boxes && excl_cnt && inter_cnt && ev_kind && ev_b1 && ev_b2, "gapro_enumerate_events: null pointer");
    GAPRO_REQUIRE(B >= 0 && stride >= B, "gapro_enumerate_events: stride %d < B %d", stride, B);
    std::vector<double> iou((size_t)B * B);
    gapro_box_iou(boxes, B, iou.data());
    std::vector<char> visited(B, 0);
    std::vector<int> overlap;
    int n_ev = 0;
    auto push = [&](int kind, int b1, int b2) -> bool {
        if (n_ev >= capacity) return false;
        ev_kind[n_ev] = kind;
        ev_b1[n_ev] = b1;
        ev_b2[n_ev] = b2;
        ++n_ev;
        return true;
    };
    for (int b1 = 0; b1 < B; ++b1) {
        overlap.clear();
        for (int b2 = 0; b2 < B; ++b2)
            if (iou[(size_t)b1 * B + b2] > 0.0001 && !visited[b2]) overlap.push_back(b2);   // :393
        if (overlap.empty()) {
            visited[b1] = 1;
            continue;
        }
        for (int b2 : overlap) {
            int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
            if (inter_cnt[(size_t)lo * stride + hi] == 0) continue;                          // :408
            if (box_in_box(boxes + 6 * b1, boxes + 6 * b2, 0.1)) {                           // :411-416
                if (!push(GAPRO_EV_NEST_B1, b1, b2)) goto overflow;
                visited[b1] = 1;
                break;
            }
            if (box_in_box(boxes + 6 * b2, boxes + 6 * b1, 0.1)) {                           // :418-423
                if (!push(GAPRO_EV_NEST_B2, b1, b2)) goto overflow;
                visited[b2] = 1;
                continue;
            }
            if (iou[(size_t)b1 * B + b2] >= 0.6) continue;                                   // :425
            if (excl_cnt[b1] == 0 || excl_cnt[b2] == 0) continue;                            // :431
            if (!push(GAPRO_EV_GP, b1, b2)) goto overflow;
        }
        visited[b1] = 1;
    }
    return n_ev;
overflow:
    gapro_set_error("gapro_enumerate_events: more than %d events", capacity);
    return GAPRO_ERR_CAPACITY;
}
