// One CTA per SMALL region (M <= 64: a single 64 x 64 tile): all 50 training steps in ONE launch with the six
// 64 x 64 float64 matrices of the step resident in shared memory (209 KB) - no launch per phase, no trip through
// global memory between phases.  Included by gp_fit.cu inside its anonymous namespace (needs Region, Layout, GpParams,
// adam_update, softplus_d, sigmoid_d, hazard, chol_inv_64, dmma, the Gauss-Hermite tables).
//
// Same arithmetic as the batched path (hand-derived gradient): build K_zz / K_zx -> Cholesky + inverse
// -> A = L^-1 K_zx -> B = T^T A -> column statistics / Gauss-Hermite -> G_A -> dT (+Adam) -> dm (+Adam) ->
// G_C = L^-T G_A -> S = -sym Phi(G_A A^T) -> Y = S L^-1 -> G_K = L^-T Y -> kernel / inducing-point gradients -> Adam.
// The products run on the FP64 tensor pipe (mma.sync m8n8k4) straight from shared memory; the kernel values the
// gradient kernel needs are recomputed instead of stored (2 x 4096 exp per step).  Only the Adam moments of T and Z
// live in global memory (read-modify-written once per step, L2-resident).  Prediction goes through the batched pass
// with the trained parameters written back here.
#pragma once

constexpr int SM_LD = 68;                       // row stride (doubles): conflict-free fragment loads in both orientations
constexpr int SM_MAT = 64 * SM_LD;
constexpr int SM_THREADS = 128;
constexpr int SM_DMAX = 8;
constexpr int SM_SMEM = (6 * SM_MAT + 2 * 64 * SM_DMAX + 16 * 64 + 64) * (int)sizeof(double);

// acc[r][c][e]: warp w owns rows 16w + 8r + gid, columns 8c + 2 tig + e.
// C = opA * opB with opA(i,k) = TA ? A[k][i] : A[i][k], opB(k,j) = TB ? B[j][k] : B[k][j]; optional scale along k.
template <bool TA, bool TB>
__device__ __forceinline__ void mm64(double (&acc)[2][8][2], const double* __restrict__ A, const double* __restrict__ B,
                                     const double* __restrict__ kscale = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
        const int k = k0 + tig;
        double a[2], b[8];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = 16 * warp + 8 * r + gid;
            a[r] = TA ? A[k * SM_LD + i] : A[i * SM_LD + k];
        }
        if (kscale) {
            const double sc = kscale[k];
            a[0] *= sc;
            a[1] *= sc;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = 8 * c + gid;
            b[c] = TB ? B[j * SM_LD + k] : B[k * SM_LD + j];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
    }
}

#define SM_FOREACH(...)                                                                         \
    {                                                                                           \
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, gid_ = lane_ >> 2, tig_ = lane_ & 3; \
        _Pragma("unroll") for (int r_ = 0; r_ < 2; ++r_) {                                      \
            _Pragma("unroll") for (int c_ = 0; c_ < 8; ++c_) {                                  \
                _Pragma("unroll") for (int e_ = 0; e_ < 2; ++e_) {                              \
                    const int row = 16 * warp_ + 8 * r_ + gid_, col = 8 * c_ + 2 * tig_ + e_;   \
                    const double v = acc[r_][c_][e_];                                           \
                    (void)v;                                                                    \
                    __VA_ARGS__                                                                 \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
    }

__global__ void __launch_bounds__(SM_THREADS, 1)
k_small_fit(const Region* __restrict__ regs, int first, GpParams prm0, double lr, int iters, double* __restrict__ ws,
            int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double sm_[];
    double* sX = sm_;                 // L^-1
    double* sK = sX + SM_MAT;         // K_zx, later G_C
    double* sA = sK + SM_MAT;         // K_zz -> L, then A, later Y
    double* sB = sA + SM_MAT;         // B, later S
    double* sG = sB + SM_MAT;         // G_A, later G_K
    double* sT = sG + SM_MAT;         // T
    double* sZ = sT + SM_MAT;         // inducing points [64][D]
    double* sXf = sZ + 64 * SM_DMAX;  // training rows   [64][D]
    double* vec = sXf + 64 * SM_DMAX; // 16 vectors of 64
    double* v_m = vec, *v_y = vec + 64, *v_gmu = vec + 128, *v_gv = vec + 192, *v_dinv = vec + 256, *v_mm = vec + 320,
           *v_mv = vec + 384, *v_gs = vec + 448, *v_gl = vec + 512, *v_red = vec + 576;   // v_red: 6 x 64 scratch
    double* sc = vec + 16 * 64;       // scalars: c, rho_s, rho_l + Adam moments (SC_* layout)

    const Region R = regs[first + blockIdx.x];
    const int D = prm0.D, M = R.M;
    const Layout lay = make_layout(R.Mp, R.Np, R.Wp, D);
    double* base = ws + R.base;
    const int tid = threadIdx.x;
    const double invN = 1.0 / (double)M;

    // ---- state -> shared memory
    for (int e = tid; e < 64 * 64; e += SM_THREADS) sT[(e >> 6) * SM_LD + (e & 63)] = base[lay.T + e];
    for (int e = tid; e < 64 * D; e += SM_THREADS) {
        sZ[e] = base[lay.Z + e];
        sXf[e] = base[lay.X + e];
    }
    for (int i = tid; i < 64; i += SM_THREADS) {
        v_m[i] = base[lay.m + i];
        v_y[i] = base[lay.y + i];
        v_mm[i] = base[lay.mm + i];
        v_mv[i] = base[lay.mv + i];
    }
    if (tid < SC_N) sc[tid] = base[lay.scal + tid];
    __syncthreads();
    bool any_bad = false;

    for (int it = 1; it <= iters; ++it) {
        GpParams prm = prm0;
        prm.lr_over_bc1 = lr / (1.0 - pow(BETA1, (double)it));      // Driver::params()
        prm.bc2_sqrt = sqrt(1.0 - pow(BETA2, (double)it));
        const double ell = softplus_d(sc[SC_RL]), s = softplus_d(sc[SC_RS]);
        const double inv_l2 = 1.0 / (ell * ell);
        // ---- build: K_zz (+ jitter, identity on the padding diagonal) -> sA, K_zx -> sK, L^-1 buffer zeroed
        for (int e = tid; e < 64 * 64; e += SM_THREADS) {
            const int i = e >> 6, j = e & 63;
            double d2z = 0.0, d2x = 0.0;
            for (int d = 0; d < D; ++d) {
                const double zr = sZ[i * D + d];
                const double a = zr - sXf[j * D + d], b = zr - sZ[j * D + d];
                d2x = fma(a, a, d2x);
                d2z = fma(b, b, d2z);
            }
            const bool in = i < M && j < M;
            sK[i * SM_LD + j] = in ? s * exp(-0.5 * (d2x * inv_l2)) : 0.0;
            sA[i * SM_LD + j] = in ? s * exp(-0.5 * (d2z * inv_l2)) + (i == j ? prm.jitter_zz : 0.0) : (i == j ? 1.0 : 0.0);
            sX[i * SM_LD + j] = 0.0;
        }
        __syncthreads();
        // ---- Cholesky + inverse (L in sA, L^-1 in sX; strict upper parts are not read as zeros below: cleared)
        const bool bad = chol_inv_64<SM_LD>(sA, sX, v_dinv);
        any_bad = any_bad || bad;
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += SM_THREADS) {
            const int i = e >> 6, j = e & 63;
            if (j > i) sX[i * SM_LD + j] = 0.0;
        }
        __syncthreads();
        double acc[2][8][2];
        // ---- A = L^-1 K_zx -> sA (L is dead)
        mm64<false, false>(acc, sX, sK);
        __syncthreads();
        SM_FOREACH({ sA[row * SM_LD + col] = v; })
        __syncthreads();
        // ---- B = T^T A -> sB   (T lower; the padding diagonal of T is 1 and meets zero rows of A)
        mm64<true, false>(acc, sT, sA);
        SM_FOREACH({ sB[row * SM_LD + col] = v; })
        __syncthreads();
        // ---- column statistics + Gauss-Hermite (columns n < M): two threads per column
        {
            const int n = tid & 63, h = tid >> 6;
            double s_mu = 0.0, s_b = 0.0, s_a = 0.0;
            for (int k = h; k < 64; k += 2) {
                const double a = sA[k * SM_LD + n], b = sB[k * SM_LD + n];
                s_mu += a * v_m[k];
                s_b += b * b;
                s_a += a * a;
            }
            v_red[h * 64 + n] = s_mu;
            v_red[128 + h * 64 + n] = s_b;
            v_red[256 + h * 64 + n] = s_a;
            __syncthreads();
            if (h == 0) {
                s_mu = v_red[n] + v_red[64 + n];
                s_b = v_red[128 + n] + v_red[192 + n];
                s_a = v_red[256 + n] + v_red[320 + n];
                const double mu = s_mu + sc[SC_C];
                const double vv = s + prm.jitter_xx + s_b - s_a;
                const bool clamped = vv < MIN_VARIANCE;
                const double var = clamped ? MIN_VARIANCE : vv;
                double gmu = 0.0, gv = 0.0;
                if (n < M) {
                    const double y = v_y[n];
                    const double sd = sqrt(2.0 * var);
                    double a0 = 0.0, a1 = 0.0;
                    for (int k = 0; k < N_GH; ++k) {
                        const double hz = hazard(y * (sd * c_gh_t[k] + mu));
                        a0 += c_gh_w[k] * hz;
                        a1 += c_gh_w[k] * c_gh_t[k] * hz;
                    }
                    const double pref = -invN * 0.56418958354775628695;
                    gmu = pref * y * a0;
                    gv = clamped ? 0.0 : pref * y * a1 / sd;
                }
                v_gmu[n] = gmu;
                v_gv[n] = gv;
            }
            __syncthreads();
        }
        // ---- G_A = m g_mu^T + 2 (T B - A) diag(g_v) -> sG
        mm64<false, false>(acc, sT, sB);
        SM_FOREACH({ sG[row * SM_LD + col] = v_m[row] * v_gmu[col] + 2.0 * v_gv[col] * (v - sA[row * SM_LD + col]); })
        // ---- dT = tril(2 A diag(g_v) B^T + (T - diag(1/T_ii))/N), Adam on T (moments in global memory)
        mm64<false, true>(acc, sA, sB, v_gv);
        __syncthreads();                           // every warp has finished reading T (G_A product) before T changes
        SM_FOREACH({
            if (row < M && col <= row) {
                const size_t idx = (size_t)row * 64 + col;
                double p = sT[row * SM_LD + col];
                const double g = 2.0 * v + (p - (col == row ? 1.0 / p : 0.0)) * invN;
                double m1 = base[lay.Tm + idx], m2 = base[lay.Tv + idx];
                adam_update(p, m1, m2, g, prm);
                sT[row * SM_LD + col] = p;
                base[lay.Tm + idx] = m1;
                base[lay.Tv + idx] = m2;
            }
        })
        // ---- dm = A g_mu + m/N, Adam on m: one thread per row (two halves of the columns, combined through v_red)
        {
            const int i = tid & 63, h = tid >> 6;
            double sum = 0.0;
            for (int n = 32 * h; n < 32 * h + 32; ++n) sum += sA[i * SM_LD + n] * v_gmu[n];
            v_red[h * 64 + i] = sum;
            __syncthreads();
            if (h == 0 && i < M) {
                double p = v_m[i], m1 = v_mm[i], m2 = v_mv[i];
                const double g = (v_red[i] + v_red[64 + i]) + p * invN;
                adam_update(p, m1, m2, g, prm);
                v_red[128 + i] = p;                  // new m: published after G_A no longer needs the old one (it is in sG already)
                v_mm[i] = m1;
                v_mv[i] = m2;
            }
            __syncthreads();
            if (h == 0 && i < M) v_m[i] = v_red[128 + i];
            __syncthreads();
        }
        // ---- G_C = L^-T G_A -> sK (K_zx is dead: recomputed by the gradient below)
        mm64<true, false>(acc, sX, sG);
        SM_FOREACH({ sK[row * SM_LD + col] = v; })
        // ---- S = -1/2 tril(G_A A^T) mirrored -> sB (B is dead)
        mm64<false, true>(acc, sG, sA);
        __syncthreads();
        SM_FOREACH({
            if (col <= row) {
                const double val = -0.5 * v;
                sB[row * SM_LD + col] = val;
                sB[col * SM_LD + row] = val;
            }
        })
        __syncthreads();
        // ---- Y = S L^-1 -> sA (A is dead)
        mm64<false, false>(acc, sB, sX);
        SM_FOREACH({ sA[row * SM_LD + col] = v; })
        __syncthreads();
        // ---- G_K = L^-T Y (symmetric) -> sG (G_A is dead)
        mm64<true, false>(acc, sX, sA);
        SM_FOREACH({ sG[row * SM_LD + col] = v; })
        __syncthreads();
        // ---- kernel-parameter / inducing-point gradients: two threads per inducing row (halves of the columns)
        {
            const int i = tid & 63, h = tid >> 6;
            const double inv_s = 1.0 / s, m2_ell = -2.0 / ell;
            double az[SM_DMAX], as = 0.0, al = 0.0, cz = 0.0;
#pragma unroll
            for (int d = 0; d < SM_DMAX; ++d) az[d] = 0.0;
            if (i < M) {
                for (int j = 32 * h; j < 32 * h + 32; ++j) {
                    if (j >= M) break;
                    double d2z = 0.0, d2x = 0.0;
                    for (int d = 0; d < D; ++d) {
                        const double zi = sZ[i * D + d];
                        const double dz = zi - sZ[j * D + d], dx = zi - sXf[j * D + d];
                        d2z = fma(dz, dz, d2z);
                        d2x = fma(dx, dx, d2x);
                    }
                    const double kz = s * exp(-0.5 * (d2z * inv_l2)), kx = s * exp(-0.5 * (d2x * inv_l2));
                    const double gk = sG[i * SM_LD + j];
                    const double gc = sK[i * SM_LD + j];
                    const double grz2 = -gk * kz, grx = -0.5 * gc * kx;
                    as += (gk * kz + gc * kx) * inv_s;
                    cz += grz2 + grx;
                    al += (0.5 * grz2 * d2z + grx * d2x) * (inv_l2 * m2_ell);
#pragma unroll
                    for (int d = 0; d < SM_DMAX; ++d)
                        if (d < D) az[d] = fma(-grz2, sZ[j * D + d], fma(-grx, sXf[j * D + d], az[d]));
                }
            }
            // combine the two halves (h = 1 -> scratch -> h = 0)
            double* scratch = sB;                     // S is dead after Y
            __syncthreads();
            if (h == 1) {
                scratch[i * 16 + 0] = as;
                scratch[i * 16 + 1] = al;
                scratch[i * 16 + 2] = cz;
#pragma unroll
                for (int d = 0; d < SM_DMAX; ++d) scratch[i * 16 + 3 + d] = az[d];
            }
            __syncthreads();
            if (h == 0) {
                as += scratch[i * 16 + 0];
                al += scratch[i * 16 + 1];
                cz += scratch[i * 16 + 2];
                v_gs[i] = i < M ? as : 0.0;
                v_gl[i] = i < M ? al : 0.0;
                if (i < M) {
#pragma unroll
                    for (int d = 0; d < SM_DMAX; ++d) {
                        if (d < D) {
                            const double gz = 2.0 * inv_l2 * fma(cz, sZ[i * D + d], az[d] + scratch[i * 16 + 3 + d]);
                            scratch[1024 + i * SM_DMAX + d] = gz;
                        }
                    }
                }
            }
            __syncthreads();
            // Adam on Z (moments in global memory)
            for (int e = tid; e < M * D; e += SM_THREADS) {
                const int i2 = e / D, d = e - i2 * D;
                double p = sZ[e], m1 = base[lay.Zm + e], m2 = base[lay.Zv + e];
                adam_update(p, m1, m2, scratch[1024 + i2 * SM_DMAX + d], prm);
                base[lay.Zm + e] = m1;
                base[lay.Zv + e] = m2;
                scratch[2048 + e] = p;
            }
            __syncthreads();
            for (int e = tid; e < M * D; e += SM_THREADS) sZ[e] = scratch[2048 + e];
            // Adam on the three scalars: warp 0 reduces the four sums
            if (tid < 32) {
                double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
                for (int n = tid; n < M; n += 32) {
                    a += v_gmu[n];
                    b += v_gv[n];
                    c += v_gs[n];
                    d += v_gl[n];
                }
                for (int o = 16; o; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    d += __shfl_xor_sync(0xffffffffu, d, o);
                }
                if (tid == 0) {
                    const double g[3] = {a, (b + c) * sigmoid_d(sc[SC_RS]), d * sigmoid_d(sc[SC_RL])};
                    for (int k = 0; k < 3; ++k) {
                        double p = sc[k], m1 = sc[SC_M0 + k], m2 = sc[SC_V0 + k];
                        adam_update(p, m1, m2, g[k], prm);
                        sc[k] = p;
                        sc[SC_M0 + k] = m1;
                        sc[SC_V0 + k] = m2;
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- trained state -> workspace (the batched prediction pass reads it)
    for (int e = tid; e < 64 * 64; e += SM_THREADS) base[lay.T + e] = sT[(e >> 6) * SM_LD + (e & 63)];
    for (int e = tid; e < M * D; e += SM_THREADS) base[lay.Z + e] = sZ[e];
    for (int i = tid; i < 64; i += SM_THREADS) {
        base[lay.m + i] = v_m[i];
        base[lay.mm + i] = v_mm[i];
        base[lay.mv + i] = v_mv[i];
    }
    if (tid < SC_N) base[lay.scal + tid] = sc[tid];
    if (any_bad && tid == 0) atomicOr(status + R.orig, GAPRO_GP_NOT_PSD);
}
