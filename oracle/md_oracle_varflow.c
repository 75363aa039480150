/*
 * md_oracle_varflow.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the dense variational flow:
 *   OpticalFlowCalculator::varFlow       common/src/optical_flow_calculator.cpp:417-464 (parameters)
 *   VarFlow::CalcFlow                    common/src/VarFlow.cpp:600-697
 *   VarFlow::gauss_seidel_recursive      common/src/VarFlow.cpp:508-584
 *   VarFlow::gauss_seidel_iteration/step common/src/VarFlow.cpp:231-348
 *   VarFlow::calculate_residual / residual_part_step   common/src/VarFlow.cpp:373-490
 * The in-tree loops are followed statement by statement, INCLUDING the array aliasing of the residual
 * buffers when they are handed down as J13/J23 (VarFlow.cpp:537) -- "transliterate, do not fix".
 * The legacy OpenCV C calls (cvSmooth / cvFilter2D / cvResize / cvAddWeighted) are restated from the
 * OpenCV algorithm (BORDER_REPLICATE Gaussian of cvRound(8*sigma+1)|1 taps, correlation, half-pixel
 * bilinear resize with the exact-2x area shortcut) and pinned against cv2 4.13.0 in
 * tests/test_oracle_vs_cv2.py.  Float arithmetic, no FMA contraction (build with -ffp-contract=off).
 * PINNED TO THE REFERENCE'S OWN CODE: oracle/_ref/libvarflow_ref.so is the unmodified common/src/VarFlow.cpp compiled
 * against the legacy-C-API shim of oracle/ref_shim (same Gaussian / resize / filter primitives as here); orc_varflow equals
 * it bit for bit for every size and parameter set tried (tests/test_oracle_ref.py, frozen in tests/golden/golden_ref.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "md_oracle.h"

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cv::getGaussianKernel(n, sigma, CV_32F) */
static int gauss_kernel(double sigma, float *k /* >= 64 */)
{
    int n = (int)lrint(sigma * 4 * 2 + 1) | 1;
    double sum = 0, t[64];
    double scale2x = -0.5 / (sigma * sigma);
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) k[i] = (float)(t[i] * sum);
    return n;
}

/* cvSmooth(CV_GAUSSIAN, 0, 0, sigma) on 32F: separable, BORDER_REPLICATE.  Row pass sums k = 0..n-1 in order,
 * column pass uses the symmetric folded form (centre first, then pairs) like cv::SymmColumnFilter. */
void orc_gaussian_blur_f32(const float *src, int w, int h, float *dst, double sigma)
{
    float k[64];
    int n = gauss_kernel(sigma, k), r = n / 2;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const float *s = src + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float acc = k[0] * s[clampi(x - r, 0, w - 1)];
            for (int i = 1; i < n; i++) acc += k[i] * s[clampi(x + i - r, 0, w - 1)];
            tmp[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float acc = k[r] * tmp[(size_t)y * w + x];
            for (int i = 1; i <= r; i++)
                acc += k[r + i] * (tmp[(size_t)clampi(y + i, 0, h - 1) * w + x] + tmp[(size_t)clampi(y - i, 0, h - 1) * w + x]);
            dst[(size_t)y * w + x] = acc;
        }
    free(tmp);
}

/* cvResize(CV_INTER_LINEAR) on 32FC1: exact 2x2 mean when both scales are exactly 2, otherwise
 * half-pixel-centre bilinear with edge clamp (cv::resize INTER_LINEAR). */
void orc_resize_linear_f32(const float *src, int sw, int sh, float *dst, int dw, int dh)
{
    double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    if (sw == dw && sh == dh) { memcpy(dst, src, sizeof(float) * (size_t)sw * sh); return; }
    if (scale_x == 2.0 && scale_y == 2.0) {
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const float *s0 = src + (size_t)(2 * y) * sw + 2 * x, *s1 = s0 + sw;
                dst[(size_t)y * dw + x] = (s0[0] + s0[1] + s1[0] + s1[1]) * 0.25f;
            }
        return;
    }
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    float *xa = (float *)malloc(sizeof(float) * (size_t)dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx; xa[dx] = fx;
    }
    float *r0 = (float *)malloc(sizeof(float) * (size_t)dw * 2), *r1 = r0 + dw;
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        if (sy < 0) { fy = 0; sy = 0; }
        if (sy >= sh - 1) { fy = 0; sy = sh - 1; }
        int sy1 = sy + 1 < sh ? sy + 1 : sh - 1;
        const float *S0 = src + (size_t)sy * sw, *S1 = src + (size_t)sy1 * sw;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            float a1 = xa[dx], a0 = 1.f - a1;
            r0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            r1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        float b1 = fy, b0 = 1.f - fy;
        for (int dx = 0; dx < dw; dx++) dst[(size_t)dy * dw + dx] = r0[dx] * b0 + r1[dx] * b1;
    }
    free(xofs); free(xa); free(r0);
}

/* ------------------------------------------------------------------------------------------ */

typedef struct { int w, h; float *d; } fimg;

typedef struct {
    int nlev, n1, n2;
    float alpha;
    int literal;
    fimg *J11, *J12, *J13, *J22, *J23, *U, *V, *Ur, *Vr;
} vf_state;

/* VarFlow::gauss_seidel_step, VarFlow.cpp:231-285 */
static inline float gs_step(const fimg *u, int x, int y, float h, float alpha, float J11, float J12, float J13, float vi)
{
    int n = 0;
    float t = 0;
    if (y - 1 > -1) { t += u->d[(size_t)(y - 1) * u->w + x]; n++; }
    if (y + 1 < u->h) { t += u->d[(size_t)(y + 1) * u->w + x]; n++; }
    if (x - 1 > -1) { t += u->d[(size_t)y * u->w + x - 1]; n++; }
    if (x + 1 < u->w) { t += u->d[(size_t)y * u->w + x + 1]; n++; }
    t = t - (h * h / alpha) * (J12 * vi + J13);
    t = t / (n + (h * h / alpha) * J11);
    return t;
}

/* VarFlow::gauss_seidel_iteration, VarFlow.cpp:298-348 */
static void gs_iteration(vf_state *s, int lvl, float h, int iters, fimg *J13a, fimg *J23a)
{
    fimg *U = &s->U[lvl], *V = &s->V[lvl];
    const float *fxfx = s->J11[lvl].d, *fxfy = s->J12[lvl].d, *fxft = J13a[lvl].d, *fyfy = s->J22[lvl].d, *fyft = J23a[lvl].d;
    int w = U->w, hh = U->h;
    for (int k = 0; k < iters; k++) {
        size_t i = 0;
        for (int y = 0; y < hh; y++)
            for (int x = 0; x < w; x++, i++) {
                U->d[i] = gs_step(U, x, y, h, s->alpha, fxfx[i], fxfy[i], fxft[i], V->d[i]);
                V->d[i] = gs_step(V, x, y, h, s->alpha, fyfy[i], fxfy[i], fyft[i], U->d[i]);
            }
    }
}

/* VarFlow::residual_part_step, VarFlow.cpp:373-429 */
static inline float res_step(const fimg *u, int x, int y, float h, float alpha, float J11, float J12, float vi)
{
    float ih2 = 1 / (h * h);
    float t = 0;
    int n = 0;
    float cu = u->d[(size_t)y * u->w + x];
    if (y - 1 > -1) { t += u->d[(size_t)(y - 1) * u->w + x]; n++; }
    if (y + 1 < u->h) { t += u->d[(size_t)(y + 1) * u->w + x]; n++; }
    if (x - 1 > -1) { t += u->d[(size_t)y * u->w + x - 1]; n++; }
    if (x + 1 < u->w) { t += u->d[(size_t)y * u->w + x + 1]; n++; }
    t = n * cu - t;
    t *= ih2;
    t -= (1 / alpha) * (J11 * cu + J12 * vi);
    return t;
}

/* VarFlow::calculate_residual, VarFlow.cpp:441-490 (cvAddWeighted = a*(1/alpha) + r*(-1) + 0 in f32) */
static void calc_residual(vf_state *s, int lvl, float h, fimg *J13a, fimg *J23a)
{
    fimg *U = &s->U[lvl], *V = &s->V[lvl];
    const float *fxfx = s->J11[lvl].d, *fxfy = s->J12[lvl].d, *fyfy = s->J22[lvl].d;
    float *ur = s->Ur[lvl].d, *vr = s->Vr[lvl].d;
    const float *fxft = J13a[lvl].d, *fyft = J23a[lvl].d;     /* may alias ur / vr (VarFlow.cpp:537) */
    int w = U->w, hh = U->h;
    size_t i = 0;
    for (int y = 0; y < hh; y++)
        for (int x = 0; x < w; x++, i++) {
            ur[i] = res_step(U, x, y, h, s->alpha, fxfx[i], fxfy[i], V->d[i]);
            vr[i] = res_step(V, x, y, h, s->alpha, fyfy[i], fxfy[i], U->d[i]);
        }
    float ia = 1 / s->alpha;
    size_t n = (size_t)w * hh;
    for (i = 0; i < n; i++) ur[i] = fxft[i] * ia + ur[i] * -1.f + 0.f;
    for (i = 0; i < n; i++) vr[i] = fyft[i] * ia + vr[i] * -1.f + 0.f;
}

static void resize_img(const fimg *a, fimg *b) { orc_resize_linear_f32(a->d, a->w, a->h, b->d, b->w, b->h); }
static void zero_img(fimg *a) { memset(a->d, 0, sizeof(float) * (size_t)a->w * a->h); }
static void add_img(fimg *a, const fimg *b) { size_t n = (size_t)a->w * a->h; for (size_t i = 0; i < n; i++) a->d[i] = a->d[i] + b->d[i]; }

/* VarFlow::gauss_seidel_recursive, VarFlow.cpp:508-584 */
static void gs_recursive(vf_state *s, int lvl, int max_level, float h, fimg *J13a, fimg *J23a)
{
    if (lvl == max_level) { gs_iteration(s, lvl, h, s->n1, J13a, J23a); return; }
    gs_iteration(s, lvl, h, s->n1, J13a, J23a);
    for (int cyc = 0; cyc < 2; cyc++) {
        if (s->literal) {
            calc_residual(s, lvl, h, J13a, J23a);
            resize_img(&s->Ur[lvl], &s->Ur[lvl + 1]);
            resize_img(&s->Vr[lvl], &s->Vr[lvl + 1]);
            zero_img(&s->U[lvl + 1]);
            zero_img(&s->V[lvl + 1]);
            gs_recursive(s, lvl + 1, max_level, 2 * h, s->Ur, s->Vr);
            resize_img(&s->U[lvl + 1], &s->Ur[lvl]);
            resize_img(&s->V[lvl + 1], &s->Vr[lvl]);
            add_img(&s->U[lvl], &s->Ur[lvl]);
            add_img(&s->V[lvl], &s->Vr[lvl]);
        }
        gs_iteration(s, lvl, h, cyc == 0 ? s->n1 + s->n2 : s->n2, J13a, J23a);
    }
}

static fimg *alloc_pyr(int w, int h, int n)
{
    fimg *p = (fimg *)malloc(sizeof(fimg) * (size_t)n);
    for (int i = 0; i < n; i++) {
        p[i].w = (int)floor(w / pow(2.0, (double)i));
        p[i].h = (int)floor(h / pow(2.0, (double)i));
        p[i].d = (float *)calloc((size_t)p[i].w * p[i].h, sizeof(float));
    }
    return p;
}
static void free_pyr_f(fimg *p, int n) { for (int i = 0; i < n; i++) free(p[i].d); free(p); }

/* VarFlow ctor (VarFlow.cpp:27-162) + CalcFlow(imgA, imgB, imgU, imgV, saved_data = 0) (VarFlow.cpp:600-697).
 * U is +x, V is y-UP (mask_y sign, VarFlow.cpp:103-107).  Only start_level == 0 (the reference's constant,
 * optical_flow_calculator.cpp:423) is restated.  literal_corrections = 0 drops the residual/restrict/
 * recurse/prolong steps and is used only to measure how inert they are. */
int orc_varflow(const uint8_t *A, const uint8_t *B, int w, int h, int pitch, int max_level, int start_level,
                int n1, int n2, float rho, float alpha, float sigma, float *Uo, float *Vo, int literal_corrections)
{
    if (start_level != 0 || w < 1 || h < 1) return -1;
    while (max_level > 0 && ((int)floor(w / pow(2.0, (double)max_level)) < 1 || (int)floor(h / pow(2.0, (double)max_level)) < 1)) max_level--;
    int nl = max_level + 1;
    size_t n = (size_t)w * h;
    float *Af = (float *)malloc(sizeof(float) * n), *Bf = (float *)malloc(sizeof(float) * n);
    float *fx = (float *)malloc(sizeof(float) * n), *fy = (float *)malloc(sizeof(float) * n), *ft = (float *)malloc(sizeof(float) * n);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            Af[(size_t)y * w + x] = (float)A[(size_t)y * pitch + x];
            Bf[(size_t)y * w + x] = (float)B[(size_t)y * pitch + x];
        }
    orc_gaussian_blur_f32(Af, w, h, Af, sigma);
    orc_gaussian_blur_f32(Bf, w, h, Bf, sigma);
    /* cvFilter2D = correlation, anchor centre, BORDER_REPLICATE; masks VarFlow.cpp:97-107 (zero tap skipped) */
    const float mx[5] = {0.08333f, -0.66666f, 0.f, 0.66666f, -0.08333f};
    const float my[5] = {-0.08333f, 0.66666f, 0.f, -0.66666f, 0.08333f};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float sx = 0, sy = 0;
            for (int i = 0; i < 5; i++) {
                if (i == 2) continue;
                sx += mx[i] * Af[(size_t)y * w + clampi(x + i - 2, 0, w - 1)];
                sy += my[i] * Af[(size_t)clampi(y + i - 2, 0, h - 1) * w + x];
            }
            fx[(size_t)y * w + x] = sx;
            fy[(size_t)y * w + x] = sy;
            ft[(size_t)y * w + x] = Bf[(size_t)y * w + x] - Af[(size_t)y * w + x];
        }
    vf_state s;
    s.nlev = nl; s.n1 = n1; s.n2 = n2; s.alpha = alpha; s.literal = literal_corrections;
    s.J11 = alloc_pyr(w, h, nl); s.J12 = alloc_pyr(w, h, nl); s.J13 = alloc_pyr(w, h, nl);
    s.J22 = alloc_pyr(w, h, nl); s.J23 = alloc_pyr(w, h, nl);
    s.U = alloc_pyr(w, h, nl); s.V = alloc_pyr(w, h, nl); s.Ur = alloc_pyr(w, h, nl); s.Vr = alloc_pyr(w, h, nl);
    for (size_t i = 0; i < n; i++) {
        s.J11[0].d[i] = fx[i] * fx[i];
        s.J12[0].d[i] = fx[i] * fy[i];
        s.J13[0].d[i] = fx[i] * ft[i];
        s.J22[0].d[i] = fy[i] * fy[i];
        s.J23[0].d[i] = fy[i] * ft[i];
    }
    fimg *Js[5] = {s.J11, s.J12, s.J13, s.J22, s.J23};
    for (int j = 0; j < 5; j++) orc_gaussian_blur_f32(Js[j][0].d, w, h, Js[j][0].d, rho);
    for (int i = 1; i < nl; i++)
        for (int j = 0; j < 5; j++) resize_img(&Js[j][i - 1], &Js[j][i]);

    int k = max_level;
    for (;;) {
        gs_recursive(&s, k, max_level, (float)pow(2.0, (double)k), s.J13, s.J23);
        if (k > 0) {
            resize_img(&s.U[k], &s.U[k - 1]);
            resize_img(&s.V[k], &s.V[k - 1]);
            k--;
        } else break;
    }
    memcpy(Uo, s.U[0].d, sizeof(float) * n);
    memcpy(Vo, s.V[0].d, sizeof(float) * n);
    free_pyr_f(s.J11, nl); free_pyr_f(s.J12, nl); free_pyr_f(s.J13, nl); free_pyr_f(s.J22, nl); free_pyr_f(s.J23, nl);
    free_pyr_f(s.U, nl); free_pyr_f(s.V, nl); free_pyr_f(s.Ur, nl); free_pyr_f(s.Vr, nl);
    free(Af); free(Bf); free(fx); free(fy); free(ft);
    return 1;
}
