/* md_oracle_draw.c -- CPU ORACLE (test infrastructure only; never linked into or called by the product library).
 *
 * OpticalFlowVisualizer::showOpticalFlowVectors, common/src/optical_flow_visualizer.cpp:23-71 (call sites
 * ros/src/motion_detection_node.cpp:83,101): the original image with one anti-aliased arrow per flow vector.
 *
 * The drawing itself is cv::line(img, p1, p2, colour, 1, CV_AA, 0) of the un-vendored, un-pinned OpenCV the reference links
 * (CMakeLists.txt:14).  Restated here is OpenCV's published algorithm for that call -- imgproc/src/drawing.cpp: cv::line ->
 * ThickLine (thickness 1, shift 0) -> LineAA: Point2f -> Point by cvRound, end points in 16.16 fixed point, clipLine against
 * the image, then one step per pixel of the major axis touching three pixels of the minor axis, coverage from the 64-entry
 * filter table scaled by the slope / end-point correction tables, each touched pixel blended TWICE with
 * c += ((colour - c) * a + 127) >> 8.  Pinned against this image's cv2 4.13 (tests/test_oracle_vs_cv2.py, golden vectors in
 * tests/golden/golden_cv2.npz): bit-exact on random lines incl. clipped ones, 1 and 3 channels.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "md_oracle.h"

#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)

static const uint8_t kSlopeCorr[32] = {181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
                                       203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};
static const uint8_t kFilter[64] = {168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
                                    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
                                    158, 149, 140, 131, 122, 114, 105, 97,  89,  82,  75,  68,  62,  56,  50,  45,
                                    40,  36,  32,  28,  25,  22,  19,  16,  14,  12,  11,  9,   8,   7,   5,   5};

/* cv::clipLine(Size2l, Point2l&, Point2l&): Cohen-Sutherland with truncating 64-bit divisions */
static int clip_line64(int64_t w, int64_t h, int64_t *x1, int64_t *y1, int64_t *x2, int64_t *y2)
{
    const int64_t right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return 0;
    int c1 = (*x1 < 0) + (*x1 > right) * 2 + (*y1 < 0) * 4 + (*y1 > bottom) * 8;
    int c2 = (*x2 < 0) + (*x2 > right) * 2 + (*y2 < 0) * 4 + (*y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            *x1 += (int64_t)((double)(a - *y1) * (*x2 - *x1) / (*y2 - *y1));
            *y1 = a;
            c1 = (*x1 < 0) + (*x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            *x2 += (int64_t)((double)(a - *y2) * (*x2 - *x1) / (*y2 - *y1));
            *y2 = a;
            c2 = (*x2 < 0) + (*x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                *y1 += (int64_t)((double)(a - *x1) * (*y2 - *y1) / (*x2 - *x1));
                *x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                *y2 += (int64_t)((double)(a - *x2) * (*y2 - *y1) / (*x2 - *x1));
                *x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

static void put_point(uint8_t *img, int w, int h, int pitch, int nch, int64_t x, int64_t y, const uint8_t *colour, int a)
{
    if (x < 0 || x >= w || y < 0 || y >= h) return;
    uint8_t *p = img + (size_t)y * pitch + (size_t)x * nch;
    for (int c = 0; c < nch; c++) {
        int v = p[c];
        v += ((colour[c] - v) * a + 127) >> 8;
        v += ((colour[c] - v) * a + 127) >> 8;
        p[c] = (uint8_t)v;
    }
}

/* LineAA for integer end points (the Point arguments of cv::line with shift 0) */
void orc_line_aa(uint8_t *img, int w, int h, int pitch, int nch, int px1, int py1, int px2, int py2, const uint8_t *colour)
{
    int64_t x1 = (int64_t)px1 << XY_SHIFT, y1 = (int64_t)py1 << XY_SHIFT, x2 = (int64_t)px2 << XY_SHIFT, y2 = (int64_t)py2 << XY_SHIFT;
    if (!clip_line64((int64_t)w << XY_SHIFT, (int64_t)h << XY_SHIFT, &x1, &y1, &x2, &y2)) return;
    int64_t dx = x2 - x1, dy = y2 - y1;
    int64_t j = dx < 0 ? -1 : 0, ax = (dx ^ j) - j;
    int64_t i = dy < 0 ? -1 : 0, ay = (dy ^ i) - i;
    int64_t x_step, y_step;
    int ecount, scount = 0, slope;
    const int xmajor = ax > ay;
    if (xmajor) {
        dy = (dy ^ j) - j;
        if (j) { int64_t t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        x_step = XY_ONE;
        y_step = (dy * XY_ONE) / (ax | 1);                 /* (dy << XY_SHIFT) / (ax | 1), truncating */
        x2 += XY_ONE;
        ecount = (int)((x2 >> XY_SHIFT) - (x1 >> XY_SHIFT));
        j = -(x1 & (XY_ONE - 1));
        y1 += ((y_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((y_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (y_step < 0 ? 0x3f : 0);
        i = (x1 >> (XY_SHIFT - 7)) & 0x78;
        j = (x2 >> (XY_SHIFT - 7)) & 0x78;
    } else {
        dx = (dx ^ i) - i;
        if (i) { int64_t t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        x_step = (dx * XY_ONE) / (ay | 1);
        y_step = XY_ONE;
        y2 += XY_ONE;
        ecount = (int)((y2 >> XY_SHIFT) - (y1 >> XY_SHIFT));
        j = -(y1 & (XY_ONE - 1));
        x1 += ((x_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((x_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (x_step < 0 ? 0x3f : 0);
        i = (y1 >> (XY_SHIFT - 7)) & 0x78;
        j = (y2 >> (XY_SHIFT - 7)) & 0x78;
    }
    slope = (slope & 0x20) ? 0x100 : kSlopeCorr[slope];
    int ep[9];
    {
        const int t0 = slope << 7, t1 = ((0x78 - (int)i) | 4) * slope, t2 = ((int)j | 4) * slope;
        ep[0] = 0;
        ep[8] = slope;
        ep[1] = ep[3] = ((((int)(j - i) & 0x78) | 4) * slope >> 8) & 0x1ff;
        ep[2] = (t1 >> 8) & 0x1ff;
        ep[4] = ((((int)(j - i) + 0x80) | 4) * slope >> 8) & 0x1ff;
        ep[5] = ((t1 + t0) >> 8) & 0x1ff;
        ep[6] = (t2 >> 8) & 0x1ff;
        ep[7] = ((t2 + t0) >> 8) & 0x1ff;
    }
    if (xmajor) {
        int64_t x = x1 >> XY_SHIFT;
        while (ecount >= 0) {
            const int64_t y = (y1 >> XY_SHIFT) - 1;
            const int epc = ep[(((scount >= 2) + 1) & (scount | 2)) * 3 + (((ecount >= 2) + 1) & (ecount | 2))];
            const int dist = (int)((y1 >> (XY_SHIFT - 5)) & 31);
            put_point(img, w, h, pitch, nch, x, y, colour, (epc * kFilter[dist + 32] >> 8) & 0xff);
            put_point(img, w, h, pitch, nch, x, y + 1, colour, (epc * kFilter[dist] >> 8) & 0xff);
            put_point(img, w, h, pitch, nch, x, y + 2, colour, (epc * kFilter[63 - dist] >> 8) & 0xff);
            y1 += y_step; x++; scount++; ecount--;
        }
    } else {
        int64_t y = y1 >> XY_SHIFT;
        while (ecount >= 0) {
            const int64_t x = (x1 >> XY_SHIFT) - 1;
            const int epc = ep[(((scount >= 2) + 1) & (scount | 2)) * 3 + (((ecount >= 2) + 1) & (ecount | 2))];
            const int dist = (int)((x1 >> (XY_SHIFT - 5)) & 31);
            put_point(img, w, h, pitch, nch, x, y, colour, (epc * kFilter[dist + 32] >> 8) & 0xff);
            put_point(img, w, h, pitch, nch, x + 1, y, colour, (epc * kFilter[dist] >> 8) & 0xff);
            put_point(img, w, h, pitch, nch, x + 2, y, colour, (epc * kFilter[63 - dist] >> 8) & 0xff);
            x1 += x_step; y++; scount++; ecount--;
        }
    }
}

/* The three integer segments of one arrow (shaft, two head strokes), exactly as optical_flow_visualizer.cpp:33-66 forms them:
 * Point2f start / end (double -> float), atan2 on float arguments (the float overload <cmath> selects), the head points in
 * double rounded to float, every Point2f -> Point by cvRound.  Returns 0 when the vector is not drawn (:37). */
int orc_arrow_segments(const double *elem, int pixel_step, double min_vector_size, int *seg /* [3][4] */)
{
    if (!((fabs(elem[2]) > min_vector_size || fabs(elem[3]) > min_vector_size) && fabs(elem[2]) < pixel_step * 5 &&
          fabs(elem[3]) < pixel_step * 5))
        return 0;
    const float sx = (float)elem[0], sy = (float)elem[1];
    const float ex = (float)((double)sx + elem[2]), ey = (float)((double)sy + elem[3]);
    const double back = (double)atan2f(sy - ey, sx - ex);
    const double a1 = back + M_PI / 4.0, a2 = back - M_PI / 4.0;
    const float h1x = (float)((double)ex + 3.0 * cos(a1)), h1y = (float)((double)ey + 3.0 * sin(a1));
    const float h2x = (float)((double)ex + 3.0 * cos(a2)), h2y = (float)((double)ey + 3.0 * sin(a2));
    const int isx = (int)lrintf(sx), isy = (int)lrintf(sy), iex = (int)lrintf(ex), iey = (int)lrintf(ey);
    seg[0] = isx; seg[1] = isy; seg[2] = iex; seg[3] = iey;
    seg[4] = iex; seg[5] = iey; seg[6] = (int)lrintf(h1x); seg[7] = (int)lrintf(h1y);
    seg[8] = iex; seg[9] = iey; seg[10] = (int)lrintf(h2x); seg[11] = (int)lrintf(h2y);
    return 1;
}

/* showOpticalFlowVectors over a list of flow-field elements in the order the reference visits them (row-major over the
 * CV_64FC4 field: y outer, x inner) */
int orc_draw_flow(const uint8_t *src, int w, int h, int nch, int pitch, const double *vec4, int n, int pixel_step,
                  double min_vector_size, const uint8_t *colour, uint8_t *dst, int dpitch)
{
    int drawn = 0;
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dpitch, src + (size_t)y * pitch, (size_t)w * nch);
    for (int k = 0; k < n; k++) {
        int seg[12];
        if (!orc_arrow_segments(vec4 + 4 * k, pixel_step, min_vector_size, seg)) continue;
        for (int l = 0; l < 3; l++) orc_line_aa(dst, w, h, dpitch, nch, seg[4 * l], seg[4 * l + 1], seg[4 * l + 2], seg[4 * l + 3], colour);
        drawn++;
    }
    return drawn;
}
