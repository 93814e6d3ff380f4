/*
 * xw_oracle_fpv.c -- CPU restatement of the first-person view (--visible_radius > 0) of the XWorld2D render
 * path.  TEST INFRASTRUCTURE ONLY (see xw_oracle.h).
 *
 * Reference call sequence (all under /root/reference):
 *   XWorldSimulator::get_screen          games/xworld/xworld_simulator.cpp:278-285
 *   -> get_screen_rgb                    :287-307   to_image, cv::resize(view -> H*64 x W*64), HWC -> CHW
 *      -> XMap::to_image                 games/xworld/xworld/xmap.cpp:125-205
 *         -> XItem::get_item_image       games/xworld/xworld/xitem.cpp:33-63  (per item: getRotationMatrix2D + warpAffine)
 *         -> XMap::image_masking         xmap.cpp:273-362 (ROI ahead of the agent + wall shadows)
 *         -> copyMakeBorder (black), crop, black shadow cells, warpAffine by 90 + yaw (view rotation)
 *   -> down_sample_image                 xworld_simulator.cpp:508-545  CHW -> HWC, cv::resize, [cvtColor], HWC -> CHW
 *
 * Third-party arithmetic (OpenCV 3.2.0, cmake/opencv.cmake:5-6; not under /root/reference), restated from the
 * published algorithm in modules/imgproc/src/imgwarp.cpp and pinned against the real cv2 (tests/golden/gen_fpv_golden.py,
 * tests/test_oracle_fpv.py):
 *   cv::getRotationMatrix2D      angle *= CV_PI/180; alpha = cos*scale; beta = sin*scale; centre as Point2f
 *   cv::warpAffine               the matrix is inverted in double; coordinates in fixed point: AB_BITS = 10,
 *                                round_delta = 16, INTER_BITS = 5 (1/32 pixel); remap INTER_LINEAR with the 32x32 table of
 *                                2x2 short weights (INTER_REMAP_COEF_BITS = 15, sum forced to 32768), result
 *                                (sum + 2^14) >> 15; BORDER_CONSTANT taps outside the source read the border value
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "xw_oracle.h"

#define G XW_ICON_SIZE

/* ------------------------------------------------------------------------------------------ */
/* cv::warpAffine                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* initInterTab2D(INTER_LINEAR, fixpt = true) (imgwarp.cpp): tab1[i] = (1 - i/32, i/32) in float; the 2x2 weights are
 * saturate_cast<short>(v * 32768) and, where they do not add up to 32768, the difference is put on one entry.  That
 * correction scans itab[k1*ksize + k2] for k1, k2 in [ksize/2, ksize/2 + 2) -- for ksize == 2 these are entries 3..6,
 * of which 4..6 belong to the NEXT table cell, not yet written (zero: the table is a static array filled in order).
 * Only cell (0, 0) needs it: 32768 saturates to 32767 and entry 3 becomes 1, i.e. the weights (32767, 0, 0, 1). */
static int16_t g_itab[32 * 32 * 4 + 8];
__attribute__((constructor)) static void build_inter_tab(void) {
    float tab1[32][2];
    for (int i = 0; i < 32; ++i) {
        float x = (float)i * (1.f / 32);
        tab1[i][0] = 1.f - x;
        tab1[i][1] = x;
    }
    memset(g_itab, 0, sizeof g_itab);
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            int16_t* itab = g_itab + (i * 32 + j) * 4;
            int isum = 0;
            for (int k1 = 0; k1 < 2; ++k1) {
                float vy = tab1[i][k1];
                for (int k2 = 0; k2 < 2; ++k2) {
                    float v = vy * tab1[j][k2];
                    long iv = lrintf(v * 32768.f);
                    if (iv > 32767) iv = 32767;
                    if (iv < -32768) iv = -32768;
                    itab[k1 * 2 + k2] = (int16_t)iv;
                    isum += (int)iv;
                }
            }
            if (isum != 32768) {
                int diff = isum - 32768;
                int Mk = 3, mk = 3;
                for (int k1 = 1; k1 < 3; ++k1)
                    for (int k2 = 1; k2 < 3; ++k2) {
                        int idx = k1 * 2 + k2;
                        if (itab[idx] < itab[mk]) mk = idx;
                        else if (itab[idx] > itab[Mk]) Mk = idx;
                    }
                if (diff < 0) itab[Mk] = (int16_t)(itab[Mk] - diff);
                else itab[mk] = (int16_t)(itab[mk] - diff);
            }
        }
}

void xo_rotation_matrix(float cx, float cy, double angle_deg, double scale, double M[6]) {
    double angle = angle_deg * (3.1415926535897932384626433832795 / 180);
    double alpha = cos(angle) * scale;
    double beta = sin(angle) * scale;
    M[0] = alpha; M[1] = beta; M[2] = (1 - alpha) * cx - beta * cy;
    M[3] = -beta; M[4] = alpha; M[5] = beta * cx + (1 - alpha) * cy;
}

void xo_warp_affine_8uc3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, const double Min[6],
                         const uint8_t border[3]) {
    double M[6];
    memcpy(M, Min, sizeof M);
    { /* !(flags & WARP_INVERSE_MAP): invert */
        double D = M[0] * M[4] - M[1] * M[3];
        D = D != 0 ? 1. / D : 0;
        double A11 = M[4] * D, A22 = M[0] * D;
        M[0] = A11; M[1] *= -D;
        M[3] *= -D; M[4] = A22;
        double b1 = -M[0] * M[2] - M[1] * M[5];
        double b2 = -M[3] * M[2] - M[4] * M[5];
        M[2] = b1; M[5] = b2;
    }
    int* adelta = (int*)malloc(sizeof(int) * 2 * (size_t)dw);
    int* bdelta = adelta + dw;
    for (int x = 0; x < dw; ++x) { /* saturate_cast<int>(double) == cvRound == lrint */
        adelta[x] = (int)lrint(M[0] * x * 1024);
        bdelta[x] = (int)lrint(M[3] * x * 1024);
    }
    for (int y = 0; y < dh; ++y) {
        int X0 = (int)lrint((M[1] * y + M[2]) * 1024) + 16;
        int Y0 = (int)lrint((M[4] * y + M[5]) * 1024) + 16;
        for (int x = 0; x < dw; ++x) {
            int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            int sx = X >> 5, sy = Y >> 5; /* (saturate_cast<short>: sizes here are far below 32768) */
            const int16_t* w = g_itab + ((Y & 31) * 32 + (X & 31)) * 4;
            for (int c = 0; c < 3; ++c) {
                int acc = 0;
                for (int k1 = 0; k1 < 2; ++k1)
                    for (int k2 = 0; k2 < 2; ++k2) {
                        int yy = sy + k1, xx = sx + k2;
                        int p = (yy >= 0 && yy < sh && xx >= 0 && xx < sw) ? src[((size_t)yy * sw + xx) * 3 + c] : border[c];
                        acc += p * w[k1 * 2 + k2];
                    }
                dst[((size_t)y * dw + x) * 3 + c] = (uint8_t)((acc + (1 << 14)) >> 15);
            }
        }
    }
    free(adelta);
}

/* XItem::get_item_image (xitem.cpp:47-60) after the cached imread + identity resize */
void xo_item_image(const uint8_t* icon, double yaw, double scale, double offset, uint8_t* out) {
    static const uint8_t white[3] = {255, 255, 255};
    double M[6];
    xo_rotation_matrix((float)(G / 2.0), (float)(G / 2.0), 90 - yaw * 180 / M_PI, scale, M);
    M[2] += (offset + scale / 2 - 0.5) * G;
    M[5] += (offset + scale / 2 - 0.5) * G;
    xo_warp_affine_8uc3(icon, G, G, out, G, G, M, white);
}

/* XItem::get_item_facing_dir (xitem.cpp:65-78) */
int xo_facing_dir(double yaw) {
    const double eps = 1e-4;
    if (fabs(yaw) < eps) return 0;            /* "right" */
    if (fabs(yaw - M_PI / 2) < eps) return 1; /* "down"  */
    if (fabs(yaw - M_PI) < eps) return 2;     /* "left"  */
    return 3;                                 /* "up"    */
}

int xo_visible_radius(const xw_config* cfg) { /* xworld_simulator.cpp:63-64 */
    int m = cfg->height > cfg->width ? cfg->height : cfg->width;
    return cfg->visible_radius < m ? cfg->visible_radius : m;
}

/* set_property (xworld_env.py:211-223): yaw = uniform(0, PI_2 * 4), scale = uniform(0.5, 1), offset = uniform(0, 1 - scale),
 * random.uniform(a, b) = a + (b - a) * random().  The reference's random() is Python's unseeded generator; here
 * random() = (Philox draw) * 2^-32 for scale and offset, and (draw >> 20) / 4096 for the yaw: the engine evaluates
 * cos / sin of the 4096 possible yaws on the HOST (same libm as this file), because device and host libm do not agree to
 * the last bit and the matrix feeds lrint(). */
void xo_goal_pose(uint64_t seed, int64_t gid, uint32_t ep, uint32_t att, int k, double* yaw, double* scale, double* offset) {
    const double PI_2 = 1.5707963;
    double u0 = (double)(xo_draw(seed, gid, ep, att, XO_SITE_GOAL_POSE, 4u * (uint32_t)k + 0) >> 20) / (double)XO_YAW_STEPS;
    double u1 = (double)xo_draw(seed, gid, ep, att, XO_SITE_GOAL_POSE, 4u * (uint32_t)k + 1) * (1.0 / 4294967296.0);
    double u2 = (double)xo_draw(seed, gid, ep, att, XO_SITE_GOAL_POSE, 4u * (uint32_t)k + 2) * (1.0 / 4294967296.0);
    *yaw = 0 + (PI_2 * 4 - 0) * u0;
    *scale = 0.5 + (1 - 0.5) * u1;
    *offset = 0 + ((1 - *scale) - 0) * u2;
}

/* ------------------------------------------------------------------------------------------ */
/* XMap::image_masking (xmap.cpp:273-362)                                                      */
/* ------------------------------------------------------------------------------------------ */
static int is_block(const xo_env* e, int x, int y) {
    return x >= 0 && x < e->W && y >= 0 && y < e->H && e->grid[y * e->W + x] == XW_CELL_BLOCK;
}

void xo_image_masking(const xo_env* e, int vr, int* x_st_out, int* y_st_out, uint8_t* shadow) {
    int xa = e->agent_x + vr, ya = e->agent_y + vr;
    int major_inc_x = 0, major_inc_y = 0, minor_inc_x = 0, minor_inc_y = 0, scan_x = 0, scan_y = 0;
    int dir = xo_facing_dir(e->agent_yaw);
    if (dir == 0) { xa += vr / 2; major_inc_y = 1; minor_inc_x = 1; }
    else if (dir == 3) { ya -= vr / 2; major_inc_x = 1; minor_inc_y = -1; scan_y = vr - 1; }
    else if (dir == 2) { xa -= vr / 2; major_inc_y = 1; minor_inc_x = -1; scan_x = vr - 1; }
    else { ya += vr / 2; major_inc_x = 1; minor_inc_y = 1; }
    int x_st = xa - vr / 2, y_st = ya - vr / 2;
    /* which grids the agent's ray can start going forward */
    uint8_t ray_starts[XW_MAX_DIM];
    memset(ray_starts, 1, sizeof ray_starts);
    for (int o = -1; o <= 1; o += 2) {
        int block = 0, ray_x = e->agent_x, ray_y = e->agent_y;
        for (int k = 1; k <= vr / 2; ++k) {
            ray_x += o * major_inc_x;
            ray_y += o * major_inc_y;
            if (block) ray_starts[vr / 2 + o * k] = 0;
            if (is_block(e, ray_x, ray_y)) block = 1;
        }
    }
    /* shadow grids due to the occlusion of the wall blocks */
    memset(shadow, 0, (size_t)vr * vr);
    for (int k = 0; k < vr; ++k) {
        int block = !ray_starts[k];
        int cur_x = scan_x, cur_y = scan_y;
        for (int j = 0; j < vr; ++j) {
            if (block) shadow[cur_y * vr + cur_x] = 1;
            int g_x = x_st - vr + cur_x, g_y = y_st - vr + cur_y;
            if (is_block(e, g_x, g_y)) block = 1;
            cur_x = (cur_x + minor_inc_x + vr) % vr;
            cur_y = (cur_y + minor_inc_y + vr) % vr;
        }
        scan_x += major_inc_x;
        scan_y += major_inc_y;
    }
    *x_st_out = x_st;
    *y_st_out = y_st;
}

/* ------------------------------------------------------------------------------------------ */
/* The frame                                                                                   */
/* ------------------------------------------------------------------------------------------ */
void xo_gray_or_planes(const xw_config* cfg, const uint8_t* img_out, int oh, int ow, uint8_t* out); /* xw_oracle.c */

void xo_render_fpv(const xw_config* cfg, const xw_catalog* cat, const xo_env* e, uint8_t* out) {
    const int vr = xo_visible_radius(cfg);
    const int H = e->H, W = e->W;
    const int ch = H * G, cw = W * G;
    const int ph = (H + 2 * vr) * G, pw = (W + 2 * vr) * G; /* copyMakeBorder, xmap.cpp:154-161 */
    const int vs = vr * G;
    int oh, ow;
    xo_frame_dims(cfg, &oh, &ow);
    uint8_t* padded = (uint8_t*)calloc((size_t)ph * pw * 3, 1); /* Scalar(0, 0, 0) border */
    uint8_t* view = (uint8_t*)malloc((size_t)vs * vs * 3 * 2);
    uint8_t* rotated = view + (size_t)vs * vs * 3;
    uint8_t* screen = (uint8_t*)malloc((size_t)ch * cw * 3 + (size_t)oh * ow * 3);
    uint8_t* img_out = screen + (size_t)ch * cw * 3;
    uint8_t item[G * G * 3];
    /* XMap::to_image (xmap.cpp:129-146): white canvas, every item's transformed icon copied to its cell */
    for (int i = 0; i < ch; ++i) memset(padded + ((size_t)(i + vr * G) * pw + (size_t)vr * G) * 3, 255, (size_t)cw * 3);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            int code = e->grid[i * W + j];
            if (code == XW_CELL_EMPTY) continue;
            int icon;
            double yaw = 1.5707963, scale = 1.0, offset = 0.0; /* Entity defaults (xworld_env.py:41-42): blocks keep them */
            if (code == XW_CELL_BLOCK) icon = cat->brick_icon;
            else if (code == XW_CELL_AGENT) { icon = cat->agent_icon; yaw = e->agent_yaw; }
            else {
                int g = code - XW_CELL_GOAL0;
                icon = e->goal_icon[g]; yaw = e->goal_yaw[g]; scale = e->goal_scale[g]; offset = e->goal_offset[g];
            }
            xo_item_image(cat->atlas64 + (size_t)icon * G * G * 3, yaw, scale, offset, item);
            for (int r = 0; r < G; ++r)
                memcpy(padded + ((size_t)((i + vr) * G + r) * pw + (size_t)(j + vr) * G) * 3, item + (size_t)r * G * 3, (size_t)G * 3);
        }
    /* ROI + shadows (xmap.cpp:150-152), crop (:162-167), black shadow cells (:170-185; flag_illustration is false) */
    int x_st, y_st;
    uint8_t shadow[XW_MAX_DIM * XW_MAX_DIM];
    xo_image_masking(e, vr, &x_st, &y_st, shadow);
    for (int r = 0; r < vs; ++r)
        memcpy(view + (size_t)r * vs * 3, padded + ((size_t)(y_st * G + r) * pw + (size_t)x_st * G) * 3, (size_t)vs * 3);
    for (int x = 0; x < vr; ++x)
        for (int y = 0; y < vr; ++y)
            if (shadow[y * vr + x])
                for (int r = 0; r < G; ++r) memset(view + ((size_t)(y * G + r) * vs + (size_t)x * G) * 3, 0, (size_t)G * 3);
    { /* rotate to the agent's heading (xmap.cpp:196-200); warpAffine's defaults: INTER_LINEAR, BORDER_CONSTANT, Scalar() */
        static const uint8_t black[3] = {0, 0, 0};
        double M[6];
        xo_rotation_matrix((float)(vs / 2.0), (float)(vs / 2.0), 90 + e->agent_yaw * 180 / M_PI, 1.0, M);
        xo_warp_affine_8uc3(view, vs, vs, rotated, vs, vs, M, black);
    }
    /* get_screen_rgb: resize to (img_height_, img_width_) = the whole map's pixel size (xworld_simulator.cpp:293-295) */
    xo_resize_linear_8uc3(rotated, vs, vs, screen, ch, cw);
    /* down_sample_image (:508-545) */
    xo_resize_linear_8uc3(screen, ch, cw, img_out, oh, ow);
    xo_gray_or_planes(cfg, img_out, oh, ow, out);
    free(screen); free(view); free(padded);
}
