// xw_fpv.cuh -- the first-person view (--visible_radius > 0): XWorldSimulator::get_screen for every env when the agent
// sees a vr x vr window of cells ahead of it.
//
// Replaces (reference file:line):
//   XMap::to_image, visible_radius_unit > 0   games/xworld/xworld/xmap.cpp:125-200  (canvas, black border, crop, shadow cells,
//                                                                                   rotation by 90 + yaw)
//   XMap::image_masking                       xmap.cpp:273-362                      (ROI + wall shadows)
//   XItem::get_item_image                     games/xworld/xworld/xitem.cpp:33-63   (per-item warpAffine: yaw, scale, offset)
//   XWorldSimulator::get_screen_rgb           games/xworld/xworld_simulator.cpp:287-307  (cv::resize view -> map size)
//   XWorldSimulator::down_sample_image        xworld_simulator.cpp:508-545               (cv::resize -> frame)
//
// None of the reference's images exists here.  An output pixel is resize2(resize1(view)) evaluated backwards: its 2x2 taps in
// the map-sized intermediate image, each of which is 2x2 taps in the rotated view; a view pixel is a pure function of
//   (heading, crop cell code, pixel in the cell):  the view rotation is an exact pixel permutation for the four headings
//   (a quarter turn about the pixel corner (N/2, N/2): index i -> N - i, one black row / column; tests/test_oracle_fpv.py),
//   a crop cell is white / black (outside the map or in a wall's shadow) / brick / the agent (its icon turned by the same
//   kind of permutation) / goal g (its icon warped with the goal's own yaw, scale and offset).
// Goal icons are warped ONCE per episode (k_fpv_warp_goals, after the reset kernel) into a per-env cache in HBM with
// cv::warpAffine's fixed-point arithmetic; the matrix set-up runs in IEEE doubles on the device, with cos / sin taken from a
// host-evaluated table (xw_common.cuh XW_YAW_STEPS).
//
// Frame kernel (k_render_fpv): one CTA per env at a time, frame composed in shared memory, one TMA bulk store per frame.
// Pixels whose whole footprint lies inside one crop cell of a uniform class come from per-heading tables built once per
// handle (pmap: which crop cell; Tb / Ta: the finished pixel if that cell is a brick / the agent; white and black are
// constants); the rest (footprints that straddle cells, goal cells) are queued and evaluated exactly, tap by tap.
#pragma once
#include "xw_common.cuh"
#include "xw_render.cuh"
#include "xw_reset.cuh"

// crop-cell classes (the grid's cell codes + black)
#define XW_FPV_BLACK 0x80
// goal counters of one render launch (XwFpv::goal_count[2][XW_FPV_SLOTS]): [0 .. 8] entries listed per chunk slot, [9] entries
// claimed by the streaming goal kernel, [10] frame-kernel groups that have finished
#define XW_FPV_SLOTS 11
#define XW_FPV_HEAD 9
#define XW_FPV_DONE 10

struct XwFpv {
    int32_t vr, N;            // window side in cells / pixels (N = 64 * vr)
    int32_t CH;               // side of the intermediate image = map side * 64 (xworld_simulator.cpp:293)
    int32_t OH, OW, FB;       // frame: rows, columns, bytes (3 * OH * OW)
    int32_t ident1, ident2;   // resize 1 / 2 is a copy (cv::resize with equal sizes)
    int32_t G;                // goal slots per env
    // cv::resize tables (INTER_LINEAR, 11-bit weights): columns clamp index and weight, rows keep the raw weight and
    // clip the two row indices (oracle/xw_oracle.c xo_resize_linear_8uc3 has the derivation)
    const int16_t *x1ofs, *x1a0, *x1a1, *y1ofs, *y1a0, *y1a1;  // [CH]  view -> intermediate
    const int16_t *x2ofs, *x2a0, *x2a1, *y2ofs, *y2a0, *y2a1;  // [OW] / [OH]  intermediate -> frame
    const uint8_t* atlas64;   // [n_icons][64][64][3] BGR
    int32_t brick_icon, agent_icon;
    const uint8_t* agent4;    // [4][64][64][3] the agent's icon for headings right, down, left, up
    uint32_t* gcache;         // [n][G][64][64] warped goal icons of the current episode, B | G << 8 | R << 16 per pixel
    // regular geometries (every frame pixel takes all its taps from one cell, cell blocks a multiple of 4 pixels wide --
    // e.g. the 84x84 view of visible_radius 7 or 3 on a map it divides evenly): k_render_fpv_cells
    int32_t regular, bs;      // bs = pixels per cell block (OW / vr)
    const uint8_t* cell2block;// [4][vr * vr]: crop cell -> its block (by * vr + bx) in the frame, per heading (derived from pmap)
    const uint8_t* block2cell;// [4][vr * vr]: the inverse
    const uint16_t* taps;     // [4][OH][OW][32]: per heading and frame pixel the 16 in-cell pixel offsets (y * 64 + x) of its taps and
                              // the 12 weights of the two resizes (xw_fpv_tap_entry) -- the host mirror's layout (tests/hostsim)
    int32_t* goal_count;      // [2][slots] k_render_fpv_cells -> k_fpv_goal_cells: the launch's visible goal cells (ping-pong),
    uint64_t* goal_list;      // [n * G] env | (block | slot << 8 | heading << 16) << 32
    const void* taps4;        // (uint4) the same entries on the device (the 16 offsets in bytes: x 4), quarter-major: [4][4 quarters of 8 values][OH * OW], so that the lanes
                              // of a warp (consecutive pixels) read consecutive 16-byte words
    const int16_t* itab;      // [32][32][4] cv::warpAffine's bilinear weights (sum 32768)
    const uint8_t* pmap;      // [4][OH][OW] crop cell (cy * vr + cx) that holds the whole footprint of the pixel, 0xff = none
    const uint8_t* Tb;        // [4][3][OH][OW] the pixel if that cell is a brick
    const uint8_t* Ta;        // [4][3][OH][OW] the pixel if that cell is the agent
};

// ---------------------------------------------------------------------------------------- geometry
// View rotation (xmap.cpp:196-200), inverted: pixel (vy, vx) of the rotated view is pixel (cy, cx) of the crop;
// false = the black row / column the quarter turn about (N/2, N/2) leaves.
XW_HD bool xw_fpv_unrotate(int N, int facing, int vy, int vx, int* cy, int* cx) {
    int y, x;
    if (facing == 3) { y = vy; x = vx; }
    else if (facing == 0) { y = vx; x = N - vy; }
    else if (facing == 1) { y = N - vy; x = N - vx; }
    else { y = N - vx; x = vy; }
    *cy = y; *cx = x;
    return y < N && x < N;
}
// The agent's icon (xitem.cpp:47-60 with the agent's yaw): pixel (y, x) of the turned icon is pixel (sy, sx) of the file's;
// false = the white row / column.
XW_HD bool xw_fpv_agent_src(int facing, int y, int x, int* sy, int* sx) {
    int a, b;
    if (facing == 1) { a = y; b = x; }
    else if (facing == 0) { a = x; b = 64 - y; }
    else if (facing == 2) { a = 64 - x; b = y; }
    else { a = 64 - y; b = 64 - x; }
    *sy = a; *sx = b;
    return a < 64 && b < 64;
}

// XMap::image_masking (xmap.cpp:273-362) + the crop: class of every cell of the vr x vr window, ccode[cy * vr + cx].
// One call computes the major line `k` (a column of the window when the agent looks up / down, a row otherwise): the
// lines are independent, so vr lanes do them side by side.
XW_HD void xw_fpv_cells_line_at(const uint8_t* g, int W, int H, int vr, int ax, int ay, int facing, int k, uint8_t* ccode) {
    const int h = vr / 2;
    // window origin in map cells: x_st - vr, y_st - vr of the reference (the padded canvas is shifted by vr)
    int ox = ax - h, oy = ay - h;
    int major_x = 0, major_y = 0, minor_x = 0, minor_y = 0, scan_x = 0, scan_y = 0;
    if (facing == 0) { ox += h; major_y = 1; minor_x = 1; }
    else if (facing == 3) { oy -= h; major_x = 1; minor_y = -1; scan_y = vr - 1; }
    else if (facing == 2) { ox -= h; major_y = 1; minor_x = -1; scan_x = vr - 1; }
    else { oy += h; major_x = 1; minor_y = 1; }
    // "which grids the agent's ray can start going forward": line k starts blocked iff a brick sits between the agent and it
    bool block = false;
    {
        const int o = k < h ? -1 : 1, steps = k < h ? h - k : k - h;
        int rx = ax, ry = ay;
        for (int j = 1; j < steps; ++j) {
            rx += o * major_x; ry += o * major_y;
            if (rx >= 0 && rx < W && ry >= 0 && ry < H && g[ry * W + rx] == XW_CELL_BLOCK) block = true;
        }
    }
    int cx = scan_x + k * major_x, cy = scan_y + k * major_y;
    for (int j = 0; j < vr; ++j) {
        const int gx = ox + cx, gy = oy + cy;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const int code = in ? g[gy * W + gx] : XW_FPV_BLACK;          // copyMakeBorder(..., Scalar(0, 0, 0))
        ccode[cy * vr + cx] = (uint8_t)(block ? XW_FPV_BLACK : code);  // shadow cells are painted black (xmap.cpp:170-185)
        if (code == XW_CELL_BLOCK) block = true;
        cx += minor_x; cy += minor_y;  // (never wraps: vr steps from one edge of the window to the other)
    }
}
XW_HD void xw_fpv_cells_line(const XwDev& d, int e, int k, uint8_t* ccode) {
    xw_fpv_cells_line_at(d.grid + (size_t)e * d.CS, d.W, d.H, d.vr, d.agent_x[e], d.agent_y[e], d.facing[e], k, ccode);
}

// ---------------------------------------------------------------------------------------- pixels
XW_HD uint32_t xw_px3(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }

// A view pixel of one env (B | G << 8 | R << 16).
struct XwFpvEnvFetch {
    const XwFpv* F; const uint8_t* ccode; const uint32_t* gc; int facing;
    XW_HD uint32_t operator()(int vy, int vx) const {
        int cy, cx;
        if (!xw_fpv_unrotate(F->N, facing, vy, vx, &cy, &cx)) return 0u;
        const int code = ccode[(cy >> 6) * F->vr + (cx >> 6)];
        if (code == XW_CELL_EMPTY) return 0xffffffu;
        if (code == XW_FPV_BLACK) return 0u;
        const int off = (((cy & 63) << 6) + (cx & 63)) * 3;
        if (code == XW_CELL_BLOCK) return xw_px3(F->atlas64 + (size_t)F->brick_icon * 12288 + off);
        if (code == XW_CELL_AGENT) return xw_px3(F->agent4 + facing * 12288 + off);
        return gc[(size_t)(code - XW_CELL_GOAL0) * 4096 + (off / 3)];
    }
};
// Table building: every cell holds `src` (a 64x64x3 icon); records which crop cells the taps touch.
struct XwFpvProbeFetch {
    const XwFpv* F; const uint8_t* src; int facing;
    mutable int cell, mixed;
    XW_HD uint32_t operator()(int vy, int vx) const {
        int cy, cx;
        if (!xw_fpv_unrotate(F->N, facing, vy, vx, &cy, &cx)) { mixed = 1; return 0u; }
        const int c = (cy >> 6) * F->vr + (cx >> 6);
        if (cell < 0) cell = c; else if (cell != c) mixed = 1;
        return xw_px3(src + (((cy & 63) << 6) + (cx & 63)) * 3);
    }
};

XW_HD int xw_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
XW_HD uint32_t xw_resize_px3(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, int a0, int a1, int b0, int b1) {
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s = 8 * c;
        o |= (uint32_t)xw_resize_px((p00 >> s) & 255, (p01 >> s) & 255, (p10 >> s) & 255, (p11 >> s) & 255, a0, a1, b0, b1) << s;
    }
    return o;
}
// pixel (jy, jx) of the intermediate image: cv::resize(view -> CH x CH)
template <class Fetch>
XW_HD uint32_t xw_fpv_mid(const XwFpv& F, int jy, int jx, const Fetch& fetch) {
    if (F.ident1) return fetch(jy, jx);
    const int sx0 = F.x1ofs[jx], sx1 = sx0 + 1 < F.N ? sx0 + 1 : F.N - 1;
    const int sy = F.y1ofs[jy], sy0 = xw_clampi(sy, 0, F.N - 1), sy1 = xw_clampi(sy + 1, 0, F.N - 1);
    return xw_resize_px3(fetch(sy0, sx0), fetch(sy0, sx1), fetch(sy1, sx0), fetch(sy1, sx1), F.x1a0[jx], F.x1a1[jx], F.y1a0[jy], F.y1a1[jy]);
}
// frame pixel (oy, ox), three channels: cv::resize(intermediate -> OH x OW)
template <class Fetch>
XW_HD uint32_t xw_fpv_px(const XwFpv& F, int oy, int ox, const Fetch& fetch) {
    if (F.ident2) return xw_fpv_mid(F, oy, ox, fetch);
    const int jx0 = F.x2ofs[ox], jx1 = jx0 + 1 < F.CH ? jx0 + 1 : F.CH - 1;
    const int jy = F.y2ofs[oy], jy0 = xw_clampi(jy, 0, F.CH - 1), jy1 = xw_clampi(jy + 1, 0, F.CH - 1);
    return xw_resize_px3(xw_fpv_mid(F, jy0, jx0, fetch), xw_fpv_mid(F, jy0, jx1, fetch), xw_fpv_mid(F, jy1, jx0, fetch),
                         xw_fpv_mid(F, jy1, jx1, fetch), F.x2a0[ox], F.x2a1[ox], F.y2a0[oy], F.y2a1[oy]);
}

// One entry of the per-heading tables: i = ((f * OH) + oy) * OW + ox.
XW_HD void xw_fpv_table_entry(const XwFpv& F, size_t i, uint8_t* pmap, uint8_t* Tb, uint8_t* Ta) {
    const int ox = (int)(i % F.OW), oy = (int)((i / F.OW) % F.OH), f = (int)(i / ((size_t)F.OW * F.OH));
    const size_t plane = (size_t)F.OH * F.OW;
    XwFpvProbeFetch pb;
    pb.F = &F; pb.facing = f; pb.cell = -1; pb.mixed = 0;
    pb.src = F.atlas64 + (size_t)F.brick_icon * 12288;
    const uint32_t vb = xw_fpv_px(F, oy, ox, pb);
    pmap[i] = (uint8_t)(pb.mixed || pb.cell < 0 ? 0xff : pb.cell);
    pb.src = F.agent4 + f * 12288;
    const uint32_t va = xw_fpv_px(F, oy, ox, pb);
    for (int c = 0; c < 3; ++c) {
        Tb[((size_t)f * 3 + c) * plane + (size_t)oy * F.OW + ox] = (uint8_t)(vb >> (8 * c));
        Ta[((size_t)f * 3 + c) * plane + (size_t)oy * F.OW + ox] = (uint8_t)(va >> (8 * c));
    }
}

// Regular geometries: the taps of frame pixel (oy, ox) under heading f, as offsets inside the one cell that holds them all,
// and the weights of the two resizes.  e[0..15]: tap (s2 * 4 + s1) -- s2 = (row, column) of the pixel's 2x2 taps in the
// intermediate image, s1 = (row, column) of that pixel's 2x2 taps in the view -- as y * 64 + x inside the cell;
// e[16..19]: horizontal weights (a0, a1) of resize 1 for the left / right intermediate column, e[20..23]: vertical (b0, b1)
// for the upper / lower intermediate row; e[24..27]: a0, a1, b0, b1 of resize 2.  A copy (equal sizes) is weights (2048, 0).
struct XwFpvTapFetch {
    const XwFpv* F; int facing; mutable uint16_t* out; mutable int n;
    XW_HD uint32_t operator()(int vy, int vx) const {
        int cy, cx;
        xw_fpv_unrotate(F->N, facing, vy, vx, &cy, &cx);
        if (n < 16) out[n] = (uint16_t)(((cy & 63) << 6) | (cx & 63));
        ++n;
        return 0u;
    }
};
XW_HD void xw_fpv_tap_values(const XwFpv& F, size_t i, uint16_t* e) {   // e[32]: entry i = (f * OH + oy) * OW + ox
    const int ox = (int)(i % F.OW), oy = (int)((i / F.OW) % F.OH), f = (int)(i / ((size_t)F.OW * F.OH));
    for (int k = 0; k < 32; ++k) e[k] = 0;
    XwFpvTapFetch tf;
    tf.F = &F; tf.facing = f; tf.out = e; tf.n = 0;
    int jx[2] = {ox, ox}, jy[2] = {oy, oy};
    if (!F.ident2) {
        jx[0] = F.x2ofs[ox]; jx[1] = jx[0] + 1 < F.CH ? jx[0] + 1 : F.CH - 1;
        jy[0] = xw_clampi(F.y2ofs[oy], 0, F.CH - 1); jy[1] = xw_clampi(F.y2ofs[oy] + 1, 0, F.CH - 1);
        e[24] = (uint16_t)F.x2a0[ox]; e[25] = (uint16_t)F.x2a1[ox]; e[26] = (uint16_t)F.y2a0[oy]; e[27] = (uint16_t)F.y2a1[oy];
    } else { e[24] = 2048; e[25] = 0; e[26] = 2048; e[27] = 0; }
    for (int s2 = 0; s2 < 4; ++s2) {
        const int my = jy[s2 >> 1], mx = jx[s2 & 1];
        if (F.ident1) { for (int s1 = 0; s1 < 4; ++s1) tf(my, mx); }
        else {
            const int sx0 = F.x1ofs[mx], sx1 = sx0 + 1 < F.N ? sx0 + 1 : F.N - 1;
            const int sy = F.y1ofs[my], sy0 = xw_clampi(sy, 0, F.N - 1), sy1 = xw_clampi(sy + 1, 0, F.N - 1);
            tf(sy0, sx0); tf(sy0, sx1); tf(sy1, sx0); tf(sy1, sx1);
        }
    }
    for (int q = 0; q < 2; ++q) {
        if (F.ident1) { e[16 + 2 * q] = 2048; e[17 + 2 * q] = 0; e[20 + 2 * q] = 2048; e[21 + 2 * q] = 0; }
        else {
            e[16 + 2 * q] = (uint16_t)F.x1a0[jx[q]]; e[17 + 2 * q] = (uint16_t)F.x1a1[jx[q]];
            e[20 + 2 * q] = (uint16_t)F.y1a0[jy[q]]; e[21 + 2 * q] = (uint16_t)F.y1a1[jy[q]];
        }
    }
}
XW_HD void xw_fpv_tap_entry(const XwFpv& F, size_t i, uint16_t* taps) { xw_fpv_tap_values(F, i, taps + i * 32); }
// frame pixel from its tap entry and the (64 x 64, 4 bytes per pixel) icon of its cell
XW_HD uint32_t xw_fpv_px_taps(const uint16_t* e, const uint32_t* icon) {
    uint32_t mid[4];
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2)
        mid[s2] = xw_resize_px3(icon[e[4 * s2]], icon[e[4 * s2 + 1]], icon[e[4 * s2 + 2]], icon[e[4 * s2 + 3]],
                                (int16_t)e[16 + 2 * (s2 & 1)], (int16_t)e[17 + 2 * (s2 & 1)], (int16_t)e[20 + 2 * (s2 >> 1)], (int16_t)e[21 + 2 * (s2 >> 1)]);
    return xw_resize_px3(mid[0], mid[1], mid[2], mid[3], (int16_t)e[24], (int16_t)e[25], (int16_t)e[26], (int16_t)e[27]);
}

// ---------------------------------------------------------------------------------------- goal icons (cv::warpAffine)
// Fixed-point set-up of cv::warpAffine for XItem::get_item_image's matrix (xitem.cpp:47-60; OpenCV imgproc/imgwarp.cpp):
// getRotationMatrix2D(centre (32, 32), 90 - yaw * 180 / pi, scale), the offset shift, the inversion, then for index i
//   adelta = rint(M0 * i * 1024), bdelta = rint(M3 * i * 1024), X0 = rint((M1 * i + M2) * 1024) + 16, Y0 = rint((M4 * i + M5) * 1024) + 16.
// Every product / sum is a separately rounded IEEE double, as in the generic C++ OpenCV is built from.
XW_HD int xw_d2i(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)__builtin_lrint(v);
#endif
}
XW_HD void xw_fpv_warp_coeffs(const double* yaw_cs, int yaw_idx, double scale, double offset, int i, int* ad, int* bd, int* X0, int* Y0) {
    const double alpha = xw_dmul(yaw_cs[2 * yaw_idx], scale), beta = xw_dmul(yaw_cs[2 * yaw_idx + 1], scale);
    const double c = 32.0;
    const double t = xw_dmul(xw_dadd(xw_dadd(offset, xw_dmul(scale, 0.5)), -0.5), 64.0);  // (offset + scale / 2 - 0.5) * icon.cols
    double m0 = alpha, m1 = beta, m2 = xw_dadd(xw_dadd(xw_dmul(xw_dadd(1.0, -alpha), c), -xw_dmul(beta, c)), t);
    double m3 = -beta, m4 = alpha, m5 = xw_dadd(xw_dadd(xw_dmul(beta, c), xw_dmul(xw_dadd(1.0, -alpha), c)), t);
    double D = xw_dadd(xw_dmul(m0, m4), -xw_dmul(m1, m3));
#if defined(__CUDA_ARCH__)
    D = D != 0 ? __ddiv_rn(1.0, D) : 0;
#else
    D = D != 0 ? 1. / D : 0;
#endif
    const double A11 = xw_dmul(m4, D), A22 = xw_dmul(m0, D);
    m0 = A11; m1 = xw_dmul(m1, -D); m3 = xw_dmul(m3, -D); m4 = A22;
    const double b1 = xw_dadd(xw_dmul(-m0, m2), -xw_dmul(m1, m5)), b2 = xw_dadd(xw_dmul(-m3, m2), -xw_dmul(m4, m5));
    m2 = b1; m5 = b2;
    const double di = (double)i;
    *ad = xw_d2i(xw_dmul(xw_dmul(m0, di), 1024.0));
    *bd = xw_d2i(xw_dmul(xw_dmul(m3, di), 1024.0));
    *X0 = xw_d2i(xw_dmul(xw_dadd(xw_dmul(m1, di), m2), 1024.0)) + 16;
    *Y0 = xw_d2i(xw_dmul(xw_dadd(xw_dmul(m4, di), m5), 1024.0)) + 16;
}
// remap, INTER_LINEAR, BORDER_CONSTANT white: destination pixel with fixed-point source coordinates X, Y (5 fractional bits)
XW_HD uint32_t xw_fpv_warp_px(const uint8_t* icon, const int16_t* itab, int X, int Y) {
    const int sx = X >> 5, sy = Y >> 5;
    const int16_t* w = itab + (((Y & 31) << 5) + (X & 31)) * 4;
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int acc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = sy + (k >> 1), xx = sx + (k & 1);
            const int p = (yy >= 0 && yy < 64 && xx >= 0 && xx < 64) ? icon[((yy << 6) + xx) * 3 + c] : 255;
            acc += p * w[k];
        }
        o |= (uint32_t)(uint8_t)((acc + (1 << 14)) >> 15) << (8 * c);
    }
    return o;
}

// initInterTab2D(INTER_LINEAR, fixpt) (imgwarp.cpp): float weights * 32768 rounded to short; the only cell whose weights do
// not add up to 32768 is (0, 0) (32768 saturates to 32767), and the correction lands on its last entry: (32767, 0, 0, 1)
// (oracle/xw_oracle_fpv.c build_inter_tab emulates the original loop, including where it reads; the tests compare the two)
static inline void xw_fpv_build_itab(int16_t* itab) {
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            const float y1 = (float)i * (1.f / 32), x1 = (float)j * (1.f / 32);
            const float wy[2] = {1.f - y1, y1}, wx[2] = {1.f - x1, x1};
            int sum = 0;
            for (int k = 0; k < 4; ++k) {
                long v = __builtin_lrintf(wy[k >> 1] * wx[k & 1] * 32768.f);
                if (v > 32767) v = 32767;
                itab[(i * 32 + j) * 4 + k] = (int16_t)v;
                sum += (int)v;
            }
            if (sum != 32768) itab[(i * 32 + j) * 4 + 3] = (int16_t)(itab[(i * 32 + j) * 4 + 3] - (sum - 32768));
        }
}

#if defined(__CUDACC__)
__global__ void k_fpv_build_tables(XwFpv F, uint8_t* pmap, uint8_t* Tb, uint8_t* Ta) {
    const size_t total = (size_t)4 * F.OH * F.OW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        xw_fpv_table_entry(F, i, pmap, Tb, Ta);
}

// Warps the goal icons of the envs that were just reset (list mode: the step's auto-reset queue; else every env,
// optionally filtered by mask).  One CTA per env at a time.
__global__ void __launch_bounds__(256) k_fpv_warp_goals(XwDev d, XwFpv F, const uint8_t* mask, const int32_t* list, const int32_t* count) {
    __shared__ int co[4][64];
    const int cnt = list ? *count : d.n;
    for (int i = blockIdx.x; i < cnt * F.G; i += gridDim.x) {   // one CTA per (env, goal)
        const int ei = i / F.G, g = i - ei * F.G;
        const int e = list ? list[ei] : ei;
        if (!list && mask && !mask[e]) continue;
        const size_t k = (size_t)g * d.n + e;
        __syncthreads();
        if (threadIdx.x < 64)
            xw_fpv_warp_coeffs(d.yaw_cs, d.goal_yaw[k], d.goal_scale[k], d.goal_offset[k], threadIdx.x, &co[0][threadIdx.x],
                               &co[1][threadIdx.x], &co[2][threadIdx.x], &co[3][threadIdx.x]);
        __syncthreads();
        const uint8_t* icon = F.atlas64 + (size_t)d.goal_icon[k] * 12288;
        uint32_t* dst = F.gcache + ((size_t)e * F.G + g) * 4096;
        for (int p = threadIdx.x; p < 4096; p += blockDim.x) {
            const int y = p >> 6, x = p & 63;
            dst[p] = xw_fpv_warp_px(icon, F.itab, (co[2][y] + co[0][x]) >> 5, (co[3][y] + co[1][x]) >> 5);
        }
    }
}

// Any frame size: one thread per frame pixel, straight to global memory.
__global__ void __launch_bounds__(256) k_render_fpv_generic(XwDev d, XwFpv F, uint8_t* __restrict__ frames, size_t env_stride) {
    __shared__ uint8_t ccode[XW_MAX_DIM * XW_MAX_DIM];
    const int plane = F.OH * F.OW;
    for (int e = blockIdx.x; e < d.n; e += gridDim.x) {
        __syncthreads();
        if ((int)threadIdx.x < F.vr) xw_fpv_cells_line(d, e, threadIdx.x, ccode);
        __syncthreads();
        XwFpvEnvFetch fe;
        fe.F = &F; fe.ccode = ccode; fe.gc = F.gcache + (size_t)e * F.G * 4096; fe.facing = d.facing[e];
        uint8_t* out = frames + (size_t)e * env_stride;
        for (int p = threadIdx.x; p < plane; p += blockDim.x) {
            const uint32_t v = xw_fpv_px(F, p / F.OW, p % F.OW, fe);
            out[p] = (uint8_t)v; out[plane + p] = (uint8_t)(v >> 8); out[2 * plane + p] = (uint8_t)(v >> 16);
        }
    }
}

// Frame kernel for OW % 4 == 0, FB % 16 == 0.  Dynamic shared memory: frame buffer [FB] | slow list u16[OH * OW] | ccode[256].
template <int NT>
__global__ void __launch_bounds__(NT) k_render_fpv(XwDev d, XwFpv F, uint8_t* __restrict__ frames, size_t env_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* fb = (uint32_t*)smem;
    uint16_t* slow = (uint16_t*)(smem + F.FB);
    uint8_t* ccode = smem + F.FB + 2 * F.OH * F.OW;
    __shared__ int n_slow;
    const int WR = F.OW >> 2, plane_w = F.OH * WR, plane = F.OH * F.OW;
    const int tid = threadIdx.x;
    for (int e = blockIdx.x; e < d.n; e += gridDim.x) {
        if (tid == 0) { tma_wait_read<0>(); n_slow = 0; }   // the previous frame has left the buffer
        if (tid < F.vr) xw_fpv_cells_line(d, e, tid, ccode);
        __syncthreads();
        const int facing = d.facing[e];
        const uint32_t* pm = (const uint32_t*)(F.pmap + (size_t)facing * plane);
        const uint32_t* tb = (const uint32_t*)(F.Tb + (size_t)facing * 3 * plane);
        const uint32_t* ta = (const uint32_t*)(F.Ta + (size_t)facing * 3 * plane);
        // pass 1: the words of uniform cells; everything else goes on the list (one entry per pixel)
        for (int w = tid; w < plane_w; w += NT) {
            const uint32_t cells = pm[w];
            uint32_t sel_w = 0, sel_b = 0, sel_a = 0, todo = 0;  // byte masks: white, brick, agent; pixels for pass 2
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t id = (cells >> (8 * b)) & 255u;
                const int code = id == 255u ? -1 : ccode[id];
                const uint32_t m = 0xffu << (8 * b);
                if (code == XW_CELL_EMPTY) sel_w |= m;
                else if (code == XW_CELL_BLOCK) sel_b |= m;
                else if (code == XW_CELL_AGENT) sel_a |= m;
                else if (code != XW_FPV_BLACK) todo |= 1u << b;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t v = sel_w;
                if (sel_b) v |= tb[c * plane_w + w] & sel_b;
                if (sel_a) v |= ta[c * plane_w + w] & sel_a;
                fb[c * plane_w + w] = v;
            }
            if (todo) {
                const int k = atomicAdd(&n_slow, __popc(todo));
                int j = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (todo & (1u << b)) slow[k + j++] = (uint16_t)(w * 4 + b);
            }
        }
        __syncthreads();
        // pass 2: exact evaluation, tap by tap
        {
            XwFpvEnvFetch fe;
            fe.F = &F; fe.ccode = ccode; fe.gc = F.gcache + (size_t)e * F.G * 4096; fe.facing = facing;
            uint8_t* fb8 = (uint8_t*)fb;
            const int ns = n_slow;
            for (int i = tid; i < ns; i += NT) {
                const int p = slow[i];
                const uint32_t v = xw_fpv_px(F, p / F.OW, p % F.OW, fe);
                fb8[p] = (uint8_t)v; fb8[plane + p] = (uint8_t)(v >> 8); fb8[2 * plane + p] = (uint8_t)(v >> 16);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) { tma_store_1d(frames + (size_t)e * env_stride, fb, (uint32_t)F.FB); tma_commit(); }
    }
    if (tid == 0) tma_wait_all<0>();
}

__global__ void k_fpv_build_taps(XwFpv F, uint4* taps4) {
    const size_t plane = (size_t)F.OH * F.OW, total = 4 * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        __align__(16) uint16_t e[32];
        xw_fpv_tap_values(F, i, e);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = (uint16_t)(e[k] * 4);   // byte offsets into the staged icon (k_fpv_goal_cells)
        const size_t f = i / plane, p = i - f * plane;
#pragma unroll
        for (int q = 0; q < 4; ++q) taps4[(f * 4 + q) * plane + p] = ((const uint4*)e)[q];
    }
}

// Frame kernels for regular geometries (XwFpv::regular): a frame is vr x vr blocks of bs x bs pixels, each a pure function of
// its crop cell and the heading.
//
// k_render_fpv_cells -- everything but the goal cells.  One persistent CTA per SM, NG groups of NT threads, one env at a time per group:
//   * warp 0 keeps the NEXT env's pose and grid row in registers (loaded one env time ahead), and between the frame's TMA store
//     and the next paint turns them into the window's cell codes (image_masking: one lane per major line) and appends the
//     visible goal cells to the launch's goal list in HBM, while thread 32 waits for the store to have read the buffer --
//     nothing on the per-env path waits for HBM;
//   * paint: work item = (band of cells, colour plane); lane = word column, so a lane's cell is fixed per item: white / black
//     (outside the map, wall shadow) are constant words, brick / agent words go from the per-heading tables (L1 / L2) to the
//     frame buffer as 4-byte cp.async copies (SASS LDGSTS) -- all of a warp's items in flight at once, one wait per env;
//     goal blocks are left as they are (k_fpv_goal_cells overwrites them);
//   * fence.proxy.async, barrier, ONE TMA bulk store per frame (evict-first).
// k_fpv_goal_cells -- the goal cells of the launch, one CTA per (env, cell) of the list: the goal's warped icon (16 KB of the
//   per-env cache) comes into shared memory with one TMA bulk load, each thread evaluates pixels through the precomputed tap
//   entries (16 taps, five fixed-point bilinear steps) and writes them into the finished frame.  Kept out of the frame kernel
//   because it needs 16 KB of shared memory and ~60 registers that would halve the frame kernel's occupancy, and because taken
//   straight from global memory the scattered taps made the L1 data pipe the frame kernel's bound (profiles/r02_summary.md).
// (48 registers: a CTA of 1,024 threads then leaves 16 K registers of its SM to the CTAs of k_fpv_warp_goals, which runs on the
// reset stream beside this kernel.)
// list != NULL: the envs list[0 .. *count) instead of envs [env0, env0 + env_n) (the re-paint of a step's auto-reset queue).
// A launch covers one chunk of the batch: its goal cells go to goal_list + list_base, counted in goal_count[chunk], and
// k_fpv_goal_cells of chunk k runs on a second stream beside the frame kernel of chunk k + 1 (the frame kernel is bound by
// HBM writes, the goal kernel by instruction issue).
// Dynamic shared memory of k_render_fpv_cells, per group: frame [FB] | grid row [CS] | cell codes [ceil16(vr^2)] | misc [4 x i32].
// (BS_T, VR_T: compile-time block size and window side of the common geometries; 0 = read them from F)
// bar.sync with an IMMEDIATE barrier id per group: with the id in a register ptxas books all 16 barriers for the CTA, and the
// barrier slots are an SM resource -- no CTA of another kernel (which needs one for __syncthreads) could then share the SM
__device__ __forceinline__ void xw_group_bar_imm(int grp, int nthreads) {
    switch (grp) {
        case 0: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
        case 3: asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory"); break;
        case 4: asm volatile("bar.sync 5, %0;" ::"r"(nthreads) : "memory"); break;
        case 5: asm volatile("bar.sync 6, %0;" ::"r"(nthreads) : "memory"); break;
        case 6: asm volatile("bar.sync 7, %0;" ::"r"(nthreads) : "memory"); break;
        default: asm volatile("bar.sync 8, %0;" ::"r"(nthreads) : "memory"); break;
    }
}
template <int NT, int NG, int BS_T, int VR_T>
__global__ void __maxnreg__(48) k_render_fpv_cells(XwDev d, XwFpv F, uint8_t* __restrict__ frames, size_t env_stride,
                                                                 const int32_t* __restrict__ list, const int32_t* __restrict__ count,
                                                                 int env0, int env_n, int chunk, int list_base, int zero_base, int zero_n,
                                                                 int stream_seq) {
    extern __shared__ __align__(128) uint8_t smem_all[];
    // (programmatic dependent launch: the step's auto-reset kernel, queued behind this one, may start once every CTA is resident --
    //  see k_render_sp)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // the goal counters of the PREVIOUS launch (the other half of the ping-pong; its goal kernel is done: stream order) -> 0,
    // so that no memset node sits between the step kernel and this one
    if (blockIdx.x == 0 && (int)threadIdx.x < zero_n) F.goal_count[zero_base + threadIdx.x] = 0;
    const int vr = VR_T ? VR_T : F.vr, bs = BS_T ? BS_T : F.bs, bw = bs >> 2;
    const int OW = bs * vr, WR = OW >> 2, plane_w = OW * WR, plane = OW * OW;  // (regular frames are square)
    const int n_cells = vr * vr, nc16 = (n_cells + 15) & ~15;
    // NG groups of NT threads, each a frame pipeline of its own (buffer, barrier, TMA stores): one CTA per SM, so that a launch on
    // fewer CTAs than SMs leaves whole SMs to the auto-reset kernels that run beside it (xw_engine.cu step_xworld)
    const int grp = threadIdx.x / NT, tid = threadIdx.x - grp * NT;
    uint8_t* smem = smem_all + (size_t)grp * ((F.FB + d.CS + nc16 + 16 + 128 + 127) & ~127);
    const int vcta = blockIdx.x * NG + grp, vgrid = gridDim.x * NG;
    uint32_t* fb = (uint32_t*)smem;
    uint8_t* row = smem + F.FB;
    uint8_t* ccode = row + d.CS;
    int* misc = (int*)(ccode + nc16);   // [1] heading, [2] goal cells of the env being prepared
    // streaming (stream_seq != 0: k_fpv_goal_stream runs BESIDE this kernel): the list entries of an env become valid for the goal
    // kernel only when the env's frame has landed in HBM, i.e. when its TMA store is complete -- thread 32 marks them one env later
    uint32_t* pend = (uint32_t*)(misc + 4);   // [4][6] list positions of the goal cells of the last four envs, [24 .. 27] their counts
    const uint8_t* __restrict__ b2c = F.block2cell;
    const uint8_t* __restrict__ c2b = F.cell2block;
    const int lane = tid & 31, warp = tid >> 5, n_warps = NT >> 5;
    const int total = list ? *count : env_n;
    int32_t* goal_count = F.goal_count + chunk;
    uint64_t* goal_list = F.goal_list + list_base;
    // warp 0: pose + grid row of the env after this one, in registers
    const int row_words = d.CS >> 2;
    uint32_t pre0 = 0, pre1 = 0;
    int pre_ax = 0, pre_ay = 0, pre_f = 0;
    auto prefetch = [&](int idx) {
        if (idx >= total) return;
        const int e = list ? list[idx] : env0 + idx;
        const uint32_t* g = (const uint32_t*)(d.grid + (size_t)e * d.CS);
        if (lane < row_words) pre0 = g[lane];
        if (lane + 32 < row_words) pre1 = g[lane + 32];
        pre_ax = d.agent_x[e]; pre_ay = d.agent_y[e]; pre_f = d.facing[e];
    };
    if (warp == 0) prefetch(vcta);
    auto mark_ready = [&](int buf) {   // (thread 32) the env's frame is in HBM: its goal cells may be taken
        const int np = (int)pend[24 + buf];
        for (int j = 0; j < np; ++j) {
            volatile uint32_t* hi = (volatile uint32_t*)(goal_list + pend[buf * 6 + j]) + 1;
            *hi = *hi | 0x80000000u;
        }
    };
    uint32_t it = 0;
    for (int idx = vcta; idx < total; idx += vgrid, ++it) {
        const int e = list ? list[idx] : env0 + idx;
        const int pb = (int)(it & 3u);
        if (warp == 0) {
            if (lane < row_words) ((uint32_t*)row)[lane] = pre0;
            if (lane + 32 < row_words) ((uint32_t*)row)[lane + 32] = pre1;
            if (lane == 0) { misc[1] = pre_f; misc[2] = 0; }
            __syncwarp();
            for (int k = lane; k < vr; k += 32) xw_fpv_cells_line_at(row, d.W, d.H, vr, pre_ax, pre_ay, pre_f, k, ccode);
            __syncwarp();
            for (int c = lane; c < n_cells; c += 32) {
                const int code = ccode[c];
                if (code >= XW_CELL_GOAL0 && code != XW_FPV_BLACK) {
                    const int k = atomicAdd(goal_count, 1);
                    goal_list[k] = (uint64_t)(uint32_t)e | ((uint64_t)((uint32_t)c2b[pre_f * n_cells + c] | ((uint32_t)(code - XW_CELL_GOAL0) << 8) | ((uint32_t)pre_f << 16) |
                                                                       ((uint32_t)stream_seq << 18)) << 32);
                    if (stream_seq) { const int j = atomicAdd(&misc[2], 1); if (j < 6) pend[pb * 6 + j] = (uint32_t)k; }
                }
            }
            __syncwarp();
            if (lane == 0) pend[24 + pb] = (uint32_t)(misc[2] < 6 ? misc[2] : 6);
            prefetch(idx + vgrid);
        } else if (tid == 32) {
            tma_wait_read<0>();   // the previous frame has left the buffer
        }
        xw_group_bar_imm(grp, NT);
        const int facing = misc[1];
        const uint32_t* tb = (const uint32_t*)(F.Tb + (size_t)facing * 3 * plane);
        const uint32_t* ta = (const uint32_t*)(F.Ta + (size_t)facing * 3 * plane);
        // ---- bands: item = (band by, plane pl); lane = word column
        for (int item = warp; item < 3 * vr; item += n_warps) {
            const int by = item / 3, pl = item - 3 * by;
            for (int col = lane; col < WR; col += 32) {
                const int bx = col / bw;
                const int code = ccode[b2c[facing * n_cells + by * vr + bx]];
                const int w0 = pl * plane_w + by * bs * WR + col;
                if (code == XW_CELL_BLOCK || code == XW_CELL_AGENT) {
                    const uint32_t* src = (code == XW_CELL_BLOCK ? tb : ta) + w0;
#pragma unroll 12
                    for (int r = 0; r < bs; ++r) cp_async_4(fb + w0 + r * WR, src + r * WR);
                } else if (code == XW_CELL_EMPTY || code == XW_FPV_BLACK) {
                    const uint32_t fill = code == XW_CELL_EMPTY ? 0xffffffffu : 0u;
#pragma unroll 12
                    for (int r = 0; r < bs; ++r) fb[w0 + r * WR] = fill;
                }
            }
        }
        cp_async_wait_all();
        fence_async_smem();
        xw_group_bar_imm(grp, NT);
        if (tid == 32) {
            tma_store_1d(frames + (size_t)e * env_stride, fb, (uint32_t)F.FB);
            tma_commit();
            if (stream_seq && it > 1) {   // the store of the env before the previous one (committed two env times ago) is complete by now
                tma_wait_all<2>();
                __threadfence();
                mark_ready((int)((it + 2u) & 3u));
            }
        }
    }
    if (tid == 32) {
        tma_wait_all<0>();
        if (stream_seq) {
            __threadfence();
            if (it > 1) mark_ready((int)((it + 2u) & 3u));   // the group's last two envs
            if (it > 0) mark_ready((int)((it + 3u) & 3u));
            __threadfence();
            atomicAdd(F.goal_count + (chunk - chunk % XW_FPV_SLOTS) + XW_FPV_DONE, 1);
        }
    }
}

// One bilinear step of cv::resize on three packed channels (B | G << 8 | R << 16): the horizontal pass of a row is one
// two-way dot product (DP2A: 16-bit weights x 8-bit taps), a01 = a0 | a1 << 16.  Same integers as xw_resize_px3.
__device__ __forceinline__ uint32_t xw_bilin3_dev(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t a01, uint32_t b0, uint32_t b1) {
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t sel = 0x40u + 0x11u * c;   // bytes (c of p?0, c of p?1)
        const uint32_t s0 = __dp2a_lo(a01, __byte_perm(p00, p01, sel), 0u);
        const uint32_t s1 = __dp2a_lo(a01, __byte_perm(p10, p11, sel), 0u);
        const uint32_t v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2u) >> 2;
        o |= (v & 255u) << (8 * c);
    }
    return o;
}
// frame pixel from its device tap entry (xw_fpv_tap_values with the 16 offsets scaled to bytes) and the staged icon
__device__ __forceinline__ uint32_t xw_fpv_px_taps_dev(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const uint8_t* icon) {
    const uint32_t ow[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};   // 16 byte offsets
    const uint32_t a1[2] = {q2.x, q2.y}, b1[2] = {q2.z, q2.w};                  // resize 1: (a0, a1) left / right, (b0, b1) upper / lower
    uint32_t mid[4];
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
        const uint32_t w01 = ow[2 * s2], w23 = ow[2 * s2 + 1], b = b1[s2 >> 1];
        mid[s2] = xw_bilin3_dev(*(const uint32_t*)(icon + (w01 & 0xffffu)), *(const uint32_t*)(icon + (w01 >> 16)),
                                *(const uint32_t*)(icon + (w23 & 0xffffu)), *(const uint32_t*)(icon + (w23 >> 16)),
                                a1[s2 & 1], b & 0xffffu, b >> 16);
    }
    return xw_bilin3_dev(mid[0], mid[1], mid[2], mid[3], q3.x, q3.y & 0xffffu, q3.y >> 16);
}

// The goal cells k_render_fpv_cells listed for chunk `chunk`.  Dynamic shared memory: icon [16 KB] | mbarrier.
template <int BS_T, int VR_T>
__global__ void __launch_bounds__(160) k_fpv_goal_cells(XwFpv F, uint8_t* __restrict__ frames, size_t env_stride, int chunk, int list_base) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = (uint64_t*)(smem + 16384);
    const int vr = VR_T ? VR_T : F.vr, bs = BS_T ? BS_T : F.bs, OW = bs * vr, plane = OW * OW, bb = bs * bs;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int total = F.goal_count[chunk];
    const uint64_t* __restrict__ glist = F.goal_list + list_base;
    uint32_t phase = 0;
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        const uint64_t ent = glist[i];
        const uint32_t lo = (uint32_t)ent, hi = (uint32_t)(ent >> 32);
        const int e = (int)lo, blk = hi & 255, slot = (hi >> 8) & 255, facing = (hi >> 16) & 3;
        if (tid == 0) {
            mbar_expect_tx(bar, 16384u);
            tma_load_1d(smem, F.gcache + ((size_t)e * F.G + slot) * 4096, 16384u, bar);
        }
        const int by = blk / vr, bx = blk - by * vr;
        const uint4* t4 = (const uint4*)F.taps4 + (size_t)facing * 4 * plane;
        uint8_t* out = frames + (size_t)e * env_stride;
        bool waited = false;
        for (int q = tid; q < bb; q += nt) {
            const int qy = q / bs, p = (by * bs + qy) * OW + bx * bs + (q - qy * bs);
            const uint4 q0 = __ldg(t4 + p), q1 = __ldg(t4 + (size_t)plane + p), q2 = __ldg(t4 + 2 * (size_t)plane + p), q3 = __ldg(t4 + 3 * (size_t)plane + p);
            if (!waited) { mbar_wait(bar, phase); waited = true; }
            const uint32_t v = xw_fpv_px_taps_dev(q0, q1, q2, q3, smem);
            out[p] = (uint8_t)v; out[plane + p] = (uint8_t)(v >> 8); out[2 * (size_t)plane + p] = (uint8_t)(v >> 16);
        }
        if (!waited) mbar_wait(bar, phase);
        phase ^= 1;
        __syncthreads();   // every thread is done with the icon before the next load lands
    }
}

// The goal cells of a launch, taken WHILE k_render_fpv_cells runs (one CTA per SM beside the frame kernel's): warp 0 claims list
// positions (atomic head), waits until the position holds an entry of this launch whose frame has landed (ready bit + launch
// sequence number, set by the frame kernel when the env's TMA store is complete) and streams the goal's icon into a three-stage
// shared-memory ring with TMA bulk loads; each stage has its own five warps that evaluate the cell's 144 pixels (full / empty mbarriers).  The
// frame kernel is bound by HBM writes and this one by instruction issue and icon reads, so side by side they cost little more
// than the frame kernel alone.  Ends when every group of the frame kernel has reported done and the list is exhausted.
// Dynamic shared memory: 3 x icon [16 KB] | entries [3 x u64] | full [3], empty [3] mbarriers.
template <int BS_T, int VR_T>
__global__ void __maxnreg__(32) k_fpv_goal_stream(XwFpv F, uint8_t* __restrict__ frames, size_t env_stride, int ctr_base, int list_base,
                                                  int seq, int expected_groups, int list_cap) {
    // (512 threads x 32 registers: exactly what a 1,024-thread x 48-register frame CTA leaves of an SM's register file)
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int S = 3;   // stages = compute groups: warp 0 loads, threads 32 + 160 g .. 32 + 160 g + 159 evaluate the cells of stage g
    uint64_t* ent_s = (uint64_t*)(smem + S * 16384);
    uint64_t* full = ent_s + S;
    uint64_t* empty = full + S;
    const int vr = VR_T ? VR_T : F.vr, bs = BS_T ? BS_T : F.bs, OW = bs * vr, plane = OW * OW, bb = bs * bs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 5); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    volatile int32_t* ctr = F.goal_count + ctr_base;
    uint64_t* __restrict__ glist = F.goal_list + list_base;
    const uint64_t EXIT = ~0ull;
    if (warp == 0) {
        // the loader warp: 32 list positions are claimed at once and polled side by side (a claim and a poll are each an L2 round
        // trip: one lane doing them one after the other was the whole kernel's pace), then handed to the stages in order
        const long long t_start = clock64();
        uint32_t it = 0;
        for (bool more = true; more;) {
            int base = 0;
            if (lane == 0) base = atomicAdd((int32_t*)F.goal_count + ctr_base + XW_FPV_HEAD, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int i = base + lane;
            uint32_t e_lo = 0xffffffffu, e_hi = 0xffffffffu;   // EXIT
            for (;;) {
                if (i < list_cap) {
                    const uint32_t hi = ((volatile uint32_t*)(glist + i))[1];
                    if ((hi >> 31) && ((hi >> 18) & 0x1fffu) == (uint32_t)seq) {
                        __threadfence();
                        e_lo = ((volatile uint32_t*)(glist + i))[0]; e_hi = hi;
                        ((volatile uint32_t*)(glist + i))[1] = 0u;   // consumed: a later launch never finds it ready
                        break;
                    }
                }
                if (ctr[XW_FPV_DONE] >= expected_groups) {
                    __threadfence();
                    if (i >= ctr[0] || i >= list_cap) break;   // the list is exhausted
                }
                if (clock64() - t_start > 400000000ll) break;  // (0.2 s: never hang the device on a protocol error)
                __nanosleep(64);
            }
            __syncwarp();
            for (int l = 0; l < 32 && more; ++l) {
                const uint32_t lo = __shfl_sync(0xffffffffu, e_lo, l), hi = __shfl_sync(0xffffffffu, e_hi, l);
                if (lo == 0xffffffffu && hi == 0xffffffffu) { more = false; break; }   // (positions past the end: every later one too)
                if (lane == 0) {
                    const int s = (int)(it % S);
                    if (it >= S) mbar_wait(empty + s, ((it / S) & 1u) ^ 1u);
                    ent_s[s] = (uint64_t)lo | ((uint64_t)hi << 32);
                    mbar_expect_tx(full + s, 16384u);
                    tma_load_1d(smem + s * 16384, F.gcache + ((size_t)lo * F.G + ((hi >> 8) & 255u)) * 4096, 16384u, full + s);
                }
                ++it;
            }
        }
        if (lane == 0)
            for (int k = 0; k < S; ++k, ++it) {   // every compute group gets its own EXIT
                const int s = (int)(it % S);
                if (it >= S) mbar_wait(empty + s, ((it / S) & 1u) ^ 1u);
                ent_s[s] = EXIT;
                mbar_arrive(full + s);
            }
        return;
    }
    const int g = (tid - 32) / 160, ct = tid - 32 - g * 160;
    if (g >= S) return;
    const uint8_t* icon = smem + g * 16384;
    for (uint32_t k = 0;; ++k) {   // the group's k-th cell = stage g of iteration k * S + g
        mbar_wait(full + g, k & 1u);
        const uint64_t ent = ent_s[g];
        if (ent == EXIT) break;
        const uint32_t lo = (uint32_t)ent, hi = (uint32_t)(ent >> 32);
        const int e = (int)lo, blk = hi & 255, facing = (hi >> 16) & 3;
        const int by = blk / vr, bx = blk - by * vr;
        const uint4* t4 = (const uint4*)F.taps4 + (size_t)facing * 4 * plane;
        uint8_t* out = frames + (size_t)e * env_stride;
        for (int q = ct; q < bb; q += 160) {
            const int qy = q / bs, p = (by * bs + qy) * OW + bx * bs + (q - qy * bs);
            const uint4 q0 = __ldg(t4 + p), q1 = __ldg(t4 + (size_t)plane + p), q2 = __ldg(t4 + 2 * (size_t)plane + p), q3 = __ldg(t4 + 3 * (size_t)plane + p);
            const uint32_t v = xw_fpv_px_taps_dev(q0, q1, q2, q3, icon);
            out[p] = (uint8_t)v; out[plane + p] = (uint8_t)(v >> 8); out[2 * (size_t)plane + p] = (uint8_t)(v >> 16);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + g);   // (the group's five warps are whole warps: 32 + 160 g is a multiple of 32)
    }
}

// --color=false: cv::cvtColor(BGR2GRAY) of the finished frame, OpenCV 3.2.0's coefficients (include/xworld_b200.h xw_config.gray)
__global__ void k_gray(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ out, int n, int plane, size_t in_stride, size_t out_stride) {
    const size_t total = (size_t)n * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i / plane, p = i % plane;
        const uint8_t* s = bgr + e * in_stride + p;
        out[e * out_stride + p] = (uint8_t)((s[0] * 1868 + s[plane] * 9617 + s[2 * (size_t)plane] * 4899 + (1 << 13)) >> 14);
    }
}
#endif  // __CUDACC__
