// xw_render_host.hpp -- host-side construction of the renderer's lookup tables.
// cv::resize(INTER_LINEAR, 8U) coefficient tables as OpenCV 3.2 computes them (imgproc/resize.cpp:
// scale = 1/(dst/src) in double, fx in float, cvFloor, 11-bit rounded weights), plus the plan the
// compositing kernel executes: one "band-column" item per (cell row, output word column).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

// One compositor work item = one 4-pixel word column of one colour plane over a run of rows.  Items are
// typed so that a warp executes one code path (the plan lists them sorted by type, planes expanded):
//   M1  rows of a band (cell row), word inside ONE cell:            word = table(A)
//   M2  rows of a band, word spanning two cells:                    word = PRMT(table(A), table(B))
//   M3  like M2, and one byte lies on a straddling column (its two tap columns are in different
//       cells): that byte is evaluated exactly from the cells' edge-tap tables
//   R   one word of a straddling row (tap rows in different cell rows): separable U(top)+V(bottom) rule
//   RC  like R, and one byte is a corner (straddling row x straddling column): exact from the atlas
// The M items of a band exclude the band's straddling last row, so every frame word has one writer.
enum { XW_ITEM_M1 = 0, XW_ITEM_M2 = 1, XW_ITEM_M3 = 2, XW_ITEM_R = 3, XW_ITEM_RC = 4, XW_ITEM_TYPES = 5 };

// Packed plan entry (one LDS.64):
//   x = cellA | cellB << 8 | sel << 16        cell indices ty*W + tx owning the word's first / last byte
//                                             (top cell row for R/RC), PRMT selector merging word(A), word(B)
//   y = woff | nrows << 16 | type << 24 | c << 27     word offset of the first row in the FRAME, rows, plane
// Items of type >= M3 also have an aux entry (index = plan index - seg[M3]):
//   x = y0 | dx << 8 | scell << 16 | sbyte << 24     first row; the straddling column, its left cell
//                                                    (the right one is scell + 1), its byte in the word
//   y = q | k << 8                                   R/RC: index of the straddling row; word column
struct alignas(8) XwU2 { uint32_t x, y; };

struct XwRenderTables {
    int H = 0, W = 0, OH = 0, OW = 0, WR = 0, FB = 0;
    bool fast_ok = false;  // the shared-memory compositor applies (else: generic kernel)
    std::vector<int16_t> xofs, xa0, xa1, yofs, ya0, ya1, sc, sr;
    std::vector<XwU2> plan, aux;         // plan sorted by type; aux for plan[seg[M3]..]
    int seg[XW_ITEM_TYPES + 1] = {0};    // items of type t are plan[seg[t] .. seg[t+1])
};

inline void xw_resize_tables(int src, int dst, int16_t* ofs, int16_t* a0, int16_t* a1) {
    const double inv_scale = (double)dst / (double)src;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        ofs[d] = (int16_t)s;
        a0[d] = (int16_t)lrintf((1.f - f) * 2048.f);
        a1[d] = (int16_t)lrintf(f * 2048.f);
    }
}

inline XwRenderTables xw_build_render_tables(int H, int W, int OH, int OW) {
    XwRenderTables t;
    t.H = H; t.W = W; t.OH = OH; t.OW = OW;
    t.FB = 3 * OH * OW;
    t.xofs.resize(OW); t.xa0.resize(OW); t.xa1.resize(OW);
    t.yofs.resize(OH); t.ya0.resize(OH); t.ya1.resize(OH);
    xw_resize_tables(W * 64, OW, t.xofs.data(), t.xa0.data(), t.xa1.data());
    xw_resize_tables(H * 64, OH, t.yofs.data(), t.ya0.data(), t.ya1.data());
    std::vector<uint8_t> is_sc(OW, 0), is_sr(OH, 0);
    for (int dx = 0; dx < OW; ++dx)
        if (t.xa1[dx] != 0 && ((t.xofs[dx] + 1) >> 6) != (t.xofs[dx] >> 6)) { t.sc.push_back((int16_t)dx); is_sc[dx] = 1; }
    for (int dy = 0; dy < OH; ++dy)
        if (t.ya1[dy] != 0 && ((t.yofs[dy] + 1) >> 6) != (t.yofs[dy] >> 6)) { t.sr.push_back((int16_t)dy); is_sr[dy] = 1; }
    bool ok = (OW % 4 == 0) && (t.FB % 16 == 0) && OH <= 255 && OW <= 252 && OH * (OW / 4) <= 65535 && t.sr.size() <= 255;
    // a straddling row is the last row of its band (yofs is monotone); two in one band (upscaling)
    // would break the plan
    for (size_t q = 0; q < t.sr.size() && ok; ++q) {
        int dy = t.sr[q];
        if (dy + 1 < OH && (t.yofs[dy + 1] >> 6) == (t.yofs[dy] >> 6)) ok = false;
    }
    if (!ok) return t;
    t.WR = OW / 4;
    struct Item { int cellA, cellB, sel, woff, nrows, type, y0, k, sbyte, scell, dx, q; };
    std::vector<Item> all;
    for (int ty = 0; ty < H && ok; ++ty) {
        int y0 = -1, y1 = -1;  // rows owned by cell row ty
        for (int dy = 0; dy < OH; ++dy)
            if ((t.yofs[dy] >> 6) == ty) { if (y0 < 0) y0 = dy; y1 = dy + 1; }
        if (y0 < 0) continue;
        const bool srow = is_sr[y1 - 1] != 0;
        int q = 0;
        if (srow) while (t.sr[q] != y1 - 1) ++q;
        for (int k = 0; k < t.WR && ok; ++k) {
            Item it;
            memset(&it, 0, sizeof it);
            int tx[4];
            for (int i = 0; i < 4; ++i) tx[i] = t.xofs[4 * k + i] >> 6;
            const int A = tx[0], B = tx[3];
            int n_sc = 0;
            for (int i = 0; i < 4; ++i) {
                if (tx[i] == A) it.sel |= i << (4 * i);
                else if (tx[i] == B) it.sel |= (4 + i) << (4 * i);
                else ok = false;  // a 4-pixel word spans 3 cells: cells narrower than 2 px
                if (is_sc[4 * k + i]) { ++n_sc; it.sbyte = i; it.scell = ty * W + tx[i]; it.dx = 4 * k + i; }
            }
            if (n_sc > 1) ok = false;  // two straddling columns in one word: cells narrower than 4 px
            it.cellA = ty * W + A; it.cellB = ty * W + B;
            it.woff = y0 * t.WR + k;
            it.nrows = y1 - y0 - (srow ? 1 : 0);
            it.y0 = y0; it.k = k;
            it.type = n_sc ? XW_ITEM_M3 : (A != B ? XW_ITEM_M2 : XW_ITEM_M1);
            if (it.nrows > 0) all.push_back(it);
            if (srow) {
                Item r = it;
                r.woff = (y1 - 1) * t.WR + k;
                r.nrows = 1; r.y0 = y1 - 1; r.q = q;
                r.type = n_sc ? XW_ITEM_RC : XW_ITEM_R;
                all.push_back(r);
            }
        }
    }
    if (3 * OH * t.WR > 65535) ok = false;  // woff is 16 bits
    for (int ty = 0; ty < XW_ITEM_TYPES && ok; ++ty) {
        t.seg[ty] = (int)t.plan.size();
        for (int c = 0; c < 3; ++c)
            for (const Item& it : all) {
                if (it.type != ty) continue;
                XwU2 e, a;
                e.x = (uint32_t)it.cellA | ((uint32_t)it.cellB << 8) | ((uint32_t)it.sel << 16);
                e.y = (uint32_t)(c * OH * t.WR + it.woff) | ((uint32_t)it.nrows << 16) | ((uint32_t)it.type << 24) | ((uint32_t)c << 27);
                t.plan.push_back(e);
                if (ty >= XW_ITEM_M3) {
                    a.x = (uint32_t)it.y0 | ((uint32_t)it.dx << 8) | ((uint32_t)it.scell << 16) | ((uint32_t)it.sbyte << 24);
                    a.y = (uint32_t)it.q | ((uint32_t)it.k << 8);
                    t.aux.push_back(a);
                }
            }
    }
    t.seg[XW_ITEM_TYPES] = (int)t.plan.size();
    t.fast_ok = ok;
    return t;
}
