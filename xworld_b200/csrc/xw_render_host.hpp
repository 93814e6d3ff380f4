// xw_render_host.hpp -- host-side construction of the renderer's lookup tables.
// cv::resize(INTER_LINEAR, 8U) coefficient tables as OpenCV 3.2 computes them (imgproc/resize.cpp:
// scale = 1/(dst/src) in double, fx in float, cvFloor, 11-bit rounded weights), plus the cell
// ownership maps the compositing kernel uses.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

struct XwRenderTables {
    int H = 0, W = 0, OH = 0, OW = 0, WR = 0, FB = 0, R = 1, rpg = 0, threads = 0;
    bool fast_ok = false;  // the shared-memory compositor applies (else: generic kernel)
    std::vector<int16_t> xofs, xa0, xa1, yofs, ya0, ya1, sc, sr;
    std::vector<uint8_t> rowcell, bandend;
    std::vector<uint32_t> colpair;
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
    for (int dx = 0; dx < OW; ++dx)
        if (t.xa1[dx] != 0 && ((t.xofs[dx] + 1) >> 6) != (t.xofs[dx] >> 6)) t.sc.push_back((int16_t)dx);
    for (int dy = 0; dy < OH; ++dy)
        if (t.ya1[dy] != 0 && ((t.yofs[dy] + 1) >> 6) != (t.yofs[dy] >> 6)) t.sr.push_back((int16_t)dy);
    t.rowcell.resize(OH);
    t.bandend.assign(H, (uint8_t)0);
    bool ok = (OW % 4 == 0) && (t.FB % 16 == 0) && OH <= 255 && OW <= 1020;
    for (int dy = 0; dy < OH; ++dy) t.rowcell[dy] = (uint8_t)(t.yofs[dy] >> 6);
    for (int ty = 0; ty < H; ++ty) {
        int e = OH;
        for (int dy = 0; dy < OH; ++dy) if (t.rowcell[dy] > ty) { e = dy; break; }
        t.bandend[ty] = (uint8_t)(e > 255 ? 255 : e);
    }
    if (ok) {
        t.WR = OW / 4;
        t.colpair.resize(t.WR);
        for (int k = 0; k < t.WR && ok; ++k) {
            int tx[4];
            for (int i = 0; i < 4; ++i) tx[i] = t.xofs[4 * k + i] >> 6;
            const int A = tx[0], B = tx[3];
            uint32_t sel = 0;
            for (int i = 0; i < 4; ++i) {
                if (tx[i] == A) sel |= (uint32_t)i << (4 * i);
                else if (tx[i] == B) sel |= (uint32_t)(4 + i) << (4 * i);
                else ok = false;  // a 4-pixel word spans 3 cells: cells narrower than 2 px
            }
            t.colpair[k] = (uint32_t)A | ((uint32_t)B << 8) | (sel << 16);
        }
    }
    if (ok) {
        int R = (int)lround(256.0 / (3.0 * t.WR));
        if (R < 1) R = 1;
        while (3 * R * t.WR > 1024) --R;
        if (R < 1) ok = false;
        t.R = R < 1 ? 1 : R;
        t.rpg = (OH + t.R - 1) / t.R;
        int items = 3 * t.R * t.WR;
        if (items < H * W) items = H * W;
        t.threads = (items + 31) / 32 * 32;
        if (t.threads > 1024) ok = false;
    }
    t.fast_ok = ok;
    return t;
}
