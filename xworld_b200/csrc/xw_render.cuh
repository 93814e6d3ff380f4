// xw_render.cuh -- the observation renderer: XWorldSimulator::get_screen for every env.
//
// Replaces (reference file:line):
//   XMap::to_image (fully observed)       games/xworld/xworld/xmap.cpp:125-146,201-205
//   XItem::get_item_image                 games/xworld/xworld/xitem.cpp:33-63   (identity warp)
//   XWorldSimulator::get_screen_rgb       games/xworld/xworld_simulator.cpp:287-307
//   XWorldSimulator::down_sample_image    games/xworld/xworld_simulator.cpp:508-545  (cv::resize INTER_LINEAR, 8U)
//
// The reference paints a (H*64)x(W*64) canvas and bilinearly resizes it.  Here the canvas never
// exists.  cv::resize's sample positions depend only on the output coordinate, so every output
// pixel whose 2x2 taps fall inside one cell is a pure function of (icon in that cell, pixel):
//     T[icon][c][dy][dx] = resize(canvas tiled with `icon`)[c][dy][dx]         ("phase atlas")
// is built once per handle with the exact fixed-point arithmetic, and a frame is then a per-cell
// SELECT from those tables, done one 4-pixel word at a time: word(A) and word(B) of the two cells a
// word can touch are merged with one PRMT.  The few output columns / rows whose taps straddle a
// cell border (2 of 84 each at 11x11->84, none at 7x7->84) are evaluated exactly from two small
// per-icon edge tables: raw edge taps for straddling columns, and for straddling rows the two
// vertical partial terms U, V of cv::resize's  (U + V + 2) >> 2  -- the rule is separable there.
//
// Kernel shape (k_render): one persistent CTA per SM split into G warp groups; each group composes
// one env's frame at a time in its own shared-memory frame buffer, following a precomputed plan of
// typed work items (xw_render_host.hpp) so that warps do not diverge, and hands the frame to the
// TMA engine (cp.async.bulk.global.shared::cta, one bulk store per frame, full-line HBM writes)
// while the other groups keep composing.  The brick table -- the icon most non-empty cells hold --
// is staged into shared memory once per CTA by a TMA bulk load; other tables stay L2-resident.
#pragma once
#include <stdlib.h>
#include "xw_common.cuh"
#include "xw_render_host.hpp"

#define XW_MAX_OUT 252        // max frame side
#define XW_RENDER_THREADS 1024
#define XW_RENDER_MAX_GROUPS 12   // groups per CTA (named barriers 1..12)
#define XW_TABLE_PAD 8192     // bytes past each table the compositor may read (and never use)

// The painter's experiment switches.  XwRender::debug can only become non-zero in a -DXW_SP_DEBUG build (xw_engine.cu reads
// XW_RENDER_DEBUG there and nowhere else); a shipped build always passes 0.  The tests of the (always zero) field stay in
// the kernel on purpose: compiling them away (-DXW_SP_NODEBUG) changes ptxas's schedule of k_render_sp (123 -> 116
// registers, +6 % executed instructions) and costs 14 % of the kernel's time (profiles/r02_summary.md, r02c vs r02d).
#if defined(XW_SP_NODEBUG)
#define XW_DBG(mask) 0
#else
#define XW_DBG(mask) (r.debug & (mask))
#endif

struct XwTaps { const int16_t *xofs, *xa0, *xa1, *yofs, *ya0, *ya1; };  // cv::resize tables

// shared-memory layout of k_render_sp (byte offsets; filled by the host with xw_render_sp_smem so that the kernel
// reads them from the constant bank instead of keeping a dozen derived pointers in registers)
struct XwRenderSpSmem { int ctab, fb, cellgeo, wcol, wshare, srq, sca, yb, pair, cell, bar, total; };

struct XwRender {
    int32_t OH, OW, WR, FB;   // frame rows, cols, words per row, bytes per frame (3*OH*OW)
    int32_t H, W;
    int32_t n_plan;           // plan slots (a multiple of GT)
    int32_t n_plan1;          // the first n_plan1 slots (straddling rows) do not read staged words
    int32_t G, GT;            // warp groups per CTA, threads per group
    int32_t n_icons, brick_icon, agent_icon;
    int32_t n_sr;             // straddling rows
    int32_t debug;            // -DXW_SP_DEBUG builds only (tuning experiments, env XW_RENDER_DEBUG): 1 skip compose, 2 skip staging /
                              // the white fill, 4 skip the frame store, 8 no special cells, 16 no straddling-row words, 32 no special
                              // finish, 64 no issue.  The shipped build compiles every test of it away (XW_DBG == 0): an environment
                              // variable can never switch parts of the frame off.
    XwTaps taps;
    const int16_t* sr;        // [n_sr] the straddling rows
    const XwU4* plan;         // [n_plan] packed items (xw_render_host.hpp)
    const uint8_t* T;         // [n_icons][3][OH][OW] phase atlas
    const uint32_t* cellinfo; // [H*W] staging geometry of each cell (xw_stage_column)
    // edge tables, indexed by cell descriptor (0 = white, icon + 1 otherwise)
    const uint16_t* ecol;     // [n_icons+1][2][3][H][RB]  role 0: taps of icon column 63, role 1: column 0, for
                              //   row j of band ty; low byte = tap row of yofs[dy], high byte = the row below it
    const int16_t* band_y0;   // [H] first output row of each band
    int32_t RB;               // rows per band in ecol (a multiple of 4: one 8-byte load = 4 rows)
    const uint16_t* uv;       // [n_icons+1][n_sr][2][3][OW]  role 0: U from icon row 63, role 1: V from row 0
    const uint32_t* corner;   // [n_icons+1][3]  icon pixels (63,63) | (63,0) << 8 | (0,63) << 16 | (0,0) << 24
    // pair tables: the finished pixel when the cell on the other side of the border is white (cls 0) or
    // a brick (cls 1) -- all but a few per cent of the borders of a maze
    const uint8_t *colL, *colR;  // [n_icons+1][2][n_sc][3][H][RB]  this cell left / right of straddling column s
    const uint8_t *rowT, *rowB;  // [n_icons+1][2][n_sr][3][OW]     this cell above / below straddling row q
    const uint8_t* cornerWB;     // [16][n_sr][n_sc][3]  corner pixel when the four cells are white / brick:
                                 //   combo = cls(top-left) | cls(top-right) << 1 | cls(bottom-left) << 2 | cls(bottom-right) << 3
    const int16_t* sc;        // [n_sc] the straddling columns
    int32_t n_sc;
    const uint8_t* atlas64;   // [n_icons][64][64][3] BGR
    // sparse painter (k_render_sp): static geometry, see xw_build_paint_tables (xw_render_host.hpp)
    const XwU4* cellgeo;      // [H*W]
    const uint32_t* wcol;     // [WR]
    const uint8_t* white;     // [FB] 0xff: source of the TMA pre-fill
    const uint8_t* wshare;    // [WR]
    const uint8_t* sr_ty;     // [n_sr]
    const uint32_t* ctab;     // class tables (xw_ctab_words), built by k_build_class_tables
    int32_t ns;               // word columns shared by two cell columns that have variant class tables (all of them, or --
                              //   ctab_merge -- only those that hold a straddling pixel)
    int32_t ctab_merge;       // 1: a brick|white / white|brick word of a shared column WITHOUT a straddling pixel is the brick|brick
                              //   word with the white cell's bytes set (one PRMT with the column's byte-ownership selector) instead
                              //   of a table of its own.  Chosen when that is what makes a third frame buffer fit (15x15 -> 128x128:
                              //   class tables 86 -> 55 KB); costs an instruction per such word, so not otherwise.
    int32_t nwc;              // word-column slots per cell
    int32_t slot_magic;       // slot / nwc == (slot * slot_magic) >> 16 for slot < 4096
    const uint32_t* cornerP;  // [n_icons+1][4] corner tap of role (top-left cell's (63,63), top-right's (63,0), bottom-left's
                              //   (0,63), bottom-right's (0,0)), planes 0..2 in bytes 0..2
    const uint32_t *rowT2, *rowB2;  // [n_icons+1][2][n_sr][WR][3] the painter's copy of rowT / rowB: the three planes of a word are
                              //   adjacent (one L2 sector per cell column of a straddling-row word instead of three)
    const uint32_t* TC;       // [n_icons][H*W][3][nwc][tc_rows] cell-major copy of the phase atlas: the words a special slot-plane
                              //   reads (band rows of word column wc of the cell in plane p) are tc_rows consecutive words --
                              //   one or two 32-byte sectors instead of one per row (the painter is bound by the number of L2
                              //   requests in flight, not by their bytes)
    int32_t tc_rows;          // 8 or 12 = the kernel's XW_SP_ROWS
    XwRenderSpSmem sp;        // = xw_render_sp_smem(r, G)
    unsigned int* prof;       // -DXW_SP_PROF builds only: per-phase clock sums of group 0 of CTA 0 (tools/sweep_render.py)
    int32_t sp_fill;          // 0: white pre-fill with vector stores (warp 1), 1: with a TMA bulk load of `white`
    // list mode of the painter (k_render_sp<..., LIST = true>): the envs env_list[0 .. *env_count) instead of all of them --
    // the frames of the step's auto-reset queue, painted after the reset launch that ran beside the painter (xw_engine.cu step_xworld)
    const int32_t* env_list;
    const int32_t* env_count;
};

// ---- exact cv::resize arithmetic --------------------------------------------------------
// One output pixel of the bilinear resize given its four taps (HResizeLinear + VResizeLinear,
// INTER_RESIZE_COEF_BITS = 11, FixedPtCast shift 22 split as >>4, >>16, +2, >>2).
XW_HD int xw_vterm(int p0, int p1, int a0, int a1, int b) { return (b * ((p0 * a0 + p1 * a1) >> 4)) >> 16; }
XW_HD uint8_t xw_resize_px(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    return (uint8_t)((xw_vterm(p00, p01, a0, a1, b0) + xw_vterm(p10, p11, a0, a1, b1) + 2) >> 2);
}

// The cells of one env as the compositor reads them: the grid's cell codes (one byte per cell, so a
// whole 16x16 map spans 64 banks' worth of words at most twice -- lookups rarely conflict) and the
// descriptor (0 = white, icon + 1) each code stands for in this env.
struct XwCells {
    const uint8_t* code;    // [XW_CELL_STRIDE] XW_CELL_* per cell; zero past the map
    const uint32_t* icon;   // [XW_CELL_GOAL0 + XW_MAX_GOALS] descriptor per code
    XW_HD uint32_t operator()(int cell) const { return icon[code[cell]]; }
};

// One tap of the virtual canvas: cell descriptor (icon+1, 0 = empty = white) -> pixel.
XW_HD int xw_canvas_tap(const XwRender& r, uint32_t dsc, int sy, int sx, int c) {
    if (dsc == 0) return 255;
    return r.atlas64[(size_t)(dsc - 1) * (64 * 64 * 3) + ((sy & 63) * 64 + (sx & 63)) * 3 + c];
}

// Phase-atlas entry: canvas tiled with `icon` everywhere.
XW_HD uint8_t xw_phase_px(const XwRender& r, int icon, int c, int dy, int dx) {
    const XwTaps& t = r.taps;
    const uint32_t dsc = (uint32_t)icon + 1;
    int sx0 = t.xofs[dx], sy0 = t.yofs[dy];
    int a1 = t.xa1[dx], b1 = t.ya1[dy];
    int sx1 = a1 ? sx0 + 1 : sx0, sy1 = b1 ? sy0 + 1 : sy0;
    return xw_resize_px(xw_canvas_tap(r, dsc, sy0, sx0, c), xw_canvas_tap(r, dsc, sy0, sx1, c),
                        xw_canvas_tap(r, dsc, sy1, sx0, c), xw_canvas_tap(r, dsc, sy1, sx1, c),
                        t.xa0[dx], a1, t.ya0[dy], b1);
}

// Edge-table entries (built once per handle for every descriptor).
XW_HD uint16_t xw_ecol_entry(const XwRender& r, uint32_t dsc, int role, int c, int band, int j) {
    const XwTaps& t = r.taps;
    const int dy = r.band_y0[band] + j;
    if (r.band_y0[band] < 0 || dy >= r.OH) return 0;
    int sy0 = t.yofs[dy], sy1 = t.ya1[dy] ? sy0 + 1 : sy0;
    int sx = role == 0 ? 63 : 0;
    return (uint16_t)(xw_canvas_tap(r, dsc, sy0, sx, c) | (xw_canvas_tap(r, dsc, sy1, sx, c) << 8));
}
XW_HD uint16_t xw_uv_entry(const XwRender& r, uint32_t dsc, int q, int role, int c, int dx) {
    const XwTaps& t = r.taps;
    const int dy = r.sr[q];
    int sx0 = t.xofs[dx], sx1 = t.xa1[dx] ? sx0 + 1 : sx0;
    int sy = role == 0 ? 63 : 0;
    return (uint16_t)xw_vterm(xw_canvas_tap(r, dsc, sy, sx0, c), xw_canvas_tap(r, dsc, sy, sx1, c), t.xa0[dx], t.xa1[dx],
                              role == 0 ? t.ya0[dy] : t.ya1[dy]);
}

// descriptor class: 0 white, 1 brick, 2 anything else ("special": agent, goals)
XW_HD int xw_cls(const XwRender& r, uint32_t dsc) { return dsc == 0 ? 0 : ((int)dsc - 1 == r.brick_icon ? 1 : 2); }
XW_HD uint32_t xw_cls_desc(const XwRender& r, int cls) { return cls == 0 ? 0u : (uint32_t)r.brick_icon + 1; }
// role 0: `dsc` is the LEFT cell and class `cls` the right one; role 1: the other way round
XW_HD uint8_t xw_colpair_entry(const XwRender& r, int role, uint32_t dsc, int cls, int s, int c, int band, int j) {
    const XwTaps& t = r.taps;
    const int dy = r.band_y0[band] + j, dx = r.sc[s];
    if (r.band_y0[band] < 0 || dy >= r.OH) return 0;
    const uint32_t dl = role == 0 ? dsc : xw_cls_desc(r, cls), dr = role == 0 ? xw_cls_desc(r, cls) : dsc;
    const int sy0 = t.yofs[dy], sy1 = t.ya1[dy] ? sy0 + 1 : sy0;
    return xw_resize_px(xw_canvas_tap(r, dl, sy0, 63, c), xw_canvas_tap(r, dr, sy0, 0, c), xw_canvas_tap(r, dl, sy1, 63, c),
                        xw_canvas_tap(r, dr, sy1, 0, c), t.xa0[dx], t.xa1[dx], t.ya0[dy], t.ya1[dy]);
}
// role 0: `dsc` is the cell ABOVE straddling row q and class `cls` the one below; role 1: the other way round
XW_HD uint8_t xw_rowpair_entry(const XwRender& r, int role, uint32_t dsc, int cls, int q, int c, int dx) {
    const XwTaps& t = r.taps;
    const int dy = r.sr[q];
    const uint32_t dt = role == 0 ? dsc : xw_cls_desc(r, cls), db = role == 0 ? xw_cls_desc(r, cls) : dsc;
    const int sx0 = t.xofs[dx], sx1 = t.xa1[dx] ? sx0 + 1 : sx0;
    return xw_resize_px(xw_canvas_tap(r, dt, 63, sx0, c), xw_canvas_tap(r, dt, 63, sx1, c), xw_canvas_tap(r, db, 0, sx0, c),
                        xw_canvas_tap(r, db, 0, sx1, c), t.xa0[dx], t.xa1[dx], t.ya0[dy], t.ya1[dy]);
}

XW_HD uint8_t xw_cornerwb_entry(const XwRender& r, int combo, int q, int s, int c) {
    const XwTaps& t = r.taps;
    const int dy = r.sr[q], dx = r.sc[s];
    return xw_resize_px(xw_canvas_tap(r, xw_cls_desc(r, combo & 1), 63, 63, c), xw_canvas_tap(r, xw_cls_desc(r, (combo >> 1) & 1), 63, 0, c),
                        xw_canvas_tap(r, xw_cls_desc(r, (combo >> 2) & 1), 0, 63, c), xw_canvas_tap(r, xw_cls_desc(r, (combo >> 3) & 1), 0, 0, c),
                        t.xa0[dx], t.xa1[dx], t.ya0[dy], t.ya1[dy]);
}
XW_HD uint32_t xw_cornerP_entry(const XwRender& r, uint32_t dsc, int role) {
    uint32_t v = 0;
    for (int c = 0; c < 3; ++c) v |= (uint32_t)xw_canvas_tap(r, dsc, role < 2 ? 63 : 0, (role & 1) ? 0 : 63, c) << (8 * c);
    return v;
}
XW_HD uint32_t xw_corner_entry(const XwRender& r, uint32_t dsc, int c) {
    return (uint32_t)xw_canvas_tap(r, dsc, 63, 63, c) | ((uint32_t)xw_canvas_tap(r, dsc, 63, 0, c) << 8) |
           ((uint32_t)xw_canvas_tap(r, dsc, 0, 63, c) << 16) | ((uint32_t)xw_canvas_tap(r, dsc, 0, 0, c) << 24);
}

// Value of output pixel (c,dy,dx) on the real canvas, from the 64-px atlas (any four cells).
template <class CellDesc>
XW_HD uint8_t xw_exact_px(const XwRender& r, const CellDesc& celldesc, int c, int dy, int dx) {
    const XwTaps& t = r.taps;
    int sx0 = t.xofs[dx], sy0 = t.yofs[dy];
    int a1 = t.xa1[dx], b1 = t.ya1[dy];
    int sx1 = a1 ? sx0 + 1 : sx0, sy1 = b1 ? sy0 + 1 : sy0;
    uint32_t d00 = celldesc((sy0 >> 6) * r.W + (sx0 >> 6)), d01 = celldesc((sy0 >> 6) * r.W + (sx1 >> 6));
    uint32_t d10 = celldesc((sy1 >> 6) * r.W + (sx0 >> 6)), d11 = celldesc((sy1 >> 6) * r.W + (sx1 >> 6));
    return xw_resize_px(xw_canvas_tap(r, d00, sy0, sx0, c), xw_canvas_tap(r, d01, sy0, sx1, c),
                        xw_canvas_tap(r, d10, sy1, sx0, c), xw_canvas_tap(r, d11, sy1, sx1, c),
                        t.xa0[dx], a1, t.ya0[dy], b1);
}

XW_HD uint32_t xw_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t o = 0;
    for (int i = 0; i < 4; ++i) o |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return o;
#endif
}

XW_HD int xw_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// What the compositor reads besides the per-env cells: shared-memory copies on the device.
struct XwComposeCtx {
    const uint8_t* hot;     // brick phase table [FB]
    const uint32_t* yb;     // [OH] ya0 | ya1 << 16
    const uint8_t *colL_hot, *rowT_hot;  // the white and brick entries of colL / rowT: [2 descs][2 cls][stride]
    const uint8_t* cornerWB;             // copy of XwRender::cornerWB
};
XW_HD size_t xw_colpair_stride(const XwRender& r) { return (size_t)r.n_sc * 3 * r.H * r.RB; }  // per (desc, cls)
XW_HD size_t xw_rowpair_stride(const XwRender& r) { return (size_t)r.n_sr * 3 * r.OW; }
XW_HD int xw_pair_hot_bytes(const XwRender& r) {  // shared-memory copies: desc in {white, brick} x cls x ...
    return (int)(2 * 2 * (xw_colpair_stride(r) + xw_rowpair_stride(r))) + 16 * r.n_sr * r.n_sc * 3;
}

// Source of a cell's words: word w of the frame comes from *(base + 4*w) | wmask.  White cells read
// the brick table and OR it to 0xffffffff, so that every lane runs the same instructions.  The words
// of the few other icons of an env (agent, goals: "special" cells) are staged into the frame buffer
// itself before the items run (xw_stage_column), at the offsets where the items then read them.
struct XwSrc { const uint8_t* base; uint32_t wmask; };
XW_HD bool xw_special(const XwRender& r, uint32_t dsc) { return dsc != 0 && (int)dsc - 1 != r.brick_icon; }
// staged_elsewhere: the word was staged from the OTHER cell's table (both cells of the word are
// special): read this one straight from the phase atlas.
XW_HD XwSrc xw_src_of(const XwRender& r, const XwComposeCtx& x, uint32_t dsc, const uint32_t* fb, bool staged_elsewhere) {
    XwSrc s;
    s.wmask = dsc == 0 ? 0xffffffffu : 0u;
    if (!xw_special(r, dsc)) s.base = x.hot;
    else s.base = staged_elsewhere ? r.T + (size_t)(dsc - 1) * r.FB : (const uint8_t*)fb;
    return s;
}

// Staging plan of word column wc of special cell `cell` in plane c: false = nothing to copy (the cell
// has fewer word columns, or the column is shared with a special left neighbour, which stages it).
// Otherwise rows j < *nrows: frame word (*w0 + j*WR) <- word (*w0 + j*WR) of the cell's phase table.
// cellinfo[cell] = woff | nrows << 16 | nwords << 24 | first_shared << 26 (xw_render_host.hpp).
XW_HD bool xw_stage_column(const XwRender& r, const XwCells& cells, const uint32_t* cellinfo, int cell, int c, int wc,
                           uint32_t* w0, int* nrows, const uint32_t** src) {
    const uint32_t info = cellinfo[cell];
    if (wc >= (int)((info >> 24) & 3)) return false;
    if (wc == 0 && ((info >> 26) & 1) && xw_special(r, cells(cell - 1))) return false;
    *w0 = (uint32_t)c * r.OH * r.WR + (info & 0xffffu) + wc;
    *nrows = (info >> 16) & 0xff;
    *src = (const uint32_t*)(r.T + (size_t)(cells(cell) - 1) * r.FB) + *w0;
    return true;
}
#define XW_STAGE_SLOTS (1 + XW_MAX_GOALS)  // agent + goals
#define XW_STAGE_COLS 9                    // 3 planes x up to 3 word columns per cell

// the corner pixel (straddling row x straddling column): taps (63,63) of the top-left cell, (63,0) of the
// top-right, (0,63) of the bottom-left, (0,0) of the bottom-right
XW_HD uint32_t xw_corner_px(const XwRender& r, const XwComposeCtx& x, const XwCells& celldesc, int scell, int c, int dy, int dx) {
    const uint32_t t00 = r.corner[celldesc(scell) * 3 + c], t01 = r.corner[celldesc(scell + 1) * 3 + c];
    const uint32_t t10 = r.corner[celldesc(scell + r.W) * 3 + c], t11 = r.corner[celldesc(scell + r.W + 1) * 3 + c];
    const uint32_t b = x.yb[dy];
    return xw_resize_px(t00 & 255, (t01 >> 8) & 255, (t10 >> 16) & 255, t11 >> 24, r.taps.xa0[dx], r.taps.xa1[dx], b & 0xffff, b >> 16);
}

// ---- exact (any cells) versions of the straddling items: used when BOTH cells across a border are
// special, which the pair tables do not cover.  Same item decode as xw_compose_item.
template <int WR_T>
XW_HD void xw_item_m3_exact(const XwRender& r, const XwComposeCtx& x, const XwU4 e, const XwCells& celldesc, uint32_t* fb) {
    const int WR = WR_T ? WR_T : r.WR;
    const int PW = r.OH * WR;
    const int nrows = (e.y >> 16) & 0xff, nc = e.y >> 29, c0 = (e.y >> 27) & 3;
    const uint32_t w0 = e.y & 0xffffu, sel = e.x >> 16;
    const int cellA = e.x & 0xff, cellB = (e.x >> 8) & 0xff;
    const int y0 = e.z & 0xff, dx = (e.z >> 8) & 0xff, scell = (e.z >> 16) & 0xff, sh = (e.z >> 24) * 8;
    const uint32_t dA = celldesc(cellA), dB = celldesc(cellB);
    const XwSrc sA = xw_src_of(r, x, dA, fb, false), sB = xw_src_of(r, x, dB, fb, xw_special(r, dA));
    const uint32_t wmask = xw_prmt(sA.wmask, sB.wmask, sel);
    const int band = (e.w >> 16) & 0xff;
    const uint16_t* eL0 = r.ecol + ((((size_t)celldesc(scell) * 2 + 0) * 3 + c0) * r.H + band) * r.RB;
    const uint16_t* eR0 = r.ecol + ((((size_t)celldesc(scell + 1) * 2 + 1) * 3 + c0) * r.H + band) * r.RB;
    const int a0 = r.taps.xa0[dx], a1 = r.taps.xa1[dx];
    const uint32_t keep = ~(0xffu << sh);
    for (int cc = 0; cc < nc; ++cc) {
        const uint32_t* pA = (const uint32_t*)sA.base + w0 + cc * PW;
        const uint32_t* pB = (const uint32_t*)sB.base + w0 + cc * PW;
        uint32_t* dst = fb + w0 + cc * PW;
        const uint16_t* eL = eL0 + (size_t)cc * r.H * r.RB;
        const uint16_t* eR = eR0 + (size_t)cc * r.H * r.RB;
        for (int j = 0; j < nrows; ++j) {
            const uint32_t tl = eL[j], tr = eR[j], b = x.yb[y0 + j];
            const uint32_t v = xw_resize_px(tl & 255, tr & 255, tl >> 8, tr >> 8, a0, a1, b & 0xffff, b >> 16);
            dst[j * WR] = ((xw_prmt(pA[j * WR], pB[j * WR], sel) | wmask) & keep) | (v << sh);
        }
    }
}
// one word of straddling row q: out = (U[top cell] + V[bottom cell] + 2) >> 2; corner byte from the corner taps
template <int WR_T>
XW_HD void xw_item_r_exact(const XwRender& r, const XwComposeCtx& x, const XwU4 e, const XwCells& celldesc, uint32_t* fb) {
    const int WR = WR_T ? WR_T : r.WR;
    const int PW = r.OH * WR;
    const int nc = e.y >> 29, c0 = (e.y >> 27) & 3;
    const uint32_t w0 = e.y & 0xffffu, sel = e.x >> 16;
    const int cellA = e.x & 0xff, cellB = (e.x >> 8) & 0xff;
    const int y0 = e.z & 0xff, dx = (e.z >> 8) & 0xff, scell = (e.z >> 16) & 0xff, sh = (e.z >> 24) * 8;
    const int q = e.w & 0xff, k = (e.w >> 8) & 0xff;
    const bool corner = (e.w >> 31) != 0;
    const size_t per_desc = (size_t)r.n_sr * 2 * 3 * r.OW;
    const uint16_t* u0 = r.uv + ((size_t)q * 2 * 3 + c0) * r.OW + 4 * k;
    const uint16_t* v0 = u0 + (size_t)3 * r.OW;
    const uint16_t *uAp = u0 + celldesc(cellA) * per_desc, *uBp = u0 + celldesc(cellB) * per_desc;
    const uint16_t *vAp = v0 + celldesc(cellA + r.W) * per_desc, *vBp = v0 + celldesc(cellB + r.W) * per_desc;
    for (int cc = 0; cc < nc; ++cc) {
        const XwU2 uA = *(const XwU2*)(uAp + cc * r.OW), uB = *(const XwU2*)(uBp + cc * r.OW);
        const XwU2 vA = *(const XwU2*)(vAp + cc * r.OW), vB = *(const XwU2*)(vBp + cc * r.OW);
        // packed u16 pairs: no carry between halves (U + V + 2 <= 2042)
        const uint32_t a_lo = ((uA.x + vA.x + 0x00020002u) >> 2) & 0x00ff00ffu, a_hi = ((uA.y + vA.y + 0x00020002u) >> 2) & 0x00ff00ffu;
        const uint32_t b_lo = ((uB.x + vB.x + 0x00020002u) >> 2) & 0x00ff00ffu, b_hi = ((uB.y + vB.y + 0x00020002u) >> 2) & 0x00ff00ffu;
        uint32_t word = xw_prmt(xw_prmt(a_lo, a_hi, 0x6420), xw_prmt(b_lo, b_hi, 0x6420), sel);
        if (corner) word = (word & ~(0xffu << sh)) | (xw_corner_px(r, x, celldesc, scell, c0 + cc, y0, dx) << sh);
        fb[w0 + cc * PW] = word;
    }
}

// ---- compose one plan item into the frame being built (fb, words) ----------------------------
// WR_T = words per frame row when known at compile time (row offsets become immediates), 0 = use r.WR.
// Rows go four at a time, all loads before the stores: the tables and the frame buffer may alias as
// far as the compiler knows, and a load-store-load chain would expose one memory latency per row.
template <int WR_T, int SMALL = 0>
XW_HD void xw_compose_item(const XwRender& r, const XwComposeCtx& x, const XwU4 e, const XwCells& celldesc, uint32_t* fb) {
    // rows per load batch: every source is shared memory once the special cells are staged, so small
    // batches (SMALL: half the registers, for CTAs of more than 512 threads) cost little
    constexpr int R1 = SMALL ? 4 : 8, R2 = SMALL ? 2 : 4;
    const int WR = WR_T ? WR_T : r.WR;
    const int PW = r.OH * WR;  // words per plane
    const int type = (e.y >> 24) & 7, nrows = (e.y >> 16) & 0xff, nc = e.y >> 29;
    const uint32_t w0 = e.y & 0xffffu, sel = e.x >> 16;
    const int cellA = e.x & 0xff, cellB = (e.x >> 8) & 0xff;
    // The goal / agent tables live in L2: every batch of loads below costs one L2 round trip for the
    // warp, so each batch is as wide as the register budget allows (24 words in flight).  Loads are
    // unconditional -- rows past the band read table bytes that are never stored (the tables are
    // padded by XW_TABLE_PAD) -- and only the stores are predicated, one predicate per row.
    if (nc == 0) return;  // padding slot
    // plane order of this lane: (rot + i) % 3, chosen by the planner so that the lanes of a bundle stay
    // on different banks (xw_build_plan); po[i] = word offset of the i-th plane visited
    const int rot = (e.w >> 24) & 3;
    const int pl[3] = {rot, rot == 2 ? 0 : rot + 1, rot == 0 ? 2 : rot - 1};
    const int po[3] = {pl[0] * PW, pl[1] * PW, pl[2] * PW};
    if (type == XW_ITEM_M1) {
        const XwSrc sA = xw_src_of(r, x, celldesc(cellA), fb, false);
        const uint32_t* pA = (const uint32_t*)sA.base + w0;
        uint32_t* dst = fb + w0;
        for (int i0 = 0; i0 < nrows; i0 += R1, pA += R1 * WR, dst += R1 * WR) {
            uint32_t v[3][R1];
#pragma unroll
            for (int j = 0; j < R1; ++j)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) v[cc][j] = pA[po[cc] + j * WR];
#pragma unroll
            for (int j = 0; j < R1; ++j)
                if (i0 + j < nrows) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) dst[po[cc] + j * WR] = v[cc][j] | sA.wmask;
                }
        }
        return;
    }
    if (type == XW_ITEM_M2) {
        const uint32_t dA = celldesc(cellA), dB = celldesc(cellB);
        const XwSrc sA = xw_src_of(r, x, dA, fb, false), sB = xw_src_of(r, x, dB, fb, xw_special(r, dA));
        const uint32_t wmask = xw_prmt(sA.wmask, sB.wmask, sel);
        const uint32_t* pA = (const uint32_t*)sA.base + w0;
        const uint32_t* pB = (const uint32_t*)sB.base + w0;
        uint32_t* dst = fb + w0;
        for (int i0 = 0; i0 < nrows; i0 += R2, pA += R2 * WR, pB += R2 * WR, dst += R2 * WR) {
            uint32_t va[3][R2], vb[3][R2];
#pragma unroll
            for (int j = 0; j < R2; ++j)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) { va[cc][j] = pA[po[cc] + j * WR]; vb[cc][j] = pB[po[cc] + j * WR]; }
#pragma unroll
            for (int j = 0; j < R2; ++j)
                if (i0 + j < nrows) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) dst[po[cc] + j * WR] = xw_prmt(va[cc][j], vb[cc][j], sel) | wmask;
                }
        }
        return;
    }
    const int c0 = (e.y >> 27) & 3;
    const int y0 = e.z & 0xff, dx = (e.z >> 8) & 0xff, scell = (e.z >> 16) & 0xff, sh = (e.z >> 24) * 8;
    if (type == XW_ITEM_M3) {
        // an M2 word whose byte `sh` comes from a pair table: (left cell, class of the right cell) or
        // (right cell, class of the left cell); both special -> the exact version
        const uint32_t dA = celldesc(cellA), dB = celldesc(cellB);
        const XwSrc sA = xw_src_of(r, x, dA, fb, false), sB = xw_src_of(r, x, dB, fb, xw_special(r, dA));
        const uint32_t wmask = xw_prmt(sA.wmask, sB.wmask, sel);
        const uint32_t* pA = (const uint32_t*)sA.base + w0;
        const uint32_t* pB = (const uint32_t*)sB.base + w0;
        uint32_t* dst = fb + w0;
        const int band = (e.w >> 16) & 0xff, s_idx = e.w & 0xff;
        const uint32_t dL = celldesc(scell), dR = celldesc(scell + 1);
        const int cL = xw_cls(r, dL), cR = xw_cls(r, dR);
        const size_t cs = xw_colpair_stride(r);
        const uint8_t* tab = x.colL_hot;
        if (cR < 2) tab = (cL < 2 ? x.colL_hot + (size_t)cL * 2 * cs : r.colL + (size_t)dL * 2 * cs) + (size_t)cR * cs;
        else if (cL < 2) tab = r.colR + (size_t)dR * 2 * cs + (size_t)cL * cs;
        const bool need_exact = cL == 2 && cR == 2;
        tab += ((size_t)(s_idx * 3 + c0) * r.H + band) * r.RB;
        const int pstride = r.H * r.RB;  // bytes between planes
        const uint32_t keep = ~(0xffu << sh);
        for (int i0 = 0; i0 < nrows; i0 += 4, pA += 4 * WR, pB += 4 * WR, dst += 4 * WR, tab += 4) {
            uint32_t pb[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                if (cc < nc) pb[cc] = *(const uint32_t*)(tab + pl[cc] * pstride);  // the straddling byte of four rows
#pragma unroll
            for (int h0 = 0; h0 < 4; h0 += R2) {
                uint32_t va[3][R2], vb[3][R2];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    if (cc < nc) {
#pragma unroll
                        for (int j = 0; j < R2; ++j) { va[cc][j] = pA[po[cc] + (h0 + j) * WR]; vb[cc][j] = pB[po[cc] + (h0 + j) * WR]; }
                    }
#pragma unroll
                for (int j = 0; j < R2; ++j)
                    if (i0 + h0 + j < nrows) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc)
                            if (cc < nc)
                                dst[po[cc] + (h0 + j) * WR] =
                                    ((xw_prmt(va[cc][j], vb[cc][j], sel) | wmask) & keep) | (((pb[cc] >> (8 * (h0 + j))) & 0xffu) << sh);
                    }
            }
        }
        if (need_exact) xw_item_m3_exact<WR_T>(r, x, e, celldesc, fb);
        return;
    }
    // R: one word of straddling row q, from the row pair tables: per cell column (top cell, class of the
    // bottom cell) or (bottom cell, class of the top cell); both special -> the exact version
    const int q = e.w & 0xff, k = (e.w >> 8) & 0xff;
    const bool corner = (e.w >> 31) != 0;
    const size_t rs = xw_rowpair_stride(r);
    bool need_exact = false;
    const uint8_t* tabs[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cell = h ? cellB : cellA;
        const uint32_t dT = celldesc(cell), dBt = celldesc(cell + r.W);
        const int cT = xw_cls(r, dT), cB = xw_cls(r, dBt);
        const uint8_t* tab = x.rowT_hot;
        if (cB < 2) tab = (cT < 2 ? x.rowT_hot + (size_t)cT * 2 * rs : r.rowT + (size_t)dT * 2 * rs) + (size_t)cB * rs;
        else if (cT < 2) tab = r.rowB + (size_t)dBt * 2 * rs + (size_t)cT * rs;
        else need_exact = true;
        tabs[h] = tab + (size_t)(q * 3 + c0) * r.OW + 4 * k;
    }
    if (need_exact) { xw_item_r_exact<WR_T>(r, x, e, celldesc, fb); return; }
    int corner_combo = -1;  // all four corner cells white / brick: one table byte per plane
    const int s_idx = (e.w >> 26) & 31;
    if (corner) {
        const int c00 = xw_cls(r, celldesc(scell)), c01 = xw_cls(r, celldesc(scell + 1));
        const int c10 = xw_cls(r, celldesc(scell + r.W)), c11 = xw_cls(r, celldesc(scell + r.W + 1));
        if ((c00 | c01 | c10 | c11) < 2) corner_combo = c00 | (c01 << 1) | (c10 << 2) | (c11 << 3);
    }
    uint32_t wa[3], wb[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
        if (cc < nc) { wa[cc] = *(const uint32_t*)(tabs[0] + pl[cc] * r.OW); wb[cc] = *(const uint32_t*)(tabs[1] + pl[cc] * r.OW); }
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
        if (cc < nc) {
            uint32_t word = xw_prmt(wa[cc], wb[cc], sel);
            if (corner) {
                const uint32_t v = corner_combo >= 0 ? x.cornerWB[((corner_combo * r.n_sr + q) * r.n_sc + s_idx) * 3 + c0 + pl[cc]]
                                                     : xw_corner_px(r, x, celldesc, scell, c0 + pl[cc], y0, dx);
                word = (word & ~(0xffu << sh)) | (v << sh);
            }
            fb[w0 + po[cc]] = word;
        }
}

// descriptor of a grid code in env e: grid code -> icon + 1
XW_HD uint32_t xw_cell_desc(const XwDev& d, int e, int code) {
    if (code == XW_CELL_EMPTY) return 0;
    if (code == XW_CELL_BLOCK) return (uint32_t)d.brick_icon + 1;
    if (code == XW_CELL_AGENT) return (uint32_t)d.agent_icon + 1;
    return (uint32_t)d.goal_icon[(size_t)(code - XW_CELL_GOAL0) * d.n + e] + 1;
}
#define XW_CODE_SLOTS (XW_CELL_GOAL0 + XW_MAX_GOALS + 5)  // 16 descriptors per env


// Dynamic shared memory layout (bytes), all sections 16-byte aligned:
//   brick phase table (TMA bulk load, once) | G frame buffers | plan | yb | G cell arrays | mbarrier
#define XW_CELL_STRIDE (XW_MAX_DIM * XW_MAX_DIM + 2 * XW_MAX_DIM)  // bytes: the map + the (never drawn) row below it
// one cell buffer: codes | descriptors per code | cell of the agent / of goal g (XW_STAGE_SLOTS bytes)
#define XW_CELLBUF_BYTES (XW_CELL_STRIDE + XW_CODE_SLOTS * 4 + 16)
struct XwRenderSmem { int hot, fb, plan, yb, cellinfo, pair, cell, bar, total; };
XW_HD int xw_align16(int v) { return (v + 15) & ~15; }
// nbuf = frame buffers (and cell buffers) in the CTA: G in the single-buffer kernel, 2*G in the pipelined one
XW_HD XwRenderSmem xw_render_smem(const XwRender& r, int nbuf) {
    XwRenderSmem s;
    int o = 0;
    s.hot = o; o += xw_align16(r.FB);
    s.fb = o; o += nbuf * xw_align16(r.FB);
    s.plan = o; o += r.n_plan * 16;
    s.yb = o; o += xw_align16(r.OH * 4);
    s.cellinfo = o; o += XW_MAX_DIM * XW_MAX_DIM * 4;
    s.pair = o; o += xw_align16(xw_pair_hot_bytes(r));
    s.cell = o; o += nbuf * XW_CELLBUF_BYTES;
    s.bar = o; o += 16;
    s.total = o;
    return s;
}


// ---- sparse painter (k_render_sp) -----------------------------------------------------------------
// The frame buffer is pre-filled with white; only the words a non-white cell touches are then written.
// Three kinds of work, each with its own uniform code path (geometry: xw_build_paint_tables):
//   brick slots    (brick cell, word column of the cell): the rows of the cell's band in the three planes.
//                  A word touches at most two cell columns lo | hi, so with only bricks and white around
//                  there are three cases -- brick|brick (or a word inside one cell), brick|white,
//                  white|brick -- and each has a precomputed, exact (straddling byte included) class
//                  table in shared memory, stored column-major so that the rows of a word column are
//                  consecutive words: one LDS + one STS per frame word, no merging.
//   special slots  (agent / goal cell, word column): the general two-source merge (PRMT) with the
//                  straddling byte from the pair tables or the edge taps; the cell's own words come from
//                  the L2-resident phase atlas.  Computed into registers while the previous frame of
//                  the group drains, stored after the pre-fill.
//   straddling-row words: every word of every straddling row (dense, they are few), from the row pair
//                  tables; likewise precomputed into registers.
// Ownership: a word that touches a special cell belongs to that cell's slot (the left one if both are
// special); otherwise to the brick on its left, else to the brick on its right.
#define XW_SP_ROWS_MAX 12  // most rows of a band without its straddling row; a special slot keeps 3 x ROWS words in
                           // registers, ROWS = 8 or 12 (template parameter)
#define XW_SP_RR 2     // straddling-row words a lane of warp 0 precomputes per env; the rest are written in place
struct XwPaintCtx {    // shared-memory copies on the device
    const XwU4* cellgeo;
    const uint32_t* wcol;
    const uint8_t* wshare;
    const uint32_t* ctab;   // class tables, column-major: full [WR][3][OH] | brick|white [ns][3][OH] | white|brick [ns][3][OH]
    const uint8_t* sr_ty;   // [n_sr] cell row above straddling row q
    const uint8_t* sr_dy;   // [n_sr] its output row
    const uint32_t* sc_a;   // [n_sc] horizontal weights xa0 | xa1 << 16 of straddling column s
    const uint32_t* row2_hot;  // the white and brick entries of rowT2: [2 descs][2 cls][n_sr][WR][3]
};
// words per class-table column: 3 planes x OH rows, made odd so that the columns of one band (same rows, different
// word columns) sit on different shared-memory banks
XW_HD int xw_ctab_stride(const XwRender& r) { return (3 * r.OH) | 1; }
XW_HD size_t xw_ctab_words(const XwRender& r) {  // (a multiple of 4: the tables are one TMA bulk load)
    return ((size_t)(r.WR + 2 * r.ns) * xw_ctab_stride(r) + 3) & ~(size_t)3;
}

// cells of a class-table entry: cell column `lo` and the other one hold a brick or nothing
struct XwClassCells {
    int W, lo;
    uint32_t dlo, dhi;
    XW_HD uint32_t operator()(int cell) const { return cell % W == lo ? dlo : dhi; }
};
// word (k, p, dy) of class table `variant` (0 brick|brick, 1 brick|white, 2 white|brick); only the rows
// whose taps stay inside one cell row are used
XW_HD uint32_t xw_ctab_word(const XwRender& r, uint32_t wk, int variant, int k, int p, int dy) {
    XwClassCells cc;
    cc.W = r.W; cc.lo = (int)(wk & 15);
    cc.dlo = variant == 2 ? 0u : (uint32_t)r.brick_icon + 1;
    cc.dhi = variant == 1 ? 0u : (uint32_t)r.brick_icon + 1;
    uint32_t w = 0;
    for (int i = 0; i < 4; ++i) w |= (uint32_t)xw_exact_px(r, cc, p, dy, 4 * k + i) << (8 * i);
    return w;
}

// The three planes of word k of straddling row q, in two steps like the special slots:
//   MODE 0 (issue):  the two row-pair-table words per plane -> wa, wb and, for a corner next to a special cell,
//                    the corner taps -> t; loads only (from L2 when a special cell is involved), no use;
//   MODE 1 (finish): merge, insert the corner byte, store.  (Special above special: everything here, exactly.)
template <int WR_T, int MODE>
XW_HD void xw_sp_rword(const XwRender& r, const XwComposeCtx& x, const XwPaintCtx& g, const XwCells& cells, int q, int k,
                       uint32_t wa[3], uint32_t wb[3], uint32_t t[4], uint32_t* fb) {
    const int WR = WR_T ? WR_T : r.WR;
    const int OH = WR_T ? 4 * WR_T : r.OH;  // (a compile-time row stride is only used for square frames)
    const uint32_t wk = g.wcol[k];
    const int lo = wk & 15, hi = (wk >> 4) & 15;
    const uint32_t sel = (wk >> 8) & 0xffffu;
    const int top = g.sr_ty[q] * r.W;
    const int kTa = cells.code[top + lo], kBa = cells.code[top + r.W + lo], kTb = cells.code[top + hi], kBb = cells.code[top + r.W + hi];
    const int cTa = kTa < 2 ? kTa : 2, cBa = kBa < 2 ? kBa : 2, cTb = kTb < 2 ? kTb : 2, cBb = kBb < 2 ? kBb : 2;  // = xw_cls
    const bool exact = (cTa == 2 && cBa == 2) || (cTb == 2 && cBb == 2);
    const bool corner = ((wk >> 24) & 1) != 0, wb_corner = (cTa | cTb | cBa | cBb) < 2;
    const int sh = (int)((wk >> 25) & 3) * 8, sidx = (int)(wk >> 27);
    if (MODE == 0) {
        if (exact) return;
        // per cell column: (top cell, class of the bottom cell) or (bottom cell, class of the top cell)
        const uint32_t rs = (uint32_t)(r.n_sr * WR * 3), off = (uint32_t)((q * WR + k) * 3);  // words per (desc, class); this word
        const uint32_t *tA, *tB;
        if (cBa < 2) tA = (cTa < 2 ? g.row2_hot + cTa * 2 * rs : r.rowT2 + cells.icon[kTa] * 2 * rs) + cBa * rs;
        else tA = r.rowB2 + cells.icon[kBa] * 2 * rs + cTa * rs;
        if (cBb < 2) tB = (cTb < 2 ? g.row2_hot + cTb * 2 * rs : r.rowT2 + cells.icon[kTb] * 2 * rs) + cBb * rs;
        else tB = r.rowB2 + cells.icon[kBb] * 2 * rs + cTb * rs;
#pragma unroll
        for (int p = 0; p < 3; ++p) { wa[p] = tA[off + p]; wb[p] = tB[off + p]; }
        if (corner && !wb_corner) {
            // taps (63,63) of the top-left cell, (63,0) top-right, (0,63) bottom-left, (0,0) bottom-right
            t[0] = r.cornerP[cells.icon[kTa] * 4 + 0]; t[1] = r.cornerP[cells.icon[kTb] * 4 + 1];
            t[2] = r.cornerP[cells.icon[kBa] * 4 + 2]; t[3] = r.cornerP[cells.icon[kBb] * 4 + 3];
        }
        return;
    }
    uint32_t out[3];
    if (exact) {
        // a special cell above a special cell: separable rule (U(top) + V(bottom) + 2) >> 2 on packed u16 pairs
        const size_t per_desc = (size_t)r.n_sr * 2 * 3 * r.OW;
        const uint32_t dTa = cells.icon[kTa], dBa = cells.icon[kBa], dTb = cells.icon[kTb], dBb = cells.icon[kBb];
        for (int p = 0; p < 3; ++p) {
            const uint16_t* u0 = r.uv + ((size_t)q * 2 * 3 + p) * r.OW + 4 * k;
            const uint16_t* v0 = u0 + (size_t)3 * r.OW;
            const XwU2 uA = *(const XwU2*)(u0 + dTa * per_desc), uB = *(const XwU2*)(u0 + dTb * per_desc);
            const XwU2 vA = *(const XwU2*)(v0 + dBa * per_desc), vB = *(const XwU2*)(v0 + dBb * per_desc);
            const uint32_t a_lo = ((uA.x + vA.x + 0x00020002u) >> 2) & 0x00ff00ffu, a_hi = ((uA.y + vA.y + 0x00020002u) >> 2) & 0x00ff00ffu;
            const uint32_t b_lo = ((uB.x + vB.x + 0x00020002u) >> 2) & 0x00ff00ffu, b_hi = ((uB.y + vB.y + 0x00020002u) >> 2) & 0x00ff00ffu;
            out[p] = xw_prmt(xw_prmt(a_lo, a_hi, 0x6420), xw_prmt(b_lo, b_hi, 0x6420), sel);
            if (corner) out[p] = (out[p] & ~(0xffu << sh)) | ((uint32_t)xw_corner_px(r, x, cells, top + lo, p, g.sr_dy[q], r.sc[sidx]) << sh);
        }
    } else {
#pragma unroll
        for (int p = 0; p < 3; ++p) out[p] = xw_prmt(wa[p], wb[p], sel);
        if (corner) {  // corner byte: four cells
            uint32_t v[3];
            if (wb_corner) {
#pragma unroll
                for (int p = 0; p < 3; ++p) v[p] = x.cornerWB[(((cTa | (cTb << 1) | (cBa << 2) | (cBb << 3)) * r.n_sr + q) * r.n_sc + sidx) * 3 + p];
            } else {
                const uint32_t a = g.sc_a[sidx], b = x.yb[g.sr_dy[q]];
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    v[p] = xw_resize_px((t[0] >> (8 * p)) & 255, (t[1] >> (8 * p)) & 255, (t[2] >> (8 * p)) & 255, (t[3] >> (8 * p)) & 255, a & 0xffff, a >> 16,
                                        b & 0xffff, b >> 16);
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) out[p] = (out[p] & ~(0xffu << sh)) | (v[p] << sh);
        }
    }
    uint32_t* dst = fb + g.sr_dy[q] * WR + k;
#pragma unroll
    for (int p = 0; p < 3; ++p) dst[p * OH * WR] = out[p];
}

// Word column wc of special cell `cell` in plane p, in two steps so that the L2 latency of the first is never waited for:
//   MODE 0 (issue):  the cell's own phase-table words of the band rows -> m[j], the straddling byte (if any) of
//                    four rows per word -> pb[j / 4]; loads only, no use;
//   MODE 1 (finish): merge with the other cell column of the word (PRMT), insert the straddling byte, store
//                    to the frame buffer.  Recomputes the (shared-memory) geometry instead of carrying it.
// False: nothing to do (the cell has fewer word columns, or the word belongs to the special cell on its left).
template <int WR_T, int XW_SP_ROWS, int MODE>
XW_HD bool xw_sp_special(const XwRender& r, const XwComposeCtx& x, const XwPaintCtx& g, const XwCells& cells, int cell, int wc, int p,
                         uint32_t m[XW_SP_ROWS], uint32_t pb[XW_SP_ROWS / 4], uint32_t* fb) {
    const int WR = WR_T ? WR_T : r.WR;
    const int OH = WR_T ? 4 * WR_T : r.OH;  // (a compile-time row stride is only used for square frames)
    const XwU4 cg = g.cellgeo[cell];
    if (wc >= (int)((cg.x >> 24) & 7)) return false;
    const int k = (int)(cg.y & 0xff) + wc, ty = (int)((cg.y >> 8) & 0xff);
    const uint32_t wk = g.wcol[k];
    const int lo = wk & 15, hi = (wk >> 4) & 15;
    const int row = cell - (int)((cg.z >> 8) & 0xff);  // ty * W
    const int kA = cells.code[row + lo], kB = cells.code[row + hi];
    if (row + lo != cell && kA >= XW_CELL_AGENT) return false;
    const uint32_t dA = cells.icon[kA], dB = cells.icon[kB];
    const int nrows = (int)((cg.x >> 16) & 0xff), y0 = (int)(cg.z & 0xff);
    const int PW = OH * WR;
    const bool meA = row + lo == cell;
    const bool has_sc = ((wk >> 24) & 1) != 0;
    const int sh = (int)((wk >> 25) & 3) * 8, sidx = (int)(wk >> 27);
    const int cL = kA < 2 ? kA : 2, cR = kB < 2 ? kB : 2;
    const bool exact = has_sc && cL == 2 && cR == 2;
    if (MODE == 0) {
        // this cell's own phase table (row-major, L2); the straddling byte from the pair table of the special
        // cell and the class of the other one (special | special: exact, in the finish step)
        const uint32_t dM = meA ? dA : dB;
        const uint8_t* tab = (const uint8_t*)g.ctab;  // (no straddling byte: any readable address)
        if (has_sc && !exact) {
            const size_t cs = xw_colpair_stride(r);
            tab = (cR < 2 ? r.colL + ((size_t)dA * 2 + cR) * cs : r.colR + ((size_t)dB * 2 + cL) * cs) + ((size_t)(sidx * 3 + p) * r.H + ty) * r.RB;
        }
        // (unconditional loads: rows past the band are inside the padded tables and never used)
        if (r.TC == nullptr) {  // big maps: no cell-major table (xw_create), the phase atlas row by row
            const uint32_t* pT = (const uint32_t*)(r.T + (size_t)(dM - 1) * r.FB) + p * PW + y0 * WR + k;
#pragma unroll
            for (int j = 0; j < XW_SP_ROWS; ++j) m[j] = pT[j * WR];
        } else {  // 16-byte loads (the blocks are 32- or 48-byte aligned)
            const uint32_t* pM = r.TC + ((((size_t)(dM - 1) * (r.H * r.W) + cell) * 3 + p) * r.nwc + wc) * XW_SP_ROWS;
#pragma unroll
            for (int j4 = 0; j4 < XW_SP_ROWS; j4 += 4) {
                const XwU4 q4 = *(const XwU4*)(pM + j4);
                m[j4] = q4.x; m[j4 + 1] = q4.y; m[j4 + 2] = q4.z; m[j4 + 3] = q4.w;
            }
        }
#pragma unroll
        for (int j4 = 0; j4 < XW_SP_ROWS; j4 += 4) pb[j4 / 4] = *(const uint32_t*)(tab + j4);
        return true;
    }
    // the other cell column of the word: the brick class table (column-major, shared memory; white = brick |
    // 0xffffffff) unless it is special too
    const bool spO = lo != hi && (meA ? kB : kA) >= XW_CELL_AGENT;
    const uint32_t dO = meA ? dB : dA;
    const uint32_t sel = ((wk >> 8) & 0xffffu) ^ (meA ? 0u : 0x4444u);  // PRMT(mine, other)
    const uint32_t* pO = spO ? (const uint32_t*)(r.T + (size_t)(dO - 1) * r.FB) + p * PW + y0 * WR + k
                             : g.ctab + (size_t)k * ((3 * OH) | 1) + p * OH + y0;
    const int rsO = spO ? WR : 1;
    const uint32_t wmask = xw_prmt(0u, dO == 0 ? 0xffffffffu : 0u, sel);
    const uint32_t keep = has_sc ? ~(0xffu << sh) : 0xffffffffu;
    const int bsh = has_sc ? sh : 0;
    uint32_t* dst = fb + p * PW + y0 * WR + k;
    uint32_t o[XW_SP_ROWS];
#pragma unroll
    for (int j = 0; j < XW_SP_ROWS; ++j) o[j] = pO[j * rsO];  // (unconditional, as in the issue step)
#pragma unroll
    for (int j = 0; j < XW_SP_ROWS; ++j)
        if (j < nrows) {
            const uint32_t byte = has_sc ? (pb[j / 4] >> (8 * (j & 3))) & 0xffu : 0u;
            dst[j * WR] = ((xw_prmt(m[j], o[j], sel) | wmask) & keep) | (byte << bsh);
        }
    if (exact) {  // special | special across the straddling column: from the edge taps
        const int dx = r.sc[sidx], a0 = r.taps.xa0[dx], a1 = r.taps.xa1[dx];
        const uint16_t* eL = r.ecol + ((((size_t)dA * 2 + 0) * 3 + p) * r.H + ty) * r.RB;
        const uint16_t* eR = r.ecol + ((((size_t)dB * 2 + 1) * 3 + p) * r.H + ty) * r.RB;
        for (int j = 0; j < nrows; ++j) {
            const uint32_t tl = eL[j], tr = eR[j], b = x.yb[y0 + j];
            const uint32_t v = xw_resize_px(tl & 255, tr & 255, tl >> 8, tr >> 8, a0, a1, b & 0xffff, b >> 16);
            dst[j * WR] = (dst[j * WR] & ~(0xffu << sh)) | (v << sh);
        }
    }
    return true;
}

// Word column wc of brick cell `cell`: one class-table column per plane, copied to the frame.
template <int WR_T, int XW_SP_ROWS>
XW_HD void xw_sp_brick_slot(const XwRender& r, const XwPaintCtx& g, const XwCells& cells, int cell, int wc, uint32_t* fb) {
    const int WR = WR_T ? WR_T : r.WR;
    const int OH = WR_T ? 4 * WR_T : r.OH;  // (a compile-time row stride is only used for square frames)
    const XwU4 cg = g.cellgeo[cell];
    if (wc >= (int)((cg.x >> 24) & 7)) return;
    const int k = (int)(cg.y & 0xff) + wc;
    const uint32_t wk = g.wcol[k];
    const int lo = wk & 15, hi = (wk >> 4) & 15;
    int col = k;  // class-table column: brick | brick
    uint32_t selm = 0;  // merge mode: != 0 = the brick|brick word with the white cell's bytes set (PRMT selector)
    if (lo != hi) {
        const int row = cell - (int)((cg.z >> 8) & 0xff);
        const bool left = row + lo == cell;
        const int ko = cells.code[row + (left ? hi : lo)];  // the other cell of the word
        if (ko >= XW_CELL_AGENT || (!left && ko == XW_CELL_BLOCK)) return;  // a special cell's, or the left brick's word
        if (ko == XW_CELL_EMPTY) {
            const int ws = g.wshare[k];
            if (ws != 0xff) col = r.WR + (left ? 0 : r.ns) + ws;
            else selm = ((wk >> 8) & 0xffffu) ^ (left ? 0u : 0x4444u);  // (ctab_merge only: every shared column has a table otherwise)
        }
    }
    const int nrows = (int)((cg.x >> 16) & 0xff), y0 = (int)(cg.z & 0xff);
    const uint32_t* src = g.ctab + (size_t)col * ((3 * OH) | 1) + y0;
    uint32_t* dst = fb + y0 * WR + k;
    // (rows past the band are loaded -- the tables are followed by other shared memory -- and not stored)
    if (selm == 0) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = src[p * OH + j];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nrows) dst[p * OH * WR + j * WR] = v[j];
            if (XW_SP_ROWS > 8 && nrows > 8) {
                uint32_t u[XW_SP_ROWS > 8 ? XW_SP_ROWS - 8 : 1];
#pragma unroll
                for (int j = 8; j < XW_SP_ROWS; ++j) u[j - 8] = src[p * OH + j];
#pragma unroll
                for (int j = 8; j < XW_SP_ROWS; ++j)
                    if (j < nrows) dst[p * OH * WR + j * WR] = u[j - 8];
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = src[p * OH + j];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nrows) dst[p * OH * WR + j * WR] = xw_prmt(v[j], 0xffffffffu, selm);
            if (XW_SP_ROWS > 8 && nrows > 8) {
                uint32_t u[XW_SP_ROWS > 8 ? XW_SP_ROWS - 8 : 1];
#pragma unroll
                for (int j = 8; j < XW_SP_ROWS; ++j) u[j - 8] = src[p * OH + j];
#pragma unroll
                for (int j = 8; j < XW_SP_ROWS; ++j)
                    if (j < nrows) dst[p * OH * WR + j * WR] = xw_prmt(u[j - 8], 0xffffffffu, selm);
            }
        }
    }
}
// brick slot s -> (list index, word column); exact for s < 4096 (xw_build_paint_tables: nwc <= 4)
XW_HD void xw_sp_decode(const XwRender& r, int s, int* i, int* wc) {
    *i = (int)(((uint32_t)s * (uint32_t)r.slot_magic) >> 16);
    *wc = s - *i * r.nwc;
}

// TC[icon][cell][p][wc][j] = T[icon][p][y0(cell) + j][kfirst(cell) + wc]  (0 past the frame / the cell's columns)
XW_HD uint32_t xw_tc_word(const XwRender& r, size_t i) {
    const int j = (int)(i % r.tc_rows); i /= r.tc_rows;
    const int wc = (int)(i % r.nwc); i /= r.nwc;
    const int p = (int)(i % 3); i /= 3;
    const int cell = (int)(i % (r.H * r.W)); const size_t icon = i / (r.H * r.W);
    const XwU4 cg = r.cellgeo[cell];
    const int k = (int)(cg.y & 0xff) + wc, dy = (int)(cg.z & 0xff) + j;
    if (wc >= (int)((cg.x >> 24) & 7) || dy >= r.OH) return 0;
    return ((const uint32_t*)(r.T + icon * r.FB))[(p * r.OH + dy) * r.WR + k];
}
// rowT2 / rowB2 [desc][cls][q][k][p] <- rowT / rowB [desc][cls][q][p][OW bytes]
XW_HD uint32_t xw_row2_word(const XwRender& r, const uint8_t* src, size_t i) {
    const int p = (int)(i % 3); i /= 3;
    const int k = (int)(i % r.WR); i /= r.WR;
    const int q = (int)(i % r.n_sr); i /= r.n_sr;  // i = desc * 2 + cls
    return *(const uint32_t*)(src + i * xw_rowpair_stride(r) + (size_t)(q * 3 + p) * r.OW + 4 * k);
}
#define XW_SP_LIST_BYTES (XW_MAX_DIM * XW_MAX_DIM + 16)   // brick cell list + its length (u32 at the end)
XW_HD XwRenderSpSmem xw_render_sp_smem(const XwRender& r, int G) {
    XwRenderSpSmem s;
    int o = 0;
    s.ctab = o; o += xw_align16((int)xw_ctab_words(r) * 4);
    s.fb = o; o += G * xw_align16(r.FB);
    s.cellgeo = o; o += r.H * r.W * 16;
    s.wcol = o; o += xw_align16(r.WR * 4);
    s.wshare = o; o += xw_align16(r.WR);
    s.srq = o; o += xw_align16(2 * (r.n_sr + 1));
    s.sca = o; o += xw_align16(4 * (r.n_sc + 1));
    s.yb = o; o += xw_align16(r.OH * 4);
    s.pair = o; o += xw_align16(4 * r.n_sr * r.WR * 3 * 4) + xw_align16(16 * r.n_sr * r.n_sc * 3);
    s.cell = o; o += 2 * G * (XW_CELLBUF_BYTES + XW_SP_LIST_BYTES);
    s.bar = o; o += 16 + 16 * G;
    s.total = o;
    return s;
}

// Merge mode of the class tables (XwRender::ctab_merge)?  Yes when the full tables leave room for two frame buffers only and the
// reduced ones for a third (15x15 -> 128x128: 195 KB at two groups, 245 KB at three with 86 KB of tables, 214 KB with 55 KB);
// XW_RENDER_CTAB_MERGE=0|1 overrides.  r: OH, OW, WR, FB, H, W, n_sr, n_sc set.
inline bool xw_want_ctab_merge(const XwRender& r, const XwRenderTables& t, int max_smem) {
    if (const char* e = getenv("XW_RENDER_CTAB_MERGE")) return atoi(e) != 0 && t.ns_strad < t.ns;
    if (t.ns_strad >= t.ns) return false;
    XwRender a = r, b = r;
    a.ns = t.ns; b.ns = t.ns_strad;
    return xw_render_sp_smem(a, 3).total > max_smem && xw_render_sp_smem(b, 3).total <= max_smem;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------ kernels
__global__ void k_build_phase_atlas(XwRender r) {
    const size_t total = (size_t)r.n_icons * r.FB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int icon = (int)(i / r.FB), rem = (int)(i % r.FB);
        int c = rem / (r.OH * r.OW), p = rem % (r.OH * r.OW);
        ((uint8_t*)r.T)[i] = xw_phase_px(r, icon, c, p / r.OW, p % r.OW);
    }
}
__global__ void k_build_edge_tables(XwRender r) {
    const size_t n_ecol = (size_t)(r.n_icons + 1) * 2 * 3 * r.H * r.RB, n_uv = (size_t)(r.n_icons + 1) * r.n_sr * 2 * 3 * r.OW;
    const size_t n_cor = (size_t)(r.n_icons + 1) * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ecol + n_uv + n_cor; i += (size_t)gridDim.x * blockDim.x) {
        if (i >= n_ecol + n_uv) {
            const size_t j = i - n_ecol - n_uv;
            ((uint32_t*)r.corner)[j] = xw_corner_entry(r, (uint32_t)(j / 3), (int)(j % 3));
        } else if (i < n_ecol) {
            size_t j = i;
            const int row = (int)(j % r.RB); j /= r.RB;
            const int band = (int)(j % r.H); j /= r.H;
            const int c = (int)(j % 3); j /= 3;
            ((uint16_t*)r.ecol)[i] = xw_ecol_entry(r, (uint32_t)(j / 2), (int)(j % 2), c, band, row);
        } else {
            size_t j = i - n_ecol;
            const int dx = (int)(j % r.OW); j /= r.OW;
            const int c = (int)(j % 3); j /= 3;
            const int role = (int)(j % 2); j /= 2;
            ((uint16_t*)r.uv)[i - n_ecol] = xw_uv_entry(r, (uint32_t)(j / r.n_sr), (int)(j % r.n_sr), role, c, dx);
        }
    }
}

__global__ void k_build_pair_tables(XwRender r) {
    const size_t cs = xw_colpair_stride(r), rs = xw_rowpair_stride(r);
    const size_t n_col = (size_t)(r.n_icons + 1) * 2 * cs, n_row = (size_t)(r.n_icons + 1) * 2 * rs;
    const size_t n_cwb = (size_t)16 * r.n_sr * r.n_sc * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * (n_col + n_row) + n_cwb; i += (size_t)gridDim.x * blockDim.x) {
        if (i >= 2 * (n_col + n_row)) {
            size_t j = i - 2 * (n_col + n_row);
            const int c = (int)(j % 3); j /= 3;
            const int sidx = (int)(j % r.n_sc); j /= r.n_sc;
            ((uint8_t*)r.cornerWB)[i - 2 * (n_col + n_row)] = xw_cornerwb_entry(r, (int)(j / r.n_sr), (int)(j % r.n_sr), sidx, c);
        } else if (i < 2 * n_col) {
            const int role = i >= n_col;
            size_t j = i - role * n_col;
            const int row = (int)(j % r.RB); j /= r.RB;
            const int band = (int)(j % r.H); j /= r.H;
            const int c = (int)(j % 3); j /= 3;
            const int s = (int)(j % r.n_sc); j /= r.n_sc;
            ((uint8_t*)(role ? r.colR : r.colL))[i - role * n_col] = xw_colpair_entry(r, role, (uint32_t)(j / 2), (int)(j % 2), s, c, band, row);
        } else {
            const size_t i2 = i - 2 * n_col;
            const int role = i2 >= n_row;
            size_t j = i2 - role * n_row;
            const int dx = (int)(j % r.OW); j /= r.OW;
            const int c = (int)(j % 3); j /= 3;
            const int q = (int)(j % r.n_sr); j /= r.n_sr;
            ((uint8_t*)(role ? r.rowB : r.rowT))[i2 - role * n_row] = xw_rowpair_entry(r, role, (uint32_t)(j / 2), (int)(j % 2), q, c, dx);
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copies (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// frames are written once and not read by this kernel: evict-first in L2, so that the streaming
// 1.4 GB of frame lines does not push the render tables out
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {  // SASS: LDGSTS
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent CTAs.  Group g of CTA b renders envs (b*G + g) + i * gridDim.x*G; the frame of env e goes
// to frames + e*env_stride.  Each group owns TWO frame buffers and runs a two-stage pipeline with one
// group barrier per env:
//   warp 0 of the group is also the loader: while the group composes env i into buffer i&1 it stores the
//   (register-prefetched) cells of env i+1 into the other cell buffer, prefetches env i+2, waits until
//   the TMA store of env i-1 has drained buffer (i+1)&1 and stages env i+1's agent / goal table words
//   into it (LDGSTS, no registers).  After the barrier thread 0 hands buffer i&1 to the TMA engine and
//   everybody starts env i+1 at once.
template <int WR_T, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, 1)
k_render(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int G = r.G, GT = r.GT;
    const XwRenderSmem L = xw_render_smem(r, 2 * G);
    uint8_t* hot = smem + L.hot;
    XwU4* s_plan = (XwU4*)(smem + L.plan);
    uint32_t* s_yb = (uint32_t*)(smem + L.yb);
    uint64_t* bar = (uint64_t*)(smem + L.bar);
    const int tid = threadIdx.x, nt = blockDim.x;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {  // stage the brick table: one TMA bulk load per CTA
        mbar_expect_tx(bar, (uint32_t)r.FB);
        tma_load_1d(hot, r.T + (size_t)r.brick_icon * r.FB, (uint32_t)r.FB, bar);
    }
    {  // plan + row weights + cell geometry -> shared memory
        for (int i = tid; i < r.n_plan; i += nt) s_plan[i] = r.plan[i];
        for (int i = tid; i < r.H * r.W; i += nt) ((uint32_t*)(smem + L.cellinfo))[i] = r.cellinfo[i];
        for (int i = tid; i < r.OH; i += nt) s_yb[i] = (uint32_t)(uint16_t)r.taps.ya0[i] | ((uint32_t)(uint16_t)r.taps.ya1[i] << 16);
        for (int i = tid; i < G * 2 * XW_CELLBUF_BYTES / 4; i += nt) ((uint32_t*)(smem + L.cell))[i] = 0;
        // white / brick entries of the pair tables: [desc in {white, brick}][cls][...]
        const int cs2 = (int)(2 * xw_colpair_stride(r)), rs2 = (int)(2 * xw_rowpair_stride(r));
        uint8_t* pc = smem + L.pair;
        uint8_t* pr = pc + 2 * cs2;
        for (int i = tid; i < 2 * cs2; i += nt) pc[i] = r.colL[(size_t)(i < cs2 ? 0 : r.brick_icon + 1) * cs2 + (i < cs2 ? i : i - cs2)];
        for (int i = tid; i < 2 * rs2; i += nt) pr[i] = r.rowT[(size_t)(i < rs2 ? 0 : r.brick_icon + 1) * rs2 + (i < rs2 ? i : i - rs2)];
        for (int i = tid; i < 16 * r.n_sr * r.n_sc * 3; i += nt) pr[2 * rs2 + i] = r.cornerWB[i];
    }
    mbar_wait(bar, 0);
    __syncthreads();

    const int g = tid / GT, gt = tid - g * GT;
    if (g >= G) return;  // spare warps (G*GT < blockDim.x)
    const uint32_t* s_cellinfo = (const uint32_t*)(smem + L.cellinfo);
    uint8_t* cellbuf = smem + L.cell + (size_t)g * 2 * XW_CELLBUF_BYTES;
    uint8_t* fbbuf = smem + L.fb + (size_t)g * 2 * xw_align16(r.FB);
    XwComposeCtx x;
    x.hot = hot; x.yb = s_yb;
    x.colL_hot = smem + L.pair; x.rowT_hot = smem + L.pair + 4 * xw_colpair_stride(r);
    x.cornerWB = x.rowT_hot + 4 * xw_rowpair_stride(r);
    const int bar_id = 1 + g;
    const int n_plan = r.n_plan;
    const int gstride = gridDim.x * G;
    const bool loader = gt < 32;  // warp 0 of the group
    const int lane = gt & 31;
    const int row_words = d.CS >> 2;  // <= 64: two words per loader lane
    const int WR = WR_T ? WR_T : r.WR;

    // loader registers: the grid row + goal icons of the env after next
    uint32_t nq0 = 0, nq1 = 0, ni = 0;
    auto load_cells = [&](int e) {
        if (e < d.n) {
            const uint32_t* row = (const uint32_t*)(d.grid + (size_t)e * d.CS);
            if (lane < row_words) nq0 = row[lane];
            if (lane + 32 < row_words) nq1 = row[lane + 32];
            if (lane < d.G) ni = (uint32_t)d.goal_icon[(size_t)lane * d.n + e] + 1;
        }
    };
    auto store_cells = [&](int buf) {  // registers -> cell buffer `buf`: codes, descriptors, where the specials are
        uint8_t* code = cellbuf + buf * XW_CELLBUF_BYTES;
        uint32_t* icon = (uint32_t*)(code + XW_CELL_STRIDE);
        uint8_t* special = (uint8_t*)(icon + XW_CODE_SLOTS);
        if (lane == XW_CELL_BLOCK) icon[lane] = (uint32_t)d.brick_icon + 1;
        if (lane == XW_CELL_AGENT) icon[lane] = (uint32_t)d.agent_icon + 1;
        if (lane < d.G) icon[XW_CELL_GOAL0 + lane] = ni;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t q = h ? nq1 : nq0;
            const int wi = lane + 32 * h;
            if (wi < row_words) {
                ((uint32_t*)code)[wi] = q;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t c = (q >> (8 * b)) & 0xff;
                    if (c >= XW_CELL_AGENT) special[c - XW_CELL_AGENT] = (uint8_t)(4 * wi + b);
                }
            }
        }
        __syncwarp();
    };
    auto stage = [&](int buf) {  // LDGSTS the special cells' table words of cell buffer `buf` into frame buffer `buf`
        XwCells cells;
        cells.code = cellbuf + buf * XW_CELLBUF_BYTES;
        cells.icon = (const uint32_t*)(cells.code + XW_CELL_STRIDE);
        const uint8_t* special = (const uint8_t*)(cells.icon + XW_CODE_SLOTS);
        uint32_t* fb = (uint32_t*)(fbbuf + (size_t)buf * xw_align16(r.FB));
        for (int i = lane; i < (1 + d.G) * XW_STAGE_COLS; i += 32) {
            const int slot = i / XW_STAGE_COLS, col = i - slot * XW_STAGE_COLS, c = col / 3, wc = col - 3 * c;
            uint32_t w0; int nrows; const uint32_t* src;
            if (xw_stage_column(r, cells, s_cellinfo, special[slot], c, wc, &w0, &nrows, &src))
                for (int j = 0; j < nrows; ++j) cp_async_4(fb + w0 + j * WR, src + j * WR);
        }
    };

    int env = blockIdx.x * G + g;
    if (env >= d.n) return;
    if (loader) {  // prologue: env 0 of this group goes in directly, env 1 into the registers
        load_cells(env);
        store_cells(0);
        stage(0);
        load_cells(env + gstride);
        cp_async_wait_all();
    }
    group_bar(bar_id, GT);
    for (int it = 0; env < d.n; env += gstride, ++it) {
        const int buf = it & 1;
        XwCells cells;
        cells.code = cellbuf + buf * XW_CELLBUF_BYTES;
        cells.icon = (const uint32_t*)(cells.code + XW_CELL_STRIDE);
        uint32_t* fb = (uint32_t*)(fbbuf + (size_t)buf * xw_align16(r.FB));
        const bool have_next = env + gstride < d.n;
        if (loader && have_next) {
            store_cells(buf ^ 1);             // env i+1's cells (every warp is past env i-1, which used this buffer)
            load_cells(env + 2 * gstride);    // env i+2 -> registers
        }
        bool staged = !(loader && have_next);
        for (int i = gt; i < n_plan; i += GT) {
            xw_compose_item<WR_T, (NT_MAX > 640)>(r, x, s_plan[i], cells, fb);
            if (!staged) {  // after the first bundle: by now the TMA store of env i-1 has usually drained
                if (lane == 0) tma_wait_read<0>();
                __syncwarp();
                stage(buf ^ 1);
                staged = true;
            }
        }
        if (!staged) { if (lane == 0) tma_wait_read<0>(); __syncwarp(); stage(buf ^ 1); }
        if (loader) cp_async_wait_all();
        fence_async_smem();  // generic-proxy writes -> visible to the async (TMA) proxy
        group_bar(bar_id, GT);
        if (gt == 0) {
            tma_store_1d(frames + (size_t)env * env_stride, fb, (uint32_t)r.FB);
            tma_commit();
        }
    }
    if (gt == 0) tma_wait_all<0>();
}

// Single-buffer variant of k_render (XW_RENDER_MODE=sb): r.G groups with ONE frame buffer each and
// three group barriers per env (cells visible / staged words visible / compose done).  More groups and
// warps per SM, more synchronisation per env.  Same compose code, same shared-memory layout function
// (called with (G+1)/2 buffer pairs).
template <int WR_T, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, 1)
k_render_sb(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int G = r.G, GT = r.GT;
    const XwRenderSmem L = xw_render_smem(r, G);
    uint8_t* hot = smem + L.hot;
    XwU4* s_plan = (XwU4*)(smem + L.plan);
    uint32_t* s_yb = (uint32_t*)(smem + L.yb);
    uint64_t* bar = (uint64_t*)(smem + L.bar);
    const int tid = threadIdx.x, nt = blockDim.x;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {  // stage the brick table: one TMA bulk load per CTA
        mbar_expect_tx(bar, (uint32_t)r.FB);
        tma_load_1d(hot, r.T + (size_t)r.brick_icon * r.FB, (uint32_t)r.FB, bar);
    }
    {  // plan + row weights + pair-table heads -> shared memory
        for (int i = tid; i < r.n_plan; i += nt) s_plan[i] = r.plan[i];
        for (int i = tid; i < r.H * r.W; i += nt) ((uint32_t*)(smem + L.cellinfo))[i] = r.cellinfo[i];
        for (int i = tid; i < r.OH; i += nt) s_yb[i] = (uint32_t)(uint16_t)r.taps.ya0[i] | ((uint32_t)(uint16_t)r.taps.ya1[i] << 16);
        const int cs2 = (int)(2 * xw_colpair_stride(r)), rs2 = (int)(2 * xw_rowpair_stride(r));
        uint8_t* pc = smem + L.pair;
        uint8_t* pr = pc + 2 * cs2;
        for (int i = tid; i < 2 * cs2; i += nt) pc[i] = r.colL[(size_t)(i < cs2 ? 0 : r.brick_icon + 1) * cs2 + (i < cs2 ? i : i - cs2)];
        for (int i = tid; i < 2 * rs2; i += nt) pr[i] = r.rowT[(size_t)(i < rs2 ? 0 : r.brick_icon + 1) * rs2 + (i < rs2 ? i : i - rs2)];
        for (int i = tid; i < 16 * r.n_sr * r.n_sc * 3; i += nt) pr[2 * rs2 + i] = r.cornerWB[i];
    }
    mbar_wait(bar, 0);
    __syncthreads();

    const int g = tid / GT, gt = tid - g * GT;
    if (g >= G) return;  // spare warps (G*GT < blockDim.x)
    uint8_t* s_code = smem + L.cell + g * XW_CELLBUF_BYTES;
    uint32_t* s_icon = (uint32_t*)(s_code + XW_CELL_STRIDE);
    uint8_t* s_special = (uint8_t*)(s_icon + XW_CODE_SLOTS);  // [XW_STAGE_SLOTS] cell of the agent / goal g
    const uint32_t* s_cellinfo = (const uint32_t*)(smem + L.cellinfo);
    uint32_t* fb = (uint32_t*)(smem + L.fb + (size_t)g * xw_align16(r.FB));
    XwComposeCtx x;
    x.hot = hot; x.yb = s_yb;
    x.colL_hot = smem + L.pair; x.rowT_hot = smem + L.pair + 4 * xw_colpair_stride(r);
    x.cornerWB = x.rowT_hot + 4 * xw_rowpair_stride(r);
    XwCells cells;
    cells.code = s_code; cells.icon = s_icon;
    const int bar_id = 1 + g;
    const int n_plan = r.n_plan;
    const int gstride = gridDim.x * G;
    int env = blockIdx.x * G + g;

    // register prefetch of the env's grid row (CS bytes = CS/4 words, CS/4 <= 64 <= GT) and goal icons
    const int row_words = d.CS >> 2;
    uint32_t nq = 0, ni = 0;
    if (env < d.n) {
        if (gt < row_words) nq = ((const uint32_t*)(d.grid + (size_t)env * d.CS))[gt];
        if (gt < d.G) ni = (uint32_t)d.goal_icon[(size_t)gt * d.n + env] + 1;
    }
    for (int i = gt; i < XW_CELL_STRIDE / 4; i += GT) ((uint32_t*)s_code)[i] = 0;
    if (gt < XW_CODE_SLOTS) s_icon[gt] = gt == XW_CELL_BLOCK ? (uint32_t)d.brick_icon + 1 : gt == XW_CELL_AGENT ? (uint32_t)d.agent_icon + 1 : 0;
    for (; env < d.n; env += gstride) {
        if (gt < row_words) {  // (every warp passed the barrier after the last compose)
            ((uint32_t*)s_code)[gt] = nq;
#pragma unroll
            for (int b = 0; b < 4; ++b) {  // where the agent and the goals are
                const uint32_t code = (nq >> (8 * b)) & 0xff;
                if (code >= XW_CELL_AGENT) s_special[code - XW_CELL_AGENT] = (uint8_t)(4 * gt + b);
            }
        }
        if (gt < d.G) s_icon[XW_CELL_GOAL0 + gt] = ni;
        if (gt == 0) tma_wait_read<0>();  // the TMA store of this group's previous frame has drained fb
        group_bar(bar_id, GT);
        // stage the agent's and goals' table words into the frame buffer (LDGSTS, no registers): one
        // thread per (cell, plane, word column), one 4-byte copy per row
        for (int i = gt; i < ((XW_DBG(2)) ? 0 : (1 + d.G) * XW_STAGE_COLS); i += GT) {
            const int slot = i / XW_STAGE_COLS, col = i - slot * XW_STAGE_COLS, c = col / 3, wc = col - 3 * c;
            uint32_t w0; int nrows; const uint32_t* src;
            if (xw_stage_column(r, cells, s_cellinfo, s_special[slot], c, wc, &w0, &nrows, &src)) {
                const int WR = WR_T ? WR_T : r.WR;
                for (int j = 0; j < nrows; ++j) cp_async_4(fb + w0 + j * WR, src + j * WR);
            }
        }
        {  // prefetch the next env's cells while this one is composed
            const int en = env + gstride;
            if (en < d.n) {
                if (gt < row_words) nq = ((const uint32_t*)(d.grid + (size_t)en * d.CS))[gt];
                if (gt < d.G) ni = (uint32_t)d.goal_icon[(size_t)gt * d.n + en] + 1;
            }
        }
        // the straddling-row items read only tables: they run while the staging copies are in flight
        for (int i = gt; i < r.n_plan1; i += GT) xw_compose_item<WR_T, (NT_MAX > 640)>(r, x, s_plan[i], cells, fb);
        cp_async_wait_all();
        group_bar(bar_id, GT);
        for (int i = r.n_plan1 + gt; i < ((XW_DBG(1)) ? 0 : n_plan); i += GT) xw_compose_item<WR_T, (NT_MAX > 640)>(r, x, s_plan[i], cells, fb);
        fence_async_smem();  // generic-proxy writes -> visible to the async (TMA) proxy
        group_bar(bar_id, GT);
        if (gt == 0 && !(XW_DBG(4))) {
            tma_store_1d(frames + (size_t)env * env_stride, fb, (uint32_t)r.FB);
            tma_commit();
        }
    }
    if (gt == 0) tma_wait_all<0>();
}


__global__ void k_build_row2_tables(XwRender r) {
    const size_t total = (size_t)(r.n_icons + 1) * 2 * r.n_sr * r.WR * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        ((uint32_t*)r.rowT2)[i] = xw_row2_word(r, r.rowT, i);
        ((uint32_t*)r.rowB2)[i] = xw_row2_word(r, r.rowB, i);
    }
}
__global__ void k_build_cell_tables(XwRender r) {
    const size_t total = (size_t)r.n_icons * r.H * r.W * 3 * r.nwc * r.tc_rows;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        ((uint32_t*)r.TC)[i] = xw_tc_word(r, i);
}
__global__ void k_build_class_tables(XwRender r) {
    const size_t per = (size_t)xw_ctab_stride(r), total = xw_ctab_words(r);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)(r.n_icons + 1) * 4; i += (size_t)gridDim.x * blockDim.x)
        ((uint32_t*)r.cornerP)[i] = xw_cornerP_entry(r, (uint32_t)(i / 4), (int)(i % 4));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(i / per), rem = (int)(i % per), p = rem / r.OH, dy = rem % r.OH;
        if (p >= 3 || col >= r.WR + 2 * r.ns) { ((uint32_t*)r.ctab)[i] = 0; continue; }  // pad words
        int variant = 0, k = col;
        if (col >= r.WR) {  // brick|white then white|brick, shared word columns only
            variant = col - r.WR < r.ns ? 1 : 2;
            const int si = col - r.WR - (variant == 2 ? r.ns : 0);
            for (k = 0; k < r.WR; ++k) if (r.wshare[k] == si) break;
        }
        ((uint32_t*)r.ctab)[i] = xw_ctab_word(r, r.wcol[k], variant, k, p, dy);
    }
}

// Sparse painter: r.G warp groups of >= 2 warps, one frame buffer each.  Per env, in a group:
//   warp 0    stores the env's cell codes, lists its brick cells, signals warp 1 (mbarrier), prefetches the next
//             env's cells and computes the special slots into registers;
//   warp 1    computes the straddling-row words into registers; then its last lane -- the thread that issued
//             the group's previous TMA store -- waits until that store has read the frame buffer, and the
//             buffer is made white again (vector stores by warp 1, or one TMA bulk load: r.sp_fill).
//             All L2 latency of an env is in these two PRE phases, which run side by side and overlap the
//             drain of the previous frame;
//   everybody then stores the precomputed words and paints the brick slots (POST: shared memory only), and
//   the TMA thread hands the frame to the TMA engine (one bulk store, evict-first in L2).
template <int WR_T, int NT_MAX, int XW_SP_ROWS, bool DISJ = false, bool LIST = false>
__global__ void __launch_bounds__(NT_MAX, 1)
k_render_sp(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    // Programmatic dependent launch: once every CTA of the painter is resident (it is one CTA per SM on fewer CTAs than SMs), the
    // step's auto-reset kernel -- launched behind it in the same stream with programmatic stream serialization -- may start on
    // the SMs the painter left free (xw_engine.cu step_xworld).  Without a dependent the instruction does nothing.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int G = r.G, GT = r.GT;
    const XwRenderSpSmem& L = r.sp;
    uint32_t* s_yb = (uint32_t*)(smem + L.yb);
    uint64_t* bar = (uint64_t*)(smem + L.bar);
    const int tid = threadIdx.x, nt = blockDim.x;

    if (tid == 0) {
        mbar_init(bar, 1);
        for (int i = 0; i < 2 * G; ++i) mbar_init(bar + 2 + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ctab_bytes = (uint32_t)xw_ctab_words(r) * 4;
    if (tid == 0) {  // the class tables: one TMA bulk load per CTA
        mbar_expect_tx(bar, ctab_bytes);
        tma_load_1d(smem + L.ctab, r.ctab, ctab_bytes, bar);
    }
    {  // geometry, row weights, row pair tables of white / brick -> shared memory
        for (int i = tid; i < r.H * r.W; i += nt) ((XwU4*)(smem + L.cellgeo))[i] = r.cellgeo[i];
        for (int i = tid; i < r.WR; i += nt) { ((uint32_t*)(smem + L.wcol))[i] = r.wcol[i]; (smem + L.wshare)[i] = r.wshare[i]; }
        for (int i = tid; i < r.n_sr; i += nt) { (smem + L.srq)[i] = r.sr_ty[i]; (smem + L.srq)[r.n_sr + i] = (uint8_t)r.sr[i]; }
        for (int i = tid; i < r.n_sc; i += nt) ((uint32_t*)(smem + L.sca))[i] = (uint32_t)(uint16_t)r.taps.xa0[r.sc[i]] | ((uint32_t)(uint16_t)r.taps.xa1[r.sc[i]] << 16);
        for (int i = tid; i < r.OH; i += nt) s_yb[i] = (uint32_t)(uint16_t)r.taps.ya0[i] | ((uint32_t)(uint16_t)r.taps.ya1[i] << 16);
        const int rs2 = 2 * r.n_sr * r.WR * 3;  // words of one descriptor in rowT2
        uint32_t* pr = (uint32_t*)(smem + L.pair);
        for (int i = tid; i < 2 * rs2; i += nt) pr[i] = r.rowT2[(size_t)(i < rs2 ? 0 : r.brick_icon + 1) * rs2 + (i < rs2 ? i : i - rs2)];
        uint8_t* pcw = smem + L.pair + xw_align16(2 * rs2 * 4);
        for (int i = tid; i < 16 * r.n_sr * r.n_sc * 3; i += nt) pcw[i] = r.cornerWB[i];
        for (int i = tid; i < 2 * G * (XW_CELLBUF_BYTES + XW_SP_LIST_BYTES) / 4; i += nt) ((uint32_t*)(smem + L.cell))[i] = 0;
    }
    mbar_wait(bar, 0);
    __syncthreads();

    const int g = tid / GT, gt = tid - g * GT;
    if (g >= G) return;  // spare warps (G*GT < blockDim.x)
    // two cell buffers per group: codes | descriptor per code | cell of the agent / goal g | brick list | count
    uint8_t* cellbuf = smem + L.cell + g * 2 * (XW_CELLBUF_BYTES + XW_SP_LIST_BYTES);
    uint32_t* fb = (uint32_t*)(smem + L.fb + (size_t)g * xw_align16(r.FB));
    uint64_t *fillbar = bar + 2 + g, *cellsbar = bar + 2 + G + g;
    XwComposeCtx x;
    x.hot = nullptr; x.yb = s_yb; x.colL_hot = nullptr;
    x.rowT_hot = nullptr;
    x.cornerWB = smem + L.pair + xw_align16(4 * r.n_sr * r.WR * 3 * 4);
    XwPaintCtx pg;
    pg.row2_hot = (const uint32_t*)(smem + L.pair);
    pg.cellgeo = (const XwU4*)(smem + L.cellgeo); pg.wcol = (const uint32_t*)(smem + L.wcol); pg.wshare = smem + L.wshare;
    pg.ctab = (const uint32_t*)(smem + L.ctab); pg.sr_ty = smem + L.srq; pg.sr_dy = smem + L.srq + r.n_sr;
    pg.sc_a = (const uint32_t*)(smem + L.sca);
    const int bar_id = 1 + g;
    const int gstride = gridDim.x * G;
    const int lane = gt & 31;
    const bool w0 = gt < 32, w1 = gt >= 32 && gt < 64;  // the two PRE warps of the group
    const bool tma_thread = gt == 63, tma_fill = r.sp_fill != 0;
    const int row_words = d.CS >> 2, HW = d.H * d.W;
    const int WR = WR_T ? WR_T : r.WR;
    const int n_sslots = (1 + d.G) * r.nwc, n_rw = r.n_sr * WR;
    // `env` counts the kernel's work items: the envs themselves, or (LIST) positions of the env list
#define XW_N_ITEMS (LIST ? n_list : d.n)
#define XW_ENV_ID(i) (LIST ? r.env_list[(i)] : (i))
    const int n_list = LIST ? *r.env_count : 0;
    int env = blockIdx.x * G + g;
    if (env >= XW_N_ITEMS) return;

    // register prefetch of an env's grid row (<= 64 words: two per lane of warp 0) and goal icons
    uint32_t nq0 = 0, nq1 = 0, ni = 0;
    auto load_cells = [&](int e) {
        const uint32_t* rowp = (const uint32_t*)(d.grid + (size_t)e * d.CS);
        if (lane < row_words) nq0 = rowp[lane];
        if (lane + 32 < row_words) nq1 = rowp[lane + 32];
        if (lane < d.G) ni = (uint32_t)d.goal_icon[(size_t)lane * d.n + e];  // (+1 at the use: no wait here)
    };
    auto cells_of = [&](int b) {
        XwCells c;
        c.code = cellbuf + b * (XW_CELLBUF_BYTES + XW_SP_LIST_BYTES);
        c.icon = (const uint32_t*)(c.code + XW_CELL_STRIDE);
        return c;
    };
    // (warp 0) registers -> cell buffer b: codes, descriptors, where the agent and the goals are, brick list
    auto store_cells = [&](int b) {
        uint8_t* s_code = cellbuf + b * (XW_CELLBUF_BYTES + XW_SP_LIST_BYTES);
        uint32_t* s_icon = (uint32_t*)(s_code + XW_CELL_STRIDE);
        uint8_t* s_special = (uint8_t*)(s_icon + XW_CODE_SLOTS);
        uint8_t* s_list = s_code + XW_CELLBUF_BYTES;
        if (lane == XW_CELL_BLOCK) s_icon[lane] = (uint32_t)d.brick_icon + 1;
        if (lane == XW_CELL_AGENT) s_icon[lane] = (uint32_t)d.agent_icon + 1;
        if (lane < d.G) s_icon[XW_CELL_GOAL0 + lane] = ni + 1;
        if (lane < XW_STAGE_SLOTS) s_special[lane] = 0;  // (absent: the cell's code will not match, see issue())
        __syncwarp();
        uint32_t m0 = 0, m1 = 0;
        if (lane < row_words) {
            ((uint32_t*)s_code)[lane] = nq0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t c = (nq0 >> (8 * q)) & 0xff;
                if (c == XW_CELL_BLOCK && 4 * lane + q < HW) m0 |= 1u << q;
                if (c >= XW_CELL_AGENT && !(XW_DBG(8))) s_special[c - XW_CELL_AGENT] = (uint8_t)(4 * lane + q);
            }
        }
        if (lane + 32 < row_words) {
            ((uint32_t*)s_code)[lane + 32] = nq1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t c = (nq1 >> (8 * q)) & 0xff;
                if (c == XW_CELL_BLOCK && 4 * (lane + 32) + q < HW) m1 |= 1u << q;
                if (c >= XW_CELL_AGENT && !(XW_DBG(8))) s_special[c - XW_CELL_AGENT] = (uint8_t)(4 * (lane + 32) + q);
            }
        }
        // brick list: exclusive prefix of the per-lane counts (4-bit masks -> one ballot per bit); the second half
        // of the row (maps of more than 128 cells) follows the first
        int total = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && row_words <= 32) break;
            const uint32_t mk = h ? m1 : m0;
            const uint32_t b0 = __ballot_sync(0xffffffffu, mk & 1u), b1 = __ballot_sync(0xffffffffu, mk & 2u);
            const uint32_t b2 = __ballot_sync(0xffffffffu, mk & 4u), b3 = __ballot_sync(0xffffffffu, mk & 8u);
            const uint32_t below = (1u << lane) - 1u;
            int pos = total + xw_popc(b0 & below) + xw_popc(b1 & below) + xw_popc(b2 & below) + xw_popc(b3 & below);
#pragma unroll
            for (int q = 0; q < 4; ++q) if (mk & (1u << q)) s_list[pos++] = (uint8_t)(4 * (lane + 32 * h) + q);
            total += xw_popc(b0) + xw_popc(b1) + xw_popc(b2) + xw_popc(b3);
        }
        if (lane == 0) *(volatile uint32_t*)(s_list + XW_MAX_DIM * XW_MAX_DIM) = (uint32_t)total;
        __syncwarp();
    };
    // Work that reads L2 (special slots, straddling-row words) is ISSUED for the next env at the end of an env's
    // POST and FINISHED in the next env's POST: the loads have a whole env time to land and nothing on the path
    // barrier -> paint -> barrier -> TMA store ever waits for L2 (under this kernel's write load an L2 round
    // trip is thousands of cycles).  Every thread of the group takes one special slot-plane (threads < 3 * n_sslots)
    // and one straddling-row word (threads < n_rw; more words than threads: the rest in place), so that both warps
    // carry the same load: a warp issues an instruction every ~7 cycles, and the instruction count of the
    // busiest warp is what an env costs.
    // role index of a thread: with three or more warps, warps 1 and 2 are swapped so that warp 1 -- which waits for
    // the drain and refills the buffer -- gets straddling-row words only and the special slot-planes go to warps 0, 2
    const int vt = (GT >= 96 && gt >= 32 && gt < 96) ? (gt < 64 ? gt + 32 : gt - 32) : gt;
    const int n_sl3 = 3 * n_sslots;                       // special slot-planes: vt = (ord * 3 + p) * nwc + wc
    const bool s_lane = vt < n_sl3;
    const int s_ord = vt / (3 * r.nwc), s_p = (vt - s_ord * 3 * r.nwc) / r.nwc, s_wc = vt - (s_ord * 3 + s_p) * r.nwc;
    const int r_idx = GT - 1 - vt;                        // straddling-row word (three planes): from the last thread down
    const bool r_lane = r_idx < n_rw && !(XW_DBG(16));
    const int r_q = r_lane ? r_idx / WR : 0, r_k = r_lane ? r_idx - r_q * WR : 0;
    // carried raw loads: special slot-plane (m[ROWS], pb[ROWS/4]) and straddling-row word (wa[3], wb[3], t[4]).
    // DISJ: no thread has both roles (3 * n_sslots + n_rw <= GT), so the two share one set of registers.
    constexpr int N_SC = XW_SP_ROWS + XW_SP_ROWS / 4;
    uint32_t carry[N_SC < 10 ? 10 : N_SC], carry2[DISJ ? 1 : 10];
    uint32_t *sm = carry, *spb = carry + XW_SP_ROWS;
    uint32_t *rwa = DISJ ? carry : carry2, *rwb = rwa + 3, *rt = rwa + 6;
    int s_cell = 0;
    bool s_have = false;
    auto issue = [&](int b) {
        const XwCells cn = cells_of(b);
        s_have = false;
        if (s_lane) {
            s_cell = ((const uint8_t*)(cn.icon + XW_CODE_SLOTS))[s_ord];
            // (every cell index 0..255 is a valid cell of a 16x16 map: presence = the cell really holds this code)
            if (cn.code[s_cell] == XW_CELL_AGENT + s_ord) s_have = xw_sp_special<WR_T, XW_SP_ROWS, 0>(r, x, pg, cn, s_cell, s_wc, s_p, sm, spb, nullptr);
        }
        if (r_lane) xw_sp_rword<WR_T, 0>(r, x, pg, cn, r_q, r_k, rwa, rwb, rt, nullptr);
    };
    // ---- prologue: env 0 of the group
    if (w0) {
        load_cells(XW_ENV_ID(env));
        store_cells(0);
        if (env + gstride < XW_N_ITEMS) load_cells(XW_ENV_ID(env + gstride));
    }
    group_bar(bar_id, GT);
    issue(0);

#if defined(XW_SP_PROF)  // phase clocks of lane 0 (a special-slot lane) and lane 63 (the TMA thread) of group 0, CTA 0
    const bool prof = r.prof != nullptr && blockIdx.x == 0 && g == 0 && (gt == 0 || gt == 63);
    unsigned int t_prev = prof ? (unsigned int)clock() : 0u;
#define XW_PROF(i) do { if (prof) { const unsigned int t_now = (unsigned int)clock(); atomicAdd(r.prof + (gt == 0 ? 0 : 8) + (i), t_now - t_prev); t_prev = t_now; } } while (0)
#else
#define XW_PROF(i) do { } while (0)
#endif
    for (uint32_t it = 0; env < XW_N_ITEMS; env += gstride, ++it) {
        const int cur = it & 1;
        const bool have_next = env + gstride < XW_N_ITEMS;
        const XwCells cells = cells_of(cur);
        XW_PROF(7);  // (loop overhead, TMA store issue)
        if (w0) {
            if (have_next) {
                store_cells(cur ^ 1);  // next env's cells (every warp is past the POST of env it-1, the buffer's last reader)
            }
        } else if (w1 && !(XW_DBG(2))) {  // the previous frame has left the buffer -> white again
            if (lane == 31) {
                tma_wait_read<0>();
                XW_PROF(0);  // lane 63: drain wait
                if (tma_fill) {
                    mbar_expect_tx(fillbar, (uint32_t)r.FB);
                    tma_load_1d(fb, r.white, (uint32_t)r.FB, fillbar);
                }
            }
            if (!tma_fill) {
                __syncwarp();
                const int4 ones = make_int4(-1, -1, -1, -1);
#pragma unroll 8
                for (int i = lane; i < r.FB / 16; i += 32) ((int4*)fb)[i] = ones;
            }
        }
        XW_PROF(1);  // lane 0: next env's cells -> shared memory; lane 63: fill
        if (tma_fill && !(XW_DBG(2))) mbar_wait(fillbar, it & 1);
        group_bar(bar_id, GT);  // buffer white, next env's cells visible
        XW_PROF(2);  // barrier A wait
        // ---- POST of this env: finish what was issued an env ago, paint the bricks, issue for the next env
        if (!(XW_DBG(1))) {
            const uint8_t* s_list = cells.code + XW_CELLBUF_BYTES;
            if (r_lane) xw_sp_rword<WR_T, 1>(r, x, pg, cells, r_q, r_k, rwa, rwb, rt, fb);
            if (s_have && !(XW_DBG(32))) xw_sp_special<WR_T, XW_SP_ROWS, 1>(r, x, pg, cells, s_cell, s_wc, s_p, sm, spb, fb);
            XW_PROF(3);  // finish (lane 0: special slot, lane 63: straddling-row word)
            if (!(XW_DBG(16)))
                for (int idx = GT + gt; idx < n_rw; idx += GT) {  // (more straddling-row words than lanes: in place)
                    uint32_t ta[3], tb[3], tt[4];
                    xw_sp_rword<WR_T, 0>(r, x, pg, cells, idx / WR, idx % WR, ta, tb, tt, nullptr);
                    xw_sp_rword<WR_T, 1>(r, x, pg, cells, idx / WR, idx % WR, ta, tb, tt, fb);
                }
            const int n_slots = (int)*(volatile const uint32_t*)(s_list + XW_MAX_DIM * XW_MAX_DIM) * r.nwc;
            // (the last warp first: warp 0 has the cells of the next env to store, so the odd round goes elsewhere)
            for (int s = GT - 32 - (gt & ~31) + lane; s < n_slots; s += GT) {
                int i, wc;
                xw_sp_decode(r, s, &i, &wc);
                xw_sp_brick_slot<WR_T, XW_SP_ROWS>(r, pg, cells, s_list[i], wc, fb);
            }
        }
        XW_PROF(4);  // brick slots
        fence_async_smem();  // generic-proxy writes -> visible to the async (TMA) proxy
        // (after the fence, which would otherwise wait for them: L2 loads for the next env and the env after it)
        if (have_next && !(XW_DBG(65))) issue(cur ^ 1);
        if (w0 && env + 2 * gstride < XW_N_ITEMS) load_cells(XW_ENV_ID(env + 2 * gstride));
        XW_PROF(5);  // issue for the next env
        group_bar(bar_id, GT);
        XW_PROF(6);  // barrier C wait
        if (tma_thread && !(XW_DBG(4))) {
            tma_store_1d(frames + (size_t)XW_ENV_ID(env) * env_stride, fb, (uint32_t)r.FB);
            tma_commit();
        }
    }
#undef XW_PROF
#undef XW_N_ITEMS
#undef XW_ENV_ID
    if (tma_thread) tma_wait_all<0>();
}

// General fallback (any frame size): one thread per output byte, straight to global memory.
__device__ __forceinline__ uint8_t xw_generic_byte(const XwDev& d, const XwRender& r, int e, int rem) {
    const XwTaps& t = r.taps;
    const int c = rem / (r.OH * r.OW), p = rem % (r.OH * r.OW), dy = p / r.OW, dx = p % r.OW;
    const int sx0 = t.xofs[dx], sy0 = t.yofs[dy];
    const int sx1 = t.xa1[dx] ? sx0 + 1 : sx0, sy1 = t.ya1[dy] ? sy0 + 1 : sy0;
    const uint8_t* g = d.grid + (size_t)e * d.CS;
    uint32_t dd[4];
    const int cy[2] = {sy0 >> 6, sy1 >> 6}, cx[2] = {sx0 >> 6, sx1 >> 6};
    for (int q = 0; q < 4; ++q) dd[q] = xw_cell_desc(d, e, g[cy[q >> 1] * r.W + cx[q & 1]]);
    if (dd[0] == dd[1] && dd[0] == dd[2] && dd[0] == dd[3]) return dd[0] == 0 ? 255 : r.T[(size_t)(dd[0] - 1) * r.FB + rem];
    const int sy[2] = {sy0, sy1}, sx[2] = {sx0, sx1};
    int px[4];
    for (int q = 0; q < 4; ++q) px[q] = xw_canvas_tap(r, dd[q], sy[q >> 1], sx[q & 1], c);
    return xw_resize_px(px[0], px[1], px[2], px[3], t.xa0[dx], t.xa1[dx], t.ya0[dy], t.ya1[dy]);
}
__global__ void k_render_generic(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    const size_t total = (size_t)d.n * r.FB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / r.FB), rem = (int)(i % r.FB);
        frames[(size_t)e * env_stride + rem] = xw_generic_byte(d, r, e, rem);
    }
}
// --context > 1 (GameSimulator::shift_context, simulator.cpp:51-60): slots 1..K-1 -> 0..K-2, for the envs that were stepped
// since their last render (flag 1).  An env that sat the step out (flag 0, XW_ACTION_NONE / not in a reset mask) keeps its
// history; an env that was reset (flag 2) starts with a zero-filled context (init_screen, simulator.cpp:110-113).
// The flags are consumed here.
__global__ void k_shift_context(uint8_t* frames, int n, int K, int FB, uint8_t* flag) {
    const int words = FB / 16;  // FB % 16 == 0 checked by the host
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int f = flag ? flag[e] : 1;
        __syncthreads();
        if (threadIdx.x == 0 && flag) flag[e] = 0;
        if (f == 0) continue;
        int4* base = (int4*)(frames + (size_t)e * K * FB);
        if (f == 2) {
            for (int i = threadIdx.x; i < (K - 1) * words; i += blockDim.x) base[i] = make_int4(0, 0, 0, 0);
            continue;
        }
        for (int s = 0; s + 1 < K; ++s) {
            for (int i = threadIdx.x; i < words; i += blockDim.x) base[(size_t)s * words + i] = base[(size_t)(s + 1) * words + i];
            __syncthreads();
        }
    }
}
#endif  // __CUDACC__
