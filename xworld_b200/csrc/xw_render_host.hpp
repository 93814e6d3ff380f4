// xw_render_host.hpp -- host-side construction of the renderer's lookup tables.
// cv::resize(INTER_LINEAR, 8U) coefficient tables as OpenCV 3.2 computes them (imgproc/resize.cpp:
// scale = 1/(dst/src) in double, fx in float, cvFloor, 11-bit rounded weights), plus the plan the
// compositing kernel executes: one "band-column" item per (cell row, output word column).
#pragma once
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

// One compositor work item = one 4-pixel word column over a run of rows, in nc consecutive colour
// planes.  Items are typed so that a warp executes one code path:
//   M1  rows of a band (cell row), word inside ONE cell:            word = table(A)
//   M2  rows of a band, word spanning two cells:                    word = PRMT(table(A), table(B))
//   M3  like M2, and one byte lies on a straddling column (its two tap columns are in different
//       cells): that byte is evaluated exactly from the cells' edge-tap tables
//   R   one word of a straddling row (tap rows in different cell rows): separable U(top)+V(bottom)
//       rule
//   RC  an R word that also holds a corner (straddling row x column): that byte depends on four cells
//       (kept in bundles of their own: there are n_sr*n_sc of them per plane)
// The M items of a band exclude the band's straddling last row, so every frame word has one writer.
enum { XW_ITEM_M1 = 0, XW_ITEM_M2 = 1, XW_ITEM_M3 = 2, XW_ITEM_R = 3, XW_ITEM_RC = 4, XW_ITEM_TYPES = 5 };

// Packed plan entry (one LDS.128):
//   x = cellA | cellB << 8 | sel << 16     cell indices ty*W + tx owning the word's first / last byte
//                                          (top cell row for R), PRMT selector merging word(A), word(B)
//   y = woff | nrows << 16 | type << 24 | c0 << 27 | nc << 29
//                                          word offset of the first row in the FRAME (plane c0), rows,
//                                          first plane, number of planes (0 = padding slot)
//   z = y0 | dx << 8 | scell << 16 | sbyte << 24    first row; M3 / R-with-corner: the straddling column,
//                                          its left cell (the right one is scell + 1), its byte in the word
//   w = q | k << 8 | band << 16 | rot << 24 | sidx << 26 | corner << 31     R: index of the straddling row (M3: of the
//                                          straddling column), word column,
//                                          corner flag; M3: band = cell row (edge tables are stored per
//                                          band); rot: the item visits planes (rot + i) % 3, i = 0,1,2
// The plan is a sequence of 32-slot bundles of one type; bundle b belongs to warp b % n_warps of a
// warp group, and the host orders bundles so that the warps of a group finish together (LPT).
struct alignas(16) XwU4 { uint32_t x, y, z, w; };
struct alignas(8) XwU2 { uint32_t x, y; };

struct XwPlanItem { int cellA, cellB, sel, woff, nrows, type, y0, k, sbyte, scell, dx, q, corner, band, sidx; };

struct XwRenderTables {
    int H = 0, W = 0, OH = 0, OW = 0, WR = 0, FB = 0;
    bool fast_ok = false;  // the shared-memory compositor applies (else: generic kernel)
    std::vector<int16_t> xofs, xa0, xa1, yofs, ya0, ya1, sc, sr;
    std::vector<uint32_t> cellinfo;      // [H*W] woff | nrows << 16 | nwords << 24 | first_shared << 26: the frame
                                         // words cell (ty,tx) owns in the M items (staging of special cells)
    std::vector<int16_t> band_y0;        // [H] first output row of each cell row's band (-1: none)
    int RB = 0;                          // rows per band in the edge table, a multiple of 4
    std::vector<XwPlanItem> items;       // per-plane items (planner input)
    std::vector<XwU4> plan;              // bundled plan for n_warps warps per group
    // sparse painter (k_render_sp): static geometry per cell / per word column, see xw_build_paint_tables
    bool sp_ok = false;
    int nwc = 0;                         // most word columns one cell column touches
    std::vector<XwU4> cellgeo;           // [H*W]
    std::vector<uint32_t> wcol;          // [WR]
    std::vector<uint8_t> wshare;         // [WR] index of word column k among the columns shared by two cell columns (0xff: not shared)
    // merge mode (XwRender::ctab_merge): variant class tables only for the shared columns that hold a straddling pixel
    std::vector<uint8_t> wshare_strad;   // [WR] index among those (0xff: not shared, or shared without a straddling pixel)
    int ns_strad = 0;
    int ns = 0;                          // shared word columns
    std::vector<uint8_t> sr_ty;          // [n_sr] cell row above straddling row q
    int max_band_rows = 0;               // most rows of a band without its straddling row
    int n_warps = 0;
    int n_plan1 = 0;                     // plan[0..n_plan1): phase 1 (R / RC bundles), the rest: phase 2
    int bank_conflicts = 0;              // lane pairs of a bundle left on one bank (0 = conflict-free plan)
    double makespan = 0, total_cost = 0; // planner's cost model (instructions per env): slowest warp, sum
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
    bool ok = (OW % 4 == 0) && (t.FB % 16 == 0) && OH <= 255 && OW <= 252 && OH * (OW / 4) <= 65535 && t.sr.size() <= 255 &&
              t.sc.size() <= 31;
    // a straddling row is the last row of its band (yofs is monotone); two in one band (upscaling)
    // would break the plan
    for (size_t q = 0; q < t.sr.size() && ok; ++q) {
        int dy = t.sr[q];
        if (dy + 1 < OH && (t.yofs[dy + 1] >> 6) == (t.yofs[dy] >> 6)) ok = false;
    }
    if (!ok) return t;
    t.WR = OW / 4;
    t.band_y0.assign(H, -1);
    for (int ty = 0; ty < H && ok; ++ty) {
        int y0 = -1, y1 = -1;  // rows owned by cell row ty
        for (int dy = 0; dy < OH; ++dy)
            if ((t.yofs[dy] >> 6) == ty) { if (y0 < 0) y0 = dy; y1 = dy + 1; }
        if (y0 < 0) continue;
        t.band_y0[ty] = (int16_t)y0;
        if (y1 - y0 > t.RB) t.RB = y1 - y0;
        const bool srow = is_sr[y1 - 1] != 0;
        int q = 0;
        if (srow) while (t.sr[q] != y1 - 1) ++q;
        for (int k = 0; k < t.WR && ok; ++k) {
            XwPlanItem it;
            memset(&it, 0, sizeof it);
            int tx[4];
            for (int i = 0; i < 4; ++i) tx[i] = t.xofs[4 * k + i] >> 6;
            const int A = tx[0], B = tx[3];
            int n_sc = 0;
            for (int i = 0; i < 4; ++i) {
                if (tx[i] == A) it.sel |= i << (4 * i);
                else if (tx[i] == B) it.sel |= (4 + i) << (4 * i);
                else ok = false;  // a 4-pixel word spans 3 cells: cells narrower than 2 px
                if (is_sc[4 * k + i]) {
                    ++n_sc; it.sbyte = i; it.scell = ty * W + tx[i]; it.dx = 4 * k + i;
                    for (size_t si = 0; si < t.sc.size(); ++si) if (t.sc[si] == 4 * k + i) it.sidx = (int)si;
                }
            }
            if (n_sc > 1) ok = false;  // two straddling columns in one word: cells narrower than 4 px
            it.cellA = ty * W + A; it.cellB = ty * W + B;
            it.woff = y0 * t.WR + k;
            it.nrows = y1 - y0 - (srow ? 1 : 0);
            it.y0 = y0; it.k = k; it.band = ty;
            it.type = n_sc ? XW_ITEM_M3 : (A != B ? XW_ITEM_M2 : XW_ITEM_M1);
            if (it.nrows > 0) t.items.push_back(it);
            if (srow) {
                XwPlanItem r = it;
                r.woff = (y1 - 1) * t.WR + k;
                r.nrows = 1; r.y0 = y1 - 1; r.q = q;
                r.type = XW_ITEM_R; r.corner = n_sc ? 1 : 0;  // (XW_ITEM_RC: corner words in bundles of their own -- measured slower)
                t.items.push_back(r);
            }
        }
    }
    if (3 * OH * t.WR > 65535) ok = false;  // woff is 16 bits
    t.RB = (t.RB + 3) / 4 * 4;
    t.cellinfo.assign((size_t)H * W, 0);
    for (const XwPlanItem& it : t.items) {  // per cell: first word column, number of columns, rows of the M items
        if (it.type >= XW_ITEM_R) continue;
        const int cells[2] = {it.cellA, it.cellB};
        for (int q = 0; q < 2; ++q) {
            uint32_t& ci = t.cellinfo[cells[q]];
            const int kfirst = ci ? (int)((ci & 0xffffu) - (uint32_t)it.y0 * t.WR) : it.k;
            const int nwords = ci ? (int)((ci >> 24) & 3) : 0;
            const int kf = it.k < kfirst ? it.k : kfirst, kl = (it.k > kfirst + nwords - 1 ? it.k : kfirst + nwords - 1);
            if (kl - kf + 1 > 3) ok = false;  // XW_STAGE_COLS assumes a cell spans at most 3 word columns
            ci = (uint32_t)(it.y0 * t.WR + kf) | ((uint32_t)it.nrows << 16) | ((uint32_t)(kl - kf + 1) << 24);
        }
    }
    for (const XwPlanItem& it : t.items)  // first column shared with the left neighbour?
        if (it.type < XW_ITEM_R && it.cellA != it.cellB) t.cellinfo[it.cellB] |= 1u << 26;
    t.fast_ok = ok;
    return t;
}

// ---- sparse painter geometry -------------------------------------------------------------------
// k_render_sp pre-fills a frame with white and then paints only the words a non-white cell touches.
// A frame word (band = cell row ty, word column k) touches at most two cell columns lo <= hi (its four
// bytes' tap-0 cells, plus the tap-1 cell of a straddling byte); its OWNER is cell (ty,lo) unless that
// cell is white, then (ty,hi).  The word of a straddling row between bands ty and ty+1 is owned by the
// owner of word (ty,k) if there is one, else by the owner of word (ty+1,k).  Every non-white word thus
// has exactly one writer and all-white words none.
//   wcol[k]    = lo | hi << 4 | sel << 8 | has_sc << 24 | sbyte << 25 | sidx << 27
//                sel: PRMT selector taking byte i from word(lo) or word(hi); the straddling column
//                sc[sidx] (if any) is byte `sbyte`, its tap-0 cell column is lo and its tap-1 column hi
//   cellgeo[c] = x: woff | nrows << 16 | nwords << 24     first word of the cell's band rows in a plane, rows of
//                                                         the band without its straddling row, word columns
//                y: kfirst | ty << 8 | q_below << 16 | q_above << 24   (straddling-row index, 0xff = none)
//                z: y0 | tx << 8                          first row of the band, cell column
inline void xw_build_paint_tables(XwRenderTables& t) {
    t.sp_ok = false;
    t.max_band_rows = 0;
    if (!t.fast_ok || t.W > 16 || t.H > 255) return;
    const int W = t.W, H = t.H, WR = t.WR;
    std::vector<uint8_t> is_sc(t.OW, 0), is_sr(t.OH, 0);
    for (int16_t v : t.sc) is_sc[v] = 1;
    for (int16_t v : t.sr) is_sr[v] = 1;
    t.wcol.assign(WR, 0);
    std::vector<int> kfirst(W, -1), klast(W, -1);
    for (int k = 0; k < WR; ++k) {
        int tx[4], lo = 1 << 30, hi = -1, nsc = 0, sbyte = 0, sidx = 0;
        for (int i = 0; i < 4; ++i) {
            tx[i] = t.xofs[4 * k + i] >> 6;
            lo = std::min(lo, tx[i]); hi = std::max(hi, tx[i]);
            if (is_sc[4 * k + i]) {
                ++nsc; sbyte = i; hi = std::max(hi, tx[i] + 1);
                for (size_t si = 0; si < t.sc.size(); ++si) if (t.sc[si] == 4 * k + i) sidx = (int)si;
            }
        }
        if (hi - lo > 1 || nsc > 1 || hi >= W) return;
        // the straddling byte must have its tap-0 cell = lo (then tap 1 = hi)
        if (nsc && tx[sbyte] != lo) return;
        uint32_t sel = 0;
        for (int i = 0; i < 4; ++i) sel |= (uint32_t)(tx[i] == lo ? i : 4 + i) << (4 * i);
        t.wcol[k] = (uint32_t)lo | ((uint32_t)hi << 4) | (sel << 8) | ((uint32_t)(nsc ? 1 : 0) << 24) | ((uint32_t)sbyte << 25) | ((uint32_t)sidx << 27);
        for (int c = lo; c <= hi; ++c) { if (kfirst[c] < 0) kfirst[c] = k; klast[c] = k; }
    }
    t.wshare.assign(WR, 0xff);
    t.ns = 0;
    for (int k = 0; k < WR; ++k) if ((t.wcol[k] & 15) != ((t.wcol[k] >> 4) & 15)) t.wshare[k] = (uint8_t)t.ns++;
    t.wshare_strad.assign(WR, 0xff);
    t.ns_strad = 0;
    for (int k = 0; k < WR; ++k)
        if ((t.wcol[k] & 15) != ((t.wcol[k] >> 4) & 15) && ((t.wcol[k] >> 24) & 1)) t.wshare_strad[k] = (uint8_t)t.ns_strad++;
    t.nwc = 0;
    for (int c = 0; c < W; ++c) {
        if (kfirst[c] < 0) return;
        t.nwc = std::max(t.nwc, klast[c] - kfirst[c] + 1);
    }
    if (t.nwc > 4) return;
    t.cellgeo.assign((size_t)H * W, XwU4{0, 0, 0, 0});
    std::vector<int> qb(H, 0xff);
    for (int ty = 0; ty < H; ++ty) {
        const int y0 = t.band_y0[ty];
        if (y0 < 0) return;
        int y1 = y0;
        while (y1 < t.OH && (t.yofs[y1] >> 6) == ty) ++y1;
        const bool srow = is_sr[y1 - 1] != 0;
        if (srow) { for (size_t q = 0; q < t.sr.size(); ++q) if (t.sr[q] == y1 - 1) qb[ty] = (int)q; if (ty + 1 >= H) return; }
        const int nrows = y1 - y0 - (srow ? 1 : 0);
        t.max_band_rows = std::max(t.max_band_rows, nrows);
        for (int tx = 0; tx < W; ++tx) {
            XwU4& g = t.cellgeo[(size_t)ty * W + tx];
            g.x = (uint32_t)(y0 * WR + kfirst[tx]) | ((uint32_t)nrows << 16) | ((uint32_t)(klast[tx] - kfirst[tx] + 1) << 24);
            g.y = (uint32_t)kfirst[tx] | ((uint32_t)ty << 8) | ((uint32_t)qb[ty] << 16) | ((uint32_t)(ty > 0 ? qb[ty - 1] : 0xff) << 24);
            g.z = (uint32_t)y0 | ((uint32_t)tx << 8);
        }
    }
    t.sr_ty.assign(t.sr.size(), 0);
    for (int ty = 0; ty < H; ++ty) if (qb[ty] != 0xff) t.sr_ty[qb[ty]] = (uint8_t)ty;
    if (t.max_band_rows > 12) return;  // XW_SP_ROWS_MAX: rows a special slot keeps in registers
    t.sp_ok = true;
}

// Cost model of one item (instructions), calibrated on ncu source counters (profiles/).
inline double xw_item_cost(const XwPlanItem& it, int nc) {
    switch (it.type) {
        case XW_ITEM_M1: return 50 + nc * (6 + it.nrows * 5.0);
        case XW_ITEM_M2: return 80 + nc * (6 + it.nrows * 10.0);
        case XW_ITEM_M3: return 110 + nc * (8 + it.nrows * 13.0);
        default: return 90 + nc * (20.0 + (it.corner ? 25.0 : 0.0));
    }
}

// Bundle the items for groups of n_warps warps.  split_m3: M3 items per plane (three times the
// items, a third of the rows each) instead of one item for the three planes.
inline void xw_build_plan(XwRenderTables& t, int n_warps, bool split_m3, bool conflict_free = false, bool warp0_loader = false,
                          bool two_phase = false) {
    struct Bundle { std::vector<XwU4> slots; double cost; };
    std::vector<Bundle> bundles;
    t.bank_conflicts = 0;
    const int PW = t.OH * t.WR;
    for (int ty = 0; ty < XW_ITEM_TYPES; ++ty) {
        std::vector<std::pair<XwPlanItem, int>> v;  // (item, first plane), nc implied
        const bool per_plane = (ty == XW_ITEM_M3 && split_m3);
        for (const XwPlanItem& it : t.items) {
            if (it.type != ty) continue;
            if (per_plane) for (int c = 0; c < 3; ++c) v.push_back({it, c});
            else v.push_back({it, 0});
        }
        const int nc = per_plane ? 1 : 3;
        std::stable_sort(v.begin(), v.end(), [&](const std::pair<XwPlanItem, int>& a, const std::pair<XwPlanItem, int>& b) {
            return xw_item_cost(a.first, nc) > xw_item_cost(b.first, nc);
        });
        // Spread the items over ceil(n/32) bundles so that the lanes of a bundle touch different
        // shared-memory banks: every row offset is the same for all lanes, so the bank of a lane's
        // frame-buffer and brick-table words at plane step i is fixed by its first word offset, and one
        // 2-way conflict doubles the wavefronts of every load and store of the bundle.  Each 3-plane
        // item may visit the planes in a rotated order ((rot + i) % 3): two lanes on the same bank in
        // plane 0 then never meet.  Randomised greedy, best of a few hundred tries (create-time only).
        if (v.empty()) continue;
        const size_t n = v.size();
        size_t nb = (n + 31) / 32;
        std::vector<uint32_t> w0s(n);
        for (size_t j = 0; j < n; ++j) w0s[j] = (uint32_t)(v[j].second * PW + v[j].first.woff);
        const int steps = nc;
        auto bank_at = [&](size_t j, int rot, int i) { return (int)((w0s[j] + (uint32_t)(((rot + i) % 3) * PW)) & 31u); };
        std::vector<int> best_b(n, 0), best_r(n, 0);
        long best_conf = -1;
        uint32_t rng = 12345u + (uint32_t)ty;
        std::vector<size_t> order(n);
        if (conflict_free) nb += 1;  // one spare bundle makes a conflict-free split much more likely
        for (int attempt = 0; attempt < 400 && best_conf != 0; ++attempt) {
            for (size_t j = 0; j < n; ++j) order[j] = j;
            for (size_t j = n; j > 1; --j) { rng = rng * 1664525u + 1013904223u; std::swap(order[j - 1], order[(rng >> 8) % j]); }
            std::vector<int> occ(nb * 3 * 32, 0), fill(nb, 0), ab(n), ar(n);
            long conf = 0;
            for (size_t oi = 0; oi < n; ++oi) {
                const size_t j = order[oi];
                long bc = -1; int bb = 0, br = 0;
                for (size_t bi = 0; bi < nb; ++bi) {
                    if (fill[bi] >= 32) continue;
                    for (int rot = 0; rot < (steps == 3 ? 3 : 1); ++rot) {
                        long c = 0;
                        for (int i = 0; i < steps; ++i) c += occ[(bi * 3 + i) * 32 + bank_at(j, rot, i)];
                        c = c * 64 + fill[bi];  // fewest conflicts first, then the emptiest bundle
                        if (bc < 0 || c < bc) { bc = c; bb = (int)bi; br = rot; }
                    }
                }
                ab[j] = bb; ar[j] = br; fill[bb]++;
                for (int i = 0; i < steps; ++i) { int& o = occ[((size_t)bb * 3 + i) * 32 + bank_at(j, br, i)]; conf += o; ++o; }
            }
            if (best_conf < 0 || conf < best_conf) { best_conf = conf; best_b = ab; best_r = ar; }
        }
        t.bank_conflicts += (int)best_conf;
        std::vector<Bundle> mine(nb);
        for (size_t bi = 0; bi < nb; ++bi) mine[bi].cost = 0;
        for (size_t j = 0; j < n; ++j) {
            const XwPlanItem& it = v[j].first;
            const int c0 = v[j].second;
            Bundle& bd = mine[best_b[j]];
            XwU4 e;
            e.x = (uint32_t)it.cellA | ((uint32_t)it.cellB << 8) | ((uint32_t)it.sel << 16);
            e.y = w0s[j] | ((uint32_t)it.nrows << 16) | ((uint32_t)ty << 24) | ((uint32_t)c0 << 27) | ((uint32_t)nc << 29);
            e.z = (uint32_t)it.y0 | ((uint32_t)it.dx << 8) | ((uint32_t)it.scell << 16) | ((uint32_t)it.sbyte << 24);
            e.w = (uint32_t)(ty == XW_ITEM_M3 ? it.sidx : it.q) | ((uint32_t)it.k << 8) | ((uint32_t)it.band << 16) | ((uint32_t)best_r[j] << 24) |
                  ((uint32_t)(it.sidx & 31) << 26) | ((uint32_t)it.corner << 31);
            bd.slots.push_back(e);
            const double c = xw_item_cost(it, nc);
            if (c > bd.cost) bd.cost = c;
        }
        for (size_t b = 0; b < nb; ++b) {
            XwU4 pad = {0, (uint32_t)ty << 24, 0, 0};  // padding slot: nc = 0
            if (mine[b].slots.empty()) continue;
            while (mine[b].slots.size() < 32) mine[b].slots.push_back(pad);
            bundles.push_back(mine[b]);
        }
    }
    // Two phases: the straddling-row bundles (R, RC) read only tables, so the single-buffer kernel runs
    // them while the staging copies of the special cells are in flight; everything else follows.  Within
    // a phase: longest-processing-time-first assignment of bundles to the warps of a group.
    t.plan.clear();
    t.total_cost = 0; t.makespan = 0;
    for (int phase = 0; phase < 2; ++phase) {
        std::vector<Bundle> ph;
        for (const Bundle& b : bundles) {
            const int ty = (int)((b.slots[0].y >> 24) & 7);
            if ((two_phase && ty >= XW_ITEM_R) == (phase == 0)) ph.push_back(b);
        }
        std::stable_sort(ph.begin(), ph.end(), [](const Bundle& a, const Bundle& b) { return a.cost > b.cost; });
        std::vector<std::vector<int>> mine(n_warps);
        std::vector<double> load(n_warps, 0.0);
        if (phase == 1 && warp0_loader) load[0] = 250.0;  // k_render (pipe): warp 0 also loads cells and stages
        for (size_t b = 0; b < ph.size(); ++b) {
            int w = 0;
            for (int i = 1; i < n_warps; ++i) if (load[i] < load[w]) w = i;
            mine[w].push_back((int)b);
            load[w] += ph[b].cost;
            t.total_cost += ph[b].cost;
        }
        size_t rounds = 0;
        double mk = 0;
        for (int w = 0; w < n_warps; ++w) { if (mine[w].size() > rounds) rounds = mine[w].size(); if (load[w] > mk) mk = load[w]; }
        t.makespan += mk;
        const size_t base = t.plan.size();
        t.plan.resize(base + rounds * n_warps * 32, XwU4{0, 0, 0, 0});
        for (int w = 0; w < n_warps; ++w)
            for (size_t j = 0; j < mine[w].size(); ++j)
                for (int l = 0; l < 32; ++l) t.plan[base + (j * n_warps + w) * 32 + l] = ph[mine[w][j]].slots[l];
        if (phase == 0) t.n_plan1 = (int)t.plan.size();
    }
    t.n_warps = n_warps;
}
