// hostsim.cpp -- TEST-ONLY host build of the engine's per-env device functions
// (xworld_b200/csrc/*.cuh compiled with XW_HD = inline).  It lets the CPU test-suite run the exact
// code the CUDA kernels inline -- reset, step, compose, straddle fix-up -- against the oracle in a
// container without a GPU.  It is NOT a product path: nothing in xworld_b200/ loads it, and the C ABI
// returns XW_ERR_NO_DEVICE when there is no GPU.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../xworld_b200/csrc/xw_common.cuh"
#include "../../xworld_b200/csrc/xw_race.cuh"
#include "../../xworld_b200/csrc/xw_render.cuh"
#include "../../xworld_b200/csrc/xw_render_host.hpp"
#include "../../xworld_b200/csrc/xw_reset.cuh"
#include "../../xworld_b200/csrc/xw_step.cuh"
#include "../../xworld_b200/csrc/xw_fpv.cuh"
#include "../../xworld_b200/csrc/xw_fpv_host.hpp"
#include "../../xworld_b200/csrc/xw_teacher_names.hpp"

struct HostSim {
    xw_config cfg;
    XwDev d;
    XwRender r;
    XwRenderTables tab;
    std::vector<std::vector<uint8_t>> bufs;
    std::vector<uint8_t> T;
    std::vector<uint16_t> ecol, uv;
    std::vector<uint32_t> corner;
    std::vector<uint8_t> colL, colR, rowT, rowB, cwb;
    std::vector<uint32_t> ctab, cornerP, TC, rowT2, rowB2;
    XwRaceCfg race;
    // first-person view
    XwFpv fpv;
    std::vector<int16_t> ft[12];
    std::vector<uint8_t> agent4, pmap, Tb, Ta, c2b;
    std::vector<uint32_t> gcache;
    std::vector<uint16_t> taps;
    std::vector<int16_t> itab;
    std::vector<double> yaw_cs;
    int32_t n_invalid = 0;
};

template <typename T>
static T* halloc(HostSim* s, size_t count) {
    s->bufs.emplace_back(count * sizeof(T) + 16, 0);
    return (T*)s->bufs.back().data();
}

static uint32_t minstd_seed(int32_t simulator_seed, int64_t thread_no) {
    int32_t seed = (int32_t)(uint32_t)std::hash<std::string>()(std::to_string((int)(simulator_seed + thread_no)));
    uint64_t x = ((uint64_t)(int64_t)seed) % 2147483647ull;
    return x == 0 ? 1u : (uint32_t)x;
}

extern "C" {

HostSim* hs_create(const xw_config* c, const xw_catalog* cat, int n) {
    HostSim* s = new HostSim();
    s->cfg = *c;
    if (c->game == XW_GAME_SIMPLE_RACE) {
        XwRaceCfg& r = s->race;
        memset(&r, 0, sizeof r);
        r.n = n; r.track_type = c->track_type; r.full_manouver = c->race_full_manouver; r.difficulty = c->difficulty;
        r.max_steps = c->max_steps; r.auto_reset = c->auto_reset; r.reward_scale = (double)c->reward_scale;
        r.mid_x = 480 / 2; r.mid_y = 720 / 2;
        if (c->track_type == 0) {
            r.length = c->track_length; r.width = c->track_width;
            r.start_y = r.mid_y - (float)(0.4 * r.length);
            r.end_y = r.mid_y + (float)(0.6 * r.length);
            r.start_px = r.mid_x - 0.0f; r.start_py = r.start_y;
        } else {
            r.inner = c->track_radius; r.width = c->track_width; r.outer = r.inner + r.width;
            r.start_px = (r.inner + r.width / 2) + r.mid_x; r.start_py = 0.0f + r.mid_y;
        }
        r.pos_x = halloc<float>(s, n); r.pos_y = halloc<float>(s, n); r.angle = halloc<float>(s, n);
        r.state = halloc<float>(s, (size_t)n * 4); r.steps = halloc<int32_t>(s, n);
        r.random = c->race_random != 0;
        if (r.random) {
            r.minstd = halloc<uint32_t>(s, n);
            for (int i = 0; i < n; ++i) r.minstd[i] = minstd_seed(c->simulator_seed, c->env_id_offset + i + 1);
        }
        return s;
    }
    XwDev& d = s->d;
    memset(&d, 0, sizeof d);
    d.n = n; d.H = c->height; d.W = c->width; d.CS = (c->height * c->width + 15) & ~15;
    d.G = c->n_goals; d.n_blocks = c->n_blocks; d.rules = c->rules; d.max_steps = c->max_steps;
    d.max_steps_factor = c->max_steps_factor; d.auto_reset = c->auto_reset; d.seed = c->seed; d.gid0 = c->env_id_offset;
    d.vr = c->visible_radius < c->height ? c->visible_radius : c->height; d.task_mode = c->task_mode;
    d.n_invalid = &s->n_invalid;
    if (d.vr > 0) {
        d.goal_yaw = halloc<uint16_t>(s, (size_t)n * XW_MAX_GOALS);
        d.goal_scale = halloc<double>(s, (size_t)n * XW_MAX_GOALS); d.goal_offset = halloc<double>(s, (size_t)n * XW_MAX_GOALS);
        s->yaw_cs = xw_fpv_yaw_table();
        d.yaw_cs = s->yaw_cs.data();
    }
    d.grid = halloc<uint8_t>(s, (size_t)n * d.CS);
    uint8_t** u8s[] = {&d.agent_x, &d.agent_y, &d.facing, &d.task, &d.stage, &d.event, &d.succ, &d.tmask, &d.aux0, &d.aux1, &d.aux2};
    for (auto p : u8s) *p = halloc<uint8_t>(s, n);
    d.goal_x = halloc<uint8_t>(s, (size_t)n * XW_MAX_GOALS); d.goal_y = halloc<uint8_t>(s, (size_t)n * XW_MAX_GOALS);
    d.goal_icon = halloc<int32_t>(s, (size_t)n * XW_MAX_GOALS); d.goal_name = halloc<int32_t>(s, (size_t)n * XW_MAX_GOALS);
    int32_t** i32s[] = {&d.steps_in_task, &d.num_steps, &d.episode, &d.n_success, &d.n_failure, &d.success_steps, &d.error};
    for (auto p : i32s) *p = halloc<int32_t>(s, n);
    d.minstd = halloc<uint32_t>(s, n);
    d.reset_count = halloc<int32_t>(s, 2); d.reset_list = halloc<int32_t>(s, n);
    if (c->curriculum != 0) {
        d.curriculum = (double)c->curriculum; d.check_period = c->curriculum_check_period > 0 ? c->curriculum_check_period : 100;
        d.level = halloc<uint8_t>(s, n); d.check_counter = halloc<int32_t>(s, n);
        d.win_len = halloc<uint8_t>(s, (size_t)n * XW_N_T3); d.win_pos = halloc<uint8_t>(s, (size_t)n * XW_N_T3);
        d.win_sum = halloc<uint8_t>(s, (size_t)n * XW_N_T3); d.win_bits = halloc<uint32_t>(s, (size_t)n * XW_N_T3 * XW_WIN_WORDS);
        for (int i = 0; i < n; ++i) d.level[i] = (uint8_t)c->start_level;
    }
    for (int i = 0; i < n; ++i) d.minstd[i] = minstd_seed(c->simulator_seed, c->env_id_offset + i + 1);
    d.n_names = cat->n_names; d.brick_icon = cat->brick_icon; d.agent_icon = cat->agent_icon;
    d.name_first = cat->name_first; d.name_icons = cat->name_icons; d.icon_colored = cat->icon_colored;
    int OH = c->out_h > 0 ? c->out_h : c->height * 12, OW = c->out_w > 0 ? c->out_w : c->width * 12;
    if (d.vr > 0) {  // create_fpv (xw_engine.cu)
        if (c->out_h <= 0) OH = d.vr * (84 / d.vr);
        if (c->out_w <= 0) OW = d.vr * (84 / d.vr);
        XwFpv& F = s->fpv;
        memset(&F, 0, sizeof F);
        F.vr = d.vr; F.N = d.vr * 64; F.CH = c->height * 64; F.OH = OH; F.OW = OW; F.FB = 3 * OH * OW;
        F.ident1 = F.N == F.CH; F.ident2 = F.CH == OH && F.CH == OW;
        F.G = c->n_goals; F.brick_icon = cat->brick_icon; F.agent_icon = cat->agent_icon;
        xw_fpv_resize_tables(F.N, F.CH, false, s->ft[0], s->ft[1], s->ft[2]);
        xw_fpv_resize_tables(F.N, F.CH, true, s->ft[3], s->ft[4], s->ft[5]);
        xw_fpv_resize_tables(F.CH, OW, false, s->ft[6], s->ft[7], s->ft[8]);
        xw_fpv_resize_tables(F.CH, OH, true, s->ft[9], s->ft[10], s->ft[11]);
        F.x1ofs = s->ft[0].data(); F.x1a0 = s->ft[1].data(); F.x1a1 = s->ft[2].data();
        F.y1ofs = s->ft[3].data(); F.y1a0 = s->ft[4].data(); F.y1a1 = s->ft[5].data();
        F.x2ofs = s->ft[6].data(); F.x2a0 = s->ft[7].data(); F.x2a1 = s->ft[8].data();
        F.y2ofs = s->ft[9].data(); F.y2a0 = s->ft[10].data(); F.y2a1 = s->ft[11].data();
        F.atlas64 = cat->atlas64;
        s->agent4 = xw_fpv_agent_icons(cat->atlas64 + (size_t)cat->agent_icon * 12288);
        F.agent4 = s->agent4.data();
        s->itab.resize(32 * 32 * 4);
        xw_fpv_build_itab(s->itab.data());
        F.itab = s->itab.data();
        s->gcache.assign((size_t)n * F.G * 4096, 0);
        F.gcache = s->gcache.data();
        s->pmap.assign((size_t)4 * OH * OW, 0); s->Tb.assign((size_t)12 * OH * OW, 0); s->Ta.assign((size_t)12 * OH * OW, 0);
        for (size_t i = 0; i < (size_t)4 * OH * OW; ++i) xw_fpv_table_entry(F, i, s->pmap.data(), s->Tb.data(), s->Ta.data());
        F.pmap = s->pmap.data(); F.Tb = s->Tb.data(); F.Ta = s->Ta.data();
        if (OW % 4 == 0 && F.FB % 16 == 0 && OW == OH && OW % F.vr == 0 && (OW / F.vr) % 4 == 0) {  // create_fpv's "regular" test
            const int bs = OW / F.vr, ncell = F.vr * F.vr;
            s->c2b.assign((size_t)4 * ncell, 0xff);
            bool regular = true;
            for (int f = 0; f < 4 && regular; ++f)
                for (int p = 0; p < OH * OW && regular; ++p) {
                    const int blk = ((p / OW) / bs) * F.vr + (p % OW) / bs;
                    const uint8_t cell = s->pmap[(size_t)f * OH * OW + p];
                    if (cell == 0xff || cell >= ncell) { regular = false; break; }
                    uint8_t& slot = s->c2b[(size_t)f * ncell + cell];
                    if (slot == 0xff) slot = (uint8_t)blk; else if (slot != blk) regular = false;
                }
            for (uint8_t v : s->c2b) if (v == 0xff) regular = false;
            if (regular) {
                s->taps.assign((size_t)4 * OH * OW * 32, 0);
                for (size_t i = 0; i < (size_t)4 * OH * OW; ++i) xw_fpv_tap_entry(F, i, s->taps.data());
                F.regular = 1; F.bs = bs; F.taps = s->taps.data(); F.cell2block = s->c2b.data();
            }
        }
        s->r.OH = OH; s->r.OW = OW; s->r.FB = F.FB;
        return s;
    }
    s->tab = xw_build_render_tables(c->height, c->width, OH, OW);
    XwRenderTables& t = s->tab;
    XwRender& r = s->r;
    memset(&r, 0, sizeof r);
    r.OH = OH; r.OW = OW; r.WR = t.WR; r.FB = t.FB; r.H = c->height; r.W = c->width;
    r.n_icons = cat->n_icons; r.brick_icon = cat->brick_icon; r.agent_icon = cat->agent_icon;
    r.taps.xofs = t.xofs.data(); r.taps.xa0 = t.xa0.data(); r.taps.xa1 = t.xa1.data();
    r.taps.yofs = t.yofs.data(); r.taps.ya0 = t.ya0.data(); r.taps.ya1 = t.ya1.data();
    if (t.fast_ok) {  // odd map sides exercise the per-plane M3 split, even ones the 3-plane items
        xw_build_plan(t, 4 + c->height % 3, c->height % 2 != 0);
        r.plan = t.plan.data(); r.n_plan = (int)t.plan.size(); r.n_plan1 = t.n_plan1; r.G = 1; r.GT = t.n_warps * 32;
        r.cellinfo = t.cellinfo.data();
    }
    xw_build_paint_tables(t);
    if (t.sp_ok) {
        r.cellgeo = t.cellgeo.data(); r.wcol = t.wcol.data(); r.wshare = t.wshare.data(); r.sr_ty = t.sr_ty.data();
        r.nwc = t.nwc; r.ns = t.ns; r.slot_magic = 65536 / t.nwc + 1;
        r.n_sc = (int)t.sc.size(); r.n_sr = (int)t.sr.size();
        r.ctab_merge = xw_want_ctab_merge(r, t, 232448) ? 1 : 0;   // (as xw_engine.cu: 227 KB of shared memory per CTA)
        if (r.ctab_merge) { t.wshare = t.wshare_strad; t.ns = t.ns_strad; r.ns = t.ns; r.wshare = t.wshare.data(); }
        r.tc_rows = 12;  // (the host build always uses the 12-row instantiation)
    }
    r.n_sr = (int)t.sr.size();
    r.sr = t.sr.data();
    r.atlas64 = cat->atlas64;
    if (t.fast_ok) {  // k_build_edge_tables
        r.band_y0 = t.band_y0.data(); r.RB = t.RB;
        const size_t n_ecol = (size_t)(cat->n_icons + 1) * 2 * 3 * c->height * t.RB;
        s->ecol.resize(n_ecol + XW_TABLE_PAD / 2);
        s->uv.resize((size_t)(cat->n_icons + 1) * r.n_sr * 2 * 3 * OW + 8);
        for (size_t i = 0; i < n_ecol; ++i) {
            size_t j = i;
            const int row = (int)(j % t.RB); j /= t.RB;
            const int band = (int)(j % c->height); j /= c->height;
            const int cc = (int)(j % 3); j /= 3;
            s->ecol[i] = xw_ecol_entry(r, (uint32_t)(j / 2), (int)(j % 2), cc, band, row);
        }
        for (size_t i = 0; i < (size_t)(cat->n_icons + 1) * r.n_sr * 2 * 3 * OW; ++i) {
            size_t j = i;
            const int dx = (int)(j % OW); j /= OW;
            const int cc = (int)(j % 3); j /= 3;
            const int role = (int)(j % 2); j /= 2;
            s->uv[i] = xw_uv_entry(r, (uint32_t)(j / r.n_sr), (int)(j % r.n_sr), role, cc, dx);
        }
        s->corner.resize((size_t)(cat->n_icons + 1) * 3);
        for (size_t j = 0; j < s->corner.size(); ++j) s->corner[j] = xw_corner_entry(r, (uint32_t)(j / 3), (int)(j % 3));
        r.ecol = s->ecol.data(); r.uv = s->uv.data(); r.corner = s->corner.data();
        r.sc = t.sc.data(); r.n_sc = (int)t.sc.size();
        {  // k_build_pair_tables
            const size_t cs = xw_colpair_stride(r), rs = xw_rowpair_stride(r);
            const size_t n_col = (size_t)(cat->n_icons + 1) * 2 * cs, n_row = (size_t)(cat->n_icons + 1) * 2 * rs;
            s->colL.assign(n_col + 64, 0); s->colR.assign(n_col + 64, 0); s->rowT.assign(n_row + 64, 0); s->rowB.assign(n_row + 64, 0);
            for (int role = 0; role < 2; ++role) {
                for (size_t i = 0; i < n_col; ++i) {
                    size_t j = i;
                    const int row = (int)(j % r.RB); j /= r.RB;
                    const int band = (int)(j % r.H); j /= r.H;
                    const int cc = (int)(j % 3); j /= 3;
                    const int si = (int)(j % r.n_sc); j /= r.n_sc;
                    (role ? s->colR : s->colL)[i] = xw_colpair_entry(r, role, (uint32_t)(j / 2), (int)(j % 2), si, cc, band, row);
                }
                for (size_t i = 0; i < n_row; ++i) {
                    size_t j = i;
                    const int dx = (int)(j % r.OW); j /= r.OW;
                    const int cc = (int)(j % 3); j /= 3;
                    const int q = (int)(j % r.n_sr); j /= r.n_sr;
                    (role ? s->rowB : s->rowT)[i] = xw_rowpair_entry(r, role, (uint32_t)(j / 2), (int)(j % 2), q, cc, dx);
                }
            }
            r.colL = s->colL.data(); r.colR = s->colR.data(); r.rowT = s->rowT.data(); r.rowB = s->rowB.data();
            s->cwb.assign((size_t)16 * r.n_sr * r.n_sc * 3 + 64, 0);
            for (size_t i = 0; i < (size_t)16 * r.n_sr * r.n_sc * 3; ++i) {
                size_t j = i;
                const int cc = (int)(j % 3); j /= 3;
                const int si = (int)(j % r.n_sc); j /= r.n_sc;
                s->cwb[i] = xw_cornerwb_entry(r, (int)(j / r.n_sr), (int)(j % r.n_sr), si, cc);
            }
            r.cornerWB = s->cwb.data();
        }
    }
    r.atlas64 = cat->atlas64;
    return s;
}

// Build the phase atlas only for the icons in use (the full 363-icon table takes a while on one core).
static void hs_build_cell_tables(HostSim* s, const int32_t* icons, int n_icons);
void hs_build_phase_atlas(HostSim* s, const int32_t* icons, int n_icons) {
    XwRender& r = s->r;
    s->T.assign((size_t)r.n_icons * r.FB + XW_TABLE_PAD, 0);
    r.T = s->T.data();
    for (int q = 0; q < n_icons; ++q) {
        int icon = icons[q];
        for (int i = 0; i < r.FB; ++i) {
            int c = i / (r.OH * r.OW), p = i % (r.OH * r.OW);
            s->T[(size_t)icon * r.FB + i] = xw_phase_px(r, icon, c, p / r.OW, p % r.OW);
        }
    }
    hs_build_cell_tables(s, icons, n_icons);
}

// k_build_cell_tables, for the icons whose phase tables exist
static void hs_build_cell_tables(HostSim* s, const int32_t* icons, int n_icons) {
    XwRender& r = s->r;
    if (!s->tab.sp_ok) return;
    const size_t per_icon = (size_t)r.H * r.W * 3 * r.nwc * r.tc_rows;
    s->TC.resize((size_t)r.n_icons * per_icon + 16, 0);
    r.TC = s->TC.data();
    for (int q = 0; q < n_icons; ++q)
        for (size_t i = 0; i < per_icon; ++i) s->TC[(size_t)icons[q] * per_icon + i] = xw_tc_word(r, (size_t)icons[q] * per_icon + i);
}
int hs_fast_ok(HostSim* s) { return s->tab.fast_ok; }
int hs_threads(HostSim* s) { return XW_RENDER_THREADS; }
void hs_destroy(HostSim* s) { delete s; }

// k_fpv_warp_goals for one env
static void hs_warp_goals(HostSim* s, int e) {
    XwDev& d = s->d;
    XwFpv& F = s->fpv;
    if (d.vr == 0) return;
    for (int g = 0; g < F.G; ++g) {
        const size_t k = (size_t)g * d.n + e;
        int co[4][64];
        for (int i = 0; i < 64; ++i)
            xw_fpv_warp_coeffs(d.yaw_cs, d.goal_yaw[k], d.goal_scale[k], d.goal_offset[k], i, &co[0][i], &co[1][i], &co[2][i], &co[3][i]);
        const uint8_t* icon = F.atlas64 + (size_t)d.goal_icon[k] * 12288;
        uint32_t* dst = F.gcache + ((size_t)e * F.G + g) * 4096;
        for (int p = 0; p < 4096; ++p)
            dst[p] = xw_fpv_warp_px(icon, F.itab, (co[2][p >> 6] + co[0][p & 63]) >> 5, (co[3][p >> 6] + co[1][p & 63]) >> 5);
    }
}
static void hs_reset_one(HostSim* s, int e) { xw_reset_env(s->d, e); hs_warp_goals(s, e); }

void hs_reset(HostSim* s, const uint8_t* mask) {
    if (s->cfg.game == XW_GAME_SIMPLE_RACE) { for (int e = 0; e < s->race.n; ++e) if (!mask || mask[e]) xw_race_reset_env(s->race, e); return; }
    for (int e = 0; e < s->d.n; ++e) if (!mask || mask[e]) hs_reset_one(s, e);
}

void hs_step(HostSim* s, const int32_t* actions, int act_rep, float* reward, int32_t* over) {
    if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        for (int e = 0; e < s->race.n; ++e) if (xw_race_step_env(s->race, e, actions[e], act_rep, &reward[e], &over[e])) xw_race_reset_env(s->race, e);
        return;
    }
    for (int e = 0; e < s->d.n; ++e)
        if (actions[e] != XW_ACTION_NONE && xw_step_env(s->d, e, actions[e], act_rep, &reward[e], &over[e])) hs_reset_one(s, e);
}

// Emulates one k_render warp group per env: celldesc, every item of the plan, copy out.  Configs the
// compositor does not take go through the generic per-pixel rule (k_render_generic's arithmetic).
// mode 0: the plan compositor (k_render / k_render_sb); mode 1: the sparse painter (k_render_sp);
// mode 2: the per-pixel rule on the 64-px atlas (k_render_generic's arithmetic) -- the ground truth
int hs_sp_ok(HostSim* s) { return s->tab.sp_ok; }
void hs_render_mode(HostSim* s, uint8_t* frames, int mode);
void hs_render(HostSim* s, uint8_t* frames) { hs_render_mode(s, frames, s->d.vr > 0 || s->tab.sp_ok ? 1 : 0); }
// k_render_fpv (mode != 2: the table pass + the exact pass) / k_render_fpv_generic (mode 2: every pixel tap by tap)
static void hs_render_fpv(HostSim* s, uint8_t* frames, int mode) {
    XwDev& d = s->d;
    XwFpv& F = s->fpv;
    const int plane = F.OH * F.OW;
    uint8_t ccode[XW_MAX_DIM * XW_MAX_DIM];
    for (int e = 0; e < d.n; ++e) {
        for (int k = 0; k < F.vr; ++k) xw_fpv_cells_line(d, e, k, ccode);
        XwFpvEnvFetch fe;
        fe.F = &F; fe.ccode = ccode; fe.gc = F.gcache + (size_t)e * F.G * 4096; fe.facing = d.facing[e];
        uint8_t* out = frames + (size_t)e * F.FB;
        const int f = d.facing[e];
        if (mode != 2 && F.regular) {  // k_render_fpv_cells: white fill, then cell blocks
            memset(out, 0x5a, F.FB);   // (every cell block is written: nothing of this survives)
            const int vr = F.vr, bs = F.bs;
            for (int c = 0; c < vr * vr; ++c) {
                const int code = ccode[c];
                const int blk = F.cell2block[f * vr * vr + c], by = blk / vr, bx = blk % vr;
                for (int i = 0; i < bs * bs; ++i) {
                    const int p = (by * bs + i / bs) * F.OW + bx * bs + i % bs;
                    uint32_t v = 0;
                    if (code >= XW_CELL_GOAL0 && code != XW_FPV_BLACK) v = xw_fpv_px_taps(F.taps + ((size_t)f * plane + p) * 32, fe.gc + (size_t)(code - XW_CELL_GOAL0) * 4096);
                    else if (code == XW_CELL_EMPTY) v = 0xffffffu;
                    else if (code != XW_FPV_BLACK) {
                        const uint8_t* t = (code == XW_CELL_BLOCK ? F.Tb : F.Ta) + (size_t)f * 3 * plane + p;
                        v = (uint32_t)t[0] | ((uint32_t)t[plane] << 8) | ((uint32_t)t[2 * (size_t)plane] << 16);
                    }
                    out[p] = (uint8_t)v; out[plane + p] = (uint8_t)(v >> 8); out[2 * plane + p] = (uint8_t)(v >> 16);
                }
            }
            continue;
        }
        for (int p = 0; p < plane; ++p) {
            const int id = F.pmap[(size_t)f * plane + p];
            const int code = id == 255 ? -1 : ccode[id];
            uint32_t v;
            if (mode != 2 && code == XW_CELL_EMPTY) v = 0xffffffu;
            else if (mode != 2 && code == XW_FPV_BLACK) v = 0;
            else if (mode != 2 && (code == XW_CELL_BLOCK || code == XW_CELL_AGENT)) {
                const uint8_t* t = (code == XW_CELL_BLOCK ? F.Tb : F.Ta) + (size_t)f * 3 * plane + p;
                v = (uint32_t)t[0] | ((uint32_t)t[plane] << 8) | ((uint32_t)t[2 * (size_t)plane] << 16);
            } else v = xw_fpv_px(F, p / F.OW, p % F.OW, fe);
            out[p] = (uint8_t)v; out[plane + p] = (uint8_t)(v >> 8); out[2 * plane + p] = (uint8_t)(v >> 16);
        }
    }
}

void hs_render_mode(HostSim* s, uint8_t* frames, int mode) {
    if (s->d.vr > 0) { hs_render_fpv(s, frames, mode); return; }
    XwRender& r = s->r;
    XwDev& d = s->d;
    std::vector<uint32_t> fb(r.FB / 4 + 4), yb(r.OH);
    std::vector<uint8_t> code(XW_CELL_STRIDE, 0);
    uint32_t icon[XW_CODE_SLOTS];
    for (int i = 0; i < r.OH; ++i) yb[i] = (uint32_t)(uint16_t)r.taps.ya0[i] | ((uint32_t)(uint16_t)r.taps.ya1[i] << 16);
    XwComposeCtx x;
    x.hot = r.T + (size_t)r.brick_icon * r.FB; x.yb = yb.data();
    std::vector<uint8_t> pair_hot;  // the kernel's shared-memory copies: [white, brick][cls][...]
    if (s->tab.fast_ok) {
        const size_t cs2 = 2 * xw_colpair_stride(r), rs2 = 2 * xw_rowpair_stride(r);
        pair_hot.resize(2 * cs2 + 2 * rs2 + 64);
        memcpy(pair_hot.data(), r.colL, cs2); memcpy(pair_hot.data() + cs2, r.colL + (size_t)(r.brick_icon + 1) * cs2, cs2);
        memcpy(pair_hot.data() + 2 * cs2, r.rowT, rs2); memcpy(pair_hot.data() + 2 * cs2 + rs2, r.rowT + (size_t)(r.brick_icon + 1) * rs2, rs2);
        x.colL_hot = pair_hot.data(); x.rowT_hot = pair_hot.data() + 2 * cs2;
        x.cornerWB = r.cornerWB;
    }
    XwCells cells;
    cells.code = code.data(); cells.icon = icon;
    for (int e = 0; e < d.n; ++e) {
        memcpy(code.data(), d.grid + (size_t)e * d.CS, d.CS);
        for (int k = 0; k < XW_CODE_SLOTS; ++k) icon[k] = k < XW_CELL_GOAL0 + d.G ? xw_cell_desc(d, e, k) : 0;
        if (s->tab.fast_ok && mode == 1 && s->tab.sp_ok) {
            // k_render_sp: white pre-fill, special slots + straddling-row words (PRE), brick slots (POST)
            if (s->ctab.empty()) {  // k_build_class_tables
                s->cornerP.resize((size_t)(r.n_icons + 1) * 4);
                for (size_t i = 0; i < s->cornerP.size(); ++i) s->cornerP[i] = xw_cornerP_entry(r, (uint32_t)(i / 4), (int)(i % 4));
                r.cornerP = s->cornerP.data();
                s->ctab.resize(xw_ctab_words(r) + 64, 0);
                r.ctab = s->ctab.data();
                const size_t per = (size_t)xw_ctab_stride(r);
                for (size_t i = 0; i < xw_ctab_words(r); ++i) {
                    const int col = (int)(i / per), rem = (int)(i % per), p = rem / r.OH, dy = rem % r.OH;
                    if (p >= 3 || col >= r.WR + 2 * r.ns) continue;
                    int variant = 0, k = col;
                    if (col >= r.WR) {
                        variant = col - r.WR < r.ns ? 1 : 2;
                        const int si = col - r.WR - (variant == 2 ? r.ns : 0);
                        for (k = 0; k < r.WR; ++k) if (r.wshare[k] == si) break;
                    }
                    s->ctab[i] = xw_ctab_word(r, r.wcol[k], variant, k, p, dy);
                }
            }
            std::fill(fb.begin(), fb.end(), 0xffffffffu);
            std::vector<uint8_t> sr_dy(r.n_sr + 1);
            for (int q = 0; q < r.n_sr; ++q) sr_dy[q] = (uint8_t)r.sr[q];
            std::vector<uint32_t> sc_a(r.n_sc + 1);
            for (int i = 0; i < r.n_sc; ++i) sc_a[i] = (uint32_t)(uint16_t)r.taps.xa0[r.sc[i]] | ((uint32_t)(uint16_t)r.taps.xa1[r.sc[i]] << 16);
            if (s->rowT2.empty()) {  // k_build_row2_tables
                const size_t tot = (size_t)(r.n_icons + 1) * 2 * r.n_sr * r.WR * 3;
                s->rowT2.resize(tot + 16); s->rowB2.resize(tot + 16);
                for (size_t i = 0; i < tot; ++i) { s->rowT2[i] = xw_row2_word(r, r.rowT, i); s->rowB2[i] = xw_row2_word(r, r.rowB, i); }
                r.rowT2 = s->rowT2.data(); r.rowB2 = s->rowB2.data();
            }
            const size_t rs2w = (size_t)2 * r.n_sr * r.WR * 3;
            std::vector<uint32_t> row2_hot(2 * rs2w + 16);
            for (size_t i = 0; i < rs2w; ++i) { row2_hot[i] = r.rowT2[i]; row2_hot[rs2w + i] = r.rowT2[(size_t)(r.brick_icon + 1) * rs2w + i]; }
            XwPaintCtx pg;
            pg.row2_hot = row2_hot.data();
            pg.sc_a = sc_a.data();
            pg.cellgeo = r.cellgeo; pg.wcol = r.wcol; pg.wshare = r.wshare; pg.ctab = r.ctab; pg.sr_ty = r.sr_ty; pg.sr_dy = sr_dy.data();
            std::vector<uint32_t> written(r.FB / 4, 0);  // every frame word has at most one writer
            const bool fixed = r.WR == 21 && r.OH == 84 && (e & 1);    // alternate the compile-time and the run-time row stride
            auto mark = [&](uint32_t w) { if (written[w]++) abort(); };
            for (int cell = 0; cell < d.H * d.W; ++cell) {
                if (code[cell] < XW_CELL_AGENT) continue;
                for (int wc = 0; wc < r.nwc; ++wc) {
                    for (int p = 0; p < 3; ++p) {
                        uint32_t m[12], pb[3];
                        memset(m, 0x5a, sizeof m); memset(pb, 0x5a, sizeof pb);
                        XwRender r_row = r;   // every other pair of envs: without the cell-major table (the big-map path)
                        if (e & 2) r_row.TC = nullptr;
                        const bool have = fixed ? xw_sp_special<21, 12, 0>(r_row, x, pg, cells, cell, wc, p, m, pb, nullptr)
                                                : xw_sp_special<0, 12, 0>(r_row, x, pg, cells, cell, wc, p, m, pb, nullptr);
                        if (!have) continue;
                        std::vector<uint32_t> before(fb);
                        const bool have2 = fixed ? xw_sp_special<21, 12, 1>(r, x, pg, cells, cell, wc, p, m, pb, fb.data())
                                                 : xw_sp_special<0, 12, 1>(r, x, pg, cells, cell, wc, p, m, pb, fb.data());
                        if (!have2) abort();
                        for (size_t w = 0; w < (size_t)r.FB / 4; ++w) if (fb[w] != before[w]) mark((uint32_t)w);
                    }
                }
            }
            for (int idx = 0; idx < r.n_sr * r.WR; ++idx) {
                uint32_t wa[3], wb[3], tt[4];
                memset(wa, 0x5a, sizeof wa); memset(wb, 0x5a, sizeof wb); memset(tt, 0x5a, sizeof tt);
                const int q = idx / r.WR, k = idx % r.WR;
                if (fixed) { xw_sp_rword<21, 0>(r, x, pg, cells, q, k, wa, wb, tt, nullptr); xw_sp_rword<21, 1>(r, x, pg, cells, q, k, wa, wb, tt, fb.data()); }
                else { xw_sp_rword<0, 0>(r, x, pg, cells, q, k, wa, wb, tt, nullptr); xw_sp_rword<0, 1>(r, x, pg, cells, q, k, wa, wb, tt, fb.data()); }
                for (int p = 0; p < 3; ++p) mark(p * r.OH * r.WR + r.sr[q] * r.WR + k);
            }
            std::vector<uint8_t> list;
            for (int cell = 0; cell < d.H * d.W; ++cell) if (code[cell] == XW_CELL_BLOCK) list.push_back((uint8_t)cell);
            for (int sl = 0; sl < (int)list.size() * r.nwc; ++sl) {
                int i, wc;
                xw_sp_decode(r, sl, &i, &wc);
                if (i != sl / r.nwc || wc != sl % r.nwc) abort();
                std::vector<uint32_t> before(fb);
                if (fixed) xw_sp_brick_slot<21, 12>(r, pg, cells, list[i], wc, fb.data()); else xw_sp_brick_slot<0, 12>(r, pg, cells, list[i], wc, fb.data());
                for (size_t w = 0; w < (size_t)r.FB / 4; ++w) if (fb[w] != before[w]) mark((uint32_t)w);
            }
            memcpy(frames + (size_t)e * r.FB, fb.data(), r.FB);
        } else if (s->tab.fast_ok && mode != 2) {
            std::fill(fb.begin(), fb.end(), 0x5a5a5a5au);
            for (int cell = 0; cell < d.H * d.W; ++cell) {  // the kernel's staging pass (LDGSTS)
                if (code[cell] < XW_CELL_AGENT) continue;
                for (int col = 0; col < XW_STAGE_COLS; ++col) {
                    uint32_t w0; int nrows; const uint32_t* src;
                    if (xw_stage_column(r, cells, r.cellinfo, cell, col / 3, col % 3, &w0, &nrows, &src))
                        for (int j = 0; j < nrows; ++j) fb[w0 + j * r.WR] = src[j * r.WR];
                }
            }
            for (int i = 0; i < r.n_plan; ++i) {  // the test alternates the compile-time and run-time row stride
                if (r.WR == 21 && (e & 1)) xw_compose_item<21>(r, x, r.plan[i], cells, fb.data());
                else xw_compose_item<0>(r, x, r.plan[i], cells, fb.data());
            }
            memcpy(frames + (size_t)e * r.FB, fb.data(), r.FB);
        } else {
            for (int i = 0; i < r.FB; ++i) {
                const int c = i / (r.OH * r.OW), p = i % (r.OH * r.OW);
                frames[(size_t)e * r.FB + i] = xw_exact_px(r, cells, c, p / r.OW, p % r.OW);
            }
        }
    }
}

// test-only: overwrite env 0's grid and goal icons (golden-frame tests)
void hs_set_grid(HostSim* s, const uint8_t* grid, const int32_t* goal_icons) {
    memcpy(s->d.grid, grid, (size_t)s->d.H * s->d.W);
    for (int k = 0; k < s->d.G; ++k) s->d.goal_icon[(size_t)k * s->d.n] = goal_icons[k];
}

void hs_set_grid_env(HostSim* s, int e, const uint8_t* grid, const int32_t* goal_icons) {
    memcpy(s->d.grid + (size_t)e * s->d.CS, grid, (size_t)s->d.H * s->d.W);
    for (int k = 0; k < s->d.G; ++k) s->d.goal_icon[(size_t)k * s->d.n + e] = goal_icons[k];
}

// test-only: overwrite one env's whole first-person scene (golden-frame tests)
void hs_set_fpv_env(HostSim* s, int e, const uint8_t* grid, const int32_t* goal_icons, const uint16_t* yaw_idx, const double* scale,
                    const double* offset, int facing) {
    XwDev& d = s->d;
    memcpy(d.grid + (size_t)e * d.CS, grid, (size_t)d.H * d.W);
    for (int c = 0; c < d.H * d.W; ++c) if (grid[c] == XW_CELL_AGENT) { d.agent_x[e] = (uint8_t)(c % d.W); d.agent_y[e] = (uint8_t)(c / d.W); }
    for (int k = 0; k < d.G; ++k) {
        const size_t i = (size_t)k * d.n + e;
        d.goal_icon[i] = goal_icons[k]; d.goal_yaw[i] = yaw_idx[k]; d.goal_scale[i] = scale[k]; d.goal_offset[i] = offset[k];
    }
    d.facing[e] = (uint8_t)facing;
    hs_warp_goals(s, e);
}

const uint8_t* hs_fpv_pmap(HostSim* s) { return s->pmap.data(); }
int hs_rec_task_of_draw(uint32_t x) { return rec_task_of_draw(x); }

int hs_get_field(HostSim* s, const char* name, void* out) {
    XwDev& d = s->d;
    std::string k(name);
    size_t n = d.n;
    if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        XwRaceCfg& r = s->race;
        n = r.n;
        if (k == "pos_x") memcpy(out, r.pos_x, 4 * n); else if (k == "pos_y") memcpy(out, r.pos_y, 4 * n);
        else if (k == "angle") memcpy(out, r.angle, 4 * n); else if (k == "steps") memcpy(out, r.steps, 4 * n);
        else if (k == "state") memcpy(out, r.state, 16 * n); else return -1;
        return 0;
    }
    if (k == "grid") { for (size_t e = 0; e < n; ++e) memcpy((uint8_t*)out + e * d.H * d.W, d.grid + e * d.CS, d.H * d.W); return 0; }
    struct { const char* nm; uint8_t* p; } u8t[] = {{"agent_x", d.agent_x}, {"agent_y", d.agent_y}, {"facing", d.facing}, {"task", d.task},
        {"stage", d.stage}, {"event", d.event}, {"action_success", d.succ}, {"target_mask", d.tmask}, {"aux0", d.aux0}, {"aux1", d.aux1}, {"aux2", d.aux2}};
    for (auto& t : u8t) if (k == t.nm) { memcpy(out, t.p, n); return 0; }
    struct { const char* nm; void* p; } i32t[] = {{"steps_in_task", d.steps_in_task}, {"num_steps", d.num_steps}, {"episode", d.episode},
        {"n_success", d.n_success}, {"n_failure", d.n_failure}, {"success_steps", d.success_steps}, {"minstd", d.minstd}, {"error", d.error}};
    for (auto& t : i32t) if (k == t.nm) { memcpy(out, t.p, 4 * n); return 0; }
    if (d.curriculum != 0) {
        if (k == "level") { memcpy(out, d.level, n); return 0; }
        if (k == "check_counter") { memcpy(out, d.check_counter, 4 * n); return 0; }
        if (k == "win_len") { memcpy(out, d.win_len, n * XW_N_T3); return 0; }
        if (k == "win_sum") { memcpy(out, d.win_sum, n * XW_N_T3); return 0; }
    }
    if (k == "goal_x" || k == "goal_y") {
        uint8_t* src = k == "goal_x" ? d.goal_x : d.goal_y;
        for (size_t g = 0; g < XW_MAX_GOALS; ++g) for (size_t e = 0; e < n; ++e) ((uint8_t*)out)[e * XW_MAX_GOALS + g] = src[g * n + e];
        return 0;
    }
    if (d.vr > 0 && k == "goal_yaw") {
        for (size_t g = 0; g < XW_MAX_GOALS; ++g) for (size_t e = 0; e < n; ++e) ((uint16_t*)out)[e * XW_MAX_GOALS + g] = d.goal_yaw[g * n + e];
        return 0;
    }
    if (d.vr > 0 && (k == "goal_scale" || k == "goal_offset")) {
        double* src = k == "goal_scale" ? d.goal_scale : d.goal_offset;
        for (size_t g = 0; g < XW_MAX_GOALS; ++g) for (size_t e = 0; e < n; ++e) ((double*)out)[e * XW_MAX_GOALS + g] = src[g * n + e];
        return 0;
    }
    if (k == "goal_icon" || k == "goal_name") {
        int32_t* src = k == "goal_icon" ? d.goal_icon : d.goal_name;
        for (size_t g = 0; g < XW_MAX_GOALS; ++g) for (size_t e = 0; e < n; ++e) ((int32_t*)out)[e * XW_MAX_GOALS + g] = src[g * n + e];
        return 0;
    }
    return -1;
}
}
