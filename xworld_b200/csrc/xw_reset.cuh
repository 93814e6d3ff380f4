// xw_reset.cuh -- episode start for one env: map generation + the teacher's first teach().
//
// Replaces, per env (reference file:line):
//   XWorld::reset                       games/xworld/xworld/xworld.cpp:109-151
//   XWorldEnv.reset/__instantiate_entities   games/xworld/maps/xworld_env.py:95-101,412-452
//   XWorldNav._configure (curriculum 0 and > 0: levels, padded worlds)  games/xworld/maps/XWorldNav.py:16-67
//   spanning_tree_maze_generator, bfs, flood_fill   python/maze2d.py:21-114
//   Teacher::reset_after_game_reset + teach -> TaskGroup::run_stage  teacher.cpp:207-237, teaching_task.cpp:204-222
//   the idle() stages of games/xworld3d/tasks/XWorld3DNav*.py and games/xworld/tasks/XWorldNav*.py
//
// One thread per env.  Occupancy is kept as 256-bit masks (cells <= 16x16) so "k-th free cell",
// neighbourhood and tile tests are popcount/bit operations rather than list scans.
#pragma once
#include "xw_common.cuh"

struct XwMask { uint64_t w[4]; };

XW_HD int xw_popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
XW_HD void m_zero(XwMask& m) { m.w[0] = m.w[1] = m.w[2] = m.w[3] = 0; }
XW_HD bool m_get(const XwMask& m, int c) { return (m.w[c >> 6] >> (c & 63)) & 1ull; }
XW_HD void m_set(XwMask& m, int c) { m.w[c >> 6] |= 1ull << (c & 63); }
XW_HD void m_clr(XwMask& m, int c) { m.w[c >> 6] &= ~(1ull << (c & 63)); }
XW_HD int m_count(const XwMask& m) { return xw_popc64(m.w[0]) + xw_popc64(m.w[1]) + xw_popc64(m.w[2]) + xw_popc64(m.w[3]); }
// index of the k-th set bit (row-major cell order); -1 if fewer
XW_HD int m_nth(const XwMask& m, int k) {
    for (int i = 0; i < 4; ++i) {
        int c = xw_popc64(m.w[i]);
        if (k < c) {
#if defined(__CUDA_ARCH__)
            uint32_t w = (uint32_t)m.w[i];
            int base = i * 64;
            const int cl = __popc(w);
            if (k >= cl) { k -= cl; w = (uint32_t)(m.w[i] >> 32); base += 32; }
            return base + (int)__fns(w, 0, k + 1);  // position of the (k+1)-th set bit
#else
            uint64_t v = m.w[i];
            for (int j = 0; j < k; ++j) v &= v - 1;  // drop k lowest set bits
            return i * 64 + __builtin_ctzll(v);
#endif
        }
        k -= c;
    }
    return -1;
}
// the cells 0 .. n-1
XW_HD void m_first(XwMask& m, int n) {
    for (int i = 0; i < 4; ++i) {
        const int r = n - 64 * i;
        m.w[i] = r >= 64 ? ~0ull : (r <= 0 ? 0ull : ((1ull << r) - 1));
    }
}

XW_HD int xw_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
// OR a row of at most 16 bits into the mask at bit offset `b` (it touches at most two words)
XW_HD void m_or_bits(XwMask& m, int b, uint64_t bits) {
    const int w = b >> 6, s = b & 63;
    m.w[w] |= bits << s;
    if (s > 48 && w < 3) m.w[w + 1] |= bits >> (64 - s);
}

struct XwMapCtx {
    int H, W;        // the world the Python side sees: the whole map, or the level's inner world (curriculum)
    int nG, nB;      // goals / blocks of this episode
    XwMask inrange;  // cells of the map
    XwMask block;    // wall bricks
    XwMask goal;     // goal cells
    int agent;       // agent cell or -1 while "deleted"
    int gcell[XW_MAX_GOALS];
    int gname[XW_MAX_GOALS];
    int gicon[XW_MAX_GOALS];
    // first-person view only: the agent's heading and the goals' poses (set_property, xworld_env.py:207-223)
    int facing;
    int gyaw[XW_MAX_GOALS];
    double gscale[XW_MAX_GOALS], goffset[XW_MAX_GOALS];
};

// IEEE double products / sums that must not be contracted into FMAs (the oracle, like OpenCV, rounds each one)
XW_HD double xw_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
XW_HD double xw_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

XW_HD bool ctx_free(const XwMapCtx& c, int x, int y) {  // (x,y,0) in env.available_grids
    if (x < 0 || y < 0 || x >= c.W || y >= c.H) return false;
    int cell = y * c.W + x;
    return !m_get(c.block, cell) && !m_get(c.goal, cell) && cell != c.agent;
}

// -------------------------------------------------------------------------- maze + entities
// Where the maze DFS keeps its stack and where its draws come from are policies: the one-thread-per-env paths (host build,
// full / masked reset) use a local array and evaluate Philox as they go; the per-step auto-reset launch (one warp per env)
// keeps the stack in shared memory, and for attempt 0 the warp's lanes evaluate all of the maze's Philox blocks side by
// side before lane 0 walks the maze (the draws are a pure function of the visit number, so only the walk is sequential).
struct XwStackLocal {
    uint32_t s[64];
    XW_HD uint32_t get(int i) const { return s[i]; }
    XW_HD void set(int i, uint32_t v) { s[i] = v; }
};
struct XwStackStrided {  // entry i of this lane's stack at p[i * 32] (lanes interleaved: no bank conflicts between lanes)
    uint32_t* p;
    XW_HD uint32_t get(int i) const { return p[i * 32]; }
    XW_HD void set(int i, uint32_t v) { p[i * 32] = v; }
};
struct XwMazeDrawsLazy {
    XwDrawSeq q;
    XW_HD uint32_t operator()(uint32_t i) { return xw_draw_next(q, i); }
};
struct XwMazeDrawsTable {  // t != NULL: the draws were evaluated up front (attempt 0 of the warp path); else as they come
    const uint32_t* t;
    XwDrawSeq q;
    XW_HD uint32_t operator()(uint32_t i) { return t ? t[i] : xw_draw_next(q, i); }
};
XW_HD int xw_maze_blocks(int D) {  // Philox blocks a D-sided maze consumes: 3 draws per node of its (nx x nx) graph
    const int X = (D % 2 == 0) ? D - 1 : D, nx = (X + 1) / 2;
    return (3 * nx * nx + 3) / 4;
}

// spanning_tree_maze_generator (python/maze2d.py:74-114): walls of a D x D maze as a bit mask, carved by an explicit-stack
// DFS with 3 draws per visited node (random.shuffle of its 4 moves).  A stack entry packs x | y << 4 | next << 8 | order << 12.
template <class Stack, class Draws>
XW_HD void xw_maze_walls(int D, Stack& st, Draws& draw, XwMask& wall) {
    m_zero(wall);
    const int pad = (D % 2 == 0), X = pad ? D - 1 : D, nx = (X + 1) / 2;
    {   // odd rows are all wall, even rows have wall at odd x (python/maze2d.py:80-86), a row at a time
        const uint64_t full = (1ull << X) - 1, odd = 0xAAAAull & full;
        for (int y = 0; y < X; ++y) m_or_bits(wall, y * D, (y & 1) ? full : odd);
    }
    uint64_t visited = 1ull;
    uint32_t visit_no = 0;
    int sp = 0;
    auto push = [&](int px, int py) {
        uint32_t ord = 0xE4u;  // moves 0 (-1,0), 1 (1,0), 2 (0,1), 3 (0,-1) in two-bit fields: identity order
        for (int i = 3; i >= 1; --i) {  // random.shuffle: Fisher-Yates from the end
            const uint32_t j = xw_randbelow(draw(visit_no * 3 + (uint32_t)(3 - i)), (uint32_t)i + 1);
            const uint32_t a = (ord >> (2 * i)) & 3u, b = (ord >> (2 * j)) & 3u;
            ord = (ord & ~((3u << (2 * i)) | (3u << (2 * j)))) | (b << (2 * i)) | (a << (2 * j));
        }
        st.set(sp, (uint32_t)px | ((uint32_t)py << 4) | (ord << 12));
        ++visit_no; ++sp;
    };
    push(0, 0);
    while (sp > 0) {
        const uint32_t top = st.get(sp - 1);
        const uint32_t next = (top >> 8) & 7u;
        if (next == 4u) { --sp; continue; }
        st.set(sp - 1, top + 0x100u);
        const uint32_t m = (top >> (12 + 2 * next)) & 3u;
        const int cx = (int)(top & 15u), cy = (int)((top >> 4) & 15u);
        const int qx = cx + (int)((0x1120u >> (4 * m)) & 3u) - 1, qy = cy + (int)((0x0211u >> (4 * m)) & 3u) - 1;
        if ((unsigned)qx < (unsigned)nx && (unsigned)qy < (unsigned)nx && !((visited >> (qy * nx + qx)) & 1ull)) {
            m_clr(wall, (cy + qy) * D + (cx + qx));  // the edge's mid-point (maze2d.py:105-108)
            visited |= 1ull << (qy * nx + qx);
            push(qx, qy);
        }
    }
    if (pad) {
        for (int i = 0; i < X; ++i) if (i & 1) m_set(wall, X * D + i);
        for (int i = 0; i < D; ++i) if (i & 1) m_set(wall, i * D + X);
    }
}

// Returns 0 on success.
template <class Stack, class Draws>
XW_HD int xw_gen_map_t(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int level, XwMapCtx& c, Stack& st, Draws& mz) {
    int D = d.H, nG = d.G, nB = d.n_blocks;
    if (d.curriculum != 0) xw_level_dims(level, D, nG, nB);  // set_dims(current_dim, current_dim), XWorldNav.py:58
    c.H = D; c.W = D; c.nG = nG; c.nB = nB;
    m_first(c.inrange, D * D); m_zero(c.block); m_zero(c.goal);
    // ---- goal names: shuffle(goal_names) then pop() per goal == partial Fisher-Yates from the end;
    //      the touched positions are kept in a tiny sparse map instead of an n_names array
    {
        int keys[2 * XW_MAX_GOALS], vals[2 * XW_MAX_GOALS], cnt = 0;
        const int n = d.n_names;
        XwDrawSeq nq = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_NAMES);
        for (int k = 0; k < nG; ++k) {
            int i = n - 1 - k, vi = i;
            for (int q = 0; q < cnt; ++q) if (keys[q] == i) vi = vals[q];
            if (i >= 1) {
                int j = (int)xw_randbelow(xw_draw_next(nq, (uint32_t)k), (uint32_t)i + 1);
                int vj = j, qj = -1;
                for (int q = 0; q < cnt; ++q) if (keys[q] == j) { vj = vals[q]; qj = q; }
                // a[i] <-> a[j]; position i is final, only a[j] needs remembering
                if (qj >= 0) vals[qj] = vi; else { keys[cnt] = j; vals[cnt] = vi; ++cnt; }
                vi = vj;
            }
            c.gname[k] = vi;
        }
    }
    // ---- maze
    XwMask wall;
    xw_maze_walls(D, st, mz, wall);
    const int nb = m_count(wall);
    if (nb < nB) return 1;
    // ---- goals (loc + icon variant), in creation order
    XwMask avail;
    for (int i = 0; i < 4; ++i) avail.w[i] = c.inrange.w[i] & ~wall.w[i];
    XwDrawSeq lq = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_GOAL_LOC), aq = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_GOAL_ASSET);
    for (int k = 0; k < nG; ++k) {
        int nf = m_count(avail);
        if (nf == 0) return 1;
        int cell = m_nth(avail, (int)xw_randbelow(xw_draw_next(lq, (uint32_t)k), (uint32_t)nf));
        m_clr(avail, cell); m_set(c.goal, cell);
        c.gcell[k] = cell;
        int f = d.name_first[c.gname[k]], nv = d.name_first[c.gname[k] + 1] - f;
        c.gicon[k] = d.name_icons[f + (int)xw_randbelow(xw_draw_next(aq, (uint32_t)k), (uint32_t)nv)];
    }
    // ---- blocks: shuffle(blocks) + pop() per block entity == partial Fisher-Yates over the wall list
    {
        XwDrawSeq bq = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_BLOCKS);
        uint8_t bl[XW_MAX_DIM * XW_MAX_DIM / 2 + 8];
        int n = 0;
        for (int w = 0; w < 4; ++w)
            for (uint64_t v = wall.w[w]; v; v &= v - 1) bl[n++] = (uint8_t)(w * 64 + xw_ctz64(v));
        for (int k = 0; k < nB; ++k) {
            int i = nb - 1 - k;
            if (i >= 1) {
                int j = (int)xw_randbelow(xw_draw_next(bq, (uint32_t)k), (uint32_t)i + 1);
                uint8_t t = bl[i]; bl[i] = bl[j]; bl[j] = t;
            }
            m_set(c.block, bl[i]);
        }
    }
    // ---- agent
    {
        int nf = m_count(avail);
        if (nf == 0) return 1;
        c.agent = m_nth(avail, (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_AGENT_LOC, 0), (uint32_t)nf));
    }
    c.facing = 1;  // Entity default yaw 1.5707963 == "down" (xworld_env.py:42, xitem.cpp:65-78)
    if (d.vr > 0) {  // "if partially observed, perturb the objects" (xworld_env.py:207-223)
        // agent: yaw = choice(range(-1, 3)) * PI_2: -PI_2 up, 0 right, PI_2 down, 2 PI_2 left
        c.facing = ((int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_AGENT_YAW, 0), 4u) + 3) & 3;
        for (int k = 0; k < nG; ++k) {
            // random.uniform(a, b) = a + (b - a) * random(); random() = draw * 2^-32 (yaw: on the XW_YAW_STEPS grid)
            const XwDraw4 u = xw_draw_block(d.seed, gid, ep, att, XW_SITE_GOAL_POSE, (uint32_t)k);
            c.gyaw[k] = (int)(u.v0 >> 20);
            const double scale = xw_dadd(0.5, xw_dmul(0.5, (double)u.v1 * (1.0 / 4294967296.0)));
            c.gscale[k] = scale;
            c.goffset[k] = xw_dmul(1.0 - scale, (double)u.v2 * (1.0 / 4294967296.0));
        }
    }
    return 0;
}

XW_HD int xw_gen_map(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int level, XwMapCtx& c) {
    XwStackLocal st;
    XwMazeDrawsLazy mz;
    mz.q = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_MAZE);
    return xw_gen_map_t(d, gid, ep, att, level, c, st, mz);
}

// -------------------------------------------------------------------------- BFS
// Breadth-first discovery from `seed`; obstacles = `obst` (the cell `pass`, if >= 0, is always
// enterable).  order[] receives the discovered cells (seed excluded) in maze2d.flood_fill order
// (moves (-1,0),(1,0),(0,-1),(0,1)).  Returns their count; stops early when `stop_at` is found (-2).
// order[] entries are packed (y << 4 | x) (no division per dequeued cell); xw_bfs_cell turns one into a cell index.
XW_HD int xw_bfs_cell(const XwMapCtx& c, uint8_t packed) { return (packed >> 4) * c.W + (packed & 15); }
XW_HD int xw_bfs(const XwMapCtx& c, const XwMask& obst, int seed, int pass, int stop_at, uint8_t* order) {
    XwMask closed;  // visited or not enterable
    for (int i = 0; i < 4; ++i) closed.w[i] = obst.w[i] | ~c.inrange.w[i];
    if (pass >= 0) m_clr(closed, pass);
    m_set(closed, seed);
    const int sx = seed % c.W, sy = seed / c.W;
    int head = -1, n = 0;  // queue = seed, order[0..n)
    while (head < n) {
        const int cx = head < 0 ? sx : (order[head] & 15), cy = head < 0 ? sy : (order[head] >> 4);
        ++head;
        const int cur = cy * c.W + cx;
        if (cur == stop_at) return -2;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int x = cx + (m == 0 ? -1 : m == 1 ? 1 : 0), y = cy + (m == 2 ? -1 : m == 3 ? 1 : 0);
            if (x < 0 || y < 0 || x >= c.W || y >= c.H) continue;
            const int q = y * c.W + x;
            if (m_get(closed, q)) continue;
            m_set(closed, q);
            order[n++] = (uint8_t)((y << 4) | x);
        }
    }
    return n;
}

// 256-bit mask helpers: value of `m` at the neighbour cell.  nf / nl = cells with x != 0 / x != W-1.
struct XwCols { XwMask nf, nl; };
XW_HD XwCols xw_cols(const XwMapCtx& c) {
    XwCols k; m_zero(k.nf); m_zero(k.nl);
    const int D = c.W;
    const uint64_t row_nf = ((1ull << D) - 1) & ~1ull, row_nl = (1ull << (D - 1)) - 1;  // a row is at most 16 bits
    for (int y = 0; y < c.H; ++y) {
        const int b = y * D, w = b >> 6, sft = b & 63;
        k.nf.w[w] |= row_nf << sft; k.nl.w[w] |= row_nl << sft;
        if (sft + D > 64 && w < 3) { k.nf.w[w + 1] |= row_nf >> (64 - sft); k.nl.w[w + 1] |= row_nl >> (64 - sft); }
    }
    return k;
}
XW_HD XwMask m_shr(const XwMask& m, int s) {  // result(c) = m(c + s), 0 < s < 64
    XwMask r;
    r.w[0] = (m.w[0] >> s) | (m.w[1] << (64 - s)); r.w[1] = (m.w[1] >> s) | (m.w[2] << (64 - s));
    r.w[2] = (m.w[2] >> s) | (m.w[3] << (64 - s)); r.w[3] = m.w[3] >> s;
    return r;
}
XW_HD XwMask m_shl(const XwMask& m, int s) {  // result(c) = m(c - s)
    XwMask r;
    r.w[3] = (m.w[3] << s) | (m.w[2] >> (64 - s)); r.w[2] = (m.w[2] << s) | (m.w[1] >> (64 - s));
    r.w[1] = (m.w[1] << s) | (m.w[0] >> (64 - s)); r.w[0] = m.w[0] << s;
    return r;
}
XW_HD XwMask m_and(const XwMask& a, const XwMask& b) { XwMask r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] & b.w[i]; return r; }
XW_HD XwMask m_or(const XwMask& a, const XwMask& b) { XwMask r; for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] | b.w[i]; return r; }
XW_HD XwMask m_right(const XwMask& m, const XwCols& k) { return m_and(m_shr(m, 1), k.nl); }            // m at (x+1, y)
XW_HD XwMask m_left(const XwMask& m, const XwCols& k) { return m_and(m_shl(m, 1), k.nf); }             // m at (x-1, y)
XW_HD XwMask m_below(const XwMask& m, int W) { return m_shr(m, W); }                                    // m at (x, y+1)
XW_HD XwMask m_above(const XwMask& m, int W) { return m_shl(m, W); }                                    // m at (x, y-1)

// Set of cells reachable from `seed` through cells outside `obst` (4-neighbourhood), seed included: the same set
// maze2d.flood_fill / bfs discover (python/maze2d.py:21-63), grown a whole frontier at a time with 256-bit shifts
// instead of a queue.  Used where only membership matters (which goals can be reached), not discovery order.
XW_HD XwMask xw_flood(const XwMapCtx& c, const XwMask& obst, int seed) {
    const int D = c.W;
    XwMask free_, reach;
    m_zero(reach);
    for (int i = 0; i < 4; ++i) free_.w[i] = c.inrange.w[i] & ~obst.w[i];
    const XwCols k = xw_cols(c);
    const XwMask& not_first = k.nf;
    const XwMask& not_last = k.nl;
    m_set(reach, seed);
    m_set(free_, seed);
    while (true) {
        XwMask nx;
        uint64_t changed = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint64_t lo = i > 0 ? reach.w[i - 1] : 0, hi = i < 3 ? reach.w[i + 1] : 0, me = reach.w[i];
            const uint64_t right = ((me << 1) | (lo >> 63)) & not_first.w[i];          // x + 1
            const uint64_t left = ((me >> 1) | (hi << 63)) & not_last.w[i];            // x - 1
            const uint64_t down = (me << D) | (lo >> (64 - D));                         // y + 1
            const uint64_t up = (me >> D) | (hi << (64 - D));                           // y - 1
            nx.w[i] = (me | right | left | down | up) & free_.w[i];
            changed |= nx.w[i] ^ me;
        }
        if (!changed) break;
        reach = nx;
    }
    return reach;
}
// is any 4-neighbour of `cell` (or the cell itself) in `set`?  == "a walk that ends by entering `cell` exists"
XW_HD bool xw_touches(const XwMapCtx& c, const XwMask& set, int cell) {
    const int x = cell % c.W, y = cell / c.W;
    if (m_get(set, cell)) return true;
    if (x > 0 && m_get(set, cell - 1)) return true;
    if (x + 1 < c.W && m_get(set, cell + 1)) return true;
    if (y > 0 && m_get(set, cell - c.W)) return true;
    if (y + 1 < c.H && m_get(set, cell + c.W)) return true;
    return false;
}

// -------------------------------------------------------------------------- tiles
XW_HD int n_free4(const XwMapCtx& c, int x, int y, int excl) {
    int n = 0;
    if (ctx_free(c, x, y - 1) && (y - 1) * c.W + x != excl) ++n;
    if (ctx_free(c, x - 1, y) && y * c.W + x - 1 != excl) ++n;
    if (ctx_free(c, x + 1, y) && y * c.W + x + 1 != excl) ++n;
    if (ctx_free(c, x, y + 1) && (y + 1) * c.W + x != excl) ++n;
    return n;
}

// The reference's p / t / l tiles (xworld3d_task.py:226-251, 253-276, 302-322) in its own order: cells row-major, and
// inside a cell the order of the Python's appends.  Each append kind ("slot") is a 256-bit mask of the cells where
// it happens, built from the free-cell mask with shifts, so counting is popcounts and the pick-th pair is found by
// walking only the cells that emit -- the scalar enumeration was two passes of ~15 free-cell tests per cell, the
// longest stretch of a reset.  Returns the count; pick >= 0 also returns the pick-th pair in (a, b).
XW_HD int xw_tiles(const XwMapCtx& c, int kind, int pick, int& a, int& b) {
    const int W = c.W;
    const XwCols k = xw_cols(c);
    XwMask F;
    for (int i = 0; i < 4; ++i) F.w[i] = c.inrange.w[i] & ~c.block.w[i] & ~c.goal.w[i];
    if (c.agent >= 0) m_clr(F, c.agent);
    const XwMask fR = m_right(F, k), fL = m_left(F, k), fD = m_below(F, W), fU = m_above(F, W);
    XwMask slot[6];
    int da[6], db[6], ns;  // the pair a slot emits at cell q: (q + da, q + db)
    if (kind == XW_T3_NEAR) {  // for k in ((x+1,y), (x,y+1), (x+1,y+1)): (p1,p2) if p2 has another free side, then (p2,p1)
        const XwMask any4 = m_or(m_or(fU, fD), m_or(fL, fR));
        const XwMask b0 = m_and(F, fR), b1 = m_and(F, fD), b2 = m_and(F, m_right(fD, k));
        slot[0] = m_and(b0, m_right(m_or(m_or(fU, fD), fR), k)); slot[1] = m_and(b0, m_or(m_or(fU, fD), fL));
        slot[2] = m_and(b1, m_below(m_or(m_or(fL, fR), fD), W)); slot[3] = m_and(b1, m_or(m_or(fL, fR), fU));
        slot[4] = m_and(b2, m_right(m_below(any4, W), k));       slot[5] = m_and(b2, any4);
        da[0] = 0; db[0] = 1;      da[1] = 1; db[1] = 0;
        da[2] = 0; db[2] = W;      da[3] = W; db[3] = 0;
        da[4] = 0; db[4] = W + 1;  da[5] = W + 1; db[5] = 0;
        ns = 6;
    } else if (kind == XW_T3_BETWEEN) {  // the free cell between two free cells, with a free cell on one of the other sides
        slot[0] = m_and(m_and(F, m_and(fL, fR)), m_or(fU, fD)); da[0] = -1; db[0] = 1;
        slot[1] = m_and(m_and(F, m_and(fU, fD)), m_or(fL, fR)); da[1] = -W; db[1] = W;
        ns = 2;
    } else {  // three free cells in a column, then in a row: both adjacent pairs of each
        const XwMask v = m_and(m_and(F, fD), m_below(fD, W)), h = m_and(m_and(F, fR), m_right(fR, k));
        slot[0] = v; da[0] = 0; db[0] = W;  slot[1] = v; da[1] = W; db[1] = 2 * W;
        slot[2] = h; da[2] = 0; db[2] = 1;  slot[3] = h; da[3] = 1; db[3] = 2;
        ns = 4;
    }
    int total = 0;
    for (int j = 0; j < ns; ++j) total += m_count(slot[j]);
    if (pick < 0 || pick >= total) return total;
    for (int w = 0; w < 4; ++w) {
        int cw = 0;
        uint64_t any = 0;
        for (int j = 0; j < ns; ++j) { cw += xw_popc64(slot[j].w[w]); any |= slot[j].w[w]; }
        if (pick >= cw) { pick -= cw; continue; }
        while (any) {
#if defined(__CUDA_ARCH__)
            const int bit = __ffsll((long long)any) - 1;
#else
            const int bit = __builtin_ctzll(any);
#endif
            any &= any - 1;
            for (int j = 0; j < ns; ++j)
                if ((slot[j].w[w] >> bit) & 1ull) {
                    if (pick == 0) { const int q = w * 64 + bit; a = q + da[j]; b = q + db[j]; return total; }
                    --pick;
                }
        }
    }
    return total;
}

// direction of `r` seen from `t` when looking along the unit axis vector (vx,vy):
// XWorld3DNavTargetDirection.__compute_triple_direction (:98-126) collapsed to integers for
// axis-aligned headings and adjacent cells (2-D world: sign>0 => "right").
XW_HD int xw_axis_direction(int vx, int vy, int dx, int dy) {
    if (dx == 0 && dy == 0) return XW_DIR_FALSE;
    if (dx == vx && dy == vy) return XW_DIR_FRONT;
    if (dx == -vx && dy == -vy) return XW_DIR_BEHIND;
    return (vy * dx - vx * dy) > 0 ? XW_DIR_RIGHT : XW_DIR_LEFT;
}

struct XwTaskOut { int tmask, aux0, aux1, aux2; };

// idle() of the five XWorld3DNav* tasks.  false = the reference's `assert ..., "map too crowded?"`.
XW_HD bool xw_idle3d(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int task, XwMapCtx& c, XwTaskOut& o) {
    o.tmask = o.aux0 = o.aux1 = o.aux2 = 0;
    const int G = c.nG;
    XwMask obst;
    for (int i = 0; i < 4; ++i) obst.w[i] = c.block.w[i] | c.goal.w[i];
    if (task == XW_T3_TARGET || task == XW_T3_AVOID) {
        // _reachable(agent.loc, g.loc) per goal (XWorld3DNavTarget.py:31): goals and blocks are obstacles, the end cell
        // is enterable -> one flood from the agent, then "does the flood touch the goal"
        int cand = 0, nc = 0;
        const XwMask reach = xw_flood(c, obst, c.agent);
        for (int g = 0; g < G; ++g)
            if (xw_touches(c, reach, c.gcell[g])) { cand |= 1 << g; ++nc; }
        if (nc == 0) return false;
        int k = (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_A, 0), (uint32_t)nc);
        int sel = 0;
        for (int g = 0; g < G; ++g) if (cand & (1 << g)) { if (k == 0) { sel = g; break; } --k; }
        if (task == XW_T3_TARGET) {
            for (int g = 0; g < G; ++g) if (c.gname[g] == c.gname[sel]) o.tmask |= 1 << g;
            o.aux0 = sel;
        } else {
            int refs = 0, nr = 0;
            for (int g = 0; g < G; ++g) if (c.gname[g] != c.gname[sel]) { refs |= 1 << g; ++nr; }
            if (nr == 0) return false;
            int kr = (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_B, 0), (uint32_t)nr);
            int ref = 0;
            for (int g = 0; g < G; ++g) if (refs & (1 << g)) { if (kr == 0) { ref = g; break; } --kr; }
            for (int g = 0; g < G; ++g) if (c.gname[g] != c.gname[ref]) o.tmask |= 1 << g;
            o.aux0 = ref;
        }
        return true;
    }
    if (G < 2) return false;
    // random.shuffle(goals); g1, g2 = goals[:2]
    int perm[XW_MAX_GOALS];
    for (int i = 0; i < G; ++i) perm[i] = i;
    for (int i = G - 1; i >= 1; --i) {
        int j = (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_SHUF, (uint32_t)(G - 1 - i)), (uint32_t)i + 1);
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    const int g1 = perm[0], g2 = perm[1];
    // delete agent, g1, g2
    c.agent = -1;
    m_clr(c.goal, c.gcell[g1]); m_clr(c.goal, c.gcell[g2]);
    int a = 0, b = 0;
    int nt = xw_tiles(c, task, -1, a, b);
    if (nt == 0) return false;
    xw_tiles(c, task, (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_A, 0), (uint32_t)nt), a, b);
    c.gcell[g1] = a; c.gcell[g2] = b;
    m_set(c.goal, a); m_set(c.goal, b);
    for (int i = 0; i < 4; ++i) obst.w[i] = c.block.w[i] | c.goal.w[i];
    const int W = c.W;
    // `agent.loc = random.choice(new_a)`: uniform over the cells _propagate_agent finds (xworld3d_task.py:344-355).  The draw
    // indexes that SET in row-major order (the reference's list order, BFS discovery, is an artefact its unseeded draw makes
    // unobservable): one mask flood + the k-th set bit instead of a queue BFS.
    if (task == XW_T3_NEAR) {
        XwMask reach = xw_flood(c, obst, b);
        m_clr(reach, b);  // flood_fill excludes its seed
        const int nf = m_count(reach);
        if (nf == 0) return false;
        c.agent = m_nth(reach, (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_AGENT, 0), (uint32_t)nf));
        // goals within 1.5 (+1e-3) of g1, g1 itself excluded: the 8-neighbourhood
        for (int g = 0; g < G; ++g) {
            int dx = c.gcell[g] % W - a % W, dy = c.gcell[g] / W - a / W;
            if ((dx | dy) != 0 && dx * dx + dy * dy <= 2) o.tmask |= 1 << g;
        }
        o.aux0 = g1;
    } else if (task == XW_T3_BETWEEN) {
        int mx = (a % W + b % W) / 2, my = (a / W + b / W) / 2;
        int mid = my * W + mx;
        XwMask reach = xw_flood(c, obst, mid);
        m_clr(reach, mid);
        const int nf = m_count(reach);
        if (nf == 0) return false;
        c.agent = m_nth(reach, (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_AGENT, 0), (uint32_t)nf));
        o.aux0 = g1 | (g2 << 4); o.aux1 = mx; o.aux2 = my;  // (g2: only the sentence channel needs it, xw_sentence.hpp)
    } else {
        int target = g1, referent = g2;
        int tx = a % W, ty = a / W;
        int ne = n_free4(c, tx, ty, -1);
        if (ne == 0) {
            tx = b % W; ty = b / W;
            ne = n_free4(c, tx, ty, -1);
            if (ne == 0) return false;
            target = g2; referent = g1;
        }
        int k = (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_B, 0), (uint32_t)ne);
        // k-th free 4-neighbour in row-major order: up, left, right, down
        int ex = 0, ey = 0;
        const int NX[4] = {0, -1, 1, 0}, NY[4] = {-1, 0, 0, 1};
        for (int m = 0; m < 4; ++m)
            if (ctx_free(c, tx + NX[m], ty + NY[m])) { if (k == 0) { ex = tx + NX[m]; ey = ty + NY[m]; break; } --k; }
        int rcell = c.gcell[referent];
        int dir = xw_axis_direction(tx - ex, ty - ey, rcell % W - tx, rcell / W - ty);
        int ecell = ey * W + ex;
        const XwMask reach = xw_flood(c, obst, ecell);  // _propagate_agent([e], inclusive=True): the seed counts
        const int nf = m_count(reach);
        c.agent = m_nth(reach, (int)xw_randbelow(xw_draw(d.seed, gid, ep, att, XW_SITE_TASK_AGENT, 0), (uint32_t)nf));
        o.aux0 = referent; o.aux1 = dir; o.aux2 = target;
    }
    return true;
}

// idle() of XWorldNav{Target,Near,ColorTarget,Between} (walls.json).  In the reference commit NavNear
// and NavBetween never leave idle (they pass 2-tuples to a 3-tuple BFS, XWorldNavNear.py:13-16,
// XWorldNavBetween.py:11-13, maze2d.py:49-63); NavTarget/NavColorTarget start iff a (coloured) goal
// is reachable with only blocks as obstacles.  reach/colour masks are per-episode constants.
XW_HD void xw_idle2d(const XwDev& d, int64_t gid, uint32_t ep, uint32_t step_no, uint32_t& minstd, int reach_mask,
                     int color_mask, int& task, int& stage, int& tmask, int& aux0, int32_t& steps_in_task) {
    task = xw_get_rand_ind(minstd, 4);
    steps_in_task = 0;
    tmask = 0; aux0 = 0;
    if (task == XW_T2_TARGET || task == XW_T2_COLOR_TARGET) {
        int cand = reach_mask & (task == XW_T2_COLOR_TARGET ? color_mask : 0xff);
        int nc = xw_popc64((uint64_t)cand);
        if (nc > 0) {
            int k = (int)xw_randbelow(xw_draw(d.seed, gid, ep, 0, XW_SITE_TASK_A, step_no), (uint32_t)nc);
            int sel = 0;
            for (int g = 0; g < XW_MAX_GOALS; ++g) if (cand & (1 << g)) { if (k == 0) { sel = g; break; } --k; }
            aux0 = sel; tmask = 1 << sel;
            stage = XW_STAGE_NAVIGATION;
        }
    }
}

// One attempt at a new episode: map + (navigation2d.json) the task's idle() stage.  A pure function of
// (env id, episode, attempt): 1 = done, 0 = the reference's `assert ..., "map too crowded?"` -> next attempt,
// 2 = the map itself cannot be built (configuration error).
XW_HD int xw_reset_attempt(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int task, int level, XwMapCtx& c, XwTaskOut& o) {
    o.tmask = o.aux0 = o.aux1 = o.aux2 = 0;
    if (xw_gen_map(d, gid, ep, att, level, c)) return 2;
    if (d.rules != XW_RULES_NAV3D) return 1;
    return xw_idle3d(d, gid, ep, att, task, c, o) ? 1 : 0;
}
template <class Stack, class Draws>
XW_HD int xw_reset_attempt_t(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int task, int level, XwMapCtx& c, XwTaskOut& o,
                             Stack& st, Draws& mz) {
    o.tmask = o.aux0 = o.aux1 = o.aux2 = 0;
    if (xw_gen_map_t(d, gid, ep, att, level, c, st, mz)) return 2;
    if (d.rules != XW_RULES_NAV3D) return 1;
    return xw_idle3d(d, gid, ep, att, task, c, o) ? 1 : 0;
}

// Write the episode a successful attempt produced into the env's state.
XW_HD void xw_reset_commit(const XwDev& d, int e, uint32_t ep, uint32_t minstd, int task, XwMapCtx& c, XwTaskOut o) {
    const int n = d.n;
    const int64_t gid = d.gid0 + e;
    int stage = d.rules == XW_RULES_NAV3D ? XW_STAGE_NAVIGATION : XW_STAGE_IDLE;
    uint8_t* g = d.grid + (size_t)e * d.CS;
    // cpp_get_entities (xworld_env.py:352-366): the world sits at (offset_w, offset_h) = ((max - dim) / 2, same) of
    // the map, everything around it is brick (__padding_walls, :454-473); no offset when the world is the map
    const int D = c.W, off = (d.W - D) / 2;
    {   // rows are 16-byte multiples (CS), 16-byte aligned: clear with 8-byte stores
        uint64_t* g8 = (uint64_t*)g;
        for (int i = 0; i < (d.CS >> 3); ++i) g8[i] = 0;  // XW_CELL_EMPTY == 0
    }
    if (D != d.W)
        for (int y = 0; y < d.H; ++y)
            for (int x = 0; x < d.W; ++x)
                if (x < off || x >= off + D || y < off || y >= off + D) g[y * d.W + x] = XW_CELL_BLOCK;
    for (int w = 0; w < 4; ++w)
        for (uint64_t v = c.block.w[w]; v; v &= v - 1) {
            const int i = w * 64 + xw_ctz64(v);
            g[D == d.W ? i : (i / D + off) * d.W + i % D + off] = XW_CELL_BLOCK;
        }
    for (int k = 0; k < d.G; ++k) {
        const bool on = k < c.nG;  // levels 0-2 hold two goals (XWorldNav.py:31): the other slots read 0
        const int gx = on ? c.gcell[k] % D + off : 0, gy = on ? c.gcell[k] / D + off : 0;
        if (on) g[gy * d.W + gx] = (uint8_t)(XW_CELL_GOAL0 + k);
        d.goal_x[(size_t)k * n + e] = (uint8_t)gx;
        d.goal_y[(size_t)k * n + e] = (uint8_t)gy;
        d.goal_icon[(size_t)k * n + e] = on ? c.gicon[k] : 0;
        d.goal_name[(size_t)k * n + e] = on ? c.gname[k] : 0;
        if (d.vr > 0) {
            d.goal_yaw[(size_t)k * n + e] = (uint16_t)(on ? c.gyaw[k] : XW_YAW_STEPS / 4);
            d.goal_scale[(size_t)k * n + e] = on ? c.gscale[k] : 1.0;
            d.goal_offset[(size_t)k * n + e] = on ? c.goffset[k] : 0.0;
        }
    }
    const int agx = c.agent % D + off, agy = c.agent / D + off;
    g[agy * d.W + agx] = XW_CELL_AGENT;
    d.agent_x[e] = (uint8_t)agx;
    d.agent_y[e] = (uint8_t)agy;
    if (task == XW_T3_BETWEEN && d.rules == XW_RULES_NAV3D) { o.aux1 += off; o.aux2 += off; }  // the middle cell, in map coordinates
    d.facing[e] = (uint8_t)c.facing;
    int32_t steps_in_task = 0;
    if (d.rules == XW_RULES_NAV2D) {
        // per-episode constants for the 2-D idle stages: goals reachable through non-block cells,
        // goals whose icon has a colour (properties.txt)
        int reach = 0, colored = 0;
        const XwMask rset = xw_flood(c, c.block, c.agent);  // only blocks are obstacles here: goals are walked over
        for (int k = 0; k < c.nG; ++k) {
            if (m_get(rset, c.gcell[k])) reach |= 1 << k;
            if (d.icon_colored[c.gicon[k]]) colored |= 1 << k;
        }
        o.aux1 = reach; o.aux2 = colored;
        // first teach(): nav group idle stage, then the XWorldRec group's engine draw
        xw_idle2d(d, gid, ep, 0, minstd, reach, colored, task, stage, o.tmask, o.aux0, steps_in_task);
        xw_minstd_next(minstd);
    }
    d.task[e] = (uint8_t)task; d.stage[e] = (uint8_t)stage; d.event[e] = XW_EVENT_NONE; d.succ[e] = 0;
    d.tmask[e] = (uint8_t)o.tmask; d.aux0[e] = (uint8_t)o.aux0; d.aux1[e] = (uint8_t)o.aux1; d.aux2[e] = (uint8_t)o.aux2;
    d.steps_in_task[e] = steps_in_task;
    d.num_steps[e] = 0;
    d.minstd[e] = minstd;
    d.error[e] = 0;  // a new game: the invalid-action flag of the old one is gone
    if (d.ctx_flag) d.ctx_flag[e] = 2;  // init_screen: the context of a new game starts zero-filled (simulator.cpp:110-113)
}

// XWorldNav._configure, curriculum != 0 (XWorldNav.py:40-56) + XWorldEnv.get_current_usage (xworld_env.py:103-110):
// every reset counts; once `check_period` resets have passed and some task class has a record, usage = the lowest
// success rate over the task classes' windows, and the level moves up when it reaches the threshold.  Runs once
// per episode (a re-drawn "too crowded" map keeps the level).  Returns the level of the new episode.
XW_HD int xw_curriculum_level(const XwDev& d, int e) {
    if (d.curriculum == 0) return 0;
    int level = d.level[e];
    int cnt = d.check_counter[e] + 1;
    double usage = 0;
    if (cnt >= d.check_period) {
        bool any = false;
        for (int t = 0; t < XW_N_T3; ++t) {
            const int len = d.win_len[(size_t)e * XW_N_T3 + t];
            if (len == 0) continue;  // current_usage only holds the classes that recorded a result
            const double u = (double)d.win_sum[(size_t)e * XW_N_T3 + t] / (double)len;
            if (!any || u < usage) usage = u;
            any = true;
        }
        if (any) cnt = 0; else usage = 0;
    }
    d.check_counter[e] = cnt;
    if (usage >= d.curriculum && level < XW_N_LEVELS - 1) d.level[e] = (uint8_t)++level;
    return level;
}

// XWorld3DTask.__record_result (xworld3d_task.py:129-133): append to the task class's success_seq, keep the last 200.
XW_HD void xw_record_result(const XwDev& d, int e, int task, int res) {
    if (d.curriculum == 0) return;  // the windows feed nothing but the curriculum
    const size_t k = (size_t)e * XW_N_T3 + task;
    int len = d.win_len[k], pos = d.win_pos[k], sum = d.win_sum[k];
    uint32_t* w = d.win_bits + k * XW_WIN_WORDS;
    uint32_t word = w[pos >> 5];
    if (len == XW_WIN_SIZE) sum -= (int)((word >> (pos & 31)) & 1u); else ++len;
    w[pos >> 5] = (word & ~(1u << (pos & 31))) | ((uint32_t)res << (pos & 31));
    sum += res;
    pos = pos + 1 == XW_WIN_SIZE ? 0 : pos + 1;
    d.win_len[k] = (uint8_t)len; d.win_pos[k] = (uint8_t)pos; d.win_sum[k] = (uint8_t)sum;
}

// SimulatorInterface::reset_game for env `e` (simulator_interface.cpp:95-105), one thread: attempts in order.
XW_HD void xw_reset_env(const XwDev& d, int e) {
    const int64_t gid = d.gid0 + e;
    const uint32_t ep = (uint32_t)(d.episode[e] + 1);
    d.episode[e] = (int32_t)ep;
    uint32_t minstd = d.minstd[e];
    const bool nav3d = d.rules == XW_RULES_NAV3D;
    const int task = nav3d ? xw_get_rand_ind(minstd, 5) : 0;  // TaskGroup::run_stage, schedule "random"
    const int level = xw_curriculum_level(d, e);
    XwMapCtx c;
    XwTaskOut o;
    int st = 0;
    for (uint32_t att = 0; att < (nav3d ? 64u : 1u) && st == 0; ++att) st = xw_reset_attempt(d, gid, ep, att, task, level, c, o);
    if (st != 1) { d.error[e] = XW_ERR_INVALID_ARG; return; }
    xw_reset_commit(d, e, ep, minstd, task, c, o);
}

#if defined(__CUDACC__)
// The same for one WARP per env.  Attempt 0 runs on lane 0 alone (it succeeds for all but a few per cent of the
// episodes, and 32 different attempts in one warp diverge: measured 1.4x slower when every episode started that
// way).  Only when it asks for a retry do the lanes evaluate attempts 1..32, then 33..63, side by side: attempts are
// independent pure functions of (env, episode, attempt), so taking the lowest-numbered lane that did not ask for a
// retry is exactly the sequential rule -- and the rare env that needs dozens of attempts no longer sets the duration
// of the whole per-step reset launch (SMs were 14 % busy at C4).
// one copy of an attempt's code for both call sites of the warp path
__device__ __noinline__ int xw_reset_attempt_w(const XwDev& d, int64_t gid, uint32_t ep, uint32_t att, int task, int level, XwMapCtx& c,
                                               XwTaskOut& o, XwStackStrided& st, XwMazeDrawsTable& mz) {
    return xw_reset_attempt_t(d, gid, ep, att, task, level, c, o, st, mz);
}

// Shared memory of one warp of the auto-reset launch: the DFS stacks of its 32 lanes (interleaved) and attempt 0's maze draws.
#define XW_RESET_STACK_WORDS (64 * 32)
#define XW_RESET_DRAW_WORDS 192
__device__ __forceinline__ void xw_reset_env_warp(const XwDev& d, int e, uint32_t* s_stack, uint32_t* s_draws) {
    const int lane = threadIdx.x & 31;
    const int64_t gid = d.gid0 + e;
    const uint32_t ep = (uint32_t)(d.episode[e] + 1);
    uint32_t minstd = d.minstd[e];
    __syncwarp();
    if (lane == 0) d.episode[e] = (int32_t)ep;
    const bool nav3d = d.rules == XW_RULES_NAV3D;
    const int task = nav3d ? xw_get_rand_ind(minstd, 5) : 0;
    int level = 0;
    if (lane == 0) level = xw_curriculum_level(d, e);
    level = __shfl_sync(0xffffffffu, level, 0);
    XwMapCtx c;
    XwTaskOut o;
    XwStackStrided stk;
    stk.p = s_stack + lane;
    {   // attempt 0's maze draws: Philox block b on lane b % 32 (the walk itself is sequential, its random numbers are not)
        const int D = d.curriculum != 0 ? 3 + level : d.H, nblk = xw_maze_blocks(D);
        for (int b = lane; b < nblk; b += 32) {
            const XwDraw4 v = xw_draw_block(d.seed, gid, ep, 0, XW_SITE_MAZE, (uint32_t)b);
            s_draws[4 * b] = v.v0; s_draws[4 * b + 1] = v.v1; s_draws[4 * b + 2] = v.v2; s_draws[4 * b + 3] = v.v3;
        }
        __syncwarp();
    }
    int st = 0;
    XwMazeDrawsTable mz;
    if (lane == 0) {
        mz.t = s_draws;
        st = xw_reset_attempt_w(d, gid, ep, 0, task, level, c, o, stk, mz);
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st != 0 || !nav3d) {
        if (lane == 0) {
            if (st == 1) xw_reset_commit(d, e, ep, minstd, task, c, o);
            else d.error[e] = XW_ERR_INVALID_ARG;
        }
        __syncwarp();
        return;
    }
    // rounds of `retry_width` attempts, then 32 at a time (a round costs up to its width in divergent lanes; most
    // retries succeed within the first few attempts, the rare hopeless map / task pair wants them all at once)
    uint32_t width = (uint32_t)d.retry_width;
    for (uint32_t base = 1; base < 64; base += width, width = 32) {
        const uint32_t att = base + lane;
        st = 0;
        if ((uint32_t)lane < width && att < 64) {
            mz.t = nullptr;
            mz.q = xw_draw_seq(d.seed, gid, ep, att, XW_SITE_MAZE);
            st = xw_reset_attempt_w(d, gid, ep, att, task, level, c, o, stk, mz);
        }
        const unsigned done = __ballot_sync(0xffffffffu, st != 0);
        if (done) {
            if (lane == __ffs(done) - 1) {
                if (st == 1) xw_reset_commit(d, e, ep, minstd, task, c, o);
                else d.error[e] = XW_ERR_INVALID_ARG;
            }
            __syncwarp();
            return;
        }
    }
    if (lane == 0) d.error[e] = XW_ERR_INVALID_ARG;
    __syncwarp();
}
#endif
