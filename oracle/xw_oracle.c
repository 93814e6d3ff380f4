/*
 * xw_oracle.c -- CPU restatement of the XWorld2D hot path.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
 * The product (xworld_b200/) never links or calls this file.
 *
 * It follows /root/reference function by function (citations on each function).  Draws that the
 * reference takes from Python's *unseeded* `random` module are taken here from Philox4x32-10
 * substreams (one per call site, see xw_oracle.h); tests/golden/gen_reference_python.py runs the
 * reference's own Python with `random` patched onto the same substreams and the results are
 * committed under tests/golden/ -- that is how this file is pinned.  Draws the reference takes from
 * its seeded C++ engine (simulator_util.cpp) are restated exactly (minstd_rand0 + libstdc++
 * std::hash + uniform_int_distribution) and pinned by tests/test_simulator_seed.cpp:23-25.
 */
#include "xw_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* RNG                                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al., SC'11), the published algorithm. */
void xo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t xo_draw(uint64_t seed, int64_t env_gid, uint32_t episode, uint32_t attempt, uint32_t site,
                 uint32_t index) {
    uint32_t ctr[4] = {(uint32_t)((uint64_t)env_gid), (uint32_t)((uint64_t)env_gid >> 32), episode,
                       ((attempt & 0xffu) << 24) | ((site & 0xffu) << 16) | ((index >> 2) & 0xffffu)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    xo_philox4x32_10(ctr, key, out);
    return out[index & 3u];
}

/* maps a 32-bit draw to [0, n): stands in for python's int(random() * n) (random.py choice/shuffle) */
uint32_t xo_randbelow(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

/* libstdc++ std::_Hash_bytes (libsupc++/hash_bytes.cc, 64-bit), what std::hash<std::string> calls
 * in simulator_util.cpp:47 */
uint64_t xo_std_hash_bytes(const void* ptr, uint64_t len) {
    const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
    const unsigned char* buf = (const unsigned char*)ptr;
    const uint64_t len_aligned = len & ~(uint64_t)0x7;
    const unsigned char* end = buf + len_aligned;
    uint64_t hash = (uint64_t)0xc70f6907UL ^ (len * mul);
    for (const unsigned char* p = buf; p != end; p += 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        uint64_t data = v * mul;
        data = (data ^ (data >> 47)) * mul;
        hash ^= data;
        hash *= mul;
    }
    if ((len & 0x7) != 0) {
        int n = (int)(len & 0x7);
        uint64_t data = 0;
        --n;
        do data = (data << 8) + end[n]; while (--n >= 0);
        hash ^= data;
        hash *= mul;
    }
    hash = (hash ^ (hash >> 47)) * mul;
    hash = hash ^ (hash >> 47);
    return hash;
}

/* ThreadCounter (simulator_util.cpp:38-52): int seed = hash(to_string(FLAGS_simulator_seed +
 * (++__num_threads))); reng_.seed(seed) with std::default_random_engine == minstd_rand0. */
uint32_t xo_minstd_seed_for_thread(int32_t simulator_seed, int32_t thread_no) {
    char buf[32];
    int n = snprintf(buf, sizeof buf, "%d", simulator_seed + thread_no);
    int32_t seed = (int32_t)(uint32_t)xo_std_hash_bytes(buf, (uint64_t)n);
    uint64_t s = (uint64_t)(int64_t)seed; /* int -> uint_fast32_t (unsigned long) */
    uint64_t x = s % 2147483647ull;
    if (x == 0) x = 1;
    return (uint32_t)x;
}

static uint32_t minstd_next(uint32_t* st) {
    *st = (uint32_t)(((uint64_t)*st * 16807ull) % 2147483647ull);
    return *st;
}

/* util::get_rand_ind (simulator_util.cpp:66-73): std::uniform_int_distribution<int>(0,size-1) on
 * minstd_rand0, libstdc++ "downscaling" branch (bits/uniform_int_dist.h). */
int32_t xo_get_rand_ind(uint32_t* st, int32_t size) {
    const uint64_t urngrange = 2147483646ull - 1ull;
    const uint64_t uerange = (uint64_t)size;
    const uint64_t scaling = urngrange / uerange;
    const uint64_t past = uerange * scaling;
    uint64_t ret;
    do ret = (uint64_t)minstd_next(st) - 1ull; while (ret >= past);
    return (int32_t)(ret / scaling);
}

/* ------------------------------------------------------------------------------------------ */
/* Map generation                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* spanning_tree_maze_generator (python/maze2d.py:74-114).  maze[y*D+x] in {' ', '#'}. */
void xo_maze(uint64_t seed, int64_t gid, uint32_t ep, uint32_t att, int D, char* maze) {
    int pad = (D % 2 == 0);
    int X = pad ? D - 1 : D;
    int nx = (X + 1) / 2;
    for (int y = 0; y < D; ++y)
        for (int x = 0; x < D; ++x) maze[y * D + x] = '#';
    for (int y = 0; y < X; ++y)
        for (int x = 0; x < X; ++x) maze[y * D + x] = (x % 2 == 0 && y % 2 == 0) ? ' ' : '#';
    /* dfs(cur): visited.add(cur); shuffle(moves); for m in moves: ... recurse   maze2d.py:90-101 */
    static const int MX[4] = {-1, 1, 0, 0}, MY[4] = {0, 0, 1, -1};
    uint8_t visited[XW_MAX_DIM * XW_MAX_DIM];
    memset(visited, 0, sizeof visited);
    struct { int x, y, next; uint8_t order[4]; } stack[XW_MAX_DIM * XW_MAX_DIM];
    int sp = 0;
    uint32_t visit_no = 0;
#define PUSH_NODE(px, py)                                                            \
    do {                                                                             \
        stack[sp].x = (px); stack[sp].y = (py); stack[sp].next = 0;                  \
        for (int q = 0; q < 4; ++q) stack[sp].order[q] = (uint8_t)q;                 \
        for (int i = 3; i >= 1; --i) { /* random.shuffle: Fisher-Yates from the end */ \
            uint32_t j = xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_MAZE, visit_no * 3 + (3 - i)), i + 1); \
            uint8_t t = stack[sp].order[i]; stack[sp].order[i] = stack[sp].order[j]; stack[sp].order[j] = t; \
        }                                                                            \
        visited[(py) * nx + (px)] = 1; ++visit_no; ++sp;                             \
    } while (0)
    PUSH_NODE(0, 0);
    while (sp > 0) {
        int top = sp - 1;
        if (stack[top].next == 4) { --sp; continue; }
        int m = stack[top].order[stack[top].next++];
        int cx = stack[top].x, cy = stack[top].y;
        int qx = cx + MX[m], qy = cy + MY[m];
        if (qx >= 0 && qx < nx && qy >= 0 && qy < nx && !visited[qy * nx + qx]) {
            maze[(cy + qy) * D + (cx + qx)] = ' '; /* edge mid-point, maze2d.py:105-108 */
            PUSH_NODE(qx, qy);
        }
    }
#undef PUSH_NODE
    if (pad) { /* maze2d.py:110-113 */
        for (int i = 0; i < X; ++i) maze[X * D + i] = (i % 2 == 0) ? ' ' : '#';
        for (int i = 0; i < D; ++i) maze[i * D + X] = (i % 2 == 0) ? ' ' : '#';
    }
}

static int cell_of(const xo_env* e, int x, int y) { return y * e->W + x; }
static int in_bounds(const xo_env* e, int x, int y) { return x >= 0 && x < e->W && y >= 0 && y < e->H; }
static int is_free(const xo_env* e, int x, int y) { /* (x,y,0) in env.available_grids */
    return in_bounds(e, x, y) && e->grid[cell_of(e, x, y)] == XW_CELL_EMPTY;
}

/* pick the k-th free cell in canonical row-major order */
static int nth_free(const xo_env* e, int k) {
    for (int c = 0; c < e->H * e->W; ++c)
        if (e->grid[c] == XW_CELL_EMPTY) { if (k == 0) return c; --k; }
    return -1;
}
static int count_free(const xo_env* e) {
    int n = 0;
    for (int c = 0; c < e->H * e->W; ++c) n += (e->grid[c] == XW_CELL_EMPTY);
    return n;
}

/* XWorldEnv.reset (xworld_env.py:95-101) = __clean_env + XWorldNav._configure (XWorldNav.py:16-67,
 * curriculum == 0 branch) + __instantiate_entities (xworld_env.py:412-452, maze_generation on). */
static int gen_map(const xw_config* cfg, const xw_catalog* cat, xo_env* e, uint32_t ep, uint32_t att) {
    const uint64_t seed = cfg->seed;
    const int64_t gid = e->env_gid;
    /* XWorldNav.py:36-58: curriculum == 0 -> the last level's numbers (cfg); else compute(current_level) */
    static const int num_goals_seq[6] = {2, 2, 2, 4, 4, 4}, num_blocks_seq[6] = {0, 3, 6, 9, 12, 16};
    const int curr = cfg->curriculum != 0;
    const int D = curr ? 3 + e->level : cfg->height;                 /* min_dim + current_level */
    const int n_goals = curr ? num_goals_seq[e->level] : cfg->n_goals;
    const int n_blocks = curr ? num_blocks_seq[e->level] : cfg->n_blocks;
    e->H = D; e->W = D; e->n_goals = n_goals; e->dim = D;            /* set_dims(current_dim, current_dim) */
    if (cfg->height != cfg->width) return XW_ERR_INVALID_ARG; /* "only support square maps" maze2d.py:78 */
    memset(e->goal_x, 0, sizeof e->goal_x); memset(e->goal_y, 0, sizeof e->goal_y);
    memset(e->goal_icon, 0, sizeof e->goal_icon); memset(e->goal_name, 0, sizeof e->goal_name);
    /* goal names: random.shuffle(goal_names); set_entity(name=goal_names.pop()) XWorldNav.py:60-62 */
    int n = cat->n_names;
    if (n < n_goals) return XW_ERR_INVALID_ARG;
    int* names = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) names[i] = i;
    for (int k = 0; k < n_goals; ++k) {
        int i = n - 1 - k;
        if (i >= 1) {
            uint32_t j = xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_NAMES, (uint32_t)k), (uint32_t)i + 1);
            int t = names[i]; names[i] = names[j]; names[j] = t;
        }
        e->goal_name[k] = names[i];
    }
    free(names);
    /* maze + block list, xworld_env.py:419-421 */
    char maze[XW_MAX_DIM * XW_MAX_DIM];
    xo_maze(seed, gid, ep, att, D, maze);
    int blocks[XW_MAX_DIM * XW_MAX_DIM], nb = 0;
    for (int y = 0; y < D; ++y)
        for (int x = 0; x < D; ++x)
            if (maze[y * D + x] == '#') blocks[nb++] = y * D + x;
    if (nb < n_blocks) return XW_ERR_INVALID_ARG; /* "too many blocks for a valid maze" :443 */
    /* "first remove all maze blocks from the available set" :427-431 : mark them temporarily */
    memset(e->grid, XW_CELL_EMPTY, sizeof e->grid);
    for (int i = 0; i < nb; ++i) e->grid[blocks[i]] = 0xff;
    /* entities in creation order: goals, blocks, agent (XWorldNav.py:61-67) */
    for (int k = 0; k < n_goals; ++k) { /* set_property -> loc, asset_path  xworld_env.py:187-199 */
        int nf = count_free(e);
        if (nf == 0) return XW_ERR_INVALID_ARG;
        int c = nth_free(e, (int)xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_GOAL_LOC, (uint32_t)k), (uint32_t)nf));
        e->grid[c] = (uint8_t)(XW_CELL_GOAL0 + k);
        e->goal_x[k] = c % D; e->goal_y[k] = c / D;
        int f = cat->name_first[e->goal_name[k]], nv = cat->name_first[e->goal_name[k] + 1] - f;
        e->goal_icon[k] = cat->name_icons[f + (int)xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_GOAL_ASSET, (uint32_t)k), (uint32_t)nv)];
    }
    /* blocks: random.shuffle(blocks); e.loc = blocks.pop()  :424,444 */
    int block_cells[XW_MAX_DIM * XW_MAX_DIM];
    for (int k = 0; k < n_blocks; ++k) {
        int i = nb - 1 - k;
        if (i >= 1) {
            uint32_t j = xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_BLOCKS, (uint32_t)k), (uint32_t)i + 1);
            int t = blocks[i]; blocks[i] = blocks[j]; blocks[j] = t;
        }
        block_cells[k] = blocks[i];
    }
    { /* agent */
        int nf = count_free(e);
        if (nf == 0) return XW_ERR_INVALID_ARG;
        int c = nth_free(e, (int)xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_AGENT_LOC, 0), (uint32_t)nf));
        e->grid[c] = XW_CELL_AGENT;
        e->agent_x = c % D; e->agent_y = c / D;
    }
    /* "add back the unused grids" :450 */
    for (int i = 0; i < nb; ++i) e->grid[blocks[i]] = XW_CELL_EMPTY;
    for (int k = 0; k < n_blocks; ++k) e->grid[block_cells[k]] = XW_CELL_BLOCK;
    e->agent_yaw = 1.5707963; /* Entity default yaw, xworld_env.py:42 (not randomised when visible_radius==0) */
    for (int k = 0; k < XW_MAX_GOALS; ++k) { e->goal_yaw[k] = 1.5707963; e->goal_scale[k] = 1.0; e->goal_offset[k] = 0.0; }
    if (cfg->visible_radius > 0) { /* set_property, xworld_env.py:207-223: "if partially observed, perturb the objects" */
        static const int yaw_range[4] = {-1, 0, 1, 2}; /* range(-1, 3) */
        e->agent_yaw = yaw_range[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_AGENT_YAW, 0), 4)] * 1.5707963;
        for (int k = 0; k < n_goals; ++k)
            xo_goal_pose(seed, gid, ep, att, k, &e->goal_yaw[k], &e->goal_scale[k], &e->goal_offset[k]);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* BFS helpers (python/maze2d.py)                                                             */
/* ------------------------------------------------------------------------------------------ */

/* flood_fill (maze2d.py:21-39): cells reachable from `seed_cell` (exclusive), in discovery order.
 * obstacle(c) = grid[c] is a block or a goal.  Returns count; out[] holds cells. */
static int flood_fill(const xo_env* e, int seed_cell, int* out) {
    static const int MX[4] = {-1, 1, 0, 0}, MY[4] = {0, 0, -1, 1};
    uint8_t visited[XW_MAX_DIM * XW_MAX_DIM];
    int que[XW_MAX_DIM * XW_MAX_DIM], qh = 0, qt = 0, n = 0;
    memset(visited, 0, sizeof visited);
    visited[seed_cell] = 1;
    que[qt++] = seed_cell;
    while (qh < qt) {
        int cur = que[qh++];
        int cx = cur % e->W, cy = cur / e->W;
        for (int m = 0; m < 4; ++m) {
            int x = cx + MX[m], y = cy + MY[m];
            if (!in_bounds(e, x, y)) continue;
            int c = cell_of(e, x, y);
            if (visited[c]) continue;
            if (e->grid[c] == XW_CELL_BLOCK || e->grid[c] >= XW_CELL_GOAL0) continue;
            visited[c] = 1;
            que[qt++] = c;
            out[n++] = c;
        }
    }
    return n;
}

/* XWorld3DTask._reachable (xworld3d_task.py:328-342): bfs(start,end) with obstacles = blocks +
 * goals except `end`.  goals_block = 0 gives XWorldTask._reachable (xworld_task.py:340-350), where
 * only blocks are obstacles. */
static int reachable(const xo_env* e, int start, int end, int goals_block) {
    static const int MX[4] = {-1, 1, 0, 0}, MY[4] = {0, 0, -1, 1};
    if (start == end) return 1;
    uint8_t visited[XW_MAX_DIM * XW_MAX_DIM];
    int que[XW_MAX_DIM * XW_MAX_DIM], qh = 0, qt = 0;
    memset(visited, 0, sizeof visited);
    visited[start] = 1;
    que[qt++] = start;
    while (qh < qt) {
        int cur = que[qh++];
        if (cur == end) return 1;
        int cx = cur % e->W, cy = cur / e->W;
        for (int m = 0; m < 4; ++m) {
            int x = cx + MX[m], y = cy + MY[m];
            if (!in_bounds(e, x, y)) continue;
            int c = cell_of(e, x, y);
            if (visited[c]) continue;
            if (c != end) {
                if (e->grid[c] == XW_CELL_BLOCK) continue;
                if (goals_block && e->grid[c] >= XW_CELL_GOAL0) continue;
            }
            visited[c] = 1;
            que[qt++] = c;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Geometry of xworld3d_task.py                                                               */
/* ------------------------------------------------------------------------------------------ */
#define T3_PI 3.1415926
#define T3_PI_2 (T3_PI / 2)
#define T3_PI_4 (T3_PI / 4)

/* XWorld3DTask._get_direction_and_distance (xworld3d_task.py:98-124), theta and dist only */
static void direction_and_distance(double p1x, double p1y, double p2x, double p2y, double yaw,
                                   double* theta, double* dist) {
    double dx = p2x - p1x, dy = p2y - p1y;
    double d = sqrt(dx * dx + dy * dy);
    *dist = d;
    if (d == 0) { *theta = 0; *dist = 0; return; }
    double v1x = cos(yaw), v1y = sin(yaw);
    double v2x = dx / d, v2y = dy / d;
    double ct = v1x * v2x + v1y * v2y; ct = fmax(-1, fmin(1, ct));
    double st = v1y * v2x - v1x * v2y; st = fmax(-1, fmin(1, st));
    *theta = acos(ct) * copysign(1, asin(st));
}

enum { DIR_FALSE = 0, DIR_FRONT = 1, DIR_BEHIND = 2, DIR_LEFT = 3, DIR_RIGHT = 4 };

/* XWorld3DNavTargetDirection.__compute_triple_direction (XWorld3DNavTargetDirection.py:98-126),
 * 2-D world branch ("3D" not in env class name => left/right swapped) */
static int triple_direction(double tx, double ty, double rx, double ry, double view_yaw) {
    double theta, dist;
    direction_and_distance(tx, ty, rx, ry, view_yaw, &theta, &dist);
    if (dist == 0) return DIR_FALSE;
    int sign = theta > 0;
    int flag = 0;
    theta = fabs(theta);
    if (theta > T3_PI_2) { flag = 1; theta = T3_PI - theta; }
    if (theta < T3_PI_4 + 1e-3) return flag ? DIR_BEHIND : DIR_FRONT;
    else if (T3_PI_2 - theta < T3_PI_4 + 1e-3) return sign ? DIR_RIGHT : DIR_LEFT;
    return DIR_FALSE;
}

/* ------------------------------------------------------------------------------------------ */
/* Task idle stages (episode start)                                                           */
/* ------------------------------------------------------------------------------------------ */

static void remove_entity_cell(xo_env* e, int x, int y) { e->grid[cell_of(e, x, y)] = XW_CELL_EMPTY; }

/* 4-neighbours of (x,y) that are free, excluding cell `excl` : _get_surrounding_empty_grids with
 * distance_threshold=1.0 (xworld3d_task.py:208-224) */
static int free_neighbours(const xo_env* e, int x, int y, int excl, int* out) {
    int n = 0;
    /* canonical row-major order of the result: (x,y-1), (x-1,y), (x+1,y), (x,y+1) */
    static const int NX[4] = {0, -1, 1, 0}, NY[4] = {-1, 0, 0, 1};
    for (int m = 0; m < 4; ++m) {
        int qx = x + NX[m], qy = y + NY[m];
        if (is_free(e, qx, qy) && cell_of(e, qx, qy) != excl) { if (out) out[n] = cell_of(e, qx, qy); ++n; }
    }
    return n;
}

typedef struct { int16_t a, b; } tile_pair;

/* XWorld3DTask._get_p_tiles (xworld3d_task.py:226-251) */
static int p_tiles(const xo_env* e, tile_pair* out) {
    int n = 0;
    static const int DX[3] = {1, 0, 1}, DY[3] = {0, 1, 1};
    for (int y = 0; y < e->H; ++y)
        for (int x = 0; x < e->W; ++x)
            for (int k = 0; k < 3; ++k) {
                int x2 = x + DX[k], y2 = y + DY[k];
                if (is_free(e, x, y) && is_free(e, x2, y2)) {
                    int p1 = cell_of(e, x, y), p2 = cell_of(e, x2, y2);
                    if (free_neighbours(e, x2, y2, p1, NULL) > 0) { out[n].a = (int16_t)p1; out[n].b = (int16_t)p2; ++n; }
                    if (free_neighbours(e, x, y, p2, NULL) > 0) { out[n].a = (int16_t)p2; out[n].b = (int16_t)p1; ++n; }
                }
            }
    return n;
}

/* XWorld3DTask._get_t_tiles (xworld3d_task.py:253-276) */
static int t_tiles(const xo_env* e, tile_pair* out) {
    int n = 0;
    for (int y = 0; y < e->H; ++y)
        for (int x = 0; x < e->W; ++x)
            if (is_free(e, x, y)) {
                if (is_free(e, x - 1, y) && is_free(e, x + 1, y) && (is_free(e, x, y - 1) || is_free(e, x, y + 1))) {
                    out[n].a = (int16_t)cell_of(e, x - 1, y); out[n].b = (int16_t)cell_of(e, x + 1, y); ++n;
                }
                if (is_free(e, x, y - 1) && is_free(e, x, y + 1) && (is_free(e, x - 1, y) || is_free(e, x + 1, y))) {
                    out[n].a = (int16_t)cell_of(e, x, y - 1); out[n].b = (int16_t)cell_of(e, x, y + 1); ++n;
                }
            }
    return n;
}

/* XWorld3DTask._get_l_tiles (xworld3d_task.py:302-322) */
static int l_tiles(const xo_env* e, tile_pair* out) {
    int n = 0;
    for (int y = 0; y < e->H; ++y)
        for (int x = 0; x < e->W; ++x) {
            if (is_free(e, x, y) && is_free(e, x, y + 1) && is_free(e, x, y + 2)) {
                out[n].a = (int16_t)cell_of(e, x, y); out[n].b = (int16_t)cell_of(e, x, y + 1); ++n;
                out[n].a = (int16_t)cell_of(e, x, y + 1); out[n].b = (int16_t)cell_of(e, x, y + 2); ++n;
            }
            if (is_free(e, x, y) && is_free(e, x + 1, y) && is_free(e, x + 2, y)) {
                out[n].a = (int16_t)cell_of(e, x, y); out[n].b = (int16_t)cell_of(e, x + 1, y); ++n;
                out[n].a = (int16_t)cell_of(e, x + 1, y); out[n].b = (int16_t)cell_of(e, x + 2, y); ++n;
            }
        }
    return n;
}

/* `agent.loc = random.choice(new_a)` (XWorld3DNavTargetNear.py:52 and siblings): uniform over the cells _propagate_agent
 * found.  The reference's list is in BFS discovery order and its draw is Python's unseeded random(); the contract is the
 * SET and the uniform choice, so the draw is mapped onto the set in canonical row-major order (as for every other site
 * whose order is an artefact: gen_reference_python.py sorts the same list before indexing it). */
static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }
static int pick_canonical(int* cells, int n, uint32_t u) {
    qsort(cells, (size_t)n, sizeof(int), cmp_int);
    return cells[xo_randbelow(u, (uint32_t)n)];
}

static void place_goal(xo_env* e, int g, int cell) {
    e->goal_x[g] = cell % e->W; e->goal_y[g] = cell / e->W;
    e->grid[cell] = (uint8_t)(XW_CELL_GOAL0 + g);
}
static void place_agent(xo_env* e, int cell) {
    e->agent_x = cell % e->W; e->agent_y = cell / e->W;
    e->grid[cell] = XW_CELL_AGENT;
}

/* "random.shuffle(goals); g1, g2 = goals[:2]" (XWorld3DNavTargetNear.py:38-39 and siblings) */
static void shuffle_first_two(const xw_config* cfg, const xo_env* e, uint32_t ep, uint32_t att, int* g1, int* g2) {
    int a[XW_MAX_GOALS];
    for (int i = 0; i < e->n_goals; ++i) a[i] = i;
    for (int i = e->n_goals - 1; i >= 1; --i) {
        uint32_t j = xo_randbelow(xo_draw(cfg->seed, e->env_gid, ep, att, XO_SITE_TASK_SHUF, (uint32_t)(e->n_goals - 1 - i)), (uint32_t)i + 1);
        int t = a[i]; a[i] = a[j]; a[j] = t;
    }
    *g1 = a[0]; *g2 = a[1];
}

/* The five idle() stages of games/xworld3d/tasks/XWorld3DNav*.py.  Returns 0, or 1 when the
 * reference would hit `assert ..., "map too crowded?"` (LOG(FATAL) there; re-draw here). */
static int idle3d(const xw_config* cfg, xo_env* e, uint32_t ep, uint32_t att) {
    const uint64_t seed = cfg->seed;
    const int64_t gid = e->env_gid;
    int G = e->n_goals;
    int agent_cell = cell_of(e, e->agent_x, e->agent_y);
    e->target_mask = 0; e->aux0 = e->aux1 = e->aux2 = 0;
    if (e->task == XW_T3_TARGET || e->task == XW_T3_AVOID) {
        /* XWorld3DNavTarget.py:28-43, XWorld3DNavTargetAvoid.py:28-44 */
        int cand[XW_MAX_GOALS], nc = 0;
        for (int g = 0; g < G; ++g)
            if (reachable(e, agent_cell, cell_of(e, e->goal_x[g], e->goal_y[g]), 1)) cand[nc++] = g;
        if (nc == 0) return 1;
        int sel = cand[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_A, 0), (uint32_t)nc)];
        if (e->task == XW_T3_TARGET) {
            for (int g = 0; g < G; ++g)
                if (e->goal_name[g] == e->goal_name[sel]) e->target_mask |= 1 << g;
            e->aux0 = sel;
        } else {
            int refs[XW_MAX_GOALS], nr = 0;
            for (int g = 0; g < G; ++g)
                if (e->goal_name[g] != e->goal_name[sel]) refs[nr++] = g;
            if (nr == 0) return 1; /* assert referents, "Identical object names?" */
            int ref = refs[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_B, 0), (uint32_t)nr)];
            for (int g = 0; g < G; ++g)
                if (e->goal_name[g] != e->goal_name[ref]) e->target_mask |= 1 << g;
            e->aux0 = ref;
        }
        return 0;
    }
    /* relocating tasks: delete agent, g1, g2 */
    if (G < 2) return 1;
    int g1, g2;
    shuffle_first_two(cfg, e, ep, att, &g1, &g2);
    remove_entity_cell(e, e->agent_x, e->agent_y);
    remove_entity_cell(e, e->goal_x[g1], e->goal_y[g1]);
    remove_entity_cell(e, e->goal_x[g2], e->goal_y[g2]);
    tile_pair* tiles = (tile_pair*)malloc(sizeof(tile_pair) * 8 * XW_MAX_DIM * XW_MAX_DIM);
    int filled[XW_MAX_DIM * XW_MAX_DIM + 1];
    int rc = 0;
    if (e->task == XW_T3_NEAR) { /* XWorld3DNavTargetNear.py:28-61 */
        int nt = p_tiles(e, tiles);
        if (nt == 0) { rc = 1; goto done; }
        tile_pair t = tiles[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_A, 0), (uint32_t)nt)];
        place_goal(e, g1, t.a); place_goal(e, g2, t.b);
        int nf = flood_fill(e, t.b, filled);
        if (nf == 0) { rc = 1; goto done; }
        place_agent(e, pick_canonical(filled, nf, xo_draw(seed, gid, ep, att, XO_SITE_TASK_AGENT, 0)));
        /* _get_surrounding_goals(refer=g1.loc), threshold 1.5 (+1e-3)  xworld3d_task.py:189-206 */
        for (int g = 0; g < G; ++g) {
            if (e->goal_x[g] == e->goal_x[g1] && e->goal_y[g] == e->goal_y[g1]) continue;
            double dx = e->goal_x[g] - e->goal_x[g1], dy = e->goal_y[g] - e->goal_y[g1];
            if (sqrt(dx * dx + dy * dy) < 1.5 + 1e-3) e->target_mask |= 1 << g;
        }
        e->aux0 = g1;
    } else if (e->task == XW_T3_BETWEEN) { /* XWorld3DNavTargetBetween.py:29-63 */
        int nt = t_tiles(e, tiles);
        if (nt == 0) { rc = 1; goto done; }
        tile_pair t = tiles[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_A, 0), (uint32_t)nt)];
        place_goal(e, g1, t.a); place_goal(e, g2, t.b);
        int mx = (e->goal_x[g1] + e->goal_x[g2]) / 2, my = (e->goal_y[g1] + e->goal_y[g2]) / 2; /* _middle_loc :324 */
        int nf = flood_fill(e, cell_of(e, mx, my), filled);
        if (nf == 0) { rc = 1; goto done; }
        place_agent(e, pick_canonical(filled, nf, xo_draw(seed, gid, ep, att, XO_SITE_TASK_AGENT, 0)));
        e->aux0 = g1 | (g2 << 4); e->aux1 = mx; e->aux2 = my; /* g1, g2: the bindings of G1, G2 (:59-62) */
    } else { /* XW_T3_DIRECTION, XWorld3DNavTargetDirection.py:29-76 */
        int nt = l_tiles(e, tiles);
        if (nt == 0) { rc = 1; goto done; }
        tile_pair t = tiles[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_A, 0), (uint32_t)nt)];
        place_goal(e, g1, t.a); place_goal(e, g2, t.b);
        int empties[4];
        int target = g1, referent = g2;
        int ne = free_neighbours(e, e->goal_x[g1], e->goal_y[g1], -1, empties);
        if (ne == 0) {
            ne = free_neighbours(e, e->goal_x[g2], e->goal_y[g2], -1, empties);
            if (ne == 0) { rc = 1; goto done; }
            target = g2; referent = g1;
        }
        int ecell = empties[xo_randbelow(xo_draw(seed, gid, ep, att, XO_SITE_TASK_B, 0), (uint32_t)ne)];
        int ex = ecell % e->W, ey = ecell / e->W;
        double view_yaw = atan2((double)(e->goal_y[target] - ey), (double)(e->goal_x[target] - ex));
        int dir = triple_direction(e->goal_x[target], e->goal_y[target], e->goal_x[referent], e->goal_y[referent], view_yaw);
        /* assert direction != "behind" */
        /* _propagate_agent([e], inclusive=True): seed first, then BFS order  xworld3d_task.py:344-355 */
        filled[0] = ecell;
        int nf = 1 + flood_fill(e, ecell, filled + 1);
        place_agent(e, pick_canonical(filled, nf, xo_draw(seed, gid, ep, att, XO_SITE_TASK_AGENT, 0)));
        e->aux0 = referent; e->aux1 = dir; e->aux2 = target;
    }
done:
    free(tiles);
    return rc;
}

/* XWorldNav{Target,Near,ColorTarget,Between}.idle (games/xworld/tasks/).  In this reference commit
 * NavNear and NavBetween can never start: they hand 2-tuples (x,y) to bfs() whose frontier holds
 * 3-tuples (x,y,0), so `cur == end` never holds (XWorldNavNear.py:13-16, XWorldNavBetween.py:11-13,
 * maze2d.py:49-63).  NavTarget / NavColorTarget start iff a (coloured) goal is reachable. */
static void idle2d(const xw_config* cfg, const xw_catalog* cat, xo_env* e) {
    e->task = xo_get_rand_ind(&e->minstd, 4); /* TaskGroup::run_stage, schedule "random" */
    e->steps_in_task = 0;                    /* Task::reset -> XWorldTask.reset */
    e->target_mask = 0; e->aux0 = 0;
    if (e->task == XW_T2_TARGET || e->task == XW_T2_COLOR_TARGET) {
        int agent_cell = cell_of(e, e->agent_x, e->agent_y);
        int cand[XW_MAX_GOALS], nc = 0;
        for (int g = 0; g < e->n_goals; ++g) {
            if (e->task == XW_T2_COLOR_TARGET && !cat->icon_colored[e->goal_icon[g]]) continue;
            if (reachable(e, agent_cell, cell_of(e, e->goal_x[g], e->goal_y[g]), 0)) cand[nc++] = g;
        }
        if (nc > 0) {
            int sel = cand[xo_randbelow(xo_draw(cfg->seed, e->env_gid, (uint32_t)e->episode, 0, XO_SITE_TASK_A, (uint32_t)e->num_steps), (uint32_t)nc)];
            e->aux0 = sel; e->target_mask = 1 << sel;
            e->stage = XW_STAGE_NAVIGATION;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Teacher::teach for one env                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* XWorld3DTask.__record_result (xworld3d_task.py:129-133) -> XWorldEnv.record_environment_usage (xworld_env.py:60-67):
 * the env keeps a reference to the task's list, so the list itself is the usage record. */
static void record_result(const xw_config* cfg, xo_env* e, int res) {
    if (cfg->curriculum == 0) return; /* nothing reads the record then */
    uint8_t* seq = e->seq[e->task];
    int n = e->seq_len[e->task];
    /* success_seq.append(res); if len > performance_window_size: pop(0) -- the oldest entry leaves first here,
     * because the array holds exactly 200 */
    if (n == 200) { memmove(seq, seq + 1, 199); n = 199; }
    seq[n++] = (uint8_t)res;
    e->seq_len[e->task] = n;
}

/* XWorldEnv.get_current_usage (xworld_env.py:103-110) */
static double current_usage(const xw_config* cfg, xo_env* e) {
    const int period = cfg->curriculum_check_period > 0 ? cfg->curriculum_check_period : 100;
    e->check_counter += 1;
    int any = 0;
    for (int t = 0; t < 5; ++t) any |= e->seq_len[t] > 0;
    if (e->check_counter < period || !any) return 0;
    double usage = 0; int first = 1;
    for (int t = 0; t < 5; ++t) {
        if (e->seq_len[t] == 0) continue;
        int sum = 0;
        for (int i = 0; i < e->seq_len[t]; ++i) sum += e->seq[t][i];
        double u = sum / (double)e->seq_len[t];             /* sum(l) / float(len(l)) */
        if (first || u < usage) usage = u;
        first = 0;
    }
    e->check_counter = 0;
    return usage;
}

/* cpp_get_entities (xworld_env.py:352-366): entity locations move by (offset_w, offset_h) = ((max - dim) / 2, same)
 * and __padding_walls (:454-473) fills the rest of the max_height x max_width map with bricks. */
static void embed_world(const xw_config* cfg, xo_env* e) {
    const int D = e->dim, M = cfg->height, off = (M - D) / 2;
    if (D == M) return;
    uint8_t inner[XW_MAX_DIM * XW_MAX_DIM];
    memcpy(inner, e->grid, sizeof inner);
    memset(e->grid, XW_CELL_EMPTY, sizeof e->grid);
    for (int y = 0; y < M; ++y)
        for (int x = 0; x < M; ++x) {
            int ix = x - off, iy = y - off;
            e->grid[y * M + x] = (ix >= 0 && ix < D && iy >= 0 && iy < D) ? inner[iy * D + ix] : XW_CELL_BLOCK;
        }
    e->H = e->W = M;
    e->agent_x += off; e->agent_y += off;
    for (int g = 0; g < e->n_goals; ++g) { e->goal_x[g] += off; e->goal_y[g] += off; }
    if (e->task == XW_T3_BETWEEN) { e->aux1 += off; e->aux2 += off; } /* the middle cell is kept in map coordinates */
}

/* One teach() after the agent's move(s).  collided = cell code of the blocking item (0 = none).
 * Returns the teacher reward as the double the reference accumulates (simulator.h:322). */
int xo_teach(const xw_config* cfg, const xw_catalog* cat, xo_env* e, int32_t collided, double* reward_out) {
    double reward = 0; /* TeachingEnvBuffer::reward after clear */
    e->event = XW_EVENT_NONE;
    if (cfg->rules == XW_RULES_NAV2D) {
        if (e->stage == XW_STAGE_IDLE) {
            idle2d(cfg, cat, e); /* returns ["...", 0.0, sentence] */
        } else {
            /* XWorldTask.simple_navigation_reward, xworld_task.py:184-223 (lang_acquisition) */
            double r = -0.1;                          /* time_penalty */
            if (!e->action_success) r += -0.2;        /* failed_action_penalty */
            e->steps_in_task += 1;
            /* one_channel only: `self.steps_in_cur_task >= h*w / 2` with h, w = get_max_dims() and Python-2 integer
             * division: time up -> _record_failure, back to idle, no event (xworld_task.py:203-210) */
            if (cfg->task_mode == XW_TASK_ONE_CHANNEL && e->steps_in_task >= cfg->height * cfg->width / 2) {
                e->steps_in_task = 0;
                e->n_failure += 1;
                e->stage = XW_STAGE_IDLE;
            }
            /* agent.loc == self.target: the target is a goal cell, goals block => never.
             * agent.loc in goal_locs: never, same reason. */
            reward += r;
        }
        /* second group XWorldRec (walls.json): schedule "weighted" -> simple_importance_sampling ->
         * get_rand_range_val -> one engine draw per teach; lang_acquisition reward 0 and its
         * record_event_in_buffer("") overwrites the nav group's event (teaching_task.cpp:100-101) */
        minstd_next(&e->minstd);
        e->event = XW_EVENT_NONE;
        *reward_out = reward;
        return 0;
    }
    /* XW_RULES_NAV3D */
    if (e->stage == XW_STAGE_TERMINAL) { *reward_out = 0; return 0; } /* xworld3d_task.py:407-408 */
    /* _time_reward, xworld3d_task.py:472-482 */
    double r = -0.01;
    e->steps_in_task += 1;
    /* h, w = self.env.get_dims(): the world the Python side sees, not the padded C++ map */
    const int side = cfg->curriculum != 0 ? e->dim : e->H;
    if (e->steps_in_task >= side * side * cfg->max_steps_factor) {
        record_result(cfg, e, 0);
        e->n_failure += 1;
        e->event = XW_EVENT_TIME_UP;
        e->stage = XW_STAGE_TERMINAL;
        *reward_out = reward + r;
        return 0;
    }
    /* objects_reach_test: _reach_object for every goal (xworld3d_task.py:451-454) */
    int reach_mask = 0;
    for (int g = 0; g < e->n_goals; ++g) {
        double theta, dist;
        direction_and_distance(e->agent_x, e->agent_y, e->goal_x[g], e->goal_y[g], e->agent_yaw, &theta, &dist);
        if (fabs(theta) < T3_PI_4 && collided == XW_CELL_GOAL0 + g) reach_mask |= 1 << g;
    }
    int correct = 0, wrong = 0;
    if (e->task == XW_T3_TARGET || e->task == XW_T3_AVOID || e->task == XW_T3_NEAR) {
        if (reach_mask & e->target_mask) correct = 1;
        else if (reach_mask) wrong = 1;
    } else if (e->task == XW_T3_BETWEEN) { /* XWorld3DNavTargetBetween.py:65-93 */
        if (reach_mask) wrong = 1;
        else {
            double dx = e->aux1 - e->agent_x, dy = e->aux2 - e->agent_y;
            if (sqrt(dx * dx + dy * dy) < 1.0 / 2) correct = 1;
        }
    } else { /* XWorld3DNavTargetDirection.py:78-96 */
        int ref = e->aux0, any = 0, ok = 0;
        /* self.target holds the referent's Entity OBJECT from the idle stage.  cpp_get_entities (xworld_env.py:359-361)
         * adds (offset_w, offset_h) to the loc of the env's entity objects IN PLACE when the C++ side fetches the map
         * after that stage, and update_entities_from_cpp then replaces the env's list with fresh objects in the
         * Python side's coordinates: from there on referent.loc is off by the padding offset against every g.loc.
         * With curriculum levels 0-3 (offset 2, 2, 1, 1) the test below therefore runs on a displaced referent: never
         * "close" at offset 2, close for two of the four true arrangements at offset 1.  Reproduced, not fixed. */
        const int off = cfg->curriculum != 0 ? (cfg->height - e->dim) / 2 : 0;
        const int ref_x = e->goal_x[ref] + off, ref_y = e->goal_y[ref] + off;
        for (int g = 0; g < e->n_goals; ++g)
            if (reach_mask & (1 << g)) {
                any = 1;
                int d = triple_direction(e->goal_x[g], e->goal_y[g], ref_x, ref_y, e->agent_yaw);
                double dx = e->goal_x[g] - ref_x, dy = e->goal_y[g] - ref_y;
                int close = sqrt(dx * dx + dy * dy) < 1.0 + 1e-3;
                if (d == e->aux1 && d != DIR_FALSE && close) ok = 1;
            }
        if (ok) correct = 1; else if (any) wrong = 1;
    }
    if (correct) { /* _successful_goal :456-462 */
        record_result(cfg, e, 1);
        e->n_success += 1; e->success_steps += e->steps_in_task;
        e->event = XW_EVENT_CORRECT_GOAL; r += 1.0; e->stage = XW_STAGE_TERMINAL;
    } else if (wrong) { /* _failed_goal :464-470 */
        record_result(cfg, e, 0);
        e->n_failure += 1;
        e->event = XW_EVENT_WRONG_GOAL; r += -1.0; e->stage = XW_STAGE_TERMINAL;
    }
    *reward_out = reward + r;
    return 0;
}

void xo_env_init(const xw_config* cfg, xo_env* e, int64_t env_gid) {
    memset(e, 0, sizeof *e);
    e->env_gid = env_gid;
    e->H = cfg->height; e->W = cfg->width; e->n_goals = cfg->n_goals;
    e->level = cfg->start_level; e->dim = cfg->height;
    e->agent_yaw = 1.5707963;
    for (int k = 0; k < XW_MAX_GOALS; ++k) { e->goal_yaw[k] = 1.5707963; e->goal_scale[k] = 1.0; e->goal_offset[k] = 0.0; }
    /* env i plays the role of the reference's i-th simulator thread (1-based) */
    e->minstd = xo_minstd_seed_for_thread(cfg->simulator_seed, (int32_t)(env_gid + 1));
}

/* SimulatorInterface::reset_game (simulator_interface.cpp:95-105) */
int xo_reset(const xw_config* cfg, const xw_catalog* cat, xo_env* e) {
    e->episode += 1;            /* episode counter = number of resets so far; keys the Philox draws */
    uint32_t ep = (uint32_t)e->episode;
    e->num_steps = 0;           /* GameSimulator::reset_game, simulator.cpp:115-117 */
    e->steps_in_task = 0;
    e->event = XW_EVENT_NONE;
    e->action_success = 0;      /* clear_agent_env_buffer, simulator.h:286-290 */
    e->stage = XW_STAGE_IDLE;   /* TaskGroup::reset, teaching_task.cpp:176-181 */
    if (cfg->rules == XW_RULES_NAV3D) {
        /* teacher_->teach(): TaskGroup::run_stage samples the task with the seeded C++ engine */
        e->task = xo_get_rand_ind(&e->minstd, 5);
        if (cfg->curriculum != 0) { /* XWorldNav._configure, XWorldNav.py:40-56; once per episode */
            if (current_usage(cfg, e) >= (double)cfg->curriculum && e->level < 6 - 1) e->level += 1;
        }
        int ok = 0;
        for (uint32_t att = 0; att < 64 && !ok; ++att) {
            int rc = gen_map(cfg, cat, e, ep, att);
            if (rc) return rc;
            if (idle3d(cfg, e, ep, att) == 0) ok = 1;
        }
        if (!ok) { e->error = 1; return XW_ERR_INVALID_ARG; }
        if (cfg->curriculum != 0) embed_world(cfg, e);
        e->stage = XW_STAGE_NAVIGATION;
    } else {
        if (cfg->curriculum != 0) { /* the same check; with these rules no task class ever records (xworld_task.py:205-214) */
            if (current_usage(cfg, e) >= (double)cfg->curriculum && e->level < 6 - 1) e->level += 1;
        }
        int rc = gen_map(cfg, cat, e, ep, 0);
        if (rc) return rc;
        double r;
        xo_teach(cfg, cat, e, 0, &r);
        /* bookkeeping only (the engine keeps these per-episode constants in aux1/aux2; the
         * restatement above recomputes reachability at every idle stage, as the reference does) */
        e->aux1 = e->aux2 = 0;
        for (int g = 0; g < e->n_goals; ++g) {
            if (reachable(e, cell_of(e, e->agent_x, e->agent_y), cell_of(e, e->goal_x[g], e->goal_y[g]), 0)) e->aux1 |= 1 << g;
            if (cat->icon_colored[e->goal_icon[g]]) e->aux2 |= 1 << g;
        }
        if (cfg->curriculum != 0) { int t = e->task; e->task = -1; embed_world(cfg, e); e->task = t; } /* (no middle cell here) */
    }
    return 0;
}

/* SimulatorInterface::take_actions (simulator_interface.cpp:126-137) + game_over (:111-113) */
int xo_step(const xw_config* cfg, const xw_catalog* cat, xo_env* e, int32_t action, int32_t act_rep,
            float* reward_out, int32_t* game_over) {
    static const int DX[4] = {0, 0, -1, 1}, DY[4] = {-1, 1, 0, 0}; /* XAgent::act, xitem.cpp:94-98 */
    const int fpv = cfg->visible_radius > 0;
    if (action < 0 || action >= (fpv ? 6 : 4)) { e->error = XW_ERR_INVALID_ACTION; return XW_ERR_INVALID_ACTION; }
    float r = 0;
    e->num_steps += 1; /* GameSimulator::take_actions counts calls, not repeats (simulator.cpp:100) */
    int collided = 0;
    for (int rep = 0; rep < act_rep; ++rep) {
        /* XWorldSimulator::take_action (xworld_simulator.cpp:200-265) -> XMap::move_item (xmap.cpp:76-101) */
        int tx, ty;
        if (!fpv) { tx = e->agent_x + DX[action]; ty = e->agent_y + DY[action]; }
        else { /* legal_actions_ = {MOVE_FORWARD, MOVE_BACKWARD, MOVE_LEFT_FPV, MOVE_RIGHT_FPV, TURN_LEFT, TURN_RIGHT} (xitem.cpp:84-86) */
            static const int FX[4] = {1, 0, -1, 0}, FY[4] = {0, 1, 0, -1}; /* right, down, left, up */
            const int dir = xo_facing_dir(e->agent_yaw);
            const int fx = FX[dir], fy = FY[dir];
            tx = e->agent_x; ty = e->agent_y;
            switch (action) {
                case 0: tx += fx; ty += fy; break;   /* MOVE_FORWARD   xitem.cpp:100-109 */
                case 1: tx -= fx; ty -= fy; break;   /* MOVE_BACKWARD  :110-119 */
                case 2: tx += fy; ty -= fx; break;   /* MOVE_LEFT_FPV  :120-129: right -> (x, y-1), down -> (x+1, y), left -> (x, y+1), up -> (x-1, y) */
                case 3: tx -= fy; ty += fx; break;   /* MOVE_RIGHT_FPV :130-139 */
                case 4: /* TURN_LEFT :146-151 */
                    e->agent_yaw -= M_PI / 2;
                    if (e->agent_yaw < -M_PI / 2 - 1e-4) e->agent_yaw += 2 * M_PI;
                    break;
                default: /* TURN_RIGHT :140-145 */
                    e->agent_yaw += M_PI / 2;
                    if (e->agent_yaw > M_PI + 1e-4) e->agent_yaw -= 2 * M_PI;
                    break;
            }
        }
        if (fpv && action >= 4) {
            /* the target is the agent's own cell: move_item finds the agent itself there, which is not reachable and
             * is not a contact (same id) -> returns false (xmap.cpp:76-101): a turn is a "failed" action */
            e->action_success = 0;
        } else if (!in_bounds(e, tx, ty)) {
            e->action_success = 0;
        } else if (e->grid[cell_of(e, tx, ty)] != XW_CELL_EMPTY) {
            e->action_success = 0;
            collided = e->grid[cell_of(e, tx, ty)];
        } else {
            e->grid[cell_of(e, e->agent_x, e->agent_y)] = XW_CELL_EMPTY;
            e->agent_x = tx; e->agent_y = ty;
            e->grid[cell_of(e, tx, ty)] = XW_CELL_AGENT;
            e->action_success = 1;
        }
        r += 0; /* "xworld rewards are given by the teacher" :264 */
    }
    double tr;
    xo_teach(cfg, cat, e, collided, &tr);
    r += tr; /* float += double, simulator_interface.cpp:132 */
    *reward_out = r;
    /* AgentSpecificSimulator::game_over (simulator.cpp:158-161) */
    int code = 0;
    if (cfg->max_steps > 0 && e->num_steps >= cfg->max_steps) code |= XW_MAX_STEP;
    if (cfg->task_mode == XW_TASK_LANG_ACQUISITION) {             /* one_channel: "all tasks until the max steps" :192-193 */
        if (e->event == XW_EVENT_CORRECT_GOAL) code |= XW_SUCCESS;    /* xworld_simulator.cpp:170-177 */
        else if (e->event == XW_EVENT_WRONG_GOAL) code |= XW_DEAD;
        else if (e->event == XW_EVENT_TIME_UP) code |= XW_MAX_STEP;
    }
    *game_over = code;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Render                                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* cv::resize INTER_LINEAR coefficient tables for 8U (OpenCV imgproc/resize.cpp, resize_ +
 * INTER_RESIZE_COEF_BITS = 11), verified against cv2 in tests/test_oracle_render.py */
void xo_resize_tables(int src, int dst, int32_t* ofs, int16_t* a0, int16_t* a1) {
    double inv_scale = (double)dst / (double)src;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        ofs[d] = s;
        a0[d] = (int16_t)lrintf((1.f - f) * 2048.f);
        a1[d] = (int16_t)lrintf(f * 2048.f);
    }
}

/* src HWC 3 channels -> dst HWC.  HResizeLinear + VResizeLinear<uchar,int,short,FixedPtCast<..,22>>.
 * Columns: index and weight clamped as in xo_resize_tables (resize.cpp: `if (sx < 0) fx = 0, sx = 0`, same at the right
 * edge).  Rows are different: the weights keep the raw fraction and only the two row INDICES are clipped to the
 * image (resizeGeneric_Invoker: `sy = clip(sy0 - ksize2 + 1 + k, 0, ssize.height)`), so above the first / below the
 * last source row both taps read the same row with weights (1 - fy, fy) -- which the truncating V pass does not
 * collapse to a copy.  Only upscaling reaches those rows (the first-person view's resize to the map size). */
void xo_resize_linear_8uc3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
    if (sh == dh && sw == dw) { memcpy(dst, src, (size_t)sh * sw * 3); return; } /* resize(): src.copyTo(dst) */
    int32_t* xofs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(dw + dh));
    int32_t* yofs = xofs + dw;
    int16_t* xa0 = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)(dw + dh));
    int16_t *xa1 = xa0 + dw, *ya0 = xa1 + dw, *ya1 = ya0 + dh;
    xo_resize_tables(sw, dw, xofs, xa0, xa1);
    {
        double scale = 1. / ((double)dh / (double)sh);
        for (int d = 0; d < dh; ++d) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int sy = (int)floorf(f);
            f -= (float)sy;
            yofs[d] = sy;
            ya0[d] = (int16_t)lrintf((1.f - f) * 2048.f);
            ya1[d] = (int16_t)lrintf(f * 2048.f);
        }
    }
    int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)dw * 3);
    int32_t* row[2] = {rows, rows + (size_t)dw * 3};
    for (int dy = 0; dy < dh; ++dy) {
        for (int k = 0; k < 2; ++k) {
            int sy = yofs[dy] + k;
            sy = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
            const uint8_t* S = src + (size_t)sy * sw * 3;
            int32_t* D = row[k];
            for (int dx = 0; dx < dw; ++dx) {
                int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                for (int c = 0; c < 3; ++c)
                    D[dx * 3 + c] = S[sx * 3 + c] * xa0[dx] + S[sx1 * 3 + c] * xa1[dx];
            }
        }
        int b0 = ya0[dy], b1 = ya1[dy];
        uint8_t* O = dst + (size_t)dy * dw * 3;
        for (int i = 0; i < dw * 3; ++i)
            O[i] = (uint8_t)((((b0 * (row[0][i] >> 4)) >> 16) + ((b1 * (row[1][i] >> 4)) >> 16) + 2) >> 2);
    }
    free(rows); free(xa0); free(xofs);
}

void xo_frame_dims(const xw_config* cfg, int* oh, int* ow) {
    /* XWorldSimulator::init (xworld_simulator.cpp:48-68): block_size 12 when fully observed, 84 / visible_radius
     * otherwise (the frame is visible_radius blocks a side) */
    int bh = cfg->height * 12, bw = cfg->width * 12;
    if (cfg->visible_radius > 0) { int vr = xo_visible_radius(cfg); bh = bw = vr * (84 / vr); }
    *oh = cfg->out_h > 0 ? cfg->out_h : bh;
    *ow = cfg->out_w > 0 ? cfg->out_w : bw;
}

static int icon_of_cell(const xw_catalog* cat, const xo_env* e, int code) {
    if (code == XW_CELL_BLOCK) return cat->brick_icon;
    if (code == XW_CELL_AGENT) return cat->agent_icon;
    if (code >= XW_CELL_GOAL0) return e->goal_icon[code - XW_CELL_GOAL0];
    return -1;
}

/* XWorldSimulator::get_screen (xworld_simulator.cpp:278-285) for visible_radius == 0, color == true.
 * XItem::get_item_image's warpAffine is the identity for yaw 1.5707963/scale 1/offset 0
 * (byte-identical for all 363 icons) and is not re-evaluated here. */
/* down_sample_image's tail (xworld_simulator.cpp:524-544): optional cv::cvtColor(BGR2GRAY), then HWC -> planar.
 * BGR2GRAY on 8-bit data in the OpenCV 3.2.0 the reference pins (cmake/opencv.cmake:5-6; imgproc/src/color.cpp, RGB2Gray<uchar>:
 * yuv_shift = 14, B2Y = 1868, G2Y = 9617, R2Y = 4899): (B*1868 + G*9617 + R*4899 + (1 << 13)) >> 14.  OpenCV 4.x changed the
 * constants (3735, 19235, 9798, shift 15): tests/test_oracle_fpv.py checks this formula's structure against cv2 with those. */
void xo_gray_or_planes(const xw_config* cfg, const uint8_t* img_out, int oh, int ow, uint8_t* out) {
    if (cfg->gray) {
        for (int i = 0; i < oh * ow; ++i)
            out[i] = (uint8_t)((img_out[i * 3] * 1868 + img_out[i * 3 + 1] * 9617 + img_out[i * 3 + 2] * 4899 + (1 << 13)) >> 14);
        return;
    }
    for (int h = 0; h < oh; ++h)
        for (int w = 0; w < ow; ++w)
            for (int c = 0; c < 3; ++c)
                out[(size_t)c * ow * oh + (size_t)h * ow + w] = img_out[((size_t)h * ow + w) * 3 + c];
}

void xo_render(const xw_config* cfg, const xw_catalog* cat, const xo_env* e, uint8_t* out) {
    if (cfg->visible_radius > 0) { xo_render_fpv(cfg, cat, e, out); return; }
    const int G = XW_ICON_SIZE;
    int ch = e->H * G, cw = e->W * G;
    int oh, ow;
    xo_frame_dims(cfg, &oh, &ow);
    size_t csz = (size_t)ch * cw * 3;
    uint8_t* world = (uint8_t*)malloc(csz * 3 + (size_t)oh * ow * 3);
    uint8_t* planar = world + csz;
    uint8_t* img = planar + csz;
    uint8_t* img_out = img + csz;
    /* XMap::to_image (xmap.cpp:125-146): white canvas, copyTo per item */
    memset(world, 255, csz);
    for (int i = 0; i < e->H; ++i)
        for (int j = 0; j < e->W; ++j) {
            int icon = icon_of_cell(cat, e, e->grid[i * e->W + j]);
            if (icon < 0) continue;
            const uint8_t* src = cat->atlas64 + (size_t)icon * G * G * 3;
            for (int r = 0; r < G; ++r)
                memcpy(world + ((size_t)(i * G + r) * cw + (size_t)j * G) * 3, src + (size_t)r * G * 3, (size_t)G * 3);
        }
    /* get_screen_rgb (xworld_simulator.cpp:287-307): identity resize, HWC -> planar B,G,R */
    for (int i = 0; i < ch; ++i)
        for (int j = 0; j < cw; ++j) {
            const uint8_t* p = world + ((size_t)i * cw + j) * 3;
            planar[(size_t)i * cw + j] = p[0];
            planar[(size_t)cw * ch + (size_t)i * cw + j] = p[1];
            planar[2 * (size_t)cw * ch + (size_t)i * cw + j] = p[2];
        }
    /* down_sample_image (xworld_simulator.cpp:508-545) */
    for (int h = 0; h < ch; ++h)
        for (int w = 0; w < cw; ++w) {
            uint8_t* p = img + ((size_t)h * cw + w) * 3;
            p[0] = planar[(size_t)h * cw + w];
            p[1] = planar[(size_t)cw * ch + (size_t)h * cw + w];
            p[2] = planar[2 * (size_t)cw * ch + (size_t)h * cw + w];
        }
    xo_resize_linear_8uc3(img, ch, cw, img_out, oh, ow);
    xo_gray_or_planes(cfg, img_out, oh, ow, out);
    free(world);
}

/* ------------------------------------------------------------------------------------------ */
/* Batch helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */
int xo_sizeof_env(void) { return (int)sizeof(xo_env); }

int xo_batch_reset(const xw_config* cfg, const xw_catalog* cat, xo_env* envs, int n, int threads) {
    int err = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int i = 0; i < n; ++i) {
        int rc = xo_reset(cfg, cat, &envs[i]);
        if (rc) {
#pragma omp atomic write
            err = rc;
        }
    }
    return err;
}

int xo_batch_step(const xw_config* cfg, const xw_catalog* cat, xo_env* envs, int n, const int32_t* actions,
                  int act_rep, float* reward, int32_t* game_over, uint8_t* frames, int threads) {
    int oh, ow;
    xo_frame_dims(cfg, &oh, &ow);
    size_t fb = (size_t)(cfg->gray ? 1 : 3) * oh * ow;
    int err = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int i = 0; i < n; ++i) {
        int rc = xo_step(cfg, cat, &envs[i], actions[i], act_rep, &reward[i], &game_over[i]);
        if (rc == 0 && cfg->auto_reset && game_over[i] != 0) rc = xo_reset(cfg, cat, &envs[i]);
        if (frames) xo_render(cfg, cat, &envs[i], frames + fb * (size_t)i);
        if (rc) {
#pragma omp atomic write
            err = rc;
        }
    }
    return err;
}

/* ------------------------------------------------------------------------------------------ */
/* SimpleGame (games/simple_game/simple_game_simulator.cpp:31-76)                              */
/* ------------------------------------------------------------------------------------------ */
void xo_sg_reset(xo_simple_game* g, int array_size) {
    g->array_size = array_size;
    g->cur_pos = array_size / 2;
    memset(g->state, 0, sizeof g->state);
    memset(g->rewards, 0, sizeof g->rewards);
    g->state[g->cur_pos] = 1;
    g->rewards[array_size - 1] = 4.0f / 2;
    g->rewards[0] = 4.0f;
}
int xo_sg_game_over(const xo_simple_game* g) { return g->cur_pos <= 0 || g->cur_pos >= g->array_size - 1; }
static float sg_reward(xo_simple_game* g) {
    float reward = -0.1f;
    if (g->cur_pos >= 0 && g->cur_pos < g->array_size && g->rewards[g->cur_pos] != 0.0) {
        reward = g->rewards[g->cur_pos];
        g->rewards[g->cur_pos] = 0.0;
    }
    return reward;
}
float xo_sg_act(xo_simple_game* g, int a) {
    if (xo_sg_game_over(g)) return sg_reward(g);
    g->state[g->cur_pos] = 0;
    if (a == 0) --g->cur_pos; else ++g->cur_pos;
    if (g->cur_pos >= 0 && g->cur_pos < g->array_size) g->state[g->cur_pos] = 1;
    return sg_reward(g);
}

/* ------------------------------------------------------------------------------------------ */
/* SimpleRace (games/simple_race/simple_race_simulator.cpp).  float storage; cos/sin/sqrt/fabs   */
/* evaluate in double and round to float, as the reference's C <math.h> overloads do (App. A.3). */
/* ------------------------------------------------------------------------------------------ */
#define RACE_PI 3.1415926 /* simple_race_simulator.h:39 */

typedef struct { float mid_x, mid_y, start_x, start_y, end_x, end_y, length, width, inner, outer; } race_track;

static void race_track_init(const xw_config* cfg, race_track* t) {
    float cx = 480 / 2, cy = 720 / 2; /* WINDOW_WIDTH/HEIGHT :31-32, :446 */
    memset(t, 0, sizeof *t);
    t->mid_x = cx; t->mid_y = cy;
    if (cfg->track_type == 0) { /* StraightTrack ctor :103-109 */
        t->length = cfg->track_length; t->width = cfg->track_width;
        t->start_x = t->mid_x - 0.0f; t->start_y = t->mid_y - (float)(0.4 * t->length);
        t->end_x = t->mid_x + 0.0f; t->end_y = t->mid_y + (float)(0.6 * t->length);
    } else { /* CircleTrack ctor :50-54 */
        t->inner = cfg->track_radius; t->width = cfg->track_width;
        t->outer = t->inner + t->width;
    }
}
static double race_norm(float x, float y) { return sqrt((double)x * x + (double)y * y); } /* cv::norm(Point2f) -> double */

/* util::get_rand_range_val (simulator_util.cpp:57-64): std::uniform_real_distribution<float>(0, upper) on the thread's
 * std::default_random_engine (minstd_rand0).  libstdc++ (bits/random.tcc generate_canonical<float, 24>): the engine's range
 * 2147483646 has floor(log2) = 30 >= 24 bits, so ONE draw: sum = float(x - min), tmp = float(1.0f * 2147483646.0L) = 2^31,
 * ret = sum / tmp, clipped to nextafterf(1, 0) when it rounds to 1; then ret * (b - a) + a. */
float xo_rand_range_val(uint32_t* st, float upper) {
    const uint32_t x = minstd_next(st);
    const float sum = (float)(x - 1u) * 1.0f;
    const float tmp = (float)(1.0L * 2147483646.0L);
    float ret = sum / tmp;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret * (upper - 0.0f) + 0.0f;
}

void xo_race_reset(const xw_config* cfg, xo_race* r) { /* RaceEngine::reset_game :267-284 */
    race_track t;
    race_track_init(cfg, &t);
    if (cfg->race_random) {
        /* :268-273 the track draw (one track in the pool: the index is 0 either way, the draw is still taken) */
        (void)xo_rand_range_val(&r->minstd, 1.0f);
        if (cfg->track_type == 0) { /* StraightTrack::get_start_pos :192-199 */
            float dy = xo_rand_range_val(&r->minstd, 1.0f) * t.length / 2;
            float dx = (float)((xo_rand_range_val(&r->minstd, 1.0f) - 0.5) * t.width);
            r->pos_x = dx + t.start_x; r->pos_y = dy + t.start_y;
        } else { /* CircleTrack::get_start_pos :78-86 */
            float theta = (float)(xo_rand_range_val(&r->minstd, 1.0f) * 2 * RACE_PI);
            float rad = t.inner + xo_rand_range_val(&r->minstd, 1.0f) * t.width;
            r->pos_x = (float)(rad * cos((double)theta)) + t.mid_x;
            r->pos_y = (float)(rad * sin((double)theta)) + t.mid_y;
        }
        r->angle = (float)(xo_rand_range_val(&r->minstd, 1.0f) * 2 * RACE_PI); /* BaseCar::set_angle :237-243 */
        r->steps = 0;
        return;
    }
    if (cfg->track_type == 0) { r->pos_x = t.start_x; r->pos_y = t.start_y; }
    else { r->pos_x = (t.inner + t.width / 2) + t.mid_x; r->pos_y = 0.0f + t.mid_y; } /* :76-79 */
    r->angle = (float)(RACE_PI / 2);
    r->steps = 0;
}

static int race_out_of_bound(const xw_config* cfg, const race_track* t, float px, float py) {
    if (cfg->track_type == 0) /* :182-186 */
        return (px < t->mid_x - t->width / 2) || (px > t->mid_x + t->width / 2) || (py < t->start_y) || (py > t->end_y);
    float rr = (float)race_norm(px - t->mid_x, py - t->mid_y); /* :72-76 */
    return rr < t->inner || rr > t->outer;
}
static float race_hdisp(const xw_config* cfg, const race_track* t, float px, float py) {
    if (cfg->track_type == 0) return 2 * (px - t->mid_x) / t->width; /* :204-206 */
    return (float)((2 * race_norm(px - t->mid_x, py - t->mid_y) - t->inner - t->outer) / t->width); /* :88-91 */
}
static float race_vdisp(const xw_config* cfg, const race_track* t, float px, float py) {
    (void)px;
    if (cfg->track_type == 0) return 2 * (py - t->mid_y) / t->length; /* :212-214 */
    return 0;
}
static void race_tangent(const xw_config* cfg, const race_track* t, float px, float py, float* tx, float* ty) {
    if (cfg->track_type == 0) { *tx = 0.0f; *ty = 1.0f; return; } /* :220-222 */
    float ax = t->mid_y - py, ay = px - t->mid_x;              /* :97-100 */
    double s = 1 / race_norm(ax, ay);
    *tx = (float)(ax * s); *ty = (float)(ay * s);
}

float xo_race_act(const xw_config* cfg, xo_race* r, int action_index, float state[4], int32_t* game_over) {
    race_track t;
    race_track_init(cfg, &t);
    int a = cfg->race_full_manouver ? action_index : (action_index == 0 ? 4 : 7); /* get_action_set :432-440 */
    const float delta_ang = (float)(RACE_PI / 10), delta_fwd = 1; /* :262-263 */
    r->steps++;
    float d_forward = 0.0f, d_turn = 0.0f;
    switch (a % 3) { case 1: d_forward = delta_fwd; break; case 2: d_forward = -delta_fwd; break; default: break; }
    a /= 3;
    switch (a % 3) { case 1: d_turn = delta_ang; break; case 2: d_turn = -delta_ang; break; default: break; }
    /* BaseCar::move :227-235 */
    r->angle += d_turn;
    if (r->angle > 2 * RACE_PI) r->angle = (float)(r->angle - 2 * RACE_PI);
    else if (r->angle < 0) r->angle = (float)(r->angle + 2 * RACE_PI);
    float cx = (float)cos((double)r->angle), sx = (float)sin((double)r->angle);
    r->pos_x += d_forward * cx;
    r->pos_y += d_forward * sx;
    /* get_reward :386-410 */
    float tx, ty;
    race_tangent(cfg, &t, r->pos_x, r->pos_y, &tx, &ty);
    float vx = (float)cos((double)r->angle), vy = (float)sin((double)r->angle);
    float reward_speed = (vx * tx + vy * ty) * d_forward;
    int finish = (cfg->track_type == 0) && (r->pos_y > t.end_y); /* race_finish :188-190 */
    float reward_finish = finish ? 2.0f : 0.0f;
    float reward_boundary;
    int oob = race_out_of_bound(cfg, &t, r->pos_x, r->pos_y);
    if (cfg->difficulty == 0) reward_boundary = (float)(-fabs((double)race_hdisp(cfg, &t, r->pos_x, r->pos_y)));
    else reward_boundary = (oob && !finish) ? -2.0f : 0.0f;
    float reward = reward_finish + reward_boundary + reward_speed;
    reward = (float)(reward * (double)cfg->reward_scale);
    /* get_screen :412-430 */
    double ca = cos((double)r->angle), sa = sin((double)r->angle);
    float cos_theta = (float)fmax(-1.0, fmin(1.0, tx * ca + ty * sa));
    float sin_theta = (float)sqrt((double)(1 - cos_theta * cos_theta));
    if (ca * ty + sa * tx < 0) sin_theta = -sin_theta;
    state[0] = cos_theta; state[1] = sin_theta;
    state[2] = race_hdisp(cfg, &t, r->pos_x, r->pos_y);
    state[3] = race_vdisp(cfg, &t, r->pos_x, r->pos_y);
    /* SimpleRaceGame::game_over :465-467 */
    int code = 0;
    if (cfg->max_steps > 0 && r->steps >= cfg->max_steps) code |= XW_MAX_STEP;
    if (oob) code |= XW_DEAD;
    *game_over = code;
    return reward;
}

/* GameSimulator::take_actions (simulator.cpp:98-108): num_steps_++ once, the action act_rep times, float rewards summed */
float xo_race_take_actions(const xw_config* cfg, xo_race* r, int action_index, int act_rep, float state[4], int32_t* game_over) {
    const int32_t steps0 = r->steps;
    float reward = 0;
    for (int i = 0; i < act_rep; i++) reward += xo_race_act(cfg, r, action_index, state, game_over);
    r->steps = steps0 + 1;
    *game_over &= ~XW_MAX_STEP;
    if (cfg->max_steps > 0 && r->steps >= cfg->max_steps) *game_over |= XW_MAX_STEP;
    return reward;
}

double xo_race_batch(const xw_config* cfg, xo_race* envs, int n, const int32_t* actions, int steps) {
    double sum = 0;
    float st[4];
    int32_t over;
    for (int s = 0; s < steps; ++s)
        for (int i = 0; i < n; ++i) {
            sum += xo_race_act(cfg, &envs[i], actions[(size_t)s * n + i], st, &over);
            if (over) xo_race_reset(cfg, &envs[i]);
        }
    return sum;
}
