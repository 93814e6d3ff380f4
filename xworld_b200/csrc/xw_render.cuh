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
// SELECT from those tables: out = T[icon(owner(dy,dx))].  The few output columns/rows whose taps
// straddle a cell border (2 of 84 at 11x11->84, none at 7x7->84) are re-evaluated exactly from
// the 64-px atlas in a fix-up pass.  Frames are composed in shared memory and leave the SM as one
// TMA bulk store (cp.async.bulk.global.shared::cta) per env; the brick table (the icon most cells
// hold) is staged into shared memory once per CTA by a TMA bulk load.
#pragma once
#include "xw_common.cuh"

#define XW_MAX_OUT 256  // max frame side

struct XwRender {
    int32_t OH, OW, WR, FB;   // frame rows, cols, words per row, bytes per frame (3*OH*OW)
    int32_t H, W;
    int32_t R, rpg;           // row groups per plane, rows per group
    int32_t n_sc, n_sr;       // straddling columns / rows
    int32_t n_icons, brick_icon, agent_icon;
    // LUTs (device): all int16 / uint8 so the whole set is < 4 KB
    const int16_t *xofs, *xa0, *xa1, *yofs, *ya0, *ya1;  // cv::resize tables
    const uint8_t *rowcell, *bandend;                    // [OH] owner cell row; [H] first row of next band
    const uint32_t* colpair;                             // [WR] txA | txB<<8 | prmt_sel<<16
    const int16_t *sc, *sr;                              // straddle column / row indices
    const uint8_t* T;                                    // [n_icons][3][OH][OW]
    const uint8_t* atlas64;                              // [n_icons][64][64][3] BGR
};

// ---- exact cv::resize arithmetic --------------------------------------------------------
// One output pixel of the bilinear resize given its four taps (HResizeLinear + VResizeLinear,
// INTER_RESIZE_COEF_BITS = 11, FixedPtCast shift 22 split as >>4, >>16, +2, >>2).
XW_HD uint8_t xw_resize_px(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    int S0 = p00 * a0 + p01 * a1;
    int S1 = p10 * a0 + p11 * a1;
    return (uint8_t)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);
}

// Phase-atlas entry: canvas tiled with `icon` everywhere.
XW_HD uint8_t xw_phase_px(const XwRender& r, int icon, int c, int dy, int dx) {
    const uint8_t* I = r.atlas64 + (size_t)icon * (64 * 64 * 3);
    int sx0 = r.xofs[dx], sy0 = r.yofs[dy];
    int a0 = r.xa0[dx], a1 = r.xa1[dx], b0 = r.ya0[dy], b1 = r.ya1[dy];
    int sx1 = a1 ? sx0 + 1 : sx0, sy1 = b1 ? sy0 + 1 : sy0;
    int p00 = I[((sy0 & 63) * 64 + (sx0 & 63)) * 3 + c], p01 = I[((sy0 & 63) * 64 + (sx1 & 63)) * 3 + c];
    int p10 = I[((sy1 & 63) * 64 + (sx0 & 63)) * 3 + c], p11 = I[((sy1 & 63) * 64 + (sx1 & 63)) * 3 + c];
    return xw_resize_px(p00, p01, p10, p11, a0, a1, b0, b1);
}

// One tap of the virtual canvas: cell descriptor (icon+1, 0 = empty = white) -> pixel.
XW_HD int xw_canvas_tap(const XwRender& r, const uint32_t* celldesc, int sy, int sx, int c) {
    uint32_t dsc = celldesc[(sy >> 6) * r.W + (sx >> 6)];
    if (dsc == 0) return 255;
    return r.atlas64[(size_t)(dsc - 1) * (64 * 64 * 3) + ((sy & 63) * 64 + (sx & 63)) * 3 + c];
}

// Exact value of output pixel (c,dy,dx) on the real canvas; `same` reports whether all four taps
// lie in cells holding the same descriptor (then the phase atlas already has the right value).
XW_HD uint8_t xw_exact_px(const XwRender& r, const uint32_t* celldesc, int c, int dy, int dx, bool* same) {
    int sx0 = r.xofs[dx], sy0 = r.yofs[dy];
    int a0 = r.xa0[dx], a1 = r.xa1[dx], b0 = r.ya0[dy], b1 = r.ya1[dy];
    int sx1 = a1 ? sx0 + 1 : sx0, sy1 = b1 ? sy0 + 1 : sy0;
    uint32_t d00 = celldesc[(sy0 >> 6) * r.W + (sx0 >> 6)], d01 = celldesc[(sy0 >> 6) * r.W + (sx1 >> 6)];
    uint32_t d10 = celldesc[(sy1 >> 6) * r.W + (sx0 >> 6)], d11 = celldesc[(sy1 >> 6) * r.W + (sx1 >> 6)];
    if (d00 == d01 && d00 == d10 && d00 == d11) { *same = true; return 0; }
    *same = false;
    return xw_resize_px(xw_canvas_tap(r, celldesc, sy0, sx0, c), xw_canvas_tap(r, celldesc, sy0, sx1, c),
                        xw_canvas_tap(r, celldesc, sy1, sx0, c), xw_canvas_tap(r, celldesc, sy1, sx1, c), a0, a1, b0, b1);
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

// ---- compose: one thread owns output word column k of plane c for a group of rows ---------
// celldesc: [H*W] u32 (icon+1 / 0); hot: the brick table in shared memory; fb: the frame being built.
XW_HD void xw_compose_thread(const XwRender& r, int tid, const uint32_t* celldesc, const uint8_t* rowcell,
                             const uint8_t* bandend, const uint32_t* colpair, const uint8_t* hot, uint32_t* fb) {
    const int items = 3 * r.R * r.WR;
    if (tid >= items) return;
    const int k = tid % r.WR, t2 = tid / r.WR, rg = t2 % r.R, c = t2 / r.R;
    const uint32_t pair = colpair[k];
    const int txA = pair & 0xff, txB = (pair >> 8) & 0xff;
    const uint32_t sel = pair >> 16;
    const int r0 = rg * r.rpg, r1 = (r0 + r.rpg < r.OH) ? r0 + r.rpg : r.OH;
    int dy = r0;
    while (dy < r1) {
        const int ty = rowcell[dy];
        int e = bandend[ty];
        if (e > r1) e = r1;
        const uint32_t dA = celldesc[ty * r.W + txA], dB = celldesc[ty * r.W + txB];
        const uint8_t* pA = dA == 0 ? nullptr : ((int)dA - 1 == r.brick_icon ? hot : r.T + (size_t)(dA - 1) * r.FB);
        const uint8_t* pB = dB == 0 ? nullptr : ((int)dB - 1 == r.brick_icon ? hot : r.T + (size_t)(dB - 1) * r.FB);
        int widx = (c * r.OH + dy) * r.WR + k;
        if (pA == nullptr && pB == nullptr) {
            for (; dy < e; ++dy, widx += r.WR) fb[widx] = 0xffffffffu;
        } else if (dA == dB) {
            for (; dy < e; ++dy, widx += r.WR) fb[widx] = *(const uint32_t*)(pA + (size_t)widx * 4);
        } else {
            for (; dy < e; ++dy, widx += r.WR) {
                uint32_t wa = pA ? *(const uint32_t*)(pA + (size_t)widx * 4) : 0xffffffffu;
                uint32_t wb = pB ? *(const uint32_t*)(pB + (size_t)widx * 4) : 0xffffffffu;
                fb[widx] = xw_prmt(wa, wb, sel);
            }
        }
    }
}

// ---- straddle fix-up: item i of the list -> one exact pixel ---------------------------------
XW_HD int xw_fix_count(const XwRender& r) { return 3 * (r.n_sc * r.OH + r.n_sr * r.OW); }
XW_HD void xw_fix_item(const XwRender& r, int i, const uint32_t* celldesc, const int16_t* sc, const int16_t* sr, uint8_t* fb) {
    const int per_plane = r.n_sc * r.OH + r.n_sr * r.OW;
    const int c = i / per_plane;
    int j = i - c * per_plane, dy, dx;
    if (j < r.n_sc * r.OH) { dx = sc[j / r.OH]; dy = j % r.OH; }
    else { j -= r.n_sc * r.OH; dy = sr[j / r.OW]; dx = j % r.OW; }
    bool same;
    uint8_t v = xw_exact_px(r, celldesc, c, dy, dx, &same);
    if (!same) fb[(c * r.OH + dy) * r.OW + dx] = v;
}

// celldesc for one cell: grid code -> icon + 1
XW_HD uint32_t xw_cell_desc(const XwDev& d, int e, int code) {
    if (code == XW_CELL_EMPTY) return 0;
    if (code == XW_CELL_BLOCK) return (uint32_t)d.brick_icon + 1;
    if (code == XW_CELL_AGENT) return (uint32_t)d.agent_icon + 1;
    return (uint32_t)d.goal_icon[(size_t)(code - XW_CELL_GOAL0) * d.n + e] + 1;
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Dynamic shared memory layout (bytes), all sections 16-byte aligned:
//   [0, FB)            brick phase table (TMA bulk load, once)
//   [FB, 2FB) [2FB,3FB) two frame buffers (compose target / TMA bulk-store source)
//   then LUTs, celldesc[256] u32, one mbarrier
struct XwRenderSmem { int hot, fb0, fb1, rowcell, bandend, colpair, sc, sr, celldesc, bar, total; };
XW_HD int xw_align16(int v) { return (v + 15) & ~15; }
XW_HD XwRenderSmem xw_render_smem(const XwRender& r) {
    XwRenderSmem s;
    int o = 0;
    s.hot = o; o += xw_align16(r.FB);
    s.fb0 = o; o += xw_align16(r.FB);
    s.fb1 = o; o += xw_align16(r.FB);
    s.rowcell = o; o += xw_align16(r.OH);
    s.bandend = o; o += xw_align16(r.H);
    s.colpair = o; o += xw_align16(r.WR * 4);
    s.sc = o; o += xw_align16(r.n_sc * 2 + 2);
    s.sr = o; o += xw_align16(r.n_sr * 2 + 2);
    s.celldesc = o; o += XW_MAX_DIM * XW_MAX_DIM * 4;
    s.bar = o; o += 16;
    s.total = o;
    return s;
}

// Persistent CTAs; CTA b renders envs b, b+gridDim.x, ...  dst frame of env e = frames + e*env_stride.
__global__ void __launch_bounds__(1024, 1)
k_render(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    const XwRenderSmem L = xw_render_smem(r);
    uint8_t* hot = smem + L.hot;
    uint8_t* s_rowcell = smem + L.rowcell;
    uint8_t* s_bandend = smem + L.bandend;
    uint32_t* s_colpair = (uint32_t*)(smem + L.colpair);
    int16_t* s_sc = (int16_t*)(smem + L.sc);
    int16_t* s_sr = (int16_t*)(smem + L.sr);
    uint32_t* s_cell = (uint32_t*)(smem + L.celldesc);
    uint64_t* bar = (uint64_t*)(smem + L.bar);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int HW = r.H * r.W;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {  // stage the brick table: one TMA bulk load per CTA
        mbar_expect_tx(bar, (uint32_t)r.FB);
        tma_load_1d(hot, r.T + (size_t)r.brick_icon * r.FB, (uint32_t)r.FB, bar);
    }
    for (int i = tid; i < r.OH; i += nt) s_rowcell[i] = r.rowcell[i];
    for (int i = tid; i < r.H; i += nt) s_bandend[i] = r.bandend[i];
    for (int i = tid; i < r.WR; i += nt) s_colpair[i] = r.colpair[i];
    for (int i = tid; i < r.n_sc; i += nt) s_sc[i] = r.sc[i];
    for (int i = tid; i < r.n_sr; i += nt) s_sr[i] = r.sr[i];

    // register prefetch of the first env's cell descriptors
    int env = blockIdx.x;
    uint32_t next_desc = 0;
    if (env < d.n && tid < HW) next_desc = xw_cell_desc(d, env, d.grid[(size_t)env * d.CS + tid]);
    mbar_wait(bar, 0);

    const int nfix = xw_fix_count(r);
    for (int it = 0; env < d.n; env += gridDim.x, ++it) {
        uint8_t* fb = smem + ((it & 1) ? L.fb1 : L.fb0);
        if (tid < HW) s_cell[tid] = next_desc;
        if (tid == 0) tma_wait_read<1>();  // the store issued two envs ago has drained this buffer
        __syncthreads();
        {  // prefetch the next env's cells while this one is composed
            const int en = env + gridDim.x;
            if (en < d.n && tid < HW) next_desc = xw_cell_desc(d, en, d.grid[(size_t)en * d.CS + tid]);
        }
        xw_compose_thread(r, tid, s_cell, s_rowcell, s_bandend, s_colpair, hot, (uint32_t*)fb);
        if (nfix > 0) {
            __syncthreads();
            for (int i = tid; i < nfix; i += nt) xw_fix_item(r, i, s_cell, s_sc, s_sr, fb);
        }
        fence_async_smem();  // generic-proxy writes -> visible to the async (TMA) proxy
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(frames + (size_t)env * env_stride, fb, (uint32_t)r.FB);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();
}

// General fallback (any frame size): one thread per output byte, straight to global memory.
__global__ void k_render_generic(XwDev d, XwRender r, uint8_t* __restrict__ frames, size_t env_stride) {
    const size_t total = (size_t)d.n * r.FB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / r.FB), rem = (int)(i % r.FB);
        const int c = rem / (r.OH * r.OW), p = rem % (r.OH * r.OW), dy = p / r.OW, dx = p % r.OW;
        const int sx0 = r.xofs[dx], sy0 = r.yofs[dy];
        const int sx1 = r.xa1[dx] ? sx0 + 1 : sx0, sy1 = r.ya1[dy] ? sy0 + 1 : sy0;
        const uint8_t* g = d.grid + (size_t)e * d.CS;
        uint32_t dd[4];
        const int cy[2] = {sy0 >> 6, sy1 >> 6}, cx[2] = {sx0 >> 6, sx1 >> 6};
        for (int q = 0; q < 4; ++q) dd[q] = xw_cell_desc(d, e, g[cy[q >> 1] * r.W + cx[q & 1]]);
        uint8_t v;
        if (dd[0] == dd[1] && dd[0] == dd[2] && dd[0] == dd[3]) {
            v = dd[0] == 0 ? 255 : r.T[(size_t)(dd[0] - 1) * r.FB + rem];
        } else {
            int px[4];
            const int sy[2] = {sy0, sy1}, sx[2] = {sx0, sx1};
            for (int q = 0; q < 4; ++q)
                px[q] = dd[q] == 0 ? 255
                                   : r.atlas64[(size_t)(dd[q] - 1) * (64 * 64 * 3) + ((sy[q >> 1] & 63) * 64 + (sx[q & 1] & 63)) * 3 + c];
            v = xw_resize_px(px[0], px[1], px[2], px[3], r.xa0[dx], r.xa1[dx], r.ya0[dy], r.ya1[dy]);
        }
        frames[(size_t)e * env_stride + rem] = v;
    }
}

// --context > 1 (GameSimulator::shift_context, simulator.cpp:51-60): slots 1..K-1 -> 0..K-2.
__global__ void k_shift_context(uint8_t* frames, int n, int K, int FB) {
    const int words = FB / 16;  // FB % 16 == 0 checked by the host
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        int4* base = (int4*)(frames + (size_t)e * K * FB);
        for (int s = 0; s + 1 < K; ++s) {
            for (int i = threadIdx.x; i < words; i += blockDim.x) base[(size_t)s * words + i] = base[(size_t)(s + 1) * words + i];
            __syncthreads();
        }
    }
}
#endif  // __CUDACC__
