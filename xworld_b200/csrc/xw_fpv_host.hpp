// xw_fpv_host.hpp -- host-side set-up of the first-person view's tables (shared by xw_engine.cu and the test-only host
// build tests/hostsim): cv::resize coefficient tables, the agent's four icons, the cos / sin table of the goal yaws.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "xw_fpv.cuh"

// cv::resize INTER_LINEAR tables as the frame kernels read them (OpenCV imgproc/resize.cpp): columns clamp index and weight
// (`if (sx < 0) fx = 0, sx = 0`, same at the right edge); rows keep the raw weight and the invoker clips the two row
// indices instead (resizeGeneric_Invoker: `clip(sy0 - ksize2 + 1 + k, 0, ssize.height)`) -- oracle/xw_oracle.c
// xo_resize_linear_8uc3 restates the same and is checked against cv2 for up- and downscaling.
static inline void xw_fpv_resize_tables(int src, int dst, bool rows, std::vector<int16_t>& ofs, std::vector<int16_t>& a0,
                                        std::vector<int16_t>& a1) {
    ofs.resize(dst); a0.resize(dst); a1.resize(dst);
    const double scale = 1. / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int sv = (int)floorf(f);
        f -= (float)sv;
        if (!rows) {
            if (sv < 0) { sv = 0; f = 0.f; }
            if (sv >= src - 1) { sv = src - 1; f = 0.f; }
        }
        ofs[d] = (int16_t)sv;
        a0[d] = (int16_t)lrintf((1.f - f) * 2048.f);
        a1[d] = (int16_t)lrintf(f * 2048.f);
    }
}

// The agent's icon for headings right, down, left, up (xitem.cpp:47-60 with the agent's yaw: quarter turns about (32, 32),
// white border; xw_fpv_agent_src)
static inline std::vector<uint8_t> xw_fpv_agent_icons(const uint8_t* icon) {
    std::vector<uint8_t> a4((size_t)4 * 12288, 255);
    for (int f = 0; f < 4; ++f)
        for (int y = 0; y < 64; ++y)
            for (int x = 0; x < 64; ++x) {
                int sy, sx;
                if (xw_fpv_agent_src(f, y, x, &sy, &sx)) memcpy(&a4[(size_t)f * 12288 + (y * 64 + x) * 3], icon + (sy * 64 + sx) * 3, 3);
            }
    return a4;
}

// cos / sin of the goals' rotation angles, by the host's libm (xw_common.cuh XW_YAW_STEPS): getRotationMatrix2D's
// `angle *= CV_PI / 180` of `90 - yaw * 180 / M_PI` (xitem.cpp:53), yaw = 0 + (PI_2 * 4 - 0) * (k / 4096) (xworld_env.py:213-215)
static inline std::vector<double> xw_fpv_yaw_table() {
    std::vector<double> cs((size_t)XW_YAW_STEPS * 2);
    for (int k = 0; k < XW_YAW_STEPS; ++k) {
        const double yaw = 0 + (1.5707963 * 4 - 0) * ((double)k / (double)XW_YAW_STEPS);
        const double angle = (90 - yaw * 180 / M_PI) * (3.1415926535897932384626433832795 / 180);
        cs[2 * k] = cos(angle); cs[2 * k + 1] = sin(angle);
    }
    return cs;
}
