// xw_teacher_names.hpp -- the task class names of the two task groups and the XWorldRec group's task sampling, for
// XWorldSimulator::get_extra_info's "task:" field (xworld_simulator.cpp:495-504).  Host only.
#pragma once
#include <math.h>
#include <stdint.h>

// conf-json order of the task groups (confs/navigation2d.json, confs/walls.json)
static const char* const kT3Names[5] = {"XWorld3DNavTarget", "XWorld3DNavTargetNear", "XWorld3DNavTargetBetween",
                                        "XWorld3DNavTargetDirection", "XWorld3DNavTargetAvoid"};
static const char* const kT2Names[4] = {"XWorldNavTarget", "XWorldNavNear", "XWorldNavColorTarget", "XWorldNavBetween"};
static const char* const kRecNames[12] = {
    "XWorldRecDirectionToObject", "XWorldRecObjectToDirection", "XWorldRecColorToObject", "XWorldRecObjectToColor",
    "XWorldRecDirectionToColor", "XWorldRecColorToDirection", "XWorldRecDirectionAndObjectToObject",
    "XWorldRecDirectionAndObjectToColor", "XWorldRecColorAndObject", "XWorldRecBetweenToObject", "XWorldRecBetweenToDirection",
    "XWorldRecBetweenToColor"};
// TaskGroup::run_stage, schedule "weighted" (teaching_task.cpp:204-222): util::simple_importance_sampling over the accumulated
// task weights of walls.json's XWorldRec group with get_rand_range_val = std::uniform_real_distribution<float>(0, 14) on the
// thread's minstd_rand0 (simulator_util.cpp:56-86; libstdc++ generate_canonical<float, 24>: one engine call, (x - 1) / 2147483646
// in float, clamped below 1).  `x` = the engine's output of that call = its state afterwards.
static inline int rec_task_of_draw(uint32_t x) {
    static const double acc[12] = {1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14};
    float r = (float)(x - 1u) / (float)2147483646.0L;
    if (r >= 1.f) r = nextafterf(1.f, 0.f);
    const float w = (14.f - 0.f) * r + 0.f;
    for (int i = 0; i < 12; ++i) if (w <= acc[i]) return i;
    return 11;
}

