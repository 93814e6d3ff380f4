// C++ host-side test of include/xworld_b200.hpp (the SimulatorInterface-shaped facade over the C ABI).
// Reads like the reference's own tests/test_simple_game_simulator.cpp:21-47 (config 1: SimpleGame, batch 1,
// CPU plumbing); with a GPU (argv[1] == "gpu") it also runs simple_race through the same class.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/xworld_b200.hpp"

#define EXPECT(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    using xworld_b200::SimulatorInterface;
    {
        xw_config cfg = SimulatorInterface::default_config();
        cfg.array_size = 8;
        SimulatorInterface game("simple_game", cfg);
        game.start();
        game.reset_game();
        size_t h, w, c;
        game.get_screen_out_dimensions(h, w, c);
        EXPECT(h == 1 && w == 8 && c == 1 && game.get_num_actions() == 2);
        int pos = 4;
        const float want[3] = {-0.1f, -0.1f, 2.0f};
        for (int i = 0; i < 3; ++i) {
            xworld_b200::State s = game.get_state(0.f);
            for (int j = 0; j < 8; ++j) EXPECT(s.screen[j] == (j == pos ? 1 : 0));
            float r = game.take_action(1, false);
            ++pos;
            EXPECT(fabsf(r - want[i]) < 1e-6f);
        }
        EXPECT(game.game_over() == XW_SUCCESS && game.game_over_string() == "success");
        {   // simulator_interface.h:72-80 for a game without a teacher: empty info, untouched dimensions, no report
            std::string info = "x";
            game.get_extra_info(info);
            double X = -1, Y = -2, Z = -3;
            game.get_world_dimensions(X, Y, Z);
            EXPECT(info.empty() && X == -1 && Y == -2 && Z == -3 && game.last_action() == "1");
            EXPECT(game.teacher_report_task_performance().empty());
        }
        EXPECT(game.get_num_steps() == 3 && game.get_lives() == 0);
        EXPECT(fabsf(game.acc_reward() - 1.8f) < 1e-5f);
        game.reset_game();
        EXPECT(game.game_over_string() == "alive");
        bool threw = false;
        try { game.take_action(5); } catch (const std::runtime_error&) { threw = true; }  // reference: LOG(FATAL)
        EXPECT(threw);
        threw = false;
        try { SimulatorInterface bad("tetris", cfg); } catch (const std::runtime_error&) { threw = true; }
        EXPECT(threw);
    }
    {   // the CUDA games refuse to run without a device: no CPU fallback
        xw_config cfg = SimulatorInterface::default_config();
        bool threw = false;
        try { SimulatorInterface race("simple_race", cfg, nullptr, 4); } catch (const std::runtime_error& e) {
            threw = true;
            if (argc > 1 && !strcmp(argv[1], "gpu")) { printf("FAIL: simple_race on a GPU box: %s\n", e.what()); return 1; }
        }
        if (argc > 1 && !strcmp(argv[1], "gpu")) {
            SimulatorInterface race("simple_race", cfg, nullptr, 4);
            race.reset_game();
            const std::vector<float>& r = race.take_actions(std::vector<int32_t>{0, 1, 0, 1});
            EXPECT(r.size() == 4 && fabsf(r[0] - 0.920154452f) < 1e-6f);  // SURVEY App. A.2, step 0 of the compiled reference
            EXPECT(race.game_over_string(0) == "alive" && race.frame_bytes() == 16);
            // xworld through the same class: a 6-icon catalog (brick, robot, four goal names), navigation2d.json rules
            std::vector<uint8_t> atlas((size_t)6 * 64 * 64 * 3);
            for (size_t i = 0; i < atlas.size(); ++i) atlas[i] = (uint8_t)((i * 2654435761u) >> 24);
            const int32_t name_first[5] = {0, 1, 2, 3, 4}, name_icons[4] = {2, 3, 4, 5};
            const uint8_t colored[6] = {0, 0, 1, 0, 1, 0};
            xw_catalog cat = {6, 0, 1, 4, name_first, name_icons, colored, atlas.data()};
            xw_config xc = SimulatorInterface::default_config();
            xc.seed = 9; xc.simulator_seed = 2; xc.max_steps_factor = 1;
            const int n = 64;
            SimulatorInterface world("xworld", xc, &cat, n);
            world.reset_game();
            size_t h, w, c;
            world.get_screen_out_dimensions(h, w, c);
            double X = 0, Y = 0, Z = -1;
            world.get_world_dimensions(X, Y, Z);
            EXPECT(h == 96 && w == 96 && c == 3 && X == 8 && Y == 8 && Z == 0 && world.get_num_actions() == 4);
            std::string info;
            world.get_extra_info(info, 5);
            EXPECT(info.compare(0, 18, "5|task:XWorld3DNav") == 0 && info.find(",event:,height:8,width:8") != std::string::npos);
            std::vector<uint8_t> mask(n);
            for (int s2 = 0; s2 < 80; ++s2) {  // max_steps_factor 1: every episode ends within 64 steps (time_up at the latest)
                world.take_actions(std::vector<int32_t>((size_t)n, s2 % 4));
                bool any = false;
                for (int i = 0; i < n; ++i) { mask[i] = world.game_over(i) != 0; any |= mask[i] != 0; }
                if (any) {
                    for (int i = 0; i < n; ++i)
                        if (mask[i]) {
                            world.get_extra_info(info, i);
                            EXPECT(info.find("event:correct_goal") != std::string::npos || info.find("event:wrong_goal") != std::string::npos ||
                                   info.find("event:time_up") != std::string::npos);
                        }
                    world.reset_game(mask.data());
                    for (int i = 0; i < n; ++i) EXPECT(world.game_over(i) == 0);
                }
            }
            EXPECT(world.last_action(3) == "3");
            std::vector<std::string> rep = world.teacher_report_task_performance();
            EXPECT(rep.size() > 5 && rep[0] == "=== XWorld3DNavTarget ===" && rep[1].compare(0, 4, "=== ") == 0 && rep[1].find("(S)/") != std::string::npos);
            bool threw2 = false;
            try { world.take_actions(std::vector<int32_t>((size_t)n, 4)); } catch (const std::runtime_error&) { threw2 = true; }
            EXPECT(threw2);  // invalid action: status + flag, not an abort
        } else {
            EXPECT(threw);
        }
    }
    printf("OK\n");
    return 0;
}
