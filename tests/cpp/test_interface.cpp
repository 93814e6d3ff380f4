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
        } else {
            EXPECT(threw);
        }
    }
    printf("OK\n");
    return 0;
}
