#!/bin/bash
# The `ncu --set full` captures of the end-of-round pass (tools/final_gpu.sh): two consecutive launches of each render kernel
# (all envs, then the auto-reset queue's envs again -- the long one is the frame kernel), one of the reset and step kernels.
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r02z}
tools/ncu_cap.sh ${T}_c3_render "k_render_sp" c3 8 2
tools/ncu_cap.sh ${T}_c2_render "k_render_sp" c2 8 2
tools/ncu_cap.sh ${T}_c4_render "k_render_sp" c4 8 2
tools/ncu_cap.sh ${T}_fpv_frame "k_render_fpv_cells" fpv 8 2
tools/ncu_cap.sh ${T}_fpv_goal "k_fpv_goal_cells" fpv 8 2
tools/ncu_cap.sh ${T}_c2_reset "k_reset_list" c2 5 1
tools/ncu_cap.sh ${T}_c3_step "k_step" c3 5 1
