#define SFB_L 8
#define SFB_DDRX 0
#define SFB_R 1
#define SFB_TN 16
#define SFB_MINB 6
#define SFB_NAME sfb_launch_step_L8_lrot
#define SFB_APPLY_INC "gen/apply_L8_lrot.inc"
#define SFB_NP 1
#include "sfb_step_kernel.cuh"
