#define SFB_L 8
#define SFB_DDRX 1
#define SFB_R 0
#define SFB_TN 64
#define SFB_MINB 2
#define SFB_NAME sfb_launch_step_L8_ddrx
#define SFB_APPLY_INC "gen/apply_L8_ddrx.inc"
#include "sfb_step_kernel4.cuh"
