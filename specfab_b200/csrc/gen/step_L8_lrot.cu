#define SFB_L 8
#define SFB_DDRX 0
#define SFB_R 2
#define SFB_TN 32
#define SFB_MINB 2
#define SFB_NAME sfb_launch_step_L8_lrot
#define SFB_APPLY_INC "gen/apply_L8_lrot.inc"
#include "sfb_step_kernel.cuh"
