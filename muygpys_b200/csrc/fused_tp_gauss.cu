#define MGP_COL_F 3
#include "fused_tp_inst.cuh"
