#define MGP_COL_F 1
#include "fused_tp_inst.cuh"
