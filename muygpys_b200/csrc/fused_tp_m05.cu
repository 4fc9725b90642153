#define MGP_COL_F 0
#include "fused_tp_inst.cuh"
