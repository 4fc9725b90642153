#define MGP_COL_F 2
#include "fused_tp_inst.cuh"
