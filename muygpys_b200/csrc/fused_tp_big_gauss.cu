#define MGP_COL_F 3
#define MGP_TP_BIG 1
#include "fused_tp_inst.cuh"
