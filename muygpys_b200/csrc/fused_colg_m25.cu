#define MGP_COL_F 2
#define MGP_COL_GRAD 1
#include "fused_col_inst.cuh"
