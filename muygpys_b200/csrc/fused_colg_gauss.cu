#define MGP_COL_F 3
#define MGP_COL_GRAD 1
#include "fused_col_inst.cuh"
