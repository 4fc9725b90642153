#define MGP_COL_F 1
#include "fused_col_inst.cuh"
