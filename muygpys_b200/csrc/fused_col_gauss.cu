#define MGP_COL_F 3
#include "fused_col_inst.cuh"
