#define MGP_COL_F 2
#include "fused_col_inst.cuh"
