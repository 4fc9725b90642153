#define MGP_COL_F 0
#include "fused_col_inst.cuh"
